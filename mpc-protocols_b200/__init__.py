"""hbmpc_b200 -- B200-native finite-field hot path of HoneyBadgerMPC secret sharing (ark_bls12_381::Fr).

This package is a thin ctypes binding over the C-ABI shared library ``libhbmpc_b200.so`` (``include/hbmpc_b200.h``).
All arithmetic runs in hand-written CUDA kernels for sm_100a; there is NO CPU fallback: importing the package
without the built library, or creating a ``Context`` without a CUDA device, raises.

Field elements are the canonical value as 4 x uint64 little-endian limbs (the reference's ``U256``,
/root/reference/mpc/src/ffi/c_bindings/mod.rs:17-49): numpy ``uint64[..., 4]`` arrays on the host, or torch CUDA
tensors of dtype int64 (same bits) with a trailing dimension of 4 on the device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HBMPC_LIB") or os.path.join(_HERE, "libhbmpc_b200.so")  # HBMPC_LIB: tuning builds only

# ShareErrorCode (ffi/c_bindings/share/mod.rs:18-37) + library codes
SUCCESS, INSUFFICIENT_SHARES, DEGREE_MISMATCH, ID_MISMATCH, INVALID_INPUT = 0, 1, 2, 3, 4
TYPE_MISMATCH, NO_SUITABLE_DOMAIN, POLYNOMIAL_OPERATION_ERROR, DECODING_ERROR = 5, 6, 7, 8
NO_DEVICE, CUDA_ERROR = 100, 101
ERROR_NAMES = {0: "ShareSuccess", 1: "InsufficientShares", 2: "DegreeMismatch", 3: "IdMismatch", 4: "InvalidInput",
               5: "TypeMismatch", 6: "NoSuitableDomain", 7: "PolynomialOperationError", 8: "DecodingError",
               100: "NoDevice", 101: "CudaError"}

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

EXPORTS = [
    "hbmpc_ctx_create", "hbmpc_ctx_destroy", "hbmpc_ctx_set_stream", "hbmpc_ctx_set_async", "hbmpc_ctx_synchronize",
    "hbmpc_ctx_launch_count", "hbmpc_last_error", "hbmpc_compute_shares_batch", "hbmpc_apply_vandermonde_batch",
    "hbmpc_apply_matrix_batch", "hbmpc_batch_recover", "hbmpc_batch_recover_secrets", "hbmpc_robust_interpolate_batch",
    "hbmpc_nonrobust_recover_batch", "hbmpc_elementwise", "hbmpc_share_algebra_fused", "hbmpc_unpack_share_records", "hbmpc_pack_share_records",
    "hbmpc_measure_imad_peak", "hbmpc_measure_wide_chains", "hbmpc_measure_mont_mul",
    "hbmpc_sample_fr_batch", "hbmpc_sample_polynomials", "hbmpc_share_secrets_batch",
    "hbmpc_batch_recover_msgs", "hbmpc_batch_recover_secrets_msgs", "hbmpc_apply_vandermonde_msgs",
    "hbmpc_gl_compute_shares_batch", "hbmpc_gl_apply_vandermonde_batch", "hbmpc_gl_batch_recover", "hbmpc_gl_nonrobust_recover_batch",
    "hbmpc_gl_elementwise", "hbmpc_ctx_device",
    "hbmpc_group_create", "hbmpc_group_destroy", "hbmpc_group_size", "hbmpc_group_ctx", "hbmpc_group_shard_range",
    "hbmpc_group_compute_shares_batch", "hbmpc_group_apply_vandermonde_batch", "hbmpc_group_batch_recover",
    "hbmpc_group_batch_recover_secrets", "hbmpc_group_robust_interpolate_batch",
]


class HbmpcError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{ERROR_NAMES.get(code, code)} ({code}) {detail}".strip())


_lib = None


def load_library():
    """Load libhbmpc_b200.so.  Raises if it has not been built (``python mpc-protocols_b200/build.py``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python mpc-protocols_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
    lib.hbmpc_ctx_create.argtypes = [ci, C.POINTER(vp)]
    lib.hbmpc_ctx_destroy.argtypes = [vp]
    lib.hbmpc_ctx_destroy.restype = None
    lib.hbmpc_ctx_set_stream.argtypes = [vp, vp]
    lib.hbmpc_ctx_set_async.argtypes = [vp, ci]
    lib.hbmpc_ctx_synchronize.argtypes = [vp]
    lib.hbmpc_ctx_launch_count.argtypes = [vp]
    lib.hbmpc_ctx_launch_count.restype = C.c_uint64
    lib.hbmpc_last_error.argtypes = [vp]
    lib.hbmpc_last_error.restype = C.c_char_p
    lib.hbmpc_compute_shares_batch.argtypes = [vp, sz, sz, sz, vp, vp]
    lib.hbmpc_apply_vandermonde_batch.argtypes = [vp, sz, sz, sz, vp, vp, ci]
    lib.hbmpc_apply_matrix_batch.argtypes = [vp, sz, sz, vp, sz, vp, vp, ci]
    lib.hbmpc_batch_recover.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp]
    lib.hbmpc_batch_recover_secrets.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp]
    lib.hbmpc_robust_interpolate_batch.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp, vp]
    lib.hbmpc_nonrobust_recover_batch.argtypes = [vp, sz, sz, sz, vp, sz, vp, ci, vp, vp, vp]
    lib.hbmpc_elementwise.argtypes = [vp, ci, sz, vp, vp, vp]
    lib.hbmpc_share_algebra_fused.argtypes = [vp, ci, sz, C.POINTER(vp), C.POINTER(vp)]
    lib.hbmpc_unpack_share_records.argtypes = [vp, sz, vp, vp, vp, vp]
    lib.hbmpc_pack_share_records.argtypes = [vp, sz, vp, sz, sz, vp]
    lib.hbmpc_measure_imad_peak.argtypes = [vp, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.hbmpc_measure_wide_chains.argtypes = [vp, ci, ci, C.POINTER(C.c_double)]
    lib.hbmpc_measure_mont_mul.argtypes = [vp, ci, ci, C.POINTER(C.c_double)]
    lib.hbmpc_sample_fr_batch.argtypes = [vp, vp, sz, vp]
    lib.hbmpc_sample_polynomials.argtypes = [vp, vp, sz, sz, vp, vp]
    lib.hbmpc_share_secrets_batch.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
    lib.hbmpc_batch_recover_msgs.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp]
    lib.hbmpc_batch_recover_secrets_msgs.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp]
    lib.hbmpc_apply_vandermonde_msgs.argtypes = [vp, sz, sz, sz, vp, vp]
    lib.hbmpc_gl_compute_shares_batch.argtypes = [vp, sz, sz, sz, vp, vp]
    lib.hbmpc_gl_apply_vandermonde_batch.argtypes = [vp, sz, sz, sz, vp, vp, ci]
    lib.hbmpc_gl_batch_recover.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp]
    lib.hbmpc_gl_nonrobust_recover_batch.argtypes = [vp, sz, sz, sz, vp, sz, vp, ci, vp, vp, vp]
    lib.hbmpc_gl_elementwise.argtypes = [vp, ci, sz, vp, vp, vp]
    lib.hbmpc_ctx_device.argtypes = [vp]
    lib.hbmpc_group_create.argtypes = [C.POINTER(ci), sz, C.POINTER(vp)]
    lib.hbmpc_group_destroy.argtypes = [vp]
    lib.hbmpc_group_destroy.restype = None
    lib.hbmpc_group_size.argtypes = [vp]
    lib.hbmpc_group_size.restype = sz
    lib.hbmpc_group_ctx.argtypes = [vp, sz]
    lib.hbmpc_group_ctx.restype = vp
    lib.hbmpc_group_shard_range.argtypes = [vp, sz, sz, C.POINTER(sz), C.POINTER(sz)]
    lib.hbmpc_group_shard_range.restype = None
    lib.hbmpc_group_compute_shares_batch.argtypes = [vp, sz, sz, sz, vp, vp]
    lib.hbmpc_group_apply_vandermonde_batch.argtypes = [vp, sz, sz, sz, vp, vp, ci]
    lib.hbmpc_group_batch_recover.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp]
    lib.hbmpc_group_batch_recover_secrets.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp]
    lib.hbmpc_group_robust_interpolate_batch.argtypes = [vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Buf:
    """Uniform view of a numpy host array or a torch CUDA tensor as (pointer, shape)."""

    def __init__(self, x):
        self.obj = x
        if _is_torch(x):
            if not x.is_contiguous():
                raise ValueError("torch tensors must be contiguous")
            self.ptr = x.data_ptr()
            self.shape = tuple(x.shape)
            self.torch = True
        else:
            x = np.ascontiguousarray(x, dtype=np.uint64)
            self.obj = x
            self.ptr = x.ctypes.data
            self.shape = x.shape
            self.torch = False

    def like(self, shape, dtype=None):
        if self.torch:
            import torch

            dt = {None: torch.int64, "i32": torch.int32, "u64": torch.int64}[dtype]
            return torch.empty(shape, dtype=dt, device=self.obj.device)
        dt = {None: np.uint64, "i32": np.int32, "u64": np.uint64}[dtype]
        return np.zeros(shape, dtype=dt)


def _ptr(x):
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def _check_out(x, shape, itemsize=8, name="out"):
    """A caller-supplied output must be contiguous and hold exactly `shape` elements of `itemsize` bytes: the library writes through
    the raw pointer (a wrong shape or a strided view would be written past its end)."""
    want = int(np.prod(shape)) * itemsize
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError(f"{name}: torch tensor must be contiguous")
        have = x.numel() * x.element_size()
    else:
        if not isinstance(x, np.ndarray) or not x.flags["C_CONTIGUOUS"] or not x.flags["WRITEABLE"]:
            raise ValueError(f"{name}: numpy array must be C-contiguous and writeable")
        have = x.nbytes
    if have != want:
        raise ValueError(f"{name}: {have} bytes supplied, {want} needed for shape {tuple(shape)}")
    return x


class Context:
    """One context per GPU (``hbmpc_ctx``): stream, cached constant tables, scratch.

    Torch CUDA tensors are processed on the CONTEXT's stream (created non-blocking: it is not ordered against torch's current
    stream).  Either call ``set_stream(torch.cuda.current_stream().cuda_stream)`` once -- what bench.py and tools/ do -- or
    ``use_torch_stream()`` before calls whose inputs were produced on torch's current stream."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.hbmpc_ctx_create(device, C.byref(h))
        if rc != 0:
            raise HbmpcError(rc, "hbmpc_ctx_create: no usable CUDA device (this library never computes on the CPU)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.hbmpc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def set_stream(self, cuda_stream: int):
        self.lib.hbmpc_ctx_set_stream(self.h, C.c_void_p(cuda_stream))

    def use_torch_stream(self):
        """Move the context onto torch's current stream (so tensors produced by torch ops are ordered before the library's kernels)."""
        import torch

        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def set_async(self, flag: bool):
        self.lib.hbmpc_ctx_set_async(self.h, int(flag))

    def synchronize(self) -> int:
        return self.lib.hbmpc_ctx_synchronize(self.h)

    @property
    def launch_count(self) -> int:
        return int(self.lib.hbmpc_ctx_launch_count(self.h))

    def last_error(self) -> str:
        return (self.lib.hbmpc_last_error(self.h) or b"").decode()

    def _check(self, rc: int, ok=(0,)):
        if rc not in ok:
            raise HbmpcError(rc, self.last_error() if rc == CUDA_ERROR else "")
        return rc

    # -- K1
    def compute_shares_batch(self, coeffs, n: int, out=None):
        """coeffs[B][d+1][4] -> shares[B][n][4]   (RobustShare::compute_shares, robust_interpolate.rs:52-82)"""
        c = _Buf(coeffs)
        B, m = c.shape[0], c.shape[1]
        out = c.like((B, n, 4)) if out is None else _check_out(out, (B, n, 4))
        self._check(self.lib.hbmpc_compute_shares_batch(self.h, n, m - 1, B, c.ptr, _ptr(out)))
        return out

    # -- K2
    def apply_vandermonde_batch(self, inp, n: int, recipient_major: bool = False, out=None):
        """in[B][cols][4] -> out[B][n][4] (or [n][B][4])   (apply_vandermonde, common/share/mod.rs:50-76)"""
        x = _Buf(inp)
        B, cols = x.shape[0], x.shape[1]
        out = x.like((n, B, 4) if recipient_major else (B, n, 4)) if out is None else _check_out(out, (n, B, 4))
        self._check(self.lib.hbmpc_apply_vandermonde_batch(self.h, n, cols, B, x.ptr, _ptr(out), int(recipient_major)))
        return out

    def apply_matrix_batch(self, matrix, inp, recipient_major: bool = False, out=None):
        M = np.ascontiguousarray(matrix, dtype=np.uint64)
        rows, cols = M.shape[0], M.shape[1]
        x = _Buf(inp)
        B = x.shape[0]
        out = x.like((rows, B, 4) if recipient_major else (B, rows, 4)) if out is None else _check_out(out, (rows, B, 4))
        self._check(self.lib.hbmpc_apply_matrix_batch(self.h, rows, cols, M.ctypes.data, B, x.ptr, _ptr(out), int(recipient_major)))
        return out

    # -- K3 / K4
    def batch_recover(self, sender_ids, evals, n: int, d: int, t: int, want_flags: bool = False, out=None):
        """evals[S][B][4] sender-major -> (rc, coeffs[B][d+1][4], path[B], flags[B][ceil(S/64)] or None)
        (batch_recover_secret, robust_interpolate.rs:284-443).  rc is 0 or DecodingError (some chunk undecodable)."""
        e = _Buf(evals)
        S, B = e.shape[0], e.shape[1]
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        if len(ids) != S:
            raise ValueError(f"{len(ids)} sender ids for {S} sender vectors")
        if out is None:
            coeffs, path = e.like((B, d + 1, 4)), e.like((B,), "i32")
            flags = e.like((B, (S + 63) // 64), "u64") if want_flags else None
        else:
            coeffs, path, flags = out
            _check_out(coeffs, (B, d + 1, 4), name="coeffs")
            _check_out(path, (B,), 4, name="path")
            if flags is not None:
                _check_out(flags, (B, (S + 63) // 64), name="flags")
        rc = self.lib.hbmpc_batch_recover(self.h, n, d, t, S, ids.ctypes.data, B, e.ptr, _ptr(coeffs), _ptr(path),
                                          _ptr(flags) if flags is not None else None)
        self._check(rc, ok=(0, DECODING_ERROR))
        return rc, coeffs, path, flags

    def batch_recover_secrets(self, sender_ids, evals, n: int, d: int, t: int, out=None):
        e = _Buf(evals)
        S, B = e.shape[0], e.shape[1]
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        if len(ids) != S:
            raise ValueError(f"{len(ids)} sender ids for {S} sender vectors")
        if out is None:
            secrets, path = e.like((B, 4)), e.like((B,), "i32")
        else:
            secrets, path = out
            _check_out(secrets, (B, 4), name="secrets")
            _check_out(path, (B,), 4, name="path")
        rc = self.lib.hbmpc_batch_recover_secrets(self.h, n, d, t, S, ids.ctypes.data, B, e.ptr, _ptr(secrets), _ptr(path))
        self._check(rc, ok=(0, DECODING_ERROR))
        return rc, secrets, path

    def robust_interpolate_batch(self, ids, shares, n: int, d: int, t: int, want_flags: bool = False, out=None):
        """shares[B][S][4] codeword-major -> (rc, coeffs[B][d+1][4], secrets[B][4], path[B], flags)
        (RobustShare::recover_secret, robust_interpolate.rs:94-157, batched)."""
        s = _Buf(shares)
        B, S = s.shape[0], s.shape[1]
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        if len(idv) != S:
            raise ValueError(f"{len(idv)} ids for codewords of {S} shares")
        if out is None:
            coeffs, secrets, path = s.like((B, d + 1, 4)), s.like((B, 4)), s.like((B,), "i32")
            flags = s.like((B, (S + 63) // 64), "u64") if want_flags else None
        else:
            coeffs, secrets, path, flags = out
            _check_out(coeffs, (B, d + 1, 4), name="coeffs")
            _check_out(secrets, (B, 4), name="secrets")
            _check_out(path, (B,), 4, name="path")
            if flags is not None:
                _check_out(flags, (B, (S + 63) // 64), name="flags")
        rc = self.lib.hbmpc_robust_interpolate_batch(self.h, n, d, t, S, idv.ctypes.data, B, s.ptr, _ptr(coeffs), _ptr(secrets),
                                                     _ptr(path), _ptr(flags) if flags is not None else None)
        self._check(rc, ok=(0, DECODING_ERROR))
        return rc, coeffs, secrets, path, flags

    # -- N1: one host array per sender / recipient (message payloads)
    @staticmethod
    def _ptr_array(arrays):
        arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        return arr

    def batch_recover_msgs(self, sender_ids, sender_evals, n: int, d: int, t: int, want_flags: bool = False, secrets_only: bool = False):
        """sender_evals: list of S host arrays (uint64 views of B x 32-byte values, e.g. payload[8:]) in arrival order"""
        S, B = len(sender_evals), sender_evals[0].size // 4
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        pa = self._ptr_array(sender_evals)
        path = np.zeros(B, dtype=np.int32)
        if secrets_only:
            secrets = np.zeros((B, 4), dtype=np.uint64)
            rc = self.lib.hbmpc_batch_recover_secrets_msgs(self.h, n, d, t, S, ids.ctypes.data, B, pa, secrets.ctypes.data, path.ctypes.data)
            self._check(rc, ok=(0, DECODING_ERROR))
            return rc, secrets, path
        coeffs = np.zeros((B, d + 1, 4), dtype=np.uint64)
        flags = np.zeros((B, (S + 63) // 64), dtype=np.uint64) if want_flags else None
        rc = self.lib.hbmpc_batch_recover_msgs(self.h, n, d, t, S, ids.ctypes.data, B, pa, coeffs.ctypes.data, path.ctypes.data,
                                               flags.ctypes.data if flags is not None else None)
        self._check(rc, ok=(0, DECODING_ERROR))
        return rc, coeffs, path, flags

    def apply_vandermonde_msgs(self, inp, n: int, recipient_out):
        """inp[B][cols][4] -> recipient_out[j] (list of n host arrays of B x 4 uint64) = the vector for recipient j"""
        x = _Buf(inp)
        B, cols = x.shape[0], x.shape[1]
        pa = self._ptr_array(recipient_out)
        self._check(self.lib.hbmpc_apply_vandermonde_msgs(self.h, n, cols, B, x.ptr, pa))
        return recipient_out

    # -- a10
    def nonrobust_recover_batch(self, ids, shares, n: int, deg: int, sender_major: bool = False, out=None):
        """shares[B][S] (or [S][B] when sender_major) -> (coeffs[B][deg+1][4], secrets[B][4], status[B])
        (NonRobustShare::recover_secret, common/share/shamir.rs:199-239, batched; status = degree or -DegreeMismatch)."""
        s = _Buf(shares)
        S, B = (s.shape[0], s.shape[1]) if sender_major else (s.shape[1], s.shape[0])
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        if len(idv) != S:
            raise ValueError(f"{len(idv)} ids for codewords of {S} shares")
        if out is None:
            coeffs, secrets, status = s.like((B, deg + 1, 4)), s.like((B, 4)), s.like((B,), "i32")
        else:
            coeffs, secrets, status = out
            _check_out(coeffs, (B, deg + 1, 4), name="coeffs")
            _check_out(secrets, (B, 4), name="secrets")
            _check_out(status, (B,), 4, name="status")
        self._check(self.lib.hbmpc_nonrobust_recover_batch(self.h, n, deg, len(idv), idv.ctypes.data, B, s.ptr, int(sender_major),
                                                            _ptr(coeffs), _ptr(secrets), _ptr(status)))
        return coeffs, secrets, status

    # -- K5
    def elementwise(self, op: int, a, b, out=None):
        x, y = _Buf(a), _Buf(b)
        if y.shape != x.shape:
            raise ValueError(f"operands differ in shape: {x.shape} and {y.shape}")
        count = int(np.prod(x.shape[:-1]))
        out = x.like(x.shape) if out is None else _check_out(out, x.shape)
        self._check(self.lib.hbmpc_elementwise(self.h, op, count, x.ptr, y.ptr, _ptr(out)))
        return out

    K5_TRIPLE_MASK, K5_BEAVER_MASK, K5_BEAVER_FINALIZE = 0, 1, 2

    def share_algebra_fused(self, op: int, inputs, out=None):
        """One pass over HBM for the share algebra of a protocol step (include/hbmpc_b200.h: hbmpc_share_algebra_fused):
        op 0: (a, b, r_2t) -> a*b - r_2t; op 1: (a, x, b, y) -> (a - x, b - y); op 2: (c, x, y, a-x, b-y) -> the Beaver product share.
        Returns one array (ops 0, 2) or a pair (op 1)."""
        nin, nout = {0: (3, 1), 1: (4, 2), 2: (5, 1)}[op]
        bufs = [_Buf(v) for v in inputs]
        if len(bufs) != nin:
            raise ValueError(f"op {op} takes {nin} input arrays, {len(bufs)} given")
        shape = bufs[0].shape
        for k, b in enumerate(bufs):
            if b.shape != shape or b.torch != bufs[0].torch:
                raise ValueError(f"input {k}: shape {b.shape} differs from input 0 {shape} (or host / device arrays are mixed)")
        count = int(np.prod(shape[:-1]))
        if out is None:
            outs = [bufs[0].like(shape) for _ in range(nout)]
        else:
            outs = list(out) if nout == 2 else [out]
            if len(outs) != nout:
                raise ValueError(f"op {op} writes {nout} arrays")
            for k, o in enumerate(outs):
                _check_out(o, shape, name=f"out[{k}]")
        pin = (C.c_void_p * nin)(*[b.ptr for b in bufs])
        pout = (C.c_void_p * nout)(*[_ptr(o) for o in outs])
        self._check(self.lib.hbmpc_share_algebra_fused(self.h, op, count, pin, pout))
        return tuple(outs) if nout == 2 else outs[0]

    # -- N1: 48-byte ark-serialize share records
    def unpack_share_records(self, records: np.ndarray, count: int):
        """records: uint8 buffer of count*48 bytes (host) -> (values uint64[count][4], ids uint64[count], degrees uint64[count])"""
        rec = np.ascontiguousarray(records, dtype=np.uint8)
        values = np.zeros((count, 4), dtype=np.uint64)
        ids, degs = np.zeros(count, dtype=np.uint64), np.zeros(count, dtype=np.uint64)
        self._check(self.lib.hbmpc_unpack_share_records(self.h, count, rec.ctypes.data, values.ctypes.data, ids.ctypes.data, degs.ctypes.data))
        return values, ids, degs

    def pack_share_records(self, values, per_id: int, degree: int) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(v.shape[0] * 48, dtype=np.uint8)
        self._check(self.lib.hbmpc_pack_share_records(self.h, v.shape[0], v.ctypes.data, per_id, degree, out.ctypes.data))
        return out

    def measure_imad_peak(self, variant: int = 0):
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.hbmpc_measure_imad_peak(self.h, variant, C.byref(g), C.byref(ms)))
        return g.value, ms.value


    # -- N4: device-side sampling (StdRng = ChaCha12 + ark-ff Fp::rand)
    def sample_fr_batch(self, seed: bytes, count: int, out=None):
        assert len(seed) == 32
        sd = np.frombuffer(seed, dtype=np.uint8).copy()
        out = np.zeros((count, 4), dtype=np.uint64) if out is None else out
        self._check(self.lib.hbmpc_sample_fr_batch(self.h, sd.ctypes.data, count, _ptr(out)))
        return out

    def sample_polynomials(self, seed: bytes, B: int, d: int, secrets=None, out=None):
        assert len(seed) == 32
        sd = np.frombuffer(seed, dtype=np.uint8).copy()
        out = np.zeros((B, d + 1, 4), dtype=np.uint64) if out is None else out
        sec = None
        if secrets is not None:
            sec = secrets if _is_torch(secrets) else np.ascontiguousarray(secrets, dtype=np.uint64)
        self._check(self.lib.hbmpc_sample_polynomials(self.h, sd.ctypes.data, B, d, _ptr(sec) if sec is not None else None, _ptr(out)))
        return out

    def share_secrets_batch(self, seed: bytes, secrets, n: int, d: int, out=None, coeffs_out=None):
        """secrets[B][4] -> shares[B][n][4], polynomials drawn on the device from StdRng::from_seed(seed)
        (RobustShare::compute_shares(secret, n, degree, None, rng), robust_interpolate.rs:52-82, for B secrets on one generator)"""
        assert len(seed) == 32
        sd = np.frombuffer(seed, dtype=np.uint8).copy()
        s = _Buf(secrets)
        B = s.shape[0]
        out = s.like((B, n, 4)) if out is None else out
        self._check(self.lib.hbmpc_share_secrets_batch(self.h, sd.ctypes.data, n, d, B, s.ptr, _ptr(out), _ptr(coeffs_out) if coeffs_out is not None else None))
        return out

    # -- N4 tail: Goldilocks (Fp64, p = 2^64 - 2^32 + 1); elements are single canonical uint64 values, host numpy arrays
    def gl_compute_shares_batch(self, coeffs, n: int):
        c = np.ascontiguousarray(coeffs, dtype=np.uint64)
        B, m = c.shape
        out = np.zeros((B, n), dtype=np.uint64)
        self._check(self.lib.hbmpc_gl_compute_shares_batch(self.h, n, m - 1, B, c.ctypes.data, out.ctypes.data))
        return out

    def gl_apply_vandermonde_batch(self, inp, n: int, recipient_major: bool = False):
        x = np.ascontiguousarray(inp, dtype=np.uint64)
        B, cols = x.shape
        out = np.zeros((n, B) if recipient_major else (B, n), dtype=np.uint64)
        self._check(self.lib.hbmpc_gl_apply_vandermonde_batch(self.h, n, cols, B, x.ctypes.data, out.ctypes.data, int(recipient_major)))
        return out

    def gl_batch_recover(self, sender_ids, evals, n: int, d: int, t: int):
        e = np.ascontiguousarray(evals, dtype=np.uint64)
        S, B = e.shape
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        coeffs, secrets, path = np.zeros((B, d + 1), dtype=np.uint64), np.zeros(B, dtype=np.uint64), np.zeros(B, dtype=np.int32)
        rc = self.lib.hbmpc_gl_batch_recover(self.h, n, d, t, S, ids.ctypes.data, B, e.ctypes.data, coeffs.ctypes.data, secrets.ctypes.data, path.ctypes.data)
        self._check(rc, ok=(0, DECODING_ERROR))
        return rc, coeffs, secrets, path

    def gl_nonrobust_recover_batch(self, ids, shares, n: int, deg: int, sender_major: bool = False):
        s = np.ascontiguousarray(shares, dtype=np.uint64)
        S, B = (s.shape[0], s.shape[1]) if sender_major else (s.shape[1], s.shape[0])
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        coeffs, secrets, status = np.zeros((B, deg + 1), dtype=np.uint64), np.zeros(B, dtype=np.uint64), np.zeros(B, dtype=np.int32)
        self._check(self.lib.hbmpc_gl_nonrobust_recover_batch(self.h, n, deg, len(idv), idv.ctypes.data, B, s.ctypes.data, int(sender_major),
                                                               coeffs.ctypes.data, secrets.ctypes.data, status.ctypes.data))
        return coeffs, secrets, status

    def gl_elementwise(self, op: int, a, b):
        x, y = np.ascontiguousarray(a, dtype=np.uint64), np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros_like(x)
        self._check(self.lib.hbmpc_gl_elementwise(self.h, op, x.size, x.ctypes.data, y.ctypes.data, out.ctypes.data))
        return out

    def measure_mont_mul(self, ilp: int, warps_per_smsp: int) -> float:
        g = C.c_double()
        self._check(self.lib.hbmpc_measure_mont_mul(self.h, ilp, warps_per_smsp, C.byref(g)))
        return g.value

    def measure_wide_chains(self, chains: int, warps_per_smsp: int) -> float:
        g = C.c_double()
        self._check(self.lib.hbmpc_measure_wide_chains(self.h, chains, warps_per_smsp, C.byref(g)))
        return g.value


class Group:
    """Single-process multi-GPU group (``hbmpc_group``): one context per device, every batch split into contiguous ranges over the
    devices on internal host threads, no collective.  Calls take host (numpy) buffers."""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.hbmpc_group_create(devs, len(devices), C.byref(h))
        if rc != 0:
            raise HbmpcError(rc, "hbmpc_group_create")
        self.h = h
        self.size = len(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hbmpc_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard_range(self, B: int, i: int):
        lo, hi = C.c_size_t(), C.c_size_t()
        self.lib.hbmpc_group_shard_range(self.h, B, i, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    @staticmethod
    def _host(x):
        return np.ascontiguousarray(x, dtype=np.uint64)

    def compute_shares_batch(self, coeffs, n: int, out=None):
        c = self._host(coeffs)
        B, m = c.shape[0], c.shape[1]
        out = np.zeros((B, n, 4), dtype=np.uint64) if out is None else out
        rc = self.lib.hbmpc_group_compute_shares_batch(self.h, n, m - 1, B, c.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise HbmpcError(rc)
        return out

    def apply_vandermonde_batch(self, inp, n: int, recipient_major: bool = False, out=None):
        x = self._host(inp)
        B, cols = x.shape[0], x.shape[1]
        out = np.zeros((n, B, 4) if recipient_major else (B, n, 4), dtype=np.uint64) if out is None else out
        rc = self.lib.hbmpc_group_apply_vandermonde_batch(self.h, n, cols, B, x.ctypes.data, out.ctypes.data, int(recipient_major))
        if rc != 0:
            raise HbmpcError(rc)
        return out

    def batch_recover(self, sender_ids, evals, n: int, d: int, t: int, want_flags: bool = False, out=None):
        e = self._host(evals)
        S, B = e.shape[0], e.shape[1]
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        if out is None:
            coeffs, path = np.zeros((B, d + 1, 4), dtype=np.uint64), np.zeros((B,), dtype=np.int32)
            flags = np.zeros((B, (S + 63) // 64), dtype=np.uint64) if want_flags else None
        else:
            coeffs, path, flags = out
        rc = self.lib.hbmpc_group_batch_recover(self.h, n, d, t, S, ids.ctypes.data, B, e.ctypes.data, coeffs.ctypes.data, path.ctypes.data,
                                                flags.ctypes.data if flags is not None else None)
        if rc not in (0, DECODING_ERROR):
            raise HbmpcError(rc)
        return rc, coeffs, path, flags

    def batch_recover_secrets(self, sender_ids, evals, n: int, d: int, t: int):
        e = self._host(evals)
        S, B = e.shape[0], e.shape[1]
        ids = np.ascontiguousarray(sender_ids, dtype=np.uint64)
        secrets, path = np.zeros((B, 4), dtype=np.uint64), np.zeros((B,), dtype=np.int32)
        rc = self.lib.hbmpc_group_batch_recover_secrets(self.h, n, d, t, S, ids.ctypes.data, B, e.ctypes.data, secrets.ctypes.data, path.ctypes.data)
        if rc not in (0, DECODING_ERROR):
            raise HbmpcError(rc)
        return rc, secrets, path

    def robust_interpolate_batch(self, ids, shares, n: int, d: int, t: int, want_flags: bool = False):
        s = self._host(shares)
        B, S = s.shape[0], s.shape[1]
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        coeffs, secrets, path = np.zeros((B, d + 1, 4), dtype=np.uint64), np.zeros((B, 4), dtype=np.uint64), np.zeros((B,), dtype=np.int32)
        flags = np.zeros((B, (S + 63) // 64), dtype=np.uint64) if want_flags else None
        rc = self.lib.hbmpc_group_robust_interpolate_batch(self.h, n, d, t, S, idv.ctypes.data, B, s.ctypes.data, coeffs.ctypes.data, secrets.ctypes.data,
                                                           path.ctypes.data, flags.ctypes.data if flags is not None else None)
        if rc not in (0, DECODING_ERROR):
            raise HbmpcError(rc)
        return rc, coeffs, secrets, path, flags


def to_limbs(values) -> np.ndarray:
    arr = np.asarray(values, dtype=object)
    out = np.zeros(arr.shape + (4,), dtype=np.uint64)
    flat = out.reshape(-1, 4)
    for i, v in enumerate(arr.reshape(-1)):
        v = int(v)
        for k in range(4):
            flat[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_limbs(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, 4)
    vals = [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in flat]
    return np.asarray(vals, dtype=object).reshape(arr.shape[:-1]).tolist()
