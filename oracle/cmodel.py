"""ctypes loader for the C oracle (oracle/hbmpc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Field arrays are numpy uint64 with a trailing dimension of 4 (canonical little-endian limbs == U256,
/root/reference/mpc/src/ffi/c_bindings/mod.rs:17-49).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_PATH = None

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def build(native: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory).  Returns the .so path."""
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")


def load(native: bool = False):
    global _LIB, _LIB_PATH
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    if _LIB is not None and _LIB_PATH == path:
        return _LIB
    src = os.path.join(_HERE, "hbmpc_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build(native)
    lib = C.CDLL(path)
    u64p, szp, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)
    sz, ci = C.c_size_t, C.c_int
    lib.orc_compute_shares.argtypes = [sz, sz, sz, u64p, u64p, ci]
    lib.orc_make_vandermonde.argtypes = [sz, sz, u64p]
    lib.orc_apply_matrix.argtypes = [sz, sz, u64p, sz, u64p, u64p, ci, ci]
    lib.orc_apply_vandermonde.argtypes = [sz, sz, sz, u64p, u64p, ci, ci]
    lib.orc_robust_recover_secret.argtypes = [sz, sz, sz, sz, szp, u64p, u64p, szp, u64p, i32p, u64p]
    lib.orc_robust_interpolate_batch.argtypes = [sz, sz, sz, sz, szp, sz, u64p, u64p, u64p, i32p, u64p, ci]
    lib.orc_batch_recover_secret.argtypes = [sz, sz, sz, sz, szp, sz, u64p, u64p, szp, i32p, u64p, ci]
    lib.orc_nonrobust_recover_secret.argtypes = [sz, sz, sz, szp, u64p, u64p, szp, u64p]
    lib.orc_gao_rs_decode.argtypes = [sz, sz, u64p, szp, sz, u64p, szp]
    lib.orc_lagrange_interpolate.argtypes = [sz, u64p, u64p, u64p, szp]
    lib.orc_elementwise.argtypes = [ci, sz, u64p, u64p, u64p, ci]
    lib.orc_share_algebra_step.argtypes = [ci, sz, C.POINTER(u64p), C.POINTER(u64p)]
    lib.orc_domain_element.argtypes = [sz, sz, u64p]
    lib.orc_domain_element.restype = None
    lib.orc_max_threads.restype = ci
    _LIB, _LIB_PATH = lib, path
    return lib


# ------------------------------------------------------------------ limb helpers
def to_limbs(values) -> np.ndarray:
    """ints (any nesting as a flat/ nested list) -> uint64[..., 4]"""
    arr = np.asarray(values, dtype=object)
    out = np.zeros(arr.shape + (4,), dtype=np.uint64)
    flat = out.reshape(-1, 4)
    for i, v in enumerate(arr.reshape(-1)):
        v = int(v)
        for k in range(4):
            flat[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_limbs(arr: np.ndarray):
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, 4)
    vals = [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in flat]
    return np.asarray(vals, dtype=object).reshape(arr.shape[:-1]).tolist()


def _p(a, ty=C.c_uint64):
    return a.ctypes.data_as(C.POINTER(ty))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ids(ids):
    return np.ascontiguousarray(ids, dtype=np.uint64)  # size_t == uint64 on LP64


# ------------------------------------------------------------------ wrappers
def compute_shares(coeffs: np.ndarray, n: int, threads: int = 1):
    """coeffs uint64[B][d+1][4] -> (rc, shares uint64[B][n][4])"""
    lib = load()
    coeffs = _u64(coeffs)
    B, m = coeffs.shape[0], coeffs.shape[1]
    out = np.zeros((B, n, 4), dtype=np.uint64)
    rc = lib.orc_compute_shares(n, m - 1, B, _p(coeffs), _p(out), threads)
    return rc, out


def make_vandermonde(n: int, t: int):
    lib = load()
    V = np.zeros((n, t + 1, 4), dtype=np.uint64)
    rc = lib.orc_make_vandermonde(n, t, _p(V))
    return rc, V


def apply_vandermonde(inp: np.ndarray, n: int, recipient_major: bool = False, threads: int = 1):
    lib = load()
    inp = _u64(inp)
    B, cols = inp.shape[0], inp.shape[1]
    out = np.zeros((n, B, 4) if recipient_major else (B, n, 4), dtype=np.uint64)
    rc = lib.orc_apply_vandermonde(n, cols, B, _p(inp), _p(out), int(recipient_major), threads)
    return rc, out


def apply_matrix(M: np.ndarray, inp: np.ndarray, recipient_major: bool = False, threads: int = 1):
    lib = load()
    M, inp = _u64(M), _u64(inp)
    rows, cols = M.shape[0], M.shape[1]
    B = inp.shape[0]
    out = np.zeros((rows, B, 4) if recipient_major else (B, rows, 4), dtype=np.uint64)
    rc = lib.orc_apply_matrix(rows, cols, _p(M), B, _p(inp), _p(out), int(recipient_major), threads)
    return rc, out


def robust_recover_secret(ids, vals: np.ndarray, n: int, t: int, degree: int):
    """One codeword.  -> dict(rc, coeffs[d+1][4], coeff_len, secret[4], path, flags[ceil(S/64)])"""
    lib = load()
    ids, vals = _ids(ids), _u64(vals)
    S = len(ids)
    coeffs = np.zeros((degree + 1, 4), dtype=np.uint64)
    secret = np.zeros(4, dtype=np.uint64)
    flags = np.zeros(max(1, (S + 63) // 64), dtype=np.uint64)
    cl, path = C.c_size_t(0), C.c_int32(0)
    rc = lib.orc_robust_recover_secret(n, t, degree, S, _p(ids, C.c_size_t), _p(vals), _p(coeffs), C.byref(cl), _p(secret), C.byref(path), _p(flags))
    return dict(rc=rc, coeffs=coeffs, coeff_len=cl.value, secret=secret, path=path.value, flags=flags)


def robust_interpolate_batch(ids, shares: np.ndarray, n: int, degree: int, t: int, threads: int = 1):
    """shares uint64[B][S][4] -> dict(rc, coeffs[B][d+1][4], secrets[B][4], path[B], flags[B][fw])"""
    lib = load()
    ids, shares = _ids(ids), _u64(shares)
    B, S = shares.shape[0], shares.shape[1]
    fw = max(1, (S + 63) // 64)
    coeffs = np.zeros((B, degree + 1, 4), dtype=np.uint64)
    secrets = np.zeros((B, 4), dtype=np.uint64)
    path = np.zeros(B, dtype=np.int32)
    flags = np.zeros((B, fw), dtype=np.uint64)
    rc = lib.orc_robust_interpolate_batch(n, degree, t, S, _p(ids, C.c_size_t), B, _p(shares), _p(coeffs), _p(secrets), _p(path, C.c_int32), _p(flags), threads)
    return dict(rc=rc, coeffs=coeffs, secrets=secrets, path=path, flags=flags)


def batch_recover_secret(sender_ids, evals: np.ndarray, n: int, degree: int, t: int, threads: int = 1):
    """evals uint64[S][B][4] sender-major -> dict(rc, coeffs[B][d+1][4], coeff_len[B], path[B], flags[B][fw])"""
    lib = load()
    ids, evals = _ids(sender_ids), _u64(evals)
    S, B = evals.shape[0], evals.shape[1]
    fw = max(1, (S + 63) // 64)
    coeffs = np.zeros((B, degree + 1, 4), dtype=np.uint64)
    clen = np.zeros(B, dtype=np.uint64)
    path = np.zeros(B, dtype=np.int32)
    flags = np.zeros((B, fw), dtype=np.uint64)
    rc = lib.orc_batch_recover_secret(n, degree, t, S, _p(ids, C.c_size_t), B, _p(evals), _p(coeffs), _p(clen, C.c_size_t), _p(path, C.c_int32), _p(flags), threads)
    return dict(rc=rc, coeffs=coeffs, coeff_len=clen, path=path, flags=flags)


def nonrobust_recover_secret(ids, vals: np.ndarray, n: int, deg: int):
    lib = load()
    ids, vals = _ids(ids), _u64(vals)
    coeffs = np.zeros((deg + 1, 4), dtype=np.uint64)
    secret = np.zeros(4, dtype=np.uint64)
    cl = C.c_size_t(0)
    rc = lib.orc_nonrobust_recover_secret(n, deg, len(ids), _p(ids, C.c_size_t), _p(vals), _p(coeffs), C.byref(cl), _p(secret))
    return dict(rc=rc, coeffs=coeffs, coeff_len=cl.value, secret=secret)


def gao_rs_decode(received: np.ndarray, k: int, n: int, erasures):
    lib = load()
    received = _u64(received)
    er = _ids(erasures) if len(erasures) else np.zeros(1, dtype=np.uint64)
    coeffs = np.zeros((k, 4), dtype=np.uint64)
    cl = C.c_size_t(0)
    rc = lib.orc_gao_rs_decode(n, k, _p(received), _p(er, C.c_size_t), len(erasures), _p(coeffs), C.byref(cl))
    return dict(rc=rc, coeffs=coeffs, coeff_len=cl.value)


def lagrange_interpolate(xs: np.ndarray, ys: np.ndarray):
    lib = load()
    xs, ys = _u64(xs), _u64(ys)
    k = xs.shape[0]
    coeffs = np.zeros((k, 4), dtype=np.uint64)
    cl = C.c_size_t(0)
    rc = lib.orc_lagrange_interpolate(k, _p(xs), _p(ys), _p(coeffs), C.byref(cl))
    return dict(rc=rc, coeffs=coeffs, coeff_len=cl.value)


def elementwise(op: int, a: np.ndarray, b: np.ndarray, threads: int = 1):
    lib = load()
    a, b = _u64(a), _u64(b)
    out = np.zeros_like(a)
    rc = lib.orc_elementwise(op, a.size // 4, _p(a), _p(b), _p(out), threads)
    return rc, out


def share_algebra_step(step: int, inputs):
    """The multi-operator steps of the share algebra in the reference's operator order (orc_share_algebra_step): step 0
    (a, b, r_2t) -> a*b - r_2t; step 1 (a, x, b, y) -> (a - x, b - y); step 2 (c, x, y, a-x, b-y) -> the Beaver product share.
    Returns (rc, outputs)."""
    lib = load()
    ins = [_u64(v) for v in inputs]
    nin, nout = {0: (3, 1), 1: (4, 2), 2: (5, 1)}[step]
    assert len(ins) == nin and all(v.shape == ins[0].shape for v in ins)
    outs = [np.zeros_like(ins[0]) for _ in range(nout)]
    u64p = C.POINTER(C.c_uint64)
    pin = (u64p * nin)(*[_p(v) for v in ins])
    pout = (u64p * nout)(*[_p(v) for v in outs])
    rc = lib.orc_share_algebra_step(step, ins[0].size // 4, pin, pout)
    return rc, outs


def domain_element(n: int, j: int) -> int:
    lib = load()
    out = np.zeros(4, dtype=np.uint64)
    lib.orc_domain_element(n, j, _p(out))
    return from_limbs(out)


def max_threads() -> int:
    return load().orc_max_threads()


# ------------------------------------------------------------------ synthetic inputs (SplitMix64 + rejection, vectorised)
def random_fr(shape, seed: int) -> np.ndarray:
    """Uniform canonical Fr values, uint64[shape..., 4]; deterministic in (shape, seed).  Counter-based SplitMix64
    per 64-bit word, top limb masked to 63 bits... then rejection (value >= r redrawn from the next counter block)."""
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    count = int(np.prod(shape)) if shape else 1
    out = np.zeros((count, 4), dtype=np.uint64)
    todo = np.arange(count)
    rnd = 0
    mod = np.array([(R_MOD >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)
    while todo.size:
        ctr = np.uint64(((seed * 0x100000001B3) + rnd * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        base = (todo.astype(np.uint64) * np.uint64(4) + ctr)
        limbs = np.stack([_splitmix(base + np.uint64(k)) for k in range(4)], axis=1)
        limbs[:, 3] >>= np.uint64(1)
        ok = _lt(limbs, mod)
        out[todo[ok]] = limbs[ok]
        todo = todo[~ok]
        rnd += 1
    return out.reshape(shape + (4,))


def _splitmix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _lt(limbs: np.ndarray, mod: np.ndarray) -> np.ndarray:
    lt = np.zeros(limbs.shape[0], dtype=bool)
    eq = np.ones(limbs.shape[0], dtype=bool)
    for k in (3, 2, 1, 0):
        lt |= eq & (limbs[:, k] < mod[k])
        eq &= limbs[:, k] == mod[k]
    return lt
