"""N1: the per-sender / per-recipient vectors of batch reconstruction taken from and written to separate message payloads
(`Vec<F>::serialize_compressed`: u64 length + 32-byte LE canonical values, batch_recon.rs:174-175,339,419; common/utils.rs:3-21) through
hbmpc_batch_recover_msgs / hbmpc_batch_recover_secrets_msgs / hbmpc_apply_vandermonde_msgs == the contiguous-array entry points == oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def payload(values_u64):
    """ark-serialize of a Vec<F>: u64 element count, then the canonical 32-byte little-endian values"""
    b = bytearray(8 + values_u64.size * 8)
    b[:8] = int(values_u64.size // 4).to_bytes(8, "little")
    b[8:] = values_u64.tobytes()
    return b


def values_of(buf):
    return np.frombuffer(buf, dtype=np.uint64, offset=8)


@pytest.mark.parametrize("n,t,B", [(4, 1, 5), (16, 5, 3000), (64, 21, 70000)])
def test_recover_from_message_payloads(hb, ctx, orc, n, t, B):
    d = t
    rng = np.random.default_rng(n + B)
    coeffs = orc.random_fr((B, d + 1), 77 + n)
    rc, shares = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    assert rc == 0
    evals = np.ascontiguousarray(shares.transpose(1, 0, 2))          # [n][B][4]
    for S, bad in ((n, []), (n, [1]), (d + t + 1 + (1 if n > 4 else 0), [0] if t > 0 and n > 4 else [])):
        ids = rng.permutation(n)[:S]
        ev = evals[ids].copy()
        for j in bad:
            ev[j, :: 7, 0] ^= np.uint64(5)
        msgs = [payload(ev[j]) for j in range(S)]                     # one message per sender, as they arrive
        views = [values_of(m) for m in msgs]
        want = ctx.batch_recover(ids, ev, n, d, t, want_flags=True)
        got = ctx.batch_recover_msgs(ids, views, n, d, t, want_flags=True)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
        ref = orc.batch_recover_secret(ids, ev[:, :256], n, d, t)
        assert np.array_equal(got[1][:256], ref["coeffs"]) and np.array_equal(got[2][:256], ref["path"])
        got2 = ctx.batch_recover_msgs(ids, views, n, d, t)            # no flags: the lean upload of the examined senders only
        assert np.array_equal(got2[1], want[1]) and np.array_equal(got2[2], want[2])
        rc3, secrets, path3 = ctx.batch_recover_msgs(ids, views, n, d, t, secrets_only=True)
        assert np.array_equal(secrets, want[1][:, 0]) and np.array_equal(path3, want[2])


@pytest.mark.parametrize("n,t,B", [(4, 1, 3), (16, 5, 1000), (64, 21, 50000)])
def test_vandermonde_into_message_payloads(hb, ctx, orc, n, t, B):
    inp = orc.random_fr((B, t + 1), 5 + n)
    want = ctx.apply_vandermonde_batch(inp, n, recipient_major=True)   # [n][B][4]
    msgs = [bytearray(8 + B * 32) for _ in range(n)]
    for m in msgs:
        m[:8] = int(B).to_bytes(8, "little")
    views = [np.frombuffer(m, dtype=np.uint64, offset=8) for m in msgs]
    ctx.apply_vandermonde_msgs(inp, n, views)
    for j in range(n):
        assert np.array_equal(values_of(msgs[j]).reshape(B, 4), want[j]) and int.from_bytes(msgs[j][:8], "little") == B
    rc, ref = orc.apply_vandermonde(inp[:64], n, recipient_major=True)
    assert rc == 0 and np.array_equal(np.stack([values_of(m).reshape(B, 4)[:64] for m in msgs]), ref)


def test_message_variants_reject_bad_pointers(hb, ctx):
    import ctypes as C

    lib = ctx.lib
    ids = np.arange(4, dtype=np.uint64)
    out = np.zeros((2, 2, 4), dtype=np.uint64)
    path = np.zeros(2, dtype=np.int32)
    arr = (C.c_void_p * 4)(None, None, None, None)
    assert lib.hbmpc_batch_recover_msgs(ctx.h, 4, 1, 1, 4, ids.ctypes.data, 2, arr, out.ctypes.data, path.ctypes.data, None) == hb.INVALID_INPUT
    assert lib.hbmpc_batch_recover_msgs(ctx.h, 4, 1, 1, 4, ids.ctypes.data, 2, None, out.ctypes.data, path.ctypes.data, None) == hb.INVALID_INPUT
