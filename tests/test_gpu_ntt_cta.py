"""GPU parity of ntt64_cta_kernel (the CTA-cooperative 64-point transform of K1 / K2 / the all-points check of K3): forced for every
batch size with HBMPC_NTT_CTA=2 and compared with the oracle bit for bit, including ragged tiles (16 items per tile), every column
count the callers use, recipient-major outputs, corrupted shares (fail flags -> decoder) and the non-robust degree check."""
import numpy as np
import pytest

from test_gpu_parity import _codewords, _compare_recover, _corrupt, _rand

pytestmark = pytest.mark.gpu


@pytest.fixture
def cta_ctx(hb, monkeypatch):
    monkeypatch.setenv("HBMPC_NTT_CTA", "2")
    c = hb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("B", [1, 15, 16, 17, 777, 3001])
@pytest.mark.parametrize("d", [21, 42, 0, 5, 62])
def test_cta_compute_shares(cta_ctx, orc, B, d):
    n = 64
    coeffs = _rand(orc, (B, d + 1), 0x5EED8000 + B + d)
    rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    assert rc == 0
    l0 = cta_ctx.launch_count
    got = cta_ctx.compute_shares_batch(coeffs, n)
    assert cta_ctx.launch_count - l0 == 1
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,cols", [(64, 1), (64, 22), (64, 43), (64, 64), (50, 17), (33, 33)])
@pytest.mark.parametrize("recipient_major", [False, True])
def test_cta_apply_vandermonde(cta_ctx, orc, n, cols, recipient_major):
    B = 1234
    x = _rand(orc, (B, cols), 0x5EED8100 + n + cols)
    rc, want = orc.apply_vandermonde(x, n, recipient_major, threads=orc.max_threads())
    assert rc == 0
    got = cta_ctx.apply_vandermonde_batch(x, n, recipient_major)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("deg_mult", [1, 2])
def test_cta_batch_recover_all_senders(cta_ctx, orc, deg_mult):
    """all 64 senders supplied: inverse transform + degree check; clean chunks, chunks with <= t errors (decoder) and > t errors"""
    n, t = 64, 21
    d = t * deg_mult
    B = 2100
    rng = np.random.default_rng(64 + deg_mult)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED8200 + deg_mult)
    nerr = np.where(rng.random(B) < 0.9, 0, rng.integers(1, 4, size=B))
    bad = _corrupt(shares, rng, nerr)
    arrival = rng.permutation(n)
    evals = np.ascontiguousarray(bad[:, arrival].transpose(1, 0, 2))
    want = orc.batch_recover_secret(arrival, evals, n, d, t, threads=orc.max_threads())
    for fl in (True, False):
        _compare_recover(cta_ctx.batch_recover(arrival, evals, n, d, t, want_flags=fl), want, B)
    rc, sec, path = cta_ctx.batch_recover_secrets(arrival, evals, n, d, t)
    assert rc == want["rc"] and np.array_equal(path, want["path"]) and np.array_equal(sec, want["coeffs"][:, 0])


def test_cta_nonrobust_recover(cta_ctx, ctx, orc, hb):
    """NonRobustShare::recover_secret with all n = 64 shares: inverse transform, DegreeMismatch == non-zero top coefficient.
    Compared with the warp-per-item kernel's outputs (same library, default route) and, on a sample, with the oracle item by item."""
    n, t, B = 64, 21, 1500
    for deg in (t, 2 * t):
        coeffs, shares = _codewords(orc, n, deg, B, 0x5EED8300 + deg)
        bad = shares.copy()
        bad[::7, 5, 0] ^= np.uint64(9)          # every 7th sharing is not of degree `deg`
        ids = np.arange(n)
        co, sec, status = cta_ctx.nonrobust_recover_batch(ids, bad, n, deg)
        co0, sec0, status0 = ctx.nonrobust_recover_batch(ids, bad, n, deg)
        assert np.array_equal(status, status0) and np.array_equal(co, co0) and np.array_equal(sec, sec0)
        assert (status[::7] == -hb.DEGREE_MISMATCH).all() and (np.delete(status, np.s_[::7]) == deg).all()
        for b in range(16):
            ref = orc.nonrobust_recover_secret(ids, bad[b], n, deg)
            if ref["rc"] == 0:
                assert np.array_equal(co[b], ref["coeffs"]) and np.array_equal(sec[b], ref["secret"])
            else:
                assert status[b] == -hb.DEGREE_MISMATCH and not co[b].any()
