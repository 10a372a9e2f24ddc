"""GPU parity of ntt16x_kernel (ntt16x.cuh: the barrier-free, single-product-site transform for N = 16 .. 128 that serves large
batches of K1 / K2 / the inverse-transform checks of K3 / a10): forced for every batch size with HBMPC_NTT16X=2 and compared with the
oracle bit for bit -- every domain size, party counts that are not powers of two, every column count the callers use (structurally
zero operands), ragged tiles, recipient-major outputs, corrupted shares (fail flags -> decoder), sender subsets (erasure-weighted
transform, MODE 2), secrets-only outputs and the non-robust degree check."""
import numpy as np
import pytest

from test_gpu_parity import _codewords, _compare_recover, _corrupt, _rand

pytestmark = pytest.mark.gpu


@pytest.fixture
def x_ctx(hb, monkeypatch):
    monkeypatch.setenv("HBMPC_NTT16X", "2")
    c = hb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [16, 13, 32, 20, 64, 50, 128, 100])
@pytest.mark.parametrize("B", [1, 7, 33, 1000])
def test_x_compute_shares(x_ctx, orc, n, B):
    N = 1 << (n - 1).bit_length()
    for d in sorted({0, 1, (n - 1) // 3, 2 * ((n - 1) // 3), n // 2, n - 1}):
        coeffs = _rand(orc, (B, d + 1), 0x5EED9000 + n * 7 + B + d)
        rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
        assert rc == 0
        l0 = x_ctx.launch_count
        got = x_ctx.compute_shares_batch(coeffs, n)
        assert x_ctx.launch_count - l0 == 1
        assert np.array_equal(got, want), (n, N, d, B)


@pytest.mark.parametrize("n,cols", [(16, 6), (16, 16), (32, 11), (64, 1), (64, 22), (64, 43), (64, 64), (50, 17), (33, 33), (128, 43), (128, 128), (100, 67)])
@pytest.mark.parametrize("recipient_major", [False, True])
def test_x_apply_vandermonde(x_ctx, orc, n, cols, recipient_major):
    B = 1234
    x = _rand(orc, (B, cols), 0x5EED9100 + n + cols)
    rc, want = orc.apply_vandermonde(x, n, recipient_major, threads=orc.max_threads())
    assert rc == 0
    got = x_ctx.apply_vandermonde_batch(x, n, recipient_major)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,t", [(16, 5), (32, 10), (64, 21), (128, 42)])
@pytest.mark.parametrize("deg_mult", [1, 2])
def test_x_batch_recover_all_senders(x_ctx, orc, n, t, deg_mult):
    """all n = N senders supplied: inverse transform + degree check (MODE 1); clean chunks, chunks with <= t errors (decoder), > t errors"""
    d = t * deg_mult
    B = 700 if n < 128 else 150
    rng = np.random.default_rng(n + deg_mult)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED9200 + n + deg_mult)
    nerr = np.where(rng.random(B) < 0.9, 0, rng.integers(1, 4, size=B))
    bad = _corrupt(shares, rng, nerr)
    arrival = rng.permutation(n)
    evals = np.ascontiguousarray(bad[:, arrival].transpose(1, 0, 2))
    want = orc.batch_recover_secret(arrival, evals, n, d, t, threads=orc.max_threads())
    for fl in (True, False):
        _compare_recover(x_ctx.batch_recover(arrival, evals, n, d, t, want_flags=fl), want, B)
    rc, sec, path = x_ctx.batch_recover_secrets(arrival, evals, n, d, t)
    assert rc == want["rc"] and np.array_equal(path, want["path"]) and np.array_equal(sec, want["coeffs"][:, 0])


@pytest.mark.parametrize("n,t,S", [(16, 5, 11), (16, 5, 13), (64, 21, 43), (64, 21, 50), (50, 16, 40), (128, 42, 85), (128, 42, 100)])
def test_x_batch_recover_sender_subsets(x_ctx, orc, n, t, S):
    """S < n senders (the first-call shape of batch reconstruction: exactly d+t+1 arrivals, and more): erasure-weighted inverse
    transform (MODE 2) + triangular recovery, with and without flags, secrets only; errors inside and beyond the examined prefix"""
    d = t
    B = 600 if n < 128 else 120
    rng = np.random.default_rng(S * 31 + n)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED9300 + n + S)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]]
    nerr = np.where(rng.random(B) < 0.85, 0, rng.integers(1, t + 2, size=B))
    bad = _corrupt(words, rng, np.minimum(nerr, S))
    evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
    want = orc.batch_recover_secret(ids[arrival], evals, n, d, t, threads=orc.max_threads())
    for fl in (False, True):
        _compare_recover(x_ctx.batch_recover(ids[arrival], evals, n, d, t, want_flags=fl), want, B)
    rc, sec, path = x_ctx.batch_recover_secrets(ids[arrival], evals, n, d, t)
    assert rc == want["rc"] and np.array_equal(path, want["path"])
    ok = want["path"] >= 0
    assert np.array_equal(sec[ok], want["coeffs"][ok, 0])


def test_x_nonrobust_recover(x_ctx, orc, hb):
    """NonRobustShare::recover_secret with all n = N shares: inverse transform, DegreeMismatch == non-zero top coefficient"""
    for n, t in ((16, 5), (64, 21), (128, 42)):
        B = 300
        for deg in (t, 2 * t):
            coeffs, shares = _codewords(orc, n, deg, B, 0x5EED9400 + deg + n)
            bad = shares.copy()
            bad[::7, 5, 0] ^= np.uint64(9)          # every 7th sharing is not of degree `deg`
            rng = np.random.default_rng(n + deg)
            arrival = rng.permutation(n)
            co, sec, status = x_ctx.nonrobust_recover_batch(arrival, bad[:, arrival], n, deg)
            assert (status[::7] == -hb.DEGREE_MISMATCH).all() and (np.delete(status, np.s_[::7]) == deg).all()
            for b in range(24):
                ref = orc.nonrobust_recover_secret(arrival, bad[b, arrival], n, deg)
                if ref["rc"] == 0:
                    assert np.array_equal(co[b], ref["coeffs"]) and np.array_equal(sec[b], ref["secret"])
                else:
                    assert status[b] == -hb.DEGREE_MISMATCH and not co[b].any()


def test_x_rejects_non_canonical_input(x_ctx, hb):
    n, d, B = 64, 21, 100
    coeffs = np.zeros((B, d + 1, 4), dtype=np.uint64)
    coeffs[37, 3] = np.array([0xFFFFFFFF00000001, 0x53BDA402FFFE5BFE, 0x3339D80809A1D805, 0x73EDA753299D7D48], dtype=np.uint64)  # == r
    with pytest.raises(hb.HbmpcError) as e:
        x_ctx.compute_shares_batch(coeffs, n)
    assert e.value.code == hb.INVALID_INPUT
