"""GPU parity: every kernel, through the C ABI, bit-exact against the CPU oracle on the same seeded inputs.

Scenarios follow SURVEY.md section 4a (the reference's own test matrix) plus the BASELINE shapes at reduced batch.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [(4, 1), (7, 2), (10, 3), (16, 5), (64, 21), (128, 42)]


def _rand(orc, shape, seed):
    return orc.random_fr(shape, seed)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("n,t", CONFIGS + [(5, 1), (13, 4), (255, 84)])
@pytest.mark.parametrize("deg_mult", [1, 2])
def test_compute_shares_matches_oracle(ctx, orc, n, t, deg_mult):
    d = t * deg_mult
    B = 777 if n <= 64 else 130
    coeffs = _rand(orc, (B, d + 1), 0x5EED0001 + n * 3 + deg_mult)
    rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    assert rc == 0
    got = ctx.compute_shares_batch(coeffs, n)
    assert np.array_equal(got, want)


def test_compute_shares_fixed_polynomial_kat(ctx, hb):
    # robust_interpolate.rs:646-680: f = 7 + 3x + 5x^2, n = 16; share KATs from SURVEY.md 8c.3
    coeffs = hb.to_limbs([[7, 3, 5]])
    got = hb.from_limbs(ctx.compute_shares_batch(coeffs, 16))[0]
    assert got[0] == 0xF
    assert got[1] == 0x0C017888577EFB9AB9C555D3B5BC31E1CDBAD5FC7989CD68368C178E5B0D73D0
    assert got[2] == 0x29188D8EE251B7713371996896F3B6B762F2C150E34BC902567FF79BC65CBE74
    assert got[3] == 0x48AA887102257E1488DBC18CBF4C77E437F86F54D8A23928628A7BD8EFBCE893


def test_compute_shares_errors(ctx, hb, orc):
    c = _rand(orc, (4, 6), 1)
    with pytest.raises(hb.HbmpcError) as e:
        ctx.compute_shares_batch(c, 5)  # n <= d
    assert e.value.code == hb.INVALID_INPUT
    with pytest.raises(hb.HbmpcError) as e:
        ctx.compute_shares_batch(c, 300)
    assert e.value.code == hb.NO_SUITABLE_DOMAIN
    bad = c.copy()
    bad[2, 3] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)  # >= r
    with pytest.raises(hb.HbmpcError) as e:
        ctx.compute_shares_batch(bad, 16)
    assert e.value.code == hb.INVALID_INPUT
    # r itself is non-canonical, r-1 is fine
    edge = hb.to_limbs([[hb.R_MOD - 1, hb.R_MOD - 1, 0, 0, 0, 1]])
    rc, want = orc.compute_shares(edge, 16)
    assert np.array_equal(ctx.compute_shares_batch(edge, 16), want)
    edge = hb.to_limbs([[hb.R_MOD, 0, 0, 0, 0, 1]])
    with pytest.raises(hb.HbmpcError):
        ctx.compute_shares_batch(edge, 16)


# ------------------------------------------------------------------ K2
@pytest.mark.parametrize("n,cols", [(4, 3), (4, 2), (10, 4), (16, 6), (64, 22), (64, 43), (64, 64), (128, 43), (7, 7), (255, 85)])
@pytest.mark.parametrize("recipient_major", [False, True])
def test_apply_vandermonde_matches_oracle(ctx, orc, n, cols, recipient_major):
    B = 301 if n <= 64 else 70
    x = _rand(orc, (B, cols), 0x5EED0100 + n + cols)
    rc, want = orc.apply_vandermonde(x, n, recipient_major, threads=orc.max_threads())
    assert rc == 0
    got = ctx.apply_vandermonde_batch(x, n, recipient_major)
    assert np.array_equal(got, want)


def test_apply_vandermonde_kat_n4(ctx, hb, orc):
    # common/share/mod.rs:139-171: coefficients [1,2,3] -> y_j = 1 + 2 a_j + 3 a_j^2 on the n=4 domain
    x = hb.to_limbs([[1, 2, 3]])
    got = hb.from_limbs(ctx.apply_vandermonde_batch(x, 4))[0]
    w = orc.domain_element(4, 1)
    for j in range(4):
        a = pow(w, j, hb.R_MOD)
        assert got[j] == (1 + 2 * a + 3 * a * a) % hb.R_MOD
    # tests/batchrecon_test.rs:28: secrets [3,4], n=4,t=1 -> y_j = 3 + 4 w^j
    y = hb.from_limbs(ctx.apply_vandermonde_batch(hb.to_limbs([[3, 4]]), 4))[0]
    assert y[0] == 7 and y[2] == hb.R_MOD - 1
    assert y[1] == 0x0235473339D80C1343B00C0009D80C00000004000000000003 or y[1] == (3 + 4 * w) % hb.R_MOD


def test_apply_matrix_matches_oracle(ctx, orc):
    M = _rand(orc, (9, 5), 77)
    x = _rand(orc, (100, 5), 78)
    rc, want = orc.apply_matrix(M, x)
    assert rc == 0
    assert np.array_equal(ctx.apply_matrix_batch(M, x), want)


# ------------------------------------------------------------------ K5
@pytest.mark.parametrize("op", [0, 1, 2])
def test_elementwise_matches_oracle(ctx, orc, hb, op):
    a = _rand(orc, (5000,), 5)
    b = _rand(orc, (5000,), 6)
    a[0] = hb.to_limbs(hb.R_MOD - 1)
    b[0] = hb.to_limbs(hb.R_MOD - 1)
    a[1] = hb.to_limbs(0)
    rc, want = orc.elementwise(op, a, b)
    assert rc == 0
    assert np.array_equal(ctx.elementwise(op, a, b), want)


# ------------------------------------------------------------------ K3 / K4
def _codewords(orc, n, d, B, seed):
    coeffs = _rand(orc, (B, d + 1), seed)
    rc, shares = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    assert rc == 0
    return coeffs, shares  # shares[B][n]


def _corrupt(shares, rng, n_err_per_item, positions=None):
    """add nonzero offsets at random positions; returns corrupted copy"""
    out = shares.copy()
    B, S = shares.shape[0], shares.shape[1]
    for b in range(B):
        e = n_err_per_item[b]
        if e == 0:
            continue
        pos = rng.choice(S if positions is None else positions, size=e, replace=False)
        for p in pos:
            out[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 20))  # stays < r: only low limb changes
    return out


def _compare_recover(got, want, B):
    rc_g, co_g, path_g, fl_g = got
    assert rc_g == want["rc"], (rc_g, want["rc"])
    assert np.array_equal(path_g, want["path"]), np.nonzero(path_g != want["path"])
    assert np.array_equal(co_g, want["coeffs"])
    if fl_g is not None:
        assert np.array_equal(fl_g, want["flags"][:, : fl_g.shape[1]])


@pytest.mark.parametrize("n,t", CONFIGS)
@pytest.mark.parametrize("deg_mult", [1, 2])
def test_batch_recover_honest(ctx, orc, n, t, deg_mult):
    d = t * deg_mult
    if d + t + 1 > n:
        pytest.skip("needed > n")
    B = 257 if n <= 64 else 66
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED0200 + n + deg_mult)
    evals = np.ascontiguousarray(shares.transpose(1, 0, 2))  # [S][B]
    ids = np.arange(n)
    want = orc.batch_recover_secret(ids, evals, n, d, t, threads=orc.max_threads())
    got = ctx.batch_recover(ids, evals, n, d, t, want_flags=True)
    _compare_recover(got, want, B)
    assert np.array_equal(got[1], coeffs)
    assert not got[2].any()
    # reversed arrival order (robust_interpolate.rs:880-927) and secrets-only variant
    rev = ids[::-1].copy()
    got_r = ctx.batch_recover(rev, np.ascontiguousarray(evals[::-1]), n, d, t, want_flags=False)
    assert np.array_equal(got_r[1], coeffs)
    rc, secrets, path = ctx.batch_recover_secrets(ids, evals, n, d, t)
    assert rc == 0 and np.array_equal(secrets, coeffs[:, 0]) and not path.any()


@pytest.mark.parametrize("n,t", [(7, 2), (10, 3), (16, 5), (64, 21)])
def test_batch_recover_with_errors(ctx, orc, n, t):
    d = t
    B = 96 if n <= 16 else 48
    rng = np.random.default_rng(n * 1000 + t)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED0300 + n)
    nerr = rng.integers(0, t + 1, size=B)
    bad = _corrupt(shares, rng, nerr)
    evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
    ids = np.arange(n)
    want = orc.batch_recover_secret(ids, evals, n, d, t, threads=orc.max_threads())
    got = ctx.batch_recover(ids, evals, n, d, t, want_flags=True)
    _compare_recover(got, want, B)
    assert np.array_equal(got[1], coeffs)  # <= t errors with all n shares: always decodable
    # permuted arrival order, same answer; flags follow arrival order
    perm = rng.permutation(n)
    want_p = orc.batch_recover_secret(ids[perm], np.ascontiguousarray(evals[perm]), n, d, t, threads=orc.max_threads())
    got_p = ctx.batch_recover(ids[perm], np.ascontiguousarray(evals[perm]), n, d, t, want_flags=True)
    _compare_recover(got_p, want_p, B)
    rc, secrets, path = ctx.batch_recover_secrets(ids, evals, n, d, t)
    assert np.array_equal(secrets, coeffs[:, 0]) and np.array_equal(path, want["path"])


def test_batch_recover_first_t_senders_corrupted(ctx, orc):
    # robust_interpolate.rs:931-967: first t senders corrupted, distinct error per chunk
    n, t, B = 10, 3, 16
    coeffs, shares = _codewords(orc, n, t, B, 4242)
    bad = shares.copy()
    for b in range(B):
        for p in range(t):
            bad[b, p, 0] += np.uint64(1 + b + p)
    evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
    want = orc.batch_recover_secret(np.arange(n), evals, n, t, t)
    got = ctx.batch_recover(np.arange(n), evals, n, t, t, want_flags=True)
    _compare_recover(got, want, B)
    assert np.array_equal(got[1], coeffs)


@pytest.mark.parametrize("n,t", [(7, 2), (10, 3)])
def test_robust_all_error_subsets(ctx, orc, hb, n, t):
    # robust_interpolate.rs:828-876 (all <= t subsets, n=7,t=2) and :728-756 (all triples, n=10,t=3), secret 42
    d = t
    coeffs = _rand(orc, (1, d + 1), 99)
    coeffs[0, 0] = hb.to_limbs(42)
    rc, sh = orc.compute_shares(coeffs, n)
    subsets = [s for k in range(0, t + 1) for s in itertools.combinations(range(n), k)]
    words = np.repeat(sh, len(subsets), axis=0)
    for b, sub in enumerate(subsets):
        for p in sub:
            words[b, p, 0] += np.uint64(1000 + p)
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, words, n, d, t, threads=orc.max_threads())
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, words, n, d, t, want_flags=True)
    assert rc == want["rc"] == 0
    assert np.array_equal(path, want["path"])
    assert np.array_equal(co, want["coeffs"]) and np.array_equal(sec, want["secrets"])
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
    assert all(v == 42 for v in hb.from_limbs(sec))
    for b, sub in enumerate(subsets):
        assert int(flags[b, 0]) == sum(1 << p for p in sub)


@pytest.mark.parametrize("n,t,S", [(10, 3, 7), (10, 3, 8), (10, 3, 9), (10, 3, 10), (13, 4, 11), (16, 5, 11), (16, 5, 13), (64, 21, 50), (64, 21, 43), (128, 42, 100)])
def test_robust_subset_of_senders(ctx, orc, n, t, S):
    """S < n supplied shares, ids not starting at 0, error counts up to and beyond what the prefix rounds can absorb:
    the GPU must reproduce the oracle's path / DecodingError pattern exactly."""
    d = t
    B = 64
    rng = np.random.default_rng(S * 77 + n)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED0400 + n + S)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]]
    nerr = rng.integers(0, t + 2, size=B)
    nerr = np.minimum(nerr, S)
    bad = _corrupt(words, rng, nerr)
    want = orc.robust_interpolate_batch(ids[arrival], bad, n, d, t, threads=orc.max_threads())
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids[arrival], bad, n, d, t, want_flags=True)
    assert np.array_equal(path, want["path"]), (path, want["path"])
    assert rc == want["rc"]
    assert np.array_equal(co, want["coeffs"]) and np.array_equal(sec, want["secrets"])
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
    evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
    want_b = orc.batch_recover_secret(ids[arrival], evals, n, d, t, threads=orc.max_threads())
    got_b = ctx.batch_recover(ids[arrival], evals, n, d, t, want_flags=True)
    _compare_recover(got_b, want_b, B)
    # same calls without flags: the optimistic check runs as an erasure-weighted inverse NTT + triangular recovery
    got_nf = ctx.batch_recover(ids[arrival], evals, n, d, t, want_flags=False)
    _compare_recover(got_nf, want_b, B)
    rc2, co2, sec2, path2, _ = ctx.robust_interpolate_batch(ids[arrival], bad, n, d, t, want_flags=False)
    assert rc2 == want["rc"] and np.array_equal(path2, want["path"]) and np.array_equal(co2, want["coeffs"]) and np.array_equal(sec2, want["secrets"])
    rc3, sec3, path3 = ctx.batch_recover_secrets(ids[arrival], evals, n, d, t)
    assert rc3 == want["rc"] and np.array_equal(path3, want["path"]) and np.array_equal(sec3, want["secrets"])


def test_robust_more_than_t_errors(ctx, orc):
    """> t errors (tests/input_test.rs:204: 4 errors with t=3, n=10): outcome must equal the oracle's, item by item."""
    n, t, d, B = 10, 3, 3, 200
    rng = np.random.default_rng(5)
    coeffs, shares = _codewords(orc, n, d, B, 31337)
    nerr = rng.integers(t + 1, t + 3, size=B)
    bad = _corrupt(shares, rng, nerr)
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
    assert np.array_equal(path, want["path"])
    assert rc == want["rc"]
    assert np.array_equal(co, want["coeffs"])
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])


def test_robust_errors_only_beyond_prefix_take_path0(ctx, orc):
    n, t, d, B = 16, 5, 5, 8
    coeffs, shares = _codewords(orc, n, d, B, 2024)
    bad = shares.copy()
    bad[:, 13, 0] += np.uint64(5)  # id 13 >= d+t+1 = 11: never examined
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, bad, n, d, t)
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
    assert rc == 0 and not path.any() and np.array_equal(co, coeffs)
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
    assert all(int(f) == 1 << 13 for f in flags[:, 0])


def test_robust_n128_t42(ctx, orc):
    """config (4) shape at reduced batch: n=128, t=42, e ~ U{0..42} errors, plus the adversarial all-errors-in-prefix case."""
    n, t, d, B = 128, 42, 42, 24
    rng = np.random.default_rng(128)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED0004)
    nerr = rng.integers(0, t + 1, size=B)
    nerr[0], nerr[1] = 0, t
    bad = _corrupt(shares, rng, nerr)
    bad[2] = _corrupt(shares[2:3], rng, [t], positions=d + t + 1)[0]  # exactly t errors all at ids < 85
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
    assert rc == want["rc"] == 0
    assert np.array_equal(path, want["path"])
    assert np.array_equal(co, coeffs) and np.array_equal(co, want["coeffs"])
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])


def test_recover_validation_errors(ctx, hb, orc):
    n, t, d = 10, 3, 3
    coeffs, shares = _codewords(orc, n, d, 4, 1)
    evals = np.ascontiguousarray(shares.transpose(1, 0, 2))
    ids = np.arange(n)
    for bad_call, code in [
        (lambda: ctx.batch_recover(ids[:6], evals[:6], n, d, t), hb.INVALID_INPUT),          # S < d+t+1
        (lambda: ctx.batch_recover(ids, evals, 9, d, t), hb.INVALID_INPUT),                  # n < 3t+1
        (lambda: ctx.batch_recover(np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 8]), evals, n, d, t), hb.INVALID_INPUT),  # duplicate id
        (lambda: ctx.batch_recover(np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 10]), evals, n, d, t), hb.INVALID_INPUT),  # id >= n
    ]:
        with pytest.raises(hb.HbmpcError) as e:
            bad_call()
        assert e.value.code == code
    # exactly S = d+t+1 with one error: optimistic fails and OEC cannot start -> DecodingError
    S = d + t + 1
    ev = evals[:S].copy()
    ev[1, 2, 0] += np.uint64(9)
    want = orc.batch_recover_secret(ids[:S], ev, n, d, t)
    rc, co, path, _ = ctx.batch_recover(ids[:S], ev, n, d, t)
    assert rc == want["rc"] == hb.DECODING_ERROR
    assert np.array_equal(path, want["path"]) and path[2] == -hb.DECODING_ERROR
    assert np.array_equal(co, want["coeffs"])


def test_device_pointers_and_async(ctx, orc, hb):
    torch = pytest.importorskip("torch")
    n, d, B = 64, 21, 4096
    coeffs = _rand(orc, (B, d + 1), 12)
    rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    dc = torch.from_numpy(coeffs.view(np.int64)).cuda()
    out = ctx.compute_shares_batch(dc, n)
    assert out.is_cuda
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    ctx.set_async(True)
    try:
        out2 = torch.empty_like(out)
        ctx.compute_shares_batch(dc, n, out=out2)
        assert ctx.synchronize() == 0
        assert torch.equal(out, out2)
        ev = out.permute(1, 0, 2).contiguous()
        rc, co, path, _ = ctx.batch_recover(np.arange(n), ev, n, d, 21)
        assert ctx.synchronize() == 0
        assert np.array_equal(co.cpu().numpy().view(np.uint64), coeffs)
    finally:
        ctx.set_async(False)


def test_imad_probe_runs(ctx):
    for v in (0, 1, 2):
        g, ms = ctx.measure_imad_peak(v)
        assert g > 100.0 and ms > 0
    assert ctx.launch_count > 0


def test_pipelined_host_chunks(hb, orc, monkeypatch):
    """Host buffers larger than one chunk go through the multi-lane copy/compute pipeline (1D and 2D chunk copies):
    results must equal the single-pass device path and the oracle."""
    monkeypatch.setenv("HBMPC_CHUNK_MB", "1")
    c = hb.Context(0)
    try:
        n, t, d, B = 64, 21, 21, 9000
        rng = np.random.default_rng(9)
        coeffs = _rand(orc, (B, d + 1), 0xC0FFEE)
        rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
        shares = c.compute_shares_batch(coeffs, n)
        assert np.array_equal(shares, want)
        rm = c.apply_vandermonde_batch(coeffs, n, recipient_major=True)
        assert np.array_equal(rm, np.ascontiguousarray(want.transpose(1, 0, 2)))
        nerr = rng.integers(0, 3, size=B)
        nerr[::7] = 0
        bad = _corrupt(shares, rng, nerr)
        evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
        ids = np.arange(n)
        ref = orc.batch_recover_secret(ids, evals, n, d, t, threads=orc.max_threads())
        got = c.batch_recover(ids, evals, n, d, t, want_flags=True)
        _compare_recover(got, ref, B)
        assert np.array_equal(got[1], coeffs)
        rc, co, sec, path, flags = c.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
        assert np.array_equal(co, coeffs) and np.array_equal(sec, coeffs[:, 0]) and np.array_equal(path, ref["path"])
        # lean host path (no flags): only the d+t+1 examined sender vectors are uploaded first; chunks with failing items
        # are re-run with all senders.  Permuted arrival order makes the uploaded rows non-contiguous.
        perm = rng.permutation(n)
        ev_p = np.ascontiguousarray(evals[perm])
        ref_p = orc.batch_recover_secret(ids[perm], ev_p, n, d, t, threads=orc.max_threads())
        got_p = c.batch_recover(ids[perm], ev_p, n, d, t, want_flags=False)
        _compare_recover(got_p, ref_p, B)
        clean = np.ascontiguousarray(shares.transpose(1, 0, 2))
        clean[50:, 4000:4100, 0] ^= np.uint64(3)          # errors only beyond the examined prefix, only in one chunk
        got_c = c.batch_recover(ids, clean, n, d, t)
        assert got_c[0] == 0 and np.array_equal(got_c[1], coeffs) and not got_c[2].any()
        rc, secrets, path = c.batch_recover_secrets(ids, evals, n, d, t)
        assert rc == 0 and np.array_equal(secrets, coeffs[:, 0]) and np.array_equal(path, ref["path"])
        a, b = _rand(orc, (B,), 1), _rand(orc, (B,), 2)
        rc, w = orc.elementwise(2, a, b)
        assert np.array_equal(c.elementwise(2, a, b), w)
    finally:
        c.close()


# ------------------------------------------------------------------ a10: NonRobustShare::recover_secret (RanDouSha checker)
@pytest.mark.parametrize("n,t", [(4, 1), (7, 2), (10, 3), (16, 5), (64, 21)])
def test_nonrobust_recover_batch(ctx, orc, hb, n, t):
    B = 40
    rng = np.random.default_rng(n)
    for deg in (t, 2 * t):
        coeffs, shares = _codewords(orc, n, deg, B, 0x5EED0500 + n + deg)
        # item 1: top coefficient zero (degree deg-1); item 2: zero polynomial; item 3..5: not a codeword (degree mismatch)
        coeffs[1, deg] = 0
        coeffs[2] = 0
        rc, shares = orc.compute_shares(coeffs, n)
        for S, ids in ((n, np.arange(n)), (max(deg + 1, n - 2), rng.permutation(n)[: max(deg + 1, n - 2)])):
            vals = shares[:, ids].copy()
            for b in (3, 4, 5):
                vals[b, rng.integers(0, S), 0] ^= np.uint64(77)
            co, sec, st = ctx.nonrobust_recover_batch(ids, vals, n, deg)
            co2, sec2, st2 = ctx.nonrobust_recover_batch(ids, np.ascontiguousarray(vals.transpose(1, 0, 2)), n, deg, sender_major=True)
            assert np.array_equal(co, co2) and np.array_equal(sec, sec2) and np.array_equal(st, st2)
            for b in range(B):
                ref = orc.nonrobust_recover_secret(ids, vals[b], n, deg)
                if ref["rc"] == 0:
                    assert st[b] == max(ref["coeff_len"] - 1, 0), (b, st[b], ref["coeff_len"])
                    assert np.array_equal(co[b], ref["coeffs"]) and np.array_equal(sec[b], ref["secret"])
                else:
                    assert ref["rc"] == hb.DEGREE_MISMATCH and st[b] == -hb.DEGREE_MISMATCH
                    assert not co[b].any() and not sec[b].any()
            if S > deg + 1:
                assert (st[3:6] == -hb.DEGREE_MISMATCH).all()
            assert st[0] == deg and st[1] == deg - 1 and st[2] == 0
    with pytest.raises(hb.HbmpcError) as e:
        ctx.nonrobust_recover_batch(np.arange(t), shares[:, :t], n, t)
    assert e.value.code == hb.INSUFFICIENT_SHARES
    with pytest.raises(hb.HbmpcError) as e:
        ctx.nonrobust_recover_batch(np.zeros(t + 1, dtype=np.uint64), shares[:, : t + 1], n, t)
    assert e.value.code == hb.INVALID_INPUT


def test_nonrobust_table_cache_is_keyed_by_the_id_set(hb, orc):
    """RanDouSha checkers see a new arrival order every session: the a10 tables are keyed by the sorted id set (per-call order
    maps), so a thousand random arrival orders of the same senders must leave free device memory flat -- and stay exact."""
    import torch

    n, deg, S, B = 16, 5, 12, 8
    c = hb.Context(0)
    rng = np.random.default_rng(10)
    coeffs, shares = _codewords(orc, n, deg, B, 0x5EED0510)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    first = c.nonrobust_recover_batch(ids, shares[:, ids], n, deg)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for it in range(1000):
        arr = rng.permutation(S)
        co, sec, st = c.nonrobust_recover_batch(ids[arr], shares[:, ids[arr]], n, deg)
        assert np.array_equal(co, first[0]) and np.array_equal(sec, first[1]) and (st == deg).all()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (1 << 20), "table cache grows with the arrival order"
    # more id sets than the cache bound: entries are evicted and their device memory returned
    for it in range(300):
        sub = np.sort(rng.choice(n, size=S, replace=False))
        c.nonrobust_recover_batch(sub, shares[:, sub], n, deg)
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)
    c.close()


# ------------------------------------------------------------------ edges
def test_vandermonde_more_columns_than_domain(ctx, orc):
    """cols > N (the exponent wraps, w^N = 1): the NTT does not apply, the dense kernel must take over."""
    n, cols, B = 4, 7, 50
    x = _rand(orc, (B, cols), 808)
    rc, want = orc.apply_vandermonde(x, n)
    assert rc == 0 and np.array_equal(ctx.apply_vandermonde_batch(x, n), want)


def test_forced_dense_equals_ntt(hb, orc, monkeypatch):
    """K1/K2 through the dense matvec kernel (HBMPC_FORCE_DENSE) and K3 without the inverse-NTT fast path
    (HBMPC_NO_FASTPATH) give the same bytes as the default NTT paths."""
    n, t, d, B = 64, 21, 21, 300
    coeffs = _rand(orc, (B, d + 1), 909)
    monkeypatch.setenv("HBMPC_FORCE_DENSE", "1")
    monkeypatch.setenv("HBMPC_NO_FASTPATH", "1")
    c1 = hb.Context(0)
    monkeypatch.delenv("HBMPC_FORCE_DENSE")
    monkeypatch.delenv("HBMPC_NO_FASTPATH")
    c2 = hb.Context(0)
    try:
        s1, s2 = c1.compute_shares_batch(coeffs, n), c2.compute_shares_batch(coeffs, n)
        assert np.array_equal(s1, s2)
        bad = s1.copy()
        bad[::5, 60, 0] ^= np.uint64(1)      # beyond the examined prefix: fast path rejects, dense check accepts (path 0)
        bad[1::5, 3, 0] ^= np.uint64(1)      # inside: robust path
        ev = np.ascontiguousarray(bad.transpose(1, 0, 2))
        r1 = c1.batch_recover(np.arange(n), ev, n, d, t, want_flags=True)
        r2 = c2.batch_recover(np.arange(n), ev, n, d, t, want_flags=True)
        for a_, b_ in zip(r1, r2):
            assert np.array_equal(a_, b_)
        assert np.array_equal(r2[1], coeffs) and (r2[2][::5] == 0).all() and (r2[2][1::5] == 1).all()
    finally:
        c1.close()
        c2.close()


def test_large_party_count_recover(ctx, orc):
    n, t, d, B = 255, 84, 84, 12
    coeffs, shares = _codewords(orc, n, d, B, 777)
    rng = np.random.default_rng(3)
    nerr = np.array([0, 1, 5, 40, 84, 84, 0, 2, 3, 10, 20, 60])
    bad = _corrupt(shares, rng, nerr)
    ids = np.arange(n)
    rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
    assert rc == 0 and np.array_equal(co, coeffs)
    for b in range(B):
        diff = np.nonzero((bad[b] != shares[b]).any(axis=1))[0]
        got = [i for i in range(n) if (int(flags[b, i >> 6]) >> (i & 63)) & 1]
        assert got == diff.tolist()
        # reference round: smallest r with #errors among the lowest d+t+1+r ids <= r (0 if the lowest d+t+1 are clean)
        pre = d + t + 1
        want = 0 if not (diff < pre).any() else next(r for r in range(1, t + 1) if (diff < pre + r).sum() <= r)
        assert path[b] == want


def test_two_contexts_from_two_threads(hb, orc):
    """The reference calls the share functions from several tokio worker threads (one node per task, many nodes per
    process in its tests): one context per thread must work concurrently on the same GPU."""
    import threading

    n, t, d, B = 16, 5, 5, 20000
    results, errors = {}, []

    def worker(k):
        try:
            c = hb.Context(0)
            coeffs = _rand(orc, (B, d + 1), 0xAB00 + k)
            for _ in range(3):
                shares = c.compute_shares_batch(coeffs, n)
                ev = np.ascontiguousarray(shares.transpose(1, 0, 2))
                rc, co, path, _ = c.batch_recover(np.arange(n), ev, n, d, t)
                assert rc == 0 and np.array_equal(co, coeffs) and not path.any()
            results[k] = shares
            c.close()
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors
    for k in range(3):
        rc, want = orc.compute_shares(_rand(orc, (B, d + 1), 0xAB00 + k), n, threads=orc.max_threads())
        assert np.array_equal(results[k], want)


def test_serialized_payload_pointer_offsets(ctx, orc):
    """ark-serialize writes Vec<F> as a u64 length followed by 32-byte little-endian canonical elements
    (batch_recon.rs:174-175): the element bytes are this library's format, only 8-byte aligned.  Host pointers at such
    offsets must be accepted as they are (no unpacking pass)."""
    n, t, d, B = 16, 5, 5, 1000
    coeffs = _rand(orc, (B, d + 1), 0xF00D)
    raw = np.zeros(8 + coeffs.nbytes, dtype=np.uint8)
    raw[:8] = np.frombuffer(np.uint64(B * (d + 1)).tobytes(), dtype=np.uint8)   # the length prefix
    raw[8:] = coeffs.view(np.uint8).reshape(-1)
    payload = raw[8:].view(np.uint64).reshape(B, d + 1, 4)                       # 8-byte aligned view, not 16
    assert payload.ctypes.data % 16 == 8 or payload.ctypes.data % 16 == 0
    rc, want = orc.compute_shares(coeffs, n)
    out_raw = np.zeros(8 + B * n * 32, dtype=np.uint8)
    out = out_raw[8:].view(np.uint64).reshape(B, n, 4)
    got = ctx.compute_shares_batch(payload, n, out=out)
    assert np.array_equal(got, want)


def test_randomized_configurations(ctx, orc):
    """Seeded sweep over random (n, t, degree, sender subset, arrival order, error pattern, flags on/off): every output of
    K1/K2/K3/K4/a10 must equal the oracle's, byte for byte."""
    rng = np.random.default_rng(20261018)
    for trial in range(40):
        n = int(rng.integers(4, 41))
        t = int(rng.integers(1, (n - 1) // 3 + 1))
        d = int(rng.choice([t, 2 * t])) if 2 * t + t + 1 <= n else t
        needed = d + t + 1
        S = int(rng.integers(needed, n + 1))
        B = int(rng.integers(1, 70))
        coeffs, shares = _codewords(orc, n, d, B, 0x5EED1000 + trial)
        cols = int(rng.integers(1, n + 1))
        x = _rand(orc, (B, cols), 0x5EED2000 + trial)
        rm = bool(rng.integers(0, 2))
        rc, want = orc.apply_vandermonde(x, n, rm)
        assert np.array_equal(ctx.apply_vandermonde_batch(x, n, rm), want), (trial, n, cols)
        ids = rng.permutation(n)[:S]
        words = shares[:, ids]
        nerr = rng.integers(0, t + 2, size=B)
        nerr[rng.integers(0, B)] = 0
        bad = _corrupt(words, rng, np.minimum(nerr, S))
        flags_on = bool(rng.integers(0, 2))
        want = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
        rc, co, sec, path, fl = ctx.robust_interpolate_batch(ids, bad, n, d, t, want_flags=flags_on)
        key = (trial, n, t, d, S, B, flags_on)
        assert rc == want["rc"], key
        assert np.array_equal(path, want["path"]), key
        assert np.array_equal(co, want["coeffs"]) and np.array_equal(sec, want["secrets"]), key
        if flags_on:
            assert np.array_equal(fl, want["flags"][:, : fl.shape[1]]), key
        ev = np.ascontiguousarray(bad.transpose(1, 0, 2))
        wb = orc.batch_recover_secret(ids, ev, n, d, t, threads=orc.max_threads())
        gb = ctx.batch_recover(ids, ev, n, d, t, want_flags=flags_on)
        _compare_recover(gb, wb, B)
        # a10 on the clean word: degree d interpolant through all S points
        co10, sec10, st10 = ctx.nonrobust_recover_batch(ids, words, n, d)
        assert np.array_equal(co10, coeffs) and np.array_equal(sec10, coeffs[:, 0]), key


def test_table_cache_eviction(hb, orc):
    """More distinct sender sets than the table cache holds (256): entries are evicted and rebuilt, results stay exact."""
    n, t, d, B = 16, 5, 5, 8
    c = hb.Context(0)
    try:
        coeffs, shares = _codewords(orc, n, d, B, 4711)
        rng = np.random.default_rng(4711)
        seen = set()
        for k in range(500):
            S = int(rng.integers(d + t + 1, n - 1))
            ids = rng.permutation(n)[:S]
            seen.add(tuple(sorted(ids.tolist())))
            ev = np.ascontiguousarray(shares[:, ids].transpose(1, 0, 2))
            if k % 3 == 0:
                ev = ev.copy()
                ev[0, 1, 0] ^= np.uint64(5)   # one error in item 1 -> robust path when S allows it
            want = orc.batch_recover_secret(ids, ev, n, d, t)
            got = c.batch_recover(ids, ev, n, d, t)
            assert got[0] == want["rc"] and np.array_equal(got[2], want["path"]) and np.array_equal(got[1], want["coeffs"]), k
        assert len(seen) > 256
    finally:
        c.close()


def test_share_record_pack_unpack(ctx, orc):
    """N1: 48-byte ark-serialize ShamirShare records (value | u64 id | u64 degree) <-> value arrays, including an
    8-byte-aligned payload offset and more records than one pipeline chunk."""
    n, d, B = 16, 5, 3
    coeffs = _rand(orc, (B, d + 1), 0xBEEF)
    shares = ctx.compute_shares_batch(coeffs, n)                      # [B][n]
    for b in range(B):
        rec = ctx.pack_share_records(shares[b], per_id=1, degree=d)  # the n shares of one sharing: ids 0..n-1
        raw = rec.reshape(n, 48)
        assert np.array_equal(raw[:, :32].copy().view(np.uint64).reshape(n, 4), shares[b])
        assert raw[:, 32:40].copy().view(np.uint64).reshape(-1).tolist() == list(range(n))
        assert (raw[:, 40:48].copy().view(np.uint64).reshape(-1) == d).all()
        payload = np.zeros(8 + rec.size, dtype=np.uint8)             # Vec length prefix + records
        payload[:8] = np.frombuffer(np.uint64(n).tobytes(), dtype=np.uint8)
        payload[8:] = rec
        vals, ids, degs = ctx.unpack_share_records(payload[8:], n)
        assert np.array_equal(vals, shares[b]) and ids.tolist() == list(range(n)) and (degs == d).all()
    big = _rand(orc, (400000,), 5)                                   # > one 16 MB chunk of 48-byte records
    rec = ctx.pack_share_records(big, per_id=100000, degree=7)
    vals, ids, degs = ctx.unpack_share_records(rec, 400000)
    assert np.array_equal(vals, big) and np.array_equal(ids, np.arange(400000, dtype=np.uint64) // 100000) and (degs == 7).all()


@pytest.mark.parametrize("n,t,d", [(1, 0, 0), (2, 0, 0), (2, 0, 1), (3, 0, 2), (4, 1, 1), (4, 1, 2), (5, 1, 3), (255, 84, 84), (255, 84, 168), (200, 66, 132)])
def test_tiny_and_maximal_party_counts(ctx, orc, n, t, d):
    B = 9
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED4000 + n + d)
    assert np.array_equal(ctx.compute_shares_batch(coeffs, n), shares)
    ids = np.arange(n)
    ev = np.ascontiguousarray(shares.transpose(1, 0, 2))
    want = orc.batch_recover_secret(ids, ev, n, d, t, threads=orc.max_threads())
    for fl in (False, True):
        got = ctx.batch_recover(ids, ev, n, d, t, want_flags=fl)
        _compare_recover(got, want, B)
    assert np.array_equal(want["coeffs"], coeffs)
    co, sec, st = ctx.nonrobust_recover_batch(ids, shares, n, d)
    assert np.array_equal(co, coeffs) and np.array_equal(sec, coeffs[:, 0])


def test_gpu_against_independent_python_model(ctx, hb):
    """The CUDA path against the second, independent restatement (oracle/pymodel.py, plain Python big ints) -- not via the C oracle."""
    from oracle import pymodel as pm

    rng = pm.SplitMix64(0xA11CE)
    for n, t, d, errs in [(7, 2, 2, [3]), (10, 3, 3, [0, 9]), (10, 3, 6, []), (13, 4, 4, [1, 2, 3, 12]), (16, 5, 5, [15])]:
        coeffs = [rng.fr() for _ in range(d + 1)]
        shares = pm.compute_shares(coeffs, n, d)
        got = hb.from_limbs(ctx.compute_shares_batch(hb.to_limbs([coeffs]), n))[0]
        assert got == shares
        vals = list(shares)
        for e in errs:
            vals[e] = (vals[e] + 12345) % pm.R_MOD
        ids = list(range(n))
        ref = pm.robust_recover_secret([(i, v, d) for i, v in zip(ids, vals)], n, t)
        rc, co, sec, path, flags = ctx.robust_interpolate_batch(ids, hb.to_limbs([vals]), n, d, t, want_flags=True)
        assert rc == 0 and path[0] == ref["path"] and hb.from_limbs(sec)[0] == ref["secret"]
        assert hb.from_limbs(co)[0][: len(ref["coeffs"])] == ref["coeffs"]
        assert [(int(flags[0, i >> 6]) >> (i & 63)) & 1 for i in range(n)] == [int(f) for f in ref["flags"]]
        out = pm.batch_recover_secret([(i, [v]) for i, v in zip(ids, vals)], n, d, t)
        rc, co2, path2, _ = ctx.batch_recover(ids, hb.to_limbs([[v] for v in vals]), n, d, t)
        want = out[0]["coeffs"] + [0] * (d + 1 - len(out[0]["coeffs"]))
        assert rc == 0 and path2[0] == out[0]["path"] and hb.from_limbs(co2)[0] == want


@pytest.mark.parametrize("n,t,S", [(16, 5, 16), (16, 5, 14), (13, 4, 13), (64, 21, 64), (64, 21, 60)])
def test_persistent_attacker_speculation_is_exact(hb, orc, monkeypatch, n, t, S):
    """Many failing chunks with the same corrupted senders trigger the speculative path (scouts -> suspected senders ->
    interpolate from the others -> verify against all supplied shares).  Mixed in: chunks whose errors sit elsewhere
    (speculation must hand them to the full decoder), clean chunks, chunks where a suspected sender happens to be right,
    and chunks with more than t errors.  Everything must equal the oracle and the run with speculation switched off."""
    d = t
    B = 3000 if n <= 16 else 1500
    rng = np.random.default_rng(n * 7 + S)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED6000 + n + S)
    ids = np.sort(rng.permutation(n)[:S])
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]].copy()                       # [B][S] in arrival order
    x = S - (d + t + 1)                                          # extra shares: OEC can absorb at most min(t, x) errors
    nbad = max(1, min(t, x))
    bad_senders = rng.permutation(S)[:nbad]
    for b in range(B):
        kind = b % 10
        if kind == 0:
            continue                                             # clean chunk
        if kind == 1:                                            # errors somewhere else
            for p in rng.choice(S, size=int(rng.integers(1, nbad + 1)), replace=False):
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
        elif kind == 2 and nbad > 1:                             # one suspected sender happens to be right
            for p in bad_senders[1:]:
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
        elif kind == 3:                                          # more than the decodable number of errors
            for p in rng.choice(S, size=min(S, t + 1), replace=False):
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
        else:                                                    # the persistent attackers
            for p in bad_senders:
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
    evals = np.ascontiguousarray(words.transpose(1, 0, 2))
    want = orc.batch_recover_secret(ids[arrival], evals, n, d, t, threads=orc.max_threads())
    monkeypatch.setenv("HBMPC_SCAN_MAX", "0")          # list mode also for this test-sized batch, so the shortcut can trigger
    c = hb.Context(0)
    monkeypatch.setenv("HBMPC_NO_SPECULATION", "1")
    c_off = hb.Context(0)
    try:
        launches = []
        for cc in (c, c_off):
            l0 = cc.launch_count
            for fl in (True, False):
                got = cc.batch_recover(ids[arrival], evals, n, d, t, want_flags=fl)
                _compare_recover(got, want, B)
            rc, co, sec, path, flags = cc.robust_interpolate_batch(ids[arrival], words, n, d, t, want_flags=True)
            assert rc == want["rc"] and np.array_equal(path, want["path"]) and np.array_equal(co, want["coeffs"])
            assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
            launches.append(cc.launch_count - l0)
        assert launches[0] > launches[1], "the speculative path did not run"
    finally:
        c.close()
        c_off.close()


@pytest.mark.parametrize("n,t,S", [(16, 5, 13), (64, 21, 50), (100, 33, 100)])
def test_flags_on_sender_subsets_both_routes(hb, orc, monkeypatch, n, t, S):
    """Calls with flags on a sender subset: by default the erasure-weighted transform over ALL supplied senders settles the chunks in
    which every share agrees (path 0, no flag) and only the others take the dense check; HBMPC_NO_ER_FLAGS=1 sends every chunk to
    the dense check.  Both must equal the oracle: clean chunks, errors inside and beyond the examined prefix, > t errors."""
    d = t
    B = 600
    rng = np.random.default_rng(n + S)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED9000 + n + S)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]]
    nerr = np.where(rng.random(B) < 0.5, 0, rng.integers(1, t + 3, size=B))
    bad = _corrupt(words, rng, np.minimum(nerr, S))
    evals = np.ascontiguousarray(bad.transpose(1, 0, 2))
    want = orc.batch_recover_secret(ids[arrival], evals, n, d, t, threads=orc.max_threads())
    c0 = hb.Context(0)
    monkeypatch.setenv("HBMPC_NO_ER_FLAGS", "1")
    c1 = hb.Context(0)
    try:
        l0, l1 = c0.launch_count, c1.launch_count
        _compare_recover(c0.batch_recover(ids[arrival], evals, n, d, t, want_flags=True), want, B)
        _compare_recover(c1.batch_recover(ids[arrival], evals, n, d, t, want_flags=True), want, B)
        assert c0.launch_count - l0 > c1.launch_count - l1, "the erasure-weighted check did not run ahead of the dense check"
        w2 = orc.robust_interpolate_batch(ids[arrival], bad, n, d, t, threads=orc.max_threads())
        for c in (c0, c1):
            rc, co, sec, path, flags = c.robust_interpolate_batch(ids[arrival], bad, n, d, t, want_flags=True)
            assert rc == w2["rc"] and np.array_equal(path, w2["path"]) and np.array_equal(co, w2["coeffs"])
            assert np.array_equal(flags, w2["flags"][:, : flags.shape[1]])
    finally:
        c0.close()
        c1.close()
