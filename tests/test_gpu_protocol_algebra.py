"""Protocol-level parity rows of SURVEY.md 4a, computed with the CUDA kernels (through the C ABI) and checked against the
reference tests' known answers and the oracle: the FSMs stay on the host, only their field algebra is exercised."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batchrecon_fig2_pipeline_n4_t1(ctx, hb, orc):
    """tests/batchrecon_test.rs:28-119: secrets [3,4], n=4, t=1.  K2 (encode) -> K3 (round 1: open y_j from the first 2t+1
    arrivals) -> K3 (round 2: recover the secrets)."""
    n, t, R = 4, 1, hb.R_MOD
    polys = hb.to_limbs([[3, 11], [4, 29]])
    shares = ctx.compute_shares_batch(polys, n)                      # shares[k][i]: party i's share of secret k
    per_party = np.ascontiguousarray(shares.transpose(1, 0, 2))      # [party i][k]
    y = ctx.apply_vandermonde_batch(per_party, n, recipient_major=True)  # y[j][i]: what party i sends to party j
    yk = hb.from_limbs(ctx.apply_vandermonde_batch(hb.to_limbs([[3, 4]]), n))[0]
    assert yk[0] == 7 and yk[2] == R - 1
    assert yk[1] == 0x0235473339D80C1343B00C0009D80C00000004000000000003
    assert yk[3] == 0x73EDA753299D7D45FDF2A4CE3195C4C1A3B1A3F927F25BFEFFFBFFFF00000004
    opened = []
    for j in range(n):
        senders = [(j + a) % n for a in range(2 * t + 1)]            # first 2t+1 arrivals at party j
        ev = np.ascontiguousarray(y[j, senders][:, None, :])         # [S][1 chunk]
        rc, sec, path = ctx.batch_recover_secrets(senders, ev, n, t, t)
        assert rc == 0 and path[0] == 0
        opened.append(hb.from_limbs(sec)[0])
    assert opened == yk
    ev2 = hb.to_limbs([[opened[0]], [opened[1]], [opened[2]]])
    rc, co, path, _ = ctx.batch_recover([0, 1, 2], ev2, n, t, t)
    assert rc == 0 and hb.from_limbs(co)[0] == [3, 4]


def test_triple_gen_algebra_n15_t3(ctx, hb, orc):
    """tests/triple_gen_test.rs:77 / triple_generation.rs:332-340,196-208: open a*b - r_2t with a degree-2t batch
    reconstruction, then c = r_t + opened must be a degree-t sharing of a*b."""
    n, t, R = 15, 3, hb.R_MOD
    G = 2 * t + 1                                                    # one group of 2t+1 triples
    a, b, r = (orc.random_fr((G,), s) for s in (1, 2, 3))
    rnd = lambda d, s: orc.random_fr((G, d), s)
    sh = lambda sec, d, s: ctx.compute_shares_batch(np.concatenate([sec[:, None, :], rnd(d, s)], axis=1), n)
    a_sh, b_sh, rt_sh, r2t_sh = sh(a, t, 10), sh(b, t, 11), sh(r, t, 12), sh(r, 2 * t, 13)   # [G][n]
    prod = ctx.elementwise(2, a_sh, b_sh)                            # share_mul: degree 2t
    masked = ctx.elementwise(1, prod, r2t_sh)                        # a*b - r_2t   (per party i: masked[:, i])
    per_party = np.ascontiguousarray(masked.transpose(1, 0, 2))      # [party][G] = one chunk of 2t+1 values per party
    y = ctx.apply_vandermonde_batch(per_party, n, recipient_major=True)      # y[j][i]
    # round 1: party j opens y_j (degree 2t: needs 3t+1 senders); round 2: the 2t+1 opened values
    senders = list(range(3 * t + 1))
    yj = []
    for j in range(n):
        rc, sec, path = ctx.batch_recover_secrets(senders, np.ascontiguousarray(y[j, senders][:, None, :]), n, 2 * t, t)
        assert rc == 0
        yj.append(sec[0])
    ev2 = np.ascontiguousarray(np.stack(yj)[senders][:, None, :])
    rc, co, path, _ = ctx.batch_recover(senders, ev2, n, 2 * t, t)
    assert rc == 0
    opened = hb.from_limbs(co)[0]
    av, bv, rv = (hb.from_limbs(x) for x in (a, b, r))
    assert opened == [(x * y_ - z) % R for x, y_, z in zip(av, bv, rv)]
    # c = r_t + opened: add the public value to every share, reconstruct -> a*b, degree t
    opened_l = np.repeat(co[0][:, None, :], n, axis=1)               # [G][n]
    c_sh = ctx.elementwise(0, rt_sh, opened_l)
    rc, cco, sec, path, _ = ctx.robust_interpolate_batch(np.arange(n), c_sh, n, t, t)
    assert rc == 0 and hb.from_limbs(sec) == [(x * y_) % R for x, y_ in zip(av, bv)]


def test_randousha_checks(ctx, hb, orc):
    """ran_dou_sha/mod.rs:568-602, tests/randousha_test.rs:467,518: after the n x n Vandermonde, a checker reconstructs both
    sharings from all n shares; wrong degree or r_t != r_2t must be detected."""
    n, t, B = 7, 2, 6
    r = orc.random_fr((B,), 5)
    mk = lambda sec, d, s: ctx.compute_shares_batch(np.concatenate([sec[:, None, :], orc.random_fr((B, d), s)], axis=1), n)
    st, s2t = mk(r, t, 20), mk(r, 2 * t, 21)
    bad_secret = mk(orc.random_fr((B,), 6), 2 * t, 22)               # r_2t of a different secret
    wrong_deg = mk(r, t + 1, 23)                                     # "degree t" sharing that is really degree t+1
    ids = np.arange(n)
    co_t, sec_t, deg_t = ctx.nonrobust_recover_batch(ids, st, n, t)
    co_2t, sec_2t, deg_2t = ctx.nonrobust_recover_batch(ids, s2t, n, 2 * t)
    ok = (deg_t == t) & (deg_2t == 2 * t) & (sec_t == sec_2t).all(axis=1)
    assert ok.all() and np.array_equal(sec_t, r)
    _, sec_bad, deg_bad = ctx.nonrobust_recover_batch(ids, bad_secret, n, 2 * t)
    assert not ((deg_bad == 2 * t) & (sec_t == sec_bad).all(axis=1)).any()
    _, _, deg_w = ctx.nonrobust_recover_batch(ids, wrong_deg, n, t)
    assert (deg_w == -hb.DEGREE_MISMATCH).all()
    # hyperinvertible step itself: n x n Vandermonde of the dealt shares (ran_dou_sha/mod.rs:392-403) == oracle
    cols = orc.random_fr((B, n), 30)
    rc, want = orc.apply_vandermonde(cols, n)
    assert np.array_equal(ctx.apply_vandermonde_batch(cols, n), want)


def test_beaver_mul_known_answers(ctx, hb, orc):
    """tests/node_test.rs:584-764 (config 1, 4 parties, t=1): 10*10 = 100 and 20*20 = 400 through Beaver's algebra
    (multiplication.rs:79-97,417-426) on the GPU kernels."""
    n, t, R = 4, 1, hb.R_MOD
    x = hb.to_limbs([10, 20])
    a, b = orc.random_fr((2,), 41), orc.random_fr((2,), 42)
    c = ctx.elementwise(2, a, b)
    mk = lambda sec, s: ctx.compute_shares_batch(np.concatenate([sec[:, None, :], orc.random_fr((2, t), s)], axis=1), n)
    xs, ys, as_, bs, cs = mk(x, 1), mk(x, 2), mk(a, 3), mk(b, 4), mk(c, 5)
    d_sh, e_sh = ctx.elementwise(1, xs, as_), ctx.elementwise(1, ys, bs)      # x - a, y - b
    ids = np.arange(n)
    _, _, d, _, _ = ctx.robust_interpolate_batch(ids, d_sh, n, t, t)
    _, _, e, _, _ = ctx.robust_interpolate_batch(ids, e_sh, n, t, t)
    rep = lambda v: np.repeat(v[:, None, :], n, axis=1)
    z = ctx.elementwise(0, cs, ctx.elementwise(2, rep(d), bs))                # c + d*b
    z = ctx.elementwise(0, z, ctx.elementwise(2, rep(e), as_))                # + e*a
    z = ctx.elementwise(0, z, rep(ctx.elementwise(2, d, e)))                  # + d*e
    rc, _, prod, path, _ = ctx.robust_interpolate_batch(ids, z, n, t, t)
    assert rc == 0 and hb.from_limbs(prod) == [100, 400]


def _edge_values(hb, orc, count, seed):
    """random canonical values with the edge cases of the field mixed in (0, 1, r-1, r-2, 2^255-ish top limb patterns)"""
    R = hb.R_MOD
    v = orc.random_fr((count,), seed)
    edges = hb.to_limbs([0, 1, R - 1, R - 2, (R - 1) // 2, (R + 1) // 2, 1 << 32, (1 << 224) + 5])
    k = min(len(edges), count)
    pos = np.random.default_rng(seed).permutation(count)[:k]
    v[pos] = edges[:k]
    return v


@pytest.mark.parametrize("count", [1, 7, 257, 70001])
def test_fused_share_algebra_matches_operator_route(ctx, hb, orc, count):
    """hbmpc_share_algebra_fused against the oracle's operator-by-operator route, which is how the reference computes
    these values: triple_generation.rs:332-340 (share_mul, then Sub), multiplication.rs:417-426 (two Subs), multiplication.rs:79-97
    (three Muls, three Subs, in the reference's order)."""
    a, b, r2t, x, y = (_edge_values(hb, orc, count, 100 + s) for s in range(5))
    ew = lambda op, u, v: orc.elementwise(op, u, v)[1]
    # triple mask
    want = ew(1, ew(2, a, b), r2t)
    got = ctx.share_algebra_fused(ctx.K5_TRIPLE_MASK, (a, b, r2t))
    assert np.array_equal(got, want)
    assert np.array_equal(got, ctx.elementwise(1, ctx.elementwise(2, a, b), r2t))
    # Beaver mask: (a - x, b - y)
    ax, by = ctx.share_algebra_fused(ctx.K5_BEAVER_MASK, (a, x, b, y))
    assert np.array_equal(ax, ew(1, a, x)) and np.array_equal(by, ew(1, b, y))
    # Beaver finalise in the reference's order: ((c - da*db) - da*[y]) - db*[x]
    c, da, db = r2t, ax, by
    want = ew(1, ew(1, ew(1, c, ew(2, da, db)), ew(2, y, da)), ew(2, x, db))
    got = ctx.share_algebra_fused(ctx.K5_BEAVER_FINALIZE, (c, x, y, da, db))
    assert np.array_equal(got, want)
    R = hb.R_MOD
    if count <= 257:  # and against Python integers
        ci, xi, yi, dai, dbi = (hb.from_limbs(v) for v in (c, x, y, da, db))
        assert hb.from_limbs(got) == [(cc - p * q - p * yy - q * xx) % R for cc, xx, yy, p, q in zip(ci, xi, yi, dai, dbi)]


def test_fused_share_algebra_device_arrays_in_place_and_errors(ctx, hb, orc):
    """device-resident operands (torch tensors), an output aliasing an input, and the call's error behaviour"""
    import torch

    count = 4099
    a, b, r = (_edge_values(hb, orc, count, 200 + s) for s in range(3))
    want = orc.elementwise(1, orc.elementwise(2, a, b)[1], r)[1]
    dev = lambda v: torch.from_numpy(v.view(np.int64)).cuda()
    ad, bd, rd = dev(a), dev(b), dev(r)
    got = ctx.share_algebra_fused(0, (ad, bd, rd), out=ad)          # in place: out0 aliases in0
    assert got.is_cuda
    assert np.array_equal(got.cpu().numpy().view(np.uint64), want)
    # a non-canonical operand (>= r) is an input error, like hbmpc_elementwise
    bad = a.copy()
    bad[3] = hb.to_limbs([hb.R_MOD])[0]
    with pytest.raises(hb.HbmpcError) as e:
        ctx.share_algebra_fused(0, (bad, b, r))
    assert e.value.code == hb.INVALID_INPUT
    with pytest.raises(ValueError):
        ctx.share_algebra_fused(2, (a, b, r))                       # op 2 takes five arrays
    with pytest.raises(ValueError):
        ctx.share_algebra_fused(0, (a, b[:5], r))
    assert np.array_equal(ctx.share_algebra_fused(0, (a, b, r)), want)   # the context is usable after the errors


def test_triple_pipeline_with_fused_algebra_n7_t2(ctx, hb, orc):
    """triple_generation.rs:332-340,196-208 with the fused mask; then Beaver multiplication of two fresh sharings with the triple
    (multiplication.rs:417-426, 79-97) through the fused mask / finalise: the product share reconstructs to x*y."""
    n, t, R = 7, 2, hb.R_MOD
    G = 2 * t + 1
    a, b, r, xs_, ys_ = (orc.random_fr((G,), s) for s in (1, 2, 3, 4, 5))
    sh = lambda sec, d, s: ctx.compute_shares_batch(np.concatenate([sec[:, None, :], orc.random_fr((G, d), s)], axis=1), n)
    a_sh, b_sh, rt_sh, r2t_sh, x_sh, y_sh = sh(a, t, 10), sh(b, t, 11), sh(r, t, 12), sh(r, 2 * t, 13), sh(xs_, t, 14), sh(ys_, t, 15)
    ids = np.arange(n)
    masked = ctx.share_algebra_fused(0, (a_sh, b_sh, r2t_sh))        # degree 2t sharing of a*b - r
    rc, _, opened, _, _ = ctx.robust_interpolate_batch(ids, masked, n, 2 * t, t)
    assert rc == 0
    av, bv, rv, xv, yv = (hb.from_limbs(v) for v in (a, b, r, xs_, ys_))
    assert hb.from_limbs(opened) == [(p * q - z) % R for p, q, z in zip(av, bv, rv)]
    rep = lambda v: np.ascontiguousarray(np.repeat(v[:, None, :], n, axis=1))
    c_sh = ctx.elementwise(0, rt_sh, rep(opened))                    # [ab]_t
    ax_sh, by_sh = ctx.share_algebra_fused(1, (a_sh, x_sh, b_sh, y_sh))
    _, _, da, _, _ = ctx.robust_interpolate_batch(ids, ax_sh, n, t, t)
    _, _, db, _, _ = ctx.robust_interpolate_batch(ids, by_sh, n, t, t)
    z_sh = ctx.share_algebra_fused(2, (c_sh, x_sh, y_sh, rep(da), rep(db)))
    rc, _, prod, path, _ = ctx.robust_interpolate_batch(ids, z_sh, n, t, t)
    assert rc == 0 and (path == 0).all()
    assert hb.from_limbs(prod) == [(p * q) % R for p, q in zip(xv, yv)]
