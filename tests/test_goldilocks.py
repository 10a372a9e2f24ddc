"""N4 tail: the share path over GoldilocksField = Fp64, p = 2^64 - 2^32 + 1 (common/math/goldilocks.rs:4-13).
CPU: field constants and the restatement's own invariants; GPU: hbmpc_gl_* through the C ABI == oracle/goldilocks.py bit for bit."""
import numpy as np
import pytest

from oracle import goldilocks as gl

P = gl.P


def rnd(shape, seed):
    rng = np.random.default_rng(seed)
    v = rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
    return (v % np.uint64(P)).astype(np.uint64)


def test_field_constants_and_domain():
    assert P == 2**64 - 2**32 + 1 == 18446744069414584321                    # goldilocks.rs:5
    assert gl.ROOT32 == 1753635133440165772                                  # 7^((p-1)/2^32): the published two-adic root of unity
    assert pow(gl.ROOT32, 1 << 32, P) == 1 and pow(gl.ROOT32, 1 << 31, P) == P - 1
    for n in (4, 5, 16, 64, 100, 128):
        N = gl.domain_size(n)
        w = gl.domain_element(n, 1)
        assert pow(w, N, P) == 1 and pow(w, N // 2, P) == P - 1 and gl.domain_element(n, 0) == 1


def test_restatement_round_trips():
    for n, t in ((4, 1), (7, 2), (16, 5)):
        d = t
        coeffs = [int(x) for x in rnd(d + 1, n)]
        sh = gl.compute_shares(coeffs, n)
        V = gl.make_vandermonde(n, d)
        assert gl.apply_vandermonde(V, coeffs) == sh and V[0] == [1] * (d + 1) and [r[0] for r in V] == [1] * n
        ids = list(range(n))[::-1]
        rc, c, path = gl.batch_recover(ids, [[sh[i]] for i in ids], n, d, t)
        assert rc == 0 and c[0] == coeffs and path == [0]
        bad = [[sh[i]] for i in ids]
        bad[-1][0] = (bad[-1][0] + 1) % P                                    # id 0: inside the examined prefix
        rc, c, path = gl.batch_recover(ids, bad, n, d, t)
        assert rc == gl.DECODING_ERROR and path == [-gl.DECODING_ERROR]
        st, c2 = gl.nonrobust_recover(list(range(n)), sh, n, d)
        assert st == max(k for k in range(d + 1) if coeffs[k]) and c2 == coeffs
        st, _ = gl.nonrobust_recover(list(range(n)), sh, n, d - 1)
        assert st == -gl.DEGREE_MISMATCH


@pytest.mark.gpu
@pytest.mark.parametrize("n,t,B", [(4, 1, 7), (7, 2, 100), (16, 5, 1000), (64, 21, 3000), (128, 42, 200)])
def test_gl_share_and_recover_match_the_restatement(hb, ctx, n, t, B):
    for d in (t, 2 * t):
        if d + t + 1 > n:
            continue
        coeffs = rnd((B, d + 1), n + d)
        shares = ctx.gl_compute_shares_batch(coeffs, n)
        for b in (0, 1, B // 2, B - 1):
            assert [int(x) for x in shares[b]] == gl.compute_shares([int(x) for x in coeffs[b]], n)
        assert np.array_equal(ctx.gl_apply_vandermonde_batch(coeffs, n), shares)
        assert np.array_equal(ctx.gl_apply_vandermonde_batch(coeffs, n, recipient_major=True), shares.T)
        rng = np.random.default_rng(B)
        S = int(rng.integers(d + t + 1, n + 1))
        ids = rng.permutation(n)[:S]
        ev = np.ascontiguousarray(shares.T[ids])
        rc, rec, secrets, path = ctx.gl_batch_recover(ids, ev, n, d, t)
        assert rc == 0 and np.array_equal(rec, coeffs) and np.array_equal(secrets, coeffs[:, 0]) and not path.any()
        # corrupt the sender with the lowest id in some chunks (always examined), and the one with the highest id when it is not examined
        lo, hi = int(np.argmin(ids)), int(np.argmax(ids))
        ev2 = ev.copy()
        ev2[lo, ::3] = (ev2[lo, ::3] + np.uint64(1)) % np.uint64(P)
        if S > d + t + 1:
            ev2[hi, 1::3] = (ev2[hi, 1::3] + np.uint64(5)) % np.uint64(P)
        rc, rec2, sec2, path2 = ctx.gl_batch_recover(ids, ev2, n, d, t)
        small = min(B, 40)
        orc_rc, want_c, want_p = gl.batch_recover([int(i) for i in ids], [[int(v) for v in row[:small]] for row in ev2], n, d, t)
        assert rc == gl.DECODING_ERROR == orc_rc
        assert [[int(v) for v in r] for r in rec2[:small]] == want_c and [int(v) for v in path2[:small]] == want_p
        assert (path2[::3] == -gl.DECODING_ERROR).all() and not path2[1::3].any() and not path2[2::3].any()
        # NonRobustShare::recover_secret through all n points, both layouts; a degree d+1 polynomial is a DegreeMismatch
        c3, s3, st3 = ctx.gl_nonrobust_recover_batch(np.arange(n), shares, n, d)
        assert np.array_equal(c3, coeffs) and np.array_equal(s3, coeffs[:, 0])
        for b in (0, B - 1):
            assert int(st3[b]) == gl.nonrobust_recover(list(range(n)), [int(x) for x in shares[b]], n, d)[0]
        if d + 2 <= n:
            c4, s4, st4 = ctx.gl_nonrobust_recover_batch(np.arange(n), np.ascontiguousarray(shares.T), n, d - 1, sender_major=True)
            topnz = coeffs[:, d] != 0
            assert (st4[topnz] == -gl.DEGREE_MISMATCH).all()


@pytest.mark.gpu
def test_gl_elementwise_and_validation(hb, ctx):
    a, b = rnd(5000, 1), rnd(5000, 2)
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for op, f in ((0, lambda x, y: (x + y) % P), (1, lambda x, y: (x - y) % P), (2, lambda x, y: x * y % P)):
        assert [int(v) for v in ctx.gl_elementwise(op, a, b)] == [f(x, y) for x, y in zip(ai, bi)]
    edge = np.array([0, 1, P - 1, P - 2, 1 << 32, (1 << 32) - 1, (1 << 63)], dtype=np.uint64)
    for op, f in ((0, lambda x, y: (x + y) % P), (1, lambda x, y: (x - y) % P), (2, lambda x, y: x * y % P)):
        got = ctx.gl_elementwise(op, np.repeat(edge, len(edge)), np.tile(edge, len(edge)))
        want = [f(int(x), int(y)) for x in edge for y in edge]
        assert [int(v) for v in got] == want
    bad = a.copy()
    bad[3] = np.uint64(P)
    with pytest.raises(hb.HbmpcError) as e:
        ctx.gl_elementwise(0, bad, b)
    assert e.value.code == hb.INVALID_INPUT
    with pytest.raises(hb.HbmpcError) as e:
        ctx.gl_compute_shares_batch(rnd((2, 5), 3), 4)      # n <= d
    assert e.value.code == hb.INVALID_INPUT
