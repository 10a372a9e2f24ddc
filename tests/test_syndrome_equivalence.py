"""CPU check of the equivalence the CUDA decoder rests on (robust.cuh:1-30): syndromes + Berlekamp-Massey + Chien + Forney with the product's
attempt structure and path rule (oracle/syndrome_model.py) returns exactly what the reference's optimistic -> OEC -> Gao route returns
(oracle/pymodel.py, robust_interpolate.rs:94-157, 456-538, 579-628): coefficients, OEC round and per-share flags, or a decoding error --
for every error count up to and beyond t, sender subsets, arrival orders and error placements (inside / outside the examined prefixes)."""
import random

import pytest

from oracle import pymodel as pm
from oracle import syndrome_model as sm

R = pm.R_MOD


def _case(rnd, n, t, d, S, nerr, where):
    coeffs = [rnd.randrange(R) for _ in range(d + 1)]
    ids = sorted(rnd.sample(range(n), S))
    vals = {i: pm.p_eval(coeffs, pm.domain_element(n, i)) for i in ids}
    if where == "prefix":       # the lowest ids: inside every examined prefix
        bad = ids[:nerr]
    elif where == "tail":       # the highest ids: outside the early prefixes
        bad = ids[S - nerr:] if nerr else []
    else:
        bad = rnd.sample(ids, nerr)
    for i in bad:
        vals[i] = (vals[i] + 1 + rnd.randrange(R - 1)) % R
    shares = [(i, vals[i], d) for i in ids]
    rnd.shuffle(shares)          # arrival order
    return shares


def _both(shares, n, t):
    try:
        want = pm.robust_recover_secret(shares, n, t)
    except pm.ShareErr as e:
        assert e.code == pm.DECODING_ERROR
        want = None
    got = sm.robust_recover_secret(shares, n, t)
    if want is None:
        assert got is None
        return None
    assert got is not None, "the syndrome route fails where the reference decodes"
    assert got["coeffs"] == pm.p_norm(want["coeffs"]) and got["path"] == want["path"] and got["flags"] == want["flags"] and got["secret"] == want["secret"]
    return got


@pytest.mark.parametrize("n,t,d", [(4, 1, 1), (7, 2, 2), (10, 3, 3), (13, 4, 4), (16, 5, 5), (10, 2, 4), (16, 3, 6)])
def test_syndrome_route_equals_oec_gao(n, t, d):
    rnd = random.Random(1000 * n + 10 * t + d)
    needed = d + t + 1
    decoded = failed = 0
    for S in range(needed, n + 1):
        for nerr in range(0, min(S, t + 2) + 1):
            for where in ("random", "prefix", "tail"):
                for _ in range(2):
                    r = _both(_case(rnd, n, t, d, S, nerr, where), n, t)
                    decoded += r is not None
                    failed += r is None
    assert decoded > 0 and (failed > 0 or n == 4)


def test_path_rule_on_chosen_error_positions():
    """n = 16, t = 5, d = 5, all 16 shares: no error among the lowest d+t+1 = 11 ids is the optimistic path (0); otherwise the round is the
    first r whose prefix of 11 + r ids holds at most r errors."""
    n, t, d = 16, 5, 5
    rnd = random.Random(77)
    coeffs = [rnd.randrange(R) for _ in range(d + 1)]
    base = [pm.p_eval(coeffs, pm.domain_element(n, i)) for i in range(n)]
    for bad, path in (((0,), 1), ((0, 1), 2), ((0, 1, 2, 3, 4), 5), ((11,), 0), ((12, 13), 0), ((0, 12), 1), ((0, 1, 12), 3), ((3, 5, 7, 9, 15), 4)):
        vals = list(base)
        for i in bad:
            vals[i] = (vals[i] + 5) % R
        shares = [(i, vals[i], d) for i in range(n)]
        got = _both(shares, n, t)
        assert got is not None and got["path"] == path and got["coeffs"] == pm.p_norm(coeffs)
        assert [i for i, f in enumerate(got["flags"]) if f] == list(bad)
    # six errors: beyond t -- both routes must refuse
    vals = list(base)
    for i in range(6):
        vals[i] = (vals[i] + 9) % R
    assert _both([(i, vals[i], d) for i in range(n)], n, t) is None


def test_error_values_without_omega():
    """the error values from Berlekamp-Massey's auxiliary polynomial (no Omega = S*Lambda mod z^L, i.e. no
    omega_kernel) are the ones Forney's formula gives -- every attempt that succeeds returns the same positions and values either way."""
    rnd = random.Random(4242)
    checked = 0
    for _ in range(300):
        n = rnd.choice([7, 10, 13, 16, 32])
        t = (n - 1) // 3
        d = rnd.choice([t, min(2 * t, n - t - 2)])
        S = rnd.randrange(d + t + 1, n + 1)
        ids = sorted(rnd.sample(range(n), S))
        xs = [pm.domain_element(n, i) for i in ids]
        coeffs = [rnd.randrange(R) for _ in range(d + 1)]
        ys = [pm.p_eval(coeffs, x) for x in xs]
        max_l = min(t, (S - d - 1) // 2)
        nerr = rnd.randrange(0, max_l + 2)
        truth = {}
        for i in rnd.sample(range(S), min(nerr, S)):
            truth[i] = 1 + rnd.randrange(R - 1)
            ys[i] = (ys[i] + truth[i]) % R
        a, b = sm.attempt(xs, ys, d, max_l), sm.attempt(xs, ys, d, max_l, values="hk")
        assert a == b
        if len(truth) <= max_l:
            assert a == sorted(truth.items())
            checked += 1
    assert checked > 100
