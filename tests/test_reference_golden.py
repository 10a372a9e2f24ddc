"""Reference-pinned parity: consumes tests/golden/reference_*.json -- vectors produced BY THE REFERENCE (arkworks) with
tools/reference_golden/dump_reference_golden.rs -- and checks the C oracle (CPU) and the CUDA path (GPU, through the C ABI) against
them.  This image has no Rust toolchain, so no such file is committed yet and those tests SKIP ("parity unpinned");
the day a maintainer runs the dumper and drops the JSON into tests/golden/, they run without further changes.

So that the consumer itself is known to work, `test_consumer_on_a_schema_sample` builds a file of the same schema from the
independent Python model (oracle/pymodel.py + oracle/chacha_fr.py; NOT reference output, it pins nothing) and runs the same checks.

Schema ("cases": list of objects; field values are 0x-hex canonical integers):
  rng_draws       seed[32], draws[k]                          StdRng::from_seed(seed), k x Fr::rand
  compute_shares  n, d, coeffs[d+1], shares[n]                RobustShare::compute_shares  (robust_interpolate.rs:52-82)
  robust_recover  n, t, d, ids[S], values[S], rc, coeffs, secret   RobustShare::recover_secret  (robust_interpolate.rs:94-157)
  batch_recover   n, t, d, ids[S], evals[S][B], rc, coeffs[B][..]  batch_recover_secret  (robust_interpolate.rs:284-443)
  vandermonde     n, t, matrix[n][t+1], inputs[t+1], outputs[n]    make_/apply_vandermonde  (common/share/mod.rs:31-76)
  nonrobust       n, d, ids, values, rc, coeffs, recovered         NonRobustShare  (common/share/shamir.rs:158-239)
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import chacha_fr
from oracle import cmodel as cm
from oracle import pymodel as pm

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "reference_*.json")))
L, I = cm.to_limbs, cm.from_limbs


def ints(xs):
    return [int(x, 16) for x in xs]


def padded(got, want):
    """reference coefficient vectors are trimmed (DensePolynomial drops trailing zeros); ours are padded to d + 1"""
    return got[: len(want)] == want and not any(got[len(want):])


class OracleBackend:
    name = "oracle"

    def draws(self, seed, k):
        return chacha_fr.sample_fr(seed, 0, k)

    def compute_shares(self, coeffs, n):
        rc, sh = cm.compute_shares(L([coeffs]), n)
        assert rc == 0
        return I(sh)[0]

    def robust_recover(self, ids, vals, n, t, d):
        r = cm.robust_recover_secret(ids, L(vals), n, t, d)
        return r["rc"], I(r["coeffs"]), I(r["secret"])

    def batch_recover(self, ids, evals, n, d, t):
        r = cm.batch_recover_secret(ids, L(evals), n, d, t)
        return r["rc"], I(r["coeffs"])

    def vandermonde(self, n, t, inputs):
        rc, V = cm.make_vandermonde(n, t)
        assert rc == 0
        rc, out = cm.apply_vandermonde(L([inputs]), n)
        assert rc == 0
        return I(V), I(out)[0]

    def nonrobust(self, ids, vals, n, d):
        r = cm.nonrobust_recover_secret(ids, L(vals), n, d)
        return r["rc"], I(r["coeffs"]), I(r["secret"])


class GpuBackend:
    """the CUDA path through the C ABI (host buffers); make_vandermonde has no export of its own: row j of the matrix is the image of
    the unit vectors"""
    name = "gpu"

    def __init__(self, hb, ctx):
        self.hb, self.ctx = hb, ctx

    def draws(self, seed, k):
        return I(self.ctx.sample_fr_batch(seed, k))

    def compute_shares(self, coeffs, n):
        return I(self.ctx.compute_shares_batch(L([coeffs]), n))[0]

    def robust_recover(self, ids, vals, n, t, d):
        try:
            rc, coeffs, secrets, path, _ = self.ctx.robust_interpolate_batch(ids, L([vals]), n, d, t)
        except self.hb.HbmpcError as e:
            return e.code, None, None
        if int(path[0]) < 0:
            return pm.DECODING_ERROR, None, None
        return 0, I(coeffs)[0], I(secrets)[0]

    def batch_recover(self, ids, evals, n, d, t):
        try:
            rc, coeffs, path, _ = self.ctx.batch_recover(ids, L(evals), n, d, t)
        except self.hb.HbmpcError as e:
            return e.code, None
        return rc, I(coeffs)

    def vandermonde(self, n, t, inputs):
        unit = [[1 if c == k else 0 for c in range(t + 1)] for k in range(t + 1)]
        cols = I(self.ctx.apply_vandermonde_batch(L(unit), n))           # cols[k][j] = V[j][k]
        V = [[cols[k][j] for k in range(t + 1)] for j in range(n)]
        return V, I(self.ctx.apply_vandermonde_batch(L([inputs]), n))[0]

    def nonrobust(self, ids, vals, n, d):
        coeffs, secrets, status = self.ctx.nonrobust_recover_batch(ids, L([vals]), n, d)
        if int(status[0]) < 0:
            return 1, None, None
        return 0, I(coeffs)[0], I(secrets)[0]


def check_cases(cases, be):
    seen = {}
    for c in cases:
        k = c["kind"]
        seen[k] = seen.get(k, 0) + 1
        if k == "rng_draws":
            assert be.draws(bytes(c["seed"]), len(c["draws"])) == ints(c["draws"])
        elif k == "compute_shares":
            assert be.compute_shares(ints(c["coeffs"]), c["n"]) == ints(c["shares"]), (k, c["n"], c["d"])
        elif k == "robust_recover":
            rc, coeffs, secret = be.robust_recover(c["ids"], ints(c["values"]), c["n"], c["t"], c["d"])
            assert rc == c["rc"], (k, c["n"], c.get("pattern"), c.get("order"))
            if rc == 0:
                assert padded(coeffs, ints(c["coeffs"])) and secret == int(c["secret"], 16), (k, c["n"], c.get("pattern"), c.get("order"))
        elif k == "batch_recover":
            rc, coeffs = be.batch_recover(c["ids"], [ints(row) for row in c["evals"]], c["n"], c["d"], c["t"])
            assert rc == c["rc"], (k, c["n"], c.get("variant"))
            if rc == 0:
                assert all(padded(g, ints(w)) for g, w in zip(coeffs, c["coeffs"])) and len(coeffs) == len(c["coeffs"])
        elif k == "vandermonde":
            V, out = be.vandermonde(c["n"], c["t"], ints(c["inputs"]))
            assert V == [ints(r) for r in c["matrix"]] and out == ints(c["outputs"])
        elif k == "nonrobust":
            rc, coeffs, secret = be.nonrobust(c["ids"], ints(c["values"]), c["n"], c["d"])
            assert (rc == 0) == (c["rc"] == 0)
            if c["rc"] == 0:
                assert secret == int(c["recovered"], 16) and padded(coeffs, ints(c["coeffs"]))
        else:
            raise AssertionError("unknown case kind " + k)
    return seen


# ---- the real thing: reference-produced vectors (skipped until a file exists)
@pytest.mark.parametrize("path", FILES or [None])
def test_oracle_matches_reference_vectors(path):
    if path is None:
        pytest.skip("no tests/golden/reference_*.json: run tools/reference_golden/dump_reference_golden.rs inside the reference crate")
    check_cases(json.load(open(path))["cases"], OracleBackend())


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or [None])
def test_cuda_path_matches_reference_vectors(path, hb, ctx):
    if path is None:
        pytest.skip("no tests/golden/reference_*.json: run tools/reference_golden/dump_reference_golden.rs inside the reference crate")
    check_cases(json.load(open(path))["cases"], GpuBackend(hb, ctx))


# ---- the consumer on a stand-in of the same schema (Python model: pins nothing, proves the plumbing)
def hx(v):
    return hex(v)


def schema_sample():
    cases = []
    seed = bytes([1, 0, 0, 0, 0, 0, 0, 0, 0xB2] + [0] * 23)
    cases.append({"kind": "rng_draws", "seed": list(seed), "draws": [hx(v) for v in chacha_fr.sample_fr(seed, 0, 16)]})
    for n, t in [(4, 1), (7, 2), (10, 3), (16, 5)]:
        for d in (t, 2 * t):
            if d + t + 1 > n:
                continue
            sd = bytes([n, d] + [0] * 30)
            coeffs = chacha_fr.sample_polynomials(sd, 1, d)[0]
            shares = pm.compute_shares(coeffs, n, d)
            cases.append({"kind": "compute_shares", "n": n, "d": d, "coeffs": [hx(c) for c in coeffs], "shares": [hx(s) for s in shares]})
            for pname, errs in [("honest", []), ("first1", [0]), (f"first{t}", list(range(t))), ("over_t", list(range(t + 1)))]:
                for oname, order in [("all", list(range(n))), ("reversed", list(range(n))[::-1]), ("tail", list(range(n - (d + t + 1), n)))]:
                    vals = [shares[i] for i in order]
                    for k, pos in enumerate(errs):
                        if pos < len(vals):
                            vals[pos] = (vals[pos] + k + 7) % pm.R_MOD
                    case = {"kind": "robust_recover", "n": n, "t": t, "d": d, "ids": order, "values": [hx(v) for v in vals], "pattern": pname, "order": oname}
                    try:
                        rec = pm.robust_recover_secret([(i, v, d) for i, v in zip(order, vals)], n, t)
                        case.update({"rc": 0, "coeffs": [hx(c) for c in pm.p_norm(rec["coeffs"])], "secret": hx(rec["secret"])})
                    except pm.ShareErr as ex:
                        case["rc"] = ex.code
                    cases.append(case)
        d = t
        polys = chacha_fr.sample_polynomials(bytes([n, 99] + [0] * 30), 5, d)
        cols = [[pm.compute_shares(p, n, d)[i] for p in polys] for i in range(n)]
        for variant in ("honest", "one_bad_sender", "subset_reversed"):
            ev = [(i, list(c)) for i, c in enumerate(cols)]
            if variant == "one_bad_sender":
                ev[1] = (1, [(v + 3) % pm.R_MOD for v in ev[1][1]])
            if variant == "subset_reversed":
                ev = ev[::-1][: d + t + 1]
            case = {"kind": "batch_recover", "n": n, "t": t, "d": d, "variant": variant, "ids": [e[0] for e in ev], "evals": [[hx(v) for v in e[1]] for e in ev]}
            try:
                res = pm.batch_recover_secret(ev, n, d, t)
                case.update({"rc": 0, "coeffs": [[hx(c) for c in p["coeffs"]] for p in res]})
            except pm.ShareErr as ex:
                case["rc"] = ex.code
            cases.append(case)
        inputs = chacha_fr.sample_fr(bytes([n, 7] + [0] * 30), 0, t + 1)
        V = pm.make_vandermonde(n, t)
        cases.append({"kind": "vandermonde", "n": n, "t": t, "matrix": [[hx(v) for v in r] for r in V], "inputs": [hx(v) for v in inputs],
                      "outputs": [hx(v) for v in pm.apply_vandermonde(V, inputs)]})
        coeffs = chacha_fr.sample_polynomials(bytes([n, 5] + [0] * 30), 1, d)[0]
        sh = pm.compute_shares(coeffs, n, d)
        cases.append({"kind": "nonrobust", "n": n, "d": d, "secret": hx(coeffs[0]), "ids": list(range(n)), "values": [hx(v) for v in sh], "rc": 0,
                      "coeffs": [hx(c) for c in pm.p_norm(coeffs)], "recovered": hx(coeffs[0])})
    return cases


def test_consumer_on_a_schema_sample():
    seen = check_cases(schema_sample(), OracleBackend())
    assert set(seen) == {"rng_draws", "compute_shares", "robust_recover", "batch_recover", "vandermonde", "nonrobust"}


@pytest.mark.gpu
def test_consumer_on_a_schema_sample_gpu(hb, ctx):
    seen = check_cases(schema_sample(), GpuBackend(hb, ctx))
    assert len(seen) == 6
