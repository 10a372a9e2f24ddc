"""Generates tests/golden/hotpath_golden.json from the independent Python big-int model (oracle/pymodel.py).

The reference is Rust and cannot be built or imported here (no rustc/cargo; arkworks not vendored), so these vectors
are NOT outputs of the reference: they are outputs of a second, independent restatement and pin the C oracle and the
CUDA path against accidental drift.  The reference's own known-answer tests are asserted separately in
tests/test_oracle_kats.py.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pymodel as pm  # noqa: E402


def hx(v):
    return hex(v)


def main():
    g = {"generator": "tests/golden/make_golden.py (oracle/pymodel.py, SplitMix64 seeds below)", "cases": []}
    # K1 + K2: compute_shares / apply_vandermonde
    for n, d, seed in [(4, 1, 11), (7, 2, 12), (10, 3, 13), (16, 5, 14), (16, 10, 15), (64, 21, 16), (128, 42, 17)]:
        rng = pm.SplitMix64(seed)
        coeffs = [rng.fr() for _ in range(d + 1)]
        shares = pm.compute_shares(coeffs, n, d)
        g["cases"].append({"kind": "compute_shares", "n": n, "d": d, "seed": seed, "coeffs": [hx(c) for c in coeffs],
                           "shares": [hx(s) for s in shares]})
    # K3/K4: robust recover with errors (n small enough for the naive model)
    for n, t, d, S, errs, seed in [(7, 2, 2, 7, [1], 21), (7, 2, 2, 7, [0, 6], 22), (10, 3, 3, 10, [2, 5, 9], 23), (10, 3, 3, 9, [0, 1], 24),
                                   (10, 3, 6, 10, [], 25), (16, 5, 5, 16, [0, 1, 2, 3, 4], 26), (16, 5, 5, 13, [3, 12], 27),
                                   (10, 3, 3, 8, [0, 1], 28), (10, 3, 3, 10, [0, 1, 2, 3], 29)]:
        rng = pm.SplitMix64(seed)
        coeffs = [rng.fr() for _ in range(d + 1)]
        shares = pm.compute_shares(coeffs, n, d)
        ids = list(range(n))[:S] if seed % 2 else list(range(n - S, n))
        vals = [shares[i] for i in ids]
        for e in errs:
            if e < len(vals):
                vals[e] = (vals[e] + 1 + rng.next() % 1000) % pm.R_MOD
        case = {"kind": "robust_recover", "n": n, "t": t, "d": d, "ids": ids, "seed": seed, "values": [hx(v) for v in vals],
                "true_coeffs": [hx(c) for c in coeffs]}
        try:
            rec = pm.robust_recover_secret([(i, v, d) for i, v in zip(ids, vals)], n, t)
            case.update({"rc": 0, "coeffs": [hx(c) for c in rec["coeffs"]], "secret": hx(rec["secret"]), "path": rec["path"],
                         "flags": [int(f) for f in rec["flags"]]})
        except pm.ShareErr as ex:
            case.update({"rc": ex.code})
        g["cases"].append(case)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", out, len(g["cases"]), "cases")


if __name__ == "__main__":
    main()
