"""K5 (share algebra), fused pass against the operator-by-operator route, device-resident, CUDA events on the context's stream:
time and achieved HBM rate of a*b - r_2t (triple_generation.rs:332-340), of the pair a-x / b-y (multiplication.rs:417-426) and of the Beaver
product share (multiplication.rs:79-97).   python tools/k5_fused_probe.py [--log2 24]"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--log2", type=int, default=24); a = ap.parse_args()
    T = 1 << a.log2
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_async(True)
    v = [random_fr_device(torch, (T,), 40 + s, dev) for s in range(5)]
    o = [torch.empty_like(v[0]) for _ in range(4)]
    ew, fu = ctx.elementwise, ctx.share_algebra_fused
    cases = {
        "triple_mask": (lambda: (ew(2, v[0], v[1], out=o[0]), ew(1, o[0], v[2], out=o[1])), lambda: fu(0, v[:3], out=o[2]), 6, 4, (1, 2)),
        "beaver_mask": (lambda: (ew(1, v[0], v[1], out=o[0]), ew(1, v[2], v[3], out=o[1])), lambda: fu(1, v[:4], out=(o[2], o[3])), 6, 6, (0, 2)),
        "beaver_finalize": (lambda: (ew(2, v[3], v[4], out=o[0]), ew(1, v[0], o[0], out=o[1]), ew(2, v[2], v[3], out=o[0]), ew(1, o[1], o[0], out=o[1]),
                                     ew(2, v[1], v[4], out=o[0]), ew(1, o[1], o[0], out=o[1])), lambda: fu(2, v, out=o[2]), 18, 6, (1, 2)),
    }
    res = {"elements": T, "bytes_per_stream": 32 * T, "cases": {}}
    for name, (unfused, fused, s_un, s_fu, cmp) in cases.items():
        row = {}
        for label, fn, streams in (("operator_route", unfused, s_un), ("fused", fused, s_fu)):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            row[label] = {"ms": round(ms, 4), "hbm_streams": streams, "GBps": round(streams * 32 * T / ms / 1e6, 1)}
        assert ctx.synchronize() == 0
        row["identical"] = bool(torch.equal(o[cmp[0]], o[cmp[1]]))
        row["speedup"] = round(row["operator_route"]["ms"] / row["fused"]["ms"], 2)
        res["cases"][name] = row
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
