"""The realistic first call of batch reconstruction: batch_recover with the first d+t+1 = 43 of 64 senders (random id subset, arrival order,
batch_recon.rs:371-409), device-resident, CUDA events: time of the call (erasure-weighted inverse NTT + triangular coefficient recovery)
and of batch_recover_secrets.   python tools/first_call_probe.py [--log2 20]    (also the ncu target of tools/gpu_ncu_first_call.sh)"""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--log2", type=int, default=20); ap.add_argument("--reps", type=int, default=5); a = ap.parse_args()
    n, t, d, B = 64, 21, 21, 1 << a.log2
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_async(True)
    coeffs = random_fr_device(torch, (B, d + 1), 0x43, dev)
    shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
    ctx.compute_shares_batch(coeffs, n, out=shares)
    arrival = np.random.default_rng(0x5EED43).permutation(n)[: d + t + 1]
    ev43 = shares.permute(1, 0, 2)[torch.as_tensor(arrival, device=dev)].contiguous()
    rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
    sec = torch.empty((B, 4), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    res = {"chunks": B, "senders": len(arrival)}
    for name, fn in (("coeffs_ms", lambda: ctx.batch_recover(arrival, ev43, n, d, t, out=(rec, path, None))),
                     ("secrets_ms", lambda: ctx.batch_recover_secrets(arrival, ev43, n, d, t, out=(sec, path)))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / a.reps, 4)
    assert ctx.synchronize() == 0 and torch.equal(rec, coeffs) and torch.equal(sec, coeffs[:, 0]) and not bool(path.any())
    res["identical_to_the_dealt_polynomials"] = True
    print(json.dumps(res))


if __name__ == "__main__":
    main()
