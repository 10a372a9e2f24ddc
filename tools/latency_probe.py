"""Per-call latency at session sizes (the reference's sessions are <= 2048 columns, honeybadger/mod.rs:1432-1435): n=64, t=21, device
pointers, synchronous and asynchronous calls, wall clock per call over many repetitions.   python tools/latency_probe.py [--chunks 4096]"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=4096)
ap.add_argument("--reps", type=int, default=300)
a = ap.parse_args()
n, t, d, B = 64, 21, 21, a.chunks
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
ctx.use_torch_stream()
coeffs = random_fr_device(torch, (B, d + 1), 9, dev)
shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
ctx.compute_shares_batch(coeffs, n, out=shares)
ev = shares.permute(1, 0, 2).contiguous()
ids43 = np.random.default_rng(1).permutation(n)[:43]
ev43 = ev[torch.from_numpy(ids43).to(dev)].contiguous()
rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
sec = torch.empty((B, 4), dtype=torch.int64, device=dev)
path = torch.empty((B,), dtype=torch.int32, device=dev)
rmaj = torch.empty((n, B, 4), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
calls = {
    "compute_shares": lambda: ctx.compute_shares_batch(coeffs, n, out=shares),
    "apply_vandermonde_recipient_major": lambda: ctx.apply_vandermonde_batch(coeffs, n, recipient_major=True, out=rmaj),
    "batch_recover_64_senders": lambda: ctx.batch_recover(np.arange(n), ev, n, d, t, out=(rec, path, None)),
    "batch_recover_43_senders": lambda: ctx.batch_recover(ids43, ev43, n, d, t, out=(rec, path, None)),
    "batch_recover_secrets_43_senders": lambda: ctx.batch_recover_secrets(ids43, ev43, n, d, t, out=(sec, path)),
}
out = {"chunks": B, "n": n, "t": t, "us_per_call": {}}
for mode in ("sync", "async"):
    ctx.set_async(mode == "async")
    for name, fn in calls.items():
        for _ in range(20):
            fn()
        ctx.synchronize(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        ctx.synchronize(); torch.cuda.synchronize()
        out["us_per_call"][f"{name}_{mode}"] = round(1e6 * (time.perf_counter() - t0) / a.reps, 1)
assert torch.equal(rec, coeffs)
print(json.dumps(out, indent=1))
