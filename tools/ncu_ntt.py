"""Minimal driver for ncu captures of the headline kernels: share generation (K1) and all-senders batch recovery (K3) at
n=64, t=21, B = 2^log2 device-resident, launched alternately `reps` times (the first pair warms tables and caches).
   python tools/ncu_ntt.py [--log2 20] [--reps 3] [--senders 64]"""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--senders", type=int, default=64)
a = ap.parse_args()
n, t, d, B = 64, 21, 21, 1 << a.log2
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_async(True)
coeffs = random_fr_device(torch, (B, d + 1), 3, dev)
shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
ctx.compute_shares_batch(coeffs, n, out=shares)
evals = shares.permute(1, 0, 2).contiguous()[: a.senders].contiguous()
ids = np.arange(a.senders)
rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
path = torch.empty((B,), dtype=torch.int32, device=dev)
for _ in range(a.reps):
    ctx.compute_shares_batch(coeffs, n, out=shares)
    ctx.batch_recover(ids, evals, n, d, t, out=(rec, path, None))
assert ctx.synchronize() == 0 and torch.equal(rec, coeffs)
print("ok", ctx.launch_count)
