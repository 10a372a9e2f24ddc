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
ap.add_argument("--what", default="gen_recon", choices=["gen_recon", "dense", "k4"],
                help="dense: the 43-sender call with flags through matvec_kernel (HBMPC_NO_ER_FLAGS=1); k4: robust_interpolate_batch n=128,t=42, e~U{0..42}")
a = ap.parse_args()
if a.what == "dense":
    os.environ["HBMPC_NO_ER_FLAGS"] = "1"
    a.senders = 43
if a.what == "k4":
    n, t, d, B = 128, 42, 42, 1 << a.log2
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_async(True)
    coeffs = random_fr_device(torch, (B, d + 1), 4, dev)
    shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
    ctx.compute_shares_batch(coeffs, n, out=shares)
    g = torch.Generator(device=dev); g.manual_seed(44)
    e = torch.randint(0, t + 1, (B,), device=dev, generator=g)
    perm = torch.rand((B, n), device=dev, generator=g).argsort(dim=1)
    mask = torch.zeros((B, n), dtype=torch.bool, device=dev)
    mask.scatter_(1, perm, torch.arange(n, device=dev)[None, :] < e[:, None])
    shares[..., 0] = torch.where(mask, shares[..., 0] ^ 0x5A5A5, shares[..., 0])
    out = (torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev), torch.empty((B, 4), dtype=torch.int64, device=dev),
           torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B, 2), dtype=torch.int64, device=dev))
    ctx.set_async(False)
    l0 = 0
    for r in range(a.reps):
        l0 = ctx.launch_count
        if r == a.reps - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()   # ncu --profile-from-start off: only the last call is captured
        ctx.robust_interpolate_batch(np.arange(n), shares, n, d, t, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    assert torch.equal(out[0], coeffs)
    print("ok launches_per_call", ctx.launch_count - l0, "total", ctx.launch_count)
    sys.exit(0)
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
flags = torch.empty((B, 1), dtype=torch.int64, device=dev) if a.what == "dense" else None
for _ in range(a.reps):
    ctx.compute_shares_batch(coeffs, n, out=shares)
    ctx.batch_recover(ids, evals, n, d, t, out=(rec, path, flags))
assert ctx.synchronize() == 0 and torch.equal(rec, coeffs)
print("ok", ctx.launch_count)
