"""Asynchronous (enqueue-only) K4 calls under attack: robust_kernel (HBMPC_ASYNC_STAGED=0) against the staged decoder in device-count
mode (the route a context takes once hbmpc_ctx_synchronize has seen a large failing set; HBMPC_ASYNC_STAGED=2 forces it), and the
honest call of the same size through both.  us per call, device tensors."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import random_fr_device
hb = importlib.import_module("mpc-protocols_b200")
dev = torch.device("cuda", 0)
def ctx_with(env):
    os.environ.update(env)
    c = hb.Context(0)
    for k in env: del os.environ[k]
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    c.set_async(True)
    return c
routes = {"robust_kernel": ctx_with({"HBMPC_ASYNC_STAGED": "0"}), "staged_device_count": ctx_with({"HBMPC_ASYNC_STAGED": "2"})}
out = {}
for n, t in ((64, 21), (128, 42)):
    for B in (4096, 16384, 65536, 262144):
        coeffs = random_fr_device(torch, (B, t + 1), 5, dev)
        shares = routes["robust_kernel"].compute_shares_batch(coeffs, n)
        g = torch.Generator(device=dev); g.manual_seed(B)
        e = torch.randint(1, t + 1, (B,), device=dev, generator=g)
        rank = torch.rand((B, n), device=dev, generator=g).argsort(dim=1).argsort(dim=1)
        bad = shares.clone(); bad[..., 0] = torch.where(rank < e[:, None], bad[..., 0] ^ 0x5A5A5, bad[..., 0])
        ids = np.arange(n)
        row = {}
        for name, c in routes.items():
            for label, words in (("attack", bad), ("honest", shares)):
                o = c.robust_interpolate_batch(ids, words, n, t, t)
                assert c.synchronize() == 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5): o = c.robust_interpolate_batch(ids, words, n, t, t)
                assert c.synchronize() == 0
                torch.cuda.synchronize()
                row[f"{name}_{label}_us"] = round((time.perf_counter() - t0) / 5 * 1e6, 1)
                assert torch.equal(o[1], coeffs)
        row["speedup_attack"] = round(row["robust_kernel_attack_us"] / row["staged_device_count_attack_us"], 2)
        out[f"n{n}_B{B}"] = row
print(json.dumps(out, indent=1))
