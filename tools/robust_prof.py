"""Phase breakdown of robust_kernel (needs a library built with -DHB_ROBUST_PROF; see tools/robust_prof.sh)."""
import ctypes, importlib, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device
n, t, d, B = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 42, 42, 1 << 16)))
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
coeffs = random_fr_device(torch, (B, d + 1), 4, dev)
shares = ctx.compute_shares_batch(coeffs, n)
g = torch.Generator(device=dev); g.manual_seed(1)
e = torch.randint(0, t + 1, (B,), device=dev, generator=g)
perm = torch.rand((B, n), device=dev, generator=g).argsort(dim=1)
mask = torch.zeros((B, n), dtype=torch.bool, device=dev)
mask.scatter_(1, perm, torch.arange(n, device=dev)[None, :] < e[:, None])
bad = shares.clone(); bad[..., 0] = torch.where(mask, bad[..., 0] ^ 0x5A5A5, bad[..., 0])
rc, co, sec, path, _ = ctx.robust_interpolate_batch(np.arange(n), bad, n, d, t)
torch.cuda.synchronize()
cudart = ctypes.CDLL("libcudart.so")
lib = ctypes.CDLL(hb.LIB_PATH)
sym = ctypes.c_void_p()
buf = (ctypes.c_ulonglong * 8)()
# read the __device__ array through cudaMemcpyFromSymbol exposed by a tiny helper in the library
lib.hbmpc_debug_read_robust_prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
lib.hbmpc_debug_read_robust_prof(buf)
names = ["syndromes", "berlekamp_massey", "omega", "chien", "forney_eval", "inversion_values"]
tot = sum(buf[i] for i in range(6))
print(json.dumps({"n": n, "t": t, "B": B, "ok": bool(torch.equal(co, coeffs)), "phase_share": {nm: round(buf[i] / tot, 3) for i, nm in enumerate(names)}}))
