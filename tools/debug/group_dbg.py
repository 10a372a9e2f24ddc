import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device
ids = np.arange(64)
for g in (0, 1):
    dv = torch.device("cuda", g)
    with torch.cuda.device(dv):
        c = hb.Context(g)
        c.set_async(True)
        B = 1 << 18
        co = random_fr_device(torch, (B, 22), 5 + g, dv)
        sh = torch.empty((B, 64, 4), dtype=torch.int64, device=dv)
        c.compute_shares_batch(co, 64, out=sh)
        torch.cuda.synchronize(dv)
        print("dev", g, "sync after gen:", c.synchronize(), c.last_error())
        ev = sh.permute(1, 0, 2).contiguous()
        rec = torch.empty((B, 22, 4), dtype=torch.int64, device=dv)
        pth = torch.empty((B,), dtype=torch.int32, device=dv)
        torch.cuda.synchronize(dv)
    for i in range(3):
        c.compute_shares_batch(co, 64, out=sh)
        c.batch_recover(ids, ev, 64, 21, 21, out=(rec, pth, None))
        print("dev", g, "step", i, "sync:", c.synchronize(), c.last_error(), bool(torch.equal(rec, co)))
