import importlib, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
hb = importlib.import_module("mpc-protocols_b200")
from oracle import cmodel as cm
ctx = hb.Context(0)
for (n, t) in ((64, 21), (128, 42), (16, 5)):
    d = t
    coeffs = cm.random_fr((64, d + 1), 1)
    sh = ctx.compute_shares_batch(coeffs, n)
    rng = np.random.default_rng(1)
    for S in (d + t + 1, n):
        ts = []
        for rep in range(4):
            ids = np.sort(rng.permutation(n)[:S]) if S < n else np.arange(n)
            ev = np.ascontiguousarray(sh[:, ids].transpose(1, 0, 2))
            t0 = time.perf_counter(); ctx.batch_recover(ids, ev, n, d, t); t1 = time.perf_counter()
            ctx.batch_recover(ids, ev, n, d, t); t2 = time.perf_counter()
            ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 3)))
        print(n, t, S, "first-call ms / cached-call ms:", ts)
