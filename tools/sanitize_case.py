"""Smallest cases for compute-sanitizer (one tool per gpurun call): __graft_entry__.smoke() (n=16: K1, K2, K3 with flags, robust decoder)
plus one n=64 case through ntt16x_kernel, ntt64_cta_kernel, the dense kernel and the staged decoder (forced on a test-sized batch)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.smoke()
os.environ["HBMPC_STAGED_MIN"] = "64"
os.environ["HBMPC_SCAN_MAX"] = "0"
os.environ["HBMPC_NTT16X"] = "2"
hb = importlib.import_module("mpc-protocols_b200")
from oracle import cmodel as cm
ctx = hb.Context(0)
n, t, d, B = 64, 21, 21, 600
coeffs = cm.random_fr((B, d + 1), 0x5EEDC0DE)
shares = ctx.compute_shares_batch(coeffs, n)                       # ntt16x_kernel<6,0>
rc, want = cm.compute_shares(coeffs, n, threads=cm.max_threads())
assert np.array_equal(shares, want)
rng = np.random.default_rng(3)
bad = shares.copy()
for b in range(0, B, 2):
    for p in rng.choice(n, size=int(rng.integers(1, t + 1)), replace=False):
        bad[b, p, 0] ^= np.uint64(0x33)
ids = np.arange(n)
ev = np.ascontiguousarray(bad.transpose(1, 0, 2))
rc, co, path, flags = ctx.batch_recover(ids, ev, n, d, t, want_flags=True)   # ntt16x<6,1> + staged decoder on ~300 failing chunks
ref = cm.batch_recover_secret(ids, ev, n, d, t, threads=cm.max_threads())
assert rc == ref["rc"] and np.array_equal(co, ref["coeffs"]) and np.array_equal(path, ref["path"]) and np.array_equal(flags, ref["flags"][:, :1])
rc, co, path, _ = ctx.batch_recover(ids[:50], ev[:50], n, d, t)                # erasure-weighted transform (MODE 2) + triangular recovery
ref = cm.batch_recover_secret(ids[:50], ev[:50], n, d, t, threads=cm.max_threads())
assert np.array_equal(path, ref["path"]) and np.array_equal(co[path >= 0], ref["coeffs"][path >= 0])
ctx.close()
os.environ["HBMPC_NTT16X"] = "0"
os.environ["HBMPC_NTT_CTA"] = "2"
ctx = hb.Context(0)
assert np.array_equal(ctx.compute_shares_batch(coeffs, n), want)               # ntt64_cta_kernel<0>
rc, co, path, flags = ctx.batch_recover(ids, ev, n, d, t, want_flags=True)     # ntt64_cta_kernel<1>
assert np.array_equal(co[path >= 0], coeffs[path >= 0])
ctx.close()
print("sanitize case ok")
