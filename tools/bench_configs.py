"""Secondary measurements (not the driver's bench line): the other BASELINE.json configs, device-resident, CUDA events.
   python tools/bench_configs.py [--log2 B] [--which c2,c4,c5]"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=20)
    ap.add_argument("--log2-c4", type=int, default=16)
    ap.add_argument("--which", default="c2,c3r,c4,c5")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_async(True)
    out = {}
    which = a.which.split(",")
    if "c2" in which:
        n, t, d, B = 16, 5, 5, 1 << a.log2
        coeffs = random_fr_device(torch, (B, d + 1), 2, dev)
        shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
        tg = timed(lambda: ctx.compute_shares_batch(coeffs, n, out=shares))
        ev = shares.permute(1, 0, 2).contiguous()
        rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev); path = torch.empty((B,), dtype=torch.int32, device=dev)
        tr = timed(lambda: ctx.batch_recover(np.arange(n), ev, n, d, t, out=(rec, path, None)))
        assert ctx.synchronize() == 0 and torch.equal(rec, coeffs)
        out["c2_n16_t5"] = {"B": B, "gen_ms": tg, "recon_ms": tr, "shares_per_s": 2 * B * n / ((tg + tr) * 1e-3)}
    if "c5" in which:
        n, t, B = 64, 21, 1 << a.log2
        x = random_fr_device(torch, (B, n), 5, dev)
        y = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
        t_nn = timed(lambda: ctx.apply_vandermonde_batch(x, n, out=y))               # RanSha / RanDouSha n x n
        x2 = random_fr_device(torch, (B, 2 * t + 1), 6, dev)
        t_43 = timed(lambda: ctx.apply_vandermonde_batch(x2, n, out=y))              # triple open encode, cols = 2t+1
        yt = torch.empty((n, B, 4), dtype=torch.int64, device=dev)
        t_43t = timed(lambda: ctx.apply_vandermonde_batch(x2, n, recipient_major=True, out=yt))
        rec = torch.empty((B, 2 * t + 1, 4), dtype=torch.int64, device=dev); path = torch.empty((B,), dtype=torch.int32, device=dev)
        t_r2t = timed(lambda: ctx.batch_recover(np.arange(n), yt, n, 2 * t, t, out=(rec, path, None)))   # d = 2t: needed = 64, m = 43
        assert ctx.synchronize() == 0 and torch.equal(rec, x2)
        a_, b_ = random_fr_device(torch, (B * 8,), 7, dev), random_fr_device(torch, (B * 8,), 8, dev)
        o_ = torch.empty_like(a_)
        t_mul = timed(lambda: ctx.elementwise(2, a_, b_, out=o_))
        out["c5_n64_t21"] = {"B": B, "vandermonde_64x64_ms": t_nn, "vandermonde_64x43_ms": t_43, "vandermonde_64x43_recipient_major_ms": t_43t,
                             "recover_d2t_ms": t_r2t, "elementwise_mul_ms": t_mul, "elementwise_GBs": B * 8 * 96 / (t_mul * 1e-3) / 1e9}
    if "c3r" in which:   # batch reconstruction under attack at n=64,t=21: every chunk carries errors from the t corrupted senders
        n, t, d, B = 64, 21, 21, 1 << a.log2
        coeffs = random_fr_device(torch, (B, d + 1), 14, dev)
        shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
        ctx.compute_shares_batch(coeffs, n, out=shares)
        ev = shares.permute(1, 0, 2).contiguous()
        bad_senders = torch.randperm(n, device=dev)[:t]
        ev[bad_senders, :, 0] ^= 0x77
        rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev); path = torch.empty((B,), dtype=torch.int32, device=dev)
        ms_async = timed(lambda: ctx.batch_recover(np.arange(n), ev, n, d, t, out=(rec, path, None)), reps=2)
        rc = ctx.synchronize()
        ctx.set_async(False)   # synchronous calls may use the persistent-attacker shortcut (needs a host decision mid-call)
        ms = timed(lambda: ctx.batch_recover(np.arange(n), ev, n, d, t, out=(rec, path, None)), reps=2)
        ctx.set_async(True)
        out["c3_under_attack_n64_t21"] = {"B": B, "corrupted_senders": t, "ms": ms, "chunks_per_s": B / (ms * 1e-3), "ms_async_full_decoder": ms_async, "chunks_per_s_full_decoder": B / (ms_async * 1e-3), "rc": rc,
                                          "all_recovered": bool(torch.equal(rec, coeffs)), "max_path": int(path.max())}
    if "lat" in which:   # per-call latency at session-sized batches (device pointers, synchronous mode, wall clock)
        import time
        n, t, d = 64, 21, 21
        ctx.set_async(False)
        lat = {}
        for B in (64, 1024, 4096, 16384):
            coeffs = random_fr_device(torch, (B, d + 1), 24, dev)
            shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
            ctx.compute_shares_batch(coeffs, n, out=shares)
            evs = shares.permute(1, 0, 2).contiguous()
            ev43 = evs[:43].contiguous()
            rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev); path = torch.empty((B,), dtype=torch.int32, device=dev)
            sec = torch.empty((B, 4), dtype=torch.int64, device=dev)
            def wall(fn, reps=50):
                fn(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps): fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / reps * 1e6
            lat[B] = {"compute_shares_us": wall(lambda: ctx.compute_shares_batch(coeffs, n, out=shares)),
                      "batch_recover_64_senders_us": wall(lambda: ctx.batch_recover(np.arange(n), evs, n, d, t, out=(rec, path, None))),
                      "batch_recover_43_senders_us": wall(lambda: ctx.batch_recover(np.arange(43), ev43, n, d, t, out=(rec, path, None))),
                      "batch_recover_secrets_43_us": wall(lambda: ctx.batch_recover_secrets(np.arange(43), ev43, n, d, t, out=(sec, path)))}
        ctx.set_async(True)
        out["latency_n64_t21"] = lat
    if "c4" in which:
        n, t, d, B = 128, 42, 42, 1 << a.log2_c4
        coeffs = random_fr_device(torch, (B, d + 1), 4, dev)
        shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
        ctx.compute_shares_batch(coeffs, n, out=shares)
        ids = np.arange(n)
        res = {}
        g = torch.Generator(device=dev); g.manual_seed(44)
        def corrupt(kind):
            bad = shares.clone()
            if kind == "clean":
                return bad
            if kind == "uniform":      # e ~ U{0..t} errors at uniform distinct positions
                e = torch.randint(0, t + 1, (B,), device=dev, generator=g)
                perm = torch.rand((B, n), device=dev, generator=g).argsort(dim=1)
            else:                      # adversarial: exactly t errors, all at ids < d+t+1
                e = torch.full((B,), t, device=dev)
                perm = torch.rand((B, d + t + 1), device=dev, generator=g).argsort(dim=1)
            mask = torch.zeros((B, n), dtype=torch.bool, device=dev)
            sel = torch.arange(perm.shape[1], device=dev)[None, :] < e[:, None]
            mask.scatter_(1, perm, sel)
            bad[..., 0] = torch.where(mask, bad[..., 0] ^ 0x5A5A5, bad[..., 0])
            return bad
        co = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev); sec = torch.empty((B, 4), dtype=torch.int64, device=dev)
        path = torch.empty((B,), dtype=torch.int32, device=dev); fl = torch.empty((B, 2), dtype=torch.int64, device=dev)
        for kind in ("clean", "uniform", "adversarial"):
            bad = corrupt(kind)
            ms_async = timed(lambda: ctx.robust_interpolate_batch(ids, bad, n, d, t, out=(co, sec, path, fl)), reps=2)
            rc = ctx.synchronize()
            ok = bool(torch.equal(co, coeffs))
            co.zero_()
            ctx.set_async(False)   # synchronous calls know the failing count on the host: large failing sets take the staged decoder
            ms = timed(lambda: ctx.robust_interpolate_batch(ids, bad, n, d, t, out=(co, sec, path, fl)), reps=2)
            ctx.set_async(True)
            ok = ok and bool(torch.equal(co, coeffs))
            res[kind] = {"ms": ms, "codewords_per_s": B / (ms * 1e-3), "ms_async_per_thread_decoder": ms_async,
                         "codewords_per_s_per_thread_decoder": B / (ms_async * 1e-3), "rc": rc, "all_recovered": ok, "max_path": int(path.max())}
        out["c4_n128_t42"] = {"B": B, **res}
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
