"""BASELINE config 5 (RanSha + DouSha + RanDouSha + Beaver triple generation, n=64, t=21): ONE party's local field work for a
batch of triples, device-resident, CUDA events.  The FSMs / RBC / network stay on the host in the reference; this measures
the kernels a drop-in would call per session.   python tools/bench_c5.py [--log2-triples 20]"""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--log2-triples", type=int, default=20); a = ap.parse_args()
    n, t = 64, 21
    T = 1 << a.log2_triples
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_async(True)
    ids = np.arange(n)
    E = lambda *shape: torch.empty(tuple(shape) + (4,), dtype=torch.int64, device=dev)
    cols_rs = -(-2 * T // (n - 2 * t))      # RanSha columns: n-2t outputs each, 2 random shares per triple
    cols_ds = -(-T // (t + 1))              # DouSha / RanDouSha columns: t+1 outputs each
    groups = -(-T // (2 * t + 1))           # triple groups of 2t+1 (one batch-recon chunk each)
    # ---- synthetic inputs (consistent sharings so that every check takes the honest path)
    c_t = random_fr_device(torch, (cols_rs, t + 1), 1, dev); sh_t = E(cols_rs, n)
    c_d = random_fr_device(torch, (cols_ds, t + 1), 2, dev); c_d2 = random_fr_device(torch, (cols_ds, 2 * t + 1), 3, dev)
    c_d2[:, 0] = c_d[:, 0]
    sh_d, sh_d2 = E(cols_ds, n), E(cols_ds, n)
    recv = random_fr_device(torch, (cols_rs, n), 4, dev); mix = E(cols_rs, n)           # shares received from the n dealers
    recv_d, mix_d = random_fr_device(torch, (cols_ds, n), 5, dev), E(cols_ds, n)
    aS, bS, r2S, rtS = (random_fr_device(torch, (T,), s, dev) for s in (6, 7, 8, 9))    # own shares of a, b, r_2t, r_t
    masked, cS = E(T), E(T)
    grp = random_fr_device(torch, (groups, 2 * t + 1), 10, dev)                          # opened values a*b - r per group
    y_enc = E(n, groups)
    y_all = E(groups, n); ctx.compute_shares_batch(grp, n, out=y_all)                    # what the n parties would send (degree 2t)
    y_sm = y_all.permute(1, 0, 2).contiguous()
    sec1, p1 = E(groups), torch.empty((groups,), dtype=torch.int32, device=dev)
    co2, p2 = E(groups, 2 * t + 1), torch.empty((groups,), dtype=torch.int32, device=dev)
    ver_co, ver_sec, ver_p = E(cols_rs, t + 1), E(cols_rs), torch.empty((cols_rs,), dtype=torch.int32, device=dev)
    ctx.compute_shares_batch(c_t, n, out=sh_t)
    chk_co, chk_sec, chk_st = E(cols_ds, t + 1), E(cols_ds), torch.empty((cols_ds,), dtype=torch.int32, device=dev)
    chk_co2 = E(cols_ds, 2 * t + 1)
    ctx.compute_shares_batch(c_d, n, out=sh_d); ctx.compute_shares_batch(c_d2, n, out=sh_d2)
    assert ctx.synchronize() == 0

    phases = {
        "ransha_deal (K1 d=t, 1 secret/column)": lambda: ctx.compute_shares_batch(c_t, n, out=sh_t),
        "ransha_mix (K2 64x64 per column)": lambda: ctx.apply_vandermonde_batch(recv, n, out=mix),
        "ransha_verify (robust recover of one opened row per column, all n shares)": lambda: ctx.robust_interpolate_batch(ids, sh_t, n, t, t, out=(ver_co, ver_sec, ver_p, None)),
        "dousha_deal (K1 d=t and d=2t per column)": lambda: (ctx.compute_shares_batch(c_d, n, out=sh_d), ctx.compute_shares_batch(c_d2, n, out=sh_d2)),
        "randousha_mix (2x K2 64x64 per column)": lambda: (ctx.apply_vandermonde_batch(recv_d, n, out=mix_d), ctx.apply_vandermonde_batch(recv_d, n, out=mix_d)),
        "randousha_check (NonRobust recover deg t and 2t, all n shares)": lambda: (ctx.nonrobust_recover_batch(ids, sh_d, n, t, out=(chk_co, chk_sec, chk_st)),
                                                                                 ctx.nonrobust_recover_batch(ids, sh_d2, n, 2 * t, out=(chk_co2, chk_sec, chk_st))),
        "triple_mask (K5 fused: a*b - r_2t per triple, one pass)": lambda: ctx.share_algebra_fused(0, (aS, bS, r2S), out=masked),
        "triple_open_encode (K2 64x43 per group, recipient-major)": lambda: ctx.apply_vandermonde_batch(grp, n, recipient_major=True, out=y_enc),
        "triple_open_round1 (batch_recover_secrets d=2t, 64 senders)": lambda: ctx.batch_recover_secrets(ids, y_sm, n, 2 * t, t, out=(sec1, p1)),
        "triple_open_round2 (batch_recover d=2t, 64 senders)": lambda: ctx.batch_recover(ids, y_sm, n, 2 * t, t, out=(co2, p2, None)),
        "triple_finish (K5: r_t + opened)": lambda: ctx.elementwise(0, rtS, masked, out=cS),
    }
    res, total = {}, 0.0
    for name, fn in phases.items():
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = round(ms, 4); total += ms
    assert ctx.synchronize() == 0
    assert torch.equal(co2, grp) and torch.equal(chk_sec, c_d[:, 0]) and int(chk_st.min()) == 2 * t and torch.equal(ver_sec, c_t[:, 0])
    print(json.dumps({"config": "C5 one party, n=64, t=21", "triples": T, "ransha_columns": cols_rs, "dousha_columns": cols_ds, "groups": groups,
                      "phase_ms": res, "total_ms": round(total, 3), "triples_per_s_per_gpu": T / (total * 1e-3)}, indent=1))

if __name__ == "__main__":
    main()
