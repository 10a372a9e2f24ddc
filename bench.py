#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 finite-field hot path (see DESIGN.md "Measurement").

Metric (BASELINE.json): Fr shares generated + reconstructed per second at n=64, t=21.
One step = one pass of the hot path over one batch of B synthetic secrets on every rank:
    gen   : hbmpc_compute_shares_batch   coeffs[B][22] -> shares[B][64]            (K1, RobustShare::compute_shares)
    recon : hbmpc_batch_recover          evals[64][B] (all 64 senders) -> coeffs[B][22], path[B]   (K3, batch_recover_secret)
    value = N * (B*64 + B*64) / max-over-ranks(t_gen + t_rec)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2-batch 22] [--scaling weak|strong]
For N > 1 launch under torchrun (one rank per GPU); the batch of independent secrets is sharded (weak scaling: every rank owns
B secrets; --scaling strong: 2^log2-batch secrets in total), no collective on the data path; NCCL only gathers the reconstructed
secrets after the timed region.

Beside the headline the JSON line carries `configs` -- every BASELINE.json configuration measured outside the timed region, each
with a sample of its outputs compared against the CPU oracle (`parity_sample`): c2 (n=16,t=5, 2^20 secrets), c3_first_call (the
43-sender call a batch-reconstruction handler makes first: coefficients / secrets only / with flags), c3_strong (2^22 secrets in
total over all ranks), c4 (n=128,t=42, 2^20 codewords clean / uniform / adversarial errors), c5 (one party's field work for
2^21 Beaver triples per rank: 2^24 on 8 GPUs).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PARTIES, T_FAULTS = 64, 21
DEG = T_FAULTS
M = DEG + 1
NEEDED = DEG + T_FAULTS + 1
# algorithmic (dense, reference-faithful) modmul per secret and IMAD per modmul
ALG_MODMUL_GEN = N_PARTIES * DEG                      # 1344 (Horner count)
ALG_MODMUL_REC = (DEG + T_FAULTS + 1) * M + M * M     # 946 + 484 = 1430
IMAD_PER_MODMUL = 256
BYTES_GEN = M * 32 + N_PARTIES * 32                   # 704 R + 2048 W
BYTES_REC = (DEG + T_FAULTS + 1) * 32 + M * 32 + 4    # 1376 R + 704 W + path
METRIC = "Fr shares generated+reconstructed per second (n=64,t=21)"
DTYPE = "u32x8 (Fr, 255-bit Montgomery)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: 2^log2-batch secrets in total, split over the ranks")
    ap.add_argument("--log2-batch", type=int, default=22, help="secrets per rank per step (device-resident leg)")
    ap.add_argument("--log2-e2e-batch", type=int, default=20, help="secrets per rank per step (host-buffer leg)")
    ap.add_argument("--cpu-log2-batch", type=int, default=17, help="secrets per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of the other BASELINE configurations")
    ap.add_argument("--no-robust-leg", action="store_true", help="(kept for older scripts) same as --no-configs")
    ap.add_argument("--log2-c2", type=int, default=20)
    ap.add_argument("--log2-c4", type=int, default=20, help="codewords of the n=128, t=42 leg")
    ap.add_argument("--log2-c5", type=int, default=21, help="Beaver triples per rank of the preprocessing leg (2^24 over 8 ranks)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def load_oracle_native():
    """Builds the C oracle with -march=native ON THIS BOX (a binary built elsewhere may not run here)."""
    from oracle import cmodel

    stale = os.path.join(ROOT, "oracle", "liboracle_native.so")
    try:
        if os.path.exists(stale):
            os.remove(stale)
        cmodel.load(native=True)
    except Exception:
        cmodel.load(native=False)
    return cmodel


def cpu_step(cm, coeffs, threads, keep=None):
    """One pass of the same hot path on the CPU oracle (FFT share generation + batch_recover_secret)."""
    t0 = time.perf_counter()
    rc, shares = cm.compute_shares(coeffs, N_PARTIES, threads=threads)
    t1 = time.perf_counter()
    evals = np.ascontiguousarray(shares.transpose(1, 0, 2))  # message re-assembly, not timed (device leg has it resident too)
    t2 = time.perf_counter()
    out = cm.batch_recover_secret(np.arange(N_PARTIES), evals, N_PARTIES, DEG, T_FAULTS, threads=threads)
    t3 = time.perf_counter()
    assert rc == 0 and out["rc"] == 0 and np.array_equal(out["coeffs"], coeffs)
    if keep is not None:
        keep["shares"], keep["coeffs"], keep["path"] = shares, out["coeffs"], out["path"]
    return (t1 - t0) + (t3 - t2)


def cpu_rate(cm, coeffs, threads, budget_s, keep=None):
    reps, tcpu = 0, 0.0
    while reps < 2 or (tcpu < budget_s and reps < 64):
        tcpu += cpu_step(cm, coeffs, threads, keep if reps == 0 else None)
        reps += 1
    return reps * 2 * coeffs.shape[0] * N_PARTIES / tcpu, reps, tcpu


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cm = load_oracle_native()
    threads = cm.max_threads()
    B = 1 << args.cpu_log2_batch
    coeffs = cm.random_fr((B, M), 0x5EED0003)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_step(cm, coeffs, threads)
    times = [cpu_step(cm, coeffs, threads) for _ in range(args.steps)]
    total = sum(times)
    value = args.steps * 2 * B * N_PARTIES / total
    one = cm.random_fr((1 << 13, M), 0x5EED0013)
    v1, _, _ = cpu_rate(cm, one, 1, 1.0)   # the reference node itself is single-threaded: print that figure too
    sample = (f"{args.steps} steps, each a bounded sample of 2^{args.cpu_log2_batch} secrets of the workload in `config` (n=64,t=21): FFT compute_shares + "
              f"batch_recover_secret, C oracle port -march=native, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shares/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "shares/s", "cores": threads, "kind": "port", "sample": sample,
                         "value_1_thread": v1, "sample_secrets_per_step": B},
        "e2e": {"value": value, "unit": "shares/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the Rust reference cannot be built here (no rustc/cargo, arkworks not vendored): this arm times the C oracle port on all host cores; "
                "throughput in shares/s does not depend on the sample size, so the ratio against the GPU arm stands",
    }
    print(json.dumps(line))


def workload_config(args):
    """Identical in both arms (the reference arm processes a bounded sample of it per step, stated in its cpu_baseline.sample)."""
    return {
        "workload": "HoneyBadgerMPC share-gen + batch reconstruction over ark_bls12_381::Fr, n=64, t=21 (BASELINE configs[2] shape)",
        "n": N_PARTIES, "t": T_FAULTS, "degree": DEG,
        "secrets_per_step": f"2^{args.log2_batch} per rank" if args.scaling == "weak" else f"2^{args.log2_batch} in total, split over the ranks",
        "gen": "compute_shares coeffs[B][22] -> shares[B][64]", "recon": "batch_recover evals[64][B] -> coeffs[B][22] (S=64 senders, 43 examined)",
        "l2": "inputs (>= 2.9 GB per kernel at 2^22 secrets) are larger than the 126 MB L2; no explicit flush",
        "e2e_secrets_per_rank_per_step": 1 << args.log2_e2e_batch,
    }


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def random_fr_device(torch, shape, seed, device):
    """Synthetic canonical Fr values on the device: limbs 0..2 uniform 64-bit, top limb uniform below r's top limb (< r guaranteed)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randint(-(1 << 63), (1 << 63) - 1, tuple(shape) + (4,), dtype=torch.int64, device=device, generator=g)
    x[..., 3] = torch.randint(0, 0x73EDA753299D7D48, tuple(shape), dtype=torch.int64, device=device, generator=g)
    return x


def _np(x):
    """device int64 limbs -> host uint64 limbs"""
    return x.cpu().numpy().view(np.uint64)


def timed(torch, stream, fn, reps=3):
    """average device time of fn over `reps` back-to-back calls (CUDA events on the launching stream), after one warm-up call"""
    fn()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record(stream)
    for _ in range(reps):
        fn()
    d1.record(stream)
    torch.cuda.synchronize()
    return d0.elapsed_time(d1) / reps * 1e-3


# ---- BASELINE configs[1]: n=16, t=5
def leg_c2(torch, hb, ctx, stream, dev, cm, log2):
    n, t, d, B = 16, 5, 5, 1 << log2
    ids = np.arange(n)
    coeffs = random_fr_device(torch, (B, d + 1), 0x5EED0002, dev)
    shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
    tg = timed(torch, stream, lambda: ctx.compute_shares_batch(coeffs, n, out=shares))
    ev = shares.permute(1, 0, 2).contiguous()
    rec = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    tr = timed(torch, stream, lambda: ctx.batch_recover(ids, ev, n, d, t, out=(rec, path, None)))
    ok = ctx.synchronize() == 0 and bool(torch.equal(rec, coeffs)) and not bool(path.any())
    out = {"workload": "n=16, t=5: compute_shares + batch_recover (all 16 senders), device-resident", "secrets": B, "gen_ms": 1e3 * tg, "recon_ms": 1e3 * tr,
           "shares_per_s": 2 * B * n / (tg + tr), "roundtrip_ok": ok,
           "hbm_gbs": {"gen": B * (d + 1 + n) * 32 / tg / 1e9, "recon": B * (n + d + 1) * 32 / tr / 1e9}}
    if cm is not None:
        Bs = 4096
        cs = _np(coeffs[:Bs])
        rc, want = cm.compute_shares(cs, n, threads=cm.max_threads())
        evs = np.ascontiguousarray(want.transpose(1, 0, 2))
        ref = cm.batch_recover_secret(ids, evs, n, d, t, threads=cm.max_threads())
        pok = rc == 0 and np.array_equal(_np(shares[:Bs]), want) and np.array_equal(_np(rec[:Bs]), ref["coeffs"]) and np.array_equal(path[:Bs].cpu().numpy(), ref["path"])
        out["parity_sample"] = {"items": Bs, "ok": bool(pok), "what": "GPU shares / coefficients / path of the first 4096 secrets == C oracle"}
    return out


# ---- BASELINE configs[2], the call a batch-reconstruction handler makes first: exactly d+t+1 = 43 senders have arrived
def leg_c3_first_call(torch, hb, ctx, stream, dev, cm, coeffs, evals, rec, path, rank):
    B = coeffs.shape[0]
    rng = np.random.default_rng(0x5EED43)
    arrival = rng.permutation(N_PARTIES)[:NEEDED]               # ids of the first 43 arrivals, in arrival order
    ev43 = evals[torch.as_tensor(arrival, device=dev)].contiguous()
    flags43 = torch.empty((B, 1), dtype=torch.int64, device=dev)
    sec = torch.empty((B, 4), dtype=torch.int64, device=dev)

    def checked(fn, what):
        s = timed(torch, stream, fn)
        assert ctx.synchronize() == 0 and what(), "first-call recovery differs from the original polynomials"
        return s

    t_co = checked(lambda: ctx.batch_recover(arrival, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, None)), lambda: torch.equal(rec, coeffs))
    t_se = checked(lambda: ctx.batch_recover_secrets(arrival, ev43, N_PARTIES, DEG, T_FAULTS, out=(sec, path)), lambda: torch.equal(sec, coeffs[:, 0]))
    t_fl = checked(lambda: ctx.batch_recover(arrival, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, flags43)), lambda: torch.equal(rec, coeffs) and not bool(flags43.any()))
    # the dense kernel on its own (roofline_dense): a context that sends calls with flags straight to the dense check, as every
    # chunk with a disagreeing share goes (HBMPC_NO_ER_FLAGS is a test knob read at context creation)
    os.environ["HBMPC_NO_ER_FLAGS"] = "1"
    ctx_dense = hb.Context(torch.cuda.current_device())
    del os.environ["HBMPC_NO_ER_FLAGS"]
    ctx_dense.set_stream(stream.cuda_stream)
    ctx_dense.set_async(True)
    t_dense = timed(torch, stream, lambda: ctx_dense.batch_recover(arrival, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, flags43)))
    assert ctx_dense.synchronize() == 0 and torch.equal(rec, coeffs)
    ctx_dense.close()
    out = {"workload": "batch_recover / batch_recover_secrets with the first d+t+1 = 43 arrivals (random id subset, arrival order), 2^22 chunks per rank",
           "chunks": B, "coeffs_ms": 1e3 * t_co, "secrets_ms": 1e3 * t_se, "coeffs_flags_ms": 1e3 * t_fl, "dense_flags_ms": 1e3 * t_dense,
           "routes": "coeffs: erasure-weighted inverse NTT + triangular recovery; secrets: same transform, one product per chunk; flags: the transform over all supplied senders, dense check only for chunks it rejects; dense: matvec_kernel on every chunk"}
    if cm is not None and rank == 0:
        Bs = 4096
        evs = _np(ev43[:, :Bs].contiguous())
        ref = cm.batch_recover_secret(arrival, evs, N_PARTIES, DEG, T_FAULTS, threads=cm.max_threads())
        pok = ref["rc"] == 0 and np.array_equal(_np(rec[:Bs]), ref["coeffs"]) and np.array_equal(_np(sec[:Bs]), ref["coeffs"][:, 0]) and not ref["path"].any()
        out["parity_sample"] = {"items": Bs, "ok": bool(pok), "what": "coefficients and secrets of the first 4096 chunks == C oracle batch_recover_secret on the same 43 sender vectors"}
    return out, t_dense


# ---- BASELINE configs[3]: n=128, t=42 robust interpolation with injected errors
def leg_c4(torch, hb, ctx, stream, dev, cm, log2):
    n, t, d, B = 128, 42, 42, 1 << log2
    ids = np.arange(n)
    coeffs = random_fr_device(torch, (B, d + 1), 0x5EED0004, dev)
    shares = torch.empty((B, n, 4), dtype=torch.int64, device=dev)
    ctx.set_async(True)
    ctx.compute_shares_batch(coeffs, n, out=shares)
    g = torch.Generator(device=dev)
    g.manual_seed(44)
    needed = d + t + 1

    def corrupt(kind):
        if kind == "clean":
            return shares.clone(), torch.zeros((B, n), dtype=torch.bool, device=dev)
        if kind == "uniform":      # e ~ U{0..t} errors at uniform distinct positions
            e = torch.randint(0, t + 1, (B,), device=dev, generator=g)
            perm = torch.rand((B, n), device=dev, generator=g).argsort(dim=1)
        else:                      # adversarial: exactly t errors, all at ids < d+t+1 (forces the reference's last OEC round)
            e = torch.full((B,), t, device=dev)
            perm = torch.rand((B, needed), device=dev, generator=g).argsort(dim=1)
        mask = torch.zeros((B, n), dtype=torch.bool, device=dev)
        sel = torch.arange(perm.shape[1], device=dev)[None, :] < e[:, None]
        mask.scatter_(1, perm, sel)
        bad = shares.clone()
        bad[..., 0] = torch.where(mask, bad[..., 0] ^ 0x5A5A5, bad[..., 0])
        return bad, mask

    def expected_path(mask):
        """OEC round the reference accepts in (robust_interpolate.rs:579-628): the first r whose prefix of d+t+1+r ids holds at most r
        errors (r = 0: the optimistic check on the lowest d+t+1 ids passes)."""
        cum = mask.to(torch.int32).cumsum(dim=1)
        r = torch.arange(0, n - needed + 1, device=dev)
        in_prefix = cum[:, needed - 1:]                       # errors among the first d+t+1+r ids, r = 0 .. n-needed
        okr = in_prefix <= r[None, :]
        return torch.where(okr.any(dim=1), okr.to(torch.int32).argmax(dim=1), torch.full((B,), -1, device=dev, dtype=torch.int64)).to(torch.int32)

    co = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
    sec = torch.empty((B, 4), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    fl = torch.empty((B, 2), dtype=torch.int64, device=dev)
    weights = (torch.ones(64, dtype=torch.int64, device=dev) << torch.arange(64, device=dev))
    res = {}
    sample = None
    for kind in ("clean", "uniform", "adversarial"):
        bad, mask = corrupt(kind)
        ctx.set_async(False)   # synchronous calls know the failing count on the host: large failing sets take the staged decoder
        l0 = ctx.launch_count
        ctx.robust_interpolate_batch(ids, bad, n, d, t, out=(co, sec, path, fl))
        launches = ctx.launch_count - l0
        s = timed(torch, stream, lambda: ctx.robust_interpolate_batch(ids, bad, n, d, t, out=(co, sec, path, fl)), reps=2)
        ctx.set_async(True)
        want_flags = torch.stack([(mask[:, :64].to(torch.int64) * weights).sum(dim=1), (mask[:, 64:].to(torch.int64) * weights).sum(dim=1)], dim=1)
        checks = {"coeffs_equal_original": bool(torch.equal(co, coeffs)), "flags_equal_injected_error_positions": bool(torch.equal(fl, want_flags)),
                  "path_equals_first_oec_round_with_at_most_r_errors_in_prefix": bool(torch.equal(path, expected_path(mask)))}
        res[kind] = {"ms": 1e3 * s, "codewords_per_s": B / s, "max_path": int(path.max()), "gpu_launches_per_call": int(launches),
                     "analytic_check": {"items": B, "ok": all(checks.values()), **checks}}
        if kind == "uniform":
            sample = (_np(bad[:64]), _np(co[:64]), path[:64].cpu().numpy(), _np(fl[:64]))
        del bad, mask
    out = {"workload": "robust_interpolate_batch n=128, t=42, d=42: clean / e~U{0..42} errors at uniform positions / exactly 42 errors inside the examined prefix",
           "codewords": B, **res}
    if cm is not None and sample is not None:
        ref = cm.robust_interpolate_batch(ids, sample[0], n, d, t, threads=cm.max_threads())
        pok = ref["rc"] == 0 and np.array_equal(sample[1], ref["coeffs"]) and np.array_equal(sample[2], ref["path"]) and np.array_equal(sample[3], ref["flags"][:, :2])
        out["parity_sample"] = {"items": 64, "ok": bool(pok), "what": "coefficients, OEC round and flags of the first 64 'uniform' codewords == C oracle (reference-faithful OEC + Gao, ~1 core-second per codeword)"}
    return out


# ---- BASELINE configs[4]: RanSha + DouSha + RanDouSha + Beaver triple generation, one party's field work for T triples
def leg_c5(torch, hb, ctx, stream, dev, cm, log2, rank):
    n, t = N_PARTIES, T_FAULTS
    T = 1 << log2
    ids = np.arange(n)
    E = lambda *shape: torch.empty(tuple(shape) + (4,), dtype=torch.int64, device=dev)
    I32 = lambda k: torch.empty((k,), dtype=torch.int32, device=dev)
    cols_rs = -(-2 * T // (n - 2 * t))      # RanSha columns: n-2t outputs each, 2 random shares per triple
    cols_ds = -(-T // (t + 1))              # DouSha / RanDouSha columns: t+1 outputs each
    groups = -(-T // (2 * t + 1))           # triple groups of 2t+1 (one batch-recon chunk each)
    seed = 0x5EED0500 + 16 * rank
    # synthetic inputs (consistent sharings so that every check takes the honest path)
    c_t = random_fr_device(torch, (cols_rs, t + 1), seed + 1, dev); sh_t = E(cols_rs, n)
    c_d = random_fr_device(torch, (cols_ds, t + 1), seed + 2, dev); c_d2 = random_fr_device(torch, (cols_ds, 2 * t + 1), seed + 3, dev)
    c_d2[:, 0] = c_d[:, 0]
    sh_d, sh_d2 = E(cols_ds, n), E(cols_ds, n)
    recv = random_fr_device(torch, (cols_rs, n), seed + 4, dev); mix = E(cols_rs, n)           # shares received from the n dealers
    recv_d, mix_d, mix_d2 = random_fr_device(torch, (cols_ds, n), seed + 5, dev), E(cols_ds, n), E(cols_ds, n)
    aS, bS, r2S, rtS = (random_fr_device(torch, (T,), seed + s, dev) for s in (6, 7, 8, 9))    # own shares of a, b, r_2t, r_t
    masked, cS = E(T), E(T)
    grp = random_fr_device(torch, (groups, 2 * t + 1), seed + 10, dev)                          # opened values a*b - r per group
    y_enc = E(n, groups)
    y_all = E(groups, n)
    ctx.compute_shares_batch(grp, n, out=y_all)                                                  # what the n parties would send (degree 2t)
    y_sm = y_all.permute(1, 0, 2).contiguous()
    sec1, p1 = E(groups), I32(groups)
    co2, p2 = E(groups, 2 * t + 1), I32(groups)
    ver_co, ver_sec, ver_p = E(cols_rs, t + 1), E(cols_rs), I32(cols_rs)
    chk_co, chk_sec, chk_st, chk_co2, chk_sec2, chk_st2 = E(cols_ds, t + 1), E(cols_ds), I32(cols_ds), E(cols_ds, 2 * t + 1), E(cols_ds), I32(cols_ds)
    phases = {
        "ransha_deal (K1 d=t, 1 secret/column)": lambda: ctx.compute_shares_batch(c_t, n, out=sh_t),
        "ransha_mix (K2 64x64 per column)": lambda: ctx.apply_vandermonde_batch(recv, n, out=mix),
        "ransha_verify (robust recover of one opened row per column, all n shares)": lambda: ctx.robust_interpolate_batch(ids, sh_t, n, t, t, out=(ver_co, ver_sec, ver_p, None)),
        "dousha_deal (K1 d=t and d=2t per column)": lambda: (ctx.compute_shares_batch(c_d, n, out=sh_d), ctx.compute_shares_batch(c_d2, n, out=sh_d2)),
        "randousha_mix (2x K2 64x64 per column)": lambda: (ctx.apply_vandermonde_batch(recv_d, n, out=mix_d), ctx.apply_vandermonde_batch(sh_d2, n, out=mix_d2)),
        "randousha_check (NonRobust recover deg t and 2t, all n shares)": lambda: (ctx.nonrobust_recover_batch(ids, sh_d, n, t, out=(chk_co, chk_sec, chk_st)),
                                                                                 ctx.nonrobust_recover_batch(ids, sh_d2, n, 2 * t, out=(chk_co2, chk_sec2, chk_st2))),
        "triple_mask (K5 fused: a*b - r_2t per triple, one pass)": lambda: ctx.share_algebra_fused(0, (aS, bS, r2S), out=masked),
        "triple_open_encode (K2 64x43 per group, recipient-major)": lambda: ctx.apply_vandermonde_batch(grp, n, recipient_major=True, out=y_enc),
        "triple_open_round1 (batch_recover_secrets d=2t, 64 senders)": lambda: ctx.batch_recover_secrets(ids, y_sm, n, 2 * t, t, out=(sec1, p1)),
        "triple_open_round2 (batch_recover d=2t, 64 senders)": lambda: ctx.batch_recover(ids, y_sm, n, 2 * t, t, out=(co2, p2, None)),
        "triple_finish (K5: r_t + opened)": lambda: ctx.elementwise(0, rtS, masked, out=cS),
    }
    ctx.set_async(True)
    res, total = {}, 0.0
    for name, fn in phases.items():
        s = timed(torch, stream, fn)
        res[name] = round(1e3 * s, 4)
        total += s
    ok = ctx.synchronize() == 0 and bool(torch.equal(co2, grp)) and bool(torch.equal(sec1, grp[:, 0])) and bool(torch.equal(chk_sec, c_d[:, 0])) \
        and bool(torch.equal(chk_sec2, c_d[:, 0])) and int(chk_st.min()) == t and int(chk_st2.min()) == 2 * t and bool(torch.equal(ver_sec, c_t[:, 0])) \
        and bool(torch.equal(y_enc, y_sm))
    out = {"workload": "one party's field work of RanSha + DouSha + RanDouSha + Beaver triple generation (n=64, t=21), device-resident, per rank",
           "triples_per_rank": T, "ransha_columns": cols_rs, "dousha_columns": cols_ds, "groups": groups, "phase_ms": res, "total_ms": round(1e3 * total, 3),
           "triples_per_s_per_gpu": T / total, "self_consistent": ok, "_total_s": total}
    if cm is not None and rank == 0:
        K, th = 256, cm.max_threads()
        rc1, w_sh = cm.compute_shares(_np(c_t[:K]), n, threads=th)
        rc2, w_mix = cm.apply_vandermonde(_np(recv[:K]), n, threads=th)
        rc3, w_mix2 = cm.apply_vandermonde(_np(sh_d2[:K]), n, threads=th)
        w_ver = cm.robust_interpolate_batch(ids, _np(sh_t[:K]), n, t, t, threads=th)
        rc4, w_enc = cm.apply_vandermonde(_np(grp[:K]), n, recipient_major=True, threads=th)
        w_r2 = cm.batch_recover_secret(ids, _np(y_sm[:, :K].contiguous()), n, 2 * t, t, threads=th)
        Ke = 4096
        _, w_prod = cm.elementwise(2, _np(aS[:Ke]), _np(bS[:Ke]), threads=th)
        _, w_mask = cm.elementwise(1, w_prod, _np(r2S[:Ke]), threads=th)
        _, w_c = cm.elementwise(0, _np(rtS[:Ke]), w_mask, threads=th)
        nr = [cm.nonrobust_recover_secret(ids, _np(sh_d2[b]), n, 2 * t) for b in range(8)]
        pok = (rc1 == rc2 == rc3 == rc4 == 0 and np.array_equal(_np(sh_t[:K]), w_sh) and np.array_equal(_np(mix[:K]), w_mix) and np.array_equal(_np(mix_d2[:K]), w_mix2)
               and np.array_equal(_np(ver_co[:K]), w_ver["coeffs"]) and np.array_equal(_np(y_enc[:, :K].contiguous()), w_enc)
               and np.array_equal(_np(co2[:K]), w_r2["coeffs"]) and np.array_equal(_np(sec1[:K]), w_r2["coeffs"][:, 0])
               and np.array_equal(_np(masked[:Ke]), w_mask) and np.array_equal(_np(cS[:Ke]), w_c)
               and all(r["rc"] == 0 and np.array_equal(_np(chk_co2[b]), r["coeffs"]) for b, r in enumerate(nr)))
        out["parity_sample"] = {"items": K, "ok": bool(pok), "what": "first 256 columns / groups of every K1/K2/K3 phase, first 4096 triples of the K5 phases, 8 degree-2t checks == C oracle"}
    return out


def leg_sampler(torch, hb, ctx, stream, dev, cm):
    """the polynomials of 2^20 sharings (n=64, t=21) drawn on the device from a 32-byte seed (StdRng = ChaCha12 + Fp::rand
    rejection sampling, bit-compatible with the reference's generator) instead of uploaded: 704 bytes per secret stay off PCIe."""
    B, d = 1 << 20, DEG
    seed = bytes(range(32))
    coeffs = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
    ctx.set_async(False)
    ctx.sample_polynomials(seed, B, d, out=coeffs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.sample_polynomials(seed, B, d, out=coeffs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    ctx.set_async(True)
    out = {"workload": "hbmpc_sample_polynomials: 2^20 sharings of degree 21 (d+2 = 23 Fp::rand draws each, ~25.4 ChaCha12 candidates), device-resident output",
           "sharings": B, "ms": 1e3 * dt, "field_draws_per_s": B * (d + 2) / dt, "upload_bytes_saved_per_secret": (d + 1) * 32}
    if cm is not None:
        from oracle import chacha_fr as cf
        want = cf.sample_polynomials(seed, 64, d)
        got = hb.from_limbs(_np(coeffs[:64]))
        out["parity_sample"] = {"items": 64, "ok": bool(got == want), "what": "first 64 coefficient vectors == oracle/chacha_fr.py (ChaCha12 pinned by RFC 7539 / strombergson vectors; arkworks byte parity unpinned)"}
    return out


def leg_goldilocks(torch, hb, ctx, dev, cm):
    """share generation + optimistic batch recovery over GoldilocksField (Fp64), n=64, t=21, 2^20 secrets, device-resident
    (hbmpc_gl_* take device pointers; 8-byte elements)."""
    B, n, d, t = 1 << 20, N_PARTIES, DEG, T_FAULTS
    P = 0xFFFFFFFF00000001
    g = torch.Generator(device=dev)
    g.manual_seed(0x601D)
    coeffs = torch.randint(0, 1 << 62, (B, d + 1), dtype=torch.int64, device=dev, generator=g)   # < p
    shares = torch.empty((B, n), dtype=torch.int64, device=dev)
    rec = torch.empty((B, d + 1), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    ids = np.arange(n, dtype=np.uint64)
    lib, h = ctx.lib, ctx.h
    torch.cuda.synchronize()

    def gen():
        assert lib.hbmpc_gl_compute_shares_batch(h, n, d, B, coeffs.data_ptr(), shares.data_ptr()) == 0

    gen()
    evals = shares.t().contiguous()
    torch.cuda.synchronize()

    def recon():
        assert lib.hbmpc_gl_batch_recover(h, n, d, t, n, ids.ctypes.data, B, evals.data_ptr(), rec.data_ptr(), None, path.data_ptr()) == 0

    recon()

    def wall(fn, reps=3):
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps   # the calls are synchronous

    tg, tr = wall(gen), wall(recon)
    ok = bool(torch.equal(rec, coeffs)) and not bool(path.any())
    out = {"workload": "GoldilocksField (p = 2^64 - 2^32 + 1): hbmpc_gl_compute_shares_batch + hbmpc_gl_batch_recover, n=64, t=21, 2^20 secrets, device-resident",
           "secrets": B, "gen_ms": 1e3 * tg, "recon_ms": 1e3 * tr, "shares_per_s": 2 * B * n / (tg + tr), "roundtrip_ok": ok,
           "hbm_gbs": {"gen": B * (d + 1 + n) * 8 / tg / 1e9, "recon": B * (d + t + 1 + d + 1) * 8 / tr / 1e9}}
    if cm is not None:
        from oracle import goldilocks as glo
        k = 64
        cs = coeffs[:k].cpu().numpy().astype(np.uint64)
        sh = shares[:k].cpu().numpy().astype(np.uint64)
        out["parity_sample"] = {"items": k, "ok": bool(all([int(v) for v in sh[b]] == glo.compute_shares([int(v) for v in cs[b]], n) for b in range(k))),
                                "what": "shares of the first 64 secrets == oracle/goldilocks.py"}
    return out


def _sync_ok(c, what):
    rc = c.synchronize()
    if rc != 0:
        raise RuntimeError("%s: hbmpc_ctx_synchronize -> %d (%s)" % (what, rc, c.last_error()))


def leg_group(torch, hb, n_dev, log2_total, cm):
    """One process, n_dev GPUs: member contexts of an hbmpc_group driven from this process."""
    ids = np.arange(N_PARTIES)
    total = 1 << log2_total
    grp = hb.Group(list(range(n_dev)))
    members = []
    for g in range(n_dev):
        lo, hi = grp.shard_range(total, g)
        dv = torch.device("cuda", g)
        with torch.cuda.device(dv):
            c = hb.Context(g)
            c.set_async(True)
            co = random_fr_device(torch, (hi - lo, M), 0x5EED0600 + g, dv)
            sh = torch.empty((hi - lo, N_PARTIES, 4), dtype=torch.int64, device=dv)
            torch.cuda.synchronize(dv)   # the context's stream is not ordered against torch's: the inputs must be complete
            c.compute_shares_batch(co, N_PARTIES, out=sh)
            torch.cuda.synchronize(dv)
            _sync_ok(c, 'group member %d set-up' % g)
            ev = sh.permute(1, 0, 2).contiguous()
            rec = torch.empty((hi - lo, M, 4), dtype=torch.int64, device=dv)
            pth = torch.empty((hi - lo,), dtype=torch.int32, device=dv)
            torch.cuda.synchronize(dv)
            members.append((c, co, sh, ev, rec, pth))

    def step():
        for c, co, sh, ev, rec, pth in members:     # enqueue-only calls: all devices run concurrently
            c.compute_shares_batch(co, N_PARTIES, out=sh)
            c.batch_recover(ids, ev, N_PARTIES, DEG, T_FAULTS, out=(rec, pth, None))
        for i, (c, *_) in enumerate(members):
            _sync_ok(c, 'group member %d step' % i)

    for _ in range(3):
        step()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    ok = all(bool(torch.equal(rec, co)) for _, co, _, _, rec, _ in members)
    out = {"workload": "ONE process, one member context per GPU: 2^%d secrets in total in contiguous ranges, gen + recon per step, device-resident, wall clock around enqueue-on-all + synchronize-all" % log2_total,
           "devices": n_dev, "secrets_total": total, "ms_per_step": 1e3 * dt, "shares_per_s": 2 * total * N_PARTIES / dt, "roundtrip_ok": ok}
    # host-buffer group calls (hbmpc_group_*: internal host threads, PCIe-bound) on 2^20 secrets in total
    Bh = 1 << 20
    hc = _np(members[0][1][: min(Bh, members[0][1].shape[0])])
    Bh = hc.shape[0]
    hs = np.zeros((Bh, N_PARTIES, 4), dtype=np.uint64)
    grp.compute_shares_batch(hc, N_PARTIES, out=hs)
    he = np.ascontiguousarray(hs.transpose(1, 0, 2))
    t0 = time.perf_counter()
    grp.compute_shares_batch(hc, N_PARTIES, out=hs)
    rc, hr, hp, _ = grp.batch_recover(ids, he, N_PARTIES, DEG, T_FAULTS)
    dth = time.perf_counter() - t0
    out["host_buffer_group_calls"] = {"secrets": Bh, "ms": 1e3 * dth, "shares_per_s": 2 * Bh * N_PARTIES / dth, "ok": bool(rc == 0 and np.array_equal(hr, hc) and not hp.any()),
                                      "note": "pageable numpy buffers, one call each of hbmpc_group_compute_shares_batch / hbmpc_group_batch_recover"}
    if cm is not None:
        Ks = 2048
        rcs, want = cm.compute_shares(hc[:Ks], N_PARTIES, threads=cm.max_threads())
        out["parity_sample"] = {"items": Ks, "ok": bool(rcs == 0 and np.array_equal(hs[:Ks], want)), "what": "group share generation of the first 2048 secrets == C oracle"}
    for c, *_ in members:
        c.close()
    grp.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    hb = importlib.import_module("mpc-protocols_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n_gpus = world
    configs_on = not (args.no_configs or args.no_robust_leg)

    ctx = hb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_async(True)
    B = 1 << args.log2_batch
    if args.scaling == "strong":
        sh = importlib.import_module("mpc-protocols_b200.sharding")
        lo_, hi_ = sh.shard_range(B, world, rank)
        B = hi_ - lo_
    ids = np.arange(N_PARTIES)

    def max_over_ranks(vals):
        tt = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.tolist()

    def sum_over_ranks(vals):
        tt = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return tt.tolist()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs, resident in HBM before the timed region
    coeffs = random_fr_device(torch, (B, M), 0x5EED0003 + rank, dev)
    shares = torch.empty((B, N_PARTIES, 4), dtype=torch.int64, device=dev)
    ctx.compute_shares_batch(coeffs, N_PARTIES, out=shares)
    evals = shares.permute(1, 0, 2).contiguous()  # [64 senders][B]: how the per-sender messages arrive
    rec = torch.empty((B, M, 4), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    assert ctx.synchronize() == 0

    def step(ev, c=coeffs, s=shares, e=evals, r=rec, p=path):
        ev[0].record(stream)
        ctx.compute_shares_batch(c, N_PARTIES, out=s)
        ev[1].record(stream)
        ctx.batch_recover(ids, e, N_PARTIES, DEG, T_FAULTS, out=(r, p, None))
        ev[2].record(stream)

    mk = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(max(args.warmup, 0)):
        step(mk())
    assert ctx.synchronize() == 0
    assert torch.equal(rec, coeffs) and not bool(path.any()), "reconstruction != original coefficients"

    imad_peak = ctx.measure_imad_peak(0)[0] * 1e9  # thread-level IMAD/s, measured in this run on this GPU
    imad_wide_peak = ctx.measure_imad_peak(1)[0] * 1e9  # IMAD.WIDE.U32(.X) SASS instructions/s (one per mad.lo/madc.hi pair of the carry chains)

    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    events = [mk() for _ in range(args.steps)]
    for ev in events:
        step(ev)
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    assert ctx.synchronize() == 0
    t_gen = sum(ev[0].elapsed_time(ev[1]) for ev in events) * 1e-3
    t_rec = sum(ev[1].elapsed_time(ev[2]) for ev in events) * 1e-3
    t_tot = sum(ev[0].elapsed_time(ev[2]) for ev in events) * 1e-3
    assert torch.equal(rec, coeffs) and not bool(path.any())
    t_tot, t_gen, t_rec = max_over_ranks([t_tot, t_gen, t_rec])
    B_all = int(sum_over_ranks([float(B)])[0])

    # ---- CPU oracle (rank 0 only): the checker of every parity sample below and, at N = 1, the timed cpu_baseline
    cm = load_oracle_native() if (rank == 0 and not args.no_cpu_baseline) else None

    configs = {}
    t_dense = None
    if configs_on:
        # the 43-sender first call (every rank measures; rank 0 reports and checks against the oracle)
        fc, t_dense = leg_c3_first_call(torch, hb, ctx, stream, dev, cm, coeffs, evals, rec, path, rank)
        fc_ms = max_over_ranks([fc["coeffs_ms"], fc["secrets_ms"], fc["coeffs_flags_ms"], fc["dense_flags_ms"]])
        fc["coeffs_ms"], fc["secrets_ms"], fc["coeffs_flags_ms"], fc["dense_flags_ms"] = fc_ms
        t_dense = fc_ms[3] * 1e-3
        configs["c3_first_call"] = fc
        # restore the headline outputs for the parity sample below
        ctx.batch_recover(ids, evals, N_PARTIES, DEG, T_FAULTS, out=(rec, path, None))
        assert ctx.synchronize() == 0

    # ---- BASELINE configs[2] as stated: 2^22 secrets in total, sharded over the ranks (strong scaling), device-resident
    if world > 1 and args.scaling == "weak" and configs_on:
        Bs = (1 << args.log2_batch) // world
        cs_, ss_, es_, rs_, ps_ = coeffs[:Bs], shares[:Bs], evals[:, :Bs].contiguous(), rec[:Bs], path[:Bs]
        for _ in range(3):
            step(mk(), cs_, ss_, es_, rs_, ps_)
        barrier()
        evs = [mk() for _ in range(10)]
        for ev in evs:
            step(ev, cs_, ss_, es_, rs_, ps_)
        barrier()
        ts = max_over_ranks([sum(ev[0].elapsed_time(ev[2]) for ev in evs) * 1e-3])[0]
        assert ctx.synchronize() == 0 and torch.equal(rs_, cs_)
        configs["c3_strong"] = {"workload": "configs[2] as stated: 2^%d secrets in TOTAL, contiguous ranges over the ranks, gen + recon per step" % args.log2_batch,
                                "secrets_total": Bs * world, "secrets_per_rank": Bs, "ms_per_step": 1e3 * ts / 10, "shares_per_s": 10 * 2 * Bs * world * N_PARTIES / ts}

    # ---- end-to-end leg: same calls with HOST (pinned) buffers, copies inside the timed region
    Be = min(1 << args.log2_e2e_batch, B)
    h_coeffs = torch.empty((Be, M, 4), dtype=torch.int64).pin_memory()
    h_coeffs.copy_(coeffs[:Be].cpu())
    h_shares = torch.empty((Be, N_PARTIES, 4), dtype=torch.int64).pin_memory()
    h_evals = torch.empty((N_PARTIES, Be, 4), dtype=torch.int64).pin_memory()
    h_evals.copy_(evals[:, :Be].cpu())
    h_rec = torch.empty((Be, M, 4), dtype=torch.int64).pin_memory()
    h_path = torch.empty((Be,), dtype=torch.int32).pin_memory()
    # asynchronous mode: both calls only enqueue (on two different sets of internal streams) and hbmpc_ctx_synchronize completes them,
    # so the download-bound share generation and the upload-bound recovery use both PCIe directions at the same time
    ctx.set_async(True)
    np_c, np_s, np_e, np_r, np_p = (x.numpy().view(np.uint64) if x.dtype == torch.int64 else x.numpy() for x in (h_coeffs, h_shares, h_evals, h_rec, h_path))

    def e2e_step_coeffs():
        ctx.compute_shares_batch(np_c, N_PARTIES, out=np_s)
        ctx.batch_recover(ids, np_e, N_PARTIES, DEG, T_FAULTS, out=(np_r, np_p, None))
        assert ctx.synchronize() == 0

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step_coeffs()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step_coeffs()
    torch.cuda.synchronize()
    t_e2e_c = time.perf_counter() - t0
    assert np.array_equal(np_r, np_c) and not np_p.any()
    t_e2e_c = max_over_ranks([t_e2e_c])[0]
    h2d_c = Be * M * 32 + Be * N_PARTIES * 32  # coefficients and all 64 sender vectors in (asynchronous calls upload every supplied sender)
    d2h = Be * N_PARTIES * 32 + Be * M * 32 + Be * 4

    # the headline e2e step: share generation with the reference's own argument meaning -- compute_shares(secret, n, degree, rng): the
    # polynomial is drawn inside the call (robust_interpolate.rs:68-69), here on the device from a StdRng seed, so the dealer uploads 32 B
    # per secret instead of 704 -- followed by the same batch_recover call on the 64 sender vectors of those sharings
    seed = bytes((17 * i + rank + 1) & 0xFF for i in range(32))
    h_sec = torch.empty((Be, 4), dtype=torch.int64).pin_memory()
    h_sec.copy_(coeffs[:Be, 0].cpu())
    np_sec = h_sec.numpy().view(np.uint64)
    ctx.set_async(False)
    ctx.share_secrets_batch(seed, np_sec, N_PARTIES, DEG, out=np_s, coeffs_out=np_c)   # np_c: the polynomials the generator drew
    np_e[...] = np_s.transpose(1, 0, 2)
    ctx.set_async(True)

    def e2e_step():
        ctx.share_secrets_batch(seed, np_sec, N_PARTIES, DEG, out=np_s)
        ctx.batch_recover(ids, np_e, N_PARTIES, DEG, T_FAULTS, out=(np_r, np_p, None))
        assert ctx.synchronize() == 0

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(np_r, np_c) and np.array_equal(np_r[:, 0], np_sec) and not np_p.any()
    assert np.array_equal(np_e, np_s.transpose(1, 0, 2)), "seeded share generation is not reproducible"
    t_e2e = max_over_ranks([t_e2e])[0]
    h2d = Be * 32 + 32 + Be * N_PARTIES * 32
    ctx.set_async(True)
    del h_coeffs, h_shares, h_evals, h_rec, h_path, h_sec

    # ---- the box's pinned-copy rates, every rank AT THE SAME TIME (the roofline of the e2e leg: with N ranks they share the host bridge)
    pcie = pcie_probe(torch, dev, barrier)
    pcie_bound_s = max(h2d, d2h) / (pcie["both_each_GBs"] * 1e9)   # both directions busy: the slower one bounds the step
    pcie_all = max_over_ranks([1.0 / pcie["h2d_GBs"], 1.0 / pcie["d2h_GBs"], 1.0 / pcie["both_each_GBs"]])
    pcie_min = {"h2d_GBs": 1.0 / pcie_all[0], "d2h_GBs": 1.0 / pcie_all[1], "both_each_GBs": 1.0 / pcie_all[2]}
    pcie_bound_s = max_over_ranks([pcie_bound_s])[0]

    # ---- NCCL gather of the reconstructed secrets (the only collective; outside the hot path)
    gather_ms = None
    if world > 1 and args.scaling == "weak":
        sh = importlib.import_module("mpc-protocols_b200.sharding")
        lo, hi = sh.shard_range(world * B, world, rank)      # this rank's contiguous range of the global batch
        assert hi - lo == B
        secrets = rec[:, 0, :].contiguous()
        sh.gather_shards(secrets[: 1024].contiguous(), world * 1024)   # communicator set-up is not part of the gather time
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allsec = sh.gather_shards(secrets, world * B)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert torch.equal(allsec[lo:hi], coeffs[:, 0, :])
        del allsec

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores; its outputs are the parity
    # sample of the headline (the GPU shares / coefficients of the same first 2^17 secrets must be identical)
    cpu, parity = None, None
    if cm is not None and n_gpus == 1:
        threads = cm.max_threads()
        Bc = min(1 << args.cpu_log2_batch, B)
        cc = _np(coeffs[:Bc])
        cpu_step(cm, cc[: 1 << 12], threads)
        keep = {}
        v_all, reps, tcpu = cpu_rate(cm, cc, threads, 4.0, keep)
        v_one, _, _ = cpu_rate(cm, cc[: 1 << 13], 1, 1.0)
        cpu = {"value": v_all, "unit": "shares/s", "cores": threads, "kind": "port", "value_1_thread": v_one,
               "sample": f"{reps} x 2^{args.cpu_log2_batch} secrets of the same workload (FFT compute_shares + batch_recover_secret), C oracle port -march=native, {threads} threads, {tcpu:.1f} s; value_1_thread: the same on one core (the reference node is single-threaded)"}
        pok = np.array_equal(_np(shares[:Bc]), keep["shares"]) and np.array_equal(_np(rec[:Bc]), keep["coeffs"]) and np.array_equal(path[:Bc].cpu().numpy(), keep["path"])
        parity = {"items": Bc, "ok": bool(pok), "what": "GPU shares[B][64], coefficients and path of the first 2^%d secrets of the timed batch == the oracle outputs computed by the cpu_baseline leg" % args.cpu_log2_batch}
    elif cm is not None:
        Bc = 4096
        rc, want = cm.compute_shares(_np(coeffs[:Bc]), N_PARTIES, threads=cm.max_threads())
        parity = {"items": Bc, "ok": bool(rc == 0 and np.array_equal(_np(shares[:Bc]), want) and np.array_equal(_np(rec[:Bc]), _np(coeffs[:Bc]))),
                  "what": "rank 0: GPU shares of the first 4096 secrets == C oracle, coefficients == inputs"}

    # ---- the other BASELINE configs (outside the timed region).  c2 / c4: rank 0 at N = 1; c5: every rank at every N
    if configs_on:
        del shares, evals
        torch.cuda.empty_cache()
        if n_gpus == 1:
            configs["c2"] = leg_c2(torch, hb, ctx, stream, dev, cm, args.log2_c2)
            configs["c4"] = leg_c4(torch, hb, ctx, stream, dev, cm, args.log2_c4)
            torch.cuda.empty_cache()
        if n_gpus == 1:
            configs["n4_device_sampling"] = leg_sampler(torch, hb, ctx, stream, dev, cm)
            configs["n4_goldilocks"] = leg_goldilocks(torch, hb, ctx, dev, cm)
        c5 = leg_c5(torch, hb, ctx, stream, dev, cm, args.log2_c5, rank)
        tot5 = max_over_ranks([c5.pop("_total_s")])[0]
        ok5 = min(max_over_ranks([0.0 if c5["self_consistent"] else 1.0])) == 0.0
        c5.update({"ranks": world, "triples_total": world * c5["triples_per_rank"], "total_ms_max_over_ranks": 1e3 * tot5,
                   "triples_per_s": world * c5["triples_per_rank"] / tot5, "self_consistent_all_ranks": ok5})
        configs["c5"] = c5

    # ---- single-process multi-GPU (rank 0 drives one member context per GPU of the node: what a one-process reference party would
    # do): configs[2] as stated, 2^22 secrets in total in contiguous ranges over the devices, device-resident, enqueue on every member
    # then wait for all; and the same through the group's host-buffer entry points
    if world > 1 and configs_on:
        barrier()
        cpu_group = dist.new_group(backend="gloo")   # the other ranks must wait on the CPU: an NCCL barrier would spin on their GPUs
        if rank == 0:
            try:
                configs["c3_group_single_process"] = leg_group(torch, hb, world, args.log2_batch, cm)
            except Exception as exc:   # never lose the headline line to this leg
                configs["c3_group_single_process"] = {"error": repr(exc)}
        dist.barrier(group=cpu_group)

    if rank == 0:
        shares_per_step = 2 * B_all * N_PARTIES
        value = args.steps * shares_per_step / t_tot
        gen_launch_s = t_gen / args.steps
        rec_launch_s = t_rec / args.steps
        hbm = measured_hbm()
        ex = executed_counts()
        traffic = ncu_traffic()

        def roof(kernel_key, label, alg_modmul, secs, nbytes, tkey, units=B):
            """frac = EXECUTED 32x32->64 multiply-adds (IMAD.WIDE SASS instructions, counted by ncu per item: profiles/r02_executed_counts.json)
            per launch / CUDA-event launch time / the IMAD.WIDE rate measured in this run.  alg_speedup = dense reference-faithful count / executed."""
            e = ex.get(kernel_key, {})
            wide = e.get("imad_wide_thread_inst_per_item")
            o = {"kernel": label, "bound": "int32 multiplier pipe (IMAD.WIDE.U32: 32x32->64 multiply-add)", "unit": "T IMAD.WIDE/s",
                 "peak": imad_wide_peak / 1e12, "peak_how": "IMAD.WIDE.U32.X carry-chain probe measured in this run (hbmpc_measure_imad_peak variant 1)",
                 "achieved": (wide * units / secs / 1e12) if wide else None, "frac": (wide * units / secs / imad_wide_peak) if wide else None,
                 "executed_imad_wide_per_item": wide, "executed_source": e.get("source"), "launch_ms": 1e3 * secs, "items_per_launch": units,
                 "alg_modmul_per_item": alg_modmul, "alg_imad_per_s_T": alg_modmul * IMAD_PER_MODMUL * units / secs / 1e12,
                 "alg_frac_of_imad_peak": alg_modmul * IMAD_PER_MODMUL * units / secs / imad_peak, "imad_peak_T": imad_peak / 1e12,
                 "alg_speedup": (alg_modmul * IMAD_PER_MODMUL / 2 / wide) if wide else None,
                 "hbm_gbs": nbytes / secs / 1e9, "hbm_frac": nbytes / secs / 1e9 / hbm, "algorithmic_bytes": nbytes,
                 "traffic": (traffic[tkey] * units / (1 << 20)) if tkey in traffic else None}
            return o

        line = {
            "metric": METRIC, "value": value, "unit": "shares/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(args),
            "breakdown": {"gen_ms": 1e3 * t_gen / args.steps, "recon_ms": 1e3 * t_rec / args.steps,
                          "gen_shares_per_s": args.steps * B_all * N_PARTIES / t_gen, "recon_shares_per_s": args.steps * B_all * N_PARTIES / t_rec,
                          "alg_modmul_per_s": args.steps * B_all * (ALG_MODMUL_GEN + ALG_MODMUL_REC) / t_tot, "gather_ms": gather_ms,
                          "recon_note": "recon_ms: all 64 senders supplied -> inverse NTT + degree check (bit-identical; items that fail fall back to the dense check and the decoder); configs.c3_first_call: the 43-sender call"},
            "roofline": roof("recon", "K3 batch_recover launch, all 64 senders: inverse 64-point NTT + degree check per chunk", ALG_MODMUL_REC, rec_launch_s,
                             B * (N_PARTIES * 32 + M * 32 + 5), "ntt_inv"),
            "roofline_gen": roof("gen", "K1 compute_shares launch: zero-padded 64-point NTT per secret", ALG_MODMUL_GEN, gen_launch_s, B * BYTES_GEN, "ntt_fwd"),
            "parity_sample": parity,
            "configs": configs,
            "cpu_baseline": cpu,
            "e2e": {"value": n_gpus * e2e_steps * 2 * Be * N_PARTIES / t_e2e, "unit": "shares/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps, "host_buffers": "pinned",
                    "mode": "hbmpc_share_secrets_batch (secrets + StdRng seed in, polynomials drawn on the device, shares out) + hbmpc_batch_recover (64 sender vectors in, coefficients out); asynchronous host-buffer calls (enqueue on two lane sets, hbmpc_ctx_synchronize per step)",
                    "coeffs_uploaded": {"value": n_gpus * e2e_steps * 2 * Be * N_PARTIES / t_e2e_c, "ms_per_step": 1e3 * t_e2e_c / e2e_steps, "h2d_bytes_per_step": h2d_c,
                                        "what": "same step with hbmpc_compute_shares_batch on host coefficients (704 B per secret uploaded): the round-1 e2e shape"},
                    "pcie_gbs": {"h2d": h2d * e2e_steps / t_e2e / 1e9, "d2h": d2h * e2e_steps / t_e2e / 1e9},
                    "pcie_probe_slowest_rank": pcie_min, "pcie_bound_ms": 1e3 * pcie_bound_s, "pcie_frac": pcie_bound_s / (t_e2e / e2e_steps),
                    "pcie_note": "pcie_probe: pinned 256 MiB copies on every rank concurrently (H2D alone, D2H alone, both at once); pcie_bound = max(h2d, d2h bytes per step) / the bidirectional per-direction rate; pcie_frac = bound / measured step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if t_dense is not None:
            line["roofline_dense"] = roof("dense", "matvec_kernel (K3 dense check with flags, 43 senders: 43x22 check+coefficient matrix per chunk; the route of every chunk with a disagreeing share)",
                                          ALG_MODMUL_REC, t_dense, B * (BYTES_REC + 8), "matvec")
        if "c4" in configs and "k4" in ex:
            # K4: the staged decoder's dominant kernel is bm_segment_kernel; executed multiply-adds of the WHOLE call per codeword
            s4 = configs["c4"]["uniform"]["ms"] * 1e-3
            line["roofline_k4"] = roof("k4", "robust_interpolate_batch n=128,t=42, e~U{0..42}: all kernels of the call (dominant: bm_segment_kernel)", 13400, s4,
                                       configs["c4"]["codewords"] * (128 * 32 + 43 * 32 + 32 + 4 + 16), "k4", units=configs["c4"]["codewords"])
            if "k4_per_2p17" in traffic:   # the staged decoder's state lives in HBM between its stages: traffic is far above the call's algorithmic bytes
                line["roofline_k4"]["traffic"] = traffic["k4_per_2p17"] * configs["c4"]["codewords"] / (1 << 17)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pcie_probe(torch, dev, barrier, nbytes=1 << 28, reps=4):
    """Pinned-memory copy bandwidth seen by this rank while every other rank does the same: H2D, D2H, and both at once on two streams."""
    h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(n):
            if up:
                with torch.cuda.stream(s1):
                    d1.copy_(h1, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return n * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    run(True, True, 1)
    return {"h2d_GBs": run(True, False, reps), "d2h_GBs": run(False, True, reps), "both_each_GBs": run(True, True, reps)}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch at B = 2^20 from the committed ncu --set full capture (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_per_2p20.json")))
    except Exception:
        return {}


def executed_counts():
    """Executed IMAD.WIDE thread-instructions per item of each headline kernel, summed over the SASS lines of the committed ncu source
    page (tools/ncu_exec_counts.py -> profiles/r02_executed_counts.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_executed_counts.json")))
    except Exception:
        return {}


def measured_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
