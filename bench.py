#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 finite-field hot path (see DESIGN.md "Measurement").

Metric (BASELINE.json): Fr shares generated + reconstructed per second at n=64, t=21.
One step = one pass of the hot path over one batch of B synthetic secrets on every rank:
    gen   : hbmpc_compute_shares_batch   coeffs[B][22] -> shares[B][64]            (K1, RobustShare::compute_shares)
    recon : hbmpc_batch_recover          evals[64][B] (all 64 senders) -> coeffs[B][22], path[B]   (K3, batch_recover_secret)
    value = N * (B*64 + B*64) / max-over-ranks(t_gen + t_rec)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2-batch 22]
For N > 1 launch under torchrun (one rank per GPU); the batch of independent secrets is sharded (weak scaling: every
rank owns B secrets), no collective on the data path; NCCL only gathers the reconstructed secrets after the timed region.
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
# algorithmic (dense, reference-faithful) modmul per secret and IMAD per modmul
ALG_MODMUL_GEN = N_PARTIES * DEG                      # 1344 (Horner count)
ALG_MODMUL_REC = (DEG + T_FAULTS + 1) * M + M * M     # 946 + 484 = 1430
IMAD_PER_MODMUL = 256
BYTES_GEN = M * 32 + N_PARTIES * 32                   # 704 R + 2048 W
BYTES_REC = (DEG + T_FAULTS + 1) * 32 + M * 32 + 4    # 1376 R + 704 W + path


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-batch", type=int, default=22, help="secrets per rank per step (device-resident leg)")
    ap.add_argument("--log2-e2e-batch", type=int, default=20, help="secrets per rank per step (host-buffer leg)")
    ap.add_argument("--cpu-log2-batch", type=int, default=17, help="secrets per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-robust-leg", action="store_true", help="skip the secondary K4 measurement (BASELINE configs[3] shape)")
    ap.add_argument("--log2-robust", type=int, default=17, help="codewords of the K4 leg (n=128, t=42, e ~ U{0..42} errors each)")
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


def cpu_step(cm, coeffs, threads):
    """One pass of the same hot path on the CPU oracle (FFT share generation + batch_recover_secret)."""
    t0 = time.perf_counter()
    rc, shares = cm.compute_shares(coeffs, N_PARTIES, threads=threads)
    t1 = time.perf_counter()
    evals = np.ascontiguousarray(shares.transpose(1, 0, 2))  # message re-assembly, not timed (device leg has it resident too)
    t2 = time.perf_counter()
    out = cm.batch_recover_secret(np.arange(N_PARTIES), evals, N_PARTIES, DEG, T_FAULTS, threads=threads)
    t3 = time.perf_counter()
    assert rc == 0 and out["rc"] == 0 and np.array_equal(out["coeffs"], coeffs)
    return (t1 - t0) + (t3 - t2)


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
    sample = f"{args.steps} steps x 2^{args.cpu_log2_batch} secrets (n=64,t=21): FFT compute_shares + batch_recover_secret, C oracle port, {threads} threads"
    line = {
        "impl": "reference", "metric": "Fr shares generated+reconstructed per second (n=64,t=21)", "value": value, "unit": "shares/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (Fr, 255-bit Montgomery)", "data": "synthetic",
        "config": workload_config(args.cpu_log2_batch, None),
        "cpu_baseline": {"value": value, "unit": "shares/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "shares/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the Rust reference cannot be built here (no rustc/cargo, arkworks not vendored): this arm times the C oracle port on all host cores",
    }
    print(json.dumps(line))


def workload_config(log2_batch, log2_e2e):
    cfg = {
        "workload": "HoneyBadgerMPC share-gen + batch reconstruction over ark_bls12_381::Fr, n=64, t=21 (BASELINE configs[2] shape)",
        "n": N_PARTIES, "t": T_FAULTS, "degree": DEG, "secrets_per_rank_per_step": 1 << log2_batch,
        "gen": "compute_shares coeffs[B][22] -> shares[B][64]", "recon": "batch_recover evals[64][B] -> coeffs[B][22] (S=64 senders, 43 examined)",
        "l2": "inputs (>= 2.9 GB per kernel) are larger than the 126 MB L2; no explicit flush",
    }
    if log2_e2e is not None:
        cfg["e2e_secrets_per_rank_per_step"] = 1 << log2_e2e
    return cfg


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

    ctx = hb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_async(True)
    B = 1 << args.log2_batch
    ids = np.arange(N_PARTIES)

    # ---- synthetic inputs, resident in HBM before the timed region
    coeffs = random_fr_device(torch, (B, M), 0x5EED0003 + rank, dev)
    shares = torch.empty((B, N_PARTIES, 4), dtype=torch.int64, device=dev)
    ctx.compute_shares_batch(coeffs, N_PARTIES, out=shares)
    evals = shares.permute(1, 0, 2).contiguous()  # [64 senders][B]: how the per-sender messages arrive
    rec = torch.empty((B, M, 4), dtype=torch.int64, device=dev)
    path = torch.empty((B,), dtype=torch.int32, device=dev)
    assert ctx.synchronize() == 0

    def step(ev):
        ev[0].record(stream)
        ctx.compute_shares_batch(coeffs, N_PARTIES, out=shares)
        ev[1].record(stream)
        ctx.batch_recover(ids, evals, N_PARTIES, DEG, T_FAULTS, out=(rec, path, None))
        ev[2].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mk = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(max(args.warmup, 0)):
        step(mk())
    assert ctx.synchronize() == 0
    assert torch.equal(rec, coeffs) and not bool(path.any()), "reconstruction != original coefficients"

    imad_peak = ctx.measure_imad_peak(0)[0] * 1e9  # thread-level IMAD/s, measured in this run on this GPU
    imad_wide_peak = ctx.measure_imad_peak(1)[0] * 1e9

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
    tt = torch.tensor([t_tot, t_gen, t_rec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_tot, t_gen, t_rec = tt.tolist()

    # ---- general path of K3 for comparison: only the first d+t+1 = 43 senders supplied (the call a batch-reconstruction
    # handler makes on its first attempt).  Without flags: erasure-weighted inverse NTT + triangular recovery; with flags:
    # the dense matvec_kernel (43 check/coefficient rows x 22 terms).
    needed = DEG + T_FAULTS + 1
    ev43, ids43 = evals[:needed], np.arange(needed)
    flags43 = torch.empty((B, 1), dtype=torch.int64, device=dev)

    def timed3(fn):
        fn()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        for _ in range(3):
            fn()
        d1.record(stream)
        assert ctx.synchronize() == 0 and torch.equal(rec, coeffs)
        return d0.elapsed_time(d1) / 3 * 1e-3

    t_gen43 = timed3(lambda: ctx.batch_recover(ids43, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, None)))
    t_flags43 = timed3(lambda: ctx.batch_recover(ids43, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, flags43)))
    # the dense kernel on its own (roofline_dense): a context that sends calls with flags straight to the dense check, as every
    # chunk with a disagreeing share goes (HBMPC_NO_ER_FLAGS is a test knob read at context creation)
    os.environ["HBMPC_NO_ER_FLAGS"] = "1"
    ctx_dense = hb.Context(local)
    del os.environ["HBMPC_NO_ER_FLAGS"]
    ctx_dense.set_stream(stream.cuda_stream)
    ctx_dense.set_async(True)

    def timed3_dense():
        fn = lambda: ctx_dense.batch_recover(ids43, ev43, N_PARTIES, DEG, T_FAULTS, out=(rec, path, flags43))
        fn()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        for _ in range(3):
            fn()
        d1.record(stream)
        assert ctx_dense.synchronize() == 0 and torch.equal(rec, coeffs)
        return d0.elapsed_time(d1) / 3 * 1e-3

    t_dense = timed3_dense()
    ctx_dense.close()

    # ---- end-to-end leg: same calls with HOST (pinned) buffers, copies inside the timed region
    Be = 1 << args.log2_e2e_batch
    h_coeffs = torch.empty((Be, M, 4), dtype=torch.int64).pin_memory()
    h_coeffs.copy_(coeffs[:Be].cpu())
    h_shares = torch.empty((Be, N_PARTIES, 4), dtype=torch.int64).pin_memory()
    h_evals = torch.empty((N_PARTIES, Be, 4), dtype=torch.int64).pin_memory()
    h_evals.copy_(evals[:, :Be].cpu())
    h_rec = torch.empty((Be, M, 4), dtype=torch.int64).pin_memory()
    h_path = torch.empty((Be,), dtype=torch.int32).pin_memory()
    ctx.set_async(False)
    np_c, np_s, np_e, np_r, np_p = (x.numpy().view(np.uint64) if x.dtype == torch.int64 else x.numpy() for x in (h_coeffs, h_shares, h_evals, h_rec, h_path))

    def e2e_step():
        ctx.compute_shares_batch(np_c, N_PARTIES, out=np_s)
        rc, _, _, _ = ctx.batch_recover(ids, np_e, N_PARTIES, DEG, T_FAULTS, out=(np_r, np_p, None))
        assert rc == 0

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(np_r, np_c) and not np_p.any()
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = te.item()
    h2d = Be * M * 32 + Be * needed * 32  # coefficients in; of the 64 sender vectors the library uploads only the 43 it examines
    d2h = Be * N_PARTIES * 32 + Be * M * 32 + Be * 4

    # ---- NCCL gather of the reconstructed secrets (the only collective; outside the hot path)
    gather_ms = None
    if world > 1:
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

    # ---- secondary leg (rank 0, N=1 only, outside the timed region): robust interpolation with injected errors, BASELINE
    # configs[3] shape (n=128, t=42) at 2^log2_robust codewords, e ~ U{0..42} errors at uniform positions in each codeword
    robust = None
    if rank == 0 and n_gpus == 1 and not args.no_robust_leg:
        n4, t4 = 128, 42
        B4 = 1 << args.log2_robust
        ids4 = np.arange(n4)
        c4 = random_fr_device(torch, (B4, t4 + 1), 0x5EED0004, dev)
        s4 = torch.empty((B4, n4, 4), dtype=torch.int64, device=dev)
        ctx.set_async(True)
        ctx.compute_shares_batch(c4, n4, out=s4)
        g = torch.Generator(device=dev)
        g.manual_seed(44)
        e4 = torch.randint(0, t4 + 1, (B4,), device=dev, generator=g)
        perm = torch.rand((B4, n4), device=dev, generator=g).argsort(dim=1)
        mask = torch.zeros((B4, n4), dtype=torch.bool, device=dev)
        mask.scatter_(1, perm, torch.arange(n4, device=dev)[None, :] < e4[:, None])
        s4[..., 0] = torch.where(mask, s4[..., 0] ^ 0x5A5A5, s4[..., 0])
        o4 = (torch.empty((B4, t4 + 1, 4), dtype=torch.int64, device=dev), torch.empty((B4, 4), dtype=torch.int64, device=dev),
              torch.empty((B4,), dtype=torch.int32, device=dev), torch.empty((B4, 2), dtype=torch.int64, device=dev))
        ctx.set_async(False)   # synchronous calls: large failing sets take the staged decoder
        l4 = ctx.launch_count
        ctx.robust_interpolate_batch(ids4, s4, n4, t4, t4, out=o4)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(3):
            ctx.robust_interpolate_batch(ids4, s4, n4, t4, t4, out=o4)
        k1.record(stream)
        torch.cuda.synchronize()
        ms4 = k0.elapsed_time(k1) / 3
        assert torch.equal(o4[0], c4), "robust interpolation did not return the original polynomials"
        robust = {"workload": "robust_interpolate n=128 t=42, e ~ U{0..42} injected errors per codeword (BASELINE configs[3] shape)", "codewords": B4,
                  "ms": ms4, "codewords_per_s": B4 / (ms4 * 1e-3), "decoded_with_errors": int((o4[2] != 0).sum()), "max_oec_round": int(o4[2].max()),
                  "gpu_launches_per_call": int((ctx.launch_count - l4) // 4)}
        del c4, s4, o4, mask, perm

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cm = load_oracle_native()
        threads = cm.max_threads()
        Bc = 1 << args.cpu_log2_batch
        cc = coeffs[:Bc].cpu().numpy().view(np.uint64)
        cpu_step(cm, cc[: 1 << 12], threads)
        reps, tcpu = 0, 0.0
        while reps < 2 or (tcpu < 4.0 and reps < 64):
            tcpu += cpu_step(cm, cc, threads)
            reps += 1
        cpu = {"value": reps * 2 * Bc * N_PARTIES / tcpu, "unit": "shares/s", "cores": threads, "kind": "port",
               "sample": f"{reps} x 2^{args.cpu_log2_batch} secrets of the same workload (FFT compute_shares + batch_recover_secret), C oracle port -march=native, {threads} threads, {tcpu:.1f} s"}

    if rank == 0:
        shares_per_step = 2 * B * N_PARTIES
        value = n_gpus * args.steps * shares_per_step / t_tot
        gen_launch_s = t_gen / args.steps
        rec_launch_s = t_rec / args.steps
        hbm = measured_hbm()
        rec_alg_imad = B * ALG_MODMUL_REC * IMAD_PER_MODMUL
        gen_alg_imad = B * ALG_MODMUL_GEN * IMAD_PER_MODMUL
        # executed IMAD.WIDE (32x32->64 multiply-add) counts per secret: one Montgomery product = 8 rows x 16 = 128 wide;
        # 64-point radix-2 NTT: 129 non-trivial twiddle products (the inverse scales its 22 coefficients by 1/N with a shift, not a product); dense: 64 per term + 64 per reduction
        NTT_MULS = 129
        gen_exec_wide = B * NTT_MULS * 128
        rec_exec_wide = B * (NTT_MULS * 128 + M * 8)   # the 1/N scaling of the 22 coefficients is 8 narrow multiplies + a shift each (fr_div_pow2)
        dense_exec_wide = B * ((T_FAULTS + M) * M + (T_FAULTS + M)) * 64
        traffic = ncu_traffic()
        def roof(kernel, alg_imad, exec_wide, secs, nbytes, note, tkey):
            return {"kernel": kernel, "bound": "int32-imad", "achieved": alg_imad / secs / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                    "frac": alg_imad / secs / imad_peak, "how": note,
                    "executed_wide_tinst": exec_wide / secs / 1e12, "imad_wide_peak_tinst": imad_wide_peak / 1e12,
                    "executed_frac_of_imad_wide_peak": exec_wide / secs / imad_wide_peak,
                    "hbm_gbs": nbytes / secs / 1e9, "hbm_frac": nbytes / secs / 1e9 / hbm, "algorithmic_bytes": nbytes,
                    "traffic": (traffic[tkey] * B / (1 << 20)) if tkey in traffic else None}
        line = {
            "metric": "Fr shares generated+reconstructed per second (n=64,t=21)", "value": value, "unit": "shares/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (Fr, 255-bit Montgomery)", "data": "synthetic",
            "config": workload_config(args.log2_batch, args.log2_e2e_batch),
            "breakdown": {"gen_ms": 1e3 * t_gen / args.steps, "recon_ms": 1e3 * t_rec / args.steps,
                          "gen_shares_per_s": n_gpus * args.steps * B * N_PARTIES / t_gen, "recon_shares_per_s": n_gpus * args.steps * B * N_PARTIES / t_rec,
                          "alg_modmul_per_s": n_gpus * args.steps * B * (ALG_MODMUL_GEN + ALG_MODMUL_REC) / t_tot, "gather_ms": gather_ms,
                          "recon_43_senders_ms": 1e3 * t_gen43, "recon_43_senders_flags_ms": 1e3 * t_flags43, "recon_43_senders_dense_ms": 1e3 * t_dense, "robust_n128_t42": robust,
                          "recon_note": "recon_ms: all 64 senders supplied -> inverse NTT + degree check (bit-identical; items that fail fall back to the dense check); recon_43_senders_ms: only d+t+1 senders supplied -> erasure-weighted inverse NTT + triangular recovery; recon_43_senders_flags_ms: same call with flags -> the erasure-weighted transform over all supplied senders, dense check only for chunks it rejects; recon_43_senders_dense_ms: the dense matvec_kernel on every chunk (HBMPC_NO_ER_FLAGS=1)"},
            "roofline": roof("ntt64_cta_kernel<1> (K3 batch_recover launch, all 64 senders: inverse 64-point NTT + degree check per chunk)", rec_alg_imad, rec_exec_wide,
                             rec_launch_s, B * (N_PARTIES * 32 + M * 32 + 5),
                             "achieved = algorithmic IMAD (B * 1430 modmul * 256, SURVEY 8d dense count) / CUDA-event launch time; peak = mad.lo.u32 probe measured in this run; executed_* = IMAD.WIDE actually issued vs the IMAD.WIDE probe",
                             "ntt_inv"),
            "roofline_gen": roof("ntt64_cta_kernel<0> (K1 compute_shares launch: zero-padded 64-point NTT per secret)", gen_alg_imad, gen_exec_wide, gen_launch_s, B * BYTES_GEN,
                                 "algorithmic IMAD = B * 1344 modmul * 256 (dense Horner count of SURVEY 8d)", "ntt_fwd"),
            "roofline_dense": roof("matvec_kernel<4> (K3 dense check with flags, 43 senders: 43x22 check+coefficient matrix per chunk; the route of every chunk with a disagreeing share)", rec_alg_imad, dense_exec_wide, t_dense,
                                   B * (BYTES_REC + 8), "same algorithmic count, dense path", "matvec"),
            "cpu_baseline": cpu,
            "e2e": {"value": n_gpus * e2e_steps * 2 * Be * N_PARTIES / t_e2e, "unit": "shares/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps, "host_buffers": "pinned"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch at B = 2^20 from the committed ncu --set full capture (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_per_2p20.json")))
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
