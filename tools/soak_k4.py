"""Randomized soak of the K4 routes against each other (GPU only, no oracle: fast enough for 10^5-codeword batches).
Every round draws a shape (n, t, degree, sender subset, arrival order), a batch with a mix of error patterns (clean, <= t errors,
errors only beyond the examined prefix, all errors inside the prefix, more errors than can be decoded, persistent attackers) and
runs hbmpc_robust_interpolate_batch / hbmpc_batch_recover through
  (a) the staged decoder (production thresholds, synchronous call),
  (b) the staged decoder with tiny waves and the direct mode switched off,
  (c) robust_kernel alone (HBMPC_STAGED_MIN huge),
  (d) asynchronous (enqueue-only) calls through the staged decoder in device-count mode (HBMPC_ASYNC_STAGED=2), status at synchronize,
and requires identical return codes, coefficients, secrets, paths and flags.  The decoded polynomials are also checked against
the ground truth wherever the number of errors is decodable.
   python tools/soak_k4.py [--seconds 120] [--seed 1]"""
import argparse, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import random_fr_device


def make_ctx(hb, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return hb.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    hb = importlib.import_module("mpc-protocols_b200")
    dev = torch.device("cuda", 0)
    base = {"HBMPC_SCAN_MAX": "0"}
    ctxs = {"staged": make_ctx(hb, {**base}),
            "staged_small_waves_no_direct": make_ctx(hb, {**base, "HBMPC_STAGED_MIN": "1", "HBMPC_STAGED_WS_MB": "64", "HBMPC_NO_STAGED_DIRECT": "1", "HBMPC_STAGED_SEG": "5"}),
            "per_thread": make_ctx(hb, {**base, "HBMPC_STAGED_MIN": str(1 << 40)}),
            "async_device_count": make_ctx(hb, {"HBMPC_ASYNC_STAGED": "2", "HBMPC_STAGED_WS_MB": "256"})}
    ctxs["async_device_count"].set_async(True)
    rng = np.random.default_rng(a.seed)
    g = torch.Generator(device=dev)
    g.manual_seed(a.seed)
    t0, rounds, items = time.time(), 0, 0
    shapes = [(16, 5), (64, 21), (128, 42), (32, 10), (13, 4), (100, 33)]
    while time.time() - t0 < a.seconds:
        n, t = shapes[rng.integers(len(shapes))]
        d = t if rng.random() < 0.75 or n < 3 * t + 1 + t else 2 * t
        if n < d + t + 2:
            d = t
        needed = d + t + 1
        S = n if rng.random() < 0.5 else int(rng.integers(needed + 1, n + 1))
        B = int(rng.choice([6000, 20000, 50000])) if n <= 64 else int(rng.choice([5000, 12000]))
        ids = np.sort(rng.choice(n, size=S, replace=False))
        arrival = rng.permutation(S)
        ids_arr = ids[arrival]
        coeffs = random_fr_device(torch, (B, d + 1), int(rng.integers(1 << 30)), dev)
        torch.cuda.synchronize()   # torch's stream and the contexts' streams are not ordered with each other
        shares = ctxs["staged"].compute_shares_batch(coeffs, n)
        words = shares[:, torch.as_tensor(ids_arr, device=dev)].contiguous()      # [B][S] in arrival order
        rmax = min(t, S - needed)
        maxdec = min(t, (S - d - 1) // 2)
        kind = torch.randint(0, 6, (B,), device=dev, generator=g)
        e = torch.zeros((B,), dtype=torch.int64, device=dev)
        e = torch.where(kind == 1, torch.randint(1, max(rmax, 1) + 1, (B,), device=dev, generator=g), e)          # decodable
        e = torch.where(kind == 2, torch.randint(0, t + 3, (B,), device=dev, generator=g), e)                     # anything up to t+2
        e = torch.where(kind == 3, torch.full((B,), max(rmax, 1), device=dev), e)                                # all inside the prefix
        e = torch.where(kind == 4, torch.randint(1, max(S - needed, 1) + 1, (B,), device=dev, generator=g), e)    # only beyond the prefix
        e = torch.clamp(e, max=S)
        score = torch.rand((B, S), device=dev, generator=g)
        pos_sorted = torch.as_tensor(np.argsort(ids_arr), device=dev)            # arrival index of sorted position i
        inside = torch.zeros((S,), dtype=torch.bool, device=dev)
        inside[pos_sorted[:needed]] = True
        score = torch.where((kind == 3)[:, None] & ~inside[None, :], torch.full_like(score, 2.0), score)
        score = torch.where((kind == 4)[:, None] & inside[None, :], torch.full_like(score, 2.0), score)
        e = torch.where(kind == 4, torch.clamp(e, max=S - needed), e)
        e = torch.where(kind == 3, torch.clamp(e, max=needed), e)
        rank = score.argsort(dim=1).argsort(dim=1)
        mask = rank < e[:, None]
        if rmax >= 1:                                                            # persistent attackers in kind 5
            bad = torch.as_tensor(rng.permutation(S)[: max(1, rmax)], device=dev)
            pm = torch.zeros((S,), dtype=torch.bool, device=dev)
            pm[bad] = True
            mask = torch.where((kind == 5)[:, None], pm[None, :].expand(B, S), mask)
        bump = torch.randint(1, 1 << 20, (B, S), device=dev, generator=g)
        words[..., 0] = torch.where(mask, words[..., 0] ^ bump, words[..., 0])
        evals = words.permute(1, 0, 2).contiguous()
        torch.cuda.synchronize()   # the contexts run on their own (non-blocking) streams: the inputs must be complete before the calls
        outs = {}
        for name, c in ctxs.items():
            for fl in (True, False):
                rc, co, sec, path, flags = c.robust_interpolate_batch(ids_arr, words, n, d, t, want_flags=fl)
                if name == "async_device_count":
                    rc = c.synchronize()
                rb = c.batch_recover(ids_arr, evals, n, d, t, want_flags=fl)
                if name == "async_device_count":
                    rb = (c.synchronize(),) + tuple(rb[1:])
                outs[(name, fl)] = (rc, co, sec, path, flags, rb)
        ref = outs[("per_thread", True)]
        for (name, fl), o in outs.items():
            tag = f"round {rounds} n={n} t={t} d={d} S={S} B={B} route={name} flags={fl}"
            assert o[0] == ref[0], tag + f": rc {o[0]} != {ref[0]}"
            assert torch.equal(o[3], ref[3]), tag + ": path"
            assert torch.equal(o[1], ref[1]) and torch.equal(o[2], ref[2]), tag + ": coefficients"
            if fl:
                assert torch.equal(o[4], ref[4]), tag + ": flags"
            rb = o[5]
            assert rb[0] == ref[5][0] and torch.equal(rb[2], ref[5][2]) and torch.equal(rb[1], ref[5][1]), tag + ": batch_recover"
        nerr = mask.sum(dim=1)
        good = nerr <= min(rmax, maxdec)
        ok = ref[3] >= 0
        assert bool((ok | ~good).all()), "a decodable codeword was rejected"
        assert torch.equal(ref[1][good], coeffs[good]), "decoded polynomial differs from the ground truth"
        rounds += 1
        items += B
    for c in ctxs.values():
        c.close()
    print(f"soak ok: {rounds} rounds, {items} codewords x 4 routes x 4 calls, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
