"""BASELINE config 3: HBMPC Fig.2 batch reconstruction at n=64, t=21 -- the field work of ALL n parties, simulated in
one process on one GPU (the network is a device-side transpose; RBC/session routing are not modelled).
Each party i holds a degree-t share of every secret.  Per chunk of t+1 secrets:
  encode   : y^(i)_j = sum_k V[j][k] * share^(i)_k            (K2, batch_recon.rs:157-165)           party i -> party j
  round 1  : party j robustly interpolates y_j from the senders' y^(i)_j (a degree-t sharing)        (K3 secrets-only, :384-391)
  round 2  : every party interpolates the chunk's t+1 secrets from the revealed y_j                  (K3, :457-467)
python tools/bench_c3_protocol.py [--log2-secrets 16] [--corrupt 0]"""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hb = importlib.import_module("mpc-protocols_b200")
from bench import random_fr_device


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-secrets", type=int, default=16)
    ap.add_argument("--corrupt", type=int, default=0, help="number of parties that send garbage in round 1 (<= t)")
    ap.add_argument("--parties-timed", type=int, default=64, help="how many recipients' decode work to run (64 = everyone)")
    a = ap.parse_args()
    n, t = 64, 21
    m = t + 1
    S = 1 << a.log2_secrets
    chunks = (S + m - 1) // m
    dev = torch.device("cuda", 0)
    ctx = hb.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_async(True)
    ids = np.arange(n)
    # the secrets and every party's shares of them: shares[secret][party]
    polys = random_fr_device(torch, (chunks * m, t + 1), 1, dev)
    shares = torch.empty((chunks * m, n, 4), dtype=torch.int64, device=dev)
    ctx.compute_shares_batch(polys, n, out=shares)
    per_party = shares.view(chunks, m, n, 4).permute(2, 0, 1, 3).contiguous()      # [party i][chunk][k]
    secrets = polys[:, 0, :].reshape(chunks, m, 4)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e = [ev() for _ in range(4)]
    # ---- encode: every party applies the n x (t+1) Vandermonde to each of its chunks (recipient-major output)
    y_from = torch.empty((n, n, chunks, 4), dtype=torch.int64, device=dev)          # [sender i][recipient j][chunk]
    e[0].record()
    for i in range(n):
        ctx.apply_vandermonde_batch(per_party[i], n, recipient_major=True, out=y_from[i])
    e[1].record()
    # ---- the network: recipient j receives y_from[:, j]  (sender-major [S][chunks] is exactly the kernels' input layout)
    y_at = y_from.permute(1, 0, 2, 3).contiguous()                                   # [recipient j][sender i][chunk]
    if a.corrupt:
        y_at[:, : a.corrupt, :, 0] ^= 0x5555                                         # corrupted senders garble everything they send
    # ---- round 1: each recipient j opens y_j (one value per chunk)
    y_open = torch.empty((n, chunks, 4), dtype=torch.int64, device=dev)
    path1 = torch.empty((n, chunks), dtype=torch.int32, device=dev)
    e[2].record()
    for j in range(a.parties_timed):
        ctx.batch_recover_secrets(ids, y_at[j], n, t, t, out=(y_open[j], path1[j]))
    # ---- round 2: everybody receives all y_j (honest here) and interpolates the t+1 secrets of every chunk
    rec = torch.empty((chunks, m, 4), dtype=torch.int64, device=dev)
    path2 = torch.empty((chunks,), dtype=torch.int32, device=dev)
    for j in range(a.parties_timed):
        ctx.batch_recover(ids, y_open, n, t, t, out=(rec, path2, None))
    e[3].record()
    rc = ctx.synchronize()
    ok = bool(torch.equal(rec, secrets)) if a.parties_timed == n else None
    t_enc, t_dec = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
    scale = n / a.parties_timed
    total = t_enc + t_dec * scale
    print(json.dumps({"config": "C3 batch reconstruction, all 64 parties simulated on one B200", "secrets": chunks * m, "chunks": chunks,
                      "corrupted_senders": a.corrupt, "rc": rc, "secrets_recovered_by_every_party": ok,
                      "encode_ms_all_parties": round(t_enc, 3), "decode_ms_all_parties": round(t_dec * scale, 3),
                      "secrets_per_s_whole_protocol": chunks * m / (total * 1e-3),
                      "secrets_per_s_per_party": chunks * m / (total / n * 1e-3),
                      "max_round1_path": int(path1[: a.parties_timed].max())}, indent=1))


if __name__ == "__main__":
    main()
