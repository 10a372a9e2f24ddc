"""N > 1 host logic on CPU: world_size-2 gloo run of the batch sharding + result gather (mpc-protocols_b200/sharding.py).
The per-shard "compute" is the CPU oracle here (test infrastructure); on GPUs the same code path runs the CUDA kernels
and NCCL."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    sh = importlib.import_module("mpc-protocols_b200.sharding")
    for total in (0, 1, 7, 64, 1000, 4194304):
        for world in (1, 2, 3, 4, 8):
            rs = [sh.shard_range(total, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sh.shard_sizes(total, world)
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sh = importlib.import_module("mpc-protocols_b200.sharding")
    from oracle import cmodel as cm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, t = 16, 5
        coeffs = cm.random_fr((total, t + 1), 0x5EED0010)      # the global batch (same on every rank)
        lo, hi = sh.shard_range(total, world, rank)
        rc, shares = cm.compute_shares(coeffs[lo:hi], n)        # this rank's shard of K1
        evals = np.ascontiguousarray(shares.transpose(1, 0, 2))
        out = cm.batch_recover_secret(np.arange(n), evals, n, t, t)  # this rank's shard of K3
        secrets = torch.from_numpy(out["coeffs"][:, 0].astype(np.int64))
        full = sh.gather_shards(secrets, total)
        ok = bool(np.array_equal(full.numpy().astype(np.uint64), coeffs[:, 0])) and rc == 0 and out["rc"] == 0
        q.put((rank, ok, tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [37, 64])
def test_two_rank_gloo_shard_and_gather(total):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == (total, 4) for r in res)
