"""hbmpc_group_* through the Python binding: a group of member contexts (every visible device; on a 1-GPU box several members share
device 0) splits host batches into contiguous ranges and must return exactly what one context returns -- and what the oracle says."""
import numpy as np
import pytest

from test_gpu_parity import _codewords, _compare_recover, _corrupt, _rand

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def grp(hb):
    import torch

    nd = torch.cuda.device_count()
    g = hb.Group([i % nd for i in range(max(nd, 3))])
    yield g
    g.close()


def test_group_matches_oracle(grp, ctx, orc):
    n, t, d, B = 16, 5, 5, 3001
    coeffs = _rand(orc, (B, d + 1), 0x5EEDA000)
    rc, want = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    assert rc == 0 and np.array_equal(grp.compute_shares_batch(coeffs, n), want)
    rc, wantv = orc.apply_vandermonde(coeffs, n, True, threads=orc.max_threads())
    assert np.array_equal(grp.apply_vandermonde_batch(coeffs, n, recipient_major=True), wantv)
    rng = np.random.default_rng(5)
    _, shares = _codewords(orc, n, d, B, 0x5EEDA001)
    bad = _corrupt(shares, rng, np.where(rng.random(B) < 0.8, 0, rng.integers(1, t + 2, size=B)))
    arrival = rng.permutation(n)
    evals = np.ascontiguousarray(bad[:, arrival].transpose(1, 0, 2))
    ref = orc.batch_recover_secret(arrival, evals, n, d, t, threads=orc.max_threads())
    _compare_recover(grp.batch_recover(arrival, evals, n, d, t, want_flags=True), ref, B)
    rc, sec, path = grp.batch_recover_secrets(arrival, evals, n, d, t)
    ok = ref["path"] >= 0
    assert rc == ref["rc"] and np.array_equal(path, ref["path"]) and np.array_equal(sec[ok], ref["coeffs"][ok, 0])
    refk = orc.robust_interpolate_batch(arrival, bad[:, arrival], n, d, t, threads=orc.max_threads())
    rc, co, se, pa, fl = grp.robust_interpolate_batch(arrival, bad[:, arrival], n, d, t, want_flags=True)
    assert rc == refk["rc"] and np.array_equal(pa, refk["path"]) and np.array_equal(co, refk["coeffs"]) and np.array_equal(fl, refk["flags"][:, : fl.shape[1]])


def test_group_ranges_and_validation(grp, hb):
    B = 1000003
    prev = 0
    for i in range(grp.size):
        lo, hi = grp.shard_range(B, i)
        assert lo == prev and hi >= lo
        prev = hi
    assert prev == B
    with pytest.raises(hb.HbmpcError) as e:
        grp.compute_shares_batch(np.zeros((4, 9, 4), dtype=np.uint64), 8)   # n <= d
    assert e.value.code == hb.INVALID_INPUT
