"""Online-retry behaviour of the batch-reconstruction handler on the GPU path.

The reference's EvalBatch / RevealBatch arms (honeybadger/batch_recon/batch_recon.rs:332-481) accumulate sender vectors in
ARRIVAL order and call batch_recover_secret on every arrival once d+t+1 have been collected, latching the result on the
first success and keeping the shares on Err (so the next arrival retries with S+1 senders).  This test replays that loop
with random arrival orders and up to t Byzantine senders, once with the CPU oracle and once with the CUDA kernels as the
decoder, and requires every single call -- return code, coefficients, path -- to be identical, and the latched secrets to
be the true ones.  The FSM itself is test scaffolding (host control flow stays in the reference's Rust)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class HandlerMirror:
    """EvalBatch arm: accumulate (sender, values) in arrival order; decode when >= d+t+1 and nothing latched yet."""

    def __init__(self, n, d, t, decode):
        self.n, self.d, self.t, self.decode = n, d, t, decode
        self.ids, self.vals, self.result, self.log = [], [], None, []

    def on_eval_batch(self, sender, values):
        if sender in self.ids:                       # duplicate sender: dropped (batch_recon.rs:363-369)
            return
        self.ids.append(sender)
        self.vals.append(values)
        if len(self.ids) >= self.d + self.t + 1 and self.result is None:
            rc, coeffs, path = self.decode(np.array(self.ids), np.ascontiguousarray(np.stack(self.vals)))
            self.log.append((len(self.ids), rc, coeffs.copy(), path.copy()))
            if rc == 0:                              # `?` on Err keeps the accumulated shares, the next arrival retries
                self.result = coeffs[:, 0].copy()    # y_j = coeffs[0] per chunk (:391)


@pytest.mark.parametrize("n,t,chunks", [(10, 3, 5), (16, 5, 9), (64, 21, 6)])
def test_online_retry_matches_oracle(ctx, orc, n, t, chunks):
    d = t
    rng = np.random.default_rng(n * 31 + t)
    for trial in range(6):
        coeffs = orc.random_fr((chunks, d + 1), 0x5EED3000 + 10 * n + trial)
        rc, shares = orc.compute_shares(coeffs, n)                 # shares[chunk][sender]
        n_bad = int(rng.integers(0, t + 1))
        bad = set(rng.permutation(n)[:n_bad].tolist())
        arrival = rng.permutation(n).tolist()
        arrival.insert(int(rng.integers(1, n)), arrival[0])        # a duplicate delivery

        def gpu_decode(ids, ev):
            rc, co, path, _ = ctx.batch_recover(ids, ev, n, d, t)
            return rc, co, path

        def cpu_decode(ids, ev):
            out = orc.batch_recover_secret(ids, ev, n, d, t)
            return out["rc"], out["coeffs"], out["path"]

        hg, hc = HandlerMirror(n, d, t, gpu_decode), HandlerMirror(n, d, t, cpu_decode)
        for s in arrival:
            vals = shares[:, s].copy()
            if s in bad:                                            # a Byzantine sender garbles every chunk differently
                vals[:, 0] ^= rng.integers(1, 1 << 40, size=chunks).astype(np.uint64)
            hg.on_eval_batch(s, vals)
            hc.on_eval_batch(s, vals)
        assert len(hg.log) == len(hc.log) >= 1
        for (sg, rg, cg, pg), (sc, rcc, cc, pc) in zip(hg.log, hc.log):
            assert sg == sc and rg == rcc, (trial, sg, rg, rcc)
            assert np.array_equal(pg, pc) and np.array_equal(cg, cc), (trial, sg)
        # with <= t corrupted senders the handler must eventually latch the true values
        assert hg.result is not None and np.array_equal(hg.result, coeffs[:, 0])
        # the first attempt happens at exactly d+t+1 arrivals; with a Byzantine sender among them it must fail
        # (optimistic check fails and OEC cannot start), and the handler retries on the next arrival
        first = hg.log[0]
        assert first[0] == d + t + 1
        if any(s in bad for s in hg.ids[: d + t + 1]):
            assert first[1] == 8 and len(hg.log) > 1
