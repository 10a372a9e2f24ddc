"""GPU parity of the staged robust decoder (robust.cuh: syndromes NTT -> segmented, re-sorted Berlekamp-Massey -> Chien /
Forney NTTs -> finish).  Production calls only take it for failing sets of thousands of items; here HBMPC_STAGED_MIN=1 and
HBMPC_SCAN_MAX=0 force every failing item of test-sized batches through it.  Outcomes (coefficients, secrets, path, flags,
return code) must equal the oracle's item by item, and therefore the per-thread decoder's (robust_interpolate.rs:94-157,
456-538, 579-628)."""
import itertools

import numpy as np
import pytest

from test_gpu_parity import _codewords, _compare_recover, _corrupt, _rand

pytestmark = pytest.mark.gpu


@pytest.fixture
def staged_ctx(hb, monkeypatch):
    made = []

    def make(seg=None, ws_mb=None, speculation=False):
        monkeypatch.setenv("HBMPC_SCAN_MAX", "0")
        monkeypatch.setenv("HBMPC_STAGED_MIN", "1")
        monkeypatch.setenv("HBMPC_NO_SPECULATION", "0" if speculation else "1")
        if seg is not None:
            monkeypatch.setenv("HBMPC_STAGED_SEG", str(seg))
        if ws_mb is not None:
            monkeypatch.setenv("HBMPC_STAGED_WS_MB", str(ws_mb))
        c = hb.Context(0)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()


def _check_all_entry_points(c, orc, ids, words, n, d, t):
    """words[B][S] in arrival order; every K3/K4 entry point against the oracle."""
    want = orc.robust_interpolate_batch(ids, words, n, d, t, threads=orc.max_threads())
    for fl in (True, False):
        rc, co, sec, path, flags = c.robust_interpolate_batch(ids, words, n, d, t, want_flags=fl)
        assert np.array_equal(path, want["path"]), np.nonzero(path != want["path"])
        assert rc == want["rc"]
        assert np.array_equal(co, want["coeffs"]) and np.array_equal(sec, want["secrets"])
        if fl:
            assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
    evals = np.ascontiguousarray(words.transpose(1, 0, 2))
    want_b = orc.batch_recover_secret(ids, evals, n, d, t, threads=orc.max_threads())
    for fl in (True, False):
        _compare_recover(c.batch_recover(ids, evals, n, d, t, want_flags=fl), want_b, words.shape[0])
    rc3, sec3, path3 = c.batch_recover_secrets(ids, evals, n, d, t)
    assert rc3 == want_b["rc"] and np.array_equal(path3, want_b["path"]) and np.array_equal(sec3, want["secrets"])
    return want


@pytest.mark.parametrize("n,t", [(7, 2), (10, 3)])
def test_staged_all_error_subsets(staged_ctx, orc, hb, n, t):
    """robust_interpolate.rs:828-876 / :728-756 sweeps: every error subset of size <= t, secret 42."""
    c = staged_ctx(seg=2)
    d = t
    coeffs = _rand(orc, (1, d + 1), 99)
    coeffs[0, 0] = hb.to_limbs(42)
    rc, sh = orc.compute_shares(coeffs, n)
    subsets = [s for k in range(0, t + 1) for s in itertools.combinations(range(n), k)]
    words = np.repeat(sh, len(subsets), axis=0)
    for b, sub in enumerate(subsets):
        for p in sub:
            words[b, p, 0] += np.uint64(1000 + p)
    l0 = c.launch_count
    want = _check_all_entry_points(c, orc, np.arange(n), words, n, d, t)
    assert c.launch_count - l0 > 40, "the staged pipeline did not run"
    rc, co, sec, path, flags = c.robust_interpolate_batch(np.arange(n), words, n, d, t, want_flags=True)
    assert all(v == 42 for v in hb.from_limbs(sec))
    for b, sub in enumerate(subsets):
        assert int(flags[b, 0]) == sum(1 << p for p in sub)


@pytest.mark.parametrize("n,t,S,B,seg", [(16, 5, 16, 3000, 8), (16, 5, 14, 3000, 3), (13, 4, 12, 1000, 1), (64, 21, 64, 1500, 8), (64, 21, 50, 700, 5),
                                         (64, 21, 44, 300, 8), (10, 3, 8, 500, 8)])
def test_staged_mixed_error_counts(staged_ctx, orc, n, t, S, B, seg):
    """Error counts from 0 to beyond what the rounds can absorb, sender subsets whose ids do not start at 0, random arrival
    order; the smallest workspace budget, so the larger batches take several waves."""
    c = staged_ctx(seg=seg, ws_mb=1)
    d = t
    rng = np.random.default_rng(S * 131 + n)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7000 + n + S)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]]
    nerr = np.minimum(rng.integers(0, t + 3, size=B), S)
    bad = _corrupt(words, rng, nerr)
    # a few items with all their errors inside the examined prefix, and a few whose errors lie beyond it (path 0)
    needed = d + t + 1
    pos_sorted = np.argsort(ids[arrival])  # arrival index of sorted position i
    for b in range(0, min(B, 40), 4):
        bad[b] = words[b]
        e = int(rng.integers(1, min(t, max(S - needed, 1)) + 1))
        for p in rng.choice(pos_sorted[:needed], size=e, replace=False):
            bad[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 20))
    if S > needed:
        for b in range(1, min(B, 40), 8):
            bad[b] = words[b]
            bad[b, pos_sorted[needed:][0], 0] ^= np.uint64(77)
    _check_all_entry_points(c, orc, ids[arrival], bad, n, d, t)


def test_staged_degree_2t(staged_ctx, orc):
    """degree-2t sharings (triple generation opens d = 2t with n = 3t + 1 + spare senders)."""
    n, t, d, B = 16, 3, 6, 1200
    c = staged_ctx(seg=4)
    rng = np.random.default_rng(17)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7100)
    nerr = rng.integers(0, t + 2, size=B)
    bad = _corrupt(shares, rng, nerr)
    _check_all_entry_points(c, orc, np.arange(n), bad, n, d, t)


def test_staged_n128_t42(staged_ctx, orc):
    """config (4) shape: n = 128, t = 42, e ~ U{0..42} plus the adversarial all-errors-in-prefix codeword."""
    n, t, d, B = 128, 42, 42, 40
    c = staged_ctx()
    rng = np.random.default_rng(4242)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7004)
    nerr = rng.integers(0, t + 1, size=B)
    nerr[0], nerr[1], nerr[3] = 0, t, 1
    bad = _corrupt(shares, rng, nerr)
    bad[2] = _corrupt(shares[2:3], rng, [t], positions=d + t + 1)[0]
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
    for fl in (True, False):
        rc, co, sec, path, flags = c.robust_interpolate_batch(ids, bad, n, d, t, want_flags=fl)
        assert rc == want["rc"] == 0
        assert np.array_equal(path, want["path"])
        assert np.array_equal(co, coeffs) and np.array_equal(co, want["coeffs"])
        if fl:
            assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])


def test_staged_n255(staged_ctx, orc):
    """largest party count (256-point transforms, 3 stages per pass in the NTT kernels)."""
    n, t, d, B = 255, 84, 84, 10
    c = staged_ctx(seg=16)
    rng = np.random.default_rng(255)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED70FF)
    nerr = np.array([0, 1, 2, 3, 5, 8, 13, 21, 4, 6])
    bad = _corrupt(shares, rng, nerr)
    ids = np.arange(n)
    want = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
    rc, co, sec, path, flags = c.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
    assert rc == want["rc"] == 0
    assert np.array_equal(path, want["path"]) and np.array_equal(co, want["coeffs"])
    assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])


def test_staged_equals_per_thread_decoder_large(hb, orc, monkeypatch):
    """Production thresholds: 2^14 codewords at n = 64 with random error counts through the staged decoder (default
    HBMPC_STAGED_MIN) and through robust_kernel alone (HBMPC_STAGED_MIN huge): identical outputs; a sample is compared with
    the oracle."""
    n, t, d, B = 64, 21, 21, 1 << 14
    rng = np.random.default_rng(64)
    coeffs = _rand(orc, (B, d + 1), 0x5EED7200)
    rc, shares = orc.compute_shares(coeffs, n, threads=orc.max_threads())
    nerr = rng.integers(0, t + 2, size=B)
    bad = _corrupt(shares, rng, nerr)
    ids = np.arange(n)
    monkeypatch.setenv("HBMPC_SCAN_MAX", "0")
    monkeypatch.setenv("HBMPC_STAGED_WS_MB", "40")   # several waves: the first one samples the wrong senders, the others append leftovers
    c_new = hb.Context(0)
    monkeypatch.setenv("HBMPC_STAGED_MIN", str(1 << 40))
    c_old = hb.Context(0)
    try:
        l0 = c_new.launch_count
        a = c_new.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
        ln = c_new.launch_count - l0
        l0 = c_old.launch_count
        b = c_old.robust_interpolate_batch(ids, bad, n, d, t, want_flags=True)
        lo = c_old.launch_count - l0
        assert ln > lo + 20, "the staged pipeline did not run"
        assert a[0] == b[0]
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)
        sub = slice(0, 256)
        want = orc.robust_interpolate_batch(ids, bad[sub], n, d, t, threads=orc.max_threads())
        assert np.array_equal(a[3][sub], want["path"]) and np.array_equal(a[1][sub], want["coeffs"])
        assert np.array_equal(a[4][sub], want["flags"][:, : a[4].shape[1]])
    finally:
        c_new.close()
        c_old.close()


def test_staged_after_speculation(staged_ctx, orc):
    """persistent attackers plus scattered errors: the shortcut explains most items, the rest reach the staged decoder."""
    n, t, d, B, S = 16, 5, 5, 3000, 16
    c = staged_ctx(speculation=True, ws_mb=1)   # 1024-slot waves: the first wave's sample finds the attackers, the rest takes the shortcut
    rng = np.random.default_rng(99)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7300)
    words = shares.copy()
    bad_senders = rng.permutation(S)[:t]
    for b in range(B):
        if b % 3 == 0:
            for p in rng.choice(S, size=int(rng.integers(1, t + 2)), replace=False):
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
        else:
            for p in bad_senders:
                words[b, p, 0] ^= np.uint64(rng.integers(1, 1 << 30))
    _check_all_entry_points(c, orc, np.arange(n), words, n, d, t)


@pytest.mark.parametrize("sync_count", [True, False])
def test_mid_size_batches_switch_to_the_staged_decoder_under_attack(hb, orc, monkeypatch, sync_count):
    """Synchronous batches of 2048 .. HBMPC_SCAN_MAX chunks.  Round 2 (sync_count): the failing items are compacted, robust_kernel is
    enqueued for a small failing set and told to skip a large one, and the call's one synchronisation brings the count: the FIRST
    attacked call already takes the staged decoder.  Round 1's route (HBMPC_NO_SYNC_COUNT=1): no compaction, the first attacked call
    decodes with robust_kernel and reports the densely failing batch, the next calls of the context count their failing items, until a
    call finds (almost) nothing to decode.  Every call equals the oracle whatever the route."""
    if not sync_count:
        monkeypatch.setenv("HBMPC_NO_SYNC_COUNT", "1")
    n, t, d, B = 16, 5, 5, 6000
    rng = np.random.default_rng(2024)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7400)
    bad = _corrupt(shares, rng, rng.integers(1, t + 1, size=B))
    few = shares.copy()
    few[:300] = bad[:300]                      # a small failing set: decoded by the speculative robust_kernel launch
    ids = np.arange(n)
    want_bad = orc.robust_interpolate_batch(ids, bad, n, d, t, threads=orc.max_threads())
    want_few = orc.robust_interpolate_batch(ids, few, n, d, t, threads=orc.max_threads())
    c = hb.Context(0)
    try:
        def call(words, want=None):
            l0 = c.launch_count
            rc, co, sec, path, flags = c.robust_interpolate_batch(ids, words, n, d, t, want_flags=True)
            if want is not None:
                assert rc == want["rc"] and np.array_equal(path, want["path"]) and np.array_equal(co, want["coeffs"])
                assert np.array_equal(flags, want["flags"][:, : flags.shape[1]])
            else:
                assert rc == 0 and not path.any() and np.array_equal(co, coeffs)
            return c.launch_count - l0
        honest0 = call(shares)
        small = call(few, want_few) if sync_count else 0   # (round 1's route takes 300 consecutive failing items for an attack)
        first = call(bad, want_bad)
        second = call(bad, want_bad)
        third = call(bad, want_bad)
        if sync_count:
            assert small <= honest0 + 1 and first >= honest0 + 8 and second >= honest0 + 8 and third == second, (honest0, small, first, second, third)
        else:
            assert second >= first + 8 and third == second, (honest0, first, second, third)
        after = call(shares)                 # counted, nothing to decode: the attack is over
        honest1 = call(shares)
        assert honest1 == honest0 and after >= honest0, (honest0, after, honest1)
        if sync_count:
            assert call(few, want_few) == small
    finally:
        c.close()


@pytest.mark.parametrize("n,t,S,B,ws_mb", [(16, 5, 16, 3000, None), (16, 5, 14, 2000, 1), (64, 21, 64, 1500, None), (64, 21, 50, 700, 1), (10, 3, 8, 500, None),
                                           (128, 42, 128, 300, None)])
def test_async_calls_take_the_staged_decoder_in_device_count_mode(hb, orc, monkeypatch, n, t, S, B, ws_mb):
    """Asynchronous calls cannot ask the host how many items failed.  Under attack (HBMPC_ASYNC_STAGED=2 forces the state a context
    reaches after hbmpc_ctx_synchronize has seen a large failing set) they compact the failing items and run the staged decoder with
    the count read ON THE DEVICE by every stage (slots beyond it are dead from the start and never reach the exact path).
    Device tensors, enqueue-only calls, status at synchronize: outcomes equal the oracle's item by item."""
    import torch

    monkeypatch.setenv("HBMPC_STAGED_MIN", "1")
    monkeypatch.setenv("HBMPC_ASYNC_STAGED", "2")
    monkeypatch.setenv("HBMPC_NO_SPECULATION", "1")
    if ws_mb is not None:
        monkeypatch.setenv("HBMPC_STAGED_WS_MB", str(ws_mb))
    c = hb.Context(0)
    c.set_async(True)
    d = t
    rng = np.random.default_rng(S * 17 + n)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7A00 + n + S)
    ids = np.sort(rng.choice(n, size=S, replace=False))
    arrival = rng.permutation(S)
    words = shares[:, ids[arrival]]
    nerr = np.minimum(rng.integers(0, t + 3, size=B), S)
    nerr[::5] = 0                                       # honest items in between
    bad = _corrupt(words, rng, nerr)
    aid = ids[arrival]
    want = orc.robust_interpolate_batch(aid, bad, n, d, t, threads=orc.max_threads())
    dev = torch.device("cuda", 0)
    tw = torch.from_numpy(np.ascontiguousarray(bad).view(np.int64)).to(dev)
    te = tw.permute(1, 0, 2).contiguous()
    torch.cuda.synchronize()
    for fl in (False, True):
        l0 = c.launch_count
        rc, co, sec, path, flags = c.robust_interpolate_batch(aid, tw, n, d, t, want_flags=fl)
        assert rc == 0                                   # enqueue only
        status = c.synchronize()
        assert c.launch_count - l0 > 12, "the staged pipeline did not run"
        assert status == want["rc"]
        assert np.array_equal(path.cpu().numpy(), want["path"])
        assert np.array_equal(co.cpu().numpy().view(np.uint64), want["coeffs"]) and np.array_equal(sec.cpu().numpy().view(np.uint64), want["secrets"])
        if fl:
            assert np.array_equal(flags.cpu().numpy().view(np.uint64), want["flags"][:, : flags.shape[1]])
        rc, co2, path2, _ = c.batch_recover(aid, te, n, d, t)
        assert c.synchronize() == want["rc"]
        assert np.array_equal(path2.cpu().numpy(), want["path"]) and np.array_equal(co2.cpu().numpy().view(np.uint64), want["coeffs"])
    # an honest batch through the same route: nothing to decode, nothing handed to the exact path
    tw2 = torch.from_numpy(np.ascontiguousarray(words).view(np.int64)).to(dev)
    torch.cuda.synchronize()
    rc, co, sec, path, _ = c.robust_interpolate_batch(aid, tw2, n, d, t)
    assert c.synchronize() == 0 and not path.cpu().numpy().any() and np.array_equal(co.cpu().numpy().view(np.uint64), coeffs)
    c.close()


def test_async_context_learns_of_an_attack_at_synchronize(hb, orc, monkeypatch):
    """without the forcing knob: the first attacked asynchronous call decodes with robust_kernel and raises the attack word; after
    hbmpc_ctx_synchronize the next asynchronous calls take the staged decoder (many more launches), honest ones switch it off again"""
    import torch

    monkeypatch.setenv("HBMPC_STAGED_MIN", "64")
    monkeypatch.setenv("HBMPC_NO_SPECULATION", "1")
    c = hb.Context(0)
    c.set_async(True)
    n, t, d, B = 64, 21, 21, 4096
    rng = np.random.default_rng(5)
    coeffs, shares = _codewords(orc, n, d, B, 0x5EED7B00)
    bad = _corrupt(shares, rng, rng.integers(1, t + 1, size=B))
    want = orc.robust_interpolate_batch(np.arange(n), bad[:256], n, d, t, threads=orc.max_threads())
    dev = torch.device("cuda", 0)
    tb, tg = torch.from_numpy(bad.view(np.int64)).to(dev), torch.from_numpy(shares.view(np.int64)).to(dev)
    torch.cuda.synchronize()
    counts = []
    for words in (tb, tb, tb, tg, tg, tb):
        l0 = c.launch_count
        rc, co, sec, path, _ = c.robust_interpolate_batch(np.arange(n), words, n, d, t)
        assert c.synchronize() == 0
        counts.append(c.launch_count - l0)
        if words is tb:
            assert np.array_equal(co[:256].cpu().numpy().view(np.uint64), want["coeffs"]) and np.array_equal(path[:256].cpu().numpy(), want["path"])
        else:
            assert np.array_equal(co.cpu().numpy().view(np.uint64), coeffs)
    assert counts[0] < 12 and counts[1] > 25 and counts[2] > 25, counts      # learnt at the first synchronize
    assert counts[4] < 12 and counts[5] < 12, counts                          # the honest call switched it off; the next attacked call pays the slow route once
    c.close()
