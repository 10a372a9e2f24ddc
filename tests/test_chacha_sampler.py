"""N4: the randomness of a reference sharing (rand 0.8 StdRng = ChaCha12, ark-ff 0.5 Fp::rand).  CPU: the restatement
(oracle/chacha_fr.py) against published ChaCha vectors; GPU: the device sampler (csrc/sampler.cuh) against the restatement."""
import numpy as np
import pytest

from oracle import chacha_fr as cf


def _hex(words):
    return b"".join(w.to_bytes(4, "little") for w in words).hex()


def test_chacha_block_known_answers():
    # RFC 7539 section 2.3.2 (ChaCha20): key 00..1f, block counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00
    b = cf.chacha_block(bytes(range(32)), 1 | (0x09000000 << 32), 0x4A000000, rounds=20)
    assert b == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                 0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    # draft-strombergson-chacha-test-vectors TC1 (all-zero 256-bit key and IV), first keystream block, 12 / 8 / 20 rounds
    assert _hex(cf.chacha_block(bytes(32), 0, 0, 12)) == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                                                           "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    assert _hex(cf.chacha_block(bytes(32), 0, 0, 8)).startswith("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e")
    assert _hex(cf.chacha_block(bytes(32), 0, 0, 20)).startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")


def test_fr_rand_restatement():
    seed = bytes(range(32))
    rng = cf.StdRng(seed)
    words = [rng.next_u64() for _ in range(16)]
    blk0, blk1 = cf.chacha_block(seed, 0), cf.chacha_block(seed, 1)
    assert words[0] == blk0[0] | (blk0[1] << 32) and words[7] == blk0[14] | (blk0[15] << 32) and words[8] == blk1[0] | (blk1[1] << 32)
    vals = cf.sample_fr(seed, 0, 2000)
    assert all(0 <= v < cf.R for v in vals) and len(set(vals)) == 2000
    # the draws are the accepted candidates, in order: recompute them from the raw limbs
    rng = cf.StdRng(seed)
    acc, cand = [], 0
    while len(acc) < 2000:
        l = [rng.next_u64() for _ in range(4)]
        l[3] &= (1 << 63) - 1
        x = l[0] | (l[1] << 64) | (l[2] << 128) | (l[3] << 192)
        cand += 1
        if x < cf.R:
            acc.append(x * cf.R_INV % cf.R)
    assert acc == vals and 0.88 < 2000 / cand < 0.93          # acceptance r / 2^255 = 0.9057
    polys = cf.sample_polynomials(seed, 10, 3)
    flat = cf.sample_fr(seed, 0, 50)
    assert polys[0] == [flat[0], flat[2], flat[3], flat[4]] and polys[1][0] == flat[5] and polys[1][1] == flat[7]
    given = cf.sample_polynomials(seed, 10, 3, secrets=list(range(100, 110)))
    assert given[0] == [100, flat[1], flat[2], flat[3]] and given[2][0] == 102 and given[1][1] == flat[5]


@pytest.mark.gpu
def test_device_sampler_matches_restatement(hb, ctx):
    for seed in (bytes(32), bytes(range(32)), bytes((7 * i + 3) & 0xFF for i in range(32))):
        want = cf.sample_fr(seed, 0, 3000)
        got = ctx.sample_fr_batch(seed, 3000)
        assert hb.from_limbs(got) == want
        for B, d in ((1, 0), (257, 5), (40, 21), (9, 42)):
            wp = cf.sample_polynomials(seed, B, d)
            gp = ctx.sample_polynomials(seed, B, d)
            assert hb.from_limbs(gp) == wp
            secrets = hb.to_limbs([1000 + b for b in range(B)])
            wq = cf.sample_polynomials(seed, B, d, secrets=[1000 + b for b in range(B)])
            gq = ctx.sample_polynomials(seed, B, d, secrets=secrets)
            assert hb.from_limbs(gq) == wq


@pytest.mark.gpu
def test_device_sampler_large_and_device_resident(hb, ctx, orc):
    """2^20 sharings of degree 21 drawn on the device into a device buffer, then shared: a prefix equals the restatement, every value
    is canonical, and the call feeds compute_shares without any upload of coefficients."""
    import torch

    seed = bytes(range(1, 33))
    B, d, n = 1 << 18, 21, 64
    dev = torch.device("cuda", 0)
    coeffs = torch.empty((B, d + 1, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctx.sample_polynomials(seed, B, d, out=coeffs)
    host = coeffs[:64].cpu().numpy().view(np.uint64)
    assert hb.from_limbs(host) == cf.sample_polynomials(seed, 64, d)
    tail = coeffs[-2:].cpu().numpy().view(np.uint64)
    shares = ctx.compute_shares_batch(coeffs, n)          # non-canonical values would raise InvalidInput
    rc, want = orc.compute_shares(tail, n)
    assert rc == 0 and np.array_equal(shares[-2:].cpu().numpy().view(np.uint64), want)


@pytest.mark.gpu
def test_share_secrets_batch_is_compute_shares_on_a_seeded_generator(hb, ctx, orc):
    """hbmpc_share_secrets_batch(seed, secrets) == the oracle's compute_shares on the polynomials B consecutive
    RobustShare::compute_shares(secret, n, d, None, rng) calls would draw from StdRng::from_seed(seed) (robust_interpolate.rs:68-69):
    host buffers (pipelined chunks), device buffers, with and without the polynomials returned."""
    import torch

    seed = bytes((11 * i + 5) & 0xFF for i in range(32))
    for n, d, B in ((4, 1, 3), (16, 5, 700), (64, 21, 5000), (64, 42, 300), (128, 42, 200)):
        sec = [(123456789 * (b + 1)) % hb.R_MOD for b in range(B)]
        polys = cf.sample_polynomials(seed, B, d, secrets=sec)
        rc, want = orc.compute_shares(hb.to_limbs(polys), n)
        assert rc == 0
        cout = np.zeros((B, d + 1, 4), dtype=np.uint64)
        got = ctx.share_secrets_batch(seed, hb.to_limbs(sec), n, d, coeffs_out=cout)
        assert np.array_equal(got, want) and hb.from_limbs(cout) == polys
        dsec = torch.from_numpy(hb.to_limbs(sec).view(np.int64)).cuda()
        dgot = ctx.share_secrets_batch(seed, dsec, n, d)
        assert ctx.synchronize() == 0
        assert np.array_equal(dgot.cpu().numpy().view(np.uint64), want)
    with pytest.raises(hb.HbmpcError) as e:
        ctx.share_secrets_batch(seed, hb.to_limbs([1, 2]), 4, 4)
    assert e.value.code == hb.INVALID_INPUT
