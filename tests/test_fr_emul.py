"""Host emulation of the device limb arithmetic (mpc-protocols_b200/csrc/fr.cuh compiled with g++ under
HB_HOST_EMULATION) against Python big ints: even/odd lazy accumulator, carry counters, final reduction, the fused share algebra of K5."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
RINV = pow(1 << 256, -1, R)


def test_fr_cuh_limb_logic(tmp_path):
    exe = tmp_path / "fr_emul"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "fr_emul.cpp")], check=True)
    for seed in (1, 2, 3):
        out = subprocess.run([str(exe), str(seed)], capture_output=True, text=True, check=True).stdout.splitlines()
        n_acc = n_ops = n_k5 = 0
        for ln in out:
            p = ln.split()
            if p[0] == "k5":   # fused share algebra: a*b - c and c - da*db - da*y - db*x (operands: c, x, y, da, db)
                v = [int(x, 16) for x in p[1:8]]
                assert v[5] == (v[0] * v[1] - v[2]) % R
                assert v[6] == (v[0] - v[3] * v[4] - v[3] * v[2] - v[4] * v[1]) % R
                n_k5 += 1
            elif p[0] == "ops":
                a, b, m, s, d, mc = (int(x, 16) for x in p[1:7])
                assert m == a * b * RINV % R and s == (a + b) % R and d == (a - b) % R
                assert mc == m, "mont_mul_cios differs"
                n_ops += 1
            else:
                terms = int(p[0])
                vals = [int(x, 16) for x in p[1:]]
                acc = sum(vals[2 * k] * vals[2 * k + 1] for k in range(terms))
                assert vals[-1] == acc * RINV % R, f"terms={terms}"
                n_acc += 1
        assert n_acc == 10 and n_ops == 64 and n_k5 == 48


def test_division_by_power_of_two_identity():
    """The inverse transforms scale by 1/N with a shift (ntt.cuh: fr_div_pow2): r = 1 mod 2^32, so with k = -v mod N the sum v + k*r is
    divisible by N = 2^logn and (v + k*r) >> logn is the canonical v * N^-1 mod r.  Checked here with big ints for every domain size the
    kernels support, on edge values and random ones (the device code is covered bit for bit by the GPU parity tests of K3)."""
    import random

    assert R % (1 << 32) == 1
    rnd = random.Random(5)
    for logn in range(0, 9):
        n = 1 << logn
        ninv = pow(n, -1, R)
        for v in [0, 1, 2, n - 1, n, n + 1, R - 1, R - 2, R - n, (R - 1) // 2] + [rnd.randrange(R) for _ in range(200)]:
            k = (-v) % n
            s = v + k * R
            assert s % n == 0
            q = s >> logn
            assert q < R and q == v * ninv % R
            assert s < 1 << 288   # nine 32-bit limbs hold the sum
