"""Host emulation of the device limb arithmetic (mpc-protocols_b200/csrc/fr.cuh compiled with g++ under
HB_HOST_EMULATION) against Python big ints: even/odd lazy accumulator, carry counters, final reduction."""
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
        n_acc = n_ops = 0
        for ln in out:
            p = ln.split()
            if p[0] == "ops":
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
        assert n_acc == 10 and n_ops == 64
