"""Builds and runs the C++ host-side mirror tests (tests/host/mirror_test.cpp: the reference's unit tests re-stated against
include/hbmpc_b200.hpp and the C ABI) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = tmp_path / "mirror_test"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "mirror_test.cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_mirror_compiles_and_links(tmp_path):
    """CPU: the C++ mirror compiles against the header and links against the built library."""
    assert os.path.exists(os.path.join(ROOT, "mpc-protocols_b200", "libhbmpc_b200.so")), "build the library first"
    _compile(tmp_path)


@pytest.mark.gpu
def test_reference_unit_tests_through_cpp_mirror(tmp_path):
    exe = _compile(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all reference-shaped tests passed" in res.stdout


def _compile_c(tmp_path):
    exe = tmp_path / "secret_share_b200"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "secret_share_b200.c"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_c99_consumer_compiles_and_links(tmp_path):
    _compile_c(tmp_path)


@pytest.mark.gpu
def test_reference_c_abi_test_restated(tmp_path):
    """mpc/src/ffi/tests/secret_share.c restated in C99 against the batch C ABI."""
    exe = _compile_c(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all C-ABI round trips passed" in res.stdout


def _compile_batch_recon(tmp_path):
    exe = tmp_path / "batch_recon_test"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "batch_recon_test.cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_batch_recon_mirror_compiles_and_links(tmp_path):
    """CPU: the C++ BatchReconNode mirror (include/hbmpc_batch_recon.hpp) compiles and links against the built library; its wire
    framing (WrappedMessage::BatchRecon under bincode, ark-serialize Vec<F>) and payload validation run without a device."""
    exe = _compile_batch_recon(tmp_path)
    res = subprocess.run([str(exe), "--host-only"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "host-only checks passed" in res.stdout


@pytest.mark.gpu
def test_batch_recon_node_over_fake_network(tmp_path):
    """mpc/tests/batchrecon_test.rs restated: n parties run BatchReconNode over an in-process FakeNetwork, every decode on the GPU;
    single-value and batched arms, t Byzantine senders arriving first."""
    exe = _compile_batch_recon(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all batch reconstruction tests passed" in res.stdout


def _compile_ran_dou_sha(tmp_path):
    exe = tmp_path / "ran_dou_sha_test"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "ran_dou_sha_test.cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_ran_dou_sha_mirror_compiles_and_frames_messages(tmp_path):
    """CPU: the C++ RanDouShaNode mirror (include/hbmpc_ran_dou_sha.hpp) compiles, links, and its WrappedMessage::RanDouSha framing
    round-trips without a device."""
    exe = _compile_ran_dou_sha(tmp_path)
    res = subprocess.run([str(exe), "--host-only"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "host-only checks passed" in res.stdout


@pytest.mark.gpu
def test_ran_dou_sha_node_end_to_end(tmp_path):
    """Random double sharing through the mirror: hyperinvertible-matrix apply (K2, n x n) for every batch in one call, the checkers'
    degree / equality tests (a10) in two calls, outputs that open consistently; a dealer with inconsistent sharings makes every
    checker broadcast ok = false."""
    exe = _compile_ran_dou_sha(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all RanDouSha tests passed" in res.stdout


def _compile_ran_sha(tmp_path):
    exe = tmp_path / "ran_sha_test"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "ran_sha_test.cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_ran_sha_mirror_compiles_and_frames_messages(tmp_path):
    """CPU: the C++ RanShaNode mirror (include/hbmpc_ran_sha.hpp) compiles, links, and its WrappedMessage::RanSha framing round-trips."""
    exe = _compile_ran_sha(tmp_path)
    res = subprocess.run([str(exe), "--host-only"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "host-only checks passed" in res.stdout


@pytest.mark.gpu
def test_ran_sha_node_end_to_end(tmp_path):
    """Random single sharing through the mirror: batched share generation (K1), hyperinvertible apply (K2), the verifiers' robust
    recovery + degree test for every batch column in one call (K3/K4); a dealer of degree t+1 makes every verifier say false."""
    exe = _compile_ran_sha(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all RanSha tests passed" in res.stdout


def _compile_double_share(tmp_path):
    exe = tmp_path / "double_share_test"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "double_share_test.cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_double_share_mirror_compiles_and_frames_messages(tmp_path):
    """CPU: the C++ DoubleShareNode mirror (include/hbmpc_double_share.hpp) compiles, links, and its message framing round-trips."""
    exe = _compile_double_share(tmp_path)
    res = subprocess.run([str(exe), "--host-only"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "host-only checks passed" in res.stdout


@pytest.mark.gpu
def test_double_share_into_ran_dou_sha(tmp_path):
    """DoubleShareNode (two batched share-generation calls per dealer) chained into RanDouShaNode through the C++ mirrors: the collected
    double shares are accepted by every checker and the outputs open consistently."""
    exe = _compile_double_share(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all DoubleShare tests passed" in res.stdout


def _compile_plain_c(tmp_path, src, name, extra=()):
    exe = tmp_path / name
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), *extra, src, "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_group_and_compat_consumers_compile_and_link(tmp_path):
    """CPU: the C consumers of the group entry points and of the reference's own share symbols link against the built library."""
    _compile_plain_c(tmp_path, os.path.join(ROOT, "tests", "host", "group_test.c"), "group_test")
    _compile_plain_c(tmp_path, os.path.join(ROOT, "tests", "host", "secret_share_compat.c"), "secret_share_compat")


REF_TEST = "/root/reference/mpc/src/ffi/tests/secret_share.c"


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="the reference checkout is only present in the build container")
def test_reference_secret_share_c_links_unchanged(tmp_path):
    """The reference's own C test, UNMODIFIED and against the reference's own generated header, compiles and links against
    libhbmpc_b200.so: every share symbol it uses is exported with the reference's signature (it is run on the GPU box through its
    restatement tests/host/secret_share_compat.c -- /root/reference does not travel)."""
    exe = tmp_path / "ref_secret_share"
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    subprocess.run(["gcc", "-O1", "-w", REF_TEST, "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    assert exe.exists()


@pytest.mark.gpu
def test_reference_share_symbols_round_trip(tmp_path):
    exe = _compile_plain_c(tmp_path, os.path.join(ROOT, "tests", "host", "secret_share_compat.c"), "secret_share_compat")
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all round trips passed" in res.stdout


@pytest.mark.gpu
def test_group_calls_from_c(tmp_path):
    """hbmpc_group_*: every visible device (at least two member contexts) -- identical to the single-context results."""
    exe = _compile_plain_c(tmp_path, os.path.join(ROOT, "tests", "host", "group_test.c"), "group_test")
    for members in ("0", "3"):
        res = subprocess.run([str(exe), members], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "identical to the single-context results" in res.stdout


def _compile_cpp(tmp_path, name):
    exe = tmp_path / name
    lib_dir = os.path.join(ROOT, "mpc-protocols_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", name + ".cpp"),
           "-L", lib_dir, "-lhbmpc_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def test_triple_mul_mirror_compiles_and_links(tmp_path):
    _compile_cpp(tmp_path, "triple_mul_test")


@pytest.mark.gpu
def test_triple_generation_and_multiplication_over_fake_network(tmp_path):
    """BASELINE configs[0]: 4 parties, t = 1, Beaver triples from TripleGenNode, 10*10 = 100 and 20*20 = 400 through Multiply
    (mpc/tests/node_test.rs:584-764), five multiplications with a remainder value and a Byzantine party, and n = 16, t = 5."""
    exe = _compile_cpp(tmp_path, "triple_mul_test")
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all triple generation / multiplication tests passed" in res.stdout
