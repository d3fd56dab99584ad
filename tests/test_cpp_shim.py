"""The header-only C++ mirror of helib::DoubleCRT (include/helib_b200_doublecrt.hpp) compiles against
the C ABI and links the product library; on a GPU it runs the reference's unit-level properties."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def build_exe(name="test_shim", sim=False):
    """sim=True links the same test against the CPU kernel-logic simulator build of the library (tests/cusim):
    the mirror's host logic and the kernels' arithmetic are then exercised without a GPU."""
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    exe = os.path.join(ROOT, "tests", "cpp", name + ("_sim" if sim else ""))
    if sim:
        from conftest import build_sim
        libpath = build_sim()
    else:
        libpath = os.path.join(ROOT, "helib_b200", "libhelib_b200.so")
    deps = [src, os.path.join(ROOT, "include", "helib_b200_doublecrt.hpp"), os.path.join(ROOT, "include", "helib_b200_ctxt.hpp"), libpath]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, libpath,
                               "-Wl,-rpath," + os.path.dirname(libpath)])
    return exe


def test_shim_compiles_links_and_refuses_without_gpu():
    from helib_b200 import load_library
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    if load_library().hb_device_count() <= 0:
        assert r.returncode == 3, r.stdout + r.stderr      # no device: loud refusal, no CPU fallback
    else:
        assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_shim_runs_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "shim OK" in r.stdout, r.stdout + r.stderr
    _check_randomize_line(r.stdout)


def test_ctxt_mirror_compiles_and_refuses_without_gpu():
    from helib_b200 import load_library
    exe = build_exe("test_ctxt")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == (3 if load_library().hb_device_count() <= 0 else 0), r.stdout + r.stderr


@pytest.mark.gpu
def test_ctxt_multiplyBy_decrypts_on_gpu():
    """helib::Ctxt mirror: BGV encrypt -> multiplyBy (noise-driven prime sets, device norms) -> decrypt == product."""
    exe = build_exe("test_ctxt")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ctxt OK" in r.stdout, r.stdout + r.stderr


def _splitmix_bytes():
    state = [0x1234567]

    def get(n):
        out = bytearray()
        while len(out) < n:
            state[0] = (state[0] + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
            z = state[0]
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
            z ^= z >> 31
            out += z.to_bytes(8, "little")
        return bytes(out[:n])
    return get


def _check_randomize_line(stdout):
    """DoubleCRT::randomize in the mirror vs the oracle's restatement on the same byte stream (FNV-1a hash per row)."""
    import pyoracle as po
    line = [ln for ln in stdout.splitlines() if ln.startswith("randomize:")][0]
    got = {int(t.split(":")[0]): int(t.split(":")[1]) for t in line.split()[1:]}
    ch = po.build_mod_chain(4096, 257, 1, 120, 2)
    rows = po.randomize_rows(ch, ch.ctxt, _splitmix_bytes())
    for i, row in rows.items():
        h = 1469598103934665603
        for v in row:
            h = ((h ^ v) * 1099511628211) & (2 ** 64 - 1)
        assert got[i] == h, f"row {i}"


def test_shim_logic_on_simulator():
    """The DoubleCRT mirror's unit-level properties with the kernels compiled for the CPU simulator."""
    r = subprocess.run([build_exe("test_shim", sim=True)], capture_output=True, text=True)
    assert r.returncode == 0 and "shim OK" in r.stdout, r.stdout + r.stderr
    _check_randomize_line(r.stdout)


def test_ctxt_encrypt_multiply_decrypt_on_simulator():
    """PubKey::Encrypt -> Ctxt::multiplyBy -> SecKey::Decrypt through the mirror, kernels on the CPU simulator."""
    r = subprocess.run([build_exe("test_ctxt", sim=True)], capture_output=True, text=True)
    assert r.returncode == 0 and "ctxt OK" in r.stdout, r.stdout + r.stderr


def test_ckks_multiply_through_mirror_on_simulator():
    """CKKS branches of the Ctxt mirror (scaling factors, CKKS interval for the common prime set, relin_CKKS_adjust)."""
    r = subprocess.run([build_exe("test_ckks", sim=True)], capture_output=True, text=True)
    assert r.returncode == 0 and "ckks OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_ckks_multiply_through_mirror_on_gpu():
    r = subprocess.run([build_exe("test_ckks")], capture_output=True, text=True)
    assert r.returncode == 0 and "ckks OK" in r.stdout, r.stdout + r.stderr


def test_raw_mod_switch_through_mirror_on_simulator():
    """Ctxt::rawModSwitch (powerful basis, general and power-of-two m): the switched ciphertext decrypts mod q to the same plaintext."""
    r = subprocess.run([build_exe("test_rawmodswitch", sim=True)], capture_output=True, text=True)
    assert r.returncode == 0 and "rawmodswitch OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_raw_mod_switch_through_mirror_on_gpu():
    r = subprocess.run([build_exe("test_rawmodswitch")], capture_output=True, text=True)
    assert r.returncode == 0 and "rawmodswitch OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("seed", [1, 2])
def test_ctxt_random_walk_on_simulator(seed):
    """Random sequences of +=, -=, multiplyBy, multByConstant, addConstant, negate, smartAutomorph and
    dropSmallAndSpecialPrimes over a pool of BGV ciphertexts (tests/cpp/test_ctxt_walk.cpp): after every step the ciphertext
    decrypts to the plaintext mirror and the tracked noise bound dominates the measured noise."""
    r = subprocess.run([build_exe("test_ctxt_walk", sim=True), str(seed), "40"], capture_output=True, text=True)
    assert r.returncode == 0 and "walk OK" in r.stdout, r.stdout + r.stderr


def test_ckks_random_walk_on_simulator():
    """The same for the CKKS branches (tests/cpp/test_ckks_walk.cpp): +=, -= across different scaling factors, multiplyBy,
    negate, dropSmallAndSpecialPrimes; decode == real-valued mirror within the tracked noise bound, which in turn stays
    small against the tracked plaintext magnitude."""
    r = subprocess.run([build_exe("test_ckks_walk", sim=True), "1", "40"], capture_output=True, text=True)
    assert r.returncode == 0 and "ckks walk OK" in r.stdout, r.stdout + r.stderr
