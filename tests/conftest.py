import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def build_sim() -> str:
    """Build the kernel-logic simulator build of the engine (tests/cusim, CPU fibers).
    TEST INFRASTRUCTURE: never shipped, never loaded by helib_b200 itself."""
    out = os.path.join(ROOT, "tests", "cusim", "libhelib_b200_sim.so")
    srcs = [os.path.join(ROOT, "helib_b200", "csrc", f) for f in ("hb_engine.cu", "hb_chain.cpp", "hb_device.cuh", "hb_device_v1.cuh", "hb_device_v2.cuh", "hb_device_gen.cuh")]
    srcs.append(os.path.join(ROOT, "tests", "cusim", "cusim.h"))
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call([
            "g++", "-O2", "-std=c++17", "-x", "c++", "-DHB_SIM", "-DCUSIM_IMPLEMENTATION",
            "-I" + os.path.join(ROOT, "tests", "cusim"), "-I" + os.path.join(ROOT, "helib_b200", "csrc"),
            "-shared", "-fPIC", srcs[0], srcs[1], "-o", out])
    return out


@pytest.fixture(scope="session")
def sim_lib():
    from helib_b200.engine import load_library
    return load_library(build_sim())


@pytest.fixture(scope="session")
def cuda_lib():
    from helib_b200.engine import load_library
    lib = load_library()  # raises if libhelib_b200.so is not built: no silent fallback
    if lib.hb_device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return lib


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    import orc
    orc.build()


@pytest.fixture(scope="session", autouse=True)
def _product_built():
    """Make sure helib_b200/libhelib_b200.so exists and is current (nvcc cross-compiles without a GPU);
    a no-op when __graft_entry__.build() already ran."""
    from helib_b200.build import build_library
    build_library()
