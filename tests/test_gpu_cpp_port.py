"""Runs the C++ port of the reference's own tests (tests/cpp/test_reference_port.cpp) — the host
mirror include/erased_cells.hpp over the C ABI — on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_reference_port():
    exe = os.path.join(ROOT, "tests", "cpp", "build", "test_reference_port")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_plain_c_abi_smoke():
    exe = os.path.join(ROOT, "tests", "cpp", "build", "abi_smoke_c")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "c abi ok" in r.stdout, r.stdout + r.stderr


def test_cpp_port_builds_on_cpu():
    """the host mirror compiles and links against the C ABI without a GPU"""
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "build", "test_reference_port"))
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "build", "abi_smoke_c"))  # the header is valid C99 (-pedantic -Werror)


def test_host_copy_pool_on_cpu():
    """CPU: the thread pool that moves pageable memory to / from pinned staging (csrc/ec_hostcopy.hpp) — ragged sizes,
    every thread count, concurrent callers, and a process exit while the workers sleep."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s", "build/test_hostcopy"], check=True)
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "build", "test_hostcopy")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOSTCOPY_OK" in r.stdout, r.stdout + r.stderr
