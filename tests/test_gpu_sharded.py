"""One process, several logical devices: CellBuffer / Mask handles that are row-strip sharded behind the C ABI
(ec_init_devices). On a one-GPU box the same CUDA device is listed three times — same code path, the strips just share a
GPU; on a multi-GPU box the strips also go to distinct GPUs and the GPU-to-GPU finishes (peer mailboxes, NCCL) run.
The check itself is tools/sharded_check.py (every result bit for bit against the CPU oracle); it needs its own process
because the device list is fixed at the library's first use."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_check(devices: str, extra_env=None):
    env = dict(os.environ, EC_DEVICES=devices, EC_SHARD_MIN_CELLS="4096", **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
def test_sharded_handles_three_logical_devices_on_one_gpu():
    run_check("0,0,0")


@pytest.mark.gpu
def test_sharded_handles_under_the_red_zone_guard():
    """the same with 256-byte red zones around every device block (EC_DEBUG_GUARD): strips, views and re-partitioning
    copies must stay inside their blocks"""
    out = run_check("0,0", {"EC_DEBUG_GUARD": "1"})
    assert "SHARDED_OK" in out


@pytest.mark.gpu
def test_sharded_handles_distinct_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    out = run_check(",".join(str(i) for i in range(min(n, 8))))
    assert "finish_modes=[0, 1, 2]" in out, out  # host fold, peer mailboxes and NCCL all ran


@pytest.mark.gpu
def test_reference_test_suites_on_sharded_handles():
    """The reference's own tests — the C++ port over include/erased_cells.hpp and the Python port — with EVERY non-empty
    buffer and mask sharded over two logical devices (threshold 1 cell): the handles must behave exactly like plain ones."""
    env = dict(os.environ, EC_DEVICES="0,0", EC_SHARD_MIN_CELLS="1")
    exe = os.path.join(ROOT, "tests", "cpp", "build", "test_reference_port")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and " 0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_reference_kat.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_cpp_drives_two_strips_in_one_process():
    """tests/cpp/test_sharded.cpp: the sharded handles through the C++ mirror, CUDA device 0 listed twice"""
    exe = os.path.join(ROOT, "tests", "cpp", "build", "test_sharded")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
