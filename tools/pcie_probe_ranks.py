"""Host <-> device link bandwidth with N ranks copying AT THE SAME TIME (one rank per GPU, torchrun) — the ceiling of
bench.py's end-to-end leg. Bare pinned-memory cudaMemcpyAsync (torch copy_), no library code on the path:
D2H alone, H2D alone, both directions at once; per rank (slowest, fastest) and the box total.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 tools/pcie_probe_ranks.py
    python tools/pcie_probe_ranks.py            # N = 1
"""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host = dist.new_group(backend="gloo")  # barriers on the host: nothing of ours runs on the GPUs while they copy

block = 512 << 20
dev_a = torch.empty(block, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(block, dtype=torch.uint8, device="cuda")
h_out = torch.empty(block, dtype=torch.uint8).pin_memory()
h_in = torch.empty(block, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
side = torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(group=host)


def gather(x):
    if world == 1:
        return [x]
    out = [None] * world
    dist.all_gather_object(out, x, group=host)
    return out


def timed(fn, reps=8):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rates = gather(reps * block / dt / 1e9)
    barrier()
    return rates


def d2h():
    h_out.copy_(dev_a, non_blocking=True)


def h2d():
    dev_a.copy_(h_in, non_blocking=True)


def duplex():
    h_out.copy_(dev_a, non_blocking=True)
    with torch.cuda.stream(side):
        dev_b.copy_(h_in, non_blocking=True)


lines = []
for name, fn in (("D2H", d2h), ("H2D", h2d), ("D2H + H2D at once (each way)", duplex)):
    r = timed(fn)
    lines.append(f"{name:32s} per rank min {min(r):6.2f}  max {max(r):6.2f} GB/s   box total {sum(r):7.2f} GB/s   ranks {[round(x, 1) for x in r]}")
if rank == 0:
    print(f"# pinned host <-> HBM, cudaMemcpyAsync, {block >> 20} MiB blocks, {world} rank(s) copying at the same time")
    print("\n".join(lines))
    q = "nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader | head -8; lscpu | grep -i -E 'model name|socket|numa node\\(s\\)|^CPU\\(s\\)'; free -g | head -2"
    print(subprocess.run(q, shell=True, capture_output=True, text=True).stdout)
if world > 1:
    dist.destroy_process_group()
