"""Host<->device copy bandwidth on this box (development tool): pinned vs pageable, each way and both at once."""
import time

import torch

dev = torch.device("cuda", 0)
n = 1 << 30  # 1 GiB
d0 = torch.empty(n, dtype=torch.uint8, device=dev)
d1 = torch.empty(n, dtype=torch.uint8, device=dev)
hp0 = torch.empty(n, dtype=torch.uint8).pin_memory()
hp1 = torch.empty(n, dtype=torch.uint8).pin_memory()
hpage = torch.empty(n, dtype=torch.uint8)
hpage.fill_(1)


def t(fn, iters=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("H2D pinned   GB/s", n / t(lambda: d0.copy_(hp0, non_blocking=True)) / 1e9)
print("D2H pinned   GB/s", n / t(lambda: hp0.copy_(d0, non_blocking=True)) / 1e9)
print("H2D pageable GB/s", n / t(lambda: d0.copy_(hpage)) / 1e9)
print("D2H pageable GB/s", n / t(lambda: hpage.copy_(d0)) / 1e9)


def both():
    with torch.cuda.stream(s1):
        d0.copy_(hp0, non_blocking=True)
    with torch.cuda.stream(s2):
        hp1.copy_(d1, non_blocking=True)


print("H2D+D2H concurrent, GB/s each way", n / t(both) / 1e9)
for sz in (1 << 20, 1 << 24, 1 << 26, 1 << 28):
    print(f"D2H pinned {sz >> 20} MiB chunks GB/s", sz / t(lambda: hp0[:sz].copy_(d0[:sz], non_blocking=True), 20) / 1e9)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv; nvidia-smi topo -m | head -5; numactl -H 2>/dev/null | head -5; lscpu | grep -i -E 'numa|socket|model name'", shell=True, capture_output=True, text=True).stdout)
