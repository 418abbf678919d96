"""Timeline of ONE sharded min_max call on N GPUs (one rank per GPU, torchrun): where the microseconds between
"the shard kernel alone" and "the call returned" go. The reduction kernel stamps %globaltimer at its stages
(ec_set_reduce_trace), the host stamps CLOCK_REALTIME around the call; both are put on one axis (offset calibrated on
every rank) and printed relative to the first rank entering the call.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 tools/reduce_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellType, sharding, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
L = ec.lib()
ec._lib.check(L.ec_init(local))
L.ec_set_min_max_cache(0)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
side = int(os.environ.get("EC_SIDE", 32768))
off, ln = sharding.row_strip(side, side, world, rank)
strip = synth.device(CellType.Float32, ln, 0xEC40, index_offset=off, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
comm = sharding.Comm.create() if world > 1 else None
mn, mx = ec._lib.Value(), ec._lib.Value()
fn = L.ec_buf_min_max_sharded if comm is not None else L.ec_buf_min_max
args = (comm._h, strip._h, None, C.byref(mn), C.byref(mx)) if comm is not None else (strip._h, None, C.byref(mn), C.byref(mx))
for _ in range(30):
    fn(*args)
L.ec_set_reduce_trace(1)
rows = []
for call in range(12):
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    fn(*args)
    t = (C.c_uint64 * 12)()
    ec._lib.check(L.ec_reduce_trace_get(t))
    v = np.array(list(t), dtype=np.uint64).view(np.int64).copy()
    v[3:9] = np.where(v[3:9] != 0, v[3:9] + v[9], 0)  # %globaltimer -> CLOCK_REALTIME axis
    rows.append(v[:9])
# back-to-back calls (no barrier in between): what a loop of calls sees
back = []
for call in range(12):
    fn(*args)
    t = (C.c_uint64 * 12)()
    ec._lib.check(L.ec_reduce_trace_get(t))
    v = np.array(list(t), dtype=np.uint64).view(np.int64).copy()
    v[3:9] = np.where(v[3:9] != 0, v[3:9] + v[9], 0)
    back.append(v[:9])
L.ec_set_reduce_trace(0)
allrows = [None] * world
if world > 1:
    dist.all_gather_object(allrows, (rows, back))
else:
    allrows = [(rows, back)]
if rank == 0:
    names = ["enter", "launched", "seen", "gpu_start", "last_cta", "folded", "sent", "received", "published"]
    order = [0, 1, 3, 4, 5, 6, 7, 8, 2]
    for label, idx in (("after a barrier", 0), ("back to back", 1)):
        print(f"# sharded f32 {side}^2 min_max over {world} GPU(s), {label}: microseconds after the first rank entered the call (calls 2..11, median over calls)")
        print("rank  " + "  ".join(f"{names[k]:>9s}" for k in order))
        per_call = []
        for call in range(2, 12):
            m = np.stack([np.array(allrows[r][idx][call]) for r in range(world)])  # world x 9
            t0 = m[:, 0].min()
            rel = np.where(m != 0, (m - t0) / 1e3, np.nan)
            per_call.append(rel)
        med = np.nanmedian(np.stack(per_call), axis=0)
        for r in range(world):
            print(f"{r:4d}  " + "  ".join(f"{med[r][k]:9.1f}" for k in order))
        kern = med[:, 4] - med[:, 3]
        print(f"summary: kernel streaming (gpu_start -> last_cta) {np.nanmedian(kern):.1f} us [{np.nanmin(kern):.1f} .. {np.nanmax(kern):.1f}]; "
              f"enter -> gpu_start {np.nanmedian(med[:, 3] - med[:, 0]):.1f} us; start skew across ranks {np.nanmax(med[:, 3]) - np.nanmin(med[:, 3]):.1f} us; "
              f"last_cta -> folded {np.nanmedian(med[:, 5] - med[:, 4]):.1f} us; folded -> received (exchange incl. waiting for the slowest rank) {np.nanmedian(med[:, 7] - med[:, 5]):.1f} us; "
              f"received/folded -> published {np.nanmedian(med[:, 8] - np.where(np.isnan(med[:, 7]), med[:, 5], med[:, 7])):.1f} us; published -> seen by the host {np.nanmedian(med[:, 2] - med[:, 8]):.1f} us; "
              f"whole call (first enter -> last seen) {np.nanmax(med[:, 2]):.1f} us")
        print()
if comm is not None:
    comm.close()
if world > 1:
    dist.destroy_process_group()
