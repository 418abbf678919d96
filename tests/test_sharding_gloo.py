"""The N>1 host path on CPU: world_size-2 gloo processes shard a raster into row strips, each reduces
its strip (oracle stands in for the shard kernel here — no GPU), packs (min, max) into the product's
order-preserving keys, finishes with ONE all-reduce(MIN) and decodes. Result must equal the oracle's
min_max of the whole raster for every cell type, including NaN/inf/-0.0 under total order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from erased_cells_b200 import CellType, CellValue, sharding, synth
    from oracle import oracle as orc
    ok = True
    width, height = 256, 37  # ragged rows: the last strip takes the remainder
    for ct in CellType:
        full = synth.host(ct, width * height, 0xEC40 + int(ct))
        if ct.dtype.kind == "f":
            full[[5, 900, 4000, 7000]] = np.array([np.nan, -np.nan, np.inf, -0.0], dtype=ct.dtype)
        off, ln = sharding.row_strip(width, height, world, rank)
        assert off % 128 == 0
        strip = full[off:off + ln]
        mn, mx = orc.tight_min_max(strip)
        keys = torch.from_numpy(sharding.keys_of(CellValue(ct, mn.numpy()), CellValue(ct, mx.numpy())))
        gmn, gmx = sharding.finish_min_max(ct, keys)
        wmn, wmx = orc.tight_min_max(full)
        ok &= (gmn.bits, gmx.bits) == (wmn.bits, wmx.bits)
        # an empty strip contributes only the seeds
        e_mn, e_mx = orc.tight_min_max(strip[:0])
        k2 = torch.from_numpy(sharding.keys_of(CellValue(ct, e_mn.numpy()), CellValue(ct, e_mx.numpy()))) if rank else keys.clone()
        if rank == 0:
            k2 = torch.from_numpy(sharding.keys_of(CellValue(ct, mn.numpy()), CellValue(ct, mx.numpy())))
        a, b = sharding.finish_min_max(ct, k2)
        r0 = orc.tight_min_max(full[: sharding.row_strip(width, height, world, 0)[1]])
        ok &= (a.bits, b.bits) == (r0[0].bits, r0[1].bits)
    # statistics (extension): global min/max -> shared plan -> per-strip exact sums -> all-gather -> finish
    for ct in CellType:
        full = synth.host(ct, width * height, 0xEC70 + int(ct), kind=synth.REAL_RANGE, lo=-300.0 if ct.is_signed() else 3.0, hi=120.0)
        valid = synth.host(CellType.UInt8, width * height, 0xEC7F) < 200
        off, ln = sharding.row_strip(width, height, world, rank)
        strip, vstrip = full[off:off + ln], valid[off:off + ln]
        mn, mx = orc.tight_min_max(strip, vstrip)
        keys = torch.from_numpy(sharding.keys_of(CellValue(ct, mn.numpy()), CellValue(ct, mx.numpy())))
        gmn, gmx = sharding.finish_min_max(ct, keys)
        kind, pivot, exp2 = sharding.statistics_plan(gmn, gmx)
        ok &= kind == sharding.ST_REGULAR
        raw = orc.moments_raw(strip, vstrip, pivot, exp2)
        words = np.zeros(9, dtype=np.uint64)
        words[0] = raw[0]
        for k in range(4):
            v = raw[1 + k] & ((1 << 128) - 1)
            words[1 + 2 * k], words[2 + 2 * k] = v & (2 ** 64 - 1), v >> 64
        got = sharding.gather_statistics(words, gmn, gmx)
        want = orc.statistics(full, valid)
        ok &= got.count == want["count"]
        ok &= (np.float64(got.mean).view(np.uint64), np.float64(got.stddev).view(np.uint64)) == \
              (np.float64(want["mean"]).view(np.uint64), np.float64(want["stddev"]).view(np.uint64))
    # strips tile the raster exactly
    lens = [sharding.row_strip(width, height, world, g) for g in range(world)]
    ok &= sum(l for _, l in lens) == width * height and all(lens[g][0] + lens[g][1] == lens[g + 1][0] for g in range(world - 1))
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_min_max_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(150) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == 1
