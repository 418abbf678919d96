"""Multi-GPU check (one rank per GPU, torchrun): row-strip sharded min_max / counts finished with one all-reduce,
through both plumbing options — torch.distributed (NCCL) and the library's own ec_comm_* (dlopen'ed NCCL).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, CellValue, MaskedCellBuffer, NoData, sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
L = ec.lib()
ec._lib.check(L.ec_init(local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ec._lib.check(L.ec_set_stream(C.c_void_p(stream.cuda_stream)))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

W = H = int(os.environ.get("EC_SIDE", 16384))
off, ln = sharding.row_strip(W, H, world, rank)

# --- the library's own communicator: unique id from rank 0, shipped with torch.distributed ----------
comm_obj = sharding.Comm.create()
comm = C.c_void_p(comm_obj._h)
if rank == 0:
    print("peer exchange over NVLink (fused one-kernel sharded reductions):", comm_obj.peer_exchange)

ok = True
for ct in (CellType.Float32, CellType.Int16, CellType.UInt8, CellType.Float64, CellType.UInt64, CellType.Int32):
    kw = dict(kind=synth.FULL_BITS) if ct != CellType.Float32 else dict(kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
    strip = synth.device(ct, ln, 0xEC40 + int(ct), index_offset=off, **kw)
    # reference: the same raster reduced on ONE GPU (every rank builds the whole raster for the check)
    whole = synth.device(ct, W * H, 0xEC40 + int(ct), index_offset=0, **kw)
    wmn, wmx = whole.min_max()
    a = sharding.min_max_sharded(strip)                        # torch.distributed NCCL all-reduce(MIN) of 2 x int64
    mn, mx = ec._lib.Value(), ec._lib.Value()
    ec._lib.check(L.ec_buf_min_max_sharded(comm, strip._h, None, C.byref(mn), C.byref(mx)))   # ec_comm NCCL
    good = (a[0].bits, a[1].bits) == (wmn.bits, wmx.bits) == (mn.bits, mx.bits)
    # masked + counts
    nd = NoData.new(ct, whole.get(12345).value())
    ms, mw = MaskedCellBuffer.from_buffer_with_nodata(strip, nd), MaskedCellBuffer.from_buffer_with_nodata(whole, nd)
    b = sharding.min_max_sharded(strip, ms.mask())
    wm = mw.min_max()
    d, n = C.c_size_t(), C.c_size_t()
    ec._lib.check(L.ec_mask_counts_sharded(comm, ms.mask()._h, C.byref(d), C.byref(n)))
    good &= (b[0].bits, b[1].bits) == (wm[0].bits, wm[1].bits) and (d.value, n.value) == mw.counts() == sharding.counts_sharded(ms.mask())
    # statistics (extension): global min/max -> one plan -> per-strip exact sums -> all-gather -> finish; must equal
    # the single-GPU result bit for bit, through both min/max plumbings, masked and unmasked
    def same_stats(x, y):
        return (x.count, x.min.bits, x.max.bits) == (y.count, y.min.bits, y.max.bits) and \
            np.array_equal(np.array([x.mean, x.stddev]).view(np.uint64), np.array([y.mean, y.stddev]).view(np.uint64))
    good_stats = same_stats(sharding.statistics_sharded(strip), whole.statistics())
    good_stats &= same_stats(sharding.statistics_sharded(strip, None, None, comm_obj), whole.statistics())
    good_stats &= same_stats(sharding.statistics_sharded(strip, ms.mask(), None, comm_obj), mw.statistics())
    good_stats &= same_stats(comm_obj.statistics(strip), whole.statistics())               # all in the C ABI: ec_buf_statistics_sharded
    good_stats &= same_stats(comm_obj.statistics(strip, ms.mask()), mw.statistics())
    if rank == 0:
        print(f"{ct}: sharded x{world} == single GPU: {good}  min/max bits {a[0].bits:#x} {a[1].bits:#x} counts {(d.value, n.value)}; "
              f"statistics {good_stats} {whole.statistics()}")
    ok &= good and good_stats
    del whole, strip, ms, mw

# the sharded raster types: NDVI on the Landsat-like fixture layout, every rank uploads only its strip
rng = np.random.default_rng(7)
nir_h = rng.integers(5000, 40000, size=(1031, 512), dtype=np.uint16)
red_h = rng.integers(5000, 40000, size=(1031, 512), dtype=np.uint16)
nir_h[rng.random(nir_h.shape) < 0.01] = 0
nir_s, red_s = sharding.ShardedCellBuffer.from_host(nir_h, comm_obj), sharding.ShardedCellBuffer.from_host(red_h, comm_obj)
with ec.lazy():
    ndvi = (nir_s - red_s) / (nir_s + red_s)
one = (CellBuffer.from_vec(nir_h.reshape(-1)) - CellBuffer.from_vec(red_h.reshape(-1))) / (CellBuffer.from_vec(nir_h.reshape(-1)) + CellBuffer.from_vec(red_h.reshape(-1)))
good = tuple(v.bits for v in ndvi.min_max()) == tuple(v.bits for v in one.min_max())
good &= np.array_equal(ndvi.gather().reshape(-1).view(np.uint64), one.to_vec().view(np.uint64))
nd = NoData.new(CellType.UInt16, 0)
mn_s = sharding.ShardedMaskedCellBuffer.from_host_with_nodata(nir_h, nd, comm_obj)
mr_s = sharding.ShardedMaskedCellBuffer.from_host_with_nodata(red_h, nd, comm_obj)
m_ndvi = (mn_s - mr_s) / (mn_s + mr_s)
m_one = (MaskedCellBuffer.from_vec_with_nodata(nir_h.reshape(-1), nd) - MaskedCellBuffer.from_vec_with_nodata(red_h.reshape(-1), nd))
m_one = m_one / (MaskedCellBuffer.from_vec_with_nodata(nir_h.reshape(-1), nd) + MaskedCellBuffer.from_vec_with_nodata(red_h.reshape(-1), nd))
sharded_counts = m_ndvi.counts()  # collectives: every rank calls them
good &= sharded_counts == m_one.counts() and tuple(v.bits for v in m_ndvi.min_max()) == tuple(v.bits for v in m_one.min_max())
good &= same_stats(m_ndvi.statistics(), m_one.statistics()) and same_stats(ndvi.statistics(), one.statistics())
if rank == 0:
    print(f"ShardedCellBuffer / ShardedMaskedCellBuffer NDVI (1031 x 512, ragged strips) == single GPU: {good}, counts {sharded_counts}")
ok &= good

# fewer rows than ranks (the first strips are empty): every rank still takes part with the raster's cell type
tiny_h = rng.random((1, 256)).astype(np.float32)  # one row: only the last rank holds cells
ts = sharding.ShardedCellBuffer.from_host(tiny_h, comm_obj)
tr = (ts * 2.0 - ts) / 3.0
assert tr.strip.cell_type() == CellType.Float64, tr.strip.cell_type()
one_t = (CellBuffer.from_vec(tiny_h.reshape(-1)) * 2.0 - CellBuffer.from_vec(tiny_h.reshape(-1))) / 3.0
good = tuple(v.bits for v in tr.min_max()) == tuple(v.bits for v in one_t.min_max())
good &= same_stats(tr.statistics(), one_t.statistics())
good &= tuple(v.bits for v in (-ts).min_max()) == tuple(v.bits for v in (-CellBuffer.from_vec(tiny_h.reshape(-1))).min_max())
tm = sharding.ShardedMaskedCellBuffer.from_host_with_nodata(tiny_h, NoData.new(CellType.Float32, float(tiny_h[0, 3])), comm_obj)
tmr = (tm + tm) * 0.5
tm_one = MaskedCellBuffer.from_vec_with_nodata(tiny_h.reshape(-1), NoData.new(CellType.Float32, float(tiny_h[0, 3])))
tm_one = (tm_one + tm_one) * 0.5
good &= tmr.counts() == tm_one.counts() and tuple(v.bits for v in tmr.min_max()) == tuple(v.bits for v in tm_one.min_max())
if rank == 0:
    print(f"one-row raster over {world} ranks (empty strips keep the logical cell type): {good}")
ok &= good

# latency of the two ways to finish a sharded f32 min_max (device events around 50 calls)
strip = synth.device(CellType.Float32, ln, 0xEC40, index_offset=off, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
for label, fn in (("torch.distributed NCCL all-reduce", lambda: sharding.min_max_sharded(strip)), ("ec_comm (peer exchange in the kernel)" if comm_obj.peer_exchange else "ec_comm (NCCL)", lambda: comm_obj.min_max(strip))):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"sharded f32 min_max of {W}x{H} over {world} GPUs, {label}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per call")
del strip

t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
comm_obj.close()
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_OK" if int(t.item()) else "MULTI_GPU_FAILED")
sys.exit(0 if int(t.item()) else 1)
