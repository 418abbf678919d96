"""Single-process sharding check: ONE process drives G logical devices (EC_DEVICES, e.g. "0,0,0" on a one-GPU box —
the same CUDA device listed three times — or "0,1" on two GPUs) and every large CellBuffer / Mask lives as G row
strips behind the ordinary handles. Every result is compared bit for bit with the CPU oracle (test infrastructure).

    EC_DEVICES=0,0,0 EC_SHARD_MIN_CELLS=4096 python tools/sharded_check.py

Prints SHARDED_OK on success (run by tests/test_gpu_sharded.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erased_cells_b200 as ec
from erased_cells_b200 import CellBuffer, CellType, Mask, MaskedCellBuffer, NoData, synth
from oracle import oracle as orc

orc.build()
devs = [int(x) for x in os.environ.get("EC_DEVICES", "0,0").split(",")]
G = ec.init_devices(devs)
assert G == len(devs) and ec.device_count() == G
distinct = len(set(devs)) == len(devs)
THR = int(os.environ.get("EC_SHARD_MIN_CELLS", "4096"))
ec.set_shard_min_cells(THR)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def same(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(bits(a), bits(b)), what


def cells(ct, n, seed):
    return synth.host(ct, n, seed)  # full bit range: NaN / inf / -0.0 / MIN / MAX included


n = 3 * 40960 + 777  # ragged: the last strip takes the remainder
finish_modes = [ec.FINISH_HOST] + ([ec.FINISH_PEER, ec.FINISH_NCCL] if distinct and G > 1 else [])

# ---- construction, layout, round trip ----------------------------------------------------------------------------
a_h, b_h = cells(CellType.UInt8, n, 1), cells(CellType.UInt16, n, 2)
a, b = CellBuffer.from_vec(a_h), CellBuffer.from_vec(b_h)
assert a.shard_count() == G and b.shard_count() == G, (a.shard_count(), G)
sh = a.shards()
assert sh[0][2] == 0 and sum(s[3] for s in sh) == n and all(s[2] % 128 == 0 for s in sh)
assert [s[0] for s in sh] == list(range(G)) and [s[1] for s in sh] == devs
same(a.to_vec(), a_h, "round trip")
small = CellBuffer.from_vec(a_h[:1000])
assert small.shard_count() == 0  # below the threshold: one GPU
assert a.device_ptr() == 0       # a sharded buffer has no single device pointer

# ---- pageable host memory of >= 16 MiB: the staged copy, chunks alternating between the strips ----------------------------
big_n = (17 << 20) + 12345  # u16: 34 MiB, ragged against the 8 MiB chunks and the 128-cell strips
big_h = synth.host(CellType.UInt16, big_n, 0x51A6)
for threads in (0, 1, 3):
    prev = ec.set_host_copy_threads(threads)
    big = CellBuffer.from_vec(big_h)
    assert big.shard_count() == G
    same(big.to_vec(), big_h, f"pageable round trip, {threads} copy threads")
    same(CellBuffer.from_vec(big_h, wait=False).wait().to_vec(), big_h, f"pageable async round trip, {threads} copy threads")
    del big
    ec.set_host_copy_threads(prev)

# ---- maps: all four ops on mixed types, scalar, neg, convert, clone -------------------------------------------------
for op in range(4):
    same((a._bin(op, b)).to_vec(), orc.tight_binary(op, a_h, b_h), f"binary {op}")
r = a / b * 0.5
assert r.shard_count() == G and r.cell_type() == CellType.Float64
want = orc.tight_scalar(orc.MUL, orc.tight_binary(orc.DIV, a_h, b_h), orc.value(orc.Float64, 0.5))
same(r.to_vec(), want, "a / b * 0.5")
same((-a).to_vec(), orc.neg(a_h), "neg")
same(a.convert(CellType.Int32).to_vec(), orc.tight_convert(a_h, orc.Int32), "convert")
try:
    b.convert(CellType.UInt8)
    raise AssertionError("narrowing convert did not fail")
except ec.NarrowingError as e:
    assert (e.src, e.dst) == (int(CellType.UInt16), int(CellType.UInt8))
same(a.clone().to_vec(), a_h, "clone")
for ct in CellType:  # every cell type through a sharded op and a sharded reduction
    x_h = cells(ct, n, 0x40 + int(ct))
    x = CellBuffer.from_vec(x_h)
    same((x * 3).to_vec(), orc.tight_scalar(orc.MUL, x_h, orc.value(orc.Int32, 3)), f"scalar {ct}")
    for mode in finish_modes:
        ec.set_shard_finish(mode)
        mn, mx = x.min_max()
        omn, omx = orc.tight_min_max(x_h)
        assert (mn.bits, mx.bits) == (omn.bits, omx.bits), (ct, mode)
    ec.set_shard_finish(ec.FINISH_HOST)

# ---- zip of different lengths: the result's partition differs from the operands' (strips are re-partitioned GPU to GPU) ----
short_h = cells(CellType.Int16, n - 30000, 3)
short = CellBuffer.from_vec(short_h)
same((a - short).to_vec(), orc.tight_binary(orc.SUB, a_h[: n - 30000], short_h), "zip truncation, sharded result")
same((short - a).to_vec(), orc.tight_binary(orc.SUB, short_h, a_h[: n - 30000]), "zip truncation, other side")
same((a + small).to_vec(), orc.tight_binary(orc.ADD, a_h[:1000], a_h[:1000]), "sharded (op) plain -> plain result")
same((small * b).to_vec(), orc.tight_binary(orc.MUL, a_h[:1000], b_h[:1000]), "plain (op) sharded")

# ---- views, single cells, extend ---------------------------------------------------------------------------------------
v = a.view(40960 - 128, 2 * 40960)  # spans strip boundaries
same(v.to_vec(), a_h[40960 - 128: 40960 - 128 + 2 * 40960], "view across strips")
same((v + v).to_vec(), orc.tight_binary(orc.ADD, v.to_vec(), v.to_vec()), "op on a sharded view")
for i in (0, 1, sh[1][2] - 1, sh[1][2], n - 1):
    assert a.get(i).bits == int(a_h[i]), i
c = a.clone()
c.put(sh[-1][2] + 5, np.uint8(77))
a2 = a_h.copy(); a2[sh[-1][2] + 5] = 77
same(c.to_vec(), a2, "put into the last strip")
same(a.to_vec(), a_h, "clone is deep")
c.extend(np.arange(10, dtype=np.uint8))
same(c.to_vec(), np.concatenate([a2, np.arange(10, dtype=np.uint8)]), "extend")
assert c.len() == n + 10
same((c + c).to_vec(), orc.tight_binary(orc.ADD, c.to_vec(), c.to_vec()), "op after extend (non-canonical partition)")
same((c - a).to_vec(), orc.tight_binary(orc.SUB, a2, a_h), "extended (op) canonical")

# ---- Ord / Eq ---------------------------------------------------------------------------------------------------------------
assert a == a.clone() and a.cmp(c) == orc.buffer_cmp(a_h, c.to_vec()) and c.cmp(a) == orc.buffer_cmp(c.to_vec(), a_h)
d_h = a_h.copy(); d_h[n - 3] ^= 1
assert a.cmp(CellBuffer.from_vec(d_h)) == orc.buffer_cmp(a_h, d_h) != 0
assert a.cmp(small) == orc.buffer_cmp(a_h, a_h[:1000])

# ---- masks ------------------------------------------------------------------------------------------------------------------
rng = np.random.default_rng(5)
m1_h, m2_h = rng.random(n) < 0.7, rng.random(n) < 0.4
m1, m2 = Mask.new(m1_h), Mask.new(m2_h)
assert ec.lib().ec_mask_shard_count(m1._h) == G
assert np.array_equal(m1.to_vec(), m1_h)
assert np.array_equal((m1 & m2).to_vec(), m1_h & m2_h) and np.array_equal((m1 | m2).to_vec(), m1_h | m2_h)
assert np.array_equal((~m1).to_vec(), ~m1_h)
assert m1.counts() == (int(m1_h.sum()), int(n - m1_h.sum())) and (m1 & m2).counts() == (int((m1_h & m2_h).sum()), int(n - (m1_h & m2_h).sum()))
assert (~m1).counts()[0] == int((~m1_h).sum())
assert not m1.all(True) and Mask.fill(n, True).all(True) and Mask.fill(n, False).all(False)
assert np.array_equal(m1.slice(40960, 50000).to_vec(), m1_h[40960:90960])
short_m = Mask.new(m2_h[: n - 30001])
assert np.array_equal((m1 & short_m).to_vec(), (m1_h[: n - 30001] & m2_h[: n - 30001])), "mask zip truncation"
assert (m1 & short_m).counts()[0] == int((m1_h[: n - 30001] & m2_h[: n - 30001]).sum())
for i in (0, 31, 32, sh[1][2], n - 1):
    assert m1.get(i) == bool(m1_h[i])
mc = m1.clone()
before = mc.counts()[0]
flip = sh[1][2] + 3
mc.put(flip, not m1_h[flip])
assert mc.counts()[0] == before + (1 if not m1_h[flip] else -1), "put keeps the cached count right"
assert m1.counts()[0] == before and m1.get(flip) == bool(m1_h[flip]), "a clone shares words until it is mutated"
assert m1 == m1.clone() and m1.cmp(mc) == (-1 if not m1_h[flip] else 1) and mc.cmp(m1) == (1 if not m1_h[flip] else -1)
mc.extend([True, False, True])
assert mc.len() == n + 3 and mc.counts()[0] == before + (1 if not m1_h[flip] else -1) + 2

# ---- MaskedCellBuffer: NoData mask, masked chain, masked reductions, fill ------------------------------------------------------
i1_h = synth.host(CellType.Int16, n, 0x31, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
i2_h = synth.host(CellType.Int16, n, 0x32, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
nd = NoData.default(CellType.Int16)
ma, mb = MaskedCellBuffer.from_vec_with_nodata(i1_h, nd), MaskedCellBuffer.from_vec_with_nodata(i2_h, nd)
assert ma.buffer().shard_count() == G and ec.lib().ec_mask_shard_count(ma.mask()._h) == G
wm1, wm2 = orc.mask_from_nodata(i1_h, orc.ND_DEFAULT), orc.mask_from_nodata(i2_h, orc.ND_DEFAULT)
assert np.array_equal(ma.mask().to_vec(), wm1)
res = (ma - mb) * 0.0001
wd = orc.tight_scalar(orc.MUL, orc.tight_binary(orc.SUB, i1_h, i2_h), orc.value(orc.Float64, 0.0001))
wmask = orc.mask_and(wm1, wm2)
same(res.buffer().to_vec(), wd, "masked chain data")
assert np.array_equal(res.mask().to_vec(), wmask) and res.counts() == orc.mask_counts(wmask)
for mode in finish_modes:
    ec.set_shard_finish(mode)
    mn, mx = res.min_max()
    omn, omx = orc.tight_min_max(wd, wmask)
    assert (mn.bits, mx.bits) == (omn.bits, omx.bits), mode
ec.set_shard_finish(ec.FINISH_HOST)
filled = res.to_vec_with_nodata(NoData.new(CellType.Float64, -9999.0))
same(filled, orc.fill_nodata(wd, wmask, orc.Float64, orc.ND_VALUE, orc.value(orc.Float64, -9999.0)), "fill_nodata")
neg = -ma
same(neg.buffer().to_vec(), orc.neg(i1_h), "masked neg")
assert np.array_equal(neg.mask().to_vec(), wm1)

# ---- statistics (extension): integer one-pass route and the FP64 window route, masked and not ------------------------------------
for buf, host, mask_h, mask in ((ma.buffer(), i1_h, wm1, ma.mask()), (res.buffer(), wd, wmask, res.mask()), (ma.buffer(), i1_h, None, None)):
    st = MaskedCellBuffer(buf, mask).statistics() if mask is not None else buf.statistics()
    w = orc.statistics(host, mask_h)
    assert (st.count, st.min.bits, st.max.bits) == (w["count"], w["min"].bits, w["max"].bits), "statistics"
    assert np.array_equal(np.array([st.mean, st.stddev]).view(np.uint64), np.array([w["mean"], w["stddev"]]).view(np.uint64)), "statistics"

# ---- fused chains through the operators (lazy) on sharded rasters -------------------------------------------------------------------
nir_h = synth.host(CellType.UInt16, n, 0x50, kind=synth.INT_RANGE, lo=0, hi=40000)
red_h = synth.host(CellType.UInt16, n, 0x58, kind=synth.INT_RANGE, lo=0, hi=40000)
nir, red = CellBuffer.from_vec(nir_h), CellBuffer.from_vec(red_h)
k0 = ec.lib().ec_kernel_launches()
with ec.lazy():
    ndvi = (nir - red) / (nir + red)
    got = ndvi.to_vec()
assert ec.lib().ec_kernel_launches() == k0 + G, "one fused kernel per strip"
wn = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, nir_h, red_h), orc.tight_binary(orc.ADD, nir_h, red_h))
same(got, wn, "lazy NDVI")
same(nir.normalized_difference(red).to_vec(), wn, "fused NDVI")

# ---- chunked ingest: chunks go to the GPU that owns their strip; result identical to the one-shot upload -------------------------------
from erased_cells_b200 import raster_io
ing = raster_io.ingest(i1_h, nd, masked=True, chunk_cells=4096)
assert ing.buffer().shard_count() == G and ec.lib().ec_mask_shard_count(ing.mask()._h) == G
same(ing.buffer().to_vec(), i1_h, "sharded ingest data")
assert np.array_equal(ing.mask().to_vec(), wm1) and ing.counts() == orc.mask_counts(wm1)
assert raster_io.ingest(a_h, chunk_cells=1000) == a

# ---- strips that come out empty (fewer than 128 cells per strip): every strip keeps the raster's cell type -------------------------
ec.set_shard_min_cells(1)
t_h = cells(CellType.Float32, 100, 9)
t = CellBuffer.from_vec(t_h)
assert t.shard_count() == G and [s[3] for s in t.shards()] == [0] * (G - 1) + [100]
u = t * 2.0
assert u.cell_type() == CellType.Float64
same(u.to_vec(), orc.tight_scalar(orc.MUL, t_h, orc.value(orc.Float64, 2.0)), "tiny sharded raster")
for mode in finish_modes:
    ec.set_shard_finish(mode)
    mn, mx = u.min_max()
    omn, omx = orc.tight_min_max(u.to_vec())
    assert (mn.bits, mx.bits) == (omn.bits, omx.bits), ("empty strips", mode)
ec.set_shard_finish(ec.FINISH_HOST)
st = t.statistics()
w = orc.statistics(t_h, None)
assert (st.count, st.min.bits, st.max.bits) == (w["count"], w["min"].bits, w["max"].bits)
tm = MaskedCellBuffer.from_vec_with_nodata(t_h, NoData.new(CellType.Float32, float(t_h[7])))
assert tm.counts() == orc.mask_counts(orc.mask_from_nodata(t_h, orc.ND_VALUE, orc.value(orc.Float32, float(t_h[7]))))
ec.set_shard_min_cells(THR)

ec._lib.check(ec.lib().ec_synchronize())
print(f"SHARDED_OK devices={devs} finish_modes={finish_modes} launches={ec.lib().ec_kernel_launches()}")
