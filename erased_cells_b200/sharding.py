"""Row-strip sharding of one raster across the GPUs of a box (one process per GPU).

Element-wise ops are shard-local: a rank simply holds the CellBuffer / MaskedCellBuffer of its strip.
Only reductions cross GPUs, and only as 16 bytes: the shard kernel leaves {skey(min), ~skey(max)} in
device memory and ONE all-reduce(MIN) over int64 finishes both (order-preserving integer keys make the
result independent of the shard count). The collective is torch.distributed (NCCL over NVLink on the
GPU box; gloo in the CPU tests); the C ABI also carries its own NCCL communicator (ec_comm_*) for hosts
without torch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Value, check, lib
from .api import CellBuffer, CellType, CellValue, Mask, Statistics

ST_REGULAR, ST_EMPTY, ST_NONFINITE = range(3)


def row_strip(width: int, height: int, n_shards: int, shard: int) -> tuple[int, int]:
    """(cell_offset, cell_len) of strip `shard`: whole rows, remainder to the last strip, starts on 128-cell boundaries."""
    off, ln = C.c_size_t(), C.c_size_t()
    check(lib().ec_row_strip(width, height, n_shards, shard, C.byref(off), C.byref(ln)))
    return off.value, ln.value


def keys_of(mn: CellValue, mx: CellValue) -> np.ndarray:
    k = np.zeros(2, dtype=np.int64)
    check(lib().ec_min_max_to_keys(C.byref(mn._v), C.byref(mx._v), k.ctypes.data_as(C.POINTER(C.c_int64))))
    return k


def values_of(ct: CellType, keys) -> tuple[CellValue, CellValue]:
    k = np.ascontiguousarray(np.asarray(keys, dtype=np.int64))
    mn, mx = Value(), Value()
    check(lib().ec_min_max_from_keys(int(ct), k.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(mn), C.byref(mx)))
    return CellValue._wrap(mn), CellValue._wrap(mx)


def finish_min_max(ct: CellType, keys, group=None) -> tuple[CellValue, CellValue]:
    """all-reduce(MIN) of the packed keys tensor (int64[2], on the device the backend needs) + decode."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return values_of(ct, keys.cpu().numpy())


def min_max_sharded(shard: CellBuffer, mask: Mask | None = None, group=None) -> tuple[CellValue, CellValue]:
    """min_max of the whole raster from this rank's strip (every rank gets the result)."""
    import torch
    keys = torch.empty(2, dtype=torch.int64, device="cuda")
    check(lib().ec_buf_min_max_keys(shard._h, mask._h if mask is not None else None, C.c_void_p(keys.data_ptr())))
    check(lib().ec_synchronize())
    return finish_min_max(shard.cell_type(), keys, group)


def counts_sharded(mask: Mask, group=None) -> tuple[int, int]:
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(mask.counts()), dtype=torch.int64, device="cuda" if torch.cuda.is_available() else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0]), int(t[1])


# ---- statistics of a sharded raster (extension; definition in DESIGN.md §4.6) ------------------------------------
# global min/max (one all-reduce) -> the same (pivot, exponent) on every rank -> per-strip exact integer moment sums
# (one device pass) -> all-gather of 72 bytes per rank -> the same host finish everywhere. Exact sums make the result
# independent of the number of strips.
def statistics_plan(mn: CellValue, mx: CellValue) -> tuple[int, float, int]:
    kind, pivot, exp2 = C.c_int(), C.c_double(), C.c_int()
    check(lib().ec_statistics_plan(C.byref(mn._v), C.byref(mx._v), C.byref(kind), C.byref(pivot), C.byref(exp2)))
    return kind.value, pivot.value, exp2.value


def moments(shard: CellBuffer, mask: Mask | None, pivot: float, exp2: int) -> np.ndarray:
    """this strip's raw accumulators: uint64[9] = {count, four 128-bit two's complement sums}"""
    raw = np.zeros(_lib.MOMENT_WORDS, dtype=np.uint64)
    check(lib().ec_buf_moments(shard._h, mask._h if mask is not None else None, pivot, exp2,
                               raw.ctypes.data_as(C.POINTER(C.c_uint64))))
    return raw


def finish_statistics(raws, mn: CellValue, mx: CellValue) -> Statistics:
    raws = np.ascontiguousarray(np.asarray(raws, dtype=np.uint64).reshape(-1, _lib.MOMENT_WORDS))
    out = _lib.Statistics()
    check(lib().ec_statistics_finish(raws.ctypes.data_as(C.POINTER(C.c_uint64)), raws.shape[0], C.byref(mn._v),
                                     C.byref(mx._v), C.byref(out)))
    return Statistics(out)


def gather_statistics(raw: np.ndarray, mn: CellValue, mx: CellValue, group=None) -> Statistics:
    """all-gather every rank's raw accumulators and finish (every rank gets the same result)"""
    import torch
    import torch.distributed as dist
    parts = raw.reshape(1, -1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.from_numpy(raw.view(np.int64).copy())
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        got = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(got, t, group=group)
        parts = np.stack([g.cpu().numpy().view(np.uint64) for g in got])
    return finish_statistics(parts, mn, mx)


def statistics_sharded(shard: CellBuffer, mask: Mask | None = None, group=None, comm: "Comm | None" = None) -> Statistics:
    """statistics of the whole raster from this rank's strip; `comm` finishes min/max inside the reduction kernel"""
    mn, mx = comm.min_max(shard, mask) if comm is not None else min_max_sharded(shard, mask, group)
    kind, pivot, exp2 = statistics_plan(mn, mx)
    if kind == ST_REGULAR:
        raw = moments(shard, mask, pivot, exp2)
    else:
        raw = np.zeros(_lib.MOMENT_WORDS, dtype=np.uint64)
        raw[0] = mask.counts()[0] if mask is not None else shard.len()
    return gather_statistics(raw, mn, mx, group)


class Comm:
    """The library's own communicator (ec_comm_*): NCCL for plumbing, and — when every rank can map every other rank's
    mailbox over NVLink — sharded reductions as ONE kernel per GPU (reduction + peer exchange + final fold)."""

    def __init__(self, handle, n_ranks, rank):
        self._h, self.n_ranks, self.rank = handle, n_ranks, rank

    @staticmethod
    def create(group=None) -> "Comm":
        """Rank 0 draws the NCCL unique id; it travels through torch.distributed (any backend)."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        raw = (C.c_uint8 * 128)()
        if rank == 0:
            check(lib().ec_comm_unique_id(raw))
        t = torch.tensor(list(raw), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0, group=group)
        raw = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        check(lib().ec_comm_init_rank(raw, world, rank, C.byref(h)))
        return Comm(h.value, world, rank)

    @property
    def peer_exchange(self) -> bool:
        return bool(lib().ec_comm_peer_exchange(self._h))

    def min_max(self, shard: CellBuffer, mask: Mask | None = None) -> tuple[CellValue, CellValue]:
        mn, mx = Value(), Value()
        check(lib().ec_buf_min_max_sharded(self._h, shard._h, mask._h if mask is not None else None, C.byref(mn), C.byref(mx)))
        return CellValue._wrap(mn), CellValue._wrap(mx)

    def counts(self, mask: Mask) -> tuple[int, int]:
        d, n = C.c_size_t(), C.c_size_t()
        check(lib().ec_mask_counts_sharded(self._h, mask._h, C.byref(d), C.byref(n)))
        return d.value, n.value

    def statistics(self, shard: CellBuffer, mask: Mask | None = None) -> Statistics:
        """statistics of the whole raster (extension): fused min/max, this strip's exact sums, one all-reduce, host finish"""
        out = _lib.Statistics()
        check(lib().ec_buf_statistics_sharded(self._h, shard._h, mask._h if mask is not None else None, C.byref(out)))
        return Statistics(out)

    def close(self):
        if self._h:
            lib().ec_comm_destroy(self._h)
            self._h = None


_NEG_OUT = {CellType.UInt8: CellType.Int16, CellType.UInt16: CellType.Int32, CellType.UInt32: CellType.Float64, CellType.UInt64: CellType.Float64}


def _typed(strip: CellBuffer, ct: CellType) -> CellBuffer:
    """An op on an EMPTY strip returns UInt8([]) (FromIterator of nothing, src/buffer.rs:233-236) — right for an empty raster,
    wrong for the empty strip of a non-empty one (fewer rows than ranks): the other ranks hold Float64, and a reduction would
    mix UInt8 seed keys with Float64 keys. The strip keeps the raster's logical cell type instead."""
    if strip.len() == 0 and strip.cell_type() != ct:
        return CellBuffer.with_defaults(0, ct)
    return strip


class ShardedCellBuffer:
    """One raster, row-strip sharded over the ranks of a Comm: this rank holds `strip` (rows
    [row0, row0 + rows)). Element-wise ops, neg and convert are shard-local and return another ShardedCellBuffer;
    min_max finishes across the GPUs (one fused kernel per GPU when the Comm has the peer exchange)."""

    def __init__(self, strip, width: int, height: int, comm: Comm):
        self.strip, self.width, self.height, self.comm = strip, width, height, comm
        self.offset, local = row_strip(width, height, comm.n_ranks, comm.rank)
        assert strip.len() == local, "strip length does not match this rank's row strip"

    @staticmethod
    def from_host(raster: np.ndarray, comm: Comm) -> "ShardedCellBuffer":
        """Every rank passes the same (height, width) host raster (or a memmap of it) and uploads only its strip."""
        h, w = raster.shape
        off, ln = row_strip(w, h, comm.n_ranks, comm.rank)
        return ShardedCellBuffer(CellBuffer.from_vec(raster.reshape(-1)[off:off + ln]), w, h, comm)

    def _like(self, strip, ct: CellType):
        return type(self)(_typed(strip, ct) if self.len() else strip, self.width, self.height, self.comm)

    def len(self) -> int:
        return self.width * self.height

    def cell_type(self) -> CellType:
        return self.strip.cell_type()

    def _bin(self, op, rhs):
        return self._like(self.strip._bin(op, rhs.strip if isinstance(rhs, ShardedCellBuffer) else rhs), CellType.Float64)

    def __add__(self, r): return self._bin(0, r)
    def __sub__(self, r): return self._bin(1, r)
    def __mul__(self, r): return self._bin(2, r)
    def __truediv__(self, r): return self._bin(3, r)
    def __neg__(self): return self._like(-self.strip, _NEG_OUT.get(self.cell_type(), self.cell_type()))

    def convert(self, ct: CellType):
        return self._like(self.strip.convert(ct), ct)

    def min_max(self):
        return self.comm.min_max(self.strip)

    def statistics(self) -> Statistics:
        return self.comm.statistics(self.strip)

    def gather(self) -> np.ndarray:
        """The whole raster on every rank's host (tests / small rasters)."""
        import torch.distributed as dist
        parts = [None] * self.comm.n_ranks
        dist.all_gather_object(parts, self.strip.to_vec())
        return np.concatenate(parts).reshape(self.height, self.width)


class ShardedMaskedCellBuffer:
    """Row-strip sharded MaskedCellBuffer: ops propagate the strip's mask locally; min_max / counts cross GPUs."""

    def __init__(self, strip, width: int, height: int, comm: Comm):
        self.strip, self.width, self.height, self.comm = strip, width, height, comm

    @staticmethod
    def from_host_with_nodata(raster: np.ndarray, nodata, comm: Comm) -> "ShardedMaskedCellBuffer":
        from .api import MaskedCellBuffer
        h, w = raster.shape
        off, ln = row_strip(w, h, comm.n_ranks, comm.rank)
        buf = CellBuffer.from_vec(raster.reshape(-1)[off:off + ln], wait=False)
        return ShardedMaskedCellBuffer(MaskedCellBuffer.from_buffer_with_nodata(buf, nodata), w, h, comm)

    def _bin(self, op, rhs):
        from .api import MaskedCellBuffer
        r = self.strip._bin(op, rhs.strip if isinstance(rhs, ShardedMaskedCellBuffer) else rhs)
        if self.width * self.height:
            r = MaskedCellBuffer(_typed(r.buffer(), CellType.Float64), r.mask())
        return ShardedMaskedCellBuffer(r, self.width, self.height, self.comm)

    def __add__(self, r): return self._bin(0, r)
    def __sub__(self, r): return self._bin(1, r)
    def __mul__(self, r): return self._bin(2, r)
    def __truediv__(self, r): return self._bin(3, r)

    def min_max(self):
        return self.comm.min_max(self.strip.buffer(), self.strip.mask())

    def counts(self):
        return self.comm.counts(self.strip.mask())

    def statistics(self) -> Statistics:
        return self.comm.statistics(self.strip.buffer(), self.strip.mask())
