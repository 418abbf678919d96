"""ctypes binding of liberased_cells_b200.so (the C ABI declared in include/erased_cells_b200.h).

The shared library is the product; this module only declares its signatures and turns ec_status
codes into Python exceptions. There is no Python/numpy compute path: if the library is missing
the import fails, and if no CUDA device is usable every device call raises NoDeviceError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liberased_cells_b200.so")
CSRC = os.path.join(HERE, "csrc")

(EC_OK, EC_NARROWING, EC_OOB, EC_LEN_MISMATCH, EC_INVALID_ARG, EC_CUDA, EC_NCCL, EC_OOM, EC_NO_DEVICE,
 EC_PARSE) = range(10)


class Value(C.Structure):
    """ec_value: 16-byte tagged scalar (CellValue, reference src/value.rs:12-20)."""
    _fields_ = [("ct", C.c_uint8), ("pad", C.c_uint8 * 7), ("bits", C.c_uint64)]


class Statistics(C.Structure):
    """ec_statistics (extension: the reference has no statistics beyond min_max / counts)."""
    _fields_ = [("count", C.c_uint64), ("min", Value), ("max", Value), ("mean", C.c_double), ("stddev", C.c_double)]


MOMENT_WORDS = 9


class ShardInfo(C.Structure):
    """ec_shard_info: where one row strip of a sharded buffer / mask lives."""
    _fields_ = [("logical_device", C.c_int), ("cuda_device", C.c_int), ("offset", C.c_size_t), ("len", C.c_size_t),
                ("device_ptr", C.c_void_p)]


class DeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("l2_bytes", C.c_size_t), ("total_mem_bytes", C.c_size_t), ("name", C.c_char * 128)]


class EcError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"erased_cells_b200 status {status}: {msg}")
        self.status = status


class NarrowingError(EcError):
    """Error::NarrowingError{src, dst} — reference src/error.rs:14-15."""

    def __init__(self, status, msg, src, dst):
        super().__init__(status, msg)
        self.src, self.dst = src, dst


class NoDeviceError(EcError):
    pass


class ParseError(EcError, ValueError):
    pass


def build(jobs: int | None = None, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    jobs = jobs or os.cpu_count() or 4
    r = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building liberased_cells_b200.so failed:\n" + (r.stdout or "")[-4000:] + (r.stderr or "")[-4000:])
    return LIB_PATH


_SIGS = None


def _signatures():
    V, PV = Value, C.POINTER(Value)
    I, U8, SZ, VP, U64, I64 = C.c_int, C.c_uint8, C.c_size_t, C.c_void_p, C.c_uint64, C.c_int64
    PVP, PI, PSZ = C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_size_t)
    S = I  # ec_status
    return {
        "ec_abi_version": (I, []),
        "ec_last_error": (C.c_char_p, []),
        "ec_last_narrowing": (None, [C.POINTER(U8), C.POINTER(U8)]),
        "ec_init": (S, [I]),
        "ec_init_devices": (S, [PI, I]),
        "ec_device_count": (I, []),
        "ec_set_shard_min_cells": (SZ, [SZ]),
        "ec_set_shard_finish": (I, [I]),
        "ec_set_min_max_cache": (I, [I]),
        "ec_set_reduce_trace": (I, [I]),
        "ec_reduce_trace_get": (S, [C.POINTER(U64)]),
        "ec_buf_shard_count": (I, [VP]),
        "ec_mask_shard_count": (I, [VP]),
        "ec_buf_shard": (S, [VP, I, C.POINTER(ShardInfo), PVP]),
        "ec_mask_shard": (S, [VP, I, C.POINTER(ShardInfo), PVP]),
        "ec_device_info_get": (S, [C.POINTER(DeviceInfo)]),
        "ec_set_stream": (S, [VP]),
        "ec_get_stream": (VP, []),
        "ec_synchronize": (S, []),
        "ec_trim": (S, []),
        "ec_guard_violations": (U64, []),
        "ec_cached_bytes": (SZ, []),
        "ec_set_lazy": (S, [I]),
        "ec_set_launch_overlap": (I, [I]),
        "ec_get_lazy": (I, []),
        "ec_jit_cached_kernels": (SZ, []),
        "ec_jit_builds": (SZ, []),
        "ec_jit_dry_build": (S, [C.POINTER(U8), I, I, C.c_char_p, C.c_char_p, C.c_size_t]),
        "ec_kernel_launches": (U64, []),
        "ec_last_kernel": (C.c_char_p, []),
        "ec_event_create": (S, [PVP]),
        "ec_event_record": (S, [VP]),
        "ec_event_elapsed_ms": (S, [VP, VP, C.POINTER(C.c_float)]),
        "ec_event_destroy": (None, [VP]),
        "ec_host_alloc": (S, [SZ, PVP]),
        "ec_host_free": (None, [VP]),
        "ec_host_register": (S, [VP, SZ]),
        "ec_host_unregister": (S, [VP]),
        "ec_ctype_union": (U8, [U8, U8]),
        "ec_ctype_can_fit_into": (I, [U8, U8]),
        "ec_ctype_size_of": (SZ, [U8]),
        "ec_ctype_is_integral": (I, [U8]),
        "ec_ctype_is_signed": (I, [U8]),
        "ec_ctype_name": (C.c_char_p, [U8]),
        "ec_ctype_from_name": (S, [C.c_char_p, C.POINTER(U8)]),
        "ec_ctype_min_value": (S, [U8, PV]),
        "ec_ctype_max_value": (S, [U8, PV]),
        "ec_ctype_zero": (S, [U8, PV]),
        "ec_ctype_one": (S, [U8, PV]),
        "ec_value_convert": (S, [PV, U8, PV]),
        "ec_value_binary": (S, [I, PV, PV, PV]),
        "ec_value_neg": (S, [PV, PV]),
        "ec_value_cmp": (S, [PV, PV, PI]),
        "ec_value_to_f64": (S, [PV, C.POINTER(C.c_double), PI]),
        "ec_value_to_i64": (S, [PV, C.POINTER(I64), PI]),
        "ec_value_to_u64": (S, [PV, C.POINTER(U64), PI]),
        "ec_value_to_prim": (S, [PV, U8, PV, PI]),
        "ec_buf_from_host": (S, [U8, VP, SZ, PVP]),
        "ec_buf_from_host_async": (S, [U8, VP, SZ, PVP]),
        "ec_buf_wait": (S, [VP]),
        "ec_ingest_begin": (S, [U8, SZ, I, PV, I, SZ, PVP]),
        "ec_ingest_next_buffer": (S, [VP, PVP, PSZ]),
        "ec_ingest_submit": (S, [VP, SZ]),
        "ec_ingest_finish": (S, [VP, PVP, PVP]),
        "ec_ingest_abort": (None, [VP]),
        "ec_set_host_copy_threads": (I, [I]),
        "ec_buf_with_defaults": (S, [SZ, U8, PVP]),
        "ec_buf_fill": (S, [SZ, PV, PVP]),
        "ec_buf_wrap_device": (S, [U8, VP, SZ, PVP]),
        "ec_buf_view": (S, [VP, SZ, SZ, PVP]),
        "ec_buf_clone": (S, [VP, PVP]),
        "ec_buf_free": (None, [VP]),
        "ec_buf_len": (SZ, [VP]),
        "ec_buf_ctype": (U8, [VP]),
        "ec_buf_device_ptr": (VP, [VP]),
        "ec_buf_to_host": (S, [VP, VP, SZ]),
        "ec_buf_get": (S, [VP, SZ, PV]),
        "ec_buf_put": (S, [VP, SZ, PV]),
        "ec_buf_extend_host": (S, [VP, U8, VP, SZ]),
        "ec_buf_binary": (S, [I, VP, VP, PVP]),
        "ec_buf_scalar": (S, [I, VP, PV, PVP]),
        "ec_buf_neg": (S, [VP, PVP]),
        "ec_buf_convert": (S, [VP, U8, PVP]),
        "ec_buf_min_max": (S, [VP, VP, PV, PV]),
        "ec_buf_cmp": (S, [VP, VP, PI]),
        "ec_buf_statistics": (S, [VP, VP, C.POINTER(Statistics)]),
        "ec_statistics_plan": (S, [PV, PV, PI, C.POINTER(C.c_double), PI]),
        "ec_buf_moments": (S, [VP, VP, C.c_double, I, C.POINTER(U64)]),
        "ec_statistics_finish": (S, [C.POINTER(U64), C.c_size_t, PV, PV, C.POINTER(Statistics)]),
        "ec_buf_normalized_difference": (S, [VP, VP, PVP]),
        "ec_buf_binary_scalar": (S, [I, VP, VP, I, PV, PVP]),
        "ec_mask_from_bools": (S, [VP, SZ, PVP]),
        "ec_mask_fill": (S, [SZ, I, PVP]),
        "ec_mask_to_bools": (S, [VP, VP, SZ]),
        "ec_mask_clone": (S, [VP, PVP]),
        "ec_mask_slice": (S, [VP, SZ, SZ, PVP]),
        "ec_mask_free": (None, [VP]),
        "ec_mask_len": (SZ, [VP]),
        "ec_mask_device_words": (VP, [VP]),
        "ec_mask_get": (S, [VP, SZ, PI]),
        "ec_mask_put": (S, [VP, SZ, I]),
        "ec_mask_extend_host": (S, [VP, VP, SZ]),
        "ec_mask_not": (S, [VP, PVP]),
        "ec_mask_and": (S, [VP, VP, PVP]),
        "ec_mask_or": (S, [VP, VP, PVP]),
        "ec_mask_counts": (S, [VP, PSZ, PSZ]),
        "ec_mask_all": (S, [VP, I, PI]),
        "ec_mask_cmp": (S, [VP, VP, PI]),
        "ec_nodata_value": (S, [I, U8, PV, PV, PI]),
        "ec_mask_from_nodata": (S, [VP, I, PV, PVP]),
        "ec_buf_fill_nodata": (S, [VP, VP, U8, I, PV, PVP]),
        "ec_masked_binary": (S, [I, VP, VP, VP, VP, PVP, PVP]),
        "ec_row_strip": (S, [SZ, SZ, I, I, PSZ, PSZ]),
        "ec_buf_min_max_keys": (S, [VP, VP, VP]),
        "ec_min_max_from_keys": (S, [U8, C.POINTER(I64), PV, PV]),
        "ec_min_max_to_keys": (S, [PV, PV, C.POINTER(I64)]),
        "ec_comm_unique_id": (S, [VP]),
        "ec_comm_init_rank": (S, [VP, I, I, PVP]),
        "ec_comm_destroy": (None, [VP]),
        "ec_comm_peer_exchange": (I, [VP]),
        "ec_comm_allreduce_min_i64": (S, [VP, VP, SZ]),
        "ec_comm_allreduce_sum_u64": (S, [VP, VP, SZ]),
        "ec_buf_min_max_sharded": (S, [VP, VP, VP, PV, PV]),
        "ec_mask_counts_sharded": (S, [VP, VP, PSZ, PSZ]),
        "ec_buf_statistics_sharded": (S, [VP, VP, VP, C.POINTER(Statistics)]),
        "ec_buf_synth": (S, [U8, SZ, U64, U64, I, C.c_double, C.c_double, U64, PV, PVP]),
    }


_lib = None


def lib():
    """The loaded C-ABI library. Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(erased_cells_b200 has no CPU or pure-Python path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _signatures().items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int):
    if status == EC_OK:
        return
    L = lib()
    msg = (L.ec_last_error() or b"").decode(errors="replace")
    if status == EC_NARROWING:
        s, d = C.c_uint8(), C.c_uint8()
        L.ec_last_narrowing(C.byref(s), C.byref(d))
        raise NarrowingError(status, msg, s.value, d.value)
    if status == EC_NO_DEVICE:
        raise NoDeviceError(status, msg)
    if status == EC_OOB:
        raise IndexError(msg)
    if status == EC_LEN_MISMATCH:
        raise AssertionError(msg)  # assert_eq! in MaskedCellBuffer::new (src/masked/masked_buffer.rs:49-53)
    if status == EC_PARSE:
        raise ParseError(status, msg)
    if status == EC_OOM:
        raise MemoryError(msg)
    raise EcError(status, msg)
