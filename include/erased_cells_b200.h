/*
 * erased_cells_b200 — C ABI of the B200-native per-cell compute path of s22s/erased-cells.
 *
 * This header is the drop-in boundary. The reference crate (v0.1.1) has no FFI of its own: its
 * boundary is the public Rust API — trait `BufferOps` (src/lib.rs:104-163), the `std::ops` impls on
 * `CellBuffer` (src/buffer.rs:321-371) and `MaskedCellBuffer` (src/masked/masked_buffer.rs:323-383),
 * `Mask` (src/masked/mask.rs:14-164), `NoData` (src/masked/nodata.rs:9-68), `CellType`
 * (src/ctype.rs) and `CellValue` (src/value.rs). Each entry point below names the reference item
 * whose body it replaces; INTEGRATION.md shows the Rust `extern "C"` block and the `impl` bodies a
 * maintainer would write over it, include/erased_cells.hpp is the same mirror in C++ and
 * erased_cells_b200/ (Python, ctypes) is the one the parity tests drive.
 *
 * Conventions
 *   - Plain C types only. Buffers and masks are opaque handles to device-resident storage, owned
 *     by exactly one caller-side value; every op returns a fresh handle and never mutates or
 *     aliases its inputs (Rust ownership: `Drop` -> ec_*_free, `Clone` -> ec_*_clone).
 *   - Every function returns ec_status (0 = ok) unless noted; ec_last_error() gives the
 *     thread-local message of the last failure. No exceptions cross this boundary.
 *   - One process drives one GPU (ec_init(device)); work is enqueued on the library's current
 *     stream for the calling thread (ec_set_stream to share the caller's stream). Calls that
 *     return host-visible results (min_max, counts, get, to_host, cmp) synchronise that stream.
 *   - There is no CPU compute path: without a usable CUDA device every device entry point fails
 *     with EC_NO_DEVICE.
 *   - Cell types are the `CellType` discriminants (src/lib.rs:85-101): 0..9 =
 *     UInt8, UInt16, UInt32, UInt64, Int8, Int16, Int32, Int64, Float32, Float64.
 *   - Arithmetic semantics (src/value.rs:196-209): out[i] = (f64)l[i] OP (f64)r[i], IEEE-754
 *     binary64, round to nearest even, no FMA contraction; the result buffer is always Float64
 *     (UInt8 when empty, src/buffer.rs:234). NaN results carry the x86-64 SSE2 sign/payload the
 *     reference produces on its platform (lhs NaN wins, else rhs NaN, quieted; invalid operation
 *     -> 0xFFF8000000000000).
 */
#ifndef ERASED_CELLS_B200_H
#define ERASED_CELLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EC_ABI_VERSION 2

/* Error conventions — src/error.rs:12-27. Only NarrowingError is reachable from the hot path
 * (src/buffer.rs:155-159, src/value.rs:77-81); index panics (src/lib.rs:136-147) and the
 * MaskedCellBuffer::new length assertion (src/masked/masked_buffer.rs:49-53) surface as EC_OOB /
 * EC_LEN_MISMATCH so the host shim can panic/raise. */
typedef enum ec_status {
    EC_OK = 0,
    EC_NARROWING = 1,     /* Error::NarrowingError{src,dst}; see ec_last_narrowing() */
    EC_OOB = 2,           /* index >= len */
    EC_LEN_MISMATCH = 3,  /* buffer/mask length mismatch */
    EC_INVALID_ARG = 4,   /* bad cell type / op / null handle */
    EC_CUDA = 5,          /* CUDA runtime error, message in ec_last_error() */
    EC_NCCL = 6,          /* NCCL error or NCCL not loadable */
    EC_OOM = 7,           /* device or pinned-host allocation failed */
    EC_NO_DEVICE = 8,     /* no usable CUDA device: there is no CPU fallback */
    EC_PARSE = 9          /* Error::ParseError (CellType::from_str) */
} ec_status;

typedef enum ec_ctype {
    EC_UINT8 = 0, EC_UINT16 = 1, EC_UINT32 = 2, EC_UINT64 = 3,
    EC_INT8 = 4, EC_INT16 = 5, EC_INT32 = 6, EC_INT64 = 7,
    EC_FLOAT32 = 8, EC_FLOAT64 = 9
} ec_ctype;

typedef enum ec_op { EC_ADD = 0, EC_SUB = 1, EC_MUL = 2, EC_DIV = 3 } ec_op;

/* NoData<T> — src/masked/nodata.rs:9-17 */
typedef enum ec_nodata_kind { EC_NODATA_NONE = 0, EC_NODATA_DEFAULT = 1, EC_NODATA_VALUE = 2 } ec_nodata_kind;

/* CellValue — src/value.rs:12-20: a 16-byte tagged scalar. `bits` holds the little-endian payload
 * of the tagged primitive in its low size_of(ct) bytes, upper bytes zero. */
typedef struct ec_value {
    uint8_t ct;
    uint8_t pad[7];
    uint64_t bits;
} ec_value;

/* Statistics of the valid cells — an EXTENSION, the reference has none beyond min_max and Mask::counts (its only
 * mention is a comment quoting gdal_calc.py, src/gdal/rasterband.rs:152-156). mean / stddev (population, as GDAL's
 * STATISTICS_STDDEV) follow the order-independent definition of DESIGN.md §4.6: bit-identical for any grid and any
 * number of GPUs; within 2 ULP of the exactly rounded value on the data of tests/test_oracle_self.py. */
typedef struct ec_statistics {
    uint64_t count;      /* valid cells */
    ec_value min, max;   /* as ec_buf_min_max: total order, (T::MAX, T::MIN) when count == 0 */
    double mean, stddev; /* canonical NaN when count == 0 or a valid cell is NaN; mean = +-inf / stddev = NaN with infinities */
} ec_statistics;
enum { EC_STATS_REGULAR = 0, EC_STATS_EMPTY = 1, EC_STATS_NONFINITE = 2 };
#define EC_MOMENT_WORDS 9 /* {count, X1.lo, X1.hi, X2.lo, X2.hi, Z1.lo, Z1.hi, Z2.lo, Z2.hi}: 128-bit two's complement sums */

typedef struct ec_buf ec_buf;   /* CellBuffer: typed cells in HBM */
typedef struct ec_mask ec_mask; /* Mask: validity bits in HBM, packed 32 cells per little-endian word */
typedef struct ec_event ec_event;
typedef struct ec_ingest ec_ingest; /* a raster arriving chunk by chunk (ec_ingest_*) */
typedef struct ec_shard_info ec_shard_info;
typedef struct ec_comm ec_comm;

typedef struct ec_device_info {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t l2_bytes;
    size_t total_mem_bytes;
    char name[128];
} ec_device_info;

/* ---- library / device context -------------------------------------------------------------- */
int ec_abi_version(void);
const char* ec_last_error(void);
/* src/error.rs:14: the {src, dst} of the last EC_NARROWING on this thread */
void ec_last_narrowing(uint8_t* src, uint8_t* dst);
ec_status ec_init(int device);           /* bind this process to one GPU; idempotent */
/* One process, several GPUs (the reference's caller is one process and one CellBuffer is the whole raster,
 * src/buffer.rs:52, src/lib.rs:104-163): bind the library to `n` (1..16) CUDA devices. From then on a buffer or mask of
 * at least ec_set_shard_min_cells() cells is kept as n row strips, strip g on devices[g], and every entry point below
 * works on it unchanged: maps, casts and mask logic strip by strip with no data-path collective, reductions finishing
 * across the GPUs. Without a call, the first use binds $EC_DEVICES ("0,1,2,3"), else $EC_DEVICE, else $LOCAL_RANK, else 0.
 * The same CUDA device may be listed more than once (that is how the sharded paths are tested on a one-GPU box). */
ec_status ec_init_devices(const int* devices, int n);
int ec_device_count(void);               /* devices bound so far (0 before the first use) */
/* threshold (cells) from which new buffers / masks are sharded; default 2^24, $EC_SHARD_MIN_CELLS. Returns the previous value. */
size_t ec_set_shard_min_cells(size_t cells);
/* How a sharded reduction crosses the GPUs: 0 the host folds the strips' 16-byte partials out of mapped pinned memory
 * (default); 1 the finishing CTAs exchange them GPU to GPU through peer-mapped mailboxes, one kernel per GPU and nothing
 * else; 2 ncclAllReduce over ncclCommInitAll communicators. Identical bits either way (integer keys, exact integer sums).
 * $EC_SHARD_FINISH = host | peer | nccl. Returns the previous mode, -1 for a bad one. */
int ec_set_shard_finish(int mode);
ec_status ec_device_info_get(ec_device_info* out);
ec_status ec_set_stream(void* cuda_stream); /* NULL restores the library's own stream */
void* ec_get_stream(void);
ec_status ec_synchronize(void);
/* Freed device blocks are cached per stream for reuse (outputs are fresh allocations on every op).
 * ec_trim() synchronises the current stream and returns all cached blocks to the driver. */
ec_status ec_trim(void);
/* With EC_DEBUG_GUARD=1 in the environment every device block carries 256-byte red zones checked on free;
 * this is the number of zones found overwritten so far (0 when the mode is off). */
uint64_t ec_guard_violations(void);
size_t ec_cached_bytes(void);
uint64_t ec_kernel_launches(void);       /* number of this library's kernels launched so far */
/* Where the time of one reduction call goes (profiling aid, off by default). With tracing on, ec_buf_min_max and the
 * sharded reductions stamp %globaltimer inside the kernel; ec_reduce_trace_get returns, for the calling thread's last
 * call, in ns: [0] call entered, [1] launch returned, [2] result seen by the host (CLOCK_REALTIME); [3] first CTA
 * running, [4] last CTA of the grid done, [5] this GPU's result folded, [6] partial sent to the peers, [7] peers'
 * partials received, [8] result published (%globaltimer; 0 = stage not reached); [9] offset to add to a %globaltimer
 * value to compare it with CLOCK_REALTIME (calibrated on the spot, good to about a microsecond). */
int ec_set_reduce_trace(int on);
ec_status ec_reduce_trace_get(uint64_t out12[12]);
/* name of the last kernel family launched by this thread (for profiles/bench bookkeeping) */
const char* ec_last_kernel(void);
ec_status ec_event_create(ec_event** out);
ec_status ec_event_record(ec_event* e);  /* on the current stream */
ec_status ec_event_elapsed_ms(ec_event* start, ec_event* stop, float* ms); /* synchronises `stop` */
void ec_event_destroy(ec_event* e);
/* pinned host staging for from_vec / to_vec (src/buffer.rs:64-66, :175-185) */
ec_status ec_host_alloc(size_t bytes, void** out);
void ec_host_free(void* p);
ec_status ec_host_register(void* p, size_t bytes);
ec_status ec_host_unregister(void* p);

/* ---- CellType: host-side lattice — src/ctype.rs -------------------------------------------- */
uint8_t ec_ctype_union(uint8_t a, uint8_t b);        /* src/ctype.rs:99-126 */
int ec_ctype_can_fit_into(uint8_t a, uint8_t b);     /* src/ctype.rs:129-131 */
size_t ec_ctype_size_of(uint8_t ct);                 /* src/ctype.rs:87-96 */
int ec_ctype_is_integral(uint8_t ct);                /* src/ctype.rs:55-68 */
int ec_ctype_is_signed(uint8_t ct);                  /* src/ctype.rs:71-84 */
const char* ec_ctype_name(uint8_t ct);               /* Display, src/ctype.rs:23-27 */
ec_status ec_ctype_from_name(const char* s, uint8_t* out); /* FromStr, src/ctype.rs:29-43 */
ec_status ec_ctype_min_value(uint8_t ct, ec_value* out);   /* src/ctype.rs:158-167 */
ec_status ec_ctype_max_value(uint8_t ct, ec_value* out);   /* src/ctype.rs:170-179 */
ec_status ec_ctype_zero(uint8_t ct, ec_value* out);        /* src/ctype.rs:134-143 */
ec_status ec_ctype_one(uint8_t ct, ec_value* out);         /* src/ctype.rs:146-155 */

/* ---- CellValue: host-side scalar — src/value.rs -------------------------------------------- */
ec_status ec_value_convert(const ec_value* v, uint8_t ct, ec_value* out);             /* :74-98 */
ec_status ec_value_binary(int op, const ec_value* l, const ec_value* r, ec_value* out); /* :199-222 */
ec_status ec_value_neg(const ec_value* v, ec_value* out);                             /* :224-240 */
ec_status ec_value_cmp(const ec_value* l, const ec_value* r, int* ordering);          /* :248-265 */
ec_status ec_value_to_f64(const ec_value* v, double* out, int* is_some);              /* :144-156 */
ec_status ec_value_to_i64(const ec_value* v, int64_t* out, int* is_some);             /* :118-129 */
ec_status ec_value_to_u64(const ec_value* v, uint64_t* out, int* is_some);            /* :131-142 */
/* `self.to_<p>()`: the value-checked num-traits default chain (call sites src/value.rs:92, src/buffer.rs:212,
 * src/gdal/mod.rs:59); *is_some = 0 is the reference's `None` */
ec_status ec_value_to_prim(const ec_value* v, uint8_t ct, ec_value* out, int* is_some);

/* ---- CellBuffer — src/buffer.rs ------------------------------------------------------------- */
/* from_vec / From<Vec<T>> / From<&[T]> (:64-66, :252-276): one H2D copy of `len` cells */
ec_status ec_buf_from_host(uint8_t ct, const void* host, size_t len, ec_buf** out);
/* Same, without blocking: the H2D copy runs on the library's upload stream (full-duplex with D2H traffic of
 * the compute stream); every later op on the buffer is ordered after it. `host` must stay valid until
 * ec_buf_wait() returns — from_vec(Vec<T>) owns its Vec, so the Rust shim parks it in the buffer until then. */
ec_status ec_buf_from_host_async(uint8_t ct, const void* host, size_t len, ec_buf** out);
ec_status ec_buf_wait(const ec_buf* b);
/* Chunked ingest — the step in front of the path (RasterBandEx::read_cells / read_cells_masked, src/gdal/rasterband.rs:81-126,
 * feed whole bands to from_vec / from_vec_with_nodata). A reader that produces the band piece by piece asks for a pinned
 * staging buffer, fills it and submits it; the H2D copy of chunk k runs on the upload stream while the reader fills chunk
 * k + 1 and the NoData compare of chunk k - 1 writes packed mask words on the compute stream. `with_mask` = also build
 * the MaskedCellBuffer's mask (kind NONE: all valid). chunk_cells = 0 picks 32 MiB chunks; it is rounded up to a multiple
 * of 128 cells. On a multi-GPU library every chunk goes to the GPU that owns its row strip, over that GPU's own link.
 *   ec_ingest_begin -> { ec_ingest_next_buffer (blocks until that staging buffer is free; capacity 0 = all cells in),
 *                        fill, ec_ingest_submit(n <= capacity; short chunks a multiple of 128 cells unless last) }*
 *                   -> ec_ingest_finish (hands over buffer and mask; nothing is waited for: consumers are stream-ordered) */
ec_status ec_ingest_begin(uint8_t ct, size_t len, int nodata_kind, const ec_value* value_or_null, int with_mask, size_t chunk_cells,
                          ec_ingest** out);
ec_status ec_ingest_next_buffer(ec_ingest* g, void** host_chunk, size_t* capacity_cells);
ec_status ec_ingest_submit(ec_ingest* g, size_t n_cells);
ec_status ec_ingest_finish(ec_ingest* g, ec_buf** out_buf, ec_mask** out_mask_or_null);
void ec_ingest_abort(ec_ingest* g);
/* CellBuffer::from_vec / to_vec on a Vec<T> (pageable memory, src/buffer.rs:60-66, :175-188): copies of >= 16 MiB between
 * memory that is neither pinned nor registered and HBM are moved 8 MiB at a time through pinned staging by `threads` host
 * threads (the caller among them) while the DMA engine moves the previous chunks; with several GPUs the chunks alternate
 * between the strips' links. 0 = leave pageable copies to the CUDA driver (one staging buffer, one thread). Default:
 * $EC_HOST_COPY_THREADS, else (also for threads < 0) min(8, cores - 2), min(12, cores - 2) when the copy spans several GPUs'
 * links or lands in memory whose pages do not exist yet (a fresh Vec: page-fault bound). Pinned or registered memory is always copied directly. Returns the
 * previous setting. */
int ec_set_host_copy_threads(int threads);
/* with_defaults (:68-77) */
ec_status ec_buf_with_defaults(size_t len, uint8_t ct, ec_buf** out);
/* fill (:79-88): type = the value's type */
ec_status ec_buf_fill(size_t len, const ec_value* value, ec_buf** out);
/* wrap caller-owned device memory (row strip of a raster already in HBM); not freed by ec_buf_free */
ec_status ec_buf_wrap_device(uint8_t ct, void* device_ptr, size_t len, ec_buf** out);
/* a row strip (or any 32-byte aligned sub-range) of a buffer as a buffer of its own: shares the allocation
 * (refcounted), no copy; in-place mutation through either handle copies first */
ec_status ec_buf_view(const ec_buf* b, size_t offset_cells, size_t len, ec_buf** out);
ec_status ec_buf_clone(const ec_buf* b, ec_buf** out);   /* #[derive(Clone)] (:50) */
void ec_buf_free(ec_buf* b);                             /* Drop */
size_t ec_buf_len(const ec_buf* b);                      /* :90-99 */
uint8_t ec_buf_ctype(const ec_buf* b);                   /* :105-114 */
void* ec_buf_device_ptr(const ec_buf* b);
/* to_vec's copy-out after convert (:175-185): D2H of len*size_of(ct) bytes, synchronises */
ec_status ec_buf_to_host(const ec_buf* b, void* host, size_t host_bytes);
ec_status ec_buf_get(const ec_buf* b, size_t index, ec_value* out);     /* :125-134, EC_OOB = panic */
ec_status ec_buf_put(ec_buf* b, size_t index, const ec_value* value);   /* :136-148 */
/* Extend<C> (:205-221): append `n` host cells of type `ct`, converted with `to_<p>().unwrap()`
 * semantics (EC_NARROWING where the reference would panic) */
ec_status ec_buf_extend_host(ec_buf* b, uint8_t ct, const void* host, size_t n);
/* cb_bin_op! `&A op &B` (:324-329): zip => min(len), out Float64 (UInt8 if empty) */
ec_status ec_buf_binary(int op, const ec_buf* l, const ec_buf* r, ec_buf** out);
/* cb_bin_op! `A op scalar` (:346-352) */
ec_status ec_buf_scalar(int op, const ec_buf* l, const ec_value* r, ec_buf** out);
/* Neg (:360-371) with the per-type widening of src/value.rs:224-240 (signed MIN wraps) */
ec_status ec_buf_neg(const ec_buf* b, ec_buf** out);
/* convert (:150-167): same type = clone; illegal = EC_NARROWING before any launch */
ec_status ec_buf_convert(const ec_buf* b, uint8_t ct, ec_buf** out);
/* min_max (:169-173); with a mask: MaskedCellBuffer::min_max (src/masked/masked_buffer.rs:208-217).
 * Total order for floats, seeds (T::MAX, T::MIN) participate. */
ec_status ec_buf_min_max(const ec_buf* b, const ec_mask* mask_or_null, ec_value* min_out, ec_value* max_out);
/* Ord/Eq for CellBuffer (:373-436): cell type, then lexicographic total order, then length */
ec_status ec_buf_cmp(const ec_buf* l, const ec_buf* r, int* ordering);

/* ---- statistics (extension, see ec_statistics above) ---------------------------------------------
 * ec_buf_statistics = min_max pass + one moments pass (skipped when empty or non-finite). The three pieces below
 * are what a row-strip sharded caller composes: global min/max -> plan (host) -> per-strip moments (device, exact
 * integer sums) -> finish over all strips' raw words (host). */
ec_status ec_buf_statistics(const ec_buf* b, const ec_mask* mask_or_null, ec_statistics* out);
/* A buffer remembers the result of an unmasked ec_buf_min_max (put / extend forget it): a second min_max is free and
 * ec_buf_statistics of Float32 / 64-bit cells skips its min_max pass. $EC_MIN_MAX_CACHE=0 / ec_set_min_max_cache(0)
 * turn that off (benchmarks that want every call to touch HBM). Returns the previous setting. */
int ec_set_min_max_cache(int on);
ec_status ec_statistics_plan(const ec_value* min, const ec_value* max, int* kind, double* pivot, int* exp2);
ec_status ec_buf_moments(const ec_buf* b, const ec_mask* mask_or_null, double pivot, int exp2, uint64_t raw[EC_MOMENT_WORDS]);
ec_status ec_statistics_finish(const uint64_t* raws, size_t n_parts, const ec_value* min, const ec_value* max,
                               ec_statistics* out);

/* ---- fused op chains behind the same operators -------------------------
 * ec_set_lazy(1) (per thread, default 0): ec_buf_binary / ec_buf_scalar / ec_masked_binary return at once with a
 * pending result over refcounted snapshots of their operands; the first access evaluates it, fusing
 * `(X - Y) / (X + Y)`, `(X op1 Y) op2 scalar` and `(X op1 s1) op2 s2` into one pass over HBM. Results are bit-identical to eager
 * evaluation; later put/extend on an operand do not affect a pending result (copy on write).
 * (mode 2, the interpreted expression VM of ABI 1, is gone: it ran at 0.13 of the HBM roofline; EC_INVALID_ARG.)
 * ec_set_lazy(3) compiles a longer chain (up to 8 operands, 8 scalars, 48 ops) into ONE streaming kernel
 * specialised at run time with NVRTC (ec_jit.cu): same geometry, rounding and NaN rule as the eager kernels, cached
 * by shape (scalars are kernel parameters). Without libnvrtc the chain is evaluated op by op (same bits). */
ec_status ec_set_lazy(int mode);
int ec_get_lazy(void);
/* Launch overlap (process-wide, default on; EC_LAUNCH_OVERLAP=0 turns it off at ec_init): every op of the reference's
 * API is its own kernel launch, so a chain or a sweep pays the drain of one grid plus the ramp of the next per op.
 * With overlap on, the streaming kernels are launched with programmatic stream serialization (sm_90+): a grid's
 * CTAs may be scheduled while the grid before it on the stream drains, and block on `griddepcontrol.wait` — until
 * that grid has completed and its writes are visible — before their first global memory access. Stream order of
 * every read and write is unchanged, results are identical. Returns the previous setting, -1 without a device. */
int ec_set_launch_overlap(int on);
/* run-time specialised kernels built so far in this process */
size_t ec_jit_cached_kernels(void);
/* NVRTC builds done by this process; a kernel found in the on-disk cache ($EC_JIT_CACHE, default
 * ~/.cache/erased_cells_b200/jit, "off" disables) is loaded without one */
size_t ec_jit_builds(void);
/* build (not load, not launch) the kernel for `expr` — straight-line C over v0.. (operands as f64) and c0.. (scalars)
 * made of ecj_add/ecj_sub/ecj_mul/ecj_div calls — for sm_100a; works without a GPU. EC_NO_DEVICE = libnvrtc missing. */
ec_status ec_jit_dry_build(const uint8_t* cell_types, int n_operands, int n_scalars, const char* expr, char* log, size_t log_capacity);
/* `(&a - &b) / (&a + &b)` with the three roundings of the unfused chain; 1 pass over HBM */
ec_status ec_buf_normalized_difference(const ec_buf* a, const ec_buf* b, ec_buf** out);
/* `(l op1 r) op2 s`, e.g. README `buf1 / buf2 * 0.5` */
ec_status ec_buf_binary_scalar(int op1, const ec_buf* l, const ec_buf* r, int op2, const ec_value* s, ec_buf** out);

/* ---- Mask — src/masked/mask.rs -------------------------------------------------------------- */
ec_status ec_mask_from_bools(const uint8_t* host_bools, size_t len, ec_mask** out);  /* Mask::new :16-18 */
ec_status ec_mask_fill(size_t len, int value, ec_mask** out);                        /* :21-23 */
ec_status ec_mask_to_bools(const ec_mask* m, uint8_t* host_bools, size_t capacity);  /* IntoIterator :171-177 */
ec_status ec_mask_clone(const ec_mask* m, ec_mask** out);
/* the validity bits of a row strip (cells [offset, offset + len) of the mask, offset a multiple of 128 cells like every
 * strip start) as a mask of its own — the Mask half of a MaskedCellBuffer strip whose buffer half is ec_buf_view. A copy
 * (1/8 byte per cell): a packed mask keeps the bits past its length zero, which a window into another mask's words
 * could not guarantee. */
ec_status ec_mask_slice(const ec_mask* m, size_t offset_cells, size_t len, ec_mask** out);
void ec_mask_free(ec_mask* m);
size_t ec_mask_len(const ec_mask* m);                                               /* :37-39 */
void* ec_mask_device_words(const ec_mask* m);
ec_status ec_mask_get(const ec_mask* m, size_t index, int* out);                    /* :58-60 */
ec_status ec_mask_put(ec_mask* m, size_t index, int value);                         /* :50-52 */
ec_status ec_mask_extend_host(ec_mask* m, const uint8_t* host_bools, size_t n);     /* Extend :83-87 */
ec_status ec_mask_not(const ec_mask* m, ec_mask** out);                             /* :103-116 */
ec_status ec_mask_and(const ec_mask* l, const ec_mask* r, ec_mask** out);           /* :118-140, min(len) */
ec_status ec_mask_or(const ec_mask* l, const ec_mask* r, ec_mask** out);            /* :142-164 */
ec_status ec_mask_counts(const ec_mask* m, size_t* data, size_t* nodata);           /* :72-80 */
ec_status ec_mask_all(const ec_mask* m, int value, int* out);                       /* :67-69 */
ec_status ec_mask_cmp(const ec_mask* l, const ec_mask* r, int* ordering);           /* derive(Ord) :10 */

/* ---- MaskedCellBuffer / NoData — src/masked/masked_buffer.rs, src/masked/nodata.rs ----------- */
/* NoData::value (:23-40): returns EC_OK and *has_value = 0 for NoData::None */
ec_status ec_nodata_value(int kind, uint8_t ct, const ec_value* value_or_null, ec_value* out, int* has_value);
/* from_vec_with_nodata (:62-71): mask[i] = !(cell[i] == sentinel) under total order (bitwise) */
ec_status ec_mask_from_nodata(const ec_buf* b, int kind, const ec_value* value_or_null, ec_mask** out);
/* to_vec_with_nodata (:137-152): convert to dst_ct (EC_NARROWING if illegal) and replace cells whose
 * mask bit is 0 by the sentinel, one fused pass. kind NONE = plain convert. */
ec_status ec_buf_fill_nodata(const ec_buf* b, const ec_mask* m, uint8_t dst_ct, int kind,
                             const ec_value* value_or_null, ec_buf** out);
/* masked cb_bin_op! (:326-336): data as ec_buf_binary on all cells, mask = lmask & rmask, fused */
ec_status ec_masked_binary(int op, const ec_buf* lbuf, const ec_mask* lmask, const ec_buf* rbuf,
                           const ec_mask* rmask, ec_buf** out_buf, ec_mask** out_mask);

/* ---- row-strip sharding across GPUs (no reference counterpart) -----------------
 * Two ways to use several GPUs. (1) One process: ec_init_devices(); buffers are sharded behind the same handles and
 * nothing else changes (ec_buf_shard* below only look inside). (2) One process per GPU (torchrun): every rank holds
 * its strip as a plain buffer and finishes reductions with ec_comm_* / ec_*_sharded. */
/* the strips of a handle: 0 for a plain buffer / mask */
int ec_buf_shard_count(const ec_buf* b);
int ec_mask_shard_count(const ec_mask* m);
struct ec_shard_info {
    int logical_device;  /* index into the list given to ec_init_devices */
    int cuda_device;
    size_t offset, len;  /* cells [offset, offset + len) of the whole raster */
    void* device_ptr;    /* cells of the strip (mask: its packed words) on cuda_device */
};
/* strip `shard` (0 for a plain handle): where it lives and, if `strip` is given, a BORROWED plain handle of it that
 * stays valid as long as the parent and may be passed to any entry point (the work then runs on the strip's GPU) */
ec_status ec_buf_shard(const ec_buf* b, int shard, ec_shard_info* info_or_null, const ec_buf** strip_or_null);
ec_status ec_mask_shard(const ec_mask* m, int shard, ec_shard_info* info_or_null, const ec_mask** strip_or_null);
/* Strip `shard` of `n_shards` for a width x height row-major raster: whole rows, remainder rows to
 * the last shard, and every strip start a multiple of 128 cells (asserted: width % 128 == 0 or
 * n_shards == 1; otherwise the split falls back to 128-cell aligned cell ranges). */
ec_status ec_row_strip(size_t width, size_t height, int n_shards, int shard, size_t* cell_offset, size_t* cell_len);
/* Shard-local part of a reduction: leaves {min_key, ~max_key} as two int64 in device memory so one
 * MIN all-reduce (NCCL, e.g. torch.distributed or ec_comm_allreduce_min_i64) finishes it. Keys are
 * order-preserving integers, so the result is bit-identical for every shard count. */
ec_status ec_buf_min_max_keys(const ec_buf* b, const ec_mask* mask_or_null, int64_t* device_keys2);
ec_status ec_min_max_from_keys(uint8_t ct, const int64_t* host_keys2, ec_value* min_out, ec_value* max_out);
/* inverse of the above (host): pack a partial (min, max) of cell type ct as {skey(min), ~skey(max)} */
ec_status ec_min_max_to_keys(const ec_value* min_in, const ec_value* max_in, int64_t* host_keys2);
/* NCCL communicator over the GPUs of one box, one rank per process (libnccl is dlopen'ed lazily) */
ec_status ec_comm_unique_id(void* id128);  /* 128 bytes, created on rank 0 and shipped by the host */
ec_status ec_comm_init_rank(const void* id128, int n_ranks, int rank, ec_comm** out);
void ec_comm_destroy(ec_comm* c);
/* 1 when every rank could map every other rank's mailbox (cudaIpc over NVLink): the sharded reductions below then
 * run as ONE kernel per GPU — shard reduction, exchange of the 16-byte partials through peer memory and the final
 * fold in the finishing CTA — instead of kernel + NCCL all-reduce + copy. 0: the NCCL path is used. */
int ec_comm_peer_exchange(const ec_comm* c);
ec_status ec_comm_allreduce_min_i64(ec_comm* c, int64_t* device_buf, size_t count);
ec_status ec_comm_allreduce_sum_u64(ec_comm* c, uint64_t* device_buf, size_t count);
/* sharded min_max / counts: shard-local kernel + one all-reduce, result on every rank */
ec_status ec_buf_min_max_sharded(ec_comm* c, const ec_buf* shard, const ec_mask* mask_or_null,
                                 ec_value* min_out, ec_value* max_out);
ec_status ec_mask_counts_sharded(ec_comm* c, const ec_mask* shard, size_t* data, size_t* nodata);
/* statistics of the whole raster from this rank's strip (extension; every rank calls it and gets the same result) */
ec_status ec_buf_statistics_sharded(ec_comm* c, const ec_buf* shard, const ec_mask* mask_or_null, ec_statistics* out);

/* ---- synthetic rasters (bench/test utility): counter-based splitmix64(seed ^ index) ----------- */
/* cell i = f(h), h = splitmix64(seed ^ (index_offset + i)). kind 0: the low bits of h (uniform over
 * the full bit range of the type, NaN/inf/subnormals included); kind 1: uniform integer in [lo, hi]
 * cast to the type; kind 2: uniform real lo + u*(hi-lo), u = (h>>11)*2^-53, cast to the type. When
 * a sentinel is given, cells with splitmix64(h) % period == 0 are replaced by it. Host mirror:
 * erased_cells_b200/synth.py. */
ec_status ec_buf_synth(uint8_t ct, size_t len, uint64_t seed, uint64_t index_offset, int kind, double lo,
                       double hi, uint64_t period, const ec_value* sentinel_or_null, ec_buf** out);

#ifdef __cplusplus
}
#endif
#endif /* ERASED_CELLS_B200_H */
