// Host-side internals shared by the translation units of liberased_cells_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

#include "../../include/erased_cells_b200.h"

namespace ec {
struct DevBlock;  // refcounted device allocation (returns to the caching allocator when the last user lets go)
struct Expr;      // a deferred op (lazy mode): evaluated, possibly fused with its children, on first use
}  // namespace ec

struct ec_buf {
    uint8_t ct;
    bool owned;
    size_t len;
    size_t capacity_bytes;
    void* dptr;         // null while `expr` is pending
    cudaEvent_t ready;  // set by ec_buf_from_host_async: readers on other streams wait on it
    std::shared_ptr<ec::DevBlock> blk;
    std::shared_ptr<ec::Expr> expr;
};
struct ec_mask {
    size_t len;
    size_t capacity_bytes;
    uint32_t* words;
};
struct ec_event {
    cudaEvent_t ev;
};

namespace ec {

struct ReduceScratch;
struct PeerExchange;
struct VmProgram;

struct Launch {  // launch context handed to every launcher
    cudaStream_t stream;
    int sm_count;
    int max_grid;  // cap for the persistent grids (<= 0: one tile per CTA)
    bool overlap;  // programmatic dependent launch: the grid may be scheduled while its predecessor drains
};

void set_error(const char* fmt, ...);
ec_status cuda_fail(cudaError_t e, const char* what);
void note_launch(const char* family);
int env_int(const char* name, int dflt);

// grid for a tiled streaming kernel
inline int grid_for(size_t n, size_t tile, const Launch& L) {
    size_t full = n / tile;
    if (full == 0) full = 1;
    size_t cap = L.max_grid > 0 ? size_t(L.max_grid) : (size_t(1) << 30);
    return int(full < cap ? full : cap);
}

// Launch of a kernel whose first statement is overlap_prologue() (ec_common.cuh). With L.overlap the grid carries the
// programmatic-stream-serialization attribute; memory ordering against the previous grid is kept by the prologue.
#ifdef __CUDACC__
template <class... P, class... A>
inline cudaError_t launch_k(const Launch& L, void (*kernel)(P...), int grid, int threads, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(threads));
    cfg.stream = L.stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = L.overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<A&&>(args)...);
}
#endif

// ---- launchers (each TU instantiates its kernel family) ----------------------------------------
// binary: out[i] = (f64)l[i] op (f64)r[i]; optional fused mask AND
cudaError_t launch_binary(const Launch& L, int op, int lct, const void* l, int rct, const void* r, double* out,
                          size_t n, const uint32_t* lm, const uint32_t* rm, uint32_t* om);
cudaError_t launch_scalar(const Launch& L, int op, int lct, const void* l, double s, double* out, size_t n);
cudaError_t launch_neg(const Launch& L, int ct, const void* a, void* out, size_t n);
cudaError_t launch_convert(const Launch& L, int sct, const void* a, int dct, void* out, size_t n);
cudaError_t launch_checked_cast(const Launch& L, int sct, const void* a, int dct, void* out, size_t n, unsigned int* fail_flag);
cudaError_t launch_copy(const Launch& L, int cell_bytes, const void* a, void* out, size_t n);
cudaError_t launch_fill(const Launch& L, int ct, void* out, size_t n, uint64_t bits);
cudaError_t launch_fill_nodata(const Launch& L, int sct, const void* a, const uint32_t* m, int dct, void* out,
                               size_t n, uint64_t nodata_bits);
cudaError_t launch_normdiff(const Launch& L, int lct, const void* l, int rct, const void* r, double* out, size_t n);
cudaError_t launch_binary_scalar(const Launch& L, int op1, int lct, const void* l, int rct, const void* r, int op2,
                                 double s, double* out, size_t n);
cudaError_t launch_binary_scalar_static(const Launch& L, int op1, int lct, const void* l, int rct, const void* r, int op2,
                                        double s, double* out, size_t n);
cudaError_t launch_scalar_scalar(const Launch& L, int op1, int ct, const void* a, double s1, int op2, double s2, double* out, size_t n);
cudaError_t launch_vm(const Launch& L, const VmProgram& p, double* out, size_t n);
// run-time specialised kernel of one pending expression (ec_jit.cu): `expr` is straight-line C over v0.. (operands as
// f64) and c0.. (scalars) built from ecj_add/sub/mul/div calls. Returns 0 = launched (*err = launch status),
// 1 = not available (no NVRTC / build failed; ec_last_error says why) -> the caller evaluates op by op.
constexpr int kJitInputs = 8, kJitConsts = 8, kJitOps = 48;
struct JitProgram {
    int n_in = 0, n_const = 0;
    const void* in[kJitInputs] = {};
    uint8_t ct[kJitInputs] = {};
    double consts[kJitConsts] = {};
    std::string expr;
};
int launch_jit(const Launch& L, const JitProgram& p, double* out, size_t n, cudaError_t* err);
size_t jit_cached_kernels();
size_t jit_builds();
int jit_dry_build(const JitProgram& p, std::string* source, std::string* log);
// reductions: results land in scratch.result[0..1] (device); keys are unsigned order keys
cudaError_t launch_min_max(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n,
                           const ReduceScratch& s);
cudaError_t launch_popcount(const Launch& L, const uint32_t* words, size_t nwords, const ReduceScratch& s, uint64_t second_word);
ec_status reduce_min_max_peer(const ec_buf* b, const ec_mask* m, const PeerExchange& px, uint64_t* k0, uint64_t* k1);
ec_status reduce_popcount_peer(const ec_mask* m, const PeerExchange& px, uint64_t* ones, uint64_t* len_sum);
// statistics extension: exact fixed-point moment sums into acc[9] (zeroed by the caller), see ec_stats.cuh
cudaError_t launch_moments(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n, double pivot, double scale,
                           unsigned long long* acc);
cudaError_t launch_int_stats(const Launch& L, int ct, const void* a, const uint32_t* mask, size_t n, unsigned long long* acc);
cudaError_t launch_first_diff(const Launch& L, int cell_bytes, const void* a, const void* b, size_t n,
                              const ReduceScratch& s);
// masks
cudaError_t launch_mask_build(const Launch& L, int cell_bytes, const void* a, size_t n, uint64_t sentinel_bits,
                              bool pack_bools, uint32_t* out);
cudaError_t launch_mask_unpack(const Launch& L, const uint32_t* m, size_t n, uint8_t* out);
cudaError_t launch_mask_bitop(const Launch& L, int mop, const uint32_t* l, const uint32_t* r, size_t n, uint32_t* out);
cudaError_t launch_mask_fill(const Launch& L, uint32_t* out, size_t n, bool value);
cudaError_t launch_synth(const Launch& L, int ct, void* out, size_t n, uint64_t seed, uint64_t index_offset, int kind,
                         double lo, double hi, uint64_t period, uint64_t sentinel_bits);
// keys <-> values on the host (same functions the kernels use)
void key_seeds(int ct, uint64_t* seed_min, uint64_t* seed_max);
uint64_t key_to_bits(int ct, uint64_t key);
uint64_t key_from_bits(int ct, uint64_t bits);
int64_t key_to_signed(uint64_t key);    // order-preserving int64 for NCCL min/max
uint64_t key_from_signed(int64_t skey);

}  // namespace ec
