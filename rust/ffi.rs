//! `extern "C"` bindings of include/erased_cells_b200.h for the erased-cells crate.
//!
//! SOURCE ONLY: this image has no Rust toolchain (cargo/rustc absent), so this file has never been
//! compiled. It is the binding a maintainer adds as `src/ffi.rs`; `device.rs` next to it holds the
//! replaced bodies of `src/buffer.rs` / `src/masked/*`. Link with
//! `cargo:rustc-link-lib=dylib=erased_cells_b200` from build.rs.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone)]
pub struct ec_value {
    pub ct: u8,
    pub pad: [u8; 7],
    pub bits: u64,
}
/// extension: statistics of the valid cells (the crate itself stops at min_max / counts)
#[repr(C)]
#[derive(Copy, Clone)]
pub struct ec_statistics {
    pub count: u64,
    pub min: ec_value,
    pub max: ec_value,
    pub mean: f64,
    pub stddev: f64,
}
#[repr(C)] pub struct ec_buf { _p: [u8; 0] }
#[repr(C)] pub struct ec_mask { _p: [u8; 0] }
#[repr(C)] pub struct ec_comm { _p: [u8; 0] }
pub type ec_status = c_int;
pub const EC_OK: ec_status = 0;
pub const EC_NARROWING: ec_status = 1;
pub const EC_OOB: ec_status = 2;
pub const EC_LEN_MISMATCH: ec_status = 3;

#[link(name = "erased_cells_b200")]
extern "C" {
    pub fn ec_last_error() -> *const c_char;
    pub fn ec_last_narrowing(src: *mut u8, dst: *mut u8);
    pub fn ec_init(device: c_int) -> ec_status;
    pub fn ec_synchronize() -> ec_status;
    // CellBuffer
    pub fn ec_buf_from_host(ct: u8, host: *const c_void, len: usize, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_with_defaults(len: usize, ct: u8, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_fill(len: usize, value: *const ec_value, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_clone(b: *const ec_buf, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_free(b: *mut ec_buf);
    pub fn ec_buf_len(b: *const ec_buf) -> usize;
    pub fn ec_buf_ctype(b: *const ec_buf) -> u8;
    pub fn ec_buf_to_host(b: *const ec_buf, host: *mut c_void, host_bytes: usize) -> ec_status;
    pub fn ec_buf_get(b: *const ec_buf, index: usize, out: *mut ec_value) -> ec_status;
    pub fn ec_buf_put(b: *mut ec_buf, index: usize, value: *const ec_value) -> ec_status;
    pub fn ec_buf_extend_host(b: *mut ec_buf, ct: u8, host: *const c_void, n: usize) -> ec_status;
    pub fn ec_buf_from_host_async(ct: u8, host: *const c_void, len: usize, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_wait(b: *const ec_buf) -> ec_status;
    pub fn ec_buf_view(b: *const ec_buf, offset_cells: usize, len: usize, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_set_lazy(mode: c_int) -> ec_status;
    pub fn ec_set_launch_overlap(on: c_int) -> c_int;
    pub fn ec_buf_binary(op: c_int, l: *const ec_buf, r: *const ec_buf, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_scalar(op: c_int, l: *const ec_buf, r: *const ec_value, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_neg(b: *const ec_buf, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_convert(b: *const ec_buf, ct: u8, out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_buf_min_max(b: *const ec_buf, mask: *const ec_mask, mn: *mut ec_value, mx: *mut ec_value) -> ec_status;
    pub fn ec_buf_statistics(b: *const ec_buf, mask: *const ec_mask, out: *mut ec_statistics) -> ec_status;
    pub fn ec_buf_cmp(l: *const ec_buf, r: *const ec_buf, ordering: *mut c_int) -> ec_status;
    // Mask
    pub fn ec_mask_from_bools(bools: *const u8, len: usize, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_fill(len: usize, value: c_int, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_to_bools(m: *const ec_mask, bools: *mut u8, capacity: usize) -> ec_status;
    pub fn ec_mask_clone(m: *const ec_mask, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_extend_host(m: *mut ec_mask, host_bools: *const u8, n: usize) -> ec_status;
    pub fn ec_mask_slice(m: *const ec_mask, offset_cells: usize, len: usize, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_free(m: *mut ec_mask);
    pub fn ec_mask_len(m: *const ec_mask) -> usize;
    pub fn ec_mask_get(m: *const ec_mask, index: usize, out: *mut c_int) -> ec_status;
    pub fn ec_mask_put(m: *mut ec_mask, index: usize, value: c_int) -> ec_status;
    pub fn ec_mask_not(m: *const ec_mask, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_and(l: *const ec_mask, r: *const ec_mask, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_or(l: *const ec_mask, r: *const ec_mask, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_mask_counts(m: *const ec_mask, data: *mut usize, nodata: *mut usize) -> ec_status;
    pub fn ec_mask_all(m: *const ec_mask, value: c_int, out: *mut c_int) -> ec_status;
    pub fn ec_mask_cmp(l: *const ec_mask, r: *const ec_mask, ordering: *mut c_int) -> ec_status;
    // MaskedCellBuffer / NoData
    pub fn ec_mask_from_nodata(b: *const ec_buf, kind: c_int, value: *const ec_value, out: *mut *mut ec_mask) -> ec_status;
    pub fn ec_buf_fill_nodata(b: *const ec_buf, m: *const ec_mask, dst_ct: u8, kind: c_int, value: *const ec_value,
                              out: *mut *mut ec_buf) -> ec_status;
    pub fn ec_masked_binary(op: c_int, lb: *const ec_buf, lm: *const ec_mask, rb: *const ec_buf, rm: *const ec_mask,
                            out_buf: *mut *mut ec_buf, out_mask: *mut *mut ec_mask) -> ec_status;
    // sharding
    pub fn ec_row_strip(width: usize, height: usize, n: c_int, shard: c_int, off: *mut usize, len: *mut usize) -> ec_status;
    pub fn ec_comm_unique_id(id128: *mut c_void) -> ec_status;
    pub fn ec_comm_init_rank(id128: *const c_void, n: c_int, rank: c_int, out: *mut *mut ec_comm) -> ec_status;
    pub fn ec_buf_min_max_sharded(c: *mut ec_comm, shard: *const ec_buf, mask: *const ec_mask, mn: *mut ec_value,
                                  mx: *mut ec_value) -> ec_status;
}
