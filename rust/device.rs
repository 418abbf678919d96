//! Replaced bodies of `src/buffer.rs` and `src/masked/*` over the C ABI (SOURCE ONLY, never compiled here).
//! The public names (`CellBuffer`, `BufferOps`, `Mask`, `MaskedCellBuffer`, the std::ops impls) are unchanged;
//! what changes is the backing store: an owning device handle instead of `Vec<$p>`.
use crate::error::{Error, Result};
use crate::ffi::*;
use crate::{BufferOps, CellEncoding, CellType, CellValue};
use std::ptr;

fn ct(v: u8) -> CellType { CellType::iter().nth(v as usize).unwrap() }
fn check(s: ec_status) -> Result<()> {
    match s {
        EC_OK => Ok(()),
        EC_NARROWING => {
            let (mut a, mut b) = (0u8, 0u8);
            unsafe { ec_last_narrowing(&mut a, &mut b) };
            Err(Error::NarrowingError { src: ct(a), dst: ct(b) })          // src/error.rs:14-15
        }
        EC_OOB => panic!("index out of bounds"),                             // src/lib.rs:136-147
        EC_LEN_MISMATCH => panic!("Mask and buffer must have the same length."), // src/masked/masked_buffer.rs:49-53
        _ => panic!("erased_cells_b200: {}", unsafe { std::ffi::CStr::from_ptr(ec_last_error()) }.to_string_lossy()),
    }
}
fn to_ffi(v: CellValue) -> ec_value { /* tag + little-endian payload, with_ct! match */ unimplemented!() }
fn from_ffi(v: ec_value) -> CellValue { /* inverse */ unimplemented!() }

/// Was: `pub enum CellBuffer { UInt8(Vec<u8>), ... }` (src/buffer.rs:52). The cell type now lives in the
/// handle; `cell_type()` asks it. Send + Sync: kernels only read inputs.
pub struct CellBuffer(*mut ec_buf);
unsafe impl Send for CellBuffer {}
unsafe impl Sync for CellBuffer {}
impl Drop for CellBuffer { fn drop(&mut self) { unsafe { ec_buf_free(self.0) } } }
impl Clone for CellBuffer {
    fn clone(&self) -> Self { let mut h = ptr::null_mut(); check(unsafe { ec_buf_clone(self.0, &mut h) }).unwrap(); Self(h) }
}

impl BufferOps for CellBuffer {
    fn from_vec<T: CellEncoding>(data: Vec<T>) -> Self {                     // src/buffer.rs:64-66
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_from_host(T::cell_type() as u8, data.as_ptr().cast(), data.len(), &mut h) }).unwrap();
        check(unsafe { ec_synchronize() }).unwrap();                         // `data` is dropped on return
        Self(h)
    }
    fn with_defaults(len: usize, c: CellType) -> Self { let mut h = ptr::null_mut(); check(unsafe { ec_buf_with_defaults(len, c as u8, &mut h) }).unwrap(); Self(h) }
    fn fill(len: usize, value: CellValue) -> Self { let mut h = ptr::null_mut(); check(unsafe { ec_buf_fill(len, &to_ffi(value), &mut h) }).unwrap(); Self(h) }
    fn fill_via<T: CellEncoding, F: Fn(usize) -> T>(len: usize, f: F) -> Self { Self::from_vec((0..len).map(f).collect()) }
    fn len(&self) -> usize { unsafe { ec_buf_len(self.0) } }
    fn cell_type(&self) -> CellType { ct(unsafe { ec_buf_ctype(self.0) }) }
    fn get(&self, index: usize) -> CellValue { let mut v = unsafe { std::mem::zeroed() }; check(unsafe { ec_buf_get(self.0, index, &mut v) }).unwrap(); from_ffi(v) }
    fn put(&mut self, index: usize, value: CellValue) -> Result<()> { check(unsafe { ec_buf_put(self.0, index, &to_ffi(value)) }) }
    fn convert(&self, c: CellType) -> Result<Self> { let mut h = ptr::null_mut(); check(unsafe { ec_buf_convert(self.0, c as u8, &mut h) })?; Ok(Self(h)) }
    fn min_max(&self) -> (CellValue, CellValue) {
        let (mut a, mut b) = unsafe { (std::mem::zeroed(), std::mem::zeroed()) };
        check(unsafe { ec_buf_min_max(self.0, ptr::null(), &mut a, &mut b) }).unwrap();
        (from_ffi(a), from_ffi(b))
    }
    fn to_vec<T: CellEncoding>(self) -> Result<Vec<T>> {                      // src/buffer.rs:175-185
        let r = self.convert(T::cell_type())?;
        let mut out = Vec::<T>::with_capacity(r.len());
        check(unsafe { ec_buf_to_host(r.0, out.as_mut_ptr().cast(), r.len() * std::mem::size_of::<T>()) })?;
        unsafe { out.set_len(r.len()) };
        Ok(out)
    }
}

macro_rules! cb_bin_op {                                                      // src/buffer.rs:321-358
    ($trt:ident, $mth:ident, $code:expr) => {
        impl std::ops::$trt for &CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: Self) -> CellBuffer { let mut h = ptr::null_mut(); check(unsafe { ec_buf_binary($code, self.0, rhs.0, &mut h) }).unwrap(); CellBuffer(h) }
        }
        impl std::ops::$trt for CellBuffer { type Output = CellBuffer; fn $mth(self, rhs: Self) -> CellBuffer { std::ops::$trt::$mth(&self, &rhs) } }
        impl std::ops::$trt<&CellBuffer> for CellBuffer { type Output = CellBuffer; fn $mth(self, rhs: &CellBuffer) -> CellBuffer { std::ops::$trt::$mth(&self, rhs) } }
        impl<R: Into<CellValue>> std::ops::$trt<R> for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: R) -> CellBuffer { let mut h = ptr::null_mut(); check(unsafe { ec_buf_scalar($code, self.0, &to_ffi(rhs.into()), &mut h) }).unwrap(); CellBuffer(h) }
        }
    };
}
cb_bin_op!(Add, add, 0); cb_bin_op!(Sub, sub, 1); cb_bin_op!(Mul, mul, 2); cb_bin_op!(Div, div, 3);
impl std::ops::Neg for &CellBuffer { type Output = CellBuffer; fn neg(self) -> CellBuffer { let mut h = ptr::null_mut(); check(unsafe { ec_buf_neg(self.0, &mut h) }).unwrap(); CellBuffer(h) } }
impl Ord for CellBuffer { fn cmp(&self, o: &Self) -> std::cmp::Ordering { let mut r = 0; check(unsafe { ec_buf_cmp(self.0, o.0, &mut r) }).unwrap(); r.cmp(&0) } }
// PartialOrd / PartialEq / Eq delegate to `cmp` as in src/buffer.rs:373-388.

/// Was `Mask(Vec<bool>)` (src/masked/mask.rs:12): packed bits on the device. `Index/IndexMut -> &bool` and
/// `iter_mut` cannot point into packed device bits; they are served from a host mirror filled by
/// `ec_mask_to_bools` and written back with `ec_mask_from_bools` on drop of the guard.
pub struct Mask(*mut ec_mask);
impl Drop for Mask { fn drop(&mut self) { unsafe { ec_mask_free(self.0) } } }
impl std::ops::BitAnd for &Mask { type Output = Mask; fn bitand(self, r: Self) -> Mask { let mut h = ptr::null_mut(); check(unsafe { ec_mask_and(self.0, r.0, &mut h) }).unwrap(); Mask(h) } }
// Not / BitOr / counts / all / fill / get / put map 1:1 onto ec_mask_not / _or / _counts / _all / _fill / _get / _put.

pub struct MaskedCellBuffer(CellBuffer, Mask);
impl std::ops::Sub for &MaskedCellBuffer {                                    // src/masked/masked_buffer.rs:326-336
    type Output = MaskedCellBuffer;
    fn sub(self, rhs: Self) -> MaskedCellBuffer {
        let (mut b, mut m) = (ptr::null_mut(), ptr::null_mut());
        check(unsafe { ec_masked_binary(1, (self.0).0, (self.1).0, (rhs.0).0, (rhs.1).0, &mut b, &mut m) }).unwrap();
        MaskedCellBuffer(CellBuffer(b), Mask(m))
    }
}
// from_vec_with_nodata -> ec_mask_from_nodata; to_vec_with_nodata -> ec_buf_fill_nodata + ec_buf_to_host;
// min_max -> ec_buf_min_max(buf, mask); counts -> ec_mask_counts.
