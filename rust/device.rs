//! Replaced bodies of `src/buffer.rs` and `src/masked/*` over the C ABI.
//!
//! SOURCE ONLY — this image has no Rust toolchain, so this file has never been compiled; it is written to be
//! dropped into the crate next to `ffi.rs`. The public names (`CellBuffer`, `BufferOps`,
//! `Mask`, `MaskedCellBuffer`, `NoData`, the std::ops impls) are unchanged; what changes is the backing store:
//! an owning device handle instead of `Vec<$p>`. The same calls, compiled and tested, live in
//! `include/erased_cells.hpp` (C++) and `erased_cells_b200/api.py` (Python).
use crate::error::{Error, Result};
use crate::ffi::*;
use crate::{with_ct, BufferOps, CellEncoding, CellType, CellValue, NoData};
use std::cmp::Ordering;
use std::ops::{Add, BitAnd, BitOr, Div, Mul, Neg, Not, Sub};
use std::ptr;

// ---- plumbing ------------------------------------------------------------------------------------------
fn ct(v: u8) -> CellType {
    CellType::iter().nth(v as usize).expect("cell type discriminant")
}
fn check(s: ec_status) -> Result<()> {
    match s {
        EC_OK => Ok(()),
        EC_NARROWING => {
            let (mut a, mut b) = (0u8, 0u8);
            unsafe { ec_last_narrowing(&mut a, &mut b) };
            Err(Error::NarrowingError { src: ct(a), dst: ct(b) }) // src/error.rs:14-15
        }
        EC_OOB => panic!("index out of bounds"),                                   // src/lib.rs:136-147
        EC_LEN_MISMATCH => panic!("Mask and buffer must have the same length."),   // src/masked/masked_buffer.rs:49-53
        _ => panic!("erased_cells_b200: {}", unsafe { std::ffi::CStr::from_ptr(ec_last_error()) }.to_string_lossy()),
    }
}
/// `CellValue` <-> the 16-byte `ec_value` (tag + little-endian payload in the low bytes).
fn to_ffi(v: CellValue) -> ec_value {
    let mut out = ec_value { ct: v.cell_type() as u8, pad: [0; 7], bits: 0 };
    macro_rules! put {
        ($( ($id:ident, $p:ident) ),*) => {
            match v { $( CellValue::$id(x) => {
                let b = x.to_le_bytes();
                let mut w = [0u8; 8];
                w[..b.len()].copy_from_slice(&b);
                out.bits = u64::from_le_bytes(w);
            } )* }
        };
    }
    with_ct!(put);
    out
}
fn from_ffi(v: ec_value) -> CellValue {
    let w = v.bits.to_le_bytes();
    macro_rules! get {
        ($( ($id:ident, $p:ident) ),*) => {
            match ct(v.ct) { $( CellType::$id => {
                let mut b = [0u8; std::mem::size_of::<$p>()];
                b.copy_from_slice(&w[..std::mem::size_of::<$p>()]);
                CellValue::$id(<$p>::from_le_bytes(b))
            } )* }
        };
    }
    with_ct!(get)
}
fn ordering(r: i32) -> Ordering {
    r.cmp(&0)
}

// ---- CellBuffer — was `pub enum CellBuffer { UInt8(Vec<u8>), ... }` (src/buffer.rs:52) ----------------------
/// The cell type now lives in the handle. Send + Sync: kernels only read their inputs.
pub struct CellBuffer(*mut ec_buf);
unsafe impl Send for CellBuffer {}
unsafe impl Sync for CellBuffer {}
impl Drop for CellBuffer {
    fn drop(&mut self) {
        unsafe { ec_buf_free(self.0) }
    }
}
impl Clone for CellBuffer {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_clone(self.0, &mut h) }).unwrap();
        Self(h)
    }
}
impl CellBuffer {
    pub fn new<T: CellEncoding>(data: Vec<T>) -> Self {
        Self::from_vec(data)
    }
    /// Extension: cells `[offset, offset + len)` as a buffer sharing this allocation (a row strip of a resident
    /// raster; `offset` on a 32-byte boundary). No copy; a later `put` / `extend` through either handle copies first.
    pub fn view(&self, offset: usize, len: usize) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_view(self.0, offset, len, &mut h) }).unwrap();
        Self(h)
    }
}

impl BufferOps for CellBuffer {
    fn from_vec<T: CellEncoding>(data: Vec<T>) -> Self {
        // src/buffer.rs:64-66. `data` is owned, so the copy could also be left in flight with
        // ec_buf_from_host_async and the Vec parked in the handle until ec_buf_wait.
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_from_host(T::cell_type() as u8, data.as_ptr().cast(), data.len(), &mut h) }).unwrap();
        check(unsafe { ec_synchronize() }).unwrap();
        Self(h)
    }
    fn with_defaults(len: usize, c: CellType) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_with_defaults(len, c as u8, &mut h) }).unwrap();
        Self(h)
    }
    fn fill(len: usize, value: CellValue) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_fill(len, &to_ffi(value), &mut h) }).unwrap();
        Self(h)
    }
    fn fill_via<T: CellEncoding, F: Fn(usize) -> T>(len: usize, f: F) -> Self {
        Self::from_vec((0..len).map(f).collect()) // the closure is host code in the reference too
    }
    fn len(&self) -> usize {
        unsafe { ec_buf_len(self.0) }
    }
    fn cell_type(&self) -> CellType {
        ct(unsafe { ec_buf_ctype(self.0) })
    }
    fn get(&self, index: usize) -> CellValue {
        let mut v = ec_value { ct: 0, pad: [0; 7], bits: 0 };
        check(unsafe { ec_buf_get(self.0, index, &mut v) }).unwrap();
        from_ffi(v)
    }
    fn put(&mut self, index: usize, value: CellValue) -> Result<()> {
        check(unsafe { ec_buf_put(self.0, index, &to_ffi(value)) })
    }
    fn convert(&self, c: CellType) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_convert(self.0, c as u8, &mut h) })?;
        Ok(Self(h))
    }
    fn min_max(&self) -> (CellValue, CellValue) {
        let (mut a, mut b) = (ec_value { ct: 0, pad: [0; 7], bits: 0 }, ec_value { ct: 0, pad: [0; 7], bits: 0 });
        check(unsafe { ec_buf_min_max(self.0, ptr::null(), &mut a, &mut b) }).unwrap();
        (from_ffi(a), from_ffi(b))
    }
    fn to_vec<T: CellEncoding>(self) -> Result<Vec<T>> {
        // src/buffer.rs:175-185: convert on the device, one D2H copy
        let r = self.convert(T::cell_type())?;
        let n = if r.cell_type() == T::cell_type() { r.len() } else { 0 }; // empty non-identity convert is UInt8([])
        let mut out = Vec::<T>::with_capacity(n);
        check(unsafe { ec_buf_to_host(r.0, out.as_mut_ptr().cast(), n * std::mem::size_of::<T>()) })?;
        unsafe { out.set_len(n) };
        Ok(out)
    }
}
impl<T: CellEncoding> From<Vec<T>> for CellBuffer {
    fn from(values: Vec<T>) -> Self {
        Self::from_vec(values)
    }
}
impl<T: CellEncoding> From<&[T]> for CellBuffer {
    fn from(values: &[T]) -> Self {
        // src/buffer.rs:265-276
        Self::from_vec(values.to_vec())
    }
}
impl<C: CellEncoding> FromIterator<C> for CellBuffer {
    fn from_iter<I: IntoIterator<Item = C>>(iter: I) -> Self {
        // src/buffer.rs:223-227
        Self::from_vec(iter.into_iter().collect())
    }
}
impl FromIterator<CellValue> for CellBuffer {
    fn from_iter<I: IntoIterator<Item = CellValue>>(iterable: I) -> Self {
        // src/buffer.rs:229-250: the first element decides the cell type; nothing collected => UInt8([])
        let values: Vec<CellValue> = iterable.into_iter().collect();
        match values.first() {
            None => Self::with_defaults(0, CellType::UInt8),
            Some(first) => {
                let ct = first.cell_type();
                macro_rules! collect {
                    ($(($id:ident, $p:ident)),*) => {
                        match ct {
                            $(CellType::$id => Self::from_vec(values.iter().map(|v| v.get::<$p>().unwrap()).collect::<Vec<$p>>()),)*
                        }
                    };
                }
                with_ct!(collect)
            }
        }
    }
}
impl<C: CellEncoding> TryFrom<CellBuffer> for Vec<C> {
    type Error = Error;
    fn try_from(value: CellBuffer) -> Result<Self> {
        // src/buffer.rs:307-313
        value.to_vec::<C>()
    }
}
impl<'buf> IntoIterator for &'buf CellBuffer {
    type Item = CellValue;
    type IntoIter = std::vec::IntoIter<CellValue>;
    fn into_iter(self) -> Self::IntoIter {
        // src/buffer.rs:278-305 yields host CellValues: one D2H copy of the buffer, then iterate that
        (0..self.len()).map(|i| self.get(i)).collect::<Vec<_>>().into_iter() // a shim would download once; shown per cell for brevity
    }
}
impl<C: CellEncoding> Extend<C> for CellBuffer {
    fn extend<I: IntoIterator<Item = C>>(&mut self, iter: I) {
        // src/buffer.rs:205-221: value-checked `to_<p>().unwrap()` runs on the device; EC_NARROWING is that unwrap's panic
        let v: Vec<C> = iter.into_iter().collect();
        check(unsafe { ec_buf_extend_host(self.0, C::cell_type() as u8, v.as_ptr().cast(), v.len()) }).expect("Extend: value out of range");
    }
}

macro_rules! cb_bin_op {
    // src/buffer.rs:321-358
    ($trt:ident, $mth:ident, $code:expr) => {
        impl $trt for &CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: Self) -> CellBuffer {
                let mut h = ptr::null_mut();
                check(unsafe { ec_buf_binary($code, self.0, rhs.0, &mut h) }).unwrap(); // arithmetic never errors
                CellBuffer(h)
            }
        }
        impl $trt for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: Self) -> CellBuffer {
                $trt::$mth(&self, &rhs)
            }
        }
        impl $trt<&CellBuffer> for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: &CellBuffer) -> CellBuffer {
                $trt::$mth(&self, rhs)
            }
        }
        impl<R: Into<CellValue>> $trt<R> for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: R) -> CellBuffer {
                let mut h = ptr::null_mut();
                check(unsafe { ec_buf_scalar($code, self.0, &to_ffi(rhs.into()), &mut h) }).unwrap();
                CellBuffer(h)
            }
        }
    };
}
cb_bin_op!(Add, add, 0);
cb_bin_op!(Sub, sub, 1);
cb_bin_op!(Mul, mul, 2);
cb_bin_op!(Div, div, 3);
impl Neg for &CellBuffer {
    type Output = CellBuffer;
    fn neg(self) -> CellBuffer {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_neg(self.0, &mut h) }).unwrap();
        CellBuffer(h)
    }
}
impl Neg for CellBuffer {
    type Output = CellBuffer;
    fn neg(self) -> CellBuffer {
        Neg::neg(&self)
    }
}
impl Ord for CellBuffer {
    // src/buffer.rs:390-436 on the device: first differing cell, then total order of that pair, then length
    fn cmp(&self, o: &Self) -> Ordering {
        let mut r = 0;
        check(unsafe { ec_buf_cmp(self.0, o.0, &mut r) }).unwrap();
        ordering(r)
    }
}
impl PartialOrd for CellBuffer {
    fn partial_cmp(&self, o: &Self) -> Option<Ordering> {
        Some(self.cmp(o))
    }
}
impl PartialEq for CellBuffer {
    fn eq(&self, o: &Self) -> bool {
        self.cmp(o) == Ordering::Equal
    }
}
impl Eq for CellBuffer {}

// ---- Mask — was `Mask(Vec<bool>)` (src/masked/mask.rs:12): packed bits on the device ------------------------
/// `Index/IndexMut -> &bool` and `iter_mut` cannot point into packed device bits: they go through a host mirror
/// (`to_vec` = ec_mask_to_bools, write back with `Mask::new`), see INTEGRATION.md §4.
pub struct Mask(*mut ec_mask);
unsafe impl Send for Mask {}
unsafe impl Sync for Mask {}
impl Drop for Mask {
    fn drop(&mut self) {
        unsafe { ec_mask_free(self.0) }
    }
}
impl Clone for Mask {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_clone(self.0, &mut h) }).unwrap();
        Self(h)
    }
}
impl Mask {
    pub fn new(values: Vec<bool>) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_from_bools(values.as_ptr().cast(), values.len(), &mut h) }).unwrap(); // bool is one byte, 0/1
        check(unsafe { ec_synchronize() }).unwrap();
        Self(h)
    }
    pub fn fill(len: usize, value: bool) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_fill(len, value as i32, &mut h) }).unwrap();
        Self(h)
    }
    pub fn fill_via<F: Fn(usize) -> bool>(len: usize, f: F) -> Self {
        Self::new((0..len).map(f).collect())
    }
    /// `Extend<bool>` (src/masked/mask.rs:83-87): appended on the device (unpack, append, repack)
    pub fn extend_from<I: IntoIterator<Item = bool>>(&mut self, iter: I) {
        let v: Vec<bool> = iter.into_iter().collect();
        check(unsafe { ec_mask_extend_host(self.0, v.as_ptr().cast(), v.len()) }).unwrap();
    }
    /// Extension: the validity bits of cells `[offset, offset + len)` as a mask of its own (`offset` a multiple of 128).
    pub fn slice(&self, offset: usize, len: usize) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_slice(self.0, offset, len, &mut h) }).unwrap();
        Self(h)
    }
    pub fn len(&self) -> usize {
        unsafe { ec_mask_len(self.0) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn put(&mut self, index: usize, value: bool) {
        check(unsafe { ec_mask_put(self.0, index, value as i32) }).unwrap()
    }
    pub fn get(&self, index: usize) -> bool {
        let mut o = 0;
        check(unsafe { ec_mask_get(self.0, index, &mut o) }).unwrap();
        o != 0
    }
    pub fn all(&self, value: bool) -> bool {
        let mut o = 0;
        check(unsafe { ec_mask_all(self.0, value as i32, &mut o) }).unwrap();
        o != 0
    }
    pub fn counts(&self) -> (usize, usize) {
        let (mut d, mut n) = (0usize, 0usize);
        check(unsafe { ec_mask_counts(self.0, &mut d, &mut n) }).unwrap();
        (d, n)
    }
    pub fn to_vec(&self) -> Vec<bool> {
        let mut out = vec![false; self.len()];
        check(unsafe { ec_mask_to_bools(self.0, out.as_mut_ptr().cast(), out.len()) }).unwrap();
        out
    }
}
macro_rules! mask_op {
    ($trt:ident, $mth:ident, $f:ident) => {
        impl $trt for &Mask {
            type Output = Mask;
            fn $mth(self, rhs: Self) -> Mask {
                let mut h = ptr::null_mut();
                check(unsafe { $f(self.0, rhs.0, &mut h) }).unwrap();
                Mask(h)
            }
        }
        impl $trt for Mask {
            type Output = Mask;
            fn $mth(self, rhs: Self) -> Mask {
                $trt::$mth(&self, &rhs)
            }
        }
    };
}
mask_op!(BitAnd, bitand, ec_mask_and); // src/masked/mask.rs:118-140
mask_op!(BitOr, bitor, ec_mask_or); // src/masked/mask.rs:142-164
impl Not for &Mask {
    type Output = Mask;
    fn not(self) -> Mask {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_not(self.0, &mut h) }).unwrap();
        Mask(h)
    }
}
impl PartialEq for Mask {
    fn eq(&self, o: &Self) -> bool {
        let mut r = 0;
        check(unsafe { ec_mask_cmp(self.0, o.0, &mut r) }).unwrap();
        r == 0
    }
}

// ---- MaskedCellBuffer — src/masked/masked_buffer.rs ------------------------------------------------------------
fn nodata_ffi<T: CellEncoding>(nd: NoData<T>) -> (i32, ec_value) {
    match nd {
        NoData::None => (0, to_ffi(T::zero().into_cell_value())),
        NoData::Default => (1, to_ffi(T::zero().into_cell_value())),
        NoData::Value(v) => (2, to_ffi(v.into_cell_value())),
    }
}
#[derive(Clone, PartialEq)]
pub struct MaskedCellBuffer(CellBuffer, Mask);
impl MaskedCellBuffer {
    pub fn new(buffer: CellBuffer, mask: Mask) -> Self {
        assert_eq!(buffer.len(), mask.len(), "Mask and buffer must have the same length.");
        Self(buffer, mask)
    }
    /// src/masked/masked_buffer.rs:62-71 — the sentinel compare runs on the device and emits packed words
    pub fn from_vec_with_nodata<T: CellEncoding>(data: Vec<T>, nodata: NoData<T>) -> Self {
        let buf = CellBuffer::from_vec(data);
        let (kind, v) = nodata_ffi(nodata);
        let mut m = ptr::null_mut();
        check(unsafe { ec_mask_from_nodata(buf.0, kind, &v, &mut m) }).unwrap();
        Self(buf, Mask(m))
    }
    /// Extension: a row strip of a resident masked raster (`offset` a multiple of 128 cells, see `ec_row_strip`).
    pub fn view(&self, offset: usize, len: usize) -> Self {
        Self(self.0.view(offset, len), self.1.slice(offset, len))
    }
    pub fn buffer(&self) -> &CellBuffer {
        &self.0
    }
    pub fn buffer_mut(&mut self) -> &mut CellBuffer {
        &mut self.0
    }
    pub fn mask_mut(&mut self) -> &mut Mask {
        &mut self.1
    }
    /// src/masked/masked_buffer.rs:73-79: the closure is host code in the reference too; one upload of both halves
    pub fn fill_with_mask_via<T: CellEncoding, F: Fn(usize) -> (T, bool)>(len: usize, mv: F) -> Self {
        let (data, mask): (Vec<T>, Vec<bool>) = (0..len).map(mv).unzip();
        Self::new(CellBuffer::from_vec(data), Mask::new(mask))
    }
    /// src/masked/masked_buffer.rs:112-118
    pub fn get_with_mask(&self, index: usize) -> (CellValue, bool) {
        (self.0.get(index), self.1.get(index))
    }
    /// src/masked/masked_buffer.rs:120-130
    pub fn put_with_mask(&mut self, index: usize, value: CellValue, mask: bool) -> Result<()> {
        self.0.put(index, value)?;
        self.1.put(index, mask);
        Ok(())
    }
    pub fn mask(&self) -> &Mask {
        &self.1
    }
    pub fn counts(&self) -> (usize, usize) {
        self.1.counts()
    }
    pub fn get_masked(&self, index: usize) -> Option<CellValue> {
        if self.1.get(index) { Some(self.0.get(index)) } else { None }
    }
    /// src/masked/masked_buffer.rs:137-152 — convert + fill fused into one pass, then one D2H copy
    pub fn to_vec_with_nodata<T: CellEncoding>(self, no_data: NoData<T>) -> Result<Vec<T>> {
        let (kind, v) = nodata_ffi(no_data);
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_fill_nodata((self.0).0, (self.1).0, T::cell_type() as u8, kind, &v, &mut h) })?;
        CellBuffer(h).to_vec::<T>()
    }
    /// src/masked/masked_buffer.rs:208-217
    pub fn min_max(&self) -> (CellValue, CellValue) {
        let (mut a, mut b) = (ec_value { ct: 0, pad: [0; 7], bits: 0 }, ec_value { ct: 0, pad: [0; 7], bits: 0 });
        check(unsafe { ec_buf_min_max((self.0).0, (self.1).0, &mut a, &mut b) }).unwrap();
        (from_ffi(a), from_ffi(b))
    }
    /// Extension (not in the crate): count / min / max / mean / population stddev of the valid cells.
    pub fn statistics(&self) -> Statistics {
        let z = ec_value { ct: 0, pad: [0; 7], bits: 0 };
        let mut s = ec_statistics { count: 0, min: z, max: z, mean: 0.0, stddev: 0.0 };
        check(unsafe { ec_buf_statistics((self.0).0, (self.1).0, &mut s) }).unwrap();
        Statistics { count: s.count, min: from_ffi(s.min), max: from_ffi(s.max), mean: s.mean, stddev: s.stddev }
    }
}
impl BufferOps for MaskedCellBuffer {
    // src/masked/masked_buffer.rs:155-225: the buffer half does the work, the mask half is all-valid or carried along
    fn from_vec<T: CellEncoding>(data: Vec<T>) -> Self {
        CellBuffer::from_vec(data).into()
    }
    fn with_defaults(len: usize, ct: CellType) -> Self {
        CellBuffer::with_defaults(len, ct).into()
    }
    fn fill(len: usize, value: CellValue) -> Self {
        CellBuffer::fill(len, value).into()
    }
    fn fill_via<T: CellEncoding, F: Fn(usize) -> T>(len: usize, f: F) -> Self {
        CellBuffer::fill_via(len, f).into()
    }
    fn len(&self) -> usize {
        self.0.len()
    }
    fn cell_type(&self) -> CellType {
        self.0.cell_type()
    }
    fn get(&self, index: usize) -> CellValue {
        self.0.get(index)
    }
    fn put(&mut self, idx: usize, value: CellValue) -> Result<()> {
        self.0.put(idx, value)
    }
    fn convert(&self, cell_type: CellType) -> Result<Self> {
        Ok(Self(self.0.convert(cell_type)?, self.1.clone()))
    }
    fn min_max(&self) -> (CellValue, CellValue) {
        MaskedCellBuffer::min_max(self)
    }
    fn to_vec<T: CellEncoding>(self) -> Result<Vec<T>> {
        self.0.to_vec()
    }
}
impl From<CellBuffer> for MaskedCellBuffer {
    fn from(value: CellBuffer) -> Self {
        // src/masked/masked_buffer.rs:250-255
        let len = value.len();
        Self::new(value, Mask::fill(len, true))
    }
}
impl From<MaskedCellBuffer> for (CellBuffer, Mask) {
    fn from(value: MaskedCellBuffer) -> Self {
        (value.0, value.1)
    }
}
impl<C: CellEncoding> FromIterator<C> for MaskedCellBuffer {
    fn from_iter<I: IntoIterator<Item = C>>(iter: I) -> Self {
        // src/masked/masked_buffer.rs:257-261
        <Self as BufferOps>::from_vec(iter.into_iter().collect())
    }
}
impl<C: CellEncoding> FromIterator<(C, bool)> for MaskedCellBuffer {
    fn from_iter<I: IntoIterator<Item = (C, bool)>>(iter: I) -> Self {
        // src/masked/masked_buffer.rs:263-278 (an empty iterator still gives cells of type C)
        let (data, mask): (Vec<C>, Vec<bool>) = iter.into_iter().unzip();
        Self::new(CellBuffer::from_vec(data), Mask::new(mask))
    }
}
impl<C: CellEncoding> Extend<(C, bool)> for MaskedCellBuffer {
    fn extend<I: IntoIterator<Item = (C, bool)>>(&mut self, iter: I) {
        // src/masked/masked_buffer.rs:280-287, as two bulk appends instead of one per pair
        let (data, mask): (Vec<C>, Vec<bool>) = iter.into_iter().unzip();
        self.0.extend(data);
        self.1.extend_from(mask);
    }
}
impl Neg for &MaskedCellBuffer {
    type Output = MaskedCellBuffer;
    fn neg(self) -> MaskedCellBuffer {
        // src/masked/masked_buffer.rs:372-383: cells negated (also under a false mask), mask cloned
        MaskedCellBuffer(-&self.0, self.1.clone())
    }
}
impl Neg for MaskedCellBuffer {
    type Output = MaskedCellBuffer;
    fn neg(self) -> MaskedCellBuffer {
        MaskedCellBuffer(-self.0, self.1)
    }
}
/// Result of the `statistics()` extension.
#[derive(Debug, Clone, Copy, PartialEq)]
pub struct Statistics {
    pub count: u64,
    pub min: CellValue,
    pub max: CellValue,
    pub mean: f64,
    pub stddev: f64,
}
macro_rules! mcb_bin_op {
    // src/masked/masked_buffer.rs:323-370: data on all cells, mask = lmask & rmask, one fused launch
    ($trt:ident, $mth:ident, $code:expr) => {
        impl $trt for &MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: Self) -> MaskedCellBuffer {
                let (mut b, mut m) = (ptr::null_mut(), ptr::null_mut());
                check(unsafe { ec_masked_binary($code, (self.0).0, (self.1).0, (rhs.0).0, (rhs.1).0, &mut b, &mut m) }).unwrap();
                MaskedCellBuffer(CellBuffer(b), Mask(m))
            }
        }
        impl $trt for MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: Self) -> MaskedCellBuffer {
                $trt::$mth(&self, &rhs)
            }
        }
        impl<R: Into<CellValue>> $trt<R> for MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: R) -> MaskedCellBuffer {
                let MaskedCellBuffer(buf, mask) = self;
                MaskedCellBuffer($trt::$mth(buf, rhs), mask) // the mask moves through unchanged (:353-364)
            }
        }
    };
}
mcb_bin_op!(Add, add, 0);
mcb_bin_op!(Sub, sub, 1);
mcb_bin_op!(Mul, mul, 2);
mcb_bin_op!(Div, div, 3);
