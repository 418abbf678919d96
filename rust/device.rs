//! Replaced bodies of `src/buffer.rs` and `src/masked/*` over the C ABI.
//!
//! SOURCE ONLY — this image has no Rust toolchain, so this file has never been compiled; it is written to be
//! dropped into the crate next to `ffi.rs`. The public surface is the crate's own:
//!
//! * `pub enum CellBuffer { UInt8(..), UInt16(..), … }` keeps its ten variants (src/buffer.rs:12-55), so code that
//!   matches on them — the GDAL adapter does, src/gdal/rasterband.rs:118 — still compiles. What a variant holds changes
//!   from `Vec<$p>` to `DeviceVec<$p>`: a `#[repr(transparent)]` owning handle to `$p` cells in HBM (on one GPU, or in
//!   row strips over all the GPUs the library was initialised with — the handle hides which).
//! * `BufferOps`, the `std::ops` impls, `Mask` (with `Index` / `IndexMut` / `iter_mut` over a host mirror that is
//!   written back before the next device use), `NoData`, `MaskedCellBuffer`, and the serde wire format of the
//!   derives (src/buffer.rs:51, src/masked/mask.rs:11, src/masked/masked_buffer.rs:40) are unchanged.
//!
//! The same calls, compiled and tested, live in `include/erased_cells.hpp` (C++) and `erased_cells_b200/api.py`
//! (Python); `ffi.rs` is generated from the C header and checked against it by tests/test_abi_cpu.py.
use crate::error::{Error, Result};
use crate::ffi::*;
use crate::{with_ct, BufferOps, CellEncoding, CellType, CellValue, NoData};
use std::cmp::Ordering;
use std::fmt::{Debug, Formatter};
use std::marker::PhantomData;
use std::ops::{Add, BitAnd, BitOr, Div, Index, IndexMut, Mul, Neg, Not, Sub};
use std::ptr;
use std::sync::Mutex;

#[cfg(feature = "serde")]
use serde::{Deserialize, Deserializer, Serialize, Serializer};

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
const ZERO_VALUE: ec_value = ec_value { ct: 0, pad: [0; 7], bits: 0 };
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

/// Extension: one process, several GPUs. Call once before the first buffer is made (or set `EC_DEVICES=0,1,…`): from
/// then on buffers and masks of at least `ec_set_shard_min_cells` cells live as row strips, one per GPU, behind the
/// very same `CellBuffer` / `Mask` values; element-wise ops stay strip-local, reductions finish across the GPUs.
pub fn init_devices(devices: &[i32]) -> Result<()> {
    check(unsafe { ec_init_devices(devices.as_ptr(), devices.len() as i32) })
}

/// Extension: host threads that move a `Vec<T>` (pageable memory) to / from pinned staging while the DMA engine copies the
/// previous chunks (`from_vec`, `to_vec`, `Mask::new` of at least 16 MiB). 0 leaves such copies to the CUDA driver.
/// Returns the previous setting.
pub fn set_host_copy_threads(threads: i32) -> i32 {
    unsafe { ec_set_host_copy_threads(threads) }
}

// ---- DeviceVec<T>: what a CellBuffer variant holds instead of Vec<T> ------------------------------------------
/// `len` cells of `T` in HBM, owned. `#[repr(transparent)]` over the C handle: no cost over the raw pointer.
/// Send + Sync: kernels only read their inputs, handles are uniquely owned.
#[repr(transparent)]
pub struct DeviceVec<T: CellEncoding> {
    h: *mut ec_buf,
    _cells: PhantomData<T>,
}
unsafe impl<T: CellEncoding> Send for DeviceVec<T> {}
unsafe impl<T: CellEncoding> Sync for DeviceVec<T> {}
impl<T: CellEncoding> DeviceVec<T> {
    /// Takes ownership of a handle whose cell type is `T`'s.
    unsafe fn from_raw(h: *mut ec_buf) -> Self {
        debug_assert_eq!(ec_buf_ctype(h), T::cell_type() as u8);
        Self { h, _cells: PhantomData }
    }
    pub fn len(&self) -> usize {
        unsafe { ec_buf_len(self.h) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    /// One D2H copy (all strips side by side when the buffer is sharded).
    pub fn to_vec(&self) -> Vec<T> {
        let n = self.len();
        let mut out = Vec::<T>::with_capacity(n);
        check(unsafe { ec_buf_to_host(self.h, out.as_mut_ptr().cast(), n * std::mem::size_of::<T>()) }).unwrap();
        unsafe { out.set_len(n) };
        out
    }
    /// Row strips this buffer is kept as (0: one GPU).
    pub fn shard_count(&self) -> usize {
        unsafe { ec_buf_shard_count(self.h) as usize }
    }
    /// The cells `Debug` shows — all of up to ten, else the first and the last five (`Elided`, src/lib.rs:166-194):
    /// one small D2H copy or ten single-cell reads, never the raster.
    fn ends(&self) -> (Vec<T>, Vec<T>) {
        let n = self.len();
        if n <= 10 {
            return (self.to_vec(), Vec::new());
        }
        let cell = |i: usize| -> T {
            let mut v = ZERO_VALUE;
            check(unsafe { ec_buf_get(self.h, i, &mut v) }).unwrap();
            let mut out = T::zero();
            // the payload is the cell's little-endian bytes in the low bytes of `bits`
            unsafe { ptr::copy_nonoverlapping(v.bits.to_le_bytes().as_ptr(), (&mut out as *mut T).cast::<u8>(), std::mem::size_of::<T>()) };
            out
        };
        ((0..5).map(cell).collect(), (n - 5..n).map(cell).collect())
    }
}
impl<T: CellEncoding> From<Vec<T>> for DeviceVec<T> {
    fn from(data: Vec<T>) -> Self {
        // one H2D copy. `data` is owned, so the copy could also be left in flight with ec_buf_from_host_async and the
        // Vec parked next to the handle until ec_buf_wait.
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_from_host(T::cell_type() as u8, data.as_ptr().cast(), data.len(), &mut h) }).unwrap();
        check(unsafe { ec_synchronize() }).unwrap();
        unsafe { Self::from_raw(h) }
    }
}
impl<T: CellEncoding> From<&DeviceVec<T>> for Vec<T> {
    fn from(v: &DeviceVec<T>) -> Self {
        v.to_vec()
    }
}
impl<T: CellEncoding> Drop for DeviceVec<T> {
    fn drop(&mut self) {
        unsafe { ec_buf_free(self.h) }
    }
}
impl<T: CellEncoding> Clone for DeviceVec<T> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_clone(self.h, &mut h) }).unwrap();
        unsafe { Self::from_raw(h) }
    }
}

// ---- CellBuffer — `pub enum CellBuffer { UInt8(Vec<u8>), … }` (src/buffer.rs:12-55) with device-resident variants ----
macro_rules! cb_enum {
    ( $(($id:ident, $p:ident)),*) => {
        /// An enum over buffers of [`CellEncoding`] types, resident in HBM.
        #[derive(Clone)]
        pub enum CellBuffer { $($id(DeviceVec<$p>)),* }
        impl CellBuffer {
            /// the C handle of whichever variant this is
            fn h(&self) -> *mut ec_buf {
                match self { $(CellBuffer::$id(v) => v.h),* }
            }
            /// wrap a handle the library returned: its cell type picks the variant
            fn wrap(h: *mut ec_buf) -> Self {
                match ct(unsafe { ec_buf_ctype(h) }) {
                    $(CellType::$id => CellBuffer::$id(unsafe { DeviceVec::<$p>::from_raw(h) })),*
                }
            }
        }
        $(impl From<DeviceVec<$p>> for CellBuffer {
            fn from(v: DeviceVec<$p>) -> Self { CellBuffer::$id(v) }
        })*
    }
}
with_ct!(cb_enum);

impl CellBuffer {
    pub fn new<T: CellEncoding>(data: Vec<T>) -> Self {
        Self::from_vec(data)
    }
    /// Extension: cells `[offset, offset + len)` as a buffer sharing this allocation (a row strip of a resident
    /// raster; `offset` on a 32-byte boundary). No copy; a `put` through either handle is seen by both.
    pub fn view(&self, offset: usize, len: usize) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_view(self.h(), offset, len, &mut h) }).unwrap();
        Self::wrap(h)
    }
    fn op2(code: i32, l: &CellBuffer, r: &CellBuffer) -> CellBuffer {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_binary(code, l.h(), r.h(), &mut h) }).unwrap(); // arithmetic never errors
        Self::wrap(h)
    }
    fn op_scalar(code: i32, l: &CellBuffer, r: CellValue) -> CellBuffer {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_scalar(code, l.h(), &to_ffi(r), &mut h) }).unwrap();
        Self::wrap(h)
    }
}

impl BufferOps for CellBuffer {
    fn from_vec<T: CellEncoding>(data: Vec<T>) -> Self {
        // src/buffer.rs:64-66
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_from_host(T::cell_type() as u8, data.as_ptr().cast(), data.len(), &mut h) }).unwrap();
        check(unsafe { ec_synchronize() }).unwrap();
        Self::wrap(h)
    }
    fn with_defaults(len: usize, c: CellType) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_with_defaults(len, c as u8, &mut h) }).unwrap();
        Self::wrap(h)
    }
    fn fill(len: usize, value: CellValue) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_fill(len, &to_ffi(value), &mut h) }).unwrap();
        Self::wrap(h)
    }
    fn fill_via<T: CellEncoding, F: Fn(usize) -> T>(len: usize, f: F) -> Self {
        Self::from_vec((0..len).map(f).collect()) // the closure is host code in the reference too
    }
    fn len(&self) -> usize {
        unsafe { ec_buf_len(self.h()) }
    }
    fn cell_type(&self) -> CellType {
        ct(unsafe { ec_buf_ctype(self.h()) })
    }
    fn get(&self, index: usize) -> CellValue {
        let mut v = ZERO_VALUE;
        check(unsafe { ec_buf_get(self.h(), index, &mut v) }).unwrap();
        from_ffi(v)
    }
    fn put(&mut self, index: usize, value: CellValue) -> Result<()> {
        check(unsafe { ec_buf_put(self.h(), index, &to_ffi(value)) })
    }
    fn convert(&self, c: CellType) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { ec_buf_convert(self.h(), c as u8, &mut h) })?;
        Ok(Self::wrap(h))
    }
    fn min_max(&self) -> (CellValue, CellValue) {
        let (mut a, mut b) = (ZERO_VALUE, ZERO_VALUE);
        check(unsafe { ec_buf_min_max(self.h(), ptr::null(), &mut a, &mut b) }).unwrap();
        (from_ffi(a), from_ffi(b))
    }
    fn to_vec<T: CellEncoding>(self) -> Result<Vec<T>> {
        // src/buffer.rs:175-185: convert on the device, one D2H copy
        let r = self.convert(T::cell_type())?;
        let n = if r.cell_type() == T::cell_type() { r.len() } else { 0 }; // an empty non-identity convert is UInt8([])
        let mut out = Vec::<T>::with_capacity(n);
        check(unsafe { ec_buf_to_host(r.h(), out.as_mut_ptr().cast(), n * std::mem::size_of::<T>()) })?;
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
        // src/buffer.rs:278-305 yields host CellValues: ONE D2H copy of the typed cells, wrapped on the host
        macro_rules! download {
            ($(($id:ident, $p:ident)),*) => {
                match self { $(CellBuffer::$id(v) => v.to_vec().into_iter().map(CellValue::$id).collect::<Vec<_>>(),)* }
            };
        }
        with_ct!(download).into_iter()
    }
}
impl<C: CellEncoding> Extend<C> for CellBuffer {
    fn extend<I: IntoIterator<Item = C>>(&mut self, iter: I) {
        // src/buffer.rs:205-221: value-checked `to_<p>().unwrap()` runs on the device; EC_NARROWING is that unwrap's panic
        let v: Vec<C> = iter.into_iter().collect();
        check(unsafe { ec_buf_extend_host(self.h(), C::cell_type() as u8, v.as_ptr().cast(), v.len()) }).expect("Extend: value out of range");
    }
}

macro_rules! cb_bin_op {
    // src/buffer.rs:321-358
    ($trt:ident, $mth:ident, $code:expr) => {
        impl $trt for &CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: Self) -> CellBuffer {
                CellBuffer::op2($code, self, rhs)
            }
        }
        impl $trt for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: Self) -> CellBuffer {
                CellBuffer::op2($code, &self, &rhs)
            }
        }
        impl $trt<&CellBuffer> for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: &CellBuffer) -> CellBuffer {
                CellBuffer::op2($code, &self, rhs)
            }
        }
        impl<R: Into<CellValue>> $trt<R> for CellBuffer {
            type Output = CellBuffer;
            fn $mth(self, rhs: R) -> CellBuffer {
                CellBuffer::op_scalar($code, &self, rhs.into())
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
        check(unsafe { ec_buf_neg(self.h(), &mut h) }).unwrap();
        CellBuffer::wrap(h)
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
        check(unsafe { ec_buf_cmp(self.h(), o.h(), &mut r) }).unwrap();
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
impl Debug for CellBuffer {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        // src/buffer.rs:188-203: `{type}CellBuffer(a, b, …)` with more than ten cells elided to five at either end
        use crate::Elided;
        let basename = self.cell_type().to_string();
        macro_rules! render {
            ( $(($id:ident, $_p:ident)),*) => {{
                f.write_fmt(format_args!("{basename}CellBuffer("))?;
                match self {
                    $(CellBuffer::$id(b) => {
                        let (head, tail) = b.ends();
                        if !head.is_empty() {
                            f.write_fmt(format_args!("{:?}", Elided(&head)))?;
                        }
                        if !tail.is_empty() {
                            f.write_str(", ... ")?;
                            f.write_fmt(format_args!("{:?}", Elided(&tail)))?;
                        }
                    })*
                };
                f.write_str(")")
            }}
        }
        with_ct!(render)
    }
}

// serde: the derive on `enum CellBuffer { UInt8(Vec<u8>), … }` (src/buffer.rs:51) is an externally tagged enum of
// sequences. The same derive on a host-side twin with the same name and variants emits exactly that format.
#[cfg(feature = "serde")]
mod wire {
    use super::*;
    macro_rules! wire_enum {
        ( $(($id:ident, $p:ident)),*) => {
            #[derive(Serialize, Deserialize)]
            #[serde(rename = "CellBuffer")]
            pub enum CellBufferWire { $($id(Vec<$p>)),* }
            impl From<&CellBuffer> for CellBufferWire {
                fn from(b: &CellBuffer) -> Self {
                    match b { $(CellBuffer::$id(v) => CellBufferWire::$id(v.to_vec())),* }
                }
            }
            impl From<CellBufferWire> for CellBuffer {
                fn from(w: CellBufferWire) -> Self {
                    match w { $(CellBufferWire::$id(v) => CellBuffer::$id(v.into())),* }
                }
            }
        }
    }
    with_ct!(wire_enum);
    #[derive(Serialize, Deserialize)]
    #[serde(rename = "Mask")]
    pub struct MaskWire(pub Vec<bool>); // src/masked/mask.rs:11-12
    #[derive(Serialize, Deserialize)]
    #[serde(rename = "MaskedCellBuffer")]
    pub struct MaskedWire(pub CellBufferWire, pub MaskWire); // src/masked/masked_buffer.rs:40-41
}
#[cfg(feature = "serde")]
impl Serialize for CellBuffer {
    fn serialize<S: Serializer>(&self, s: S) -> std::result::Result<S::Ok, S::Error> {
        wire::CellBufferWire::from(self).serialize(s)
    }
}
#[cfg(feature = "serde")]
impl<'de> Deserialize<'de> for CellBuffer {
    fn deserialize<D: Deserializer<'de>>(d: D) -> std::result::Result<Self, D::Error> {
        wire::CellBufferWire::deserialize(d).map(CellBuffer::from)
    }
}

// ---- Mask — was `Mask(Vec<bool>)` (src/masked/mask.rs:12): packed bits on the device ------------------------
/// Validity bits packed in HBM (one eighth of a byte per cell instead of one byte).
///
/// `Index` / `IndexMut` / `iter_mut` hand out `&bool` / `&mut bool` (src/masked/mask.rs:62-64, 89-101), which cannot
/// point into packed device words: they go through a host mirror (`Vec<bool>`), downloaded on first use. Mutable
/// access marks the mirror dirty; every device use first writes a dirty mirror back (one upload + pack). `&mut bool`
/// needs `&mut self`, so no device use can run while such a reference lives.
pub struct Mask {
    state: Mutex<MaskState>,
}
struct MaskState {
    h: *mut ec_mask,
    mirror: Option<Vec<bool>>, // host copy; authoritative while `dirty`
    dirty: bool,
}
unsafe impl Send for Mask {}
unsafe impl Sync for Mask {}
impl Drop for Mask {
    fn drop(&mut self) {
        let st = self.state.get_mut().unwrap();
        unsafe { ec_mask_free(st.h) }
    }
}
impl Mask {
    fn own(h: *mut ec_mask) -> Self {
        Mask { state: Mutex::new(MaskState { h, mirror: None, dirty: false }) }
    }
    /// The device handle, with any pending host edits written back first.
    fn h(&self) -> *mut ec_mask {
        let mut st = self.state.lock().unwrap();
        if st.dirty {
            let bools = st.mirror.as_ref().unwrap();
            let mut fresh = ptr::null_mut();
            check(unsafe { ec_mask_from_bools(bools.as_ptr().cast(), bools.len(), &mut fresh) }).unwrap(); // bool is one byte, 0/1
            check(unsafe { ec_synchronize() }).unwrap();
            unsafe { ec_mask_free(st.h) };
            st.h = fresh;
            st.dirty = false;
        }
        st.h
    }
    /// The host mirror, downloaded once. The returned pointer stays valid until the mirror is dropped, which only
    /// `&mut self` methods do.
    fn mirror_ptr(&self) -> *mut Vec<bool> {
        let mut st = self.state.lock().unwrap();
        if st.mirror.is_none() {
            let n = unsafe { ec_mask_len(st.h) };
            let mut out = vec![false; n];
            check(unsafe { ec_mask_to_bools(st.h, out.as_mut_ptr().cast(), n) }).unwrap();
            st.mirror = Some(out);
        }
        st.mirror.as_mut().unwrap() as *mut Vec<bool>
    }
    /// Forget the mirror after the device copy changed under it.
    fn drop_mirror(&mut self) {
        let st = self.state.get_mut().unwrap();
        debug_assert!(!st.dirty);
        st.mirror = None;
    }
    pub fn new(values: Vec<bool>) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_from_bools(values.as_ptr().cast(), values.len(), &mut h) }).unwrap();
        check(unsafe { ec_synchronize() }).unwrap();
        Self::own(h)
    }
    pub fn fill(len: usize, value: bool) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_fill(len, value as i32, &mut h) }).unwrap();
        Self::own(h)
    }
    pub fn fill_via<F: Fn(usize) -> bool>(len: usize, f: F) -> Self {
        Self::new((0..len).map(f).collect())
    }
    /// Extension: the validity bits of cells `[offset, offset + len)` as a mask of its own (`offset` a multiple of 128).
    pub fn slice(&self, offset: usize, len: usize) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_slice(self.h(), offset, len, &mut h) }).unwrap();
        Self::own(h)
    }
    pub fn len(&self) -> usize {
        unsafe { ec_mask_len(self.h()) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn put(&mut self, index: usize, value: bool) {
        let h = self.h();
        check(unsafe { ec_mask_put(h, index, value as i32) }).unwrap();
        self.drop_mirror();
    }
    pub fn get(&self, index: usize) -> bool {
        let mut o = 0;
        check(unsafe { ec_mask_get(self.h(), index, &mut o) }).unwrap();
        o != 0
    }
    /// src/masked/mask.rs:62-64: a mutable iterator over the values, in sequence (host mirror, written back later).
    pub fn iter_mut(&mut self) -> impl Iterator<Item = &'_ mut bool> {
        let p = self.mirror_ptr();
        self.state.get_mut().unwrap().dirty = true;
        unsafe { (*p).iter_mut() }
    }
    pub fn all(&self, value: bool) -> bool {
        let mut o = 0;
        check(unsafe { ec_mask_all(self.h(), value as i32, &mut o) }).unwrap();
        o != 0
    }
    /// `(data, nodata)` — free when a kernel produced this mask (it counted as it wrote), one popcount pass otherwise.
    pub fn counts(&self) -> (usize, usize) {
        let (mut d, mut n) = (0usize, 0usize);
        check(unsafe { ec_mask_counts(self.h(), &mut d, &mut n) }).unwrap();
        (d, n)
    }
    pub fn to_vec(&self) -> Vec<bool> {
        unsafe { (*self.mirror_ptr()).clone() }
    }
}
impl Clone for Mask {
    fn clone(&self) -> Self {
        // shares the packed words on the device (refcounted; a later put / extend on either side copies first)
        let mut h = ptr::null_mut();
        check(unsafe { ec_mask_clone(self.h(), &mut h) }).unwrap();
        Self::own(h)
    }
}
impl Default for Mask {
    fn default() -> Self {
        Mask::fill(0, true) // #[derive(Default)] on Mask(Vec<bool>) (src/masked/mask.rs:10): the empty mask
    }
}
impl Debug for Mask {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        // src/masked/mask.rs:165-169, elided like a buffer: at most ten bits leave the device
        use crate::Elided;
        let n = self.len();
        let (head, tail): (Vec<bool>, Vec<bool>) =
            if n <= 10 { (self.to_vec(), Vec::new()) } else { ((0..5).map(|i| self.get(i)).collect(), (n - 5..n).map(|i| self.get(i)).collect()) };
        f.write_str("Mask(")?;
        if !head.is_empty() {
            f.write_fmt(format_args!("{:?}", Elided(&head)))?;
        }
        if !tail.is_empty() {
            f.write_str(", ... ")?;
            f.write_fmt(format_args!("{:?}", Elided(&tail)))?;
        }
        f.write_str(")")
    }
}
impl Extend<bool> for Mask {
    fn extend<T: IntoIterator<Item = bool>>(&mut self, iter: T) {
        // src/masked/mask.rs:83-87: appended on the device (unpack, append, repack)
        let v: Vec<bool> = iter.into_iter().collect();
        let h = self.h();
        check(unsafe { ec_mask_extend_host(h, v.as_ptr().cast(), v.len()) }).unwrap();
        self.drop_mirror();
    }
}
impl Index<usize> for Mask {
    type Output = bool;
    fn index(&self, index: usize) -> &bool {
        // src/masked/mask.rs:89-95. The mirror outlives `&self`: only `&mut self` methods replace it.
        unsafe { &(*self.mirror_ptr())[index] }
    }
}
impl IndexMut<usize> for Mask {
    fn index_mut(&mut self, index: usize) -> &mut bool {
        // src/masked/mask.rs:97-101
        let p = self.mirror_ptr();
        self.state.get_mut().unwrap().dirty = true;
        unsafe { &mut (*p)[index] }
    }
}
impl IntoIterator for Mask {
    type Item = bool;
    type IntoIter = std::vec::IntoIter<bool>;
    fn into_iter(self) -> Self::IntoIter {
        // src/masked/mask.rs:171-177
        self.to_vec().into_iter()
    }
}
macro_rules! mask_op {
    ($trt:ident, $mth:ident, $f:ident) => {
        impl $trt for &Mask {
            type Output = Mask;
            fn $mth(self, rhs: Self) -> Mask {
                let mut h = ptr::null_mut();
                check(unsafe { $f(self.h(), rhs.h(), &mut h) }).unwrap();
                Mask::own(h)
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
        check(unsafe { ec_mask_not(self.h(), &mut h) }).unwrap();
        Mask::own(h)
    }
}
impl Not for Mask {
    type Output = Mask;
    fn not(self) -> Mask {
        // src/masked/mask.rs:103-109
        Not::not(&self)
    }
}
impl Ord for Mask {
    fn cmp(&self, o: &Self) -> Ordering {
        // #[derive(Ord)] on Vec<bool> (src/masked/mask.rs:10): lexicographic, false < true, then length
        let mut r = 0;
        check(unsafe { ec_mask_cmp(self.h(), o.h(), &mut r) }).unwrap();
        ordering(r)
    }
}
impl PartialOrd for Mask {
    fn partial_cmp(&self, o: &Self) -> Option<Ordering> {
        Some(self.cmp(o))
    }
}
impl PartialEq for Mask {
    fn eq(&self, o: &Self) -> bool {
        self.cmp(o) == Ordering::Equal
    }
}
impl Eq for Mask {}
#[cfg(feature = "serde")]
impl Serialize for Mask {
    fn serialize<S: Serializer>(&self, s: S) -> std::result::Result<S::Ok, S::Error> {
        wire::MaskWire(self.to_vec()).serialize(s)
    }
}
#[cfg(feature = "serde")]
impl<'de> Deserialize<'de> for Mask {
    fn deserialize<D: Deserializer<'de>>(d: D) -> std::result::Result<Self, D::Error> {
        wire::MaskWire::deserialize(d).map(|w| Mask::new(w.0))
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
#[derive(Clone, PartialEq, PartialOrd)]
pub struct MaskedCellBuffer(CellBuffer, Mask);
impl MaskedCellBuffer {
    pub fn new(buffer: CellBuffer, mask: Mask) -> Self {
        assert_eq!(buffer.len(), mask.len(), "Mask and buffer must have the same length.");
        Self(buffer, mask)
    }
    /// src/masked/masked_buffer.rs:62-71 — the sentinel compare runs on the device and emits packed words. Takes what
    /// the reference takes (`Vec<T>`, one upload) or cells that are already resident (`DeviceVec<T>`, which is what
    /// `match buf { CellBuffer::$id(v) => … }` in the GDAL adapter now yields, src/gdal/rasterband.rs:118).
    pub fn from_vec_with_nodata<T: CellEncoding, V: Into<DeviceVec<T>>>(data: V, nodata: NoData<T>) -> Self
    where
        CellBuffer: From<DeviceVec<T>>,
    {
        let buf: CellBuffer = data.into().into();
        let (kind, v) = nodata_ffi(nodata);
        let mut m = ptr::null_mut();
        check(unsafe { ec_mask_from_nodata(buf.h(), kind, &v, &mut m) }).unwrap();
        Self(buf, Mask::own(m))
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
        check(unsafe { ec_buf_fill_nodata(self.0.h(), self.1.h(), T::cell_type() as u8, kind, &v, &mut h) })?;
        CellBuffer::wrap(h).to_vec::<T>()
    }
    /// src/masked/masked_buffer.rs:208-217
    pub fn min_max(&self) -> (CellValue, CellValue) {
        let (mut a, mut b) = (ZERO_VALUE, ZERO_VALUE);
        check(unsafe { ec_buf_min_max(self.0.h(), self.1.h(), &mut a, &mut b) }).unwrap();
        (from_ffi(a), from_ffi(b))
    }
    /// Extension (not in the crate): count / min / max / mean / population stddev of the valid cells.
    pub fn statistics(&self) -> Statistics {
        let mut s = ec_statistics { count: 0, min: ZERO_VALUE, max: ZERO_VALUE, mean: 0.0, stddev: 0.0 };
        check(unsafe { ec_buf_statistics(self.0.h(), self.1.h(), &mut s) }).unwrap();
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
        Ok(Self(self.0.convert(cell_type)?, self.1.clone())) // the clone shares the mask words
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
impl<'a> From<&'a MaskedCellBuffer> for (&'a CellBuffer, &'a Mask) {
    fn from(value: &'a MaskedCellBuffer) -> Self {
        (&value.0, &value.1) // src/masked/masked_buffer.rs:243-247
    }
}
impl Debug for MaskedCellBuffer {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        // src/masked/masked_buffer.rs:227-235
        let basename = self.cell_type().to_string();
        f.debug_tuple(&format!("{basename}MaskedCellBuffer")).field(self.buffer()).field(self.mask()).finish()
    }
}
impl<'buf> IntoIterator for &'buf MaskedCellBuffer {
    type Item = (CellValue, bool);
    type IntoIter = std::iter::Zip<std::vec::IntoIter<CellValue>, std::vec::IntoIter<bool>>;
    fn into_iter(self) -> Self::IntoIter {
        // src/masked/masked_buffer.rs:289-317 yields host pairs: one D2H copy of the cells, one of the unpacked mask
        (&self.0).into_iter().zip(self.1.to_vec())
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
        self.1.extend(mask);
    }
}
impl Neg for &MaskedCellBuffer {
    type Output = MaskedCellBuffer;
    fn neg(self) -> MaskedCellBuffer {
        // src/masked/masked_buffer.rs:372-383: cells negated (also under a false mask), mask shared
        MaskedCellBuffer(-&self.0, self.1.clone())
    }
}
impl Neg for MaskedCellBuffer {
    type Output = MaskedCellBuffer;
    fn neg(self) -> MaskedCellBuffer {
        MaskedCellBuffer(-self.0, self.1)
    }
}
#[cfg(feature = "serde")]
impl Serialize for MaskedCellBuffer {
    fn serialize<S: Serializer>(&self, s: S) -> std::result::Result<S::Ok, S::Error> {
        wire::MaskedWire(wire::CellBufferWire::from(&self.0), wire::MaskWire(self.1.to_vec())).serialize(s)
    }
}
#[cfg(feature = "serde")]
impl<'de> Deserialize<'de> for MaskedCellBuffer {
    fn deserialize<D: Deserializer<'de>>(d: D) -> std::result::Result<Self, D::Error> {
        wire::MaskedWire::deserialize(d).map(|w| MaskedCellBuffer::new(w.0.into(), Mask::new((w.1).0)))
    }
}
/// Extension: chunked ingest — what `read_cells` / `read_cells_masked` (src/gdal/rasterband.rs:81-126) become when the
/// band is read block by block instead of into one `Vec`. The reader fills pinned staging slices handed out one at a
/// time; the upload of chunk k, the read of chunk k + 1 and the NoData compare of chunk k − 1 overlap, and on a
/// multi-GPU library every chunk goes straight to the GPU that owns its row strip.
/// ```ignore
/// let mut ingest = Ingest::<u16>::masked(len, NoData::new(0));
/// while let Some(chunk) = ingest.next_chunk() {
///     let n = band.read_block_into(chunk)?;   // fills chunk[..n]
///     ingest.submit(n);
/// }
/// let band: MaskedCellBuffer = ingest.finish_masked();
/// ```
pub struct Ingest<T: CellEncoding> {
    g: *mut ec_ingest,
    masked: bool,
    _cells: PhantomData<T>,
}
impl<T: CellEncoding> Ingest<T> {
    /// `read_cells`: cells only.
    pub fn new(len: usize) -> Self {
        let mut g = ptr::null_mut();
        check(unsafe { ec_ingest_begin(T::cell_type() as u8, len, 0, ptr::null(), 0, 0, &mut g) }).unwrap();
        Self { g, masked: false, _cells: PhantomData }
    }
    /// `read_cells_masked`: the validity mask is built from `nodata` chunk by chunk (`NoData::None`: all valid).
    pub fn masked(len: usize, nodata: NoData<T>) -> Self {
        let (kind, v) = nodata_ffi(nodata);
        let mut g = ptr::null_mut();
        check(unsafe { ec_ingest_begin(T::cell_type() as u8, len, kind, &v, 1, 0, &mut g) }).unwrap();
        Self { g, masked: true, _cells: PhantomData }
    }
    /// The next staging slice to fill, `None` once every cell has been submitted. Blocks while the upload that last
    /// used this slice is still reading it.
    pub fn next_chunk(&mut self) -> Option<&mut [T]> {
        let (mut p, mut cap) = (ptr::null_mut(), 0usize);
        check(unsafe { ec_ingest_next_buffer(self.g, &mut p, &mut cap) }).unwrap();
        if p.is_null() {
            None
        } else {
            Some(unsafe { std::slice::from_raw_parts_mut(p.cast::<T>(), cap) })
        }
    }
    /// The first `n_cells` of the slice handed out last are valid (a short chunk must be a multiple of 128 cells unless
    /// it is the last one).
    pub fn submit(&mut self, n_cells: usize) {
        check(unsafe { ec_ingest_submit(self.g, n_cells) }).unwrap();
    }
    pub fn finish(mut self) -> CellBuffer {
        let mut b = ptr::null_mut();
        check(unsafe { ec_ingest_finish(self.g, &mut b, ptr::null_mut()) }).unwrap();
        self.g = ptr::null_mut();
        CellBuffer::wrap(b)
    }
    pub fn finish_masked(mut self) -> MaskedCellBuffer {
        assert!(self.masked, "Ingest::new reads cells only; use Ingest::masked");
        let (mut b, mut m) = (ptr::null_mut(), ptr::null_mut());
        check(unsafe { ec_ingest_finish(self.g, &mut b, &mut m) }).unwrap();
        self.g = ptr::null_mut();
        MaskedCellBuffer(CellBuffer::wrap(b), Mask::own(m))
    }
}
impl<T: CellEncoding> Drop for Ingest<T> {
    fn drop(&mut self) {
        if !self.g.is_null() {
            unsafe { ec_ingest_abort(self.g) } // abandoned half way: frees what was uploaded so far
        }
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
    // src/masked/masked_buffer.rs:323-370: data on all cells, mask = lmask & rmask, one fused launch that also
    // counts the valid cells of the result
    ($trt:ident, $mth:ident, $code:expr) => {
        impl $trt for &MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: Self) -> MaskedCellBuffer {
                let (mut b, mut m) = (ptr::null_mut(), ptr::null_mut());
                check(unsafe { ec_masked_binary($code, self.0.h(), self.1.h(), rhs.0.h(), rhs.1.h(), &mut b, &mut m) }).unwrap();
                MaskedCellBuffer(CellBuffer::wrap(b), Mask::own(m))
            }
        }
        impl $trt for MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: Self) -> MaskedCellBuffer {
                $trt::$mth(&self, &rhs)
            }
        }
        impl $trt<&MaskedCellBuffer> for MaskedCellBuffer {
            type Output = MaskedCellBuffer;
            fn $mth(self, rhs: &MaskedCellBuffer) -> MaskedCellBuffer {
                $trt::$mth(&self, rhs) // :345-351
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
