//! Device tensor and allocator for the Rust host: a few `libcudart` entry points bound by hand (no cudarc) and two RAII
//! owners -- `DeviceBuffer<T>` (HBM) and `PinnedBuffer<T>` (page-locked host memory the GPU can pull from, which is what
//! `rcn_cuda_train_epoch_host` wants for its streaming path). Every librcn_cuda entry point accepts host OR device
//! pointers (detected), so these types only exist to keep big, long-lived tensors (a data set's features, its labels)
//! resident in HBM instead of being staged on every call.
use std::marker::PhantomData;
use std::os::raw::{c_int, c_uint, c_void};
use std::{mem, ptr, slice};

const CUDA_MEMCPY_HOST_TO_DEVICE: c_int = 1;
const CUDA_MEMCPY_DEVICE_TO_HOST: c_int = 2;
const CUDA_HOST_ALLOC_PORTABLE: c_uint = 1;

#[link(name = "cudart")]
extern "C" {
    fn cudaSetDevice(device: c_int) -> c_int;
    fn cudaMalloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn cudaFree(ptr: *mut c_void) -> c_int;
    fn cudaHostAlloc(ptr: *mut *mut c_void, bytes: usize, flags: c_uint) -> c_int;
    fn cudaFreeHost(ptr: *mut c_void) -> c_int;
    fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
    fn cudaGetErrorString(code: c_int) -> *const std::os::raw::c_char;
}

fn cuda(rc: c_int, what: &str) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(cudaGetErrorString(rc)) }.to_string_lossy().into_owned();
        panic!("{what} failed: {msg}");   // there is no CPU fallback: a missing / failing device is fatal
    }
}

/// `n` elements of `T` in the HBM of `device`. Plain-old-data element types only (f64, i64, u8).
pub struct DeviceBuffer<T: Copy> {
    ptr: *mut T,
    len: usize,
    device: i32,
    _own: PhantomData<T>,
}

impl<T: Copy> DeviceBuffer<T> {
    pub fn new(device: i32, len: usize) -> Self {
        let mut p: *mut c_void = ptr::null_mut();
        cuda(unsafe { cudaSetDevice(device) }, "cudaSetDevice");
        cuda(unsafe { cudaMalloc(&mut p, len.max(1) * mem::size_of::<T>()) }, "cudaMalloc");
        DeviceBuffer { ptr: p as *mut T, len, device, _own: PhantomData }
    }
    pub fn from_slice(device: i32, host: &[T]) -> Self {
        let buf = Self::new(device, host.len());
        cuda(unsafe { cudaMemcpy(buf.ptr as *mut c_void, host.as_ptr() as *const c_void, mem::size_of_val(host),
                                 CUDA_MEMCPY_HOST_TO_DEVICE) }, "cudaMemcpy H2D");
        buf
    }
    pub fn to_vec(&self) -> Vec<T> {
        let mut v = Vec::<T>::with_capacity(self.len);
        cuda(unsafe { cudaSetDevice(self.device) }, "cudaSetDevice");
        cuda(unsafe { cudaMemcpy(v.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, self.len * mem::size_of::<T>(),
                                 CUDA_MEMCPY_DEVICE_TO_HOST) }, "cudaMemcpy D2H");
        unsafe { v.set_len(self.len) };
        v
    }
    pub fn as_ptr(&self) -> *const T { self.ptr }
    pub fn as_mut_ptr(&mut self) -> *mut T { self.ptr }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
}

impl<T: Copy> Drop for DeviceBuffer<T> {
    fn drop(&mut self) {
        unsafe { cudaSetDevice(self.device); cudaFree(self.ptr as *mut c_void); }
    }
}
unsafe impl<T: Copy + Send> Send for DeviceBuffer<T> {}

/// Page-locked host memory (portable across the box's GPUs), zero-initialised, usable as a slice.
pub struct PinnedBuffer<T: Copy> {
    ptr: *mut T,
    len: usize,
}

impl<T: Copy> PinnedBuffer<T> {
    pub fn new(len: usize) -> Self {
        let mut p: *mut c_void = ptr::null_mut();
        let bytes = len.max(1) * mem::size_of::<T>();
        cuda(unsafe { cudaHostAlloc(&mut p, bytes, CUDA_HOST_ALLOC_PORTABLE) }, "cudaHostAlloc");
        unsafe { ptr::write_bytes(p as *mut u8, 0, bytes) };
        PinnedBuffer { ptr: p as *mut T, len }
    }
    pub fn as_slice(&self) -> &[T] { unsafe { slice::from_raw_parts(self.ptr, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { slice::from_raw_parts_mut(self.ptr, self.len) } }
    pub fn as_ptr(&self) -> *const T { self.ptr }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
}

impl<T: Copy> Drop for PinnedBuffer<T> {
    fn drop(&mut self) { unsafe { cudaFreeHost(self.ptr as *mut c_void); } }
}
unsafe impl<T: Copy + Send> Send for PinnedBuffer<T> {}
unsafe impl<T: Copy + Sync> Sync for PinnedBuffer<T> {}
