//! Drop-in facade: the public API of `rcn` (rcn/src/rcn.rs, rcn/src/utils/kernel.rs) with the compute in
//! librcn_cuda.so. Same type names, argument meaning and panic behaviour as the CPU crate, so
//! `use rcn_cuda::{rcn::RCN, utils::kernel::*}` replaces `use rcn::{...}` in main.rs / benches / backend.
//! NOT compiled in this repository's environment (no rustc); kept as a thin, mechanical mirror of the C header.
pub mod ffi;

use nalgebra::{DMatrix, DVector};
use std::ffi::CStr;
use std::ptr;

fn check(rc: i32) {
    if rc != ffi::RCN_OK {
        // contract violations panic exactly like the CPU crate (kernel.rs:127,133,200,247,284)
        let msg = unsafe { CStr::from_ptr(ffi::rcn_cuda_last_error()) }.to_string_lossy().into_owned();
        panic!("{msg}");
    }
}

pub mod utils {
    pub mod kernel {
        use super::super::{check, ffi};
        use nalgebra::DMatrix;
        use std::ptr;

        #[derive(Clone, Copy)]
        pub enum SeparableOperator { Top, Bottom, Left, Right }      // kernel.rs:16-21
        #[derive(Clone, Copy)]
        pub enum Padding { None, Same }                              // kernel.rs:25-28
        #[derive(Clone, Copy)]
        pub enum Pooling { Average, Max }                            // kernel.rs:32-35

        /// kernel.rs:61-100 -- implemented for DMatrix<f64> (the type the model uses).
        pub trait Convolve2D {
            fn convolve_2d(&self, kernel: &DMatrix<f64>, padding: &Padding) -> DMatrix<f64>;
            fn convolve_2d_separated(&self, op: SeparableOperator, padding: &Padding) -> DMatrix<f64>;
            fn relu(&self) -> DMatrix<f64>;
        }
        /// kernel.rs:219-236
        pub trait Pool2D {
            fn pool_2d(&self, padding: &Padding, pooling: &Pooling) -> DMatrix<f64>;
        }

        impl Convolve2D for DMatrix<f64> {
            fn convolve_2d(&self, kernel: &DMatrix<f64>, padding: &Padding) -> DMatrix<f64> {
                let (h, w) = self.shape();
                let (kh, kw) = kernel.shape();
                let (oh, ow) = match padding { Padding::Same => (h, w), Padding::None => (h + 1 - kh.min(h + 1), w + 1 - kw.min(w + 1)) };
                let mut out = DMatrix::<f64>::zeros(oh, ow);
                check(unsafe { ffi::rcn_cuda_convolve_2d(0, ptr::null_mut(), self.as_ptr(), h, w, kernel.as_ptr(), kh, kw,
                                                         *padding as i32, out.as_mut_ptr()) });
                out
            }
            fn convolve_2d_separated(&self, op: SeparableOperator, padding: &Padding) -> DMatrix<f64> {
                let (h, w) = self.shape();
                let (oh, ow) = match padding { Padding::Same => (h, w), Padding::None => (h.saturating_sub(2), w.saturating_sub(2)) };
                let mut out = DMatrix::<f64>::zeros(oh, ow);
                check(unsafe { ffi::rcn_cuda_convolve_2d_separated(0, ptr::null_mut(), self.as_ptr(), h, w, op as i32,
                                                                   *padding as i32, out.as_mut_ptr()) });
                out
            }
            fn relu(&self) -> DMatrix<f64> {
                let mut out = DMatrix::<f64>::zeros(self.nrows(), self.ncols());
                check(unsafe { ffi::rcn_cuda_relu(0, ptr::null_mut(), self.as_ptr(), self.len(), out.as_mut_ptr()) });
                out
            }
        }
        impl Pool2D for DMatrix<f64> {
            fn pool_2d(&self, padding: &Padding, pooling: &Pooling) -> DMatrix<f64> {
                let (h, w) = self.shape();
                let (oh, ow) = match padding { Padding::Same => ((h + 1) / 2, (w + 1) / 2), Padding::None => (h / 2, w / 2) };
                let mut out = DMatrix::<f64>::zeros(oh, ow);
                check(unsafe { ffi::rcn_cuda_pool_2d(0, ptr::null_mut(), self.as_ptr(), h, w, *padding as i32, *pooling as i32,
                                                     out.as_mut_ptr(), ptr::null_mut()) });
                out
            }
        }
    }
}

pub mod rcn {
    use super::utils::kernel::{Padding, Pooling};
    use super::{check, ffi, DMatrix, DVector};
    use image::{io::Reader as ImageReader, ImageError};
    use std::ptr;

    /// rcn.rs:35-38
    pub enum RCNLayer { Convolve2D(Padding), Pool2D(Pooling) }
    /// rcn.rs:28-31
    pub struct Weights(pub DMatrix<f64>);
    pub struct Bias(pub DVector<f64>);

    /// rcn.rs:15-25 -- the parameters live on the GPU behind the handle.
    pub struct RCN<'a> {
        handle: ffi::rcn_cuda_handle,
        classes: usize,
        training_path: &'a str,
        testing_path: &'a str,
    }

    impl<'a> RCN<'a> {
        /// rcn.rs:58-64
        pub fn new(classes: usize, convpool_cfg: Vec<RCNLayer>, feedforward_cfg: Vec<usize>, training_path: &'a str,
                   testing_path: &'a str) -> Self {
            let codes: Vec<i32> = convpool_cfg.iter().map(|l| match l {
                RCNLayer::Convolve2D(p) => *p as i32,          // RCN_LAYER_CONV_NONE / _SAME
                RCNLayer::Pool2D(p) => 2 + *p as i32,          // RCN_LAYER_POOL_AVERAGE / _MAX
            }).collect();
            let mut handle = ptr::null_mut();
            check(unsafe { ffi::rcn_cuda_create(classes, codes.as_ptr(), codes.len(), feedforward_cfg.as_ptr(),
                                                feedforward_cfg.len(), 0, &mut handle) });
            RCN { handle, classes, training_path, testing_path }
        }

        /// rcn.rs:82-98: decode on the host, everything else on the GPU.
        pub fn classify(&self, img_path: &str) -> Result<usize, Box<dyn std::error::Error>> {
            let img = ImageReader::open(img_path)?.decode()?.grayscale().into_luma8();
            let (w, h) = img.dimensions();
            let mut label = 0i64;
            check(unsafe { ffi::rcn_cuda_classify(self.handle, img.as_raw().as_ptr() as *const _, ffi::RCN_PIXELS_U8_ROWMAJOR,
                                                  1, h as usize, w as usize, &mut label) });
            Ok(label as usize)
        }

        /// rcn.rs:126-133: load_data (host: directory walk + PNG decode into one u8 buffer per set), then the epoch loop
        /// of rcn.rs:144-165 issuing rcn_cuda_train_batch_images per chunk and rcn_cuda_evaluate per epoch.
        pub fn train(&mut self, batch_size: usize, epochs: usize, eta: f64, training_class_size_limit: usize,
                     testing_class_size_limit: usize) -> Result<(), ImageError> {
            let _ = (batch_size, epochs, eta, training_class_size_limit, testing_class_size_limit,
                     self.training_path, self.testing_path, self.classes);
            unimplemented!("host-side data loading: see mercer_research_b200/data.py + rcn.py::train for the reference order of calls")
        }

        /// Weights / Bias accessors used by the serde impls (serialization.rs:11-151): column-major, zero-copy layout.
        pub fn layer_weights(&self, layer: usize) -> Weights {
            let (mut r, mut c) = (0usize, 0usize);
            check(unsafe { ffi::rcn_cuda_layer_shape(self.handle, layer, &mut r, &mut c) });
            let mut m = DMatrix::<f64>::zeros(r, c);
            check(unsafe { ffi::rcn_cuda_get_weights(self.handle, layer, m.as_mut_ptr()) });
            Weights(m)
        }
    }

    impl<'a> Drop for RCN<'a> {
        fn drop(&mut self) { unsafe { ffi::rcn_cuda_destroy(self.handle); } }
    }
    // &self methods only read device state and serialise on the model's stream; mutation needs &mut self, so the
    // Sync contract of the CPU crate (rayon workers share &RCN, rcn.rs:190-191) is preserved.
    unsafe impl<'a> Send for RCN<'a> {}
}
