//! Multi-GPU data-parallel trainer for the Rust host: the rayon + mutex gradient sum of
//! rcn.rs:190-205, spread over the GPUs of one box. One replica per GPU in THIS process, one host thread per replica;
//! each global minibatch is split into `world` contiguous shards, every rank runs features + forward + backward on its
//! shard, and the gradient sums are exchanged and applied by ONE kernel over NVLink peer memory
//! (`rcn_cuda_dp_connect_local`) -- no host-side collective on the step. Replicas stay bit-identical to each other; the
//! result differs from one GPU only by summation order.
use crate::cuda::PinnedBuffer;
use crate::rcn::RCN;
use crate::{check, ffi};
use std::os::raw::c_void;
use std::{ptr, thread};

pub struct DataParallel<'a> {
    replicas: Vec<RCN<'a>>,
}

impl<'a> DataParallel<'a> {
    /// `replicas[r]` lives on GPU r and already holds the same parameters and scale_set as every other replica
    /// (e.g. each deserialised from the same `rcn.bin`, or rank 0's parameters copied to the rest).
    pub fn new(replicas: Vec<RCN<'a>>) -> Self {
        let world = replicas.len() as i32;
        assert!(world >= 1, "need at least one replica");
        if world > 1 {
            for (rank, m) in replicas.iter().enumerate() {
                check(unsafe { ffi::rcn_cuda_dp_init(m.raw(), world, rank as i32, ptr::null_mut()) });
            }
            let group: Vec<ffi::rcn_cuda_handle> = replicas.iter().map(|m| m.raw()).collect();
            for m in &replicas {
                check(unsafe { ffi::rcn_cuda_dp_connect_local(m.raw(), group.as_ptr()) });
            }
        }
        DataParallel { replicas }
    }

    pub fn world(&self) -> usize { self.replicas.len() }
    pub fn replica(&self, rank: usize) -> &RCN<'a> { &self.replicas[rank] }

    /// One epoch of `for batch in training_set.chunks_exact(global_batch) { train_batch(batch, eta) }` (rcn.rs:147-149)
    /// over an already shuffled host data set of `n` images (h x w Luma8 bytes, page-locked). Rank r trains on samples
    /// `[k*global_batch + r*b, k*global_batch + (r+1)*b)` of chunk k with `b = global_batch / world`; its shard of every
    /// chunk is gathered into one contiguous page-locked buffer first so that the GPU can stream it.
    /// Returns the number of steps taken.
    pub fn train_epoch(&mut self, pixels: &[u8], labels: &[i64], n: usize, h: usize, w: usize, global_batch: usize,
                       eta: f64) -> usize {
        let world = self.replicas.len();
        assert!(global_batch % world == 0, "the global minibatch must divide evenly over the GPUs");
        let (b, img) = (global_batch / world, h * w);
        let steps = n / global_batch;                              // chunks_exact: the remainder is dropped
        if steps == 0 { return 0; }
        let shards: Vec<(PinnedBuffer<u8>, Vec<i64>)> = (0..world).map(|r| {
            let mut px = PinnedBuffer::<u8>::new(steps * b * img);
            let mut lb = vec![0i64; steps * b];
            for k in 0..steps {
                let src = k * global_batch + r * b;
                px.as_mut_slice()[k * b * img..(k + 1) * b * img].copy_from_slice(&pixels[src * img..(src + b) * img]);
                lb[k * b..(k + 1) * b].copy_from_slice(&labels[src..src + b]);
            }
            (px, lb)
        }).collect();
        // every rank must issue the same sequence of steps; the exchange kernels meet over NVLink, not on the host
        thread::scope(|scope| {
            for (m, (px, lb)) in self.replicas.iter().zip(shards.iter()) {
                let handle = m.raw() as usize;                     // raw pointers are not Send; the handle itself is
                scope.spawn(move || {
                    let mut done = 0usize;
                    check(unsafe { ffi::rcn_cuda_train_epoch_host(handle as ffi::rcn_cuda_handle, px.as_ptr() as *const c_void,
                                                                  ffi::RCN_PIXELS_U8_ROWMAJOR, lb.as_ptr(), steps * b, h, w, b,
                                                                  eta, global_batch, ptr::null_mut(), ptr::null_mut(),
                                                                  &mut done) });
                    assert_eq!(done, steps);
                });
            }
        });
        steps
    }
}

impl<'a> Drop for DataParallel<'a> {
    fn drop(&mut self) {
        if self.replicas.len() > 1 {
            for m in &self.replicas { unsafe { ffi::rcn_cuda_dp_shutdown(m.raw()); } }
        }
    }
}
