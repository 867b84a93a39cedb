// opctx.cuh -- shared plumbing of the model-free op-level entry points (rcn_cuda_convolve_2d, rcn_cuda_ext_*):
// device selection, host<->device staging of caller buffers through grow-only thread-local scratch.
#pragma once
#include "common.cuh"

namespace rcn {

// per-thread scratch for host-pointer callers (inputs a/b/c, outputs, auxiliary)
inline thread_local DevBuf tl_op_in, tl_op_k, tl_op_out, tl_op_aux, tl_op_in2, tl_op_out2, tl_op_ws;

struct OpCtx {
    cudaStream_t stream = nullptr;
    int enter(int device, void* cuda_stream) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(RCN_ERR_CUDA, "no CUDA device available; librcn_cuda has no CPU fallback");
        }
        if (device < 0 || device >= ndev) return fail(RCN_ERR_INVALID, "device %d out of range", device);
        RCN_CUDA_TRY(cudaSetDevice(device));
        stream = (cudaStream_t)cuda_stream;  // NULL = legacy default stream
        return RCN_OK;
    }
    int in(const void* src, size_t bytes, DevBuf& buf, const void** dev) {
        if (is_device_ptr(src)) { *dev = src; return RCN_OK; }
        RCN_TRY(buf.reserve(bytes));
        RCN_CUDA_TRY(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, stream));
        *dev = buf.p;
        return RCN_OK;
    }
    int out(void* dst, size_t bytes, DevBuf& buf, void** dev, bool* host) {
        if (is_device_ptr(dst)) { *dev = dst; *host = false; return RCN_OK; }
        RCN_TRY(buf.reserve(bytes));
        *dev = buf.p; *host = true;
        return RCN_OK;
    }
    int finish(void* dst, const void* dev, size_t bytes, bool host) {
        if (host) {
            RCN_CUDA_TRY(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, stream));
            RCN_CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return RCN_OK;
    }
};

}  // namespace rcn
