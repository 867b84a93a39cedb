// Microbenchmark: sustained throughput of DMMA.8x8x4 (mma.sync.m8n8k4.f64) vs DFMA on one B200.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_loop(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_loop(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms;
        dmma_loop<<<148, warps * 32>>>(out, 100);
        cudaEventRecord(e0); dmma_loop<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dmma = 148.0 * warps * iters * 8.0;
        printf("DMMA  warps/SM=%2d: %.3f ms  %.2f TFLOP/s  (%.1f SM-cycles per DMMA per SMSP at 1.9 GHz)\n", warps, ms,
               dmma * 512 / (ms * 1e-3) / 1e12, (ms * 1e-3 * 1.9e9) / (warps / 4.0 * iters * 8.0));
        dfma_loop<<<148, warps * 32>>>(out, 100);
        cudaEventRecord(e0); dfma_loop<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * warps * 32 * iters * 16.0 * 2;
        printf("DFMA  warps/SM=%2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / (ms * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
