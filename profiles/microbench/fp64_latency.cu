// Microbenchmark: dependent-issue latency (cycles) of the f64 instructions kernel A's serial chains are made of, one
// warp on one SM: DADD, DFMA, DMMA.8x8x4 (accumulator chain), exp(), 1/(1+exp(-z)), IEEE division, LDS.64.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 256
template <int OP>
__global__ void chain(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001, c0 = 0, c1 = 0;
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < CHAIN; ++i) {
        if (OP == 0) x = __dadd_rn(x, y);
        if (OP == 1) x = fma(x, y, y);
        if (OP == 2) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(x), "d"(y));
        if (OP == 3) x = exp(-x * 1e-3);
        if (OP == 4) x = 1.0 / (1.0 + exp(-x));
        if (OP == 5) x = y / (x + 1.5);
        if (OP == 6) { x = sm[idx & 31]; idx = (int)x; }
        if (OP == 7) x = __dmul_rn(x, y);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + c0 + c1 + idx;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * sizeof(double));
    cudaMallocManaged(&cyc, 8 * sizeof(long long));
    const char* names[8] = {"DADD", "DFMA", "DMMA.8x8x4 (accumulator chain)", "exp()", "sigmoid 1/(1+exp(-z))", "IEEE f64 division", "LDS.64 (pointer chase)", "DMUL"};
    for (int rep = 0; rep < 2; ++rep) {
        chain<0><<<1, 32>>>(out, cyc, 1.0); chain<1><<<1, 32>>>(out, cyc, 1.0); chain<2><<<1, 32>>>(out, cyc, 1.0);
        chain<3><<<1, 32>>>(out, cyc, 1.0); chain<4><<<1, 32>>>(out, cyc, 1.0); chain<5><<<1, 32>>>(out, cyc, 1.0);
        chain<6><<<1, 32>>>(out, cyc, 0.0); chain<7><<<1, 32>>>(out, cyc, 1.0);
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < 8; ++i) printf("%-34s %8.1f cycles per dependent op\n", names[i], (double)cyc[i] / CHAIN);
    return 0;
}
