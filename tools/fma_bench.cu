// Microbenchmark: issue rate of FFMA (3-register form) vs FFMA2 (fma.rn.f32x2) on one SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fma_bench tools/fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a0, float b0) {
    float a[6], t[3];
    for (int i = 0; i < 6; ++i) a[i] = a0 + i + threadIdx.x;
    for (int i = 0; i < 3; ++i) t[i] = b0 + i;
    if (MODE == 0) {
        float acc[36];
        for (int i = 0; i < 36; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int s = 0; s < 6; ++s) {
                    acc[(i * 6 + s) * 2] = fmaf(a[s], t[i], acc[(i * 6 + s) * 2]);
                    acc[(i * 6 + s) * 2 + 1] = fmaf(a[s], -t[i], acc[(i * 6 + s) * 2 + 1]);
                }
        }
        float r = 0;
        for (int i = 0; i < 36; ++i) r += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {
        unsigned long long acc[18], aa[6], tt[3];
        for (int i = 0; i < 18; ++i) acc[i] = i;
        for (int i = 0; i < 6; ++i) asm("mov.b64 %0, {%1, %1};" : "=l"(aa[i]) : "f"(a[i]));
        for (int i = 0; i < 3; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(tt[i]) : "f"(t[i]), "f"(-t[i]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int s = 0; s < 6; ++s) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i * 6 + s]) : "l"(aa[s]), "l"(tt[i]));
        }
        unsigned long long r = 0;
        for (int i = 0; i < 18; ++i) r ^= acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
    }
}
int main() {
    float* d; cudaMalloc(&d, 1 << 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.f, 2.f); else k<1><<<148, warps * 32>>>(d, iters, 1.f, 2.f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 36.0 * iters * warps * 32 * 148;      // scalar FMAs
            printf("warps/SM %2d mode %s: %.3f ms, %.2f TFMA/s, %.1f FMA/clk/SM @1.965GHz\n", warps, mode ? "FFMA2" : "FFMA ", ms,
                   fma / ms / 1e9, fma / (ms * 1e-3) / 1.965e9 / 148);
        }
    }
    return 0;
}
