// Microbenchmark: shared-memory pipe cost of the access patterns the tiled kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_micro smem_micro.cu && ./smem_micro
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void k(float* out, int warps_per_block) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = i * 0.001f;
    __syncthreads();
    float acc = 0.f;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (tid >> 5) * 64;
    const unsigned a64 = base + lane * 8;          // conflict-free 64-bit
    const unsigned a128 = base + lane * 16;        // conflict-free 128-bit
    const unsigned au = base;                      // uniform address
    const unsigned a32 = base + lane * 4;
    int src = (lane + 1) & 31;
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        float x, y, z, w;
        if (MODE == 0) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a64 + (it & 15) * 256)); acc += x + y; }
        if (MODE == 1) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(au + (it & 15) * 16)); acc += x + y + z + w; }
        if (MODE == 2) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(au + (it & 15) * 4)); acc += x; }
        if (MODE == 3) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(au + (it & 15) * 8)); acc += x + y; }
        if (MODE == 4) { acc += __shfl_sync(0xffffffffu, acc, src); }
        if (MODE == 5) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a128 + (it & 7) * 512)); acc += x + y + z + w; }
        if (MODE == 6) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a64 + (it & 15) * 256), "f"(acc), "f"(acc) : "memory"); acc += 1.f; }
        if (MODE == 7) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a32 + (it & 15) * 128)); acc += x; }
        if (MODE == 8) {   // LDS.64 data + uniform LDS.128 mix (interp-like: 6 data loads + 2 uniform 128 + 2 scalar)
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a64 + (it & 15) * 256)); acc += x + y;
            if ((it % 3) == 0) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(au + (it & 15) * 16)); acc += x + w; }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, int threads) {
    float* d; cudaMalloc(&d, 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * blocks_per_sm;
    k<MODE><<<grid, threads, 40000>>>(d, threads / 32);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads, 40000>>>(d, threads / 32);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * dev_clk * 1e3;
    double winstr_per_sm = (double)blocks_per_sm * (threads / 32) * ITERS * (MODE == 8 ? 1.3333 : 1.0);
    printf("%-34s blocks/SM %d thr %d : %.3f ms, %.2f cycles per warp-instruction per SM (@%d MHz nominal)\n", name,
           blocks_per_sm, threads, ms, cycles / winstr_per_sm, dev_clk / 1000);
    cudaFree(d);
}

int main() {
    for (int thr : {256, 1024}) {
        int b = thr == 256 ? 4 : 2;
        run<0>("LDS.64 conflict-free", b, thr);
        run<5>("LDS.128 conflict-free", b, thr);
        run<7>("LDS.32 conflict-free", b, thr);
        run<1>("LDS.128 uniform (broadcast)", b, thr);
        run<3>("LDS.64 uniform (broadcast)", b, thr);
        run<2>("LDS.32 uniform (broadcast)", b, thr);
        run<6>("STS.64 conflict-free", b, thr);
        run<4>("SHFL.IDX", b, thr);
        run<8>("LDS.64 + 1/3 uniform LDS.128", b, thr);
    }
    return 0;
}
