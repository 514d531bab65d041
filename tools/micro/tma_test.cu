// Minimal check of the tensor-map plane copy used by k_interp_col: 3-D float32 map over a [K0][K1][2 K2] grid, box 1 x 9 x 20.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const CUtensorMap* map, float2* out, int x, int y, int z, int bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2048);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                         ::"r"(smem_u32(sm)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                         ::"r"(smem_u32(sm)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
    }
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    const float2* s = reinterpret_cast<const float2*>(sm);
    for (int i = threadIdx.x; i < 90; i += 32) out[i] = s[i];
}
int main(int argc, char** argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 3, b0 = argc > 2 ? atoi(argv[2]) : 20, b1 = argc > 3 ? atoi(argv[3]) : 9, x0 = argc > 4 ? atoi(argv[4]) : 10;
    const int K0 = 64, K1 = 64, K2 = 64;
    std::vector<float2> h((size_t)K0 * K1 * K2);
    for (int a = 0; a < K0; ++a) for (int b = 0; b < K1; ++b) for (int c = 0; c < K2; ++c) h[((size_t)a * K1 + b) * K2 + c] = make_float2(a * 10000 + b * 100 + c, -1.f);
    float2 *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 90 * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, f);
    CUtensorMap m;
    const cuuint64_t dims[3] = {2 * K2, K1, K0}; const cuuint64_t strides[2] = {(cuuint64_t)K2 * 8, (cuuint64_t)K1 * K2 * 8};
    const cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}; const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)f)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    printf("rank %d box %d x %d\n", rank, b0, b1);
    if (rank == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096); k<3><<<1, 32, 4096>>>(dm, o, x0, 4, 3, b0 * b1 * 4); }
    else { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096); k<2><<<1, 32, 4096>>>(dm, o, x0, 4, 3, b0 * b1 * 4); }
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    float2 ho[90]; cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("first row: %.0f %.0f .. %.0f ; second row first: %.0f (expect 30405 30406 .. 30414 ; 30505)\n", ho[0].x, ho[1].x, ho[9].x, ho[10].x);
    return 0;
}
