// Column-sweep gridding kernel for 3-D, J = 6: the register-resident replacement of pELL_spmvh_mCoil +
// atomic_add_float2 (src/re_subroutine.py:527-596, 275-287).
//
// The shared-memory tiled kernel (grid_tiled.cu) is bound by the shared-memory pipe: every one of the 216 neighbours
// of a sample is an LDS + STS.  Here the grid cells a warp works on live in REGISTERS:
//
//  * a second copy of the samples is sorted by (column, first plane): a column is a 4 x 5 cross-section (dims 1, 2)
//    of first-neighbour cells, swept along dim 0 in increasing plane order (key = column * K0 + first plane);
//  * the footprint of a sample lies inside the 9 x 10 box of its column (4 + 5 rows, 5 + 5 columns) and covers, in
//    dim 0, exactly one plane of every residue class mod 6.  Lane (g, c), g = 0..2, c = 0..9 (30 lanes), owns the
//    box cells (rows g, g + 3, g + 6; column c) of the six planes of the current window [p0, p0 + 6): plane p sits in
//    slot p mod 6, 18 complex accumulators per lane.  Samples arrive in plane order, so the window only slides
//    forward; the planes it leaves are flushed by ALL lanes at once (warp-uniform, three 8-byte REDs per lane that
//    cover three 80-byte row segments per instruction) and their slots are cleared;
//  * the inner loop is branch-free and every weight is a plain register operand: the record holds the dim-0 weights
//    already rotated to the slots, the dim-1 weights as a zero-padded table over the box rows and the dim-2 weights
//    zero-padded over the box columns, so a sample is 8 FMUL + 36 FFMA per lane on static registers;
//  * all weights are REAL: the grid this kernel produces is phase-modulated, G'[g] = G[g] * prod_d e^{i s_d g_d}
//    (s_d = gamma_d (N_d - 1) / 2), which turns the reference's complex min-max coefficients
//    u_j = c_j e^{i om N/2} e^{-i s (dk - j)} (helper.py:148-162, 606-618) into c_j times ONE phase per sample (folded
//    into the pre-gathered data).  The modulation is undone for free by the inverse FFT passes (fft256.cu) or by
//    k_demodulate below; neighbours that wrap around the periodic grid pick up the sign e^{i s K} = (-1)^(N-1);
//  * sample records (128 B, precomputed at plan time in sweep order) and the pre-gathered data arrive by
//    double-buffered TMA bulk copies; warps are persistent and fetch work items (segments of a column) from a
//    per-coil counter.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int CT1 = COL_T1;               // column cross-section, dim 1 (first-neighbour cells): 4
constexpr int CT2 = COL_T2;               // column cross-section, dim 2: 5
constexpr int CROWS = CT1 + 5;            // 9 box rows
constexpr int CCOLS = CT2 + 5;            // 10 box columns
constexpr int CNR = CROWS / 3;            // rows per lane: 3
constexpr int CRECW = COL_RECW;           // words per record: 32
constexpr int CCH = 16;                   // samples per chunk
constexpr int CWARPS = 4;                 // warps per CTA (each warp works on its own items)
constexpr int REC_BYTES = CCH * CRECW * 4;            // 2048
constexpr int YS_BYTES = CCH * 16;
constexpr int GWARP_BYTES = 4736;                     // 2 * REC + 2 * YS + mbar, rounded to 128
static_assert(CT1 == 4 && CT2 == 5 && CRECW == 32 && CROWS == 3 * CNR, "record layout below assumes 4 x 5 columns");
static_assert(2 * REC_BYTES + 2 * YS_BYTES + 16 <= GWARP_BYTES, "per-warp shared memory");

// record words: [c1g[0..11] | c0rot[0..5] | p0 | p0 mod 6 | w10[0..9] | 0 0]   (plan.cu k_col_records)

struct ColGeom {
    int K0, K1, K2, nq2;
    float sg0, sg1, sg2;        // e^{i s_d K_d} = (-1)^(N_d - 1): factor of neighbours that wrap
    long long Kprod;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void red_v2(float2* addr, float2 a) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a.x), "f"(a.y) : "memory");
}

// cell + off elements as ONE 64-bit multiply-add (the compiler would re-derive the pointer from the grid base)
__device__ __forceinline__ float2* cell_at(const float2* cell, int off) {
    float2* r;
    asm("mad.wide.s32 %0, %1, 8, %2;\n" : "=l"(r) : "r"(off), "l"(cell));
    return r;
}

// next work item of this warp (dynamic: one atomic per item, broadcast from lane 0)
__device__ __forceinline__ int next_item(int* counter, int lane) {
    int it = 0;
    if (lane == 0) it = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
}

// Pre-pass of the scatter, one kernel: ys[c][i] = (conj(P''_i) * y[perm[i], c], 0, 0) (16-byte slots so that any sample
// range is TMA-aligned), the grid zero-fill (the gather is latency-bound on the random reads of y and leaves the
// bandwidth to the stores) and the reset of the persistent kernel's work counters.
__global__ void k_gather_sorted_col(const float4* __restrict__ side, long long M, const float2* __restrict__ y,
                                    float4* __restrict__ ys, int nb, float4* __restrict__ grid4, long long n4,
                                    int* __restrict__ counters) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long T = gridDim.x * (long long)blockDim.x;
    if (i < nb) counters[i] = 0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long j = i; j < n4; j += T) grid4[j] = z;
    if (i >= M) return;
    const float4 h = __ldg(side + i);                 // P''.re, P''.im, original index
    const int m = __float_as_int(h.z);
    for (int c = 0; c < nb; ++c) {
        const float2 v = cmulc(make_float2(h.x, h.y), y[(long long)m * nb + c]);
        ys[(long long)c * M + i] = make_float4(v.x, v.y, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(CWARPS * 32, 5)
k_gridding_col(ColGeom g, const WorkItem* __restrict__ work, int n_work, int* __restrict__ counter,
               const float* __restrict__ rec, const float4* __restrict__ ys, long long M, float2* __restrict__ grid) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ws = smem_raw + warp * GWARP_BYTES;
    unsigned char* ybuf = ws + 2 * REC_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + 2 * REC_BYTES + 2 * YS_BYTES);
    const int c = blockIdx.y;
    float2* gc = grid + (long long)c * g.Kprod;
    const float4* ysc = ys + (long long)c * M;
    // lanes 30, 31 shadow lane 29's cells with the always-zero record word 30 as their column weight: their
    // accumulators stay exactly zero and their REDs add 0 to valid cells, so the flush needs no predicate
    const bool active = lane < 3 * CCOLS;
    const int lg = active ? lane / CCOLS : 2;           // row group: box rows lg, lg + 3, lg + 6
    const int lc = active ? lane - CCOLS * lg : CCOLS - 1;   // box column
    const int lw = active ? lc : CCOLS;                 // word 20 + lw of the record: this lane's column weight
    const int KK = g.K1 * g.K2;
    const unsigned sbit0 = g.sg0 < 0.f ? 0x80000000u : 0u, sbit1 = g.sg1 < 0.f ? 0x80000000u : 0u,
                   sbit2 = g.sg2 < 0.f ? 0x80000000u : 0u;
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    unsigned gk = 0;                                    // chunks consumed by this warp so far (mbarrier phases)

    for (int item = next_item(counter + c, lane); item < n_work; item = next_item(counter + c, lane)) {
        const WorkItem wi = work[item];
        const int q1 = wi.tile / g.nq2, q2 = wi.tile - q1 * g.nq2;
        // this lane's three cells inside plane 0, and the sign bits of their wrap factors
        float2* cell[CNR];
        unsigned rbit[CNR];
        {
            int col = q2 * CT2 + lc;
            unsigned cbit = 0u;
            if (col >= g.K2) { col -= g.K2; cbit = sbit2; }
#pragma unroll
            for (int i = 0; i < CNR; ++i) {
                int row = q1 * CT1 + lg + 3 * i;
                unsigned b = cbit;
                if (row >= g.K1) { row -= g.K1; b ^= sbit1; }
                cell[i] = gc + (row * g.K2 + col);
                rbit[i] = b;
            }
        }
        const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
        auto issue = [&](int k) {      // lane 0 only
            const int s = wi.begin + k * CCH;
            const int ns = min(CCH, wi.end - s);
            const unsigned b = (gk + k) & 1;
            mbar_expect(&mbar[b], (unsigned)(ns * (CRECW * 4 + 16)));
            tma_bulk(ws + b * REC_BYTES, rec + (long long)s * CRECW, (unsigned)(ns * CRECW * 4), &mbar[b]);
            tma_bulk(ybuf + b * YS_BYTES, ysc + s, (unsigned)(ns * 16), &mbar[b]);
        };
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) issue(0);

        float2 A[6][CNR];               // [plane slot][row]
#pragma unroll
        for (int s = 0; s < 6; ++s)
#pragma unroll
            for (int i = 0; i < CNR; ++i) A[s][i] = make_float2(0.f, 0.f);
        int pbase = -1, slot = 0;       // first plane of the window and its slot (warp-uniform)

        // plane p (slot s) leaves the window: add this lane's cells to the grid (vector REDs), clear the slot.
        // The wrap factors are +-1: applied as sign-bit flips.
#define COL_FLUSH_SLOT(SS)                                                                         \
    case SS: {                                                                                     \
        _Pragma("unroll") for (int i = 0; i < CNR; ++i) {                                          \
            const unsigned fb = pbit ^ rbit[i];                                                    \
            const float2 v = make_float2(__uint_as_float(__float_as_uint(A[SS][i].x) ^ fb),        \
                                         __uint_as_float(__float_as_uint(A[SS][i].y) ^ fb));       \
            red_v2(cell_at(cell[i], poff), v);                                                     \
            A[SS][i] = make_float2(0.f, 0.f);                                                      \
        }                                                                                          \
    } break;
        auto flush = [&](int s, int p) {
            unsigned pbit = 0u;
            if (p >= g.K0) { p -= g.K0; pbit = sbit0; }
            const int poff = p * KK;                    // < prod(Kd) < 2^31
            switch (s) {
                COL_FLUSH_SLOT(0)
                COL_FLUSH_SLOT(1)
                COL_FLUSH_SLOT(2)
                COL_FLUSH_SLOT(3)
                COL_FLUSH_SLOT(4)
                COL_FLUSH_SLOT(5)
            }
        };
#undef COL_FLUSH_SLOT

        for (int k = 0; k < nchunks; ++k) {
            const int ns = min(CCH, wi.end - (wi.begin + k * CCH));
            const unsigned b = (gk + k) & 1;
            if (k + 1 < nchunks) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) issue(k + 1);
            }
            mbar_wait(&mbar[b], ((gk + k) >> 1) & 1);
            const float* Rb = reinterpret_cast<const float*>(ws + b * REC_BYTES);
            const float4* Y = reinterpret_cast<const float4*>(ybuf + b * YS_BYTES);
#pragma unroll 2
            for (int u = 0; u < ns; ++u) {
                const float* R = Rb + u * CRECW;
                const float4 C1 = *reinterpret_cast<const float4*>(R + 4 * lg);       // c1 of rows lg, lg+3, lg+6
                const float4 C0a = *reinterpret_cast<const float4*>(R + 12);          // c0rot[0..3]
                const float4 C0b = *reinterpret_cast<const float4*>(R + 16);          // c0rot[4..5], p0, p0 mod 6
                const float w = R[20 + lw];
                const float2 yv = *reinterpret_cast<const float2*>(Y + u);            // conj(P'') * y
                const int p0 = __float_as_int(C0b.z);
                if (p0 != pbase) {                        // warp-uniform: the window moved
                    if (pbase >= 0) {
                        const int n = min(p0 - pbase, 6);
                        int s = slot, p = pbase;
                        for (int e = 0; e < n; ++e) {
                            flush(s, p);
                            s = (s == 5) ? 0 : s + 1;
                            ++p;
                        }
                    }
                    pbase = p0;
                    slot = __float_as_int(C0b.w);
                }
                const float vx = w * yv.x, vy = w * yv.y;
                const float c1v[CNR] = {C1.x, C1.y, C1.z};
                const float c0v[6] = {C0a.x, C0a.y, C0a.z, C0a.w, C0b.x, C0b.y};
#pragma unroll
                for (int i = 0; i < CNR; ++i) {
                    const float tx = c1v[i] * vx, ty = c1v[i] * vy;
#pragma unroll
                    for (int s = 0; s < 6; ++s) {
                        A[s][i].x = fmaf(c0v[s], tx, A[s][i].x);
                        A[s][i].y = fmaf(c0v[s], ty, A[s][i].y);
                    }
                }
            }
            __syncwarp();
        }
        if (pbase >= 0) {
            int s = slot, p = pbase;
            for (int e = 0; e < 6; ++e) {
                flush(s, p);
                s = (s == 5) ? 0 : s + 1;
                ++p;
            }
        }
        gk += nchunks;
    }
}

// ---------------------------------------------------------------------------------------------------------
// grid demodulation: grid[g] *= conj(prod_d m_d[g_d]); one CTA per (g0, g1) row
// ---------------------------------------------------------------------------------------------------------
__global__ void k_demodulate(float2* __restrict__ grid, const float2* __restrict__ mod, int K0, int K1, int K2) {
    const int row = blockIdx.x;                        // g0 * K1 + g1
    const int g0 = row / K1, g1 = row - g0 * K1;
    const long long cb = (long long)blockIdx.y * K0 * K1 * K2;
    const float2 m01 = cmul(__ldg(mod + g0), __ldg(mod + K0 + g1));
    const float2* m2 = mod + K0 + K1;
    float2* dst = grid + cb + (long long)row * K2;
    for (int g2 = threadIdx.x; g2 < K2; g2 += blockDim.x) dst[g2] = cmulc(cmul(m01, __ldg(m2 + g2)), dst[g2]);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool col3d_supported(const Geom& g) {
    if (g.ndim != 3) return false;
    for (int d = 0; d < 3; ++d)
        if (g.J[d] != 6) return false;
    return g.K[0] >= 6 && g.K[1] >= CROWS && g.K[2] >= CCOLS;
}

static ColGeom col_geom(const Geom& g) {
    ColGeom c;
    c.K0 = g.K[0]; c.K1 = g.K[1]; c.K2 = g.K[2];
    c.nq2 = (g.K[2] + CT2 - 1) / CT2;
    c.sg0 = ((g.N[0] - 1) & 1) ? -1.f : 1.f;
    c.sg1 = ((g.N[1] - 1) & 1) ? -1.f : 1.f;
    c.sg2 = ((g.N[2] - 1) & 1) ? -1.f : 1.f;
    c.Kprod = g.Kprod;
    return c;
}

// CTAs per coil of the persistent kernel: all SMs x 5 resident CTAs, shared between the coils of the launch
static int col_ctas(b200nufft_plan_t p, int nb) {
    if (p->n_sm == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) p->n_sm = prop.multiProcessorCount;
        if (p->n_sm <= 0) p->n_sm = 148;
    }
    return std::max(1, (p->n_sm * 5 + nb - 1) / nb);
}

int col3d_demodulate(b200nufft_plan_t p, float2* grid, int nb, cudaStream_t st) {
    const Geom& g = p->g;
    dim3 gr((unsigned)(g.K[0] * g.K[1]), nb);
    k_demodulate<<<gr, 128, 0, st>>>(grid, p->d_mod, g.K[0], g.K[1], g.K[2]);
    LAUNCH_CHECK();
    return B200_OK;
}

// grid receives the phase-modulated adjoint (it is zeroed here, by the pre-pass)
int col3d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    if (!p->attr_col) {
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * GWARP_BYTES));
        p->attr_col = true;
    }
    const long long nel = p->g.Kprod * nb;
    if (p->n_cwork == 0) {
        CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * nel, st));
        return B200_OK;
    }
    if (p->ys_nb < nb) {
        if (p->d_ys) { CUDA_TRY(cudaFree(p->d_ys)); p->d_ys = nullptr; p->ys_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ys, sizeof(float4) * p->M * nb));
        p->ys_nb = nb;
    }
    if (p->ccount_nb < nb) {
        if (p->d_ccount) { CUDA_TRY(cudaFree(p->d_ccount)); p->d_ccount = nullptr; p->ccount_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ccount, sizeof(int) * nb));
        p->ccount_nb = nb;
    }
    // float4 stores need a 16-byte aligned grid and an even element count; otherwise plain memset
    const bool vec = (reinterpret_cast<uintptr_t>(grid) & 15) == 0 && (nel & 1) == 0;
    if (!vec) CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * nel, st));
    {
        const int TB = 256;
        k_gather_sorted_col<<<(unsigned)((p->M + TB - 1) / TB), TB, 0, st>>>(
            p->d_cside, p->M, y, p->d_ys, nb, reinterpret_cast<float4*>(grid), vec ? nel / 2 : 0, p->d_ccount);
        LAUNCH_CHECK();
    }
    dim3 gr((unsigned)std::min((p->n_cwork + CWARPS - 1) / CWARPS, col_ctas(p, nb)), nb);
    k_gridding_col<<<gr, CWARPS * 32, CWARPS * GWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                                 p->d_crec, p->d_ys, p->M, grid);
    LAUNCH_CHECK();
    return B200_OK;
}
