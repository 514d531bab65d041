// Tiled interpolation (gather) kernel for 3-D, J = 6: the replacement of pELL_spmv_mCoil
// (src/re_subroutine.py:751-835) on the headline configuration (128^3 / 256^3 / 6^3).
//
// One CTA per work item = (8 x 16 x 16 slab of first-neighbour cells, range of bin-sorted samples).
//  * the 13 x 21 x 21 box of grid values the slab's samples can touch (slab + J-1 halo, periodic
//    wrap) is staged once into shared memory, one warp per box row (coalesced 168-byte row
//    reads).  Every staged value is multiplied by the last-dimension phase Fl[column], which turns
//    the last-dimension interpolation weights into REAL numbers (the remaining per-sample phase
//    Gl[rel] is folded into the sample's phase P).  Layout: row pitch 21, plane pitch 446 complex
//    -> the 16 rows read by one half-warp phase fall into 16 different 8-byte bank pairs;
//  * the plan's sample records (96 B each, contiguous in bin order) arrive in shared memory by ONE
//    TMA bulk copy per sub-chunk (cp.async.bulk + mbarrier) and are fixed up in place
//    (tile-relative base address, P * Gl[rel]);
//  * a warp takes 8 samples at a time: lane l owns footprint row (j0, j1) = divmod(l, 6), reads its
//    6 contiguous grid values (LDS.64 x6) and contracts them with the warp-uniform real
//    last-dim weights (12 FFMA); rows 32..35 of the 8 samples are packed into one extra pass
//    (4 lanes per sample); the 8 partial sums are reduced with a transposed butterfly
//    (18 shuffles per 8 samples instead of 80) and scattered to y through the sort permutation.
// No atomics, no global traffic inside the loop besides the y store.
//
// MOD = true: the kernel reads the phase-modulated grid G'[g] = G[g] * prod_d e^{i s_d g_d} that the column-sweep
// gridding kernel produces (col3d.cu), so that k-space solvers can iterate on modulated vectors without converting.
// The staged box then only needs the wrap signs (-1)^(N-1) of neighbours that wrap around the periodic grid, every
// interpolation weight is real, and the sample's phase becomes P * prod_d e^{i s_d (1 - k_d)} (k_d = first neighbour).
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int TJ = 6;
constexpr int TT = 16;                    // tile edge (dims 1, 2)
constexpr int TS = 8;                     // slab thickness (dim 0)
constexpr int BOX = TT + TJ - 1;          // 21
constexpr int NPL = TS + TJ - 1;          // 13 planes
constexpr int RP = TILE_RP;               // row pitch (complex), odd: 21
constexpr int PP = TILE_PP;               // plane pitch (complex) >= 21*21, == 6*RP (mod 16): 446
constexpr int TILE_ELEMS = NPL * PP;      // 5798
#ifndef IT_SUBCHUNK
#define IT_SUBCHUNK 256
#endif
#ifndef IT_CTAS
#define IT_CTAS 3
#endif
constexpr int SUBCHUNK = IT_SUBCHUNK;     // samples per record sub-chunk
constexpr int SRW = 24;                   // words per record
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int RECW = TILE_RECW;           // words per gather record (plan.cu k_build_records): 24
constexpr size_t SMEM_BYTES = TILE_ELEMS * sizeof(float2) + SUBCHUNK * SRW * sizeof(float) + 16;

// gather record: [c0[6] | c1[6] | c2[6] | base, perm | P'.re, P'.im | P''.re, P''.im]   (built at plan time)
//                 0       6       12      18    19     20     21      22      23

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

__device__ __forceinline__ float2 row_dot(const float2* __restrict__ tp, const float4 C0, const float2 C1) {
    const float2 k0 = tp[0], k1 = tp[1], k2 = tp[2], k3 = tp[3], k4 = tp[4], k5 = tp[5];
    float2 rs;
    rs.x = C0.x * k0.x;
    rs.y = C0.x * k0.y;
    rs.x = fmaf(C0.y, k1.x, rs.x);
    rs.y = fmaf(C0.y, k1.y, rs.y);
    rs.x = fmaf(C0.z, k2.x, rs.x);
    rs.y = fmaf(C0.z, k2.y, rs.y);
    rs.x = fmaf(C0.w, k3.x, rs.x);
    rs.y = fmaf(C0.w, k3.y, rs.y);
    rs.x = fmaf(C1.x, k4.x, rs.x);
    rs.y = fmaf(C1.x, k4.y, rs.y);
    rs.x = fmaf(C1.y, k5.x, rs.x);
    rs.y = fmaf(C1.y, k5.y, rs.y);
    return rs;
}

__device__ __forceinline__ int wrap2(int i, int K) {   // i in [0, 3K)
    i -= (i >= K) ? K : 0;
    i -= (i >= K) ? K : 0;
    return i;
}
__device__ __forceinline__ int wrap2s(int i, int K, float sg, float& sign) {   // same; sign *= sg per wrap
    if (i >= K) { i -= K; sign *= sg; }
    if (i >= K) { i -= K; sign *= sg; }
    return i;
}

template <bool MOD>
__global__ void __launch_bounds__(NTHREADS, IT_CTAS)
k_interp_tiled(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
               const float2* __restrict__ grid, float2* __restrict__ y, int nb, const float2* __restrict__ mod) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* tile = reinterpret_cast<float2*>(smem_raw);
    float* srec = reinterpret_cast<float*>(smem_raw + TILE_ELEMS * sizeof(float2));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + TILE_ELEMS * sizeof(float2) + SUBCHUNK * SRW * sizeof(float));

    const WorkItem wi = work[blockIdx.x];
    const int c = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int t = wi.tile;
    const int q2 = t % g.ntile[2];
    t /= g.ntile[2];
    const int q1 = t % g.ntile[1];
    const int q0 = t / g.ntile[1];
    const int T0 = q0 * TT + wi.pad, T1 = q1 * TT, T2 = q2 * TT;
    const float2* gc = grid + (long long)c * g.Kprod;

    // ---- records of the first sub-chunk: one TMA bulk copy ----
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    unsigned phase = 0;
    if (tid == 0) {
        const int ns0 = min(SUBCHUNK, wi.end - wi.begin);
        tma_bulk_load(srec, rec + (long long)wi.begin * RECW, (unsigned)(ns0 * RECW * sizeof(float)), mbar);
    }

    // ---- stage the 13 x 21 x 21 box (periodic) times Fl[column]: one warp per row, 7 loads in flight ----
    {
        // warp w owns rows r = w, w+8, w+16 of every plane; 32-bit element offsets (prod(Kd) < 2^31)
        const int K0 = g.K[0], K1 = g.K[1], K2 = g.K[2];
        const int KK = K1 * K2;
        const bool act = lane < BOX;
        // MOD: the staged values only take the wrap signs (-1)^(N-1) per wrapped dimension
        const float sg0 = ((g.N[0] - 1) & 1) ? -1.f : 1.f, sg1 = ((g.N[1] - 1) & 1) ? -1.f : 1.f,
                    sg2 = ((g.N[2] - 1) & 1) ? -1.f : 1.f;
        float s2 = 1.f;
        const int i2 = wrap2s(T2 + (act ? lane : 0), K2, sg2, s2);
        const float2 F = g.Fl[act ? lane : 0];
        constexpr int RQ = (BOX + NWARPS - 1) / NWARPS;   // 3
        int roff[RQ], soff[RQ];
        float rsg[RQ];
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const int r = warp + NWARPS * q;
            const bool ok = act && r < BOX;
            rsg[q] = s2;
            roff[q] = ok ? wrap2s(T1 + (r < BOX ? r : 0), K1, sg1, rsg[q]) * K2 + i2 : -1;
            soff[q] = r * RP + lane;
        }
        constexpr int PU = 4;                             // planes per batch -> 12 loads in flight per lane
#pragma unroll
        for (int p0 = 0; p0 < NPL; p0 += PU) {
            float2 v[PU][RQ];
#pragma unroll
            for (int pp = 0; pp < PU; ++pp) {
                if (p0 + pp < NPL) {
                    const int pb = wrap2(T0 + p0 + pp, K0) * KK;
#pragma unroll
                    for (int q = 0; q < RQ; ++q)
                        if (roff[q] >= 0) v[pp][q] = __ldg(gc + (unsigned)(pb + roff[q]));
                }
            }
#pragma unroll
            for (int pp = 0; pp < PU; ++pp) {
                if (p0 + pp < NPL) {
                    float sp = 1.f;
                    if (MOD) wrap2s(T0 + p0 + pp, K0, sg0, sp);
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        if (roff[q] >= 0) {
                            const float f = sp * rsg[q];
                            tile[(p0 + pp) * PP + soff[q]] =
                                MOD ? make_float2(v[pp][q].x * f, v[pp][q].y * f) : cmul(v[pp][q], F);
                        }
                    }
                }
            }
        }
    }

    // ---- per-lane constants ----
    const int j0l = lane / 6 > 5 ? 5 : lane / 6, j1l = lane % 6;      // rows 0..31
    const int rowoff = j0l * PP + j1l * RP;
    const float2 E01 = MOD ? make_float2(1.f, 0.f) : cmul(g.E[0][j0l], g.E[1][j1l]);
    const int rr = lane & 3;                                           // rows 32..35 = (5, 2+rr)
    const int rowoff_r = 5 * PP + (2 + rr) * RP;
    const float2 E01r = MOD ? make_float2(1.f, 0.f) : cmul(g.E[0][5], g.E[1][2 + rr]);
    const int ur = lane >> 2;

    for (int sb = wi.begin; sb < wi.end; sb += SUBCHUNK) {
        const int ns = min(SUBCHUNK, wi.end - sb);
        const int nsr = (ns + 7) & ~7;
        if (sb != wi.begin) {
            // previous sub-chunk fully consumed (trailing __syncthreads); order generic writes before the TMA write
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) tma_bulk_load(srec, rec + (long long)sb * RECW, (unsigned)(ns * RECW * sizeof(float)), mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        // the rows that pad the chunk to a multiple of 8 samples hold stale (first chunk: uninitialised) words: give them a
        // valid box address; their results are never stored
        if (tid < nsr - ns) srec[(ns + tid) * SRW + 18] = 0.f;
        __syncthreads();    // tile staged, chunk complete

        // ---- main loop: 8 samples per warp pass ----
        for (int b = warp; b < (nsr >> 3); b += NWARPS) {
            float2 acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* R = srec + (b * 8 + u) * SRW;
                const float w01 = R[j0l] * R[6 + j1l];
                const float4 C0 = *reinterpret_cast<const float4*>(R + 12);
                const float4 C1b = *reinterpret_cast<const float4*>(R + 16);          // c2[4], c2[5], base, perm
                const int base = __float_as_int(C1b.z);
                const float2 C1 = make_float2(C1b.x, C1b.y);
                const float2 rs = row_dot(tile + base + rowoff, C0, C1);
                const float2 tt = cmul(rs, E01);
                acc[u] = make_float2(tt.x * w01, tt.y * w01);
            }
            {   // rows 32..35 of the 8 samples: lane -> (sample ur, row 32 + rr)
                const float* R = srec + (b * 8 + ur) * SRW;
                const float w01 = R[5] * R[6 + 2 + rr];
                const float4 C0 = *reinterpret_cast<const float4*>(R + 12);
                const float4 C1b = *reinterpret_cast<const float4*>(R + 16);
                const int base = __float_as_int(C1b.z);
                const float2 C1 = make_float2(C1b.x, C1b.y);
                const float2 rs = row_dot(tile + base + rowoff_r, C0, C1);
                const float2 tt = cmul(rs, E01r);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ur == u) {
                        acc[u].x = fmaf(tt.x, w01, acc[u].x);
                        acc[u].y = fmaf(tt.y, w01, acc[u].y);
                    }
                }
            }
            // ---- transposed butterfly: 8 values x 32 lanes -> 8 sums ----
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool hi = lane & 16;
                const float2 send = hi ? acc[u] : acc[u + 4];
                const float2 keep = hi ? acc[u + 4] : acc[u];
                acc[u].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 16);
                acc[u].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 16);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const bool hi = lane & 8;
                const float2 send = hi ? acc[u] : acc[u + 2];
                const float2 keep = hi ? acc[u + 2] : acc[u];
                acc[u].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 8);
                acc[u].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 8);
            }
            {
                const bool hi = lane & 4;
                const float2 send = hi ? acc[0] : acc[1];
                const float2 keep = hi ? acc[1] : acc[0];
                acc[0].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 4);
                acc[0].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 4);
            }
            acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, 2);
            acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, 2);
            acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, 1);
            acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, 1);
            if ((lane & 3) == 0) {
                const int sid = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                const int s = b * 8 + sid;
                if (s < ns) {
                    const float* R = srec + s * SRW;
                    const int m = __float_as_int(R[19]);
                    y[(long long)m * nb + c] = cmul(MOD ? make_float2(R[22], R[23]) : make_float2(R[20], R[21]), acc[0]);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

bool tiled_supported(const Geom& g) {
    if (g.ndim != 3) return false;
    for (int d = 0; d < 3; ++d)
        if (g.J[d] != TJ || g.tile[d] != TT || g.sub[d] != TS) return false;
    return true;
}

int interp_tiled_launch(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st, bool modulated) {
    // tuning knob: extra dynamic shared memory per CTA (lowers the number of resident CTAs per SM; co-scheduling experiments)
    static const size_t pad = [] { const char* e = getenv("B200NUFFT_IT_PAD"); return e ? (size_t)atoi(e) : (size_t)0; }();
    const size_t smem = SMEM_BYTES + pad;
    if (!p->attr_interp) {      // once per plan (the attribute is per device, plans are per device)
        CUDA_TRY(cudaFuncSetAttribute(k_interp_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(k_interp_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        p->attr_interp = true;
    }
    if (modulated && !p->d_mod) {
        b200_set_error("interp: modulated grid requested but the plan has no modulation tables");
        return B200_ERR_UNSUPPORTED;
    }
    if (p->n_work == 0) return B200_OK;
    dim3 gr(p->n_work, nb);
    if (modulated)
        k_interp_tiled<true><<<gr, NTHREADS, smem, st>>>(p->g, p->d_work, p->d_trec, grid, y, nb, p->d_mod);
    else
        k_interp_tiled<false><<<gr, NTHREADS, smem, st>>>(p->g, p->d_work, p->d_trec, grid, y, nb, nullptr);
    LAUNCH_CHECK();
    return B200_OK;
}
