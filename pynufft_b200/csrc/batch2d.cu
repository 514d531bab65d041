// 2-D multi-coil kernels with the COIL ON THE LANES: configurations 2 and 4 (256^2 / 512^2, 32 coils).
// Used whenever a 2-D J=6 call has nb >= 8 coils.  The grids stay coil-major in global memory (so the batched
// cuFFT and the pad/crop kernels are unchanged); the tile is transposed into shared memory as
// box[position][coil] (the reference's batch-innermost order, `vec[col*Reps + nc]`, src/re_subroutine.py:808),
// so that every neighbour access of a warp is one conflict-free 8*32-byte row and the interpolation weights of
// a sample are computed once for all coils.
//
//   k_interp2d_bi   replaces pELL_spmv_mCoil  (re_subroutine.py:751-835): one CTA per (8 x 16 sub-tile, sample
//                   range); the 13 x 21 x 32 box is staged already multiplied by the separable phase
//                   F0[row]*F1[col], so all weights are real; one warp per sample, lane = coil,
//                   36 x (LDS.64 + 2 FFMA); y[m, :] is stored as one coalesced row.
//   k_gridding2d_bi replaces pELL_spmvh_mCoil + atomic_add_float2 (:527-596, :275-287): same box as
//                   accumulator; warp w owns the box rows p = w (mod 8), so no two warps ever touch the same
//                   address and every warp walks the whole sample list picking the one row it owns of each
//                   footprint: plain LDS/FFMA/STS, no atomics; the box is flushed once with vector REDs
//                   (halos of neighbouring boxes overlap).
#include "common.cuh"

namespace {

constexpr int BJ = 6;
constexpr int BS0 = 8, BS1 = 16;          // sub-tile
constexpr int BR = BS0 + BJ - 1;          // 13 box rows
constexpr int BC = BS1 + BJ - 1;          // 21 box columns
constexpr int NW = 8;                     // warps per CTA
constexpr int RECW2 = 20;                 // words per 2-D plan record: c0[6] c1[6] P(2) ks0 ks1 perm pad3
constexpr int SCH = 32;                   // samples per record chunk
constexpr int BP = 33;                    // coil pitch of a box position (complex): transposing stores stay 2-way

__device__ __forceinline__ int wrap2(int i, int K) {
    i -= (i >= K) ? K : 0;
    i -= (i >= K) ? K : 0;
    return i;
}

// ys[i, c] = y[perm[i], c]   (rows of 8*nb bytes)
__global__ void k_gather_rows(const int* __restrict__ perm, long long M, const float2* __restrict__ y,
                              float2* __restrict__ ys, int nb) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= M * nb) return;
    const long long i = gid / nb;
    const int c = (int)(gid - i * nb);
    ys[gid] = y[(long long)perm[i] * nb + c];
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the two tile kernels
// ---------------------------------------------------------------------------------------------
struct TileCtx {
    int O0, O1;        // box origin in the grid
    int c0;            // first coil of this CTA
    int nc;            // coils handled by this CTA (<= 32)
};

__device__ __forceinline__ TileCtx tile_ctx(const Geom& g, const WorkItem& wi, int nb) {
    TileCtx t;
    int bin = wi.tile;
    int sb = bin % g.nsubprod;
    int tl = bin / g.nsubprod;
    const int s1 = sb % g.nsub[1], s0 = sb / g.nsub[1];
    const int q1 = tl % g.ntile[1], q0 = tl / g.ntile[1];
    t.O0 = q0 * g.tile[0] + s0 * BS0;
    t.O1 = q1 * g.tile[1] + s1 * BS1;
    t.c0 = blockIdx.y * 32;
    t.nc = min(32, nb - t.c0);
    return t;
}

// record chunk -> shared memory, fixed up: [c0[6] c1[6] | P' (2) | rel0 rel1 perm | ..]
//   P' = P * G0[rel0] * Gl[rel1]  (per-sample part of the separable phase)
__device__ __forceinline__ void load_records(const Geom& g, const float* __restrict__ rec, int s0, int ns,
                                             const TileCtx& t, float* srec, int tid, int nthreads) {
    for (int q = tid; q < ns * (RECW2 / 4); q += nthreads)
        reinterpret_cast<float4*>(srec)[q] = __ldg(reinterpret_cast<const float4*>(rec + (long long)s0 * RECW2) + q);
    __syncthreads();
    if (tid < ns) {
        float* R = srec + tid * RECW2;
        const int ks0 = __float_as_int(R[14]), ks1 = __float_as_int(R[15]);
        const int rel0 = ks0 - t.O0, rel1 = ks1 - t.O1;
        const float2 Pp = cmul(cmul(make_float2(R[12], R[13]), g.G0[rel0]), g.Gl[rel1]);
        R[12] = Pp.x;
        R[13] = Pp.y;
        R[14] = __int_as_float(rel0);
        R[15] = __int_as_float(rel1);
    }
    __syncthreads();
}

constexpr size_t BOX_BYTES = (((size_t)BR * BC * BP * sizeof(float2)) + 127) / 128 * 128;   // 72,192 (records need 16-byte alignment)
constexpr size_t INTERP_SMEM = BOX_BYTES + SCH * RECW2 * sizeof(float);
constexpr size_t GRID_SMEM = BOX_BYTES + SCH * RECW2 * sizeof(float) + SCH * 32 * sizeof(float2);

// ---------------------------------------------------------------------------------------------
// interpolation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NW * 32)
k_interp2d_bi(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
              const float2* __restrict__ grid, float2* __restrict__ y, int nb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);                        // [BR][BC][32]
    float* srec = reinterpret_cast<float*>(smem_raw + BOX_BYTES);
    const WorkItem wi = work[blockIdx.x];
    const TileCtx t = tile_ctx(g, wi, nb);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cact = lane < t.nc;
    // ---- stage the box from the coil-major grids: one warp per (coil, box row), lanes = columns (168-byte
    //      coalesced reads), times F0[row]*F1[col], transposed into box[position][coil] ----
    {
        const bool act = lane < BC;
        const int i1 = wrap2(t.O1 + (act ? lane : 0), g.K[1]);
        const float2 F1 = g.Fl[act ? lane : 0];
        constexpr int UN = 4;
        for (int task0 = warp * UN; task0 < t.nc * BR; task0 += NW * UN) {
            float2 v[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const int task = task0 + q;
                if (task < t.nc * BR && act) {
                    const int c = task / BR, p = task - c * BR;
                    const int i0 = wrap2(t.O0 + p, g.K[0]);
                    v[q] = __ldg(grid + (long long)(t.c0 + c) * g.Kprod + (long long)i0 * g.K[1] + i1);
                }
            }
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const int task = task0 + q;
                if (task < t.nc * BR && act) {
                    const int c = task / BR, p = task - c * BR;
                    box[(p * BC + lane) * BP + c] = cmul(v[q], cmul(g.F0[p], F1));
                }
            }
        }
        // coils beyond nc (partial last chunk) read as zero
        if (t.nc < 32)
            for (int e = tid; e < BR * BC * (32 - t.nc); e += NW * 32)
                box[(e / (32 - t.nc)) * BP + t.nc + e % (32 - t.nc)] = make_float2(0.f, 0.f);
    }
    for (int s0 = wi.begin; s0 < wi.end; s0 += SCH) {
        const int ns = min(SCH, wi.end - s0);
        __syncthreads();                                     // previous chunk consumed / box staged
        load_records(g, rec, s0, ns, t, srec, tid, NW * 32);
        for (int u = warp; u < ns; u += NW) {
            const float* R = srec + u * RECW2;
            const float4 a0 = *reinterpret_cast<const float4*>(R), a1 = *reinterpret_cast<const float4*>(R + 4),
                         a2 = *reinterpret_cast<const float4*>(R + 8), a3 = *reinterpret_cast<const float4*>(R + 12);
            const float c0[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
            const float c1[6] = {a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
            const int rel0 = __float_as_int(a3.z), rel1 = __float_as_int(a3.w);
            const float2* bp = box + ((rel0 * BC) + rel1) * BP + lane;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int j0 = 0; j0 < 6; ++j0) {
                float2 rs = make_float2(0.f, 0.f);
#pragma unroll
                for (int j1 = 0; j1 < 6; ++j1) {
                    const float2 kv = bp[(j0 * BC + j1) * BP];
                    rs.x = fmaf(c1[j1], kv.x, rs.x);
                    rs.y = fmaf(c1[j1], kv.y, rs.y);
                }
                acc.x = fmaf(c0[j0], rs.x, acc.x);
                acc.y = fmaf(c0[j0], rs.y, acc.y);
            }
            if (cact) {
                const int m = __float_as_int(R[16]);
                y[(long long)m * nb + t.c0 + lane] = cmul(make_float2(a3.x, a3.y), acc);   // coalesced row of y (M, B)
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// gridding
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NW * 32)
k_gridding2d_bi(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
                const float2* __restrict__ ys, float2* __restrict__ grid, int nb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);                        // [BR][BC][32]
    float* srec = reinterpret_cast<float*>(smem_raw + BOX_BYTES);
    float2* sy = reinterpret_cast<float2*>(smem_raw + BOX_BYTES + SCH * RECW2 * sizeof(float));   // [SCH][32]
    const WorkItem wi = work[blockIdx.x];
    const TileCtx t = tile_ctx(g, wi, nb);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cact = lane < t.nc;
    for (int e = tid; e < BR * BC * BP; e += NW * 32) box[e] = make_float2(0.f, 0.f);
    for (int s0 = wi.begin; s0 < wi.end; s0 += SCH) {
        const int ns = min(SCH, wi.end - s0);
        __syncthreads();
        // sorted data of the chunk: [ns][nc] -> sy[u][lane]
        for (int u = warp; u < ns; u += NW)
            sy[u * 32 + lane] = cact ? __ldg(ys + (long long)(s0 + u) * nb + t.c0 + lane) : make_float2(0.f, 0.f);
        load_records(g, rec, s0, ns, t, srec, tid, NW * 32);
        // every warp walks the chunk and handles the footprint row it owns: box row p == warp (mod 8)
        for (int u = 0; u < ns; ++u) {
            const float* R = srec + u * RECW2;
            const int rel0 = __float_as_int(R[14]);
            const int j0 = (warp - rel0) & 7;                 // p = rel0 + j0 == warp (mod 8)
            if (j0 < 6) {                                      // warp-uniform
                const float4 a1 = *reinterpret_cast<const float4*>(R + 4), a2 = *reinterpret_cast<const float4*>(R + 8);
                const float c1[6] = {a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
                const float w0 = R[j0];
                const int rel1 = __float_as_int(R[15]);
                const float2 Yp = cmulc(make_float2(R[12], R[13]), sy[u * 32 + lane]);    // conj(P') * y
                const float2 wv = make_float2(w0 * Yp.x, w0 * Yp.y);
                float2* bp = box + (((rel0 + j0) * BC) + rel1) * BP + lane;
#pragma unroll
                for (int j1 = 0; j1 < 6; ++j1) {
                    float2 v = bp[j1 * BP];
                    v.x = fmaf(c1[j1], wv.x, v.x);
                    v.y = fmaf(c1[j1], wv.y, v.y);
                    bp[j1 * BP] = v;
                }
            }
        }
    }
    __syncthreads();
    // ---- flush: box * conj(F0[row]*F1[col]) -> coil-major grids (periodic): one warp per (coil, box row),
    //      lanes = columns, vector REDs on 168-byte row segments ----
    {
        const bool act = lane < BC;
        const int i1 = wrap2(t.O1 + (act ? lane : 0), g.K[1]);
        float2 F1 = g.Fl[act ? lane : 0];
        for (int task = warp; task < t.nc * BR; task += NW) {
            const int c = task / BR, p = task - c * BR;
            if (act) {
                const float2 v = box[(p * BC + lane) * BP + c];
                if (v.x != 0.f || v.y != 0.f) {
                    const int i0 = wrap2(t.O0 + p, g.K[0]);
                    atomicAdd(grid + (long long)(t.c0 + c) * g.Kprod + (long long)i0 * g.K[1] + i1,
                              cmulc(cmul(g.F0[p], F1), v));
                }
            }
        }
    }
}

}  // namespace

// nb >= 8 coils on a 2-D J=6 plan -> batch-innermost kernels
bool batch2d_supported(const Geom& g, int nb) {
    return g.ndim == 2 && nb >= 8 && g.J[0] == BJ && g.J[1] == BJ && g.sub[0] == BS0 && g.sub[1] == BS1 &&
           g.recw == RECW2;
}

int batch2d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st) {
    if (!p->attr_b2d) {      // once per plan (the attribute is per device, plans are per device)
        CUDA_TRY(cudaFuncSetAttribute(k_interp2d_bi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTERP_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(k_gridding2d_bi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRID_SMEM));
        p->attr_b2d = true;
    }
    if (p->n_gwork == 0) return B200_OK;
    dim3 gr(p->n_gwork, (nb + 31) / 32);
    k_interp2d_bi<<<gr, NW * 32, INTERP_SMEM, st>>>(p->g, p->d_gwork, p->d_rec, grid, y, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

int batch2d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    if (!p->attr_b2d) {      // once per plan (the attribute is per device, plans are per device)
        CUDA_TRY(cudaFuncSetAttribute(k_interp2d_bi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INTERP_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(k_gridding2d_bi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRID_SMEM));
        p->attr_b2d = true;
    }
    if (p->n_gwork == 0) return B200_OK;
    if (p->ysb_elems < p->M * nb) {
        if (p->d_ysb) { CUDA_TRY(cudaFree(p->d_ysb)); p->d_ysb = nullptr; p->ysb_elems = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ysb, sizeof(float2) * p->M * nb));
        p->ysb_elems = p->M * nb;
    }
    const long long tot = p->M * nb;
    k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p->d_perm, p->M, y, p->d_ysb, nb);
    LAUNCH_CHECK();
    dim3 gr(p->n_gwork, (nb + 31) / 32);
    k_gridding2d_bi<<<gr, NW * 32, GRID_SMEM, st>>>(p->g, p->d_gwork, p->d_rec, p->d_ysb, grid, nb);
    LAUNCH_CHECK();
    return B200_OK;
}
