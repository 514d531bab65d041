// Tiled gridding (scatter) kernel for 3-D, J = 6: the replacement of pELL_spmvh_mCoil +
// atomic_add_float2 (src/re_subroutine.py:527-596, 275-287) on the headline configuration.
//
// Shared-memory float atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN), so the accumulation is
// made conflict-free by construction instead:
//  * one WARP per work item = (8^3 sub-tile of first-neighbour cells, range of bin-sorted samples);
//    the warp owns a private 13^3 accumulation box in shared memory (row pitch 13, plane pitch 174
//    complex -> the 16 rows of a half-warp phase hit 16 different bank pairs);
//  * samples are processed one at a time: lane l owns footprint row (j0, j1) = divmod(l, 6) and does
//    a plain read-modify-write of its 6 contiguous elements (LDS.64 / 4 FFMA / STS.64 each); the
//    remaining rows 32..35 (24 elements) take one element per lane.  All lanes of one step touch
//    distinct addresses and no other warp shares the box -> no atomics, no races;
//  * when the range is done the box is flushed once to the global grid with vector reductions
//    (REDG.E.ADD.F32x2), because the halos of neighbouring boxes overlap.
// The 2*M*prod(J) global float atomics of the reference become ~4.3 * prod(Kd) vector REDs.
#include "common.cuh"

namespace {

constexpr int GJ = 6;
constexpr int GT = 8;
constexpr int GBOX = GT + GJ - 1;        // 13
constexpr int GRP = 13;                  // row pitch (complex), odd
constexpr int GPP = 174;                 // plane pitch >= 13*13, == 6*GRP (mod 16)
constexpr int GBOX_ELEMS = GBOX * GPP;   // 2262
constexpr int GBATCH = 8;                // samples expanded per pass
constexpr int GSRW = 28;                 // words per expanded record
constexpr int RECW = 24;
constexpr size_t GSMEM_BYTES = GBOX_ELEMS * sizeof(float2) + GBATCH * GSRW * sizeof(float);

__device__ __forceinline__ void rmw(float2* p, float2 b, float2 w) {
    float2 v = *p;
    cfma(v, b, w);
    *p = v;
}

__global__ void __launch_bounds__(32)
k_gridding_tiled(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
                 const float2* __restrict__ y, float2* __restrict__ grid, int nb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);
    float* srec = reinterpret_cast<float*>(smem_raw + GBOX_ELEMS * sizeof(float2));

    const WorkItem wi = work[blockIdx.x];
    const int c = blockIdx.y;
    const int lane = threadIdx.x;
    // bin id -> tile coords and sub-tile coords -> box origin
    int bin = wi.tile;
    int sb = bin % g.nsubprod;
    int t = bin / g.nsubprod;
    const int s2 = sb % g.nsub[2];
    sb /= g.nsub[2];
    const int s1 = sb % g.nsub[1];
    const int s0 = sb / g.nsub[1];
    const int q2 = t % g.ntile[2];
    t /= g.ntile[2];
    const int q1 = t % g.ntile[1];
    const int q0 = t / g.ntile[1];
    const int O0 = q0 * g.tile[0] + s0 * GT, O1 = q1 * g.tile[1] + s1 * GT, O2 = q2 * g.tile[2] + s2 * GT;

    for (int e = lane; e < GBOX_ELEMS; e += 32) box[e] = make_float2(0.f, 0.f);

    // per-lane constants
    const int j0l = lane / 6 > 5 ? 5 : lane / 6, j1l = lane % 6;                // rows 0..31
    const int rowoff = j0l * GPP + j1l * GRP;
    float2 E01c = cmul(g.E[0][j0l], g.E[1][j1l]);
    E01c.y = -E01c.y;                                                            // conj
    const int rr = lane / 6 > 3 ? 3 : lane / 6, j2r = lane % 6;                  // rows 32..35: (5, 2+rr), one element per lane
    const int remoff = 5 * GPP + (2 + rr) * GRP + j2r;
    float2 E01rc = cmul(g.E[0][5], g.E[1][2 + rr]);
    E01rc.y = -E01rc.y;
    __syncwarp();

    for (int s0i = wi.begin; s0i < wi.end; s0i += GBATCH) {
        const int ns = min(GBATCH, wi.end - s0i);
        // ---- expand records: [c0[6] c1[6] | b2[6] complex = c2*conj(E2)*conj(P)*y | base | pad3] ----
        if (lane < ns) {
            float4* R4 = reinterpret_cast<float4*>(srec + lane * GSRW);
            const float4* src = reinterpret_cast<const float4*>(rec + (long long)(s0i + lane) * RECW);
            const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3),
                         v4 = __ldg(src + 4), v5 = __ldg(src + 5);
            R4[0] = v0;
            R4[1] = v1;
            R4[2] = v2;
            const float c2[6] = {v3.x, v3.y, v3.z, v3.w, v4.x, v4.y};
            const float2 P = make_float2(v4.z, v4.w);
            const int m = __float_as_int(v5.w);
            const float2 yv = cmulc(P, y[(long long)m * nb + c]);                // conj(P) * y
            float2 b2[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float2 e = cmulc(g.E[2][j], yv);                                // conj(E2) * conj(P) * y
                b2[j] = make_float2(c2[j] * e.x, c2[j] * e.y);
            }
            R4[3] = make_float4(b2[0].x, b2[0].y, b2[1].x, b2[1].y);
            R4[4] = make_float4(b2[2].x, b2[2].y, b2[3].x, b2[3].y);
            R4[5] = make_float4(b2[4].x, b2[4].y, b2[5].x, b2[5].y);
            const int ks0 = __float_as_int(v5.x), ks1 = __float_as_int(v5.y), ks2 = __float_as_int(v5.z);
            const int base = (ks0 - O0) * GPP + (ks1 - O1) * GRP + (ks2 - O2);
            R4[6] = make_float4(__int_as_float(base), 0.f, 0.f, 0.f);
        }
        __syncwarp();
        for (int u = 0; u < ns; ++u) {
            const float* R = srec + u * GSRW;
            const int base = __float_as_int(R[24]);
            const float w01 = R[j0l] * R[6 + j1l];
            const float4 B0 = *reinterpret_cast<const float4*>(R + 12);
            const float4 B1 = *reinterpret_cast<const float4*>(R + 16);
            const float4 B2 = *reinterpret_cast<const float4*>(R + 20);
            const float2 wl = make_float2(E01c.x * w01, E01c.y * w01);
            float2* tp = box + base + rowoff;
            rmw(tp + 0, make_float2(B0.x, B0.y), wl);
            rmw(tp + 1, make_float2(B0.z, B0.w), wl);
            rmw(tp + 2, make_float2(B1.x, B1.y), wl);
            rmw(tp + 3, make_float2(B1.z, B1.w), wl);
            rmw(tp + 4, make_float2(B2.x, B2.y), wl);
            rmw(tp + 5, make_float2(B2.z, B2.w), wl);
            if (lane < 24) {                                                     // rows 32..35
                const float wr = R[5] * R[6 + 2 + rr];
                const float2 br = *reinterpret_cast<const float2*>(R + 12 + 2 * j2r);
                rmw(box + base + remoff, br, make_float2(E01rc.x * wr, E01rc.y * wr));
            }
            __syncwarp();
        }
    }

    // ---- flush the box (periodic) ----
    {
        const int K0 = g.K[0], K1 = g.K[1], K2 = g.K[2];
        float2* gc = grid + (long long)c * g.Kprod;
        for (int e = lane; e < GBOX * GBOX * GBOX; e += 32) {
            int p = e / (GBOX * GBOX);
            int rem = e - p * (GBOX * GBOX);
            int r = rem / GBOX;
            int cc = rem - r * GBOX;
            const float2 v = box[p * GPP + r * GRP + cc];
            if (v.x != 0.f || v.y != 0.f) {
                int i0 = O0 + p, i1 = O1 + r, i2 = O2 + cc;
                while (i0 >= K0) i0 -= K0;
                while (i1 >= K1) i1 -= K1;
                while (i2 >= K2) i2 -= K2;
                atomicAdd(gc + ((long long)i0 * K1 + i1) * K2 + i2, v);
            }
        }
    }
}

}  // namespace

int gridding_tiled_launch(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)GSMEM_BYTES));
        configured = true;
    }
    if (p->n_gwork == 0) return B200_OK;
    dim3 gr(p->n_gwork, nb);
    k_gridding_tiled<<<gr, 32, GSMEM_BYTES, st>>>(p->g, p->d_gwork, p->d_rec, y, grid, nb);
    LAUNCH_CHECK();
    return B200_OK;
}
