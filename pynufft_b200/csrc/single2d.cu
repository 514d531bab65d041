// 2-D kernels for fewer than 8 coils (configuration 1: 256^2 / 512^2, PROPELLER, one coil), J = 6.
// One work item = (8 x 32 slab of first-neighbour cells, sample range); its 13 x 37 box of grid values fits in
// 3.8 KB of shared memory and is stored already multiplied by the separable phase F0[row]*F1[col], so all
// interpolation weights are real.
//
//   k_interp2d_s   replaces pELL_spmv_mCoil (src/re_subroutine.py:751-835): the 36 neighbours of a sample are
//                  too few to spread over a warp, so ONE THREAD takes one sample: 36 x (LDS.64 + 2 FFMA) from
//                  the box, record read straight from global memory (80 B, 5 x LDG.128).
//   k_gridding2d_s replaces pELL_spmvh_mCoil + atomic_add_float2 (:527-596, :275-287): one WARP per 8 x 16
//                  sub-tile (<= 256 samples) owns a private 13 x 21 accumulation box; a sample's 36 elements go to lanes 0..31 (+4 in a second step), plain
//                  LDS/FFMA/STS, no atomics; records and data of 32 samples are loaded one chunk ahead, one
//                  sample per lane; one RED flush per box.
#include "common.cuh"

namespace {

constexpr int SJ = 6;
constexpr int SR = 8 + SJ - 1;            // 13 box rows (slab of 8)
constexpr int SCMAX = 32 + SJ - 1;        // 37 box columns (tile edge <= 32)
constexpr int RECW2 = 20;                 // c0[6] c1[6] P(2) ks0 ks1 perm pad3
constexpr int ITHREADS = 128;
constexpr int GC = 16 + SJ - 1;           // 21 columns of the gridding box (8 x 16 sub-tile)

__device__ __forceinline__ int wrap2(int i, int K) {
    i -= (i >= K) ? K : 0;
    i -= (i >= K) ? K : 0;
    return i;
}

struct Slab {
    int O0, O1, ncol;
};

__device__ __forceinline__ Slab slab_of(const Geom& g, const WorkItem& wi) {
    Slab s;
    const int q1 = wi.tile % g.ntile[1], q0 = wi.tile / g.ntile[1];
    s.O0 = q0 * g.tile[0] + wi.pad;
    s.O1 = q1 * g.tile[1];
    s.ncol = g.tile[1] + SJ - 1;
    return s;
}

__global__ void __launch_bounds__(ITHREADS)
k_interp2d_s(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
             const float2* __restrict__ grid, float2* __restrict__ y, int nb) {
    __shared__ float2 box[SR * SCMAX];
    const WorkItem wi = work[blockIdx.x];
    const int c = blockIdx.y;
    const Slab sl = slab_of(g, wi);
    const float2* gc = grid + (long long)c * g.Kprod;
    for (int e = threadIdx.x; e < SR * sl.ncol; e += ITHREADS) {
        const int p = e / sl.ncol, cc = e - p * sl.ncol;
        const int i0 = wrap2(sl.O0 + p, g.K[0]), i1 = wrap2(sl.O1 + cc, g.K[1]);
        box[p * SCMAX + cc] = cmul(__ldg(gc + (long long)i0 * g.K[1] + i1), cmul(g.F0[p], g.Fl[cc]));
    }
    __syncthreads();
    for (int i = wi.begin + threadIdx.x; i < wi.end; i += ITHREADS) {
        const float4* R = reinterpret_cast<const float4*>(rec + (long long)i * RECW2);
        const float4 a0 = __ldg(R), a1 = __ldg(R + 1), a2 = __ldg(R + 2), a3 = __ldg(R + 3), a4 = __ldg(R + 4);
        const float c0[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
        const float c1[6] = {a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
        const int rel0 = __float_as_int(a3.z) - sl.O0, rel1 = __float_as_int(a3.w) - sl.O1;
        const float2 Pp = cmul(cmul(make_float2(a3.x, a3.y), g.G0[rel0]), g.Gl[rel1]);
        const float2* bp = box + rel0 * SCMAX + rel1;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j0 = 0; j0 < 6; ++j0) {
            float2 rs = make_float2(0.f, 0.f);
#pragma unroll
            for (int j1 = 0; j1 < 6; ++j1) {
                const float2 kv = bp[j0 * SCMAX + j1];
                rs.x = fmaf(c1[j1], kv.x, rs.x);
                rs.y = fmaf(c1[j1], kv.y, rs.y);
            }
            acc.x = fmaf(c0[j0], rs.x, acc.x);
            acc.y = fmaf(c0[j0], rs.y, acc.y);
        }
        y[(long long)__float_as_int(a4.x) * nb + c] = cmul(Pp, acc);
    }
}

struct RecRegs {
    float4 a0, a1, a2, a3;
    float2 yv;
};

__global__ void __launch_bounds__(32)
k_gridding2d_s(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
               const float2* __restrict__ y, float2* __restrict__ grid, int nb) {
    __shared__ float2 box[SR * GC];
    __shared__ __align__(16) float srec[32 * 16];          // per sample: c0[6] c1[6] Y'(2) base pad
    const WorkItem wi = work[blockIdx.x];                   // one 8 x 16 sub-tile (bin id), <= 256 samples
    const int c = blockIdx.y;
    const int lane = threadIdx.x;
    Slab sl;
    {
        const int sb = wi.tile % g.nsubprod, tl = wi.tile / g.nsubprod;
        const int s1 = sb % g.nsub[1], s0 = sb / g.nsub[1];
        const int q1 = tl % g.ntile[1], q0 = tl / g.ntile[1];
        sl.O0 = q0 * g.tile[0] + s0 * 8;
        sl.O1 = q1 * g.tile[1] + s1 * g.sub[1];
        sl.ncol = g.sub[1] + SJ - 1;
    }
    for (int e = lane; e < SR * GC; e += 32) box[e] = make_float2(0.f, 0.f);
    // lane -> footprint element: step 1 elements 0..31, step 2 elements 32..35 (lanes 0..3)
    const int j0a = lane / 6, j1a = lane % 6;
    const int offa = j0a * GC + j1a;
    const int j1b = 2 + (lane & 3);                          // element 32 + lane -> (5, 2 + lane)
    const int offb = 5 * GC + j1b;
    auto fetch = [&](int s0) {
        RecRegs r;
        const int i = s0 + lane;
        if (i < wi.end) {
            const float4* R = reinterpret_cast<const float4*>(rec + (long long)i * RECW2);
            r.a0 = __ldg(R); r.a1 = __ldg(R + 1); r.a2 = __ldg(R + 2); r.a3 = __ldg(R + 3);
            const int m = __float_as_int(__ldg(rec + (long long)i * RECW2 + 16));
            r.yv = __ldg(y + (long long)m * nb + c);
        } else {
            r.a0 = r.a1 = r.a2 = r.a3 = make_float4(0.f, 0.f, 0.f, 0.f);
            r.yv = make_float2(0.f, 0.f);
        }
        return r;
    };
    RecRegs nxt = fetch(wi.begin);
    __syncwarp();
    for (int s0 = wi.begin; s0 < wi.end; s0 += 32) {
        const RecRegs cur = nxt;
        const int ns = min(32, wi.end - s0);
        if (s0 + 32 < wi.end) nxt = fetch(s0 + 32);         // next chunk's loads stay in flight during this chunk
        {   // publish this lane's sample: real factors, Y' = conj(P') y, box offset
            float4* S = reinterpret_cast<float4*>(srec + lane * 16);
            S[0] = cur.a0;
            S[1] = cur.a1;
            S[2] = cur.a2;
            int base = 0;
            float2 Yp = make_float2(0.f, 0.f);
            if (lane < ns) {
                const int rel0 = __float_as_int(cur.a3.z) - sl.O0, rel1 = __float_as_int(cur.a3.w) - sl.O1;
                const float2 Pp = cmul(cmul(make_float2(cur.a3.x, cur.a3.y), g.G0[rel0]), g.Gl[rel1]);
                Yp = cmulc(Pp, cur.yv);
                base = rel0 * GC + rel1;
            }
            S[3] = make_float4(Yp.x, Yp.y, __int_as_float(base), 0.f);
        }
        __syncwarp();
        for (int u = 0; u < ns; ++u) {
            const float* S = srec + u * 16;
            const float4 t = *reinterpret_cast<const float4*>(S + 12);       // Y'.re Y'.im base
            const int base = __float_as_int(t.z);
            const float wa = S[j0a < 6 ? j0a : 5] * S[6 + j1a];
            if (lane < 32 && j0a < 6) {
                float2 v = box[base + offa];
                v.x = fmaf(wa, t.x, v.x);
                v.y = fmaf(wa, t.y, v.y);
                box[base + offa] = v;
            }
            if (lane < 4) {
                const float wb = S[5] * S[6 + j1b];
                float2 v = box[base + offb];
                v.x = fmaf(wb, t.x, v.x);
                v.y = fmaf(wb, t.y, v.y);
                box[base + offb] = v;
            }
            __syncwarp();
        }
    }
    // flush: box * conj(F0[row]*F1[col]) -> grid (periodic)
    float2* gc = grid + (long long)c * g.Kprod;
    for (int e = lane; e < SR * sl.ncol; e += 32) {
        const int p = e / sl.ncol, cc = e - p * sl.ncol;
        const float2 v = box[p * GC + cc];
        if (v.x != 0.f || v.y != 0.f) {
            const int i0 = wrap2(sl.O0 + p, g.K[0]), i1 = wrap2(sl.O1 + cc, g.K[1]);
            atomicAdd(gc + (long long)i0 * g.K[1] + i1, cmulc(cmul(g.F0[p], g.Fl[cc]), v));
        }
    }
}

}  // namespace

bool single2d_supported(const Geom& g) {
    return g.ndim == 2 && g.J[0] == SJ && g.J[1] == SJ && g.sub[0] == 8 && g.tile[0] % 8 == 0 && g.tile[1] <= 32 &&
           g.sub[1] <= 16 && g.recw == RECW2;
}

int single2d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st) {
    if (p->n_work == 0) return B200_OK;
    dim3 gr(p->n_work, nb);
    k_interp2d_s<<<gr, ITHREADS, 0, st>>>(p->g, p->d_work, p->d_rec, grid, y, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

int single2d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    if (p->n_gwork == 0) return B200_OK;
    dim3 gr(p->n_gwork, nb);
    k_gridding2d_s<<<gr, 32, 0, st>>>(p->g, p->d_gwork, p->d_rec, y, grid, nb);
    LAUNCH_CHECK();
    return B200_OK;
}
