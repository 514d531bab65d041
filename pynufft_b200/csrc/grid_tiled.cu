// Tiled gridding (scatter) kernel for 3-D, J = 6: the replacement of pELL_spmvh_mCoil +
// atomic_add_float2 (src/re_subroutine.py:527-596, 275-287) on the headline configuration.
//
// Shared-memory float atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN), so the accumulation is
// made conflict-free by construction instead:
//  * a pre-pass gathers y into bin-sorted order (ys[i] = y[perm[i]]), so that everything the
//    main kernel reads is contiguous;
//  * one WARP per work item = (8^3 sub-tile of first-neighbour cells, range of bin-sorted samples).
//    The warp owns a private 13^3 accumulation box in shared memory (row pitch 13, plane pitch 174
//    complex -> the 16 rows of a half-warp phase hit 16 different bank pairs);
//  * sample records and sorted data arrive in 16-sample chunks by TMA bulk copies
//    (cp.async.bulk + mbarrier), double buffered, and are fixed up in place by 16 lanes;
//  * samples are processed one at a time: lane l owns footprint row (j0, j1) = divmod(l, 6) and does
//    a plain read-modify-write of its 6 contiguous elements (LDS.64 / 2 FFMA / STS.64 each: the
//    last-dimension phase is factored out of the box, so the last-dim weights are real); the
//    remaining rows 32..35 (24 elements) take one element per lane.  All lanes of one step touch
//    distinct addresses and no other warp shares the box -> no atomics, no races;
//  * when the range is done the box is multiplied by the last-dimension phase conj(Fl[column]) and
//    flushed once to the global grid with vector reductions (REDG.E.ADD.F32x2), because the halos
//    of neighbouring boxes overlap.
// The 2*M*prod(J) global float atomics of the reference become ~4.3 * prod(Kd) vector REDs.
#include "common.cuh"

namespace {

constexpr int GJ = 6;
constexpr int GT = 8;
constexpr int GBOX = GT + GJ - 1;        // 13
constexpr int GRP = 13;                  // row pitch (complex), odd
constexpr int GPP = 174;                 // plane pitch >= 13*13, == 6*GRP (mod 16)
constexpr int GBOX_ELEMS = GBOX * GPP;   // 2262
constexpr int GB = 16;                   // samples per chunk
constexpr int RECW = 24;
constexpr int CHUNK_BYTES = GB * RECW * 4 + GB * 16;      // records + sorted data (16-byte slots: TMA alignment)
constexpr size_t GSMEM_BYTES = GBOX_ELEMS * sizeof(float2) + 2 * CHUNK_BYTES + 16;

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

__device__ __forceinline__ float2 lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

__device__ __forceinline__ int wrap2(int i, int K) {
    i -= (i >= K) ? K : 0;
    i -= (i >= K) ? K : 0;
    return i;
}

// ys[c][i] = (y[perm[i], c], 0, 0)   -- one 16-byte slot per sample so that any range is TMA-aligned
__global__ void k_gather_sorted(const int* __restrict__ perm, long long M,
                                const float2* __restrict__ y, float4* __restrict__ ys, int nb) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int c = blockIdx.y;
    const int m = perm[i];
    const float2 v = y[(long long)m * nb + c];
    ys[(long long)c * M + i] = make_float4(v.x, v.y, 0.f, 0.f);
}

struct SampleRegs {
    float4 C0;      // c2[0..3]
    float4 C1Y;     // c2[4], c2[5], Y'.re, Y'.im
    int base;
    float w01;      // rows 0..31
    float wr;       // rows 32..35: c0[5] * c1[2+rr] * c2[j2r]
};

__global__ void __launch_bounds__(32)
k_gridding_tiled(Geom g, const WorkItem* __restrict__ work, const float* __restrict__ rec,
                 const float4* __restrict__ ys, long long M, float2* __restrict__ grid) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);
    unsigned char* cbuf = smem_raw + GBOX_ELEMS * sizeof(float2);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(cbuf + 2 * CHUNK_BYTES);

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
    const float4* ysc = ys + (long long)c * M;

    const int nchunks = (wi.end - wi.begin + GB - 1) / GB;
    auto issue = [&](int k) {      // lane 0 only
        const int s = wi.begin + k * GB;
        const int ns = min(GB, wi.end - s);
        unsigned char* dst = cbuf + (k & 1) * CHUNK_BYTES;
        mbar_expect(&mbar[k & 1], (unsigned)(ns * (RECW * 4 + 16)));
        tma_bulk(dst, rec + (long long)s * RECW, (unsigned)(ns * RECW * 4), &mbar[k & 1]);
        tma_bulk(dst + GB * RECW * 4, ysc + s, (unsigned)(ns * 16), &mbar[k & 1]);
    };
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) issue(0);

    {   // zero the box (16-byte stores; GBOX_ELEMS is even)
        float4* b4 = reinterpret_cast<float4*>(box);
        for (int e = lane; e < GBOX_ELEMS / 2; e += 32) b4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // per-lane constants
    const int j0l = lane / 6 > 5 ? 5 : lane / 6, j1l = lane % 6;                // rows 0..31
    const int rowoff = j0l * GPP + j1l * GRP;
    float2 E01c = cmul(g.E[0][j0l], g.E[1][j1l]);
    E01c.y = -E01c.y;                                                            // conj
    const int rr = lane / 6 > 3 ? 3 : lane / 6, j2r = lane % 6;                  // rows 32..35: (5, 2+rr), one element per lane
    const int remoff = 5 * GPP + (2 + rr) * GRP + j2r;
    float2 E01rc = cmul(g.E[0][5], g.E[1][2 + rr]);
    E01rc.y = -E01rc.y;
    const bool remlane = lane < 24;
    const unsigned box_s = smem_u32(box);
    __syncwarp();

    for (int k = 0; k < nchunks; ++k) {
        const int ns = min(GB, wi.end - (wi.begin + k * GB));
        if (k + 1 < nchunks) {
            fence_proxy_async();         // generic-proxy writes (fix-up of chunk k-1) before the async-proxy overwrite
            __syncwarp();
            if (lane == 0) issue(k + 1);
        }
        mbar_wait(&mbar[k & 1], (unsigned)((k >> 1) & 1));
        float* srec = reinterpret_cast<float*>(cbuf + (k & 1) * CHUNK_BYTES);
        // ---- in-place fix-up, one sample per lane: [.. c2[4] c2[5] | P | ks0 ks1 ks2 perm] ->
        //      [.. c2[4] c2[5] | Y'.re Y'.im | base ...],  Y' = conj(P * Gl[rel2]) * y
        if (lane < ns) {
            float* R = srec + lane * RECW;
            const float2 Pr = *reinterpret_cast<const float2*>(R + 18);
            const float4 v5 = *reinterpret_cast<const float4*>(R + 20);
            const float2 yv = *reinterpret_cast<const float2*>(srec + GB * RECW + 4 * lane);
            const int ks0 = __float_as_int(v5.x), ks1 = __float_as_int(v5.y), ks2 = __float_as_int(v5.z);
            const int base = (ks0 - O0) * GPP + (ks1 - O1) * GRP + (ks2 - O2);
            const float2 Pp = cmul(Pr, g.Gl[ks2 - O2]);
            const float2 Yp = cmulc(Pp, yv);
            *reinterpret_cast<float2*>(R + 18) = Yp;
            R[20] = __int_as_float(base);
        }
        __syncwarp();

        auto load = [&](int u) {
            SampleRegs s;
            const float* R = srec + u * RECW;
            s.C0 = *reinterpret_cast<const float4*>(R + 12);
            s.C1Y = *reinterpret_cast<const float4*>(R + 16);
            s.base = __float_as_int(R[20]);
            s.w01 = R[j0l] * R[6 + j1l];
            s.wr = R[5] * R[6 + 2 + rr] * R[12 + j2r];
            return s;
        };
        auto accumulate = [&](const SampleRegs& cur) {
            const float2 Yp = make_float2(cur.C1Y.z, cur.C1Y.w);
            float2 wl = cmul(E01c, Yp);
            wl.x *= cur.w01;
            wl.y *= cur.w01;
            const unsigned a = box_s + (unsigned)(cur.base + rowoff) * 8u;
            float2 v0 = lds64(a), v1 = lds64(a + 8), v2 = lds64(a + 16), v3 = lds64(a + 24), v4 = lds64(a + 32),
                   v5 = lds64(a + 40);
            v0.x = fmaf(cur.C0.x, wl.x, v0.x); v0.y = fmaf(cur.C0.x, wl.y, v0.y);
            v1.x = fmaf(cur.C0.y, wl.x, v1.x); v1.y = fmaf(cur.C0.y, wl.y, v1.y);
            v2.x = fmaf(cur.C0.z, wl.x, v2.x); v2.y = fmaf(cur.C0.z, wl.y, v2.y);
            v3.x = fmaf(cur.C0.w, wl.x, v3.x); v3.y = fmaf(cur.C0.w, wl.y, v3.y);
            v4.x = fmaf(cur.C1Y.x, wl.x, v4.x); v4.y = fmaf(cur.C1Y.x, wl.y, v4.y);
            v5.x = fmaf(cur.C1Y.y, wl.x, v5.x); v5.y = fmaf(cur.C1Y.y, wl.y, v5.y);
            sts64(a, v0); sts64(a + 8, v1); sts64(a + 16, v2); sts64(a + 24, v3); sts64(a + 32, v4); sts64(a + 40, v5);
            if (remlane) {                                // rows 32..35, one element per lane
                const float2 wlr = cmul(E01rc, Yp);
                const unsigned e = box_s + (unsigned)(cur.base + remoff) * 8u;
                float2 v = lds64(e);
                v.x = fmaf(cur.wr, wlr.x, v.x);
                v.y = fmaf(cur.wr, wlr.y, v.y);
                sts64(e, v);
            }
            __syncwarp();
        };
        // two samples per trip, registers ping-pong (the next sample's weights load while this one accumulates)
        SampleRegs ra = load(0), rb;
        int u = 0;
        for (; u + 2 <= ns; u += 2) {
            rb = load(u + 1);
            accumulate(ra);
            if (u + 2 < ns) ra = load(u + 2);
            accumulate(rb);
        }
        if (u < ns) accumulate(ra);
    }

    // ---- flush: box * conj(Fl[column]) -> global grid (periodic), vector REDs ----
    {
        const int K0 = g.K[0], K1 = g.K[1], K2 = g.K[2];
        const int KK = K1 * K2;
        float2* gc = grid + (long long)c * g.Kprod;
        // lane -> (row parity h, column cc): 26 lanes busy, two rows per step
        const int h = lane >= GBOX ? 1 : 0;
        const int cc = lane - h * GBOX;
        const bool act = lane < 2 * GBOX;
        float2 Fc = g.Fl[act ? cc : 0];
        Fc.y = -Fc.y;
        const int i2 = wrap2(O2 + (act ? cc : 0), K2);
        int goff[7];                      // global row offsets of rows h, h+2, ..
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int r = h + 2 * q;
            goff[q] = (act && r < GBOX) ? wrap2(O1 + (r < GBOX ? r : 0), K1) * K2 + i2 : -1;
        }
        const unsigned lane_s = box_s + (unsigned)(h * GRP + cc) * 8u;
        for (int p = 0; p < GBOX; ++p) {
            float2* gp = gc + (unsigned)(wrap2(O0 + p, K0) * KK);
            const unsigned ps = lane_s + (unsigned)(p * GPP) * 8u;
            float2 v[7];
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (goff[q] >= 0) v[q] = lds64(ps + q * 2 * GRP * 8);
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (goff[q] >= 0) atomicAdd(gp + goff[q], cmul(v[q], Fc));
        }
    }
}

}  // namespace

int gridding_tiled_launch(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    if (!p->attr_grid) {      // once per plan (the attribute is per device, plans are per device)
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)GSMEM_BYTES));
        p->attr_grid = true;
    }
    if (p->n_gwork == 0) return B200_OK;
    // sorted copy of the data (plan-owned scratch)
    if (p->ys_nb < nb) {
        if (p->d_ys) { CUDA_TRY(cudaFree(p->d_ys)); p->d_ys = nullptr; p->ys_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ys, sizeof(float4) * p->M * nb));
        p->ys_nb = nb;
    }
    {
        const int TB = 256;
        dim3 gr((unsigned)((p->M + TB - 1) / TB), nb);
        k_gather_sorted<<<gr, TB, 0, st>>>(p->d_perm, p->M, y, p->d_ys, nb);
        LAUNCH_CHECK();
    }
    dim3 gr(p->n_gwork, nb);
    k_gridding_tiled<<<gr, 32, GSMEM_BYTES, st>>>(p->g, p->d_gwork, p->d_rec, p->d_ys, p->M, grid);
    LAUNCH_CHECK();
    return B200_OK;
}
