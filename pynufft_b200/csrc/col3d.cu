// Column-sweep interpolation / gridding kernels for 3-D, J = 6: the register-resident replacement of
// pELL_spmv_mCoil / pELL_spmvh_mCoil + atomic_add_float2 (src/re_subroutine.py:751-835, 527-596, 275-287).
//
// The shared-memory tiled kernels (interp_tiled.cu / grid_tiled.cu) are bound by the shared-memory pipe: every one
// of the 216 neighbours of a sample is an 8-byte LDS (gather) or an LDS + STS (scatter).  Here the grid values a
// warp works on live in REGISTERS and no neighbour ever goes through shared memory:
//
//  * samples are sorted by (column, first plane): a column is a 5 x 4 cross-section (dims 1, 2) of first-neighbour
//    cells, swept along dim 0 in increasing plane order (the sort contract: key = column * K0 + first plane);
//  * the 6 x 6 x 6 footprint of a sample covers, in dim 0, exactly one plane of every residue class mod 6, and in
//    dim 1 one or two rows of every residue class mod 5.  Lane (r, t) = (plane class, row class), 30 lanes, owns the
//    two box rows b = t, t + 5 (9 columns each: 4 cells + 5 halo) of the one plane p = r (mod 6) inside the current
//    6-plane window: 18 complex registers.  Because samples arrive in plane order the window only slides forward:
//    when it leaves plane p the five lanes of that class reload (gather) or flush (scatter) their rows once;
//  * the inner loop is branch-free: the record holds the last-dimension weights already shifted to box columns
//    and zero-padded (9 values, three warp-uniform LDS.128 together with plane / offsets), and c1 for both row
//    slots comes from a zero-padded table, so every sample is 36 FFMA on static registers.  The gather reduces
//    the 30 partial sums through a 16-sample shared-memory transpose; the scatter needs no reduction at all;
//  * all weights are REAL: the grid handed to these kernels is phase-modulated, G'[g] = G[g] * prod_d e^{i s_d g_d}
//    (s_d = gamma_d (N_d - 1) / 2), which turns the reference's complex min-max coefficients
//    u_j = c_j e^{i om N/2} e^{-i s (dk - j)} (helper.py:148-162, 606-618) into c_j times ONE phase per sample.  The
//    modulation is applied by the FFT passes (fft256.cu) or by k_modulate below; neighbours that wrap around the
//    periodic grid pick up the sign e^{i s K} = (-1)^(N-1);
//  * sample records (144 B, precomputed at plan time in sweep order) arrive by double-buffered TMA bulk copies;
//    warps are persistent and fetch work items (segments of a column) from a per-coil counter.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int CT1 = COL_T1;               // column cross-section, dim 1 (first-neighbour cells): 5
constexpr int CT2 = COL_T2;               // column cross-section, dim 2: 4
constexpr int CL = CT2 + 5;               // 9 box columns
constexpr int CRECW = COL_RECW;           // words per record: 36
constexpr int CCH = 16;                   // samples per chunk
constexpr int CWARPS = 4;                 // warps per CTA (each warp works on its own items)
constexpr int REC_BYTES = CCH * CRECW * 4;            // 2304
constexpr int PBUF_PITCH = 33;
constexpr int PBUF_BYTES = CCH * PBUF_PITCH * 8;      // 4224
constexpr int STG_BYTES = 2 * 5 * 512;                // staging: 2 rows x 5 pieces x (32 lanes x 16 B)
constexpr int IWARP_BYTES = 14080;                    // 2 * REC + PBUF + STG + mbar, rounded to 128
constexpr int YS_BYTES = CCH * 16;
constexpr int GWARP_BYTES = 5248;                     // 2 * REC + 2 * YS + mbar, rounded to 128
static_assert(CT2 == 4 && CRECW == 36, "record layout below assumes 5 x 4 columns");

// record words: [w9[0..8] | p0 | info | perm | P''.re P''.im | - - | c0[0..5] | 0 0 0 0 | c1[0..5] | 0 0 0 0]
//                0          9    10     11     12              14    16         22        26         32
// w9[c] = c2[c - k2rel] (0 outside the footprint); info = 4 (26 - k1rel) (byte offset of c1[-k1rel]) | (p0 mod 6) << 8

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

__device__ __forceinline__ void red_v4(float2* addr, float2 a, float2 b) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
                 : "memory");
}
__device__ __forceinline__ void red_v2(float2* addr, float2 a) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a.x), "f"(a.y) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

// geometry of one warp's column: which global rows / columns its two box rows map to, and the wrap signs.
// Box columns 0..3 (segment A), 4..7 (B), 8 (C): K2 % 4 == 0, so each segment is contiguous in memory.
struct LaneGeom {
    int off0, off1;       // element offset of (row, first column of segment A) inside a plane, rows b = t and t + 5
    int dB, dC;           // element offsets of segments B and C relative to segment A
    float s0, s1, sB, sC; // wrap signs of row 0, row 1, segment B, segment C
    bool wraps;           // any of them != 1
    bool colwrap;         // segments B / C are not contiguous with A (last columns of the grid): warp-uniform
};

__device__ __forceinline__ LaneGeom lane_geom(const ColGeom& g, int q1, int q2, int t) {
    LaneGeom L;
    int row0 = q1 * CT1 + t, row1 = row0 + CT1;
    L.s0 = 1.f; L.s1 = 1.f; L.sB = 1.f; L.sC = 1.f;
    if (row0 >= g.K1) { row0 -= g.K1; L.s0 = g.sg1; }
    if (row1 >= g.K1) { row1 -= g.K1; L.s1 = g.sg1; }
    const int colA = q2 * CT2;
    int colB = colA + 4, colC = colA + 8;
    if (colB >= g.K2) { colB -= g.K2; L.sB = g.sg2; }
    if (colC >= g.K2) { colC -= g.K2; L.sC = g.sg2; }
    L.off0 = row0 * g.K2 + colA;
    L.off1 = row1 * g.K2 + colA;
    L.dB = colB - colA;
    L.dC = colC - colA;
    L.wraps = (L.s0 != 1.f) || (L.s1 != 1.f) || (L.sB != 1.f) || (L.sC != 1.f);
    L.colwrap = (L.dB != 4) || (L.dC != 8);
    return L;
}

__device__ __forceinline__ void load_row(const float2* __restrict__ rowA, int dB, int dC, float2 (&D)[CL]) {
    const float4* a = reinterpret_cast<const float4*>(rowA);
    const float4* b = reinterpret_cast<const float4*>(rowA + dB);
    const float4 v0 = __ldg(a), v1 = __ldg(a + 1), v2 = __ldg(b), v3 = __ldg(b + 1);
    const float2 v4 = __ldg(rowA + dC);
    D[0] = make_float2(v0.x, v0.y); D[1] = make_float2(v0.z, v0.w);
    D[2] = make_float2(v1.x, v1.y); D[3] = make_float2(v1.z, v1.w);
    D[4] = make_float2(v2.x, v2.y); D[5] = make_float2(v2.z, v2.w);
    D[6] = make_float2(v3.x, v3.y); D[7] = make_float2(v3.z, v3.w);
    D[8] = v4;
}

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_4() { asm volatile("cp.async.wait_group 4;\n" ::: "memory"); }

// global row (9 values in segments A, B, C) -> this lane's staging slot, asynchronously
// (piece-major layout: piece i of every lane is contiguous, 16 B per lane -> the five lanes that move together touch
//  80 contiguous bytes per instruction, one shared-memory wavefront)
__device__ __forceinline__ void stage_row(unsigned dst, const float2* __restrict__ rowA, int dB, int dC) {
    cp_async16(dst, rowA);
    cp_async16(dst + 512, rowA + 2);
    cp_async16(dst + 1024, rowA + dB);
    cp_async16(dst + 1536, rowA + dB + 2);
    cp_async8(dst + 2048, rowA + dC);
}
__device__ __forceinline__ void unstage_row(const unsigned char* src, float2 (&D)[CL]) {
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 512);
    const float4 v2 = *reinterpret_cast<const float4*>(src + 1024), v3 = *reinterpret_cast<const float4*>(src + 1536);
    const float2 v4 = *reinterpret_cast<const float2*>(src + 2048);
    D[0] = make_float2(v0.x, v0.y); D[1] = make_float2(v0.z, v0.w);
    D[2] = make_float2(v1.x, v1.y); D[3] = make_float2(v1.z, v1.w);
    D[4] = make_float2(v2.x, v2.y); D[5] = make_float2(v2.z, v2.w);
    D[6] = make_float2(v3.x, v3.y); D[7] = make_float2(v3.z, v3.w);
    D[8] = v4;
}

__device__ __forceinline__ void scale_row(float2 (&D)[CL], float fA, float fB, float fC) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { D[c].x *= fA; D[c].y *= fA; }
#pragma unroll
    for (int c = 4; c < 8; ++c) { D[c].x *= fB; D[c].y *= fB; }
    D[8].x *= fC; D[8].y *= fC;
}

__device__ __forceinline__ void flush_row(float2* __restrict__ rowA, int dB, int dC, const float2 (&D)[CL]) {
    red_v4(rowA, D[0], D[1]);
    red_v4(rowA + 2, D[2], D[3]);
    red_v4(rowA + dB, D[4], D[5]);
    red_v4(rowA + dB + 2, D[6], D[7]);
    red_v2(rowA + dC, D[8]);
}

// next work item of this warp (dynamic: one atomic per item, broadcast from lane 0)
__device__ __forceinline__ int next_item(int* counter, int lane) {
    int it = 0;
    if (lane == 0) it = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
}

// ---------------------------------------------------------------------------------------------------------
// interpolation (gather)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CWARPS * 32, 4)
k_interp_col(ColGeom g, const WorkItem* __restrict__ work, int n_work, int* __restrict__ counter,
             const float* __restrict__ rec, const float2* __restrict__ grid, float2* __restrict__ y, int nb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ws = smem_raw + warp * IWARP_BYTES;
    float2* pbuf = reinterpret_cast<float2*>(ws + 2 * REC_BYTES);
    unsigned char* stg = ws + 2 * REC_BYTES + PBUF_BYTES + lane * 16;           // this lane's staging slots
    const unsigned stg_s = smem_u32(stg);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + 2 * REC_BYTES + PBUF_BYTES + STG_BYTES);
    const int c = blockIdx.y;
    const float2* gc = grid + (long long)c * g.Kprod;
    const int r = lane / CT1, t = lane - CT1 * r;
    const bool active = lane < 30;
    const int KK = g.K1 * g.K2;
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
        const LaneGeom L = lane_geom(g, q1, q2, t);
        const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
        auto issue = [&](int k) {      // lane 0 only
            const int s = wi.begin + k * CCH;
            const int ns = min(CCH, wi.end - s);
            const unsigned b = (gk + k) & 1;
            mbar_expect(&mbar[b], (unsigned)(ns * CRECW * 4));
            tma_bulk(ws + b * REC_BYTES, rec + (long long)s * CRECW, (unsigned)(ns * CRECW * 4), &mbar[b]);
        };
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) issue(0);

        float2 D0[CL], D1[CL];
#pragma unroll
        for (int q = 0; q < CL; ++q) { D0[q] = make_float2(0.f, 0.f); D1[q] = make_float2(0.f, 0.f); }
        int a_cur = -1000, p0_prev = -1000, j0 = 0, staged = -1000;
        int ev = 0, ev_staged = 0;                      // commit groups of this warp so far / when `staged` was issued
        const float2* row0 = gc + L.off0;               // this lane's two rows in plane 0
        const float2* row1 = gc + L.off1;

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
#pragma unroll 2
            for (int u = 0; u < ns; ++u) {
                const float* R = Rb + u * CRECW;
                const float4 W0 = *reinterpret_cast<const float4*>(R);
                const float4 W1 = *reinterpret_cast<const float4*>(R + 4);
                const float4 W2 = *reinterpret_cast<const float4*>(R + 8);      // w8, p0, info, perm
                const int p0 = __float_as_int(W2.y);
                const int info = __float_as_int(W2.z);
                if (p0 != p0_prev) {                      // warp-uniform: the window moved
                    p0_prev = p0;
                    j0 = r - (info >> 8);
                    j0 += j0 < 0 ? 6 : 0;
                    const int need = p0 + j0;
                    if (active && need != a_cur) {
                        // this lane's plane left the window: take the next plane of its class.  It was staged in
                        // shared memory (cp.async) when the lane took the previous one, 6 window steps ago.
                        a_cur = need;
                        int pw = need;
                        float sp = 1.f;
                        if (pw >= g.K0) { pw -= g.K0; sp = g.sg0; }
                        // copy groups complete in order and every event commits exactly one (warp-wide) group:
                        // a plane staged at least 5 events ago is complete once at most 4 groups are pending
                        const bool old_enough = ev - ev_staged >= 5;
                        if (__all_sync(__activemask(), old_enough)) cp_async_wait_4(); else cp_async_wait_all();
                        if (staged == need) {
                            unstage_row(stg, D0);
                            unstage_row(stg + 2560, D1);
                        } else {
                            const long long po = (long long)pw * KK;
                            load_row(row0 + po, L.dB, L.dC, D0);
                            load_row(row1 + po, L.dB, L.dC, D1);
                        }
                        if (L.wraps || sp != 1.f) {
                            scale_row(D0, sp * L.s0, sp * L.s0 * L.sB, sp * L.s0 * L.sC);
                            scale_row(D1, sp * L.s1, sp * L.s1 * L.sB, sp * L.s1 * L.sC);
                        }
                        staged = need + 6;
                        int pn = pw + 6;
                        if (pn >= g.K0) pn -= g.K0;
                        const long long pno = (long long)pn * KK;
                        stage_row(stg_s, row0 + pno, L.dB, L.dC);
                        stage_row(stg_s + 2560, row1 + pno, L.dB, L.dC);
                        cp_async_commit();
                        ev_staged = ev;
                    }
                    ++ev;           // (uniform) one commit per event; events without a retiring lane commit nothing: harmless
                }
                const float c0 = R[16 + j0];
                const float* c1p = reinterpret_cast<const float*>(reinterpret_cast<const char*>(R) + (info & 0xff)) + t;
                const float c1a = c1p[0], c1b = c1p[5];
                const float w[CL] = {W0.x, W0.y, W0.z, W0.w, W1.x, W1.y, W1.z, W1.w, W2.x};
                float2 d0, d1;
                d0.x = w[0] * D0[0].x; d0.y = w[0] * D0[0].y;
                d1.x = w[0] * D1[0].x; d1.y = w[0] * D1[0].y;
#pragma unroll
                for (int q = 1; q < CL; ++q) {
                    d0.x = fmaf(w[q], D0[q].x, d0.x); d0.y = fmaf(w[q], D0[q].y, d0.y);
                    d1.x = fmaf(w[q], D1[q].x, d1.x); d1.y = fmaf(w[q], D1[q].y, d1.y);
                }
                // lanes 30, 31 hold zero rows (never loaded): their partial sums are exact zeros
                float2 part;
                part.x = c0 * fmaf(c1a, d0.x, c1b * d1.x);
                part.y = c0 * fmaf(c1a, d0.y, c1b * d1.y);
                pbuf[u * PBUF_PITCH + lane] = part;
            }
            // ---- reduce the 32 partial sums of every sample of the chunk: lane -> (sample s, half h) ----
            __syncwarp();
            {
                const int s = lane & 15, h = lane >> 4;
                const float2* pb = pbuf + s * PBUF_PITCH + h * 16;
                float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                for (int l = 0; l < 16; ++l) {
                    const float2 v = pb[l];
                    sum.x += v.x;
                    sum.y += v.y;
                }
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, 16);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, 16);
                if (lane < ns) {
                    const float* Rs = Rb + lane * CRECW;
                    const float2 P = *reinterpret_cast<const float2*>(Rs + 12);
                    const int m = __float_as_int(Rs[11]);
                    y[(long long)m * nb + c] = cmul(P, sum);
                }
            }
            __syncwarp();
        }
        gk += nchunks;
        cp_async_wait_all();                            // the last staged planes of this item are dropped
    }
}

// ---------------------------------------------------------------------------------------------------------
// gridding (scatter)
// ---------------------------------------------------------------------------------------------------------
// ys[c][i] = (conj(P''_i) * y[perm[i], c], 0, 0): 16-byte slots so that any sample range is TMA-aligned
__global__ void k_gather_sorted_col(const float* __restrict__ rec, long long M, const float2* __restrict__ y,
                                    float4* __restrict__ ys, int nb) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int c = blockIdx.y;
    const float4 h = __ldg(reinterpret_cast<const float4*>(rec + i * CRECW + 8));    // w8, p0, info, perm
    const float2 P = __ldg(reinterpret_cast<const float2*>(rec + i * CRECW + 12));
    const int m = __float_as_int(h.w);
    const float2 v = cmulc(P, y[(long long)m * nb + c]);
    ys[(long long)c * M + i] = make_float4(v.x, v.y, 0.f, 0.f);
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
    const int r = lane / CT1, t = lane - CT1 * r;
    const bool active = lane < 30;
    const int KK = g.K1 * g.K2;
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    unsigned gk = 0;

    for (int item = next_item(counter + c, lane); item < n_work; item = next_item(counter + c, lane)) {
        const WorkItem wi = work[item];
        const int q1 = wi.tile / g.nq2, q2 = wi.tile - q1 * g.nq2;
        const LaneGeom L = lane_geom(g, q1, q2, t);
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

        float2 A0[CL], A1[CL];
#pragma unroll
        for (int q = 0; q < CL; ++q) { A0[q] = make_float2(0.f, 0.f); A1[q] = make_float2(0.f, 0.f); }
        int a_cur = -1000, p0_prev = -1000, j0 = 0;

        float2* grow0 = gc + L.off0;
        float2* grow1 = gc + L.off1;
        auto flush = [&]() {           // rows of plane a_cur -> global grid (vector REDs), then clear
            int pw = a_cur;
            float sp = 1.f;
            if (pw >= g.K0) { pw -= g.K0; sp = g.sg0; }
            if (L.wraps || sp != 1.f) {
                scale_row(A0, sp * L.s0, sp * L.s0 * L.sB, sp * L.s0 * L.sC);
                scale_row(A1, sp * L.s1, sp * L.s1 * L.sB, sp * L.s1 * L.sC);
            }
            const long long po = (long long)pw * KK;
            flush_row(grow0 + po, L.dB, L.dC, A0);
            flush_row(grow1 + po, L.dB, L.dC, A1);
#pragma unroll
            for (int q = 0; q < CL; ++q) { A0[q] = make_float2(0.f, 0.f); A1[q] = make_float2(0.f, 0.f); }
        };

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
                const float4 W0 = *reinterpret_cast<const float4*>(R);
                const float4 W1 = *reinterpret_cast<const float4*>(R + 4);
                const float4 W2 = *reinterpret_cast<const float4*>(R + 8);      // w8, p0, info, perm
                const float2 wl = *reinterpret_cast<const float2*>(Y + u);      // conj(P'') * y
                const int p0 = __float_as_int(W2.y);
                const int info = __float_as_int(W2.z);
                if (p0 != p0_prev) {                      // warp-uniform: the window moved
                    p0_prev = p0;
                    j0 = r - (info >> 8);
                    j0 += j0 < 0 ? 6 : 0;
                    const int need = p0 + j0;
                    if (active && need != a_cur) {
                        if (a_cur >= 0) flush();
                        a_cur = need;
                    }
                }
                const float c0 = R[16 + j0];
                const float* c1p = reinterpret_cast<const float*>(reinterpret_cast<const char*>(R) + (info & 0xff)) + t;
                const float ta = c0 * c1p[0], tb = c0 * c1p[5];
                const float ax = ta * wl.x, ay = ta * wl.y, bx = tb * wl.x, by = tb * wl.y;
                const float w[CL] = {W0.x, W0.y, W0.z, W0.w, W1.x, W1.y, W1.z, W1.w, W2.x};
#pragma unroll
                for (int q = 0; q < CL; ++q) {
                    A0[q].x = fmaf(w[q], ax, A0[q].x); A0[q].y = fmaf(w[q], ay, A0[q].y);
                    A1[q].x = fmaf(w[q], bx, A1[q].x); A1[q].y = fmaf(w[q], by, A1[q].y);
                }
            }
            __syncwarp();
        }
        if (active && a_cur >= 0) flush();
        gk += nchunks;
    }
}

// ---------------------------------------------------------------------------------------------------------
// grid modulation: out[g] = in[g] * prod_d m_d[g_d]  (conj -> conjugated factors); one CTA row per (g0, g1)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_modulate(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ mod,
                           int K0, int K1, int K2, int conj_) {
    const int row = blockIdx.x;                        // g0 * K1 + g1
    const int g0 = row / K1, g1 = row - g0 * K1;
    const long long cb = (long long)blockIdx.y * K0 * K1 * K2;
    float2 m01 = cmul(__ldg(mod + g0), __ldg(mod + K0 + g1));
    const float2* m2 = mod + K0 + K1;
    const float2* src = in + cb + (long long)row * K2;
    float2* dst = out + cb + (long long)row * K2;
    for (int g2 = threadIdx.x; g2 < K2; g2 += blockDim.x) {
        float2 m = cmul(m01, __ldg(m2 + g2));
        if (conj_) m.y = -m.y;
        dst[g2] = cmul(src[g2], m);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool col3d_supported(const Geom& g) {
    if (g.ndim != 3) return false;
    for (int d = 0; d < 3; ++d)
        if (g.J[d] != 6) return false;
    return g.K[0] >= 6 && g.K[1] >= 10 && g.K[2] >= 12 && g.K[2] % 4 == 0;
}


static ColGeom col_geom(const Geom& g) {
    ColGeom c;
    c.K0 = g.K[0]; c.K1 = g.K[1]; c.K2 = g.K[2];
    c.nq2 = g.K[2] / CT2;
    c.sg0 = ((g.N[0] - 1) & 1) ? -1.f : 1.f;
    c.sg1 = ((g.N[1] - 1) & 1) ? -1.f : 1.f;
    c.sg2 = ((g.N[2] - 1) & 1) ? -1.f : 1.f;
    c.Kprod = g.Kprod;
    return c;
}

// per-coil work counters of the persistent kernels (zeroed on the stream before every launch)
static int col_counters(b200nufft_plan_t p, int nb, cudaStream_t st) {
    if (p->ccount_nb < nb) {
        if (p->d_ccount) { CUDA_TRY(cudaFree(p->d_ccount)); p->d_ccount = nullptr; p->ccount_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ccount, sizeof(int) * nb));
        p->ccount_nb = nb;
    }
    CUDA_TRY(cudaMemsetAsync(p->d_ccount, 0, sizeof(int) * nb, st));
    return B200_OK;
}
// CTAs per coil of the persistent kernels: all SMs x 5 resident CTAs, shared between the coils of the launch
static int col_ctas(b200nufft_plan_t p, int nb) {
    if (p->n_sm == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) p->n_sm = prop.multiProcessorCount;
        if (p->n_sm <= 0) p->n_sm = 148;
    }
    return std::max(1, (p->n_sm * 5 + nb - 1) / nb);
}

int col3d_modulate(b200nufft_plan_t p, const float2* in, float2* out, int nb, int conj_, cudaStream_t st) {
    const Geom& g = p->g;
    dim3 gr((unsigned)(g.K[0] * g.K[1]), nb);
    k_modulate<<<gr, 128, 0, st>>>(in, out, p->d_mod, g.K[0], g.K[1], g.K[2], conj_);
    LAUNCH_CHECK();
    return B200_OK;
}

// grid: phase-modulated grid (see the header comment)
int col3d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st) {
    if (!p->attr_col) {
        CUDA_TRY(cudaFuncSetAttribute(k_interp_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * IWARP_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * GWARP_BYTES));
        p->attr_col = true;
    }
    if (p->n_cwork == 0) return B200_OK;
    ARG_CHECK((reinterpret_cast<uintptr_t>(grid) & 15) == 0, "interp: grid must be 16-byte aligned");
    int rc = col_counters(p, nb, st);
    if (rc) return rc;
    dim3 gr((unsigned)std::min((p->n_cwork + CWARPS - 1) / CWARPS, col_ctas(p, nb)), nb);
    k_interp_col<<<gr, CWARPS * 32, CWARPS * IWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                               p->d_crec, grid, y, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

// grid (zero on entry) receives the phase-modulated adjoint
int col3d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st) {
    if (!p->attr_col) {
        CUDA_TRY(cudaFuncSetAttribute(k_interp_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * IWARP_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * GWARP_BYTES));
        p->attr_col = true;
    }
    if (p->n_cwork == 0) return B200_OK;
    ARG_CHECK((reinterpret_cast<uintptr_t>(grid) & 15) == 0, "gridding: grid must be 16-byte aligned");
    if (p->ys_nb < nb) {
        if (p->d_ys) { CUDA_TRY(cudaFree(p->d_ys)); p->d_ys = nullptr; p->ys_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ys, sizeof(float4) * p->M * nb));
        p->ys_nb = nb;
    }
    {
        const int TB = 256;
        dim3 gr((unsigned)((p->M + TB - 1) / TB), nb);
        k_gather_sorted_col<<<gr, TB, 0, st>>>(p->d_crec, p->M, y, p->d_ys, nb);
        LAUNCH_CHECK();
    }
    int rc = col_counters(p, nb, st);
    if (rc) return rc;
    dim3 gr((unsigned)std::min((p->n_cwork + CWARPS - 1) / CWARPS, col_ctas(p, nb)), nb);
    k_gridding_col<<<gr, CWARPS * 32, CWARPS * GWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                                 p->d_crec, p->d_ys, p->M, grid);
    LAUNCH_CHECK();
    return B200_OK;
}
