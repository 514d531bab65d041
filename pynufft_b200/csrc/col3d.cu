// Column-sweep interpolation (gather) and gridding (scatter) kernels for 3-D, J = 6: the register-resident replacement
// of pELL_spmv_mCoil and pELL_spmvh_mCoil + atomic_add_float2 (src/re_subroutine.py:751-835, 527-596, 275-287).
//
// The shared-memory tiled kernels (interp_tiled.cu, grid_tiled.cu) are bound by the shared-memory pipe: every one of the
// 216 neighbours of a sample is an LDS (gather) or an LDS + STS (scatter).  Here the grid cells a warp works on live in
// REGISTERS and the inner loop is packed FP32 math (FFMA2, two lanes of a complex number per instruction):
//
//  * a second copy of the samples is sorted by (column, first plane): a column is a 4 x 5 cross-section (dims 1, 2)
//    of first-neighbour cells, swept along dim 0 in increasing plane order (key = column * K0 + first plane);
//  * the footprint of a sample lies inside the 9 x 10 box of its column (4 + 5 rows, 5 + 5 columns).  Lane (g, c),
//    g = 0..2, c = 0..9 (30 lanes), owns the box cells (rows g, g + 3, g + 6; column c) of the planes of the current
//    window [p, p + 6).  Samples arrive in plane order, so the window only slides forward, one plane at a time;
//  * STATIC PLANE PHASES: the code is unrolled over the residue of the window's first plane (p mod 6), so that inside a
//    phase every accumulator / window slot is a fixed register and the sample's dim-0 weights c0[j] need no rotation.
//    A phase = "take every sample whose first plane is p, then retire plane p" and falls through into the next one; the
//    first plane and the remaining run length of the NEXT sample are carried in registers, so a plane without samples
//    costs no shared-memory load and no chunk test;
//  * scatter: the retired plane is flushed by ALL lanes at once, three 8-byte REDs per lane (running cell pointers) that
//    cover three 80-byte row segments per instruction, and its registers are cleared.  gather: the retired plane's
//    registers receive plane p + 6 from a per-warp shared-memory ring that 8-byte cp.async copies fill PRING planes
//    ahead of the window (each lane fetches and reads back its own three cells, so the per-thread cp.async groups are the
//    only synchronisation): the L2 / DRAM latency of a plane is covered without registers that hold loads in flight;
//    the 30 partial sums of a sample go through a 16-sample shared-memory transpose;
//  * a sample is 5 LDS + 22 packed FP32 instructions per lane: the record holds the dim-1 weights as a zero-padded
//    table over the box rows and the dim-2 weights zero-padded over the box columns, all REAL, because the grid these
//    kernels work on is phase-modulated, G'[g] = G[g] * prod_d e^{i s_d g_d} (s_d = gamma_d (N_d - 1) / 2), which turns
//    the reference's complex min-max coefficients u_j = c_j e^{i om N/2} e^{-i s (dk - j)} (helper.py:148-162, 606-618)
//    into c_j times ONE phase per sample.  Neighbours that wrap around the periodic grid pick up the sign
//    e^{i s K} = (-1)^(N-1): folded into the record's weights at plan time (plan.cu k_col_records);
//  * the modulation is applied / undone for free by the fused FFT passes (fft256.cu) or by k_modulate / k_demodulate;
//  * sample records (128 B, precomputed at plan time in sweep order) and the scatter's pre-gathered data arrive by
//    double-buffered TMA bulk copies; warps are persistent and fetch work items (segments of a column) from a per-coil
//    counter.
#include <algorithm>
#include <climits>
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
#ifndef COL_COILS_MAX
#define COL_COILS_MAX 4
#endif
constexpr int COL_COILS = COL_COILS_MAX;  // coils that share one launch of the persistent kernels
constexpr int REC_BYTES = CCH * CRECW * 4;            // 2048
constexpr int YS_BYTES = 160;                         // up to 18 pre-gathered values (8 B) per chunk, 16-byte multiple
constexpr int GWARP_BYTES = 4608;                     // scatter: 2 * REC + 2 * YS + 128 spare + mbar, rounded to 128
#ifndef COL_I_LOADS
#define COL_I_LOADS 1       // 0: skip the plane copies (timing experiments; wrong results)
#endif
#ifndef COL_I_CTAS
#define COL_I_CTAS 4
#endif
#ifndef COL_I_PRING
#define COL_I_PRING 4
#endif
#ifndef COL_I_PB
#define COL_I_PB 16
#endif
#ifndef COL_I_PAIR
#define COL_I_PAIR 1
#endif
constexpr int PRING = COL_I_PRING;                    // gather: planes in flight ahead of the window (shared-memory ring)
constexpr int PSLOT = 32 * CNR * 8;                   // gather: bytes per ring plane (3 cells per lane)
constexpr int PB = COL_I_PB;                          // gather: samples per reduction group (transpose buffer rows)
constexpr int PBUF_PITCH = 33;                        // gather: partial sums pbuf[sample mod PB][lane], float2
constexpr int PBUF_BYTES = PB * PBUF_PITCH * 8;
constexpr int SIDE_BYTES = CCH * 16;                   // gather: (P'', original index) of the chunk's samples
#ifndef COL_I_NBUF
#define COL_I_NBUF 2
#endif
constexpr int INB = COL_I_NBUF;                       // gather: record chunks in flight + the one in use
constexpr int IWARP_BYTES = (INB * (REC_BYTES + SIDE_BYTES) + PBUF_BYTES + PRING * PSLOT + 8 * INB + 127) / 128 * 128;   // per warp
static_assert(CT1 == 4 && CT2 == 5 && CRECW == 32 && CROWS == 3 * CNR, "record layout below assumes 4 x 5 columns");
static_assert(2 * REC_BYTES + 2 * YS_BYTES + 128 + 16 <= GWARP_BYTES, "per-warp shared memory (scatter)");
static_assert(INB >= 2 && INB <= 4 && IWARP_BYTES % 128 == 0 && PBUF_BYTES % 16 == 0, "per-warp shared memory (gather)");
static_assert(PRING >= 1 && PRING <= 6 && (PB == 8 || PB == 16) && CCH % PB == 0 && (!COL_I_PAIR || PB == CCH), "gather ring / reduction groups");

// record words: [c1g[0..11] | c0[0..5] | p0 | run | w10[0..9] | 0 0]   (plan.cu k_col_records)

struct ColGeom {
    int K0, K1, K2, nq2;
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

// ---- packed FP32 (a complex number = one 64-bit register pair): FFMA2 / FMUL2 through the sm_100 intrinsics ----
// bc2(a) = {a, a}: ptxas folds it into the scalar-broadcast operand form (FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2)
typedef float2 P2;      // packed (re, im) pair
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void ffma2_acc(float2& acc, float2 a, float2 b) { acc = __ffma2_rn(a, b, acc); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

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

// Pre-pass of the scatter, one kernel: ys[c][i] = conj(P''_i) * y[perm[i], c] (8-byte slots, coil stride Mpad), the grid
// zero-fill (when the caller's grid is not known to be zero) and the reset of the persistent kernel's work counters.
// The data gather is a chain of two dependent loads per sample (side entry -> y[perm]): a thread takes GU samples at
// a time, all side loads first, then all y loads (GU independent chains in flight per thread), and the zero-fill
// stores of the thread are issued between the two so that the store stream covers part of the latency.
#ifndef COL_GU
#define COL_GU 2       // measured 2 / 4 / 8: gridding stage 227.4 / 229.3 / 233.7 us
#endif
constexpr int GU = COL_GU;
__global__ void k_gather_sorted_col(const float4* __restrict__ side, long long M, long long Mpad,
                                    const float2* __restrict__ y, float2* __restrict__ ys, int nb,
                                    float4* __restrict__ grid4, long long n4, int* __restrict__ counters) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long T = gridDim.x * (long long)blockDim.x;
    for (long long j = i; j < nb; j += T) counters[j] = 0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    bool filled = false;
    for (long long s0 = i; s0 < M || !filled; s0 += GU * T) {
        float4 h[GU];
        float2 v[GU];
#pragma unroll
        for (int j = 0; j < GU; ++j) {
            const long long s = s0 + j * T;
            h[j] = s < M ? __ldg(side + s) : z;                 // P''.re, P''.im, original index
        }
        if (!filled) {
            const long long half = (n4 / T / 2) * T;
            for (long long j = i; j < half; j += T) grid4[j] = z;
#pragma unroll
            for (int j = 0; j < GU; ++j) {
                const long long s = s0 + j * T;
                v[j] = s < M ? y[(long long)__float_as_int(h[j].z) * nb] : make_float2(0.f, 0.f);
            }
            for (long long j = half + i; j < n4; j += T) grid4[j] = z;
            filled = true;
        } else {
#pragma unroll
            for (int j = 0; j < GU; ++j) {
                const long long s = s0 + j * T;
                v[j] = s < M ? y[(long long)__float_as_int(h[j].z) * nb] : make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < GU; ++j) {
            const long long s = s0 + j * T;
            if (s < M) {
                const float2 ph = make_float2(h[j].x, h[j].y);
                const long long m = __float_as_int(h[j].z);
                ys[s] = cmulc(ph, v[j]);
                for (int c = 1; c < nb; ++c) ys[(long long)c * Mpad + s] = cmulc(ph, y[m * nb + c]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gridding (scatter): produces the phase-modulated grid
// ---------------------------------------------------------------------------------------------------------
#ifndef COL_S_RED
#define COL_S_RED 1         // 0: skip the REDs (timing experiments; wrong results)
#endif
#ifndef COL_S_CTAS
#define COL_S_CTAS 4
#endif
// DEMOD = true: the retired plane is demodulated on its way out (conj(m0[plane] m1[row] m2[column]) per cell), so that the
// REDs build the TRUE grid the y2k stage of the API returns -- no separate demodulation pass over the grid.
template <bool DEMOD>
__global__ void __launch_bounds__(CWARPS * 32, COL_S_CTAS)
k_gridding_col(ColGeom g, const WorkItem* __restrict__ work, int n_work, int* __restrict__ counter,
               const float* __restrict__ rec, const float2* __restrict__ ys, long long Mpad, float2* __restrict__ grid,
               int coil0, const float2* __restrict__ mod) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ws = smem_raw + warp * GWARP_BYTES;
    unsigned char* ybuf = ws + 2 * REC_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + 2 * REC_BYTES + 2 * YS_BYTES + 128);
    const int c = blockIdx.y + coil0;
    float2* gc = grid + (long long)c * g.Kprod;
    const float2* ysc = ys + (long long)c * Mpad;
    // lanes 30, 31 shadow lane 29's cells with the always-zero record word 30 as their column weight: their
    // accumulators stay exactly zero and their REDs add 0 to valid cells, so the flush needs no predicate
    const bool active = lane < 3 * CCOLS;
    const int lg = active ? lane / CCOLS : 2;           // row group: box rows lg, lg + 3, lg + 6
    const int lc = active ? lane - CCOLS * lg : CCOLS - 1;   // box column
    const int lw = active ? lc : CCOLS;                 // word 20 + lw of the record: this lane's column weight
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
        // this lane's three cells inside plane 0
        float2* cell[CNR];
        float2 f12[CNR];               // DEMOD: conj(m1[row] m2[column]) of this lane's cells
        {
            int col = q2 * CT2 + lc;
            if (col >= g.K2) col -= g.K2;
#pragma unroll
            for (int i = 0; i < CNR; ++i) {
                int row = q1 * CT1 + lg + 3 * i;
                if (row >= g.K1) row -= g.K1;
                cell[i] = gc + (row * g.K2 + col);
                if (DEMOD) {
                    const float2 t = cmul(__ldg(mod + g.K0 + row), __ldg(mod + g.K0 + g.K1 + col));
                    f12[i] = make_float2(t.x, -t.y);
                }
            }
        }
        const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
        auto issue = [&](int k) {      // lane 0 only
            const int s = wi.begin + k * CCH;
            const int ns = min(CCH, wi.end - s);
            const unsigned b = (gk + k) & 1;
            const int sa = s & ~1;                                     // 16-byte aligned start of the 8-byte data
            const unsigned yb = (unsigned)(((ns + (s & 1) + 1) & ~1) * 8);
            mbar_expect(&mbar[b], (unsigned)(ns * CRECW * 4) + yb);
            tma_bulk(ws + b * REC_BYTES, rec + (long long)s * CRECW, (unsigned)(ns * CRECW * 4), &mbar[b]);
            tma_bulk(ybuf + b * YS_BYTES, ysc + sa, yb, &mbar[b]);
        };
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) issue(0);

        P2 A[6][CNR];                  // [plane slot][row], packed (re, im)
#pragma unroll
        for (int s = 0; s < 6; ++s)
#pragma unroll
            for (int i = 0; i < CNR; ++i) A[s][i] = make_float2(0.f, 0.f);
        int kc = 0, u = 0, ns = 0;      // next chunk, sample inside the chunk, samples of the chunk
        int K = 0;                      // phase = slot of the window's first plane
        int p = 0, pw = 0;              // first plane of the window, the same wrapped into the grid
        int plim = 0;                   // last plane that holds contributions
        int pnext = 0, nrun = 0;        // first plane of the next sample and what is left of its run (INT_MAX: no sample left)
        float2 *cp0 = cell[0], *cp1 = cell[1], *cp2 = cell[2];   // this lane's cells in plane pw
        bool started = false;
        const float* Rb = reinterpret_cast<const float*>(ws);      // set per chunk
        const float2* Y = reinterpret_cast<const float2*>(ybuf);

#define COL_ACC(SL, I, CV, TV) ffma2_acc(A[SL][I], bc2(CV), TV);
#define COL_ACC_ROW(KC, I, TV, C0a, C0b)    \
    COL_ACC((KC + 0) % 6, I, C0a.x, TV)     \
    COL_ACC((KC + 1) % 6, I, C0a.y, TV)     \
    COL_ACC((KC + 2) % 6, I, C0a.z, TV)     \
    COL_ACC((KC + 3) % 6, I, C0a.w, TV)     \
    COL_ACC((KC + 4) % 6, I, C0b.x, TV)     \
    COL_ACC((KC + 5) % 6, I, C0b.y, TV)
        // the record of sample U of the chunk into register set X (reads past the chunk stay inside this warp's
        // shared memory and are never used)
#define COL_S_LOAD(X, U)                                                                           \
    {                                                                                              \
        const float* R = Rb + (U) * CRECW;                                                         \
        X##C1 = *reinterpret_cast<const float4*>(R + 4 * lg);                                      \
        X##C0a = *reinterpret_cast<const float4*>(R + 12);                                         \
        X##C0b = *reinterpret_cast<const float2*>(R + 16);                                         \
        X##w = R[20 + lw];                                                                         \
        X##yv = *reinterpret_cast<const P2*>(Y + (U));                                            \
    }
#define COL_S_BODY(KC, X)                                                                          \
    {                                                                                              \
        const P2 v = fmul2(bc2(X##w), X##yv);                                                     \
        const P2 t0 = fmul2(bc2(X##C1.x), v), t1 = fmul2(bc2(X##C1.y), v), t2 = fmul2(bc2(X##C1.z), v); \
        COL_ACC_ROW(KC, 0, t0, X##C0a, X##C0b)                                                     \
        COL_ACC_ROW(KC, 1, t1, X##C0a, X##C0b)                                                     \
        COL_ACC_ROW(KC, 2, t2, X##C0a, X##C0b)                                                     \
    }
        // phase KC: plane p sits in slot KC.  (pnext, nrun) = first plane and remaining run length of the next sample are
        // carried in registers, so a plane without samples costs no shared-memory load and no chunk test.  Take every
        // sample whose first plane is p -- a counted loop (two samples per trip, all loads ahead of the math) -- then
        // retire plane p: three vector REDs through this lane's running cell pointers, which then move on one plane.
        // Invariant at a phase start: u < ns, or pnext == INT_MAX (all samples of the item taken, the window drains).
#define COL_S_PHASE(KC)                                                                            \
    case KC: {                                                                                     \
        if (pnext == p) {                                                                          \
            int n = min(nrun, ns - u);                                                             \
            _Pragma("unroll 1") for (; n >= 2; n -= 2, u += 2) {                                   \
                COL_S_LOAD(a, u)                                                                   \
                COL_S_LOAD(b, u + 1)                                                               \
                COL_S_BODY(KC, a)                                                                  \
                COL_S_BODY(KC, b)                                                                  \
            }                                                                                      \
            if (n) {                                                                               \
                COL_S_LOAD(a, u)                                                                   \
                COL_S_BODY(KC, a)                                                                  \
                ++u;                                                                               \
            }                                                                                      \
            plim = p + 5;                                                                          \
            if (u == ns) { K = KC; goto chunk_done; }       /* the run may go on in the next chunk */ \
            const int2 pr = *reinterpret_cast<const int2*>(Rb + u * CRECW + 18);   /* p0, run */   \
            pnext = pr.x;                                                                          \
            nrun = pr.y;                                                                           \
        }                                                                                          \
        if (COL_S_RED) {                                                                           \
            if (DEMOD) {                                                                           \
                const float2 f0 = __ldg(mod + pw);                                                 \
                red_v2(cp0, cmul(A[KC][0], cmulc(f0, f12[0])));                                    \
                red_v2(cp1, cmul(A[KC][1], cmulc(f0, f12[1])));                                    \
                red_v2(cp2, cmul(A[KC][2], cmulc(f0, f12[2])));                                    \
            } else {                                                                               \
                red_v2(cp0, A[KC][0]);                                                             \
                red_v2(cp1, A[KC][1]);                                                             \
                red_v2(cp2, A[KC][2]);                                                             \
            }                                                                                      \
        }                                                                                          \
        _Pragma("unroll") for (int i = 0; i < CNR; ++i) A[KC][i] = make_float2(0.f, 0.f);          \
        cp0 += KK; cp1 += KK; cp2 += KK;                                                           \
        ++p;                                                                                       \
        if (++pw == g.K0) { pw = 0; cp0 -= g.Kprod; cp1 -= g.Kprod; cp2 -= g.Kprod; }              \
        if (p > plim) {                 /* nothing left in the window */                           \
            if (pnext == INT_MAX) goto item_done;                                                  \
            const long long d = (long long)(pnext - p) * KK;    /* p < K0 here: pw == p */         \
            cp0 += d; cp1 += d; cp2 += d;                                                          \
            p = pnext; pw = pnext; K = pnext % 6;                                                  \
            continue;                                                                              \
        }                                                                                          \
    }

        float4 aC1, aC0a, bC1, bC0a;
        float2 aC0b, bC0b;
        float aw, bw;
        P2 ayv, byv;
        for (;;) {                      // chunks
            if (kc == nchunks) {        // all samples taken: the phases only retire what is left in the window
                pnext = INT_MAX;
            } else {
                const int s = wi.begin + kc * CCH;
                const unsigned b = (gk + kc) & 1;
                if (kc + 1 < nchunks) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) issue(kc + 1);
                }
                mbar_wait(&mbar[b], ((gk + kc) >> 1) & 1);
                Rb = reinterpret_cast<const float*>(ws + b * REC_BYTES);
                Y = reinterpret_cast<const float2*>(ybuf + b * YS_BYTES) + (s & 1);
                ns = min(CCH, wi.end - s);
                u = 0;
                ++kc;
                const int2 pr = *reinterpret_cast<const int2*>(Rb + 18);
                pnext = pr.x;
                nrun = pr.y;
            }
            if (!started) {
                p = pnext;
                K = p % 6;
                pw = p;
                const long long d = (long long)p * KK;
                cp0 += d; cp1 += d; cp2 += d;
                plim = p + 5;
                started = true;
            }
            for (;;) {                  // phases
                switch (K) {
                    COL_S_PHASE(0)
                    COL_S_PHASE(1)
                    COL_S_PHASE(2)
                    COL_S_PHASE(3)
                    COL_S_PHASE(4)
                    COL_S_PHASE(5)
                }
                K = 0;
            }
        chunk_done:;
        }
    item_done:
        gk += nchunks;
        __syncwarp();
    }
#undef COL_S_PHASE
#undef COL_S_BODY
#undef COL_S_LOAD
#undef COL_ACC_ROW
#undef COL_ACC
}

// ---------------------------------------------------------------------------------------------------------
// interpolation (gather) on the phase-modulated grid
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int mod6(int p) { p %= 6; return p < 0 ? p + 6 : p; }

// The window planes p .. p+5 live in registers (slot = plane mod 6, static inside a phase); the planes behind it arrive
// in a per-warp shared-memory ring by 8-byte cp.async copies, PRING planes ahead of the window, each lane fetching and
// later reading back its OWN three cells (no cross-lane traffic, so the per-thread cp.async groups are all the
// synchronisation there is).  The L2 / DRAM latency of a plane is covered by PRING plane advances instead of by
// registers that hold loads in flight, which is what lets six CTAs of four warps share an SM.
__global__ void __launch_bounds__(CWARPS * 32, COL_I_CTAS)
k_interp_col(ColGeom g, const WorkItem* __restrict__ work, int n_work, int* __restrict__ counter,
             const float* __restrict__ rec, const float4* __restrict__ side, const float2* __restrict__ grid,
             float2* __restrict__ y, int nb, int coil0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ws = smem_raw + warp * IWARP_BYTES;
    P2* pbuf = reinterpret_cast<P2*>(ws + INB * REC_BYTES);
    unsigned char* sbuf = ws + INB * REC_BYTES + PBUF_BYTES;            // [buffer] side entries of the chunk
    const unsigned pring = smem_u32(ws + INB * (REC_BYTES + SIDE_BYTES) + PBUF_BYTES) + lane * 8;   // this lane's entry of ring plane 0, cell 0
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + INB * (REC_BYTES + SIDE_BYTES) + PBUF_BYTES + PRING * PSLOT);
    const int c = blockIdx.y + coil0;
    const float2* gc = grid + (long long)c * g.Kprod;
    // lanes 30, 31 shadow lane 29's cells (finite values) with the always-zero record word 30 as their column weight:
    // their partial sums are exact zeros
    const bool active = lane < 3 * CCOLS;
    const int lg = active ? lane / CCOLS : 2;
    const int lc = active ? lane - CCOLS * lg : CCOLS - 1;
    const int lw = active ? lc : CCOLS;
    const int KK = g.K1 * g.K2;
    if (lane == 0) {
        for (int b = 0; b < INB; ++b) mbar_init(&mbar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    unsigned gk = 0;

    for (int item = next_item(counter + c, lane); item < n_work; item = next_item(counter + c, lane)) {
        const WorkItem wi = work[item];
        const int q1 = wi.tile / g.nq2, q2 = wi.tile - q1 * g.nq2;
        const float2* cell[CNR];
        {
            int col = q2 * CT2 + lc;
            if (col >= g.K2) col -= g.K2;
#pragma unroll
            for (int i = 0; i < CNR; ++i) {
                int row = q1 * CT1 + lg + 3 * i;
                if (row >= g.K1) row -= g.K1;
                cell[i] = gc + (row * g.K2 + col);
            }
        }
        const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
        auto issue = [&](int k) {      // lane 0 only
            const int s = wi.begin + k * CCH;
            const int ns = min(CCH, wi.end - s);
            const unsigned b = (gk + k) % INB;
            mbar_expect(&mbar[b], (unsigned)(ns * CRECW * 4) + (unsigned)(ns * 16));
            tma_bulk(ws + b * REC_BYTES, rec + (long long)s * CRECW, (unsigned)(ns * CRECW * 4), &mbar[b]);
            tma_bulk(sbuf + b * SIDE_BYTES, side + s, (unsigned)(ns * 16), &mbar[b]);
        };
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < INB - 1 && k < nchunks; ++k) issue(k);   // INB - 1 chunks ahead
        cp_async_wait_all();           // plane copies of the previous item that ran past its last sample

        P2 G[6][CNR];                  // window planes p .. p+5: [plane mod 6][row]
#pragma unroll
        for (int s = 0; s < 6; ++s)
#pragma unroll
            for (int i = 0; i < CNR; ++i) G[s][i] = make_float2(0.f, 0.f);
        int kc = 0, u = 0, ns = 0, s0 = 0;
        int K = 0, p = 0, pnext = 0, nrun = 0;
        bool started = false;
        const float* Rb = nullptr;
        const float4* Sb = nullptr;    // (P'', original index) entries of the chunk
        // sum the rows of the transpose buffer (samples base .. base + PB - 1 of the chunk): lane -> (sample, part of its row)
        auto reduce_group = [&](int base) {
            __syncwarp();
            constexpr int PARTS = 32 / PB, LEN = 32 / PARTS;      // PB = 8: 4 parts of 8 lanes' sums
            const int su = lane & (PB - 1), h = lane / PB;
            const P2* pb = pbuf + su * PBUF_PITCH + h * LEN;
            float2 sum = make_float2(0.f, 0.f);
#pragma unroll
            for (int l = 0; l < LEN; ++l) {
                const float2 v = pb[l];
                sum.x += v.x;
                sum.y += v.y;
            }
#pragma unroll
            for (int o = 16; o >= PB; o >>= 1) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
            }
            if (lane < PB && base + lane < ns) {
                const float4 sd = Sb[base + lane];
                y[(long long)__float_as_int(sd.z) * nb + c] = cmul(make_float2(sd.x, sd.y), sum);
            }
            __syncwarp();
        };

#define COL_DOT_ROW(KC, I, C0a, C0b)                                        \
    P2 e##I = fmul2(bc2(C0a.x), G[(KC + 0) % 6][I]);                           \
    ffma2_acc(e##I, bc2(C0a.y), G[(KC + 1) % 6][I]);                            \
    ffma2_acc(e##I, bc2(C0a.z), G[(KC + 2) % 6][I]);                            \
    ffma2_acc(e##I, bc2(C0a.w), G[(KC + 3) % 6][I]);                            \
    ffma2_acc(e##I, bc2(C0b.x), G[(KC + 4) % 6][I]);                            \
    ffma2_acc(e##I, bc2(C0b.y), G[(KC + 5) % 6][I]);
#define COL_I_LOAD(X, U)                                                                           \
    {                                                                                              \
        const float* R = Rb + (U) * CRECW;                                                         \
        X##C1 = *reinterpret_cast<const float4*>(R + 4 * lg);                                      \
        X##C0a = *reinterpret_cast<const float4*>(R + 12);                                         \
        X##C0b = *reinterpret_cast<const float2*>(R + 16);                                         \
        X##w = R[20 + lw];                                                                         \
    }
    // the 32 partial sums of a sample go to row (sample mod PB) of the transpose buffer; every PB samples the rows are
    // summed (lane -> (sample, part of the row)) and stored -- folding them with shuffles first was measured: each fold
    // step costs ~40 us
#define COL_I_BODY(KC, X, U)                                                                       \
    {                                                                                              \
        COL_DOT_ROW(KC, 0, X##C0a, X##C0b)                                                         \
        COL_DOT_ROW(KC, 1, X##C0a, X##C0b)                                                         \
        COL_DOT_ROW(KC, 2, X##C0a, X##C0b)                                                         \
        P2 acc = fmul2(bc2(X##C1.x), e0);                                                         \
        ffma2_acc(acc, bc2(X##C1.y), e1);                                                          \
        ffma2_acc(acc, bc2(X##C1.z), e2);                                                          \
        pbuf[((U) & (PB - 1)) * PBUF_PITCH + lane] = fmul2(bc2(X##w), acc);                        \
    }
#if COL_I_PAIR       // two samples per trip, all loads ahead of the math; needs PB == CCH (one reduction per chunk)
#define COL_I_RUN(KC)                                                                              \
    _Pragma("unroll 1") for (; n >= 2; n -= 2, u += 2) {                                           \
        COL_I_LOAD(a, u)                                                                           \
        COL_I_LOAD(b, u + 1)                                                                       \
        COL_I_BODY(KC, a, u)                                                                       \
        COL_I_BODY(KC, b, u + 1)                                                                   \
    }                                                                                              \
    if (n) {                                                                                       \
        COL_I_LOAD(a, u)                                                                           \
        COL_I_BODY(KC, a, u)                                                                       \
        ++u;                                                                                       \
    }
#else
#define COL_I_RUN(KC)                                                                              \
    _Pragma("unroll 1") for (; n > 0; --n, ++u) {                                                  \
        COL_I_LOAD(a, u)                                                                           \
        COL_I_BODY(KC, a, u)                                                                       \
        if (PB < CCH && (u & (PB - 1)) == PB - 1) reduce_group(u - (PB - 1));                      \
    }
#endif
        // phase KC: plane p sits in window slot KC.  Take every sample whose first plane is p (counted loop over the run;
        // the first plane and the remaining run length of the next sample are kept in registers, so a plane without
        // samples costs no shared-memory load), then plane p + 6 enters slot KC from the shared-memory ring and the
        // copy of plane p + 6 + PRING is started into the ring entry it leaves.
#define COL_I_PHASE(KC)                                                                            \
    case KC: {                                                                                     \
        if (u == ns) { K = KC; goto chunk_done; }                                                  \
        if (pnext == p) {                                                                          \
            int n = min(nrun, ns - u);                                                             \
            COL_I_RUN(KC)                                                                          \
            if (u == ns) { K = KC; goto chunk_done; }                                              \
            const int2 pr = *reinterpret_cast<const int2*>(Rb + u * CRECW + 18);   /* p0, run */   \
            pnext = pr.x;                                                                          \
            nrun = pr.y;                                                                           \
        }                                                                                          \
        {                                                                                          \
            cp_async_wait<PRING - 1>();                                                            \
            const unsigned sa = pring + ((unsigned)(p + 6 + 12 * PRING) % PRING) * PSLOT;                 \
            if (COL_I_LOADS) {                                                                    \
                _Pragma("unroll") for (int i = 0; i < CNR; ++i) G[KC][i] = lds64(sa + i * 256);    \
            }                                                                                      \
            int pl = p + 6 + PRING;                                                                \
            if (pl >= g.K0) pl -= g.K0;                                                            \
            const int off = pl * KK;                                                               \
            if (COL_I_LOADS) {                                                                    \
                _Pragma("unroll") for (int i = 0; i < CNR; ++i) cp_async8(sa + i * 256, cell_at(cell[i], off)); \
            }                                                                                      \
            cp_async_commit();                                                                     \
            ++p;                                                                                   \
            if (pnext - p > 6 + PRING) {    /* jump: nothing on its way is of use, prime the window again */ \
                cp_async_wait_all();                                                               \
                p = pnext - 6 - PRING;                                                             \
                K = mod6(p);                                                                       \
                continue;                                                                          \
            }                                                                                      \
        }                                                                                          \
    }

        float4 aC1, aC0a;
        float2 aC0b;
        float aw;
#if COL_I_PAIR
        float4 bC1, bC0a;
        float2 bC0b;
        float bw;
#endif
        for (;;) {                      // chunks
            {
                s0 = wi.begin + kc * CCH;
                const unsigned b = (gk + kc) % INB;
                if (kc + INB - 1 < nchunks) {   // into the buffer of the chunk that was consumed last
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) issue(kc + INB - 1);
                }
                ns = min(CCH, wi.end - s0);
                mbar_wait(&mbar[b], ((gk + kc) / INB) & 1);
                Rb = reinterpret_cast<const float*>(ws + b * REC_BYTES);
                // phase and original index of the chunk's samples (used after the reduction): they come with the
                // records -- a global load here kept the long scoreboard busy across the chunk boundary
                Sb = reinterpret_cast<const float4*>(sbuf + b * SIDE_BYTES);
                u = 0;
                ++kc;
                const int2 pr = *reinterpret_cast<const int2*>(Rb + 18);
                pnext = pr.x;
                nrun = pr.y;
            }
            if (!started) {             // prime the window: 6 + PRING empty phases bring in planes pnext .. pnext + 5
                p = pnext - 6 - PRING;
                K = mod6(p);
                started = true;
            }
            for (;;) {                  // phases
                switch (K) {
                    COL_I_PHASE(0)
                    COL_I_PHASE(1)
                    COL_I_PHASE(2)
                    COL_I_PHASE(3)
                    COL_I_PHASE(4)
                    COL_I_PHASE(5)
                }
                K = 0;
            }
        chunk_done:
            if (PB == CCH) reduce_group(0);
            else if (ns & (PB - 1)) reduce_group(ns & ~(PB - 1));   // the last, partial group of an item's last chunk
            if (kc == nchunks) break;
        }
        gk += nchunks;
    }
#undef COL_I_PHASE
#undef COL_I_RUN
#undef COL_I_BODY
#undef COL_I_LOAD
#undef COL_DOT_ROW
}

// ---------------------------------------------------------------------------------------------------------
// grid (de)modulation: grid[g] *= (conj) prod_d m_d[g_d]; one CTA per (g0, g1) row
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

// out[g] = in[g] * prod_d m_d[g_d] (in == out allowed)
__global__ void k_modulate(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ mod,
                           int K0, int K1, int K2) {
    const int row = blockIdx.x;
    const int g0 = row / K1, g1 = row - g0 * K1;
    const long long cb = (long long)blockIdx.y * K0 * K1 * K2 + (long long)row * K2;
    const float2 m01 = cmul(__ldg(mod + g0), __ldg(mod + K0 + g1));
    const float2* m2 = mod + K0 + K1;
    for (int g2 = threadIdx.x; g2 < K2; g2 += blockDim.x) out[cb + g2] = cmul(cmul(m01, __ldg(m2 + g2)), in[cb + g2]);
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
// the gather's plane copies run PRING planes ahead of the window: planes up to K0 + 5 + PRING are wrapped with one subtraction
bool col3d_interp_supported(const Geom& g) { return col3d_supported(g) && g.K[0] >= 6 + PRING; }

static ColGeom col_geom(const Geom& g) {
    ColGeom c;
    c.K0 = g.K[0]; c.K1 = g.K[1]; c.K2 = g.K[2];
    c.nq2 = (g.K[2] + CT2 - 1) / CT2;
    c.Kprod = g.Kprod;
    return c;
}

// CTAs per coil of the persistent kernels: all SMs x resident CTAs, shared between the coils of the launch
static int col_ctas(b200nufft_plan_t p, int nb, int resident) {
    if (p->n_sm == 0) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) p->n_sm = prop.multiProcessorCount;
        if (p->n_sm <= 0) p->n_sm = 148;
    }
    return std::max(1, (p->n_sm * resident + nb - 1) / nb);
}

static int col_attrs(b200nufft_plan_t p) {
    if (!p->attr_col) {
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_col<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * GWARP_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_gridding_col<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * GWARP_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_interp_col, cudaFuncAttributeMaxDynamicSharedMemorySize, CWARPS * IWARP_BYTES));
        p->attr_col = true;
    }
    return B200_OK;
}

static int col_counters(b200nufft_plan_t p, int nb) {
    if (p->ccount_nb < nb) {
        if (p->d_ccount) { CUDA_TRY(cudaFree(p->d_ccount)); p->d_ccount = nullptr; p->ccount_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ccount, sizeof(int) * nb));
        p->ccount_nb = nb;
    }
    return B200_OK;
}

int col3d_demodulate(b200nufft_plan_t p, float2* grid, int nb, cudaStream_t st) {
    const Geom& g = p->g;
    dim3 gr((unsigned)(g.K[0] * g.K[1]), nb);
    k_demodulate<<<gr, 128, 0, st>>>(grid, p->d_mod, g.K[0], g.K[1], g.K[2]);
    LAUNCH_CHECK();
    return B200_OK;
}

int col3d_modulate(b200nufft_plan_t p, const float2* in, float2* out, int nb, cudaStream_t st) {
    const Geom& g = p->g;
    dim3 gr((unsigned)(g.K[0] * g.K[1]), nb);
    k_modulate<<<gr, 128, 0, st>>>(in, out, p->d_mod, g.K[0], g.K[1], g.K[2]);
    LAUNCH_CHECK();
    return B200_OK;
}

// grid: phase-modulated grid
int col3d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st) {
    int rc = col_attrs(p);
    if (rc) return rc;
    if (p->n_cwork == 0) return B200_OK;
    rc = col_counters(p, nb);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(p->d_ccount, 0, sizeof(int) * nb, st));
    static const int ictas = [] { const char* e = getenv("B200NUFFT_COL_ICTAS"); return e ? std::max(1, atoi(e)) : COL_I_CTAS; }();   // tuning knob
    // at most COL_COILS coils share the SMs of one launch: with more, the working set of the concurrent sweeps (every
    // plane is read by 4.5 columns) falls out of L2 (32 coils at once: 312 us per coil against 233 us)
    for (int c0 = 0; c0 < nb; c0 += COL_COILS) {
        const int nbg = std::min(COL_COILS, nb - c0);
        dim3 gr((unsigned)std::min((p->n_cwork + CWARPS - 1) / CWARPS, col_ctas(p, nbg, ictas)), nbg);
        k_interp_col<<<gr, CWARPS * 32, CWARPS * IWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                                   p->d_crec, p->d_cside, grid, y, nb, c0);
        LAUNCH_CHECK();
    }
    return B200_OK;
}

// grid receives the phase-modulated adjoint -- or, with `demodulate`, the true grid -- (it is zeroed here, by the pre-pass,
// unless the caller hands in a zeroed one)
int col3d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st, bool prezeroed, bool demodulate) {
    int rc = col_attrs(p);
    if (rc) return rc;
    const long long nel = p->g.Kprod * nb;
    if (p->n_cwork == 0) {
        CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * nel, st));
        return B200_OK;
    }
    const long long Mpad = ((p->M + 1) & ~1LL) + 2;     // even coil stride (16-byte aligned TMA sources) + read slack
    if (p->ys2_nb < nb) {
        if (p->d_ys2) { CUDA_TRY(cudaFree(p->d_ys2)); p->d_ys2 = nullptr; p->ys2_nb = 0; }
        CUDA_TRY(cudaMalloc(&p->d_ys2, sizeof(float2) * Mpad * nb));
        CUDA_TRY(cudaMemsetAsync(p->d_ys2, 0, sizeof(float2) * Mpad * nb, st));
        p->ys2_nb = nb;
    }
    rc = col_counters(p, nb);
    if (rc) return rc;
    // float4 stores need a 16-byte aligned grid and an even element count; otherwise plain memset
    const bool vec = !prezeroed && (reinterpret_cast<uintptr_t>(grid) & 15) == 0 && (nel & 1) == 0;
    if (!vec && !prezeroed) CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * nel, st));
    {
        // sized from the larger of the two jobs (grid zero-fill, data gather), capped: grid-stride loops inside
        const int TB = 256;
        const long long want = std::max<long long>((p->M + TB * GU - 1) / (TB * GU), vec ? (nel / 2 + TB * 8 - 1) / (TB * 8) : 1);
        const unsigned nblk = (unsigned)std::min<long long>(std::max<long long>(want, 1), 148LL * 64);
        k_gather_sorted_col<<<nblk, TB, 0, st>>>(p->d_cside, p->M, Mpad, y, p->d_ys2, nb,
                                                 reinterpret_cast<float4*>(grid), vec ? nel / 2 : 0, p->d_ccount);
        LAUNCH_CHECK();
    }
    for (int c0 = 0; c0 < nb; c0 += COL_COILS) {     // coil groups: see col3d_interp
        const int nbg = std::min(COL_COILS, nb - c0);
        dim3 gr((unsigned)std::min((p->n_cwork + CWARPS - 1) / CWARPS, col_ctas(p, nbg, COL_S_CTAS)), nbg);
        if (demodulate)
            k_gridding_col<true><<<gr, CWARPS * 32, CWARPS * GWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                                               p->d_crec, p->d_ys2, Mpad, grid, c0, p->d_mod);
        else
            k_gridding_col<false><<<gr, CWARPS * 32, CWARPS * GWARP_BYTES, st>>>(col_geom(p->g), p->d_cwork, p->n_cwork, p->d_ccount,
                                                                                p->d_crec, p->d_ys2, Mpad, grid, c0, nullptr);
        LAUNCH_CHECK();
    }
    return B200_OK;
}
