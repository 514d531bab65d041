// 2-D multi-coil kernels with the COIL ON THE LANES and a register-resident row sweep: configurations 2 and 4
// (256^2 / 512^2, J = 6^2, 32 coils).  Used whenever a 2-D J = 6 call has an even number (>= 8) of coils.
//
// Grids are BATCH-INNERMOST here, k[K0][K1][nb] -- the reference's own layout (`vec[col*Reps + nc]`,
// src/re_subroutine.py:808; multi_Kd = Kd + (batch,), linalg/nufft_hsa.py:225-227) -- so that the nb coils of a grid cell
// are one contiguous run: the interpolation weights of a sample are computed ONCE for all coils, every grid access of
// a warp is a full 128-byte line, and y[m, :] is read / written as one coalesced row.
//
//   k_sw2_gridding replaces pELL_spmvh_mCoil + atomic_add_float2 (re_subroutine.py:527-596, 275-287)
//   k_sw2_interp   replaces pELL_spmv_mCoil (re_subroutine.py:751-835)
//   k_sw2_scale_pad / k_sw2_crop / k_sw2_crop_combine: the pad / crop / coil-combine stages in this layout
//   (cTensorMultiply + cTensorCopy, cMultiplyConjVecInplace + cAggregate; re_subroutine.py:98-201, 441-514), with a
//   strided-batch cuFFT plan in between.
//
// The sweep (the 2-D sibling of col3d.cu): samples are sorted by (strip, first row), a strip being SW2_CT = 3
// first-neighbour columns (dim 1) wide; its box is 8 columns.  A warp takes a segment of a strip and 16 coils: lane
// (h, c) = (column half, coil) owns box columns 4h .. 4h+3 of the six rows of the window [p, p+6) for coil c, 24 complex
// accumulators in registers.  The code is unrolled over p mod 6 (static plane phases, packed FFMA2 math, counted loops
// over the run of samples that share a first row, as in col3d.cu).  All weights are REAL: the kernels work on the
// phase-modulated grid G'[g] = G[g] m0[g0] m1[g1], m_d[g] = e^{i s_d g}; the modulation is applied where a row enters
// the registers (gather) and undone where a row is flushed (scatter), with factors that are uniform over the coils, so
// the grids in memory are the TRUE grids.  Neighbours that wrap around the periodic grid carry (-1)^(N-1) in their
// weights (plan.cu k_sw2_records).
//   scatter: a retired row is flushed by all lanes at once, 4 REDs per lane covering 128-byte runs; y rows of a chunk
//            arrive by cp.async one chunk ahead, records by TMA bulk copies two chunks ahead;
//   gather:  the rows of the strip arrive in a shared-memory ring by TMA bulk copies (one 128-byte run per box column
//            and row, issued 8 rows ahead), enter the registers when the window reaches them; the two column halves of
//            a sample are combined with one shuffle and y[m, 16 coils] is stored as one 128-byte run.
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace {

typedef float2 P2;
constexpr int CT = SW2_CT;                // strip width (first-neighbour columns): 3
constexpr int CB = CT + 5;                // box columns: 8
constexpr int CH = CB / 2;                // columns per lane: 4
constexpr int RECW = SW2_RECW;            // words per record: 20
constexpr int CCH = 16;                   // samples per chunk
constexpr int NCL = 16;                   // coils per warp
constexpr int SWARPS = 4;                 // warps per CTA (each warp works on its own item)
#ifndef SW2_CTAS
#define SW2_CTAS 4                         // resident CTAs per SM the register allocation aims at
#endif
constexpr int REC_BYTES = CCH * RECW * 4;             // 1280
constexpr int YCH_BYTES = CCH * NCL * 8;              // 2048: y rows of a chunk, 16 coils
constexpr int GW_BYTES = 3 * REC_BYTES + 2 * YCH_BYTES + 128 + 64;       // scatter: 8128
constexpr int GW_PITCH = 8192;
constexpr int RS = 8;                                  // gather: rows in the shared-memory ring
constexpr int JUMP = 12;                               // gather: gaps longer than this restart the window instead of stepping
constexpr int ROW_BYTES = CB * NCL * 8;                // 1024 per ring row
constexpr int IW_BYTES = 2 * REC_BYTES + RS * ROW_BYTES + 128;            // gather: 10880
constexpr int IW_PITCH = 11008;
static_assert(GW_BYTES <= GW_PITCH && IW_BYTES <= IW_PITCH && GW_PITCH % 128 == 0 && IW_PITCH % 128 == 0, "smem");

// record words: [c0[0..3] | c0[4] c0[5] p0 run | c1t[0..3] | c1t[4..7] | P''.re P''.im m 0]   (plan.cu k_sw2_records)

struct Sw2Geom {
    int K0, K1, nq;
    int nb;
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
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ P2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void ffma2_acc(P2& acc, P2 a, P2 b) { acc = __ffma2_rn(a, b, acc); }
__device__ __forceinline__ P2 fmul2(P2 a, P2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ void red_v2(float2* addr, float2 a) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a.x), "f"(a.y) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// pad / crop stages in the batch-innermost layout
// ---------------------------------------------------------------------------------------------------------
// one thread per (grid cell, coil): grid[(i0 K1 + i1) nb + c] = x * [sn] * [sens] inside the image corner, else 0
__global__ void k_sw2_scale_pad(int N0, int N1, int K0, int K1, int sn0, int sn1, const float* __restrict__ sn,
                                const float2* __restrict__ x, float2* __restrict__ grid, int nb, int apply_sn,
                                int x_single, const float2* __restrict__ sens) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= (long long)K0 * K1 * nb) return;
    const int c = (int)(gid % nb);
    const int cell = (int)(gid / nb);
    const int i1 = cell % K1, i0 = cell / K1;
    float2 v = make_float2(0.f, 0.f);
    if (i0 < N0 && i1 < N1) {
        const long long n = (long long)i0 * N1 + i1;
        float2 xv = x_single ? x[n] : x[n * nb + c];
        if (sens) xv = cmul(xv, sens[n * nb + c]);
        const float f = apply_sn ? sn[sn0 + i0] * sn[sn1 + i1] : 1.f;
        v = make_float2(xv.x * f, xv.y * f);
    }
    grid[gid] = v;
}

__global__ void k_sw2_crop(int N0, int N1, int K1, int sn0, int sn1, const float* __restrict__ sn,
                           const float2* __restrict__ grid, float2* __restrict__ x, int nb, int mode, float scale) {
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= (long long)N0 * N1 * nb) return;
    const int c = (int)(gid % nb);
    const int n = (int)(gid / nb);
    const int n1 = n % N1, n0 = n / N1;
    const float s = sn[sn0 + n0] * sn[sn1 + n1];
    const float f = scale * (mode == 1 ? s : (mode == 2 ? 1.f / s : 1.f));
    const float2 v = grid[((long long)n0 * K1 + n1) * nb + c];
    x[gid] = make_float2(v.x * f, v.y * f);
}

// one warp per pixel: s[n] = (1/nb) sum_c conj(sens[n, c]) * grid[cell(n), c] * f
__global__ void k_sw2_crop_combine(int N0, int N1, int K1, int sn0, int sn1, const float* __restrict__ sn,
                                   const float2* __restrict__ grid, float2* __restrict__ x, int nb, int mode, float scale,
                                   const float2* __restrict__ sens) {
    const int lane = threadIdx.x & 31;
    const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (n >= (long long)N0 * N1) return;
    const int n1 = (int)(n % N1), n0 = (int)(n / N1);
    const float2* gp = grid + ((long long)n0 * K1 + n1) * nb;
    float2 acc = make_float2(0.f, 0.f);
    for (int c = lane; c < nb; c += 32) {
        float2 v = gp[c];
        if (sens) v = cmulc(sens[n * nb + c], v);
        acc.x += v.x;
        acc.y += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (lane == 0) {
        const float s = sn[sn0 + n0] * sn[sn1 + n1];
        const float f = scale * (mode == 1 ? s : (mode == 2 ? 1.f / s : 1.f)) / (float)nb;
        x[n] = make_float2(acc.x * f, acc.y * f);
    }
}

// layout change between coil-major [nb][K] and batch-innermost [K][nb] (32 x 32 tiles through shared memory).
// cuFFT's strided-batch plans are slow on the batch-innermost layout (282 us against 56 us for the 32 x 512^2 grids of
// configuration 2), so the FFT runs coil-major on a scratch copy: one transposing pass on the way.
template <bool TO_BI>
__global__ void k_sw2_transpose(const float2* __restrict__ in, float2* __restrict__ out, long long K, int nb) {
    __shared__ float2 tile[32][33];
    const long long cell0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
    if (TO_BI) {                                        // in[c][cell] -> out[cell][c]
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + j;
            const long long cell = cell0 + tx;
            if (c < nb && cell < K) tile[j][tx] = in[(long long)c * K + cell];
        }
        __syncthreads();
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const long long cell = cell0 + j;
            const int c = c0 + tx;
            if (c < nb && cell < K) out[cell * nb + c] = tile[tx][j];
        }
    } else {                                            // in[cell][c] -> out[c][cell]
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const long long cell = cell0 + j;
            const int c = c0 + tx;
            if (c < nb && cell < K) tile[j][tx] = in[cell * nb + c];
        }
        __syncthreads();
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + j;
            const long long cell = cell0 + tx;
            if (c < nb && cell < K) out[(long long)c * K + cell] = tile[tx][j];
        }
    }
}

// coil-major pad (one thread per cell of one coil): cm[c][i0 K1 + i1]
__global__ void k_sw2_scale_pad_cm(int N0, int N1, int K0, int K1, int sn0, int sn1, const float* __restrict__ sn,
                                   const float2* __restrict__ x, float2* __restrict__ cm, int nb, int apply_sn,
                                   int x_single, const float2* __restrict__ sens) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (cell >= K0 * K1) return;
    const int i1 = cell % K1, i0 = cell / K1;
    float2 v = make_float2(0.f, 0.f);
    if (i0 < N0 && i1 < N1) {
        const long long n = (long long)i0 * N1 + i1;
        float2 xv = x_single ? x[n] : x[n * nb + c];
        if (sens) xv = cmul(xv, sens[n * nb + c]);
        const float f = apply_sn ? sn[sn0 + i0] * sn[sn1 + i1] : 1.f;
        v = make_float2(xv.x * f, xv.y * f);
    }
    cm[(long long)c * K0 * K1 + cell] = v;
}

// ---------------------------------------------------------------------------------------------------------
// gridding (scatter): grid (zeroed by the caller) += A^H y, true grid, batch-innermost
// ---------------------------------------------------------------------------------------------------------
#define SW2_LOAD(X, U)                                                                             \
    {                                                                                              \
        const float* R = Rb + (U) * RECW;                                                          \
        X##W0 = *reinterpret_cast<const float4*>(R);                                               \
        X##W1 = *reinterpret_cast<const float2*>(R + 4);                                           \
        X##C1 = *reinterpret_cast<const float4*>(R + 8 + 4 * lh);                                  \
        X##P = *reinterpret_cast<const float2*>(R + 16);                                           \
    }

// MODG: the grid is left phase-modulated (the inverse FFT pass along dim 0 undoes it, fftbi.cu); the flush is then a
// plain RED of the accumulators and the per-row factor is never loaded
template <bool MODG>
__global__ void __launch_bounds__(SWARPS * 32, SW2_CTAS)
k_sw2_gridding(Sw2Geom g, const WorkItem* __restrict__ work, int n_work, const float* __restrict__ rec,
               const float2* __restrict__ mod, const float2* __restrict__ y, float2* __restrict__ grid) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * SWARPS + warp;
    if (item >= n_work) return;
    unsigned char* ws = smem_raw + warp * GW_PITCH;
    float2* sy = reinterpret_cast<float2*>(ws + 3 * REC_BYTES);                   // [2][CCH][NCL]
    float* dummy = reinterpret_cast<float*>(ws + 3 * REC_BYTES + 2 * YCH_BYTES);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + 3 * REC_BYTES + 2 * YCH_BYTES + 128);
    const int lh = lane >> 4, lc = lane & 15;          // column half, coil inside the block
    const int cb = blockIdx.y * NCL;                   // first coil of this warp
    const int nc = min(NCL, g.nb - cb);
    const bool cact = lc < nc;
    const int nb = g.nb;
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    dummy[lane] = lane == 6 ? __int_as_float(INT_MAX) : 0.f;       // the record that ends the item: first row INT_MAX
    __syncwarp();

    const WorkItem wi = work[item];
    const int q3 = wi.tile * CT;                       // first box column of the strip
    // this lane's four cells inside row 0 (element offsets) and their demodulation factors conj(m1[col])
    int coff[CH];
    float2 m1c[CH];
    const float2* m0 = mod;
    const float2* m1 = mod + g.K0;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        int col = q3 + CH * lh + i;
        if (col >= g.K1) col -= g.K1;
        coff[i] = col * nb + cb + lc;
        const float2 m = __ldg(m1 + col);
        m1c[i] = make_float2(m.x, -m.y);
    }
    const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
    auto issue_rec = [&](int k) {      // lane 0 only: records of chunk k -> buffer k mod 3
        const int s = wi.begin + k * CCH;
        const int ns = min(CCH, wi.end - s);
        const int b = k % 3;
        mbar_expect(&mbar[b], (unsigned)(ns * RECW * 4));
        tma_bulk(ws + b * REC_BYTES, rec + (long long)s * RECW, (unsigned)(ns * RECW * 4), &mbar[b]);
    };
    auto wait_rec = [&](int k) { mbar_wait(&mbar[k % 3], (unsigned)((k / 3) & 1)); };
    auto issue_y = [&](int k) {        // all lanes: y rows of chunk k (records present) -> sy[k & 1]
        const int s = wi.begin + k * CCH;
        const int ns = min(CCH, wi.end - s);
        const float* R = reinterpret_cast<const float*>(ws + (k % 3) * REC_BYTES);
        float2* dst = sy + (k & 1) * (CCH * NCL);
        // lane (h, c): samples h, h+2, ...; coil c
        for (int u = lh; u < ns; u += 2) {
            const int m = __float_as_int(R[u * RECW + 18]);
            if (cact) cp_async8(dst + u * NCL + lc, y + (long long)m * nb + cb + lc);
        }
        cp_async_commit();
    };
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        issue_rec(0);
        if (nchunks > 1) issue_rec(1);
    }
    wait_rec(0);
    issue_y(0);

    P2 A[6][CH];                        // [row slot][column]
#pragma unroll
    for (int s = 0; s < 6; ++s)
#pragma unroll
        for (int i = 0; i < CH; ++i) A[s][i] = make_float2(0.f, 0.f);
    int kc = 0, u = 0, ns = 0;
    int K = 0, p = 0, pw = 0, plim = 0, pnext = 0, nrun = 0;
    bool started = false;
    const float* Rb = dummy;
    const float2* Y = sy;
    float4 aW0, aC1;
    float2 aW1, aP;

#define SW2_ACC(SL, I, CV, TV) ffma2_acc(A[SL][I], bc2(CV), TV);
#define SW2_ACC_COL(KC, I, TV, X)             \
    SW2_ACC((KC + 0) % 6, I, X##W0.x, TV)     \
    SW2_ACC((KC + 1) % 6, I, X##W0.y, TV)     \
    SW2_ACC((KC + 2) % 6, I, X##W0.z, TV)     \
    SW2_ACC((KC + 3) % 6, I, X##W0.w, TV)     \
    SW2_ACC((KC + 4) % 6, I, X##W1.x, TV)     \
    SW2_ACC((KC + 5) % 6, I, X##W1.y, TV)
#define SW2_S_BODY(KC, X, U)                                                                       \
    {                                                                                              \
        const float2 yv = Y[(U) * NCL + lc];                                                       \
        /* v = conj(P'') * y */                                                                    \
        P2 v = fmul2(bc2(X##P.x), yv);                                                             \
        ffma2_acc(v, make_float2(X##P.y, -X##P.y), make_float2(yv.y, yv.x));                       \
        const P2 t0 = fmul2(bc2(X##C1.x), v), t1 = fmul2(bc2(X##C1.y), v), t2 = fmul2(bc2(X##C1.z), v), \
                 t3 = fmul2(bc2(X##C1.w), v);                                                      \
        SW2_ACC_COL(KC, 0, t0, X)                                                                  \
        SW2_ACC_COL(KC, 1, t1, X)                                                                  \
        SW2_ACC_COL(KC, 2, t2, X)                                                                  \
        SW2_ACC_COL(KC, 3, t3, X)                                                                  \
    }
    // phase KC: row p sits in slot KC.  Take every sample whose first row is p (counted loop over the run), then retire
    // row p: demodulate, flush with REDs (128-byte runs: 16 coils of one cell), clear.
#define SW2_S_PHASE(KC)                                                                            \
    case KC: {                                                                                     \
        const float2 f0r = MODG ? make_float2(1.f, 0.f) : __ldg(m0 + pw);   /* consumed at the flush below */ \
        if (pnext == p) {               /* (pnext, nrun): first row / remaining run of the next sample, in registers */ \
            int n = min(nrun, ns - u);                                                             \
            _Pragma("unroll 1") for (; n > 0; --n, ++u) {                                          \
                SW2_LOAD(a, u)                                                                     \
                SW2_S_BODY(KC, a, u)                                                               \
            }                                                                                      \
            plim = p + 5;                                                                          \
            if (u == ns) { K = KC; goto chunk_done; }       /* the run may go on in the next chunk */ \
            const int2 pr = *reinterpret_cast<const int2*>(Rb + u * RECW + 6);     /* p0, run */   \
            pnext = pr.x;                                                                          \
            nrun = pr.y;                                                                           \
        }                                                                                          \
        {                                                                                          \
            const float2 f0 = make_float2(f0r.x, -f0r.y);                                          \
            const long long rowoff = (long long)pw * g.K1 * nb;                                    \
            _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                       \
                const float2 v = MODG ? A[KC][i] : cmul(A[KC][i], cmul(f0, m1c[i]));               \
                if (cact) red_v2(grid + rowoff + coff[i], v);                                      \
                A[KC][i] = make_float2(0.f, 0.f);                                                  \
            }                                                                                      \
        }                                                                                          \
        ++p; ++pw;                                                                                 \
        if (pw == g.K0) pw = 0;                                                                    \
        if (p > plim) {                 /* nothing left in the window */                           \
            if (pnext == INT_MAX) goto item_done;                                                  \
            p = pnext; pw = pnext; K = pnext % 6;                                                  \
            continue;                                                                              \
        }                                                                                          \
    }

    for (;;) {                          // chunks
        if (kc == nchunks) {            // all samples taken: the phases only retire what is left in the window
            pnext = INT_MAX;
        } else {
            // records two chunks ahead, y rows one chunk ahead (every lane is done with the buffers they overwrite)
            __syncwarp();
            if (kc + 2 < nchunks) {
                fence_proxy_async();
                if (lane == 0) issue_rec(kc + 2);
            }
            if (kc + 1 < nchunks) {
                wait_rec(kc + 1);
                issue_y(kc + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            Rb = reinterpret_cast<const float*>(ws + (kc % 3) * REC_BYTES);
            Y = sy + (kc & 1) * (CCH * NCL);
            ns = min(CCH, wi.end - (wi.begin + kc * CCH));
            u = 0;
            ++kc;
            const int2 pr = *reinterpret_cast<const int2*>(Rb + 6);
            pnext = pr.x;
            nrun = pr.y;
        }
        if (!started) {
            p = pnext;
            K = p % 6;
            pw = p;
            plim = p + 5;
            started = true;
        }
        for (;;) {                      // phases
            switch (K) {
                SW2_S_PHASE(0)
                SW2_S_PHASE(1)
                SW2_S_PHASE(2)
                SW2_S_PHASE(3)
                SW2_S_PHASE(4)
                SW2_S_PHASE(5)
            }
            K = 0;
        }
    chunk_done:;
    }
item_done:;
#undef SW2_S_PHASE
#undef SW2_S_BODY
#undef SW2_ACC_COL
#undef SW2_ACC
}

// ---------------------------------------------------------------------------------------------------------
// interpolation (gather): y = A k, true grid in, batch-innermost
// ---------------------------------------------------------------------------------------------------------
// MODG: the grid arrives phase-modulated (written so by the forward FFT pass along dim 0, fftbi.cu): rows enter the
// window as they are
template <bool MODG>
__global__ void __launch_bounds__(SWARPS * 32, SW2_CTAS)
k_sw2_interp(Sw2Geom g, const WorkItem* __restrict__ work, int n_work, const float* __restrict__ rec,
             const float2* __restrict__ mod, const float2* __restrict__ grid, float2* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * SWARPS + warp;
    if (item >= n_work) return;
    unsigned char* ws = smem_raw + warp * IW_PITCH;
    const float2* ring = reinterpret_cast<const float2*>(ws + 2 * REC_BYTES);     // [RS][CB][NCL]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ws + 2 * REC_BYTES + RS * ROW_BYTES);   // [0..1] records, [2..9] rows
    const int lh = lane >> 4, lc = lane & 15;
    const int cb = blockIdx.y * NCL;
    const int nc = min(NCL, g.nb - cb);
    const bool cact = lc < nc;
    const int nb = g.nb;
    if (lane == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();

    const WorkItem wi = work[item];
    const int q3 = wi.tile * CT;
    const float2* m0 = mod;
    const float2* m1 = mod + g.K0;
    float2 m1v[CH];                     // modulation factors m1[col] of this lane's columns
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        int col = q3 + CH * lh + i;
        if (col >= g.K1) col -= g.K1;
        m1v[i] = __ldg(m1 + col);
    }
    // a row of the box = 8 cells x 16 coils = 64 chunks of 16 bytes: lane l copies chunks l and l + 32 with cp.async
    // (one commit group per row; rows complete in order, so "row r has landed" = at most (newest - r) groups pending)
    const float2* fsrc[2];
    bool fok[2];
    unsigned fdst[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        int col = q3 + (j >> 3);
        if (col >= g.K1) col -= g.K1;
        fsrc[t] = grid + (long long)col * nb + cb + 2 * (j & 7);
        fok[t] = 2 * (j & 7) < nc;
        fdst[t] = smem_u32(ws + 2 * REC_BYTES) + (unsigned)(j * 16);
    }
    auto fetch_row = [&](int row) {     // row >= 0, possibly >= K0 (wrapped) -> ring slot row mod RS
        int rw = row;
        if (rw >= g.K0) rw -= g.K0;
        const long long off = (long long)rw * g.K1 * nb;
        const unsigned so = (unsigned)((row & (RS - 1)) * ROW_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (fok[t]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(fdst[t] + so), "l"(fsrc[t] + off) : "memory");
        cp_async_commit();
    };
    // the values of row `row` (landed): this lane's four columns times the modulation m0[row] m1[col]
    auto row_factor = [&](int row) {    // m0[row], issued well before it is needed
        if (MODG) return make_float2(1.f, 0.f);
        int rw = row;
        if (rw >= g.K0) rw -= g.K0;
        return __ldg(m0 + rw);
    };
    auto row_values = [&](int row, const float2 f0, float2 (&v)[CH]) {
        const float2* rp = ring + (row & (RS - 1)) * (CB * NCL) + (CH * lh) * NCL + lc;
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = MODG ? rp[i * NCL] : cmul(rp[i * NCL], cmul(f0, m1v[i]));
        __syncwarp();                   // all lanes have read the slot before it is fetched into again
    };
    const int nchunks = (wi.end - wi.begin + CCH - 1) / CCH;
    auto issue_rec = [&](int k) {
        const int s = wi.begin + k * CCH;
        const int ns = min(CCH, wi.end - s);
        const int b = k & 1;
        mbar_expect(&mbar[b], (unsigned)(ns * RECW * 4));
        tma_bulk(ws + b * REC_BYTES, rec + (long long)s * RECW, (unsigned)(ns * RECW * 4), &mbar[b]);
    };
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) issue_rec(0);

    P2 G[6][CH];                        // window rows p .. p+5 (slot = row mod 6), this lane's four columns, modulated
#pragma unroll
    for (int s = 0; s < 6; ++s)
#pragma unroll
        for (int i = 0; i < CH; ++i) G[s][i] = make_float2(0.f, 0.f);
    int kc = 0, u = 0, ns = 0;
    int K = 0, p = 0, pnext = 0, nrun = 0;
    int rfetched = 0;                   // rows < rfetched are in (or on their way to) the ring
    bool started = false;
    const float* Rb = nullptr;
    float4 aW0, aC1;
    float2 aW1, aP;

    // row ROW (fetched earlier) enters the window in the static slot SLOT
#define SW2_TAKE_ROW(SLOT, ROW, F0)                                                                \
    {                                                                                              \
        cp_async_wait<RS - 7>();            /* rows ROW + 1 .. may still be in flight */         \
        __syncwarp();                                                                              \
        float2 v_[CH];                                                                             \
        row_values(ROW, F0, v_);                                                                   \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) G[SLOT][i] = v_[i];                         \
    }
#define SW2_DOT_COL(KC, I, X)                                               \
    P2 e##I = fmul2(bc2(X##W0.x), G[(KC + 0) % 6][I]);                      \
    ffma2_acc(e##I, bc2(X##W0.y), G[(KC + 1) % 6][I]);                      \
    ffma2_acc(e##I, bc2(X##W0.z), G[(KC + 2) % 6][I]);                      \
    ffma2_acc(e##I, bc2(X##W0.w), G[(KC + 3) % 6][I]);                      \
    ffma2_acc(e##I, bc2(X##W1.x), G[(KC + 4) % 6][I]);                      \
    ffma2_acc(e##I, bc2(X##W1.y), G[(KC + 5) % 6][I]);
#define SW2_I_BODY(KC, X, U)                                                                       \
    {                                                                                              \
        SW2_DOT_COL(KC, 0, X)                                                                      \
        SW2_DOT_COL(KC, 1, X)                                                                      \
        SW2_DOT_COL(KC, 2, X)                                                                      \
        SW2_DOT_COL(KC, 3, X)                                                                      \
        P2 acc = fmul2(bc2(X##C1.x), e0);                                                          \
        ffma2_acc(acc, bc2(X##C1.y), e1);                                                          \
        ffma2_acc(acc, bc2(X##C1.z), e2);                                                          \
        ffma2_acc(acc, bc2(X##C1.w), e3);                                                          \
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);                                          \
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);                                          \
        const int m_ = __float_as_int(Rb[(U) * RECW + 18]);                                        \
        if (lh == 0 && cact) y[(long long)m_ * nb + cb + lc] = cmul(X##P, acc);                    \
    }
    // phase KC: row p sits in slot KC.  Take every sample whose first row is p, then row p + 6 replaces row p in the
    // registers and the ring is topped up (RS rows ahead of the window).
#define SW2_I_PHASE(KC)                                                                            \
    case KC: {                                                                                     \
        const float2 f0n = row_factor(p + 6);              /* consumed when row p + 6 is taken below */ \
        if (pnext == p) {               /* (pnext, nrun): first row / remaining run of the next sample, in registers */ \
            int n = min(nrun, ns - u);                                                             \
            _Pragma("unroll 1") for (; n > 0; --n, ++u) {                                          \
                SW2_LOAD(a, u)                                                                     \
                SW2_I_BODY(KC, a, u)                                                               \
            }                                                                                      \
            if (u == ns) { K = KC; goto chunk_done; }                                              \
            const int2 pr = *reinterpret_cast<const int2*>(Rb + u * RECW + 6);     /* p0, run */   \
            pnext = pr.x;                                                                          \
            nrun = pr.y;                                                                           \
        }                                                                                          \
        if (pnext - p > JUMP) {         /* long gap: restart the window at the next sample's first row */ \
            cp_async_wait<0>();                 /* rows in flight must land first */            \
            __syncwarp();                                                                          \
            p = pnext; K = pnext % 6;                                                              \
            goto prime;                                                                            \
        }                                                                                          \
        SW2_TAKE_ROW(KC, p + 6, f0n)                                                               \
        fetch_row(rfetched);                                                                       \
        ++rfetched;                                                                                \
        ++p;                                                                                       \
    }

    for (;;) {                          // chunks
        {
            const int s0 = wi.begin + kc * CCH;
            const int b = kc & 1;
            if (kc + 1 < nchunks) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) issue_rec(kc + 1);
            }
            ns = min(CCH, wi.end - s0);
            mbar_wait(&mbar[b], (unsigned)((kc >> 1) & 1));
            Rb = reinterpret_cast<const float*>(ws + b * REC_BYTES);
            u = 0;
            ++kc;
            const int2 pr = *reinterpret_cast<const int2*>(Rb + 6);
            pnext = pr.x;
            nrun = pr.y;
        }
        if (!started) {
            p = pnext;
            K = p % 6;
            started = true;
            goto prime;
        }
    dispatch:
        for (;;) {                      // phases
            switch (K) {
                SW2_I_PHASE(0)
                SW2_I_PHASE(1)
                SW2_I_PHASE(2)
                SW2_I_PHASE(3)
                SW2_I_PHASE(4)
                SW2_I_PHASE(5)
            }
            K = 0;
        }
    prime:
        // (re)start the window at row p (nothing is in flight): fetch rows p .. p + RS - 1, take rows p .. p + 5
#pragma unroll 1
        for (int r = 0; r < RS; ++r) fetch_row(p + r);
        rfetched = p + RS;
        cp_async_wait<RS - 6>();
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < 6; ++j) {
            float2 v[CH];
            row_values(p + j, row_factor(p + j), v);
            switch ((K + j) % 6) {
                case 0: for (int i = 0; i < CH; ++i) G[0][i] = v[i]; break;
                case 1: for (int i = 0; i < CH; ++i) G[1][i] = v[i]; break;
                case 2: for (int i = 0; i < CH; ++i) G[2][i] = v[i]; break;
                case 3: for (int i = 0; i < CH; ++i) G[3][i] = v[i]; break;
                case 4: for (int i = 0; i < CH; ++i) G[4][i] = v[i]; break;
                default: for (int i = 0; i < CH; ++i) G[5][i] = v[i]; break;
            }
        }
        goto dispatch;
    chunk_done:
        if (kc == nchunks) break;
    }
    // rows still in flight must land before the shared memory is released
    cp_async_wait<0>();
#undef SW2_I_PHASE
#undef SW2_I_BODY
#undef SW2_DOT_COL
#undef SW2_TAKE_ROW
}
#undef SW2_LOAD

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool sweep2d_supported(const Geom& g) {
    return g.ndim == 2 && g.J[0] == 6 && g.J[1] == 6 && g.K[0] >= 16 && g.K[1] >= CB;
}

static Sw2Geom sw2_geom(const Geom& g, int nb) {
    Sw2Geom s;
    s.K0 = g.K[0];
    s.K1 = g.K[1];
    s.nq = (g.K[1] + CT - 1) / CT;
    s.nb = nb;
    return s;
}

static int sw2_attrs(b200nufft_plan_t p) {
    if (!p->attr_sw2) {
        CUDA_TRY(cudaFuncSetAttribute(k_sw2_gridding<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWARPS * GW_PITCH));
        CUDA_TRY(cudaFuncSetAttribute(k_sw2_gridding<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWARPS * GW_PITCH));
        CUDA_TRY(cudaFuncSetAttribute(k_sw2_interp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWARPS * IW_PITCH));
        CUDA_TRY(cudaFuncSetAttribute(k_sw2_interp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWARPS * IW_PITCH));
        p->attr_sw2 = true;
    }
    return B200_OK;
}

int sweep2d_scale_pad(b200nufft_plan_t p, const float2* x, float2* grid, int nb, int apply_sn, int x_single,
                      const float2* sens, cudaStream_t st) {
    const Geom& g = p->g;
    const long long tot = g.Kprod * nb;
    const int TB = 256;
    k_sw2_scale_pad<<<(unsigned)((tot + TB - 1) / TB), TB, 0, st>>>(g.N[0], g.N[1], g.K[0], g.K[1], g.snoff[0], g.snoff[1],
                                                                   p->d_sn, x, grid, nb, apply_sn, x_single, sens);
    LAUNCH_CHECK();
    return B200_OK;
}

int sweep2d_crop_scale(b200nufft_plan_t p, const float2* grid, float2* x, int nb, int mode, int combine,
                       const float2* sens, float scale, cudaStream_t st) {
    const Geom& g = p->g;
    const int TB = 256;
    if (combine) {
        const long long thr = g.Nprod * 32;
        k_sw2_crop_combine<<<(unsigned)((thr + TB - 1) / TB), TB, 0, st>>>(g.N[0], g.N[1], g.K[1], g.snoff[0], g.snoff[1],
                                                                          p->d_sn, grid, x, nb, mode, scale, sens);
    } else {
        const long long tot = g.Nprod * nb;
        k_sw2_crop<<<(unsigned)((tot + TB - 1) / TB), TB, 0, st>>>(g.N[0], g.N[1], g.K[1], g.snoff[0], g.snoff[1], p->d_sn,
                                                                  grid, x, nb, mode, scale);
    }
    LAUNCH_CHECK();
    return B200_OK;
}

static int sw2_get_fft(b200nufft_plan_t p, int nb) {
    if (p->fft_bi_valid && p->fft_bi_nb == nb) return B200_OK;
    if (p->fft_bi_valid) { cufftDestroy(p->fft_bi); p->fft_bi_valid = false; }
    int n[2] = {p->g.K[0], p->g.K[1]};
    CUFFT_TRY(cufftPlanMany(&p->fft_bi, 2, n, nullptr, 1, (int)p->g.Kprod, nullptr, 1, (int)p->g.Kprod, CUFFT_C2C, nb));
    p->fft_bi_valid = true;
    p->fft_bi_nb = nb;
    return B200_OK;
}

static int sw2_transpose(const float2* in, float2* out, long long K, int nb, bool to_bi, cudaStream_t st) {
    dim3 gr((unsigned)((K + 31) / 32), (unsigned)((nb + 31) / 32)), blk(32, 8);
    if (to_bi) k_sw2_transpose<true><<<gr, blk, 0, st>>>(in, out, K, nb);
    else k_sw2_transpose<false><<<gr, blk, 0, st>>>(in, out, K, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

// FFT of a batch-innermost grid (in place for the caller): transpose to the plan's coil-major scratch, batched cuFFT
// there, transpose back
int sweep2d_fft(b200nufft_plan_t p, float2* grid, int nb, int inverse, cudaStream_t st) {
    int rc = ensure_scratch2(p, nb);
    if (rc) return rc;
    rc = sw2_get_fft(p, nb);
    if (rc) return rc;
    rc = sw2_transpose(grid, p->d_grid2, p->g.Kprod, nb, false, st);
    if (rc) return rc;
    CUFFT_TRY(cufftSetStream(p->fft_bi, st));
    CUFFT_TRY(cufftExecC2C(p->fft_bi, reinterpret_cast<cufftComplex*>(p->d_grid2), reinterpret_cast<cufftComplex*>(p->d_grid2),
                           inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
    return sw2_transpose(p->d_grid2, grid, p->g.Kprod, nb, true, st);
}

// grid_bi = FFT(zero-pad(x * [sn] * [sens])): pad coil-major into the scratch, cuFFT there, ONE transposing pass out
int sweep2d_pad_fft(b200nufft_plan_t p, const float2* x, float2* grid, int nb, int apply_sn, int x_single,
                    const float2* sens, cudaStream_t st) {
    int rc = ensure_scratch2(p, nb);
    if (rc) return rc;
    rc = sw2_get_fft(p, nb);
    if (rc) return rc;
    const Geom& g = p->g;
    const int TB = 256;
    dim3 gr((unsigned)((g.Kprod + TB - 1) / TB), nb);
    k_sw2_scale_pad_cm<<<gr, TB, 0, st>>>(g.N[0], g.N[1], g.K[0], g.K[1], g.snoff[0], g.snoff[1], p->d_sn, x, p->d_grid2, nb,
                                          apply_sn, x_single, sens);
    LAUNCH_CHECK();
    CUFFT_TRY(cufftSetStream(p->fft_bi, st));
    CUFFT_TRY(cufftExecC2C(p->fft_bi, reinterpret_cast<cufftComplex*>(p->d_grid2), reinterpret_cast<cufftComplex*>(p->d_grid2),
                           CUFFT_FORWARD));
    return sw2_transpose(p->d_grid2, grid, g.Kprod, nb, true, st);
}

// x = crop(IFFT(grid_bi)) * f [combined over coils]: ONE transposing pass into the coil-major scratch, cuFFT there, crop
// from the scratch.  `cm_crop(scratch)` is the caller's coil-major crop (stages.cu).
int sweep2d_ifft_to_scratch(b200nufft_plan_t p, const float2* grid, int nb, cudaStream_t st) {
    int rc = ensure_scratch2(p, nb);
    if (rc) return rc;
    rc = sw2_get_fft(p, nb);
    if (rc) return rc;
    rc = sw2_transpose(grid, p->d_grid2, p->g.Kprod, nb, false, st);
    if (rc) return rc;
    CUFFT_TRY(cufftSetStream(p->fft_bi, st));
    CUFFT_TRY(cufftExecC2C(p->fft_bi, reinterpret_cast<cufftComplex*>(p->d_grid2), reinterpret_cast<cufftComplex*>(p->d_grid2),
                           CUFFT_INVERSE));
    return B200_OK;
}

// modulated: the grid is G[g] m0[g0] m1[g1] (fftbi_forward(..., modulate) writes it so)
int sweep2d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st, bool modulated) {
    int rc = sw2_attrs(p);
    if (rc) return rc;
    if (p->n_sw_work == 0) return B200_OK;
    dim3 gr((unsigned)((p->n_sw_work + SWARPS - 1) / SWARPS), (unsigned)((nb + NCL - 1) / NCL));
    if (modulated)
        k_sw2_interp<true><<<gr, SWARPS * 32, SWARPS * IW_PITCH, st>>>(sw2_geom(p->g, nb), p->d_sw_work, p->n_sw_work,
                                                                      p->d_sw_rec, p->d_mod, grid, y);
    else
        k_sw2_interp<false><<<gr, SWARPS * 32, SWARPS * IW_PITCH, st>>>(sw2_geom(p->g, nb), p->d_sw_work, p->n_sw_work,
                                                                       p->d_sw_rec, p->d_mod, grid, y);
    LAUNCH_CHECK();
    return B200_OK;
}

// modulated: the grid is left as G[g] m0[g0] m1[g1] (fftbi_inverse(..., demodulate) takes it)
int sweep2d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st, bool modulated) {
    int rc = sw2_attrs(p);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * p->g.Kprod * nb, st));
    if (p->n_sw_work == 0) return B200_OK;
    dim3 gr((unsigned)((p->n_sw_work + SWARPS - 1) / SWARPS), (unsigned)((nb + NCL - 1) / NCL));
    if (modulated)
        k_sw2_gridding<true><<<gr, SWARPS * 32, SWARPS * GW_PITCH, st>>>(sw2_geom(p->g, nb), p->d_sw_work, p->n_sw_work,
                                                                        p->d_sw_rec, p->d_mod, y, grid);
    else
        k_sw2_gridding<false><<<gr, SWARPS * 32, SWARPS * GW_PITCH, st>>>(sw2_geom(p->g, nb), p->d_sw_work, p->n_sw_work,
                                                                         p->d_sw_rec, p->d_mod, y, grid);
    LAUNCH_CHECK();
    return B200_OK;
}
