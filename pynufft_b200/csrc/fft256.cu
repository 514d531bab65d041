// Fused, pruned 3-D FFT passes for Kd = 256^3 (configuration 3): hand-written 256-point c2c FFT,
// 16 threads per transform, radix-16 x radix-16 in registers with one shared-memory exchange.
//
// Why: the padded image is non-zero only in the [0,N)^3 corner of the grid, and the adjoint only needs that
// corner of the inverse transform.  A library 3-D plan cannot exploit either; these passes read and write
// only what is needed and fuse the neighbouring element-wise kernels:
//   forward : F1 = scale (sn, coil map) + zero-pad + FFT along dim 2 on the N0*N1 image rows  (replaces
//                  cTensorMultiply + fill + cTensorCopy, re_subroutine.py:98-201, and 1/4 of a pass)
//             F2 = FFT along dim 1 on planes i0 < N0, reading rows i1 < N1 only
//             F3 = FFT along dim 0 on all columns, reading planes i0 < N0 only
//   inverse : I3 / I2 / I1 mirror them, writing planes i0 < N0, rows i1 < N1 and finally the cropped, scaled
//             image (replaces cTensorCopy(-1) + cTensorMultiply and the 1/prod(Kd) normalisation).
// Traffic per direction at N = 128: 352 MB instead of 3 x 268 MB (+ the separate pad / crop passes).
// Same arithmetic as the reference's unnormalised forward / 1/prod(Kd)-normalised inverse DFT
// (reikna FFT, nufft/_nufft_class_methods_device.py:246-249, 358, 431).
#include "common.cuh"

namespace {

constexpr int FN = 256;
constexpr int TPB = 256;            // threads per block = 16 transforms x 16 threads
constexpr int XPITCH = 17;          // smem exchange pitch (complex) per 16-value group
constexpr int TPITCH = 16 * XPITCH + 1;   // per-transform pitch, odd: the strided layout puts 16 transforms in one phase

template <int DIR>
__device__ __forceinline__ float2 mul_mi(float2 a) {   // a * (-i) for DIR = -1 (forward), a * (+i) for inverse
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

template <int DIR>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 t0 = make_float2(a.x + c.x, a.y + c.y), t1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 t2 = make_float2(b.x + d.x, b.y + d.y);
    const float2 t3 = mul_mi<DIR>(make_float2(b.x - d.x, b.y - d.y));
    a = make_float2(t0.x + t2.x, t0.y + t2.y);
    c = make_float2(t0.x - t2.x, t0.y - t2.y);
    b = make_float2(t1.x + t3.x, t1.y + t3.y);
    d = make_float2(t1.x - t3.x, t1.y - t3.y);
}

// in-register 16-point DFT, natural order in, natural order out:  X[m] = sum_k x[k] w^(DIR*k*m), w = e^{2 pi i/16}
template <int DIR>
__device__ __forceinline__ void dft16(float2 (&x)[16]) {
    // stage 1: 4 DFT-4 over k = j + 4q (q = 0..3) for j = 0..3
#pragma unroll
    for (int j = 0; j < 4; ++j) dft4<DIR>(x[j], x[j + 4], x[j + 8], x[j + 12]);
    // now x[j + 4p] = A_j[p], p = 0..3.  twiddle w16^(j*p)
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    const float sg = DIR < 0 ? -1.f : 1.f;
    const float2 w1 = make_float2(C1, sg * S1), w2 = make_float2(R2, sg * R2), w3 = make_float2(S1, sg * C1);
    const float2 w6 = make_float2(-R2, sg * R2), w9 = make_float2(-C1, -sg * S1);
    x[1 + 4] = cmul(x[1 + 4], w1);
    x[1 + 8] = cmul(x[1 + 8], w2);
    x[1 + 12] = cmul(x[1 + 12], w3);
    x[2 + 4] = cmul(x[2 + 4], w2);
    x[2 + 8] = mul_mi<DIR>(x[2 + 8]);            // w16^4 = -+i
    x[2 + 12] = cmul(x[2 + 12], w6);
    x[3 + 4] = cmul(x[3 + 4], w3);
    x[3 + 8] = cmul(x[3 + 8], w6);
    x[3 + 12] = cmul(x[3 + 12], w9);
    // stage 2: DFT-4 over j for each p: X[p + 4r] = sum_j B[j][p] w4^(j r)
#pragma unroll
    for (int p = 0; p < 4; ++p) dft4<DIR>(x[4 * p], x[4 * p + 1], x[4 * p + 2], x[4 * p + 3]);
    // x[4p + r] holds X[p + 4r]  -> reorder to natural order
    float2 y[16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int r = 0; r < 4; ++r) y[p + 4 * r] = x[4 * p + r];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = y[i];
}

// One 256-point transform held by 16 threads: thread t owns elements t + 16k (k = 0..15) on entry and
// outputs t + 16n on exit.  `sx` is this transform's exchange area (16 * XPITCH complex), `tw` the twiddle table
// w256^j (forward sign; conjugated on the fly for the inverse).
template <int DIR>
__device__ __forceinline__ void fft256_core(float2 (&v)[16], int t, float2* sx, const float2* __restrict__ tw) {
    dft16<DIR>(v);                                   // over k: v[m] = Y[t][m]
#pragma unroll
    for (int m = 1; m < 16; ++m) {
        float2 w = tw[(t * m) & 255];
        if (DIR > 0) w.y = -w.y;
        v[m] = cmul(v[m], w);
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) sx[m * XPITCH + t] = v[m];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = sx[t * XPITCH + k];      // thread t' = t now holds Y[k][t'], k = 0..15
    dft16<DIR>(v);                                   // over t: v[n] = X[t' + 16 n]
}

struct PassArgs {
    const float2* in;       // grid (or image for F1)
    float2* out;            // grid (or image for I1)
    const float* sn;        // concatenated scaling vectors
    const float2* sens;     // coil maps (image layout) or NULL
    int N0, N1, N2;         // image size
    int nb;                 // coils
    int x_single;           // F1: broadcast one image to all coils
    int apply_sn;           // F1: multiply by sn;  I1: 1 multiply, 2 divide, 0 none
    float scale;            // I1: extra factor (1/prod(Kd))
    int sn0, sn1, sn2;      // offsets of the per-dim scaling vectors
    const float2* mod;      // NULL, or the modulation table m_d[0..255] of the dimension the pass transforms: the
                            // inverse passes multiply their inputs by conj(m_d[g]) (the column-sweep gridding
                            // kernel, col3d.cu, leaves the grid phase-modulated), the forward passes multiply their
                            // outputs by m_d[g] (the column-sweep gather reads the modulated grid)
};

// PASS: 1 = F1, 2 = F2, 3 = F3, 4 = I3, 5 = I2, 6 = I1
#ifndef FFT256_MINB
#define FFT256_MINB 4       // resident CTAs per SM the register allocation aims at (64 registers; measured 1 / 4 / 5 / 6: inverse passes 79 / 76 / 85 / 113 us)
#endif
template <int PASS>
__global__ void __launch_bounds__(TPB, FFT256_MINB) k_fft256(PassArgs a, const float2* __restrict__ twg) {
    constexpr int DIR = PASS <= 3 ? -1 : 1;
    constexpr bool CONTIG = (PASS == 1 || PASS == 6);
    __shared__ float2 sx[16 * TPITCH];
    __shared__ float2 tw[256];
    const int tid = threadIdx.x;
    tw[tid] = twg[tid];
    const int c = blockIdx.y;
    const long long KK = (long long)FN * FN;
    // thread -> (transform within block q, position t)
    const int q = CONTIG ? tid >> 4 : tid & 15;
    const int t = CONTIG ? tid & 15 : tid >> 4;
    const long long b = (long long)blockIdx.x * 16 + q;       // transform index within the coil
    // geometry of this pass
    long long base;            // element offset of index 0 of the transform inside the coil's grid
    long long stride;          // element stride of the transform
    int nin, nout;             // inputs that are non-zero / outputs that are needed
    int i0 = 0, i1 = 0;
    if (PASS == 1 || PASS == 6) {          // rows (i0 < N0, i1 < N1), along dim 2
        i0 = (int)(b / a.N1);
        i1 = (int)(b - (long long)i0 * a.N1);
        base = ((long long)i0 * FN + i1) * FN;
        stride = 1;
        nin = PASS == 1 ? a.N2 : FN;
        nout = PASS == 1 ? FN : a.N2;
    } else if (PASS == 2 || PASS == 5) {   // (i0 < N0, i2), along dim 1
        i0 = (int)(b / FN);
        const int i2 = (int)(b - (long long)i0 * FN);
        base = (long long)i0 * KK + i2;
        stride = FN;
        nin = PASS == 2 ? a.N1 : FN;
        nout = PASS == 2 ? FN : a.N1;
    } else {                               // (i1, i2), along dim 0
        base = b;
        stride = KK;
        nin = PASS == 3 ? a.N0 : FN;
        nout = PASS == 3 ? FN : a.N0;
    }
    const float2* gin = a.in + (long long)c * KK * FN;
    float2* gout = a.out + (long long)c * KK * FN;

    float2 v[16];
    if (PASS == 1) {
        const long long nrow = ((long long)i0 * a.N1 + i1) * a.N2;
        const float s01 = a.apply_sn ? a.sn[a.sn0 + i0] * a.sn[a.sn1 + i1] : 1.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = t + 16 * k;
            v[k] = make_float2(0.f, 0.f);
            if (idx < nin) {
                const long long n = nrow + idx;
                float2 xv = a.x_single ? a.in[n] : a.in[n * a.nb + c];
                if (a.sens) xv = cmul(xv, a.sens[n * a.nb + c]);
                const float f = a.apply_sn ? s01 * a.sn[a.sn2 + idx] : 1.f;
                v[k] = make_float2(xv.x * f, xv.y * f);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = t + 16 * k;
            v[k] = make_float2(0.f, 0.f);
            if (idx < nin) v[k] = __ldg(gin + base + idx * stride);
        }
        if (DIR > 0 && a.mod) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = cmulc(__ldg(a.mod + t + 16 * k), v[k]);
        }
    }
    __syncthreads();                                  // twiddle table visible
    fft256_core<DIR>(v, t, sx + q * TPITCH, tw);
    if (PASS == 6) {
        const long long nrow = ((long long)i0 * a.N1 + i1) * a.N2;
        const float s01 = a.apply_sn ? a.sn[a.sn0 + i0] * a.sn[a.sn1 + i1] : 1.f;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int idx = t + 16 * n;
            if (idx < nout) {
                float f = a.scale;
                if (a.apply_sn == 1) f *= s01 * a.sn[a.sn2 + idx];
                if (a.apply_sn == 2) f /= s01 * a.sn[a.sn2 + idx];
                a.out[(nrow + idx) * a.nb + c] = make_float2(v[n].x * f, v[n].y * f);
            }
        }
    } else {
        if (DIR < 0 && a.mod) {
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = cmul(__ldg(a.mod + t + 16 * n), v[n]);
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int idx = t + 16 * n;
            if (idx < nout) gout[base + idx * stride] = v[n];
        }
    }
}

__global__ void k_combine_coils_few(const float2* __restrict__ xc, const float2* __restrict__ sens,
                                    float2* __restrict__ s, long long N, int nb) {
    const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= N) return;
    float2 acc = make_float2(0.f, 0.f);
    for (int c = 0; c < nb; ++c) {
        float2 v = xc[n * nb + c];
        if (sens) v = cmulc(sens[n * nb + c], v);
        acc.x += v.x;
        acc.y += v.y;
    }
    const float f = 1.0f / (float)nb;
    s[n] = make_float2(acc.x * f, acc.y * f);
}

// one warp per pixel, lanes over the coils (contiguous in the image layout): s[n] = (1/nb) sum_c conj(sens[n,c]) xc[n,c]
__global__ void k_combine_coils(const float2* __restrict__ xc, const float2* __restrict__ sens,
                                float2* __restrict__ s, long long N, int nb) {
    const int lane = threadIdx.x & 31;
    const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    float2 acc = make_float2(0.f, 0.f);
    for (int c = lane; c < nb; c += 32) {
        float2 v = xc[n * nb + c];
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
        const float f = 1.0f / (float)nb;
        s[n] = make_float2(acc.x * f, acc.y * f);
    }
}

}  // namespace

bool fft256_supported(const Geom& g) {
    if (g.ndim != 3) return false;
    for (int d = 0; d < 3; ++d)
        if (g.K[d] != FN || g.N[d] > FN || g.N[d] < 1) return false;
    // whole 16-transform blocks: N0*N1 multiple of 16 (F1/I1); the strided passes always are
    return (g.N[0] * g.N[1]) % 16 == 0;
}

static int ensure_tw(b200nufft_plan_t p) {
    if (p->d_tw256) return B200_OK;
    float2 h[256];
    for (int j = 0; j < 256; ++j) {
        const double ang = -2.0 * M_PI * j / 256.0;           // forward sign
        h[j] = make_float2((float)cos(ang), (float)sin(ang));
    }
    CUDA_TRY(cudaMalloc(&p->d_tw256, sizeof(h)));
    CUDA_TRY(cudaMemcpy(p->d_tw256, h, sizeof(h), cudaMemcpyHostToDevice));
    return B200_OK;
}

// grid <- FFT3(zero-pad(x * sn * sens))
int fft256_forward(b200nufft_plan_t p, const float2* x, float2* grid, int nb, int apply_sn, int x_single,
                   const float2* sens, bool modulated, cudaStream_t st) {
    int rc = ensure_tw(p);
    if (rc) return rc;
    const Geom& g = p->g;
    PassArgs a{};
    a.sn = p->d_sn;
    a.sens = sens;
    a.N0 = g.N[0]; a.N1 = g.N[1]; a.N2 = g.N[2];
    a.nb = nb;
    a.x_single = x_single;
    a.apply_sn = apply_sn;
    a.scale = 1.f;
    a.sn0 = g.snoff[0]; a.sn1 = g.snoff[1]; a.sn2 = g.snoff[2];
    a.in = x;
    a.out = grid;
    const float2* mod = (modulated && p->d_mod) ? p->d_mod : nullptr;
    a.mod = mod ? mod + 2 * FN : nullptr;
    k_fft256<1><<<dim3(g.N[0] * g.N[1] / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    a.in = grid;
    a.mod = mod ? mod + FN : nullptr;
    k_fft256<2><<<dim3(g.N[0] * FN / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    a.mod = mod;
    k_fft256<3><<<dim3(FN * FN / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    return B200_OK;
}

// x <- crop(IFFT3(grid)) * f * scale ; the grid is overwritten (partially transformed)
int fft256_inverse(b200nufft_plan_t p, float2* grid, float2* x, int nb, int mode, float scale, bool modulated,
                   cudaStream_t st) {
    int rc = ensure_tw(p);
    if (rc) return rc;
    const Geom& g = p->g;
    PassArgs a{};
    a.sn = p->d_sn;
    a.N0 = g.N[0]; a.N1 = g.N[1]; a.N2 = g.N[2];
    a.nb = nb;
    a.apply_sn = mode;
    a.scale = scale;
    a.sn0 = g.snoff[0]; a.sn1 = g.snoff[1]; a.sn2 = g.snoff[2];
    a.in = grid;
    a.out = grid;
    const float2* mod = (modulated && p->d_mod) ? p->d_mod : nullptr;
    a.mod = mod;
    k_fft256<4><<<dim3(FN * FN / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    a.mod = mod ? mod + FN : nullptr;
    k_fft256<5><<<dim3(g.N[0] * FN / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    a.out = x;
    a.mod = mod ? mod + 2 * FN : nullptr;
    k_fft256<6><<<dim3(g.N[0] * g.N[1] / 16, nb), TPB, 0, st>>>(a, p->d_tw256);
    LAUNCH_CHECK();
    return B200_OK;
}

int combine_coils(const float2* xc, const float2* sens, float2* s, long long N, int nb, cudaStream_t st) {
    const int TB = 256;
    if (nb < 8) {       // few coils: one thread per pixel
        k_combine_coils_few<<<(unsigned)((N + TB - 1) / TB), TB, 0, st>>>(xc, sens, s, N, nb);
    } else {
        k_combine_coils<<<(unsigned)((N * 32 + TB - 1) / TB), TB, 0, st>>>(xc, sens, s, N, nb);
    }
    LAUNCH_CHECK();
    return B200_OK;
}
