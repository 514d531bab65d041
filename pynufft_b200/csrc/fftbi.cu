// Fused, pruned 2-D FFT passes on BATCH-INNERMOST grids k[K0][K1][nb] (the layout of the 2-D multi-coil kernels,
// sweep2d.cu) for power-of-two grid sizes 64 .. 1024 (configurations 2 and 4: Kd = 512^2, 32 coils).
//
// cuFFT's strided-batch plans are slow on this layout (282 us for 32 x 512^2 against 56 us coil-major), and going
// through a coil-major scratch costs a transposing pass each way.  These passes work on the layout directly: a CTA
// takes one line (a row or a column of the grid) for 16 coils -- every global access is a full 128-byte run -- and
// transforms it in shared memory (in-place decimation in frequency, radix 8 / 4 / 2, twiddles from a table; the
// digit-reversed output order is undone by the final store).  As in fft256.cu the padding / cropping is fused and
// pruned:
//   forward : A = scale (sn, coil map) + zero-pad + FFT along dim 1 on the N0 image rows only (replaces
//                 cTensorMultiply + fill + cTensorCopy, re_subroutine.py:98-201)
//             B = FFT along dim 0 on all K1 columns, reading rows i0 < N0 only
//   inverse : A' = FFT along dim 0, writing rows i0 < N0 only;  B' = FFT along dim 1 on those rows, writing the cropped,
//                 scaled image (cTensorCopy(-1) + cTensorMultiply and the 1/prod(Kd) normalisation)
// Same arithmetic as the reference's unnormalised forward / normalised inverse DFT (reikna FFT,
// nufft/_nufft_class_methods_device.py:246-249, 358, 431).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

constexpr int NCO = 16;             // coils per CTA
constexpr int FTB = 256;            // threads per CTA: 16 coils x 16 butterfly groups

template <int DIR>
__device__ __forceinline__ float2 mul_mi(float2 a) {   // a * (-i) forward, a * (+i) inverse
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
template <int DIR>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
    const float2 t = make_float2(a.x - b.x, a.y - b.y);
    a = make_float2(a.x + b.x, a.y + b.y);
    b = t;
}
// natural order in, natural order out
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
template <int DIR>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    // even / odd split: E = DFT4(v0, v2, v4, v6), O = DFT4(v1, v3, v5, v7); X[k] = E[k] + w8^k O[k], X[k+4] = E[k] - ..
    dft4<DIR>(v[0], v[2], v[4], v[6]);
    dft4<DIR>(v[1], v[3], v[5], v[7]);
    constexpr float R2 = 0.70710678118654752f;
    const float sg = DIR < 0 ? -1.f : 1.f;
    const float2 o0 = v[1];
    const float2 o1 = cmul(v[3], make_float2(R2, sg * R2));
    const float2 o2 = mul_mi<DIR>(v[5]);
    const float2 o3 = cmul(v[7], make_float2(-R2, sg * R2));
    const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = make_float2(e0.x + o0.x, e0.y + o0.y);
    v[4] = make_float2(e0.x - o0.x, e0.y - o0.y);
    v[1] = make_float2(e1.x + o1.x, e1.y + o1.y);
    v[5] = make_float2(e1.x - o1.x, e1.y - o1.y);
    v[2] = make_float2(e2.x + o2.x, e2.y + o2.y);
    v[6] = make_float2(e2.x - o2.x, e2.y - o2.y);
    v[3] = make_float2(e3.x + o3.x, e3.y + o3.y);
    v[7] = make_float2(e3.x - o3.x, e3.y - o3.y);
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct BiPass {
    const float2* in;       // grid, or the image for the forward pass A
    float2* out;            // grid, or the image for the inverse pass B'
    int n;                  // transform length
    int nrad;               // number of stages
    int rad[4];             // radices, first applied first
    int nin, nout;          // inputs that are non-zero / outputs that are needed
    long long estride;      // element stride of the transform inside the grid (in complex elements, before the coil)
    long long lstride;      // line stride
    int nb;                 // coils
    const float2* tw;       // twiddle table exp(-2 pi i t / n), t = 0 .. n-1
    const int* rev;         // rev[pos] = frequency index held at position pos after the in-place DIF stages
    // image side (mode 1 / 2)
    int mode;               // 0 grid -> grid, 1 image -> grid (scale, pad), 2 grid -> image (crop, scale)
    int Nline;              // image extent along the transformed dimension (N1)
    const float* sn_line;   // scaling vector along the line dimension (sn1), sn_other[line] along the other (sn0)
    const float* sn_other;
    const float2* sens;     // coil maps (image layout) or NULL
    int x_single, apply_sn; // as scale_pad / crop_scale (apply_sn: 0 none, 1 multiply, 2 divide)
    float scale;
    // mode 0 along dim 0 only: phase modulation of the grid (sweep2d.cu works on G[g] m0[g0] m1[g1]): the forward pass
    // multiplies its outputs by mod_t[f] * mod_l[line], the inverse pass its inputs by the conjugate
    const float2* mod_t;    // m0, indexed by the position along the transform; NULL = none
    const float2* mod_l;    // m1, indexed by the line
};

// one in-place DIF stage: radix R on sub-transforms of length L inside a line of N points (all compile-time)
template <int DIR, int R, int L, int N>
__device__ __forceinline__ void stage(float2* __restrict__ buf, const float2* __restrict__ tw, int c, int gi) {
    constexpr int M = L / R;            // sub-transform length after this stage
    constexpr int NBF = N / R;          // butterflies per line
    constexpr int TSTEP = N / L;        // w_L^x = tw[x * N / L]
#pragma unroll
    for (int bf0 = 0; bf0 < NBF; bf0 += FTB / NCO) {
        const int bf = bf0 + gi;
        if (NBF < FTB / NCO && bf >= NBF) break;
        const int j = bf / M, k = bf % M;
        float2* p = buf + (j * L + k) * NCO + c;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * M * NCO];
        if (R == 8) {
            float2 (&v8)[8] = reinterpret_cast<float2(&)[8]>(v);
            dft8<DIR>(v8);
        } else if (R == 4) {
            dft4<DIR>(v[0], v[1 % R], v[2 % R], v[3 % R]);
        } else {
            dft2<DIR>(v[0], v[1 % R]);
        }
        if (M > 1) {                    // the last stage has no twiddles
#pragma unroll
            for (int q = 1; q < R; ++q) {
                float2 w = tw[q * k * TSTEP];
                if (DIR > 0) w.y = -w.y;
                v[q] = cmul(v[q], w);
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) p[q * M * NCO] = v[q];
    }
    __syncthreads();
}

template <int DIR, int LOGN>
__device__ __forceinline__ void stages(float2* buf, const float2* tw, int c, int gi) {
    constexpr int N = 1 << LOGN;
    if (LOGN == 6) { stage<DIR, 8, 64, N>(buf, tw, c, gi); stage<DIR, 8, 8, N>(buf, tw, c, gi); }
    if (LOGN == 7) { stage<DIR, 8, 128, N>(buf, tw, c, gi); stage<DIR, 4, 16, N>(buf, tw, c, gi); stage<DIR, 4, 4, N>(buf, tw, c, gi); }
    if (LOGN == 8) { stage<DIR, 8, 256, N>(buf, tw, c, gi); stage<DIR, 8, 32, N>(buf, tw, c, gi); stage<DIR, 4, 4, N>(buf, tw, c, gi); }
    if (LOGN == 9) { stage<DIR, 8, 512, N>(buf, tw, c, gi); stage<DIR, 8, 64, N>(buf, tw, c, gi); stage<DIR, 8, 8, N>(buf, tw, c, gi); }
    if (LOGN == 10) {
        stage<DIR, 8, 1024, N>(buf, tw, c, gi); stage<DIR, 8, 128, N>(buf, tw, c, gi); stage<DIR, 8, 16, N>(buf, tw, c, gi);
        stage<DIR, 2, 2, N>(buf, tw, c, gi);
    }
}

template <int DIR, int LOGN>
__global__ void __launch_bounds__(FTB) k_fftbi(BiPass a) {
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(16) unsigned char fsm[];
    float2* buf = reinterpret_cast<float2*>(fsm);                       // [N][NCO]
    float2* tw = buf + N * NCO;                                         // [N]
    const int t = threadIdx.x, c = t & (NCO - 1), gi = t >> 4;          // coil inside the block, butterfly group
    const int line = blockIdx.x;
    float2* aux = tw + N;                                               // [N]: per-point factors of the image side
    int* rv = reinterpret_cast<int*>(aux + N);                          // [N]: digit-reversal table
    float* fsv = reinterpret_cast<float*>(aux);                         // mode 2: real factors
    for (int i = t; i < N; i += FTB) {
        tw[i] = __ldg(a.tw + i);
        rv[i] = __ldg(a.rev + i);
    }
    const int cb = blockIdx.y * NCO;
    const bool cact = cb + c < a.nb;
    // ---- load (natural order): every element of the line goes global -> shared memory by an 8-byte cp.async, so that
    //      all loads of the CTA are in flight at once (through registers the load phase was a chain of DRAM round
    //      trips: `long_scoreboard` was the top stall of these passes); the per-point factors of the image side are
    //      gathered into shared memory while the copies fly ----
    if (a.mode == 1) {
        // image -> grid: the coil-dependent factor (coil map, or the multi-coil image) is copied, then scaled in place
        const long long nrow = (long long)line * a.Nline;
        const float2* src = a.sens ? a.sens : (a.x_single ? nullptr : a.in);
        for (int i = gi; i < N; i += FTB / NCO) {
            float2* d = buf + i * NCO + c;
            if (i < a.nin && cact && src) cp_async8(d, src + (nrow + i) * a.nb + cb + c);
            else *d = make_float2(i < a.nin && cact ? 1.f : 0.f, 0.f);
        }
        const float s0 = a.apply_sn ? a.sn_other[line] : 1.f;
        for (int i = t; i < a.nin; i += FTB) {              // aux[i] = [x[i]] * sn: the same for every coil
            const float f = a.apply_sn ? s0 * a.sn_line[i] : 1.f;
            const float2 xv = a.x_single ? __ldg(a.in + nrow + i) : make_float2(1.f, 0.f);
            aux[i] = make_float2(xv.x * f, xv.y * f);
        }
        cp_async_wait_all();
        __syncthreads();
        for (int i = gi; i < a.nin; i += FTB / NCO) {
            float2* d = buf + i * NCO + c;
            const float2 v = cmul(*d, aux[i]);
            *d = cact ? v : make_float2(0.f, 0.f);
        }
    } else {
        const float2* src = a.in + (long long)line * a.lstride * a.nb + cb + c;
        const long long es = a.estride * a.nb;
        for (int i = gi; i < N; i += FTB / NCO) {
            float2* d = buf + i * NCO + c;
            if (i < a.nin && cact) cp_async8(d, src + i * es);
            else *d = make_float2(0.f, 0.f);
        }
        if (a.mode == 0 && a.mod_t) {                       // aux[i] = m_t[i] m_l[line]
            const float2 ml = __ldg(a.mod_l + line);
            for (int i = t; i < N; i += FTB) aux[i] = cmul(__ldg(a.mod_t + i), ml);
        }
        if (a.mode == 2) {                                  // fsv[f] = scale * (sn | 1 / sn) of image column f
            const float s0 = a.sn_other[line];
            for (int f = t; f < a.nout; f += FTB) {
                float fs = a.scale;
                const float sv = s0 * a.sn_line[f];
                if (a.apply_sn == 1) fs *= sv;
                if (a.apply_sn == 2) fs /= sv;
                fsv[f] = fs;
            }
        }
        cp_async_wait_all();
        if (DIR > 0 && a.mode == 0 && a.mod_t) {            // inverse: the scatter left the grid modulated
            __syncthreads();
            for (int i = gi; i < a.nin; i += FTB / NCO) {
                float2* d = buf + i * NCO + c;
                *d = cmulc(aux[i], *d);
            }
        }
    }
    __syncthreads();
    stages<DIR, LOGN>(buf, tw, c, gi);
    // ---- store: position pos holds frequency rev[pos] ----
    if (a.mode == 2) {
        const long long nrow = (long long)line * a.Nline;
#pragma unroll 4
        for (int pos = gi; pos < N; pos += FTB / NCO) {
            const int f = rv[pos];
            if (f < a.nout && cact) {
                const float fs = fsv[f];
                const float2 v = buf[pos * NCO + c];
                a.out[(nrow + f) * a.nb + cb + c] = make_float2(v.x * fs, v.y * fs);
            }
        }
    } else {
        float2* dst = a.out + (long long)line * a.lstride * a.nb + cb + c;
        const long long es = a.estride * a.nb;
#pragma unroll 8
        const bool modout = DIR < 0 && a.mod_t;             // forward: the gather reads the modulated grid
        for (int pos = gi; pos < N; pos += FTB / NCO) {
            const int f = rv[pos];
            if (f < a.nout && cact) dst[f * es] = modout ? cmul(buf[pos * NCO + c], aux[f]) : buf[pos * NCO + c];
        }
    }
}

// line buffer [n][NCO] + twiddles [n] + image-side factors [n] (float2 each) + digit-reversal table [n] (int)
size_t fbi_smem(int n) { return sizeof(float2) * (size_t)n * (NCO + 2) + sizeof(int) * (size_t)n; }

bool pow2_ok(int k) { return k >= 64 && k <= 1024 && (k & (k - 1)) == 0; }

int factor(int n, int* rad) {       // the radix lists of stages<> above
    int ns = 0;
    switch (n) {
        case 64: rad[0] = 8; rad[1] = 8; ns = 2; break;
        case 128: rad[0] = 8; rad[1] = 4; rad[2] = 4; ns = 3; break;
        case 256: rad[0] = 8; rad[1] = 8; rad[2] = 4; ns = 3; break;
        case 512: rad[0] = 8; rad[1] = 8; rad[2] = 8; ns = 3; break;
        case 1024: rad[0] = 8; rad[1] = 8; rad[2] = 8; rad[3] = 2; ns = 4; break;
    }
    return ns;
}

}  // namespace

bool fftbi_supported(const Geom& g) {
    return g.ndim == 2 && pow2_ok(g.K[0]) && pow2_ok(g.K[1]) && g.N[0] <= g.K[0] && g.N[1] <= g.K[1];
}

// twiddle and digit-reversal tables of both dimensions: [tw0 (K0) | tw1 (K1)] and [rev0 (K0) | rev1 (K1)]
static int ensure_tables(b200nufft_plan_t p) {
    if (p->d_fbi_tw) return B200_OK;
    const Geom& g = p->g;
    std::vector<float2> tw(g.K[0] + g.K[1]);
    std::vector<int> rev(g.K[0] + g.K[1]);
    int off = 0;
    for (int d = 0; d < 2; ++d) {
        const int n = g.K[d];
        int rad[4];
        const int ns = factor(n, rad);
        for (int t = 0; t < n; ++t) {
            const double ang = -2.0 * M_PI * t / n;
            tw[off + t] = make_float2((float)cos(ang), (float)sin(ang));
        }
        for (int pos = 0; pos < n; ++pos) {   // pos = sum_s q_s m_s  ->  f = q_1 + r_1 (q_2 + r_2 (...))
            int m = n, f = 0, mul = 1, rem = pos;
            for (int s = 0; s < ns; ++s) {
                m /= rad[s];
                const int q = rem / m;
                rem -= q * m;
                f += q * mul;
                mul *= rad[s];
            }
            rev[off + pos] = f;
        }
        off += n;
    }
    CUDA_TRY(cudaMalloc(&p->d_fbi_tw, sizeof(float2) * tw.size()));
    CUDA_TRY(cudaMalloc(&p->d_fbi_rev, sizeof(int) * rev.size()));
    CUDA_TRY(cudaMemcpy(p->d_fbi_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_fbi_rev, rev.data(), sizeof(int) * rev.size(), cudaMemcpyHostToDevice));
    const int smem = (int)fbi_smem(1024);
#define FBI_ATTR(LG)                                                                                         \
    CUDA_TRY(cudaFuncSetAttribute(k_fftbi<-1, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));     \
    CUDA_TRY(cudaFuncSetAttribute(k_fftbi<1, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FBI_ATTR(6) FBI_ATTR(7) FBI_ATTR(8) FBI_ATTR(9) FBI_ATTR(10)
#undef FBI_ATTR
    return B200_OK;
}

static BiPass base_pass(b200nufft_plan_t p, int dim, int nb) {
    const Geom& g = p->g;
    BiPass a{};
    a.n = g.K[dim];
    a.nrad = factor(a.n, a.rad);
    a.nb = nb;
    a.tw = p->d_fbi_tw + (dim == 0 ? 0 : g.K[0]);
    a.rev = p->d_fbi_rev + (dim == 0 ? 0 : g.K[0]);
    a.estride = dim == 0 ? g.K[1] : 1;
    a.lstride = dim == 0 ? 1 : g.K[1];
    a.scale = 1.f;
    return a;
}

template <int DIR>
static int launch(const BiPass& a, int nlines, int nb, cudaStream_t st) {
    dim3 gr((unsigned)nlines, (unsigned)((nb + NCO - 1) / NCO));
    size_t smem = fbi_smem(a.n);
    // tuning knob: resident CTAs per SM (more shared memory per CTA = fewer of them); the number of lines is fixed, so
    // the best occupancy is the one that makes the number of waves an integer
    static const int want_ctas = [] { const char* e = getenv("B200NUFFT_FBI_CTAS"); return e ? atoi(e) : 0; }();
    if (want_ctas > 0) smem = std::max(smem, std::min<size_t>((size_t)(227 * 1024) / want_ctas - 1024, fbi_smem(1024)));
    switch (a.n) {
        case 64: k_fftbi<DIR, 6><<<gr, FTB, smem, st>>>(a); break;
        case 128: k_fftbi<DIR, 7><<<gr, FTB, smem, st>>>(a); break;
        case 256: k_fftbi<DIR, 8><<<gr, FTB, smem, st>>>(a); break;
        case 512: k_fftbi<DIR, 9><<<gr, FTB, smem, st>>>(a); break;
        default: k_fftbi<DIR, 10><<<gr, FTB, smem, st>>>(a); break;
    }
    LAUNCH_CHECK();
    return B200_OK;
}

// grid_bi <- FFT2(zero-pad(x * [sn] * [sens]))
// modulate: the grid is written as G[g] m0[g0] m1[g1] (what sweep2d_interp(..., modulated) reads)
int fftbi_forward(b200nufft_plan_t p, const float2* x, float2* grid, int nb, int apply_sn, int x_single,
                  const float2* sens, cudaStream_t st, bool modulate) {
    int rc = ensure_tables(p);
    if (rc) return rc;
    const Geom& g = p->g;
    BiPass a = base_pass(p, 1, nb);             // A: along dim 1 on the N0 image rows, from the image
    a.in = x;
    a.out = grid;
    a.mode = 1;
    a.nin = g.N[1];
    a.nout = g.K[1];
    a.Nline = g.N[1];
    a.sn_line = p->d_sn + g.snoff[1];
    a.sn_other = p->d_sn + g.snoff[0];
    a.sens = sens;
    a.x_single = x_single;
    a.apply_sn = apply_sn;
    rc = launch<-1>(a, g.N[0], nb, st);
    if (rc) return rc;
    BiPass b = base_pass(p, 0, nb);             // B: along dim 0 on all columns, rows i0 < N0 are non-zero
    b.in = grid;
    b.out = grid;
    b.nin = g.N[0];
    b.nout = g.K[0];
    if (modulate) {
        b.mod_t = p->d_mod;
        b.mod_l = p->d_mod + g.K[0];
    }
    return launch<-1>(b, g.K[1], nb, st);
}

// x <- crop(IFFT2(grid_bi)) * f * scale (image layout, all coils); the grid is overwritten (partially transformed)
// demodulate: the grid is G[g] m0[g0] m1[g1] (left so by sweep2d_gridding(..., modulated))
int fftbi_inverse(b200nufft_plan_t p, float2* grid, float2* x, int nb, int mode, float scale, cudaStream_t st,
                  bool demodulate) {
    int rc = ensure_tables(p);
    if (rc) return rc;
    const Geom& g = p->g;
    BiPass a = base_pass(p, 0, nb);             // A': along dim 0 on all columns, rows i0 < N0 are needed
    a.in = grid;
    a.out = grid;
    a.nin = g.K[0];
    a.nout = g.N[0];
    if (demodulate) {
        a.mod_t = p->d_mod;
        a.mod_l = p->d_mod + g.K[0];
    }
    rc = launch<1>(a, g.K[1], nb, st);
    if (rc) return rc;
    BiPass b = base_pass(p, 1, nb);             // B': along dim 1 on the N0 rows, cropped + scaled into the image
    b.in = grid;
    b.out = x;
    b.mode = 2;
    b.nin = g.K[1];
    b.nout = g.N[1];
    b.Nline = g.N[1];
    b.sn_line = p->d_sn + g.snoff[1];
    b.sn_other = p->d_sn + g.snoff[0];
    b.apply_sn = mode;
    b.scale = scale;
    return launch<1>(b, g.N[0], nb, st);
}
