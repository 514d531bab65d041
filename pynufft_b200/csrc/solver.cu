// Fused vector kernels for the device solvers (linalg/solve_device.py:351-481 CG, :74-275
// L1TVOLS).  The reference launches cMultiplyConjVec / cMultiplyScalar / cAddVec / cDiff / cHypot /
// cAnisoShrink / cMultiplyVec (src/re_subroutine.py) plus reikna array arithmetic one pass at a
// time and fetches alpha/beta to the host each iteration; here every CG step is two streaming
// kernels and the scalars never leave the device.
#include "common.cuh"

#define RED_TB 256
#define RED_BLOCKS (148 * 4)

__device__ __forceinline__ void block_reduce_add2(double re, double im, double* out) {
    __shared__ double sre[RED_TB / 32], sim[RED_TB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sre[w] = re; sim[w] = im; }
    __syncthreads();
    if (w == 0) {
        re = l < RED_TB / 32 ? sre[l] : 0.0;
        im = l < RED_TB / 32 ? sim[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
        }
        if (l == 0) {
            atomicAdd(out, re);
            atomicAdd(out + 1, im);
        }
    }
}

// ---- streaming skeleton -------------------------------------------------------------------------------------------
// The vector kernels below walk n complex elements in 16-byte pieces (two elements), SU pieces per thread in flight:
// all loads of a trip are issued before the first use (with one 8-byte load per trip the kernels sat at 4.4 - 5.6 TB/s
// of the 6.5 TB/s copy bandwidth).  n4 = number of 16-byte pieces (0 when a pointer is not 16-byte aligned); the
// elements from 2 * n4 on are done one by one.
constexpr int SU = 4;
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 pack4(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define STREAM_LOOP(N4, PIECES, SINGLE)                                                            \
    {                                                                                              \
        const long long stride = (long long)gridDim.x * blockDim.x;                               \
        long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;                           \
        for (; i + (SU - 1) * stride < (N4); i += SU * stride) { PIECES(SU) }                      \
        for (; i < (N4); i += stride) { PIECES(1) }                                                \
        for (long long j = 2 * (N4) + blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += stride) { SINGLE }  \
    }

__device__ __forceinline__ void dot_acc(float2 x, float2 y, float& re, float& im) {
    re = fmaf(x.x, y.x, fmaf(x.y, y.y, re));
    im = fmaf(x.x, y.y, fmaf(-x.y, y.x, im));
}

__global__ void __launch_bounds__(RED_TB) k_dotc(const float2* __restrict__ a, const float2* __restrict__ b,
                                                 long long n, long long n4, double* __restrict__ out) {
    double dre = 0.0, dim_ = 0.0;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#define DOT_PIECES(U)                                                                              \
    float4 xv[U], yv[U];                                                                           \
    _Pragma("unroll") for (int u = 0; u < U; ++u) { xv[u] = a4[i + u * stride]; yv[u] = b4[i + u * stride]; }  \
    float re = 0.f, im = 0.f;                       /* bounded f32 partials: 2 U elements */        \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        dot_acc(lo2(xv[u]), lo2(yv[u]), re, im);                                                   \
        dot_acc(hi2(xv[u]), hi2(yv[u]), re, im);                                                   \
    }                                                                                              \
    dre += re; dim_ += im;
#define DOT_SINGLE                                                                                 \
    float re = 0.f, im = 0.f;                                                                      \
    dot_acc(a[j], b[j], re, im);                                                                   \
    dre += re; dim_ += im;
    STREAM_LOOP(n4, DOT_PIECES, DOT_SINGLE)
#undef DOT_PIECES
#undef DOT_SINGLE
    block_reduce_add2(dre, dim_, out);
}

__device__ __forceinline__ float2 cdivd(const double* num, const double* den) {  // (double) num/den -> c64
    double a = num[0], b = num[1], c = den[0], d = den[1];
    double m = c * c + d * d;
    return make_float2((float)((a * c + b * d) / m), (float)((b * c - a * d) / m));
}

__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float nrm2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

// r = b - Ax ; p = r ; rsold += sum |r|^2
__global__ void __launch_bounds__(RED_TB)
k_cg_init(const float2* __restrict__ b, const float2* __restrict__ Ax, float2* __restrict__ r,
          float2* __restrict__ p, double* __restrict__ rsold, long long n, long long n4) {
    double dre = 0.0;
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const float4* A4 = reinterpret_cast<const float4*>(Ax);
    float4* r4 = reinterpret_cast<float4*>(r);
    float4* p4 = reinterpret_cast<float4*>(p);
#define INIT_PIECES(U)                                                                             \
    float4 bv[U], av[U];                                                                           \
    _Pragma("unroll") for (int u = 0; u < U; ++u) { bv[u] = b4[i + u * stride]; av[u] = A4[i + u * stride]; }  \
    float re = 0.f;                                                                                \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        const float2 r0 = sub2(lo2(bv[u]), lo2(av[u])), r1 = sub2(hi2(bv[u]), hi2(av[u]));         \
        const float4 rv = pack4(r0, r1);                                                           \
        r4[i + u * stride] = rv;                                                                   \
        p4[i + u * stride] = rv;                                                                   \
        re += nrm2(r0) + nrm2(r1);                                                                 \
    }                                                                                              \
    dre += re;
#define INIT_SINGLE                                                                                \
    const float2 rv = sub2(b[j], Ax[j]);                                                           \
    r[j] = rv;                                                                                     \
    p[j] = rv;                                                                                     \
    dre += nrm2(rv);
    STREAM_LOOP(n4, INIT_PIECES, INIT_SINGLE)
#undef INIT_PIECES
#undef INIT_SINGLE
    block_reduce_add2(dre, 0.0, rsold);
}

// alpha = rsold / pAp ; x += alpha p ; r -= alpha Ap ; rsnew += sum |r|^2
__global__ void __launch_bounds__(RED_TB)
k_cg_update_xr(float2* __restrict__ x, float2* __restrict__ r, const float2* __restrict__ p,
               const float2* __restrict__ Ap, const double* __restrict__ rsold, const double* __restrict__ pAp,
               double* __restrict__ rsnew, long long n, long long n4) {
    const float2 alpha = cdivd(rsold, pAp);
    double dre = 0.0;
    float4* x4 = reinterpret_cast<float4*>(x);
    float4* r4 = reinterpret_cast<float4*>(r);
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* q4 = reinterpret_cast<const float4*>(Ap);
#define XR_PIECES(U)                                                                               \
    float4 pv[U], qv[U], xv[U], rv[U];                                                             \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        pv[u] = p4[i + u * stride]; qv[u] = q4[i + u * stride];                                    \
        xv[u] = x4[i + u * stride]; rv[u] = r4[i + u * stride];                                    \
    }                                                                                              \
    float re = 0.f;                                                                                \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        const float2 x0 = add2(lo2(xv[u]), cmul(alpha, lo2(pv[u]))), x1 = add2(hi2(xv[u]), cmul(alpha, hi2(pv[u]))); \
        const float2 r0 = sub2(lo2(rv[u]), cmul(alpha, lo2(qv[u]))), r1 = sub2(hi2(rv[u]), cmul(alpha, hi2(qv[u]))); \
        x4[i + u * stride] = pack4(x0, x1);                                                        \
        r4[i + u * stride] = pack4(r0, r1);                                                        \
        re += nrm2(r0) + nrm2(r1);                                                                 \
    }                                                                                              \
    dre += re;
#define XR_SINGLE                                                                                  \
    const float2 x0 = add2(x[j], cmul(alpha, p[j])), r0 = sub2(r[j], cmul(alpha, Ap[j]));          \
    x[j] = x0;                                                                                     \
    r[j] = r0;                                                                                     \
    dre += nrm2(r0);
    STREAM_LOOP(n4, XR_PIECES, XR_SINGLE)
#undef XR_PIECES
#undef XR_SINGLE
    block_reduce_add2(dre, 0.0, rsnew);
}

// beta = rsnew / rsold ; p = r + beta p
__global__ void __launch_bounds__(RED_TB)
k_cg_update_p(float2* __restrict__ p, const float2* __restrict__ r, const double* __restrict__ rsnew,
              const double* __restrict__ rsold, long long n, long long n4) {
    const float2 beta = cdivd(rsnew, rsold);
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* r4 = reinterpret_cast<const float4*>(r);
#define P_PIECES(U)                                                                                \
    float4 pv[U], rv[U];                                                                           \
    _Pragma("unroll") for (int u = 0; u < U; ++u) { pv[u] = p4[i + u * stride]; rv[u] = r4[i + u * stride]; }  \
    _Pragma("unroll") for (int u = 0; u < U; ++u)                                                  \
        p4[i + u * stride] = pack4(add2(lo2(rv[u]), cmul(beta, lo2(pv[u]))), add2(hi2(rv[u]), cmul(beta, hi2(pv[u]))));
#define P_SINGLE p[j] = add2(r[j], cmul(beta, p[j]));
    STREAM_LOOP(n4, P_PIECES, P_SINGLE)
#undef P_PIECES
#undef P_SINGLE
}

__global__ void k_cdiv(float2* __restrict__ a, const float2* __restrict__ b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float2 x = a[i], y = b[i];
        float m = y.x * y.x + y.y * y.y;
        a[i] = make_float2((x.x * y.x + x.y * y.y) / m, (x.y * y.x - x.x * y.y) / m);
    }
}

__global__ void k_cmul(float2* __restrict__ a, const float2* __restrict__ b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = cmul(a[i], b[i]);
}

// out = a x + b y with host scalars (the vector updates of the Krylov recurrences, pynufft_b200/krylov.py);
// HAS_Y = false: out = a x and y is never read.  out may alias x or y.
template <bool HAS_Y>
__global__ void __launch_bounds__(RED_TB)
k_axpby(float2* out, float2 a, const float2* x, float2 b, const float2* y, long long n, long long n4) {
    float4* o4 = reinterpret_cast<float4*>(out);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
#define AX_PIECES(U)                                                                               \
    float4 xv[U], yv[U];                                                                           \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        xv[u] = x4[i + u * stride];                                                                \
        if (HAS_Y) yv[u] = y4[i + u * stride];                                                     \
    }                                                                                              \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                                \
        float2 r0 = cmul(a, lo2(xv[u])), r1 = cmul(a, hi2(xv[u]));                                 \
        if (HAS_Y) { r0 = add2(r0, cmul(b, lo2(yv[u]))); r1 = add2(r1, cmul(b, hi2(yv[u]))); }     \
        o4[i + u * stride] = pack4(r0, r1);                                                        \
    }
#define AX_SINGLE                                                                                  \
    float2 r0 = cmul(a, x[j]);                                                                     \
    if (HAS_Y) r0 = add2(r0, cmul(b, y[j]));                                                       \
    out[j] = r0;
    STREAM_LOOP(n4, AX_PIECES, AX_SINGLE)
#undef AX_PIECES
#undef AX_SINGLE
}

// ---- L1TVOLS ---------------------------------------------------------------------------------
// periodic neighbour along axis d of the linear image index n
__device__ __forceinline__ long long shifted(const Geom& g, long long n, int d, int step, const int* coord,
                                             const long long* nstride) {
    int i = coord[d] + step;
    if (i < 0) i += g.N[d];
    if (i >= g.N[d]) i -= g.N[d];
    return n + (long long)(i - coord[d]) * nstride[d];
}

__device__ __forceinline__ void decode_image(const Geom& g, long long n, int* coord, long long* nstride) {
    long long s = g.Nprod;
    for (int d = 0; d < g.ndim; ++d) {
        s /= g.N[d];
        nstride[d] = s;
        coord[d] = (int)(n / s);
        n -= coord[d] * s;
    }
}

// rhs = mu*AHyk + lambda * sum_p Dt_p(d_p - b_p),  Dt_p(v)[i] = v[i + e_p] - v[i]
// (solve_device.py:133-154; dt_indx = roll(-1), helper.py:67; cDiff re_subroutine.py:936-950)
__global__ void k_tv_rhs(Geom g, const float2* __restrict__ AHyk, const float2* __restrict__ dd,
                         const float2* __restrict__ bb, float mu, float lambda, float2* __restrict__ rhs) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= g.Nprod) return;
    int coord[MAXD];
    long long ns[MAXD];
    decode_image(g, n, coord, ns);
    float2 a = AHyk[n];
    float2 acc = make_float2(mu * a.x, mu * a.y);
    for (int p = 0; p < g.ndim; ++p) {
        const float2* dp = dd + (long long)p * g.Nprod;
        const float2* bp = bb + (long long)p * g.Nprod;
        long long nn = shifted(g, n, p, +1, coord, ns);
        float2 v1 = make_float2(dp[nn].x - bp[nn].x, dp[nn].y - bp[nn].y);
        float2 v0 = make_float2(dp[n].x - bp[n].x, dp[n].y - bp[n].y);
        float2 t = make_float2((v1.x - v0.x) * lambda, (v1.y - v0.y) * lambda);
        acc.x += t.x;
        acc.y += t.y;
    }
    rhs[n] = acc;
}

// z_p = D_p(x) (D_p(v)[i] = v[i - e_p] - v[i]; d_indx = roll(+1), helper.py:66); s_p = z_p + b_p;
// s = s_0, then hypot(|s|,|s_p|) (cHypot :656-678); s += 1e-6; t = shrink(s, 1/lambda)/s
// (cAnisoShrink :973-993); d_p = s_p*t; b_p += z_p - d_p          (solve_device.py:189-262)
__global__ void k_tv_shrink(Geom g, const float2* __restrict__ x, float2* __restrict__ dd,
                            float2* __restrict__ bb, float lambda) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= g.Nprod) return;
    int coord[MAXD];
    long long ns[MAXD];
    decode_image(g, n, coord, ns);
    const float2 x0 = x[n];
    float2 z[MAXD], sp[MAXD];
    float2 s = make_float2(0.f, 0.f);
    for (int p = 0; p < g.ndim; ++p) {
        float2 xm = x[shifted(g, n, p, -1, coord, ns)];
        z[p] = make_float2(xm.x - x0.x, xm.y - x0.y);
        float2 b = bb[(long long)p * g.Nprod + n];
        sp[p] = make_float2(z[p].x + b.x, z[p].y + b.y);
        if (p == 0) {
            s = sp[0];
        } else {
            float hx = hypotf(s.x, s.y), hy = hypotf(sp[p].x, sp[p].y);
            s = make_float2(hypotf(hx, hy), 0.f);
        }
    }
    s.x += 1e-6f;
    const float thr = 1.0f / lambda;
    float2 t;
    t.x = (s.x > thr) * (s.x - thr) + (s.x < -thr) * (s.x + thr);
    t.y = (s.y > thr) * (s.y - thr) + (s.y < -thr) * (s.y + thr);
    {   // t /= s (complex)
        float m = s.x * s.x + s.y * s.y;
        t = make_float2((t.x * s.x + t.y * s.y) / m, (t.y * s.x - t.x * s.y) / m);
    }
    for (int p = 0; p < g.ndim; ++p) {
        float2 d = cmul(sp[p], t);
        long long o = (long long)p * g.Nprod + n;
        float2 b = bb[o];
        dd[o] = d;
        bb[o] = make_float2(b.x + (z[p].x - d.x), b.y + (z[p].y - d.y));
    }
}

__global__ void k_tv_bregman(float2* __restrict__ AHyk, const float2* __restrict__ zf,
                             const float2* __restrict__ AHy, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float2 a = AHyk[i], z = zf[i], h = AHy[i];
        AHyk[i] = make_float2(a.x - (z.x - h.x), a.y - (z.y - h.y));
    }
}

// ---- C ABI -------------------------------------------------------------------------------------
static inline unsigned stream_blocks(long long n) {
    long long b = (n + RED_TB - 1) / RED_TB;
    return (unsigned)(b < RED_BLOCKS ? (b > 0 ? b : 1) : RED_BLOCKS);
}

extern "C" int b200nufft_zero_scalars(double* s, int n, void* stream) {
    ARG_CHECK(s && n >= 0, "zero_scalars: bad arguments");
    ON_DEVICE(device_of(s));
    CUDA_TRY(cudaMemsetAsync(s, 0, sizeof(double) * n, as_stream(stream)));
    return B200_OK;
}

extern "C" int b200nufft_dotc(const b200_c64* a, const b200_c64* b, int64_t n, double* out, void* stream) {
    ARG_CHECK(a && b && out && n >= 0, "dotc: bad arguments");
    ON_DEVICE(device_of(a));
    const long long n4 = aligned16(a) && aligned16(b) ? n / 2 : 0;
    k_dotc<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(a),
                                                               reinterpret_cast<const float2*>(b), n, n4, out);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_cg_init(const b200_c64* b, const b200_c64* Ax, b200_c64* r, b200_c64* p, double* rsold,
                                 int64_t n, void* stream) {
    ARG_CHECK(b && Ax && r && p && rsold && n >= 0, "cg_init: bad arguments");
    ON_DEVICE(device_of(b));
    k_cg_init<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(b), reinterpret_cast<const float2*>(Ax), reinterpret_cast<float2*>(r),
        reinterpret_cast<float2*>(p), rsold, n, aligned16(b) && aligned16(Ax) && aligned16(r) && aligned16(p) ? n / 2 : 0);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_cg_update_xr(b200_c64* x, b200_c64* r, const b200_c64* p, const b200_c64* Ap,
                                      const double* rsold, const double* pAp, double* rsnew, int64_t n,
                                      void* stream) {
    ARG_CHECK(x && r && p && Ap && rsold && pAp && rsnew && n >= 0, "cg_update_xr: bad arguments");
    ON_DEVICE(device_of(x));
    k_cg_update_xr<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(
        reinterpret_cast<float2*>(x), reinterpret_cast<float2*>(r), reinterpret_cast<const float2*>(p),
        reinterpret_cast<const float2*>(Ap), rsold, pAp, rsnew, n,
        aligned16(x) && aligned16(r) && aligned16(p) && aligned16(Ap) ? n / 2 : 0);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_cg_update_p(b200_c64* p, const b200_c64* r, const double* rsnew, const double* rsold,
                                     int64_t n, void* stream) {
    ARG_CHECK(p && r && rsnew && rsold && n >= 0, "cg_update_p: bad arguments");
    ON_DEVICE(device_of(p));
    k_cg_update_p<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(
        reinterpret_cast<float2*>(p), reinterpret_cast<const float2*>(r), rsnew, rsold, n,
        aligned16(p) && aligned16(r) ? n / 2 : 0);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_cdiv(b200_c64* a, const b200_c64* b, int64_t n, void* stream) {
    ARG_CHECK(a && b && n >= 0, "cdiv: bad arguments");
    ON_DEVICE(device_of(a));
    k_cdiv<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(a),
                                                               reinterpret_cast<const float2*>(b), n);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_cmul(b200_c64* a, const b200_c64* b, int64_t n, void* stream) {
    ARG_CHECK(a && b && n >= 0, "cmul: bad arguments");
    ON_DEVICE(device_of(a));
    k_cmul<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(a),
                                                               reinterpret_cast<const float2*>(b), n);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_axpby(b200_c64* out, double a_re, double a_im, const b200_c64* x, double b_re, double b_im,
                               const b200_c64* y, int64_t n, void* stream) {
    ARG_CHECK(out && x && n >= 0, "axpby: bad arguments");
    ON_DEVICE(device_of(out));
    const float2 a = make_float2((float)a_re, (float)a_im), b = make_float2((float)b_re, (float)b_im);
    float2* o = reinterpret_cast<float2*>(out);
    const float2* xv = reinterpret_cast<const float2*>(x);
    const bool has_y = y && (b_re != 0.0 || b_im != 0.0);
    const long long n4 = aligned16(out) && aligned16(x) && (!has_y || aligned16(y)) ? n / 2 : 0;
    if (has_y)
        k_axpby<true><<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(o, a, xv, b, reinterpret_cast<const float2*>(y), n, n4);
    else
        k_axpby<false><<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(o, a, xv, b, nullptr, n, n4);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_tv_rhs(b200nufft_plan_t p, const b200_c64* AHyk, const b200_c64* d, const b200_c64* b,
                                float mu, float lambda, b200_c64* rhs, void* stream) {
    ARG_CHECK(p && AHyk && d && b && rhs, "tv_rhs: bad arguments");
    ON_DEVICE(p->device);
    const int TB = 256;
    k_tv_rhs<<<(unsigned)((p->g.Nprod + TB - 1) / TB), TB, 0, as_stream(stream)>>>(
        p->g, reinterpret_cast<const float2*>(AHyk), reinterpret_cast<const float2*>(d),
        reinterpret_cast<const float2*>(b), mu, lambda, reinterpret_cast<float2*>(rhs));
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_tv_shrink(b200nufft_plan_t p, const b200_c64* x, b200_c64* d, b200_c64* b, float lambda,
                                   void* stream) {
    ARG_CHECK(p && x && d && b && lambda != 0.f, "tv_shrink: bad arguments");
    ON_DEVICE(p->device);
    const int TB = 256;
    k_tv_shrink<<<(unsigned)((p->g.Nprod + TB - 1) / TB), TB, 0, as_stream(stream)>>>(
        p->g, reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(d), reinterpret_cast<float2*>(b), lambda);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_tv_bregman(b200_c64* AHyk, const b200_c64* zf, const b200_c64* AHy, int64_t n,
                                    void* stream) {
    ARG_CHECK(AHyk && zf && AHy && n >= 0, "tv_bregman: bad arguments");
    ON_DEVICE(device_of(AHyk));
    k_tv_bregman<<<stream_blocks(n), RED_TB, 0, as_stream(stream)>>>(
        reinterpret_cast<float2*>(AHyk), reinterpret_cast<const float2*>(zf), reinterpret_cast<const float2*>(AHy), n);
    LAUNCH_CHECK();
    return B200_OK;
}
