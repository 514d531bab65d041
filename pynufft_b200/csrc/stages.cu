// Scaling / zero-padding / cropping kernels, cuFFT wrapper, operator compositions.
// Replaces cTensorMultiply, cTensorCopy, cPopulate, cMultiplyVecInplace,
// cMultiplyConjVecInplace, cAggregate (src/re_subroutine.py:98-201, 441-514, 839-921) and the
// reikna FFT (nufft/_nufft_class_methods_device.py:246-249).
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// x2xx: out = in * sn or in / sn   (image layout, batch innermost)
// ------------------------------------------------------------------------------------------
__global__ void k_x2xx(Geom g, const float* __restrict__ sn, const float2* __restrict__ in,
                       float2* __restrict__ out, int nb, int div) {
    long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= g.Nprod * nb) return;
    long long n = gid / nb;
    // same factor order as cTensorMultiply (re_subroutine.py:169-178): dim 0 first
    float s = 1.f;
    {
        long long res = n;
        long long stride = g.Nprod;
        for (int d = 0; d < g.ndim; ++d) {
            stride /= g.N[d];
            int i = (int)(res / stride);
            res -= i * stride;
            s = s * sn[g.snoff[d] + i];
        }
    }
    float2 v = in[gid];
    if (div) { v.x = v.x / s; v.y = v.y / s; } else { v.x *= s; v.y *= s; }
    out[gid] = v;
}

// ------------------------------------------------------------------------------------------
// scale_pad: one (threadIdx.y) thread row per grid row (all leading coordinates fixed); the
// leading-index decode is uniform per row and 32-bit; every grid element is written (zeros
// outside the image corner), VEC = 2 -> float4 stores.  blockIdx.y = coil.
// ------------------------------------------------------------------------------------------
#define SP_TX 64
#define SP_TY 4
template <int VEC>
__global__ void __launch_bounds__(SP_TX * SP_TY)
k_scale_pad(Geom g, const float* __restrict__ sn, const float2* __restrict__ x, float2* __restrict__ grid,
            int nb, int apply_sn, int x_single, const float2* __restrict__ sens, int nrows) {
    const int c = blockIdx.y;
    const int row = blockIdx.x * SP_TY + threadIdx.y;
    if (row >= nrows) return;
    const int dl = g.ndim - 1;
    const int KL = g.K[dl], NL = g.N[dl];
    // decode the leading coordinates of this row
    int rem = row, n = 0;
    bool inside = true;
    float sbase = 1.f;
    for (int d = dl - 1; d >= 0; --d) {
        const int i = rem % g.K[d];
        rem /= g.K[d];
        if (i >= g.N[d]) inside = false;
    }
    if (inside) {
        rem = row;
        int nstride = 1;
        for (int d = dl - 1; d >= 0; --d) {
            const int i = rem % g.K[d];
            rem /= g.K[d];
            n += i * nstride;
            nstride *= g.N[d];
            if (apply_sn) sbase *= sn[g.snoff[d] + i];
        }
    }
    float2* dst = grid + (long long)c * g.Kprod + (long long)row * KL;
    const long long nrow = (long long)n * NL;
    for (int i2 = threadIdx.x * VEC; i2 < KL; i2 += SP_TX * VEC) {
        float2 v[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            v[q] = make_float2(0.f, 0.f);
            if (inside && i2 + q < NL) {
                const long long nn = nrow + i2 + q;
                float2 xv = x_single ? x[nn] : x[nn * nb + c];
                if (sens) xv = cmul(xv, sens[nn * nb + c]);
                const float f = apply_sn ? sbase * sn[g.snoff[dl] + i2 + q] : 1.f;
                v[q] = make_float2(xv.x * f, xv.y * f);
            }
        }
        if (VEC == 2) {
            *reinterpret_cast<float4*>(dst + i2) = make_float4(v[0].x, v[0].y, v[VEC - 1].x, v[VEC - 1].y);
        } else {
            dst[i2] = v[0];
        }
    }
}

// ------------------------------------------------------------------------------------------
// crop_scale
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int image_to_grid(const Geom& g, int n, float* s_out, const float* __restrict__ sn) {
    // 32-bit decode (prod(Nd) <= prod(Kd) < 2^31), last dimension first
    int idx = 0;
    float s = 1.f;
    for (int d = g.ndim - 1; d >= 0; --d) {
        const int i = n % g.N[d];
        n /= g.N[d];
        idx += i * (int)g.Kstride[d];
        s *= sn[g.snoff[d] + i];
    }
    *s_out = s;
    return idx;
}

__global__ void k_crop_scale(Geom g, const float* __restrict__ sn, const float2* __restrict__ grid,
                             float2* __restrict__ x, int nb, int mode, float scale) {
    long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (gid >= g.Nprod * nb) return;
    const int n = (int)(gid / nb);
    const int c = (int)(gid - (long long)n * nb);
    float s;
    const int idx = image_to_grid(g, n, &s, sn);
    float f = scale * (mode == 1 ? s : (mode == 2 ? 1.f / s : 1.f));
    float2 v = grid[(long long)c * g.Kprod + idx];
    x[gid] = make_float2(v.x * f, v.y * f);
}

__global__ void k_crop_combine(Geom g, const float* __restrict__ sn, const float2* __restrict__ grid,
                               float2* __restrict__ x, int nb, int mode, float scale,
                               const float2* __restrict__ sens) {
    const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= g.Nprod) return;
    float s;
    const int idx = image_to_grid(g, (int)n, &s, sn);
    float f = scale * (mode == 1 ? s : (mode == 2 ? 1.f / s : 1.f)) / (float)nb;   // mean over coils
    float2 acc = make_float2(0.f, 0.f);
    for (int c = 0; c < nb; ++c) {
        float2 v = grid[(long long)c * g.Kprod + idx];
        if (sens) v = cmulc(sens[n * nb + c], v);
        acc.x += v.x;
        acc.y += v.y;
    }
    x[n] = make_float2(acc.x * f, acc.y * f);
}

__global__ void k_scale_inplace(float2* __restrict__ a, long long n, float f) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float2 v = a[i];
        a[i] = make_float2(v.x * f, v.y * f);
    }
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int b200nufft_x2xx(b200nufft_plan_t p, const b200_c64* in, b200_c64* out, int nb, int div,
                              void* stream) {
    ARG_CHECK(p && in && out && nb >= 1, "x2xx: bad arguments");
    ON_DEVICE(p->device);
    long long tot = p->g.Nprod * nb;
    const int TB = 256;
    k_x2xx<<<(unsigned)((tot + TB - 1) / TB), TB, 0, as_stream(stream)>>>(
        p->g, p->d_sn, reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), nb, div);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_scale_pad(b200nufft_plan_t p, const b200_c64* x, b200_c64* grid, int nb,
                                   int apply_sn, int x_single, const b200_c64* sens, void* stream) {
    ARG_CHECK(p && x && grid && nb >= 1 && nb <= 65535, "scale_pad: bad arguments");
    ON_DEVICE(p->device);
    if (use_bi(p, nb))
        return sweep2d_scale_pad(p, reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(grid), nb, apply_sn,
                                 x_single, reinterpret_cast<const float2*>(sens), as_stream(stream));
    const Geom& g = p->g;
    const int dl = g.ndim - 1;
    const int nrows = (int)(g.Kprod / g.K[dl]);
    bool vec2 = (g.K[dl] % 2 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    dim3 blk(SP_TX, SP_TY);
    dim3 gr((unsigned)((nrows + SP_TY - 1) / SP_TY), nb);
    if (vec2) {
        k_scale_pad<2><<<gr, blk, 0, as_stream(stream)>>>(g, p->d_sn, reinterpret_cast<const float2*>(x),
                                                          reinterpret_cast<float2*>(grid), nb, apply_sn, x_single,
                                                          reinterpret_cast<const float2*>(sens), nrows);
    } else {
        k_scale_pad<1><<<gr, blk, 0, as_stream(stream)>>>(g, p->d_sn, reinterpret_cast<const float2*>(x),
                                                          reinterpret_cast<float2*>(grid), nb, apply_sn, x_single,
                                                          reinterpret_cast<const float2*>(sens), nrows);
    }
    LAUNCH_CHECK();
    return B200_OK;
}

static int get_fft(b200nufft_plan_t p, int nb, cufftHandle* out) {
    if (p->fft_valid && p->fft_nb == nb) { *out = p->fft; return B200_OK; }
    if (p->fft_valid) { cufftDestroy(p->fft); p->fft_valid = false; }
    int n[MAXD];
    for (int d = 0; d < p->g.ndim; ++d) n[d] = p->g.K[d];
    CUFFT_TRY(cufftPlanMany(&p->fft, p->g.ndim, n, nullptr, 1, (int)p->g.Kprod, nullptr, 1, (int)p->g.Kprod,
                            CUFFT_C2C, nb));
    p->fft_valid = true;
    p->fft_nb = nb;
    *out = p->fft;
    return B200_OK;
}

// Pruned 3-D transforms (two cuFFT plans): the padded image is non-zero only in planes i0 < N0, and the
// adjoint only needs planes i0 < N0 of the inverse transform.  forward: batched 2-D FFT over (dim 1, dim 2)
// on the first N0 planes, then the strided 1-D FFT along dim 0 on everything; inverse: the same two in the
// opposite order.  That is (1/2 + 1/2 + 1)/3 of the memory passes of the full 3-D plan.
static int get_pruned_fft(b200nufft_plan_t p) {
    if (p->fftp_valid) return B200_OK;
    const Geom& g = p->g;
    int n2[2] = {g.K[1], g.K[2]};
    CUFFT_TRY(cufftPlanMany(&p->fft2d, 2, n2, nullptr, 1, g.K[1] * g.K[2], nullptr, 1, g.K[1] * g.K[2], CUFFT_C2C,
                            g.N[0]));
    int n1[1] = {g.K[0]};
    int emb[1] = {g.K[0]};
    CUFFT_TRY(cufftPlanMany(&p->fft1d, 1, n1, emb, g.K[1] * g.K[2], 1, emb, g.K[1] * g.K[2], 1, CUFFT_C2C,
                            g.K[1] * g.K[2]));
    p->fftp_valid = true;
    return B200_OK;
}

static bool can_prune(const Geom& g) { return g.ndim == 3 && 2 * g.N[0] <= g.K[0] + g.K[0] / 2; }

// mode 3: forward, input zero outside planes i0 < N0; mode 4: inverse unnormalised, only planes i0 < N0 valid
static int fft_pruned(b200nufft_plan_t p, float2* grid, int nb, int mode, cudaStream_t st) {
    int rc = get_pruned_fft(p);
    if (rc) return rc;
    CUFFT_TRY(cufftSetStream(p->fft2d, st));
    CUFFT_TRY(cufftSetStream(p->fft1d, st));
    for (int c = 0; c < nb; ++c) {
        cufftComplex* gc = reinterpret_cast<cufftComplex*>(grid + (long long)c * p->g.Kprod);
        if (mode == 3) {
            CUFFT_TRY(cufftExecC2C(p->fft2d, gc, gc, CUFFT_FORWARD));
            CUFFT_TRY(cufftExecC2C(p->fft1d, gc, gc, CUFFT_FORWARD));
        } else {
            CUFFT_TRY(cufftExecC2C(p->fft1d, gc, gc, CUFFT_INVERSE));
            CUFFT_TRY(cufftExecC2C(p->fft2d, gc, gc, CUFFT_INVERSE));
        }
    }
    return B200_OK;
}

// inverse: 0 forward, 1 inverse normalised by 1/prod(Kd), 2 inverse unnormalised,
//          3 forward of a zero-padded image (pruned), 4 inverse unnormalised, image corner planes only (pruned)
extern "C" int b200nufft_fft(b200nufft_plan_t p, b200_c64* grid, int nb, int inverse, void* stream) {
    ARG_CHECK(p && grid && nb >= 1 && inverse >= 0 && inverse <= 4, "fft: bad arguments");
    ON_DEVICE(p->device);
    if (use_bi(p, nb)) {
        int rc = sweep2d_fft(p, reinterpret_cast<float2*>(grid), nb, (inverse == 0 || inverse == 3) ? 0 : 1, as_stream(stream));
        if (rc || inverse != 1) return rc;
        k_scale_inplace<<<148 * 8, 256, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(grid), p->g.Kprod * nb,
                                                               1.0f / (float)p->g.Kprod);
        LAUNCH_CHECK();
        return B200_OK;
    }
    if (inverse >= 3) {
        if (can_prune(p->g)) return fft_pruned(p, reinterpret_cast<float2*>(grid), nb, inverse, as_stream(stream));
        inverse = inverse == 3 ? 0 : 2;
    }
    cufftHandle h;
    int rc = get_fft(p, nb, &h);
    if (rc) return rc;
    CUFFT_TRY(cufftSetStream(h, as_stream(stream)));
    CUFFT_TRY(cufftExecC2C(h, reinterpret_cast<cufftComplex*>(grid), reinterpret_cast<cufftComplex*>(grid),
                           inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
    if (inverse == 1) {
        long long tot = p->g.Kprod * nb;
        k_scale_inplace<<<148 * 8, 256, 0, as_stream(stream)>>>(reinterpret_cast<float2*>(grid), tot,
                                                               1.0f / (float)p->g.Kprod);
        LAUNCH_CHECK();
    }
    return B200_OK;
}

static int crop_scale_impl(b200nufft_plan_t p, const float2* grid, float2* x, int nb, int mode, int combine,
                           const float2* sens, float scale, cudaStream_t st, bool coil_major = false) {
    if (!coil_major && use_bi(p, nb)) return sweep2d_crop_scale(p, grid, x, nb, mode, combine, sens, scale, st);
    const int TB = 256;
    if (combine) {
        k_crop_combine<<<(unsigned)((p->g.Nprod + TB - 1) / TB), TB, 0, st>>>(p->g, p->d_sn, grid, x, nb, mode,
                                                                             scale, sens);
    } else {
        long long tot = p->g.Nprod * nb;
        k_crop_scale<<<(unsigned)((tot + TB - 1) / TB), TB, 0, st>>>(p->g, p->d_sn, grid, x, nb, mode, scale);
    }
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_crop_scale(b200nufft_plan_t p, const b200_c64* grid, b200_c64* x, int nb, int mode,
                                    int combine, const b200_c64* sens, void* stream) {
    ARG_CHECK(p && grid && x && nb >= 1 && mode >= 0 && mode <= 2, "crop_scale: bad arguments");
    ON_DEVICE(p->device);
    return crop_scale_impl(p, reinterpret_cast<const float2*>(grid), reinterpret_cast<float2*>(x), nb, mode,
                           combine, reinterpret_cast<const float2*>(sens), 1.0f, as_stream(stream));
}

int ensure_scratch(b200nufft_plan_t p, int nb) {
    if (p->grid_nb >= nb) return B200_OK;
    if (p->d_grid) { CUDA_TRY(cudaFree(p->d_grid)); p->d_grid = nullptr; p->grid_nb = 0; }
    CUDA_TRY(cudaMalloc(&p->d_grid, sizeof(float2) * p->g.Kprod * nb));
    p->grid_nb = nb;
    return B200_OK;
}

int ensure_scratch2(b200nufft_plan_t p, int nb) {
    if (p->grid2_nb >= nb) return B200_OK;
    if (p->d_grid2) { CUDA_TRY(cudaFree(p->d_grid2)); p->d_grid2 = nullptr; p->grid2_nb = 0; }
    CUDA_TRY(cudaMalloc(&p->d_grid2, sizeof(float2) * p->g.Kprod * nb));
    p->grid2_nb = nb;
    return B200_OK;
}

// pad_fft: grid = FFT(zero-pad(x * [sn] * [sens])).  Fused pruned passes when the geometry allows
// (fft256.cu), else scale_pad + cuFFT (pruned plan for other 3-D sizes).
// `modulated`: leave the grid phase-modulated, G'[g] = G[g] prod_d m_d[g_d] (what the column-sweep gather reads)
static int pad_fft_impl(b200nufft_plan_t p, const b200_c64* x, b200_c64* grid, int nb, int apply_sn, int x_single,
                        const b200_c64* sens, void* stream, bool modulated) {
    ARG_CHECK(p && x && grid && nb >= 1 && nb <= 65535, "pad_fft: bad arguments");
    ON_DEVICE(p->device);
    if (modulated && !p->d_mod) {
        b200_set_error("pad_fft: modulated grid requested but the plan has no modulation tables");
        return B200_ERR_UNSUPPORTED;
    }
    if (p->fft_variant != 1 && fft256_supported(p->g))
        return fft256_forward(p, reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(grid), nb, apply_sn,
                              x_single, reinterpret_cast<const float2*>(sens), modulated, as_stream(stream));
    if (use_bi(p, nb) && p->fft_variant != 1 && fftbi_supported(p->g))     // fused pruned passes on the layout itself
        return fftbi_forward(p, reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(grid), nb, apply_sn,
                             x_single, reinterpret_cast<const float2*>(sens), as_stream(stream),
                             modulated && bi_fused_mod(p, nb));
    if (use_bi(p, nb))              // batch-innermost grid out: coil-major pad + cuFFT on a scratch, one transposing pass
        return sweep2d_pad_fft(p, reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(grid), nb, apply_sn,
                               x_single, reinterpret_cast<const float2*>(sens), as_stream(stream));
    int rc = b200nufft_scale_pad(p, x, grid, nb, apply_sn, x_single, sens, stream);
    if (rc) return rc;
    rc = b200nufft_fft(p, grid, nb, 3, stream);
    if (rc || !modulated) return rc;
    return col3d_modulate(p, reinterpret_cast<const float2*>(grid), reinterpret_cast<float2*>(grid), nb, as_stream(stream));
}

extern "C" int b200nufft_pad_fft(b200nufft_plan_t p, const b200_c64* x, b200_c64* grid, int nb, int apply_sn,
                                 int x_single, const b200_c64* sens, void* stream) {
    return pad_fft_impl(p, x, grid, nb, apply_sn, x_single, sens, stream, false);
}
// the same, leaving the grid in the form b200nufft_interp_modulated reads (only when b200nufft_kspace_modulated(plan))
extern "C" int b200nufft_pad_fft_modulated(b200nufft_plan_t p, const b200_c64* x, b200_c64* grid, int nb, int apply_sn,
                                           int x_single, const b200_c64* sens, void* stream) {
    return pad_fft_impl(p, x, grid, nb, apply_sn, x_single, sens, stream, true);
}

// ifft_crop: x = crop(IFFT(grid)) * f (mode 0/1/2 as crop_scale), optionally combined over coils.
// The grid is used as scratch: its contents are undefined afterwards.
// `modulated`: the grid comes from gridding_impl(..., modulated_ok = true) (phase-modulated iff
// gridding_modulated(p)); the fused inverse passes undo the modulation on the way in.
static int ifft_crop_impl(b200nufft_plan_t p, b200_c64* grid, b200_c64* x, int nb, int mode, int combine,
                          const b200_c64* sens, void* stream, bool modulated) {
    ARG_CHECK(p && x && grid && nb >= 1 && mode >= 0 && mode <= 2, "ifft_crop: bad arguments");
    ON_DEVICE(p->device);
    cudaStream_t st = as_stream(stream);
    const float scale = 1.0f / (float)p->g.Kprod;
    const bool mod = modulated && gridding_modulated(p);
    const bool bi_mod = modulated && p->M > 0 && bi_fused_mod(p, nb);    // 2-D multi-coil: as sweep2d_gridding(..., modulated) left it
    if (p->fft_variant != 1 && fft256_supported(p->g)) {
        if (!combine)
            return fft256_inverse(p, reinterpret_cast<float2*>(grid), reinterpret_cast<float2*>(x), nb, mode, scale,
                                  mod, st);
        if (p->xc_nb < nb) {
            if (p->d_xc) { CUDA_TRY(cudaFree(p->d_xc)); p->d_xc = nullptr; p->xc_nb = 0; }
            CUDA_TRY(cudaMalloc(&p->d_xc, sizeof(float2) * p->g.Nprod * nb));
            p->xc_nb = nb;
        }
        int rc = fft256_inverse(p, reinterpret_cast<float2*>(grid), p->d_xc, nb, mode, scale, mod, st);
        if (rc) return rc;
        return combine_coils(p->d_xc, reinterpret_cast<const float2*>(sens), reinterpret_cast<float2*>(x),
                             p->g.Nprod, nb, st);
    }
    int rc = B200_OK;
    if (use_bi(p, nb) && p->fft_variant != 1 && fftbi_supported(p->g)) {
        if (!combine)
            return fftbi_inverse(p, reinterpret_cast<float2*>(grid), reinterpret_cast<float2*>(x), nb, mode, scale, st, bi_mod);
        if (p->xc_nb < nb) {
            if (p->d_xc) { CUDA_TRY(cudaFree(p->d_xc)); p->d_xc = nullptr; p->xc_nb = 0; }
            CUDA_TRY(cudaMalloc(&p->d_xc, sizeof(float2) * p->g.Nprod * nb));
            p->xc_nb = nb;
        }
        rc = fftbi_inverse(p, reinterpret_cast<float2*>(grid), p->d_xc, nb, mode, scale, st, bi_mod);
        if (rc) return rc;
        return combine_coils(p->d_xc, reinterpret_cast<const float2*>(sens), reinterpret_cast<float2*>(x), p->g.Nprod, nb,
                             st);
    }
    if (use_bi(p, nb)) {            // batch-innermost grid in: one transposing pass, cuFFT and crop on the coil-major scratch
        rc = sweep2d_ifft_to_scratch(p, reinterpret_cast<const float2*>(grid), nb, st);
        if (rc) return rc;
        return crop_scale_impl(p, p->d_grid2, reinterpret_cast<float2*>(x), nb, mode, combine,
                               reinterpret_cast<const float2*>(sens), scale, st, true);
    }
    if (mod) {
        rc = col3d_demodulate(p, reinterpret_cast<float2*>(grid), nb, st);
        if (rc) return rc;
    }
    rc = b200nufft_fft(p, grid, nb, 4, stream);
    if (rc) return rc;
    // the 1/prod(Kd) of the inverse FFT is folded into the crop kernel
    return crop_scale_impl(p, reinterpret_cast<const float2*>(grid), reinterpret_cast<float2*>(x), nb, mode, combine,
                           reinterpret_cast<const float2*>(sens), scale, st);
}

extern "C" int b200nufft_ifft_crop(b200nufft_plan_t p, b200_c64* grid, b200_c64* x, int nb, int mode, int combine,
                                   const b200_c64* sens, void* stream) {
    return ifft_crop_impl(p, grid, x, nb, mode, combine, sens, stream, false);
}
extern "C" int b200nufft_ifft_crop_modulated(b200nufft_plan_t p, b200_c64* grid, b200_c64* x, int nb, int mode,
                                             int combine, const b200_c64* sens, void* stream) {
    return ifft_crop_impl(p, grid, x, nb, mode, combine, sens, stream, true);
}
extern "C" int b200nufft_gridding_is_modulated(b200nufft_plan_t p) { return (p && gridding_modulated(p)) ? 1 : 0; }

static int forward_impl(b200nufft_plan_t p, const float2* x, int x_single, const float2* sens, float2* y, int nb,
                        void* stream) {
    int rc = ensure_scratch(p, nb);
    if (rc) return rc;
    // the column-sweep gather reads the phase-modulated grid: the forward FFT passes produce it directly
    // (3-D); the 2-D multi-coil FFT pass along dim 0 writes the modulated grid the row-sweep gather reads
    // ("auto": only where the fused passes emit the modulated grid for free)
    const bool col_auto = p->interp_variant == 0 && interp_col_on_modulated(p) && p->fft_variant != 1 && fft256_supported(p->g);
    const bool mod = ((interp_uses_col(p) || col_auto) && !use_bi(p, nb)) || bi_fused_mod(p, nb);
    rc = pad_fft_impl(p, reinterpret_cast<const b200_c64*>(x), reinterpret_cast<b200_c64*>(p->d_grid), nb, 1,
                      x_single, reinterpret_cast<const b200_c64*>(sens), stream, mod);
    if (rc) return rc;
    return interp_impl(p, p->d_grid, y, nb, as_stream(stream), mod);
}

// The column-sweep adjoint works on its own grid, which the plan keeps zeroed between calls: the zero-fill (20 us at
// 256^3) runs on a side stream behind the inverse FFT passes of the previous call, so that it overlaps whatever the
// caller's stream does next instead of sitting in front of the scatter.  Measured: 14 us per pair while the pre-pass
// was one sample per thread; 3 us since its loads are batched (the zero-fill stores now hide behind the data gather
// anyway), for one more grid per coil.  Off by default; B200NUFFT_PREZERO=1 turns it on.
static int ensure_gridz(b200nufft_plan_t p, int nb) {
    if (!p->zs) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p->zs, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&p->gridz_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&p->gridz_use, cudaEventDisableTiming));
    }
    if (p->gridz_nb >= nb) return B200_OK;
    if (p->d_gridz) {
        CUDA_TRY(cudaStreamSynchronize(p->zs));
        CUDA_TRY(cudaFree(p->d_gridz));
        p->d_gridz = nullptr;
        p->gridz_nb = 0;
    }
    p->gridz_clean = false;
    CUDA_TRY(cudaMalloc(&p->d_gridz, sizeof(float2) * p->g.Kprod * nb));
    p->gridz_nb = nb;
    return B200_OK;
}

static int adjoint_impl(b200nufft_plan_t p, const float2* y, float2* x, int nb, int combine, const float2* sens,
                        void* stream) {
    static const bool prezero = [] { const char* e = getenv("B200NUFFT_PREZERO"); return e && atoi(e) != 0; }();
    const size_t gbytes = sizeof(float2) * (size_t)p->g.Kprod * nb;
    if (prezero && p->M > 0 && gridding_modulated(p) && gbytes <= ((size_t)4 << 30)) {
        int rc = ensure_gridz(p, nb);
        if (rc) return rc;
        cudaStream_t st = as_stream(stream);
        const bool clean = p->gridz_clean && p->gridz_nb == nb;
        if (clean) CUDA_TRY(cudaStreamWaitEvent(st, p->gridz_ev, 0));
        p->gridz_clean = false;
        rc = col3d_gridding(p, y, p->d_gridz, nb, st, clean);
        if (rc) return rc;
        rc = ifft_crop_impl(p, reinterpret_cast<b200_c64*>(p->d_gridz), reinterpret_cast<b200_c64*>(x), nb, 1, combine,
                            reinterpret_cast<const b200_c64*>(sens), stream, true);
        if (rc) return rc;
        // zero the grid for the next call, behind this call's last reader
        CUDA_TRY(cudaEventRecord(p->gridz_use, st));
        CUDA_TRY(cudaStreamWaitEvent(p->zs, p->gridz_use, 0));
        CUDA_TRY(cudaMemsetAsync(p->d_gridz, 0, gbytes, p->zs));
        CUDA_TRY(cudaEventRecord(p->gridz_ev, p->zs));
        p->gridz_clean = true;
        return B200_OK;
    }
    int rc = ensure_scratch(p, nb);
    if (rc) return rc;
    // 3-D column sweep in front of cuFFT (no fused passes at this size): ask the scatter for the true grid -- its
    // demodulating instantiation -- instead of a demodulation pass over the grid in front of the inverse FFT
    const bool col3 = p->M > 0 && gridding_modulated(p) && !use_bi(p, nb);
    const bool keep_mod = !col3 || (p->fft_variant != 1 && fft256_supported(p->g));
    rc = gridding_impl(p, y, p->d_grid, nb, as_stream(stream), keep_mod);
    if (rc) return rc;
    return ifft_crop_impl(p, reinterpret_cast<b200_c64*>(p->d_grid), reinterpret_cast<b200_c64*>(x), nb, 1,
                          combine, reinterpret_cast<const b200_c64*>(sens), stream, keep_mod);
}

extern "C" int b200nufft_forward(b200nufft_plan_t p, const b200_c64* x, b200_c64* y, int nb, void* stream) {
    ARG_CHECK(p && x && y && nb >= 1, "forward: bad arguments");
    ON_DEVICE(p->device);
    return forward_impl(p, reinterpret_cast<const float2*>(x), 0, nullptr, reinterpret_cast<float2*>(y), nb, stream);
}

extern "C" int b200nufft_adjoint(b200nufft_plan_t p, const b200_c64* y, b200_c64* x, int nb, void* stream) {
    ARG_CHECK(p && x && y && nb >= 1, "adjoint: bad arguments");
    ON_DEVICE(p->device);
    return adjoint_impl(p, reinterpret_cast<const float2*>(y), reinterpret_cast<float2*>(x), nb, 0, nullptr, stream);
}

extern "C" int b200nufft_forward_one2many(b200nufft_plan_t p, const b200_c64* s, const b200_c64* sens,
                                          b200_c64* y, int nb, void* stream) {
    ARG_CHECK(p && s && y && nb >= 1, "forward_one2many: bad arguments");
    ON_DEVICE(p->device);
    return forward_impl(p, reinterpret_cast<const float2*>(s), 1, reinterpret_cast<const float2*>(sens),
                        reinterpret_cast<float2*>(y), nb, stream);
}

extern "C" int b200nufft_adjoint_many2one(b200nufft_plan_t p, const b200_c64* y, const b200_c64* sens,
                                          b200_c64* s, int nb, void* stream) {
    ARG_CHECK(p && s && y && nb >= 1, "adjoint_many2one: bad arguments");
    ON_DEVICE(p->device);
    return adjoint_impl(p, reinterpret_cast<const float2*>(y), reinterpret_cast<float2*>(s), nb, 1,
                        reinterpret_cast<const float2*>(sens), stream);
}

static int ensure_io(b200nufft_plan_t p, int nb) {
    if (p->io_nb >= nb) return B200_OK;
    if (p->d_xin) { cudaFree(p->d_xin); p->d_xin = nullptr; }
    if (p->d_yio) { cudaFree(p->d_yio); p->d_yio = nullptr; }
    p->io_nb = 0;
    CUDA_TRY(cudaMalloc(&p->d_xin, sizeof(float2) * p->g.Nprod * nb));
    CUDA_TRY(cudaMalloc(&p->d_yio, sizeof(float2) * std::max<long long>(p->M, 1) * nb));
    p->io_nb = nb;
    return B200_OK;
}

extern "C" int b200nufft_forward_host(b200nufft_plan_t p, const b200_c64* x_host, b200_c64* y_host, int nb,
                                      void* stream) {
    ARG_CHECK(p && x_host && y_host && nb >= 1, "forward_host: bad arguments");
    ON_DEVICE(p->device);
    int rc = ensure_io(p, nb);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemcpyAsync(p->d_xin, x_host, sizeof(float2) * p->g.Nprod * nb, cudaMemcpyHostToDevice, st));
    rc = forward_impl(p, p->d_xin, 0, nullptr, p->d_yio, nb, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(y_host, p->d_yio, sizeof(float2) * p->M * nb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

extern "C" int b200nufft_adjoint_host(b200nufft_plan_t p, const b200_c64* y_host, b200_c64* x_host, int nb,
                                      void* stream) {
    ARG_CHECK(p && x_host && y_host && nb >= 1, "adjoint_host: bad arguments");
    ON_DEVICE(p->device);
    int rc = ensure_io(p, nb);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    CUDA_TRY(cudaMemcpyAsync(p->d_yio, y_host, sizeof(float2) * p->M * nb, cudaMemcpyHostToDevice, st));
    rc = adjoint_impl(p, p->d_yio, p->d_xin, nb, 0, nullptr, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(x_host, p->d_xin, sizeof(float2) * p->g.Nprod * nb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

// ---- pipelined host entry points ------------------------------------------------------------------------------
// forward_host / adjoint_host above serialise copy-in, compute and copy-out.  The *_host_async pair overlaps them
// across calls: H2D on a copy-in stream, the operator on the caller's stream, D2H on a copy-out stream, chained by
// events, with B200NUFFT_HOST_SLOTS staging slots per direction (copy-in, operator and copy-out of three consecutive
// calls can be in progress at once); nothing blocks the host until b200nufft_host_wait(op, slot).
void host_pipe_destroy(b200nufft_plan_t p) {
    b200nufft_plan_s::HostPipe& h = p->pipe;
    constexpr int NS = b200nufft_plan_s::HostPipe::NSLOT;
    for (int o = 0; o < 2; ++o)
        for (int s = 0; s < NS; ++s) {
            cudaFree(h.d_in[o][s]);
            cudaFree(h.d_out[o][s]);
            h.d_in[o][s] = h.d_out[o][s] = nullptr;
            if (h.ready) {
                cudaEventDestroy(h.ev_in[o][s]);
                cudaEventDestroy(h.ev_comp[o][s]);
                cudaEventDestroy(h.ev_out[o][s]);
            }
        }
    if (h.ready) {
        cudaStreamDestroy(h.s_in);
        cudaStreamDestroy(h.s_out);
    }
    h.ready = false;
    h.nb = 0;
}

static int ensure_pipe(b200nufft_plan_t p, int nb) {
    b200nufft_plan_s::HostPipe& h = p->pipe;
    constexpr int NS = b200nufft_plan_s::HostPipe::NSLOT;
    if (!h.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h.s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&h.s_out, cudaStreamNonBlocking));
        for (int o = 0; o < 2; ++o)
            for (int s = 0; s < NS; ++s) {
                CUDA_TRY(cudaEventCreateWithFlags(&h.ev_in[o][s], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&h.ev_comp[o][s], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&h.ev_out[o][s], cudaEventDisableTiming));
            }
        h.ready = true;
    }
    if (h.nb >= nb) return B200_OK;
    CUDA_TRY(cudaDeviceSynchronize());                  // nothing may still use the old buffers
    const size_t xb = sizeof(float2) * p->g.Nprod * nb, yb = sizeof(float2) * std::max<long long>(p->M, 1) * nb;
    for (int s = 0; s < NS; ++s) {
        cudaFree(h.d_in[0][s]); cudaFree(h.d_out[0][s]); cudaFree(h.d_in[1][s]); cudaFree(h.d_out[1][s]);
        h.d_in[0][s] = h.d_out[0][s] = h.d_in[1][s] = h.d_out[1][s] = nullptr;
        h.nb = 0;
        CUDA_TRY(cudaMalloc(&h.d_in[0][s], xb));        // forward: x in, y out
        CUDA_TRY(cudaMalloc(&h.d_out[0][s], yb));
        CUDA_TRY(cudaMalloc(&h.d_in[1][s], yb));        // adjoint: y in, x out
        CUDA_TRY(cudaMalloc(&h.d_out[1][s], xb));
    }
    h.nb = nb;
    return B200_OK;
}

static int host_async(b200nufft_plan_t p, int op, const b200_c64* in_host, b200_c64* out_host, int nb, int slot,
                      void* stream) {
    ARG_CHECK(p && in_host && out_host && nb >= 1 && slot >= 0 && slot < B200NUFFT_HOST_SLOTS, "host_async: bad arguments");
    ON_DEVICE(p->device);
    int rc = ensure_pipe(p, nb);
    if (rc) return rc;
    rc = ensure_scratch(p, nb);
    if (rc) return rc;
    b200nufft_plan_s::HostPipe& h = p->pipe;
    cudaStream_t st = as_stream(stream);
    const size_t xb = sizeof(float2) * p->g.Nprod * nb, yb = sizeof(float2) * p->M * nb;
    const size_t inb = op == 0 ? xb : yb, outb = op == 0 ? yb : xb;
    // copy-in: after the operator that last read this slot's input
    CUDA_TRY(cudaStreamWaitEvent(h.s_in, h.ev_comp[op][slot], 0));
    if (inb) CUDA_TRY(cudaMemcpyAsync(h.d_in[op][slot], in_host, inb, cudaMemcpyHostToDevice, h.s_in));
    CUDA_TRY(cudaEventRecord(h.ev_in[op][slot], h.s_in));
    // operator: after its input arrived and the previous result of this slot left
    CUDA_TRY(cudaStreamWaitEvent(st, h.ev_in[op][slot], 0));
    CUDA_TRY(cudaStreamWaitEvent(st, h.ev_out[op][slot], 0));
    rc = op == 0 ? forward_impl(p, h.d_in[0][slot], 0, nullptr, h.d_out[0][slot], nb, stream)
                 : adjoint_impl(p, h.d_in[1][slot], h.d_out[1][slot], nb, 0, nullptr, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h.ev_comp[op][slot], st));
    // copy-out
    CUDA_TRY(cudaStreamWaitEvent(h.s_out, h.ev_comp[op][slot], 0));
    if (outb) CUDA_TRY(cudaMemcpyAsync(out_host, h.d_out[op][slot], outb, cudaMemcpyDeviceToHost, h.s_out));
    CUDA_TRY(cudaEventRecord(h.ev_out[op][slot], h.s_out));
    return B200_OK;
}

extern "C" int b200nufft_forward_host_async(b200nufft_plan_t p, const b200_c64* x_host, b200_c64* y_host, int nb,
                                            int slot, void* stream) {
    return host_async(p, 0, x_host, y_host, nb, slot, stream);
}
extern "C" int b200nufft_adjoint_host_async(b200nufft_plan_t p, const b200_c64* y_host, b200_c64* x_host, int nb,
                                            int slot, void* stream) {
    return host_async(p, 1, y_host, x_host, nb, slot, stream);
}
extern "C" int b200nufft_host_wait(b200nufft_plan_t p, int op, int slot) {
    ARG_CHECK(p && (op == 0 || op == 1) && slot >= 0 && slot < B200NUFFT_HOST_SLOTS, "host_wait: bad arguments");
    if (!p->pipe.ready) return B200_OK;
    ON_DEVICE(p->device);
    CUDA_TRY(cudaEventSynchronize(p->pipe.ev_out[op][slot]));
    return B200_OK;
}
