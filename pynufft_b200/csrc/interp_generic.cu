// Generic interpolation (gather) and gridding (scatter) kernels: any ndim <= 3, any J <= 16.
// One warp per bin-sorted sample, lanes stride over the prod(J) neighbours, the tensor product
// of the per-dimension factors is formed on the fly (as pELL_spmv_mCoil / pELL_spmvh_mCoil do,
// src/re_subroutine.py:751-835, 527-596) but from real factors + constant phase tables held in
// shared memory, with bin-sorted samples so neighbouring warps hit the same L2 lines.
// These are the reference-shaped kernels; the tiled shared-memory kernels (interp_tiled.cu,
// grid_tiled.cu) replace them for the geometries they support.
#include "common.cuh"

#define GEN_WARPS 8

struct WarpFactors {
    float2 a[MAXD * MAXJ];   // a_d[j] = c_d[j] * E_d[j]
    int col[MAXD * MAXJ];    // wrapped index * stride
};

__device__ __forceinline__ void load_factors(const Geom& g, const float* __restrict__ r, WarpFactors& wf,
                                             int lane) {
    for (int t = lane; t < g.sumJ; t += 32) {
        int d = 0;
        if (g.ndim > 1 && t >= g.Joff[1]) d = 1;
        if (g.ndim > 2 && t >= g.Joff[2]) d = 2;
        int j = t - g.Joff[d];
        float cv = r[t];
        float2 E = g.E[d][j];
        wf.a[t] = make_float2(cv * E.x, cv * E.y);
        int ks = reinterpret_cast<const int*>(r)[g.sumJ + 2 + d];
        int idx = ks + j;
        if (idx >= g.K[d]) idx -= g.K[d];
        wf.col[t] = (int)(idx * g.Kstride[d]);
    }
    __syncwarp();
}

__device__ __forceinline__ void neighbour(const Geom& g, const WarpFactors& wf, int f, float2& w, int& col) {
    // row-major decode of the flat neighbour id (meshindex, helper.py:376-383)
    int t = g.Joff[g.ndim - 1] + f % g.J[g.ndim - 1];
    w = wf.a[t];
    col = wf.col[t];
    f /= g.J[g.ndim - 1];
    for (int d = g.ndim - 2; d >= 0; --d) {
        t = g.Joff[d] + f % g.J[d];
        f /= g.J[d];
        w = cmul(w, wf.a[t]);
        col += wf.col[t];
    }
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
k_interp_generic(Geom g, const float* __restrict__ rec, long long M, const float2* __restrict__ grid,
                 float2* __restrict__ y, int nb) {
    __shared__ WarpFactors swf[GEN_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = blockIdx.x * (long long)GEN_WARPS + wib;
    if (i >= M) return;
    const int c = blockIdx.y;
    const float* r = rec + i * g.recw;
    WarpFactors& wf = swf[wib];
    load_factors(g, r, wf, lane);
    const float2* gc = grid + (long long)c * g.Kprod;
    float2 acc = make_float2(0.f, 0.f);
    for (int f = lane; f < g.prodJ; f += 32) {
        float2 w;
        int col;
        neighbour(g, wf, f, w, col);
        cfma(acc, w, __ldg(gc + col));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (lane == 0) {
        float2 P = make_float2(r[g.sumJ], r[g.sumJ + 1]);
        int m = reinterpret_cast<const int*>(r)[g.sumJ + 2 + g.ndim];
        y[(long long)m * nb + c] = cmul(P, acc);
    }
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
k_gridding_generic(Geom g, const float* __restrict__ rec, long long M, const float2* __restrict__ y,
                   float2* __restrict__ grid, int nb) {
    __shared__ WarpFactors swf[GEN_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = blockIdx.x * (long long)GEN_WARPS + wib;
    if (i >= M) return;
    const int c = blockIdx.y;
    const float* r = rec + i * g.recw;
    WarpFactors& wf = swf[wib];
    load_factors(g, r, wf, lane);
    float2 P = make_float2(r[g.sumJ], r[g.sumJ + 1]);
    int m = reinterpret_cast<const int*>(r)[g.sumJ + 2 + g.ndim];
    float2 yv = cmulc(P, y[(long long)m * nb + c]);        // conj(P) * y
    float2* gc = grid + (long long)c * g.Kprod;
    for (int f = lane; f < g.prodJ; f += 32) {
        float2 w;
        int col;
        neighbour(g, wf, f, w, col);
        float2 v = cmulc(w, yv);                           // conj(w) * conj(P) * y
        atomicAdd(gc + col, v);                            // REDG.E.ADD.F32x2
    }
}

int interp_impl(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st, bool grid_modulated) {
    if (p->M == 0) return B200_OK;
    if (grid_modulated && bi_fused_mod(p, nb)) return sweep2d_interp(p, grid, y, nb, st, true);   // 2-D multi-coil
    if (grid_modulated) {
        if (!interp_takes_modulated(p)) {
            b200_set_error("interp: this plan / variant cannot read a phase-modulated grid");
            return B200_ERR_UNSUPPORTED;
        }
        if (interp_col_on_modulated(p)) return col3d_interp(p, grid, y, nb, st);
        return interp_tiled_launch(p, grid, y, nb, st, true);
    }
    if (use_bi(p, nb)) return sweep2d_interp(p, grid, y, nb, st);             // batch-innermost grid
    if (p->interp_variant != 1 && single2d_supported(p->g)) return single2d_interp(p, grid, y, nb, st);
    if (interp_uses_col(p)) {
        // true grid in: one modulation pass into the plan's scratch, then the column-sweep gather
        int rc = ensure_scratch(p, nb);
        if (rc) return rc;
        rc = col3d_modulate(p, grid, p->d_grid, nb, st);
        if (rc) return rc;
        return col3d_interp(p, p->d_grid, y, nb, st);
    }
    if (p->interp_variant != 1 && tiled_supported(p->g)) return interp_tiled_launch(p, grid, y, nb, st, false);
    if (p->interp_variant == 2) {
        b200_set_error("interp: tiled variant requested but geometry unsupported");
        return B200_ERR_UNSUPPORTED;
    }
    dim3 gr((unsigned)((p->M + GEN_WARPS - 1) / GEN_WARPS), nb);
    k_interp_generic<<<gr, GEN_WARPS * 32, 0, st>>>(p->g, p->d_rec, p->M, grid, y, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

// modulated_ok: the caller takes the grid phase-modulated when the column-sweep kernel ran (gridding_modulated(p));
// otherwise the true grid is returned (one extra pass after the column-sweep kernel).
int gridding_impl(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st, bool modulated_ok) {
    if (p->M > 0 && gridding_modulated(p))           // zeroes the grid itself, inside its pre-pass kernel; a caller that wants
        return col3d_gridding(p, y, grid, nb, st, false, !modulated_ok);   // the true grid gets the demodulating instantiation
    if (p->M > 0 && use_bi(p, nb))               // batch-innermost grid, zero-fills; modulated when the caller takes it so
        return sweep2d_gridding(p, y, grid, nb, st, modulated_ok && bi_fused_mod(p, nb));
    CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float2) * p->g.Kprod * nb, st));
    if (p->M == 0) return B200_OK;
    if (p->gridding_variant != 1 && single2d_supported(p->g)) return single2d_gridding(p, y, grid, nb, st);
    if (p->gridding_variant != 1 && tiled_supported(p->g)) return gridding_tiled_launch(p, y, grid, nb, st);
    if (p->gridding_variant == 2) {
        b200_set_error("gridding: tiled variant requested but geometry unsupported");
        return B200_ERR_UNSUPPORTED;
    }
    dim3 gr((unsigned)((p->M + GEN_WARPS - 1) / GEN_WARPS), nb);
    k_gridding_generic<<<gr, GEN_WARPS * 32, 0, st>>>(p->g, p->d_rec, p->M, y, grid, nb);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_interp(b200nufft_plan_t p, const b200_c64* grid, b200_c64* y, int nb, void* stream) {
    ARG_CHECK(p && grid && y && nb >= 1 && nb <= 65535, "interp: bad arguments");
    ON_DEVICE(p->device);
    return interp_impl(p, reinterpret_cast<const float2*>(grid), reinterpret_cast<float2*>(y), nb, as_stream(stream), false);
}

// interp on the phase-modulated grid b200nufft_gridding_modulated produces (k-space solvers); only when
// b200nufft_kspace_modulated(plan) == 1
extern "C" int b200nufft_interp_modulated(b200nufft_plan_t p, const b200_c64* grid, b200_c64* y, int nb, void* stream) {
    ARG_CHECK(p && grid && y && nb >= 1 && nb <= 65535, "interp: bad arguments");
    ON_DEVICE(p->device);
    return interp_impl(p, reinterpret_cast<const float2*>(grid), reinterpret_cast<float2*>(y), nb, as_stream(stream), true);
}
// 1 if interp_modulated / gridding_modulated / ifft_crop_modulated all work on the phase-modulated grid
extern "C" int b200nufft_kspace_modulated(b200nufft_plan_t p) {
    return (p && gridding_modulated(p) && interp_takes_modulated(p)) ? 1 : 0;
}
// the same for a call with nb coils: 2-D plans keep the grid modulated only on the batch-innermost path (even nb >= 8)
extern "C" int b200nufft_kspace_modulated_nb(b200nufft_plan_t p, int nb) {
    if (!p || nb < 1) return 0;
    if (p->M > 0 && bi_fused_mod(p, nb)) return 1;
    return (!use_bi(p, nb) && gridding_modulated(p) && interp_takes_modulated(p)) ? 1 : 0;
}

extern "C" int b200nufft_gridding(b200nufft_plan_t p, const b200_c64* y, b200_c64* grid, int nb, void* stream) {
    ARG_CHECK(p && grid && y && nb >= 1 && nb <= 65535, "gridding: bad arguments");
    ON_DEVICE(p->device);
    return gridding_impl(p, reinterpret_cast<const float2*>(y), reinterpret_cast<float2*>(grid), nb, as_stream(stream), false);
}

// gridding that leaves the grid in the form b200nufft_ifft_crop_modulated takes (phase-modulated when
// b200nufft_gridding_is_modulated(plan), the true grid otherwise)
extern "C" int b200nufft_gridding_modulated(b200nufft_plan_t p, const b200_c64* y, b200_c64* grid, int nb, void* stream) {
    ARG_CHECK(p && grid && y && nb >= 1 && nb <= 65535, "gridding: bad arguments");
    ON_DEVICE(p->device);
    return gridding_impl(p, reinterpret_cast<const float2*>(y), reinterpret_cast<float2*>(grid), nb, as_stream(stream), true);
}
