/*
 * b200nufft.h -- C ABI of libb200nufft.so: the B200 (sm_100a) replacement for PyNUFFT's
 * device NUFFT hot path.
 *
 * The reference has no FFI: its device path is a set of reikna/PyCUDA kernel launches made
 * from Python (nufft/_nufft_class_methods_device.py).  Each entry point below names the
 * reference interface it replaces (paths relative to the pynufft checkout); INTEGRATION.md
 * shows the ctypes binding a pynufft maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; b200nufft_last_error() gives the text
 *   - all device pointers are plain CUDA device pointers on the plan's device; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous and
 *     never synchronise unless documented
 *   - complex64 = interleaved float pairs (b200_c64)
 *   - "image" arrays have the reference layout Nd+(B,), batch innermost: element (n, c) at
 *     n*B + c (linalg/nufft_hsa.py:225-227); B = nb of the call
 *   - "data" arrays y have the reference layout (M, B): element (m, c) at m*B + c
 *   - "grid" arrays are COIL-MAJOR: nb contiguous Kd grids, element (idx, c) at c*prod(Kd)+idx.
 *     The Python layer exposes them as Kd+(B,) views of that storage.
 *   - one plan is not re-entrant across streams (it owns scratch grids).
 */
#ifndef B200NUFFT_H
#define B200NUFFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } b200_c64;
typedef struct b200nufft_plan_s* b200nufft_plan_t;

#define B200NUFFT_HOST_SLOTS 4
#define B200NUFFT_MAX_DIM 3
#define B200NUFFT_MAX_J 16
#define B200NUFFT_MAX_L 32

/* ---- status ------------------------------------------------------------------------------ */
const char* b200nufft_last_error(void);
int b200nufft_version(void);

/* ---- plan: replaces helper.plan(format='pELL') + _plan_device/_offload_device -------------
 * (src/_helper/helper.py:620-802, nufft/_nufft_class_methods_device.py:98-254,
 *  batch from linalg/nufft_hsa.py:152-242)
 *   om        : (M, ndim) float64, row-major, host or device pointer, radians in [-pi, pi]
 *   alpha     : ndim * B200NUFFT_MAX_L doubles; alpha[d*MAX_L + l], l = 0..alpha_len[d]-1
 *               (float32 values of helper.nufft_alpha_kb_fit :910-958 widened to double)
 *   Tmat      : ndim * MAX_J*MAX_J doubles; T_d[j1*Jd[d] + j2] at Tmat[d*MAX_J*MAX_J + ...]
 *               (helper.nufft_T :1060-1083)
 *   sn        : sum(Nd) floats, the concatenated 1-D scaling vectors (helper.Tensor_sn :246-282)
 * The per-sample part (helper.nufft_offset :900-907, OMEGA_k :164-210, nufft_r :1086-1117,
 * min_max :606-618, OMEGA_u :148-162) runs on the GPU in float64; samples are then bin-sorted
 * by the tile that holds their first neighbour (stable).  Synchronises `stream` before return. */
int b200nufft_plan_create(b200nufft_plan_t* out, int device, int ndim,
                          const int32_t* Nd, const int32_t* Kd, const int32_t* Jd,
                          int64_t M, const double* om, int batch,
                          const double* alpha, const int32_t* alpha_len,
                          const double* Tmat, const float* sn, void* stream);
/* replaces NUFFT.release (nufft/_nufft_class_methods_device.py:653-683) */
int b200nufft_plan_destroy(b200nufft_plan_t plan);

/* ---- plan introspection for parity tests (original sample order) ---------------------------
 * kindx : (M, sum(Jd)) uint32, udata : (M, sum(Jd)) complex64 -- exactly st['pELL'].kindx/.udata
 *         (helper.create_partialELL :346-393); k0 : (M, ndim) int32 = floor(om/gam - J/2);
 * perm  : (M,) int32, perm[i] = original index of the i-th sample in bin-sorted order;
 * tile  : 2*ndim int32 (HOST): tile edges then sub-tile edges; the bin key of a sample is
 *         key = tile_linear * subtiles_per_tile + subtile_linear of its first neighbour
 *         (k0+1) mod K, both C-order; perm == stable argsort(key).  Other outputs are DEVICE.   */
int b200nufft_plan_get_kindx(b200nufft_plan_t plan, uint32_t* kindx, void* stream);
int b200nufft_plan_get_udata(b200nufft_plan_t plan, b200_c64* udata, void* stream);
int b200nufft_plan_get_k0(b200nufft_plan_t plan, int32_t* k0, void* stream);
int b200nufft_plan_get_perm(b200nufft_plan_t plan, int32_t* perm, void* stream);
int b200nufft_plan_get_tile(b200nufft_plan_t plan, int32_t* tile_host);
int64_t b200nufft_plan_bytes(b200nufft_plan_t plan);

/* ---- stages ----------------------------------------------------------------------------------
 * x2xx : out = in * sn (div=0) or in / sn (div=1), image layout.  Replaces cTensorMultiply
 *        (src/re_subroutine.py:145-201) as used by _x2xx_device/_xx2x_device
 *        (nufft/_nufft_class_methods_device.py:314-331, 451-468) and the CG un-scale
 *        (linalg/solve_device.py:472-480).  in == out allowed.                               */
int b200nufft_x2xx(b200nufft_plan_t plan, const b200_c64* in, b200_c64* out, int nb, int div,
                   void* stream);
/* scale_pad : grid = zero-pad( x * [sn if apply_sn] * [sens if sens] ) into the [0,Nd) corner,
 *        writing EVERY grid element (no separate fill).  x is image layout with nb coils, or a
 *        single-coil image broadcast to nb coils when x_single != 0 (cPopulate).  sens is image
 *        layout (nb coils) or NULL.  Replaces cTensorMultiply + fill + cTensorCopy(+1)
 *        (re_subroutine.py:98-201; _x2xx_device, _xx2k_device :314-359) and s2x
 *        (cPopulate + cMultiplyVecInplace, linalg/nufft_hsa.py:358-388).                     */
int b200nufft_scale_pad(b200nufft_plan_t plan, const b200_c64* x, b200_c64* grid, int nb,
                        int apply_sn, int x_single, const b200_c64* sens, void* stream);
/* fft : in-place c2c FFT of nb coil-major grids over all axes; inverse = 0 forward, 1 inverse
 *        normalised by 1/prod(Kd) (numpy.fft.ifftn convention), 2 inverse unnormalised,
 *        3 forward of a grid that is zero outside the image corner (as scale_pad leaves it; pruned
 *        3-D plan), 4 inverse unnormalised whose result is only valid in the image corner (what
 *        crop_scale reads; pruned 3-D plan).  Replaces reikna.fft.FFT
 *        (_nufft_class_methods_device.py:246-249, 358, 431).  cuFFT.                          */
int b200nufft_fft(b200nufft_plan_t plan, b200_c64* grid, int nb, int inverse, void* stream);
/* interp : y[m,c] = sum_j w[m,j] grid_c[col(m,j)].  Replaces pELL_spmv_mCoil
 *        (re_subroutine.py:751-835; _k2y_device :361-386).                                    */
int b200nufft_interp(b200nufft_plan_t plan, const b200_c64* grid, b200_c64* y, int nb,
                     void* stream);
/* grid : grid_c[col(m,j)] = sum_m conj(w[m,j]) y[m,c]; zero-fills the grid itself.  Replaces
 *        pELL_spmvh_mCoil + atomic_add_float2 (re_subroutine.py:527-596, 275-287;
 *        _y2k_device :388-422).                                                               */
int b200nufft_gridding(b200nufft_plan_t plan, const b200_c64* y, b200_c64* grid, int nb,
                       void* stream);
/* crop_scale : image = crop(grid corner) * f, f = 1 (mode 0), sn (mode 1), 1/sn (mode 2).
 *        With combine != 0 the nb coils are reduced to ONE image:
 *        out[n] = (1/nb) sum_c conj(sens[n,c]) * crop_c[n] * f   (sens NULL = ones).
 *        Replaces cTensorCopy(-1) + cTensorMultiply (_k2xx_device/_xx2x_device :425-468) and x2s
 *        (cMultiplyConjVecInplace + cAggregate, linalg/nufft_hsa.py:628-656).                 */
int b200nufft_crop_scale(b200nufft_plan_t plan, const b200_c64* grid, b200_c64* x, int nb,
                         int mode, int combine, const b200_c64* sens, void* stream);

/* pad_fft : grid = FFT(zero-pad(x * [sn] * [sens])) = scale_pad followed by fft(3); ifft_crop : x =
 *        crop(IFFT(grid)) * f [combined over coils] = fft(4) followed by crop_scale with the 1/prod(Kd)
 *        folded in; the grid is scratch afterwards.  For Kd = 256^3 both run as three hand-written pruned
 *        FFT passes that read/write only the image corner (csrc/fft256.cu); otherwise cuFFT.
 *        Same reference lines as scale_pad / fft / crop_scale above.                                  */
int b200nufft_pad_fft(b200nufft_plan_t plan, const b200_c64* x, b200_c64* grid, int nb, int apply_sn,
                      int x_single, const b200_c64* sens, void* stream);
int b200nufft_ifft_crop(b200nufft_plan_t plan, b200_c64* grid, b200_c64* x, int nb, int mode,
                        int combine, const b200_c64* sens, void* stream);

/* Concurrency contract: a plan owns its scratch grids, sorted-data buffers, work counters, staging buffers and cuFFT
 * handles (the stream of a handle is rebound per call).  ONE plan is therefore single-stream and single-thread: calls on
 * the same plan must be issued from one host thread and ordered on one CUDA stream (the pipelined host entry points
 * manage their own copy streams and events).  Use one plan per stream / thread for concurrent work.  Every entry point
 * runs on the plan's device (or on the device that owns its first pointer argument) and restores the caller's current
 * device before it returns.                                                                                          */

/* ---- compositions (plan-owned scratch grids) -------------------------------------------------
 * forward  : _forward_device  (:555-572)  x image(nb) -> y (M, nb)
 * adjoint  : _adjoint_device  (:617-633)  y (M, nb)   -> x image(nb)        (= A^H y / prod(Kd))
 * one2many / many2one: forward_one2many / adjoint_many2one (linalg/nufft_hsa.py:723-769)        */
int b200nufft_forward(b200nufft_plan_t plan, const b200_c64* x, b200_c64* y, int nb, void* stream);
int b200nufft_adjoint(b200nufft_plan_t plan, const b200_c64* y, b200_c64* x, int nb, void* stream);
int b200nufft_forward_one2many(b200nufft_plan_t plan, const b200_c64* s, const b200_c64* sens,
                               b200_c64* y, int nb, void* stream);
int b200nufft_adjoint_many2one(b200nufft_plan_t plan, const b200_c64* y, const b200_c64* sens,
                               b200_c64* s, int nb, void* stream);
/* Host-buffer variants: the reference's NUFFT.forward / NUFFT.adjoint host wrappers
 * (_forward_host/_adjoint_host, nufft/_nufft_class_methods_cpu.py:366-379): H2D copy, device
 * path, D2H copy, stream synchronise.  x_host/y_host are host pointers (pinned or pageable).   */
int b200nufft_forward_host(b200nufft_plan_t plan, const b200_c64* x_host, b200_c64* y_host,
                           int nb, void* stream);
int b200nufft_adjoint_host(b200nufft_plan_t plan, const b200_c64* y_host, b200_c64* x_host,
                           int nb, void* stream);
/* Pipelined variants for streams of host arrays: the H2D copy, the operator (on `stream`) and the D2H copy of
 * successive calls overlap (two copy streams inside the plan, chained by events); `slot` (0 .. B200NUFFT_HOST_SLOTS-1)
 * selects one of the staging buffers per direction (three calls in flight keep all three stages busy).  Nothing blocks the host: call host_wait(op, slot) (op 0 forward, 1 adjoint) before
 * reading the output of, or reusing the host buffers handed to, the last call on that (op, slot).  Host buffers must
 * be pinned for the copies to overlap.  */
int b200nufft_forward_host_async(b200nufft_plan_t plan, const b200_c64* x_host, b200_c64* y_host, int nb, int slot,
                                 void* stream);
int b200nufft_adjoint_host_async(b200nufft_plan_t plan, const b200_c64* y_host, b200_c64* x_host, int nb, int slot,
                                 void* stream);
int b200nufft_host_wait(b200nufft_plan_t plan, int op, int slot);

/* ---- solver vector ops (fused replacements of the element-wise kernels the device solvers
 *      launch one by one, linalg/solve_device.py:351-481 and :74-275) --------------------------
 * Scalars stay on the device as double pairs (re, im): no per-iteration host round trip
 * (the reference fetches alpha/beta with .get() three times per iteration, :424,432,449).      */
/* out[0..1] += sum conj(a)*b  (cMultiplyConjVec + reikna Reduce, :375-377,410-412) ; out must be
 * zeroed by the caller (b200nufft_zero_scalars). */
int b200nufft_dotc(const b200_c64* a, const b200_c64* b, int64_t n, double* out, void* stream);
int b200nufft_zero_scalars(double* s, int n, void* stream);
/* alpha = rsold/pAp ; x += alpha p ; r -= alpha Ap ; rsnew += sum conj(r) r   (:414-446) */
int b200nufft_cg_update_xr(b200_c64* x, b200_c64* r, const b200_c64* p, const b200_c64* Ap,
                           const double* rsold, const double* pAp, double* rsnew, int64_t n,
                           void* stream);
/* beta = rsnew/rsold ; p = r + beta p   (:448-455) */
int b200nufft_cg_update_p(b200_c64* p, const b200_c64* r, const double* rsnew,
                          const double* rsold, int64_t n, void* stream);
/* r = b - Ax ; p = r ; rsold += sum conj(r) r   (:385-402) */
int b200nufft_cg_init(const b200_c64* b, const b200_c64* Ax, b200_c64* r, b200_c64* p,
                      double* rsold, int64_t n, void* stream);
/* a[i] = a[i] / b[i] (complex) -- `k /= uker`, solve_device.py:178 */
int b200nufft_cdiv(b200_c64* a, const b200_c64* b, int64_t n, void* stream);
/* a[i] = a[i] * b[i] (complex) -- `self.W * k` of the Toeplitz-style selfadjoint2,
 * nufft/_nufft_class_methods_cpu.py:216-222 */
int b200nufft_cmul(b200_c64* a, const b200_c64* b, int64_t n, void* stream);
/* out = a x + b y, host scalars a, b (complex as two doubles); y NULL or b == 0: out = a x and y is not read.
 * out may alias x or y.  The vector update of the Krylov family of the CPU solve (linalg/solve_cpu.py:226-288: scipy's
 * lsqr / lsmr / bicgstab / bicg / gmres / lgmres run these on host arrays; pynufft_b200/krylov.py keeps them on the device). */
int b200nufft_axpby(b200_c64* out, double a_re, double a_im, const b200_c64* x, double b_re, double b_im,
                    const b200_c64* y, int64_t n, void* stream);
/* L1TVOLS, all arrays single-coil images of the plan's Nd (solve_device.py:74-275):
 *   tv_rhs    : rhs = mu*AHyk + lambda * sum_p Dt_p(d_p - b_p)           (:133-154, cDiff :936-950)
 *   tv_shrink : z_p = D_p(x); s_p = z_p + b_p; s = hypot chain + 1e-6; t = shrink(s,1/lambda)/s;
 *               d_p = s_p t; b_p += z_p - d_p                               (:189-262, cHypot, cAnisoShrink)
 *   tv_bregman: AHyk -= (zf - AHy)                                          (:196-198, :269)
 * d and b are ndim arrays stored back to back: d_p at d + p*prod(Nd).                          */
int b200nufft_tv_rhs(b200nufft_plan_t plan, const b200_c64* AHyk, const b200_c64* d,
                     const b200_c64* b, float mu, float lambda, b200_c64* rhs, void* stream);
int b200nufft_tv_shrink(b200nufft_plan_t plan, const b200_c64* x, b200_c64* d, b200_c64* b,
                        float lambda, void* stream);
int b200nufft_tv_bregman(b200_c64* AHyk, const b200_c64* zf, const b200_c64* AHy, int64_t n,
                         void* stream);

/* ---- tuning / introspection ---------------------------------------------------------------- */
/* kernel variant selection: interp/gridding "auto" (0), "generic" (1), "tiled" (2); interp only: "column sweep" (3,
 * the register-resident gather of csrc/col3d.cu on the phase-modulated grid; 3-D, Jd = 6^3, Kd[0] >= 10).  "auto"
 * runs the column-sweep gather wherever the grid is phase-modulated already (forward on the fused FFT passes, the
 * k-space solvers) and the tiled gather on true grids; 3 modulates a true grid first and sweeps everywhere */
int b200nufft_set_variant(b200nufft_plan_t plan, int interp_variant, int gridding_variant);
/* Column-sweep gridding (csrc/col3d.cu).  Plans for 3-D, Jd = 6^3 keep, next to the tile-sorted samples, a second
 * copy sorted by (4 x 5 column of first-neighbour cells, first plane) for the register-resident scatter kernel.
 * set_layout_preference(0) disables it for plans created afterwards (1, the default, enables it);
 * plan_get_layout reports 1 if the plan holds the column-sweep records.  plan_get_col_perm exports the second
 * permutation (device pointer, M entries; may be NULL) and its sort key as (tile, sub-tile) edges of the generic
 * bin key (host pointer, 6 entries; may be NULL): tile = (K0, 4, 5), sub-tile = (1, 4, 5).                      */
int b200nufft_set_layout_preference(int pref);
/* Grid layout of the stage entry points (scale_pad, fft, interp, gridding, crop_scale, pad_fft, ifft_crop) for a call
 * with nb coils: 0 = coil-major (nb contiguous Kd grids), 1 = batch-innermost (k[K0][K1][nb], the reference's own
 * Kd + (batch,) order, linalg/nufft_hsa.py:225-227).  2-D Jd = 6^2 plans use 1 for even nb >= 8 (csrc/sweep2d.cu:
 * coil on the lanes, register-resident row sweep, strided-batch cuFFT); images and data are batch-innermost always. */
int b200nufft_grid_layout(b200nufft_plan_t plan, int nb);
int b200nufft_plan_get_layout(b200nufft_plan_t plan);
int b200nufft_plan_get_col_perm(b200nufft_plan_t plan, int32_t* perm, int32_t* tile_host, void* stream);
/* The column-sweep kernel produces the phase-modulated grid G'[g] = G[g] * prod_d exp(i gam_d (N_d - 1)/2 * g_d),
 * which makes every interpolation weight of the reference (helper.py:148-162, 606-618:
 * c_j e^{i om N/2} e^{-i gam (N-1)/2 (dk - j)}) a real number times one phase per sample.  b200nufft_gridding
 * returns the TRUE grid (one extra pass); gridding_modulated leaves it modulated (iff gridding_is_modulated) for
 * ifft_crop_modulated, whose inverse FFT passes undo the modulation for free.  adjoint / selfadjoint use the pair. */
int b200nufft_gridding_is_modulated(b200nufft_plan_t plan);
/* k-space solvers (linalg/solve_device.py:351-461) may iterate on modulated vectors, G' = D G D^H with D diagonal and
 * unitary has the same CG scalars: kspace_modulated(plan) == 1 means gridding_modulated, interp_modulated (the tiled
 * gather reading the modulated grid directly) and ifft_crop_modulated all agree on that form; saves one pass over the
 * grid per application of G = interp^H interp.                                                                 */
int b200nufft_kspace_modulated(b200nufft_plan_t plan);
/* The same question for a call with nb coils.  2-D Jd = 6^2 plans keep the grid modulated, G'[g] = G[g] m0[g0] m1[g1], only on
 * the batch-innermost path (even nb >= 8, power-of-two Kd): there the dim-0 FFT passes (csrc/fftbi.cu) apply / undo the
 * modulation and the row-sweep kernels (csrc/sweep2d.cu) skip it; with any other nb the *_modulated entries of a 2-D plan
 * work on the true grid. */
int b200nufft_kspace_modulated_nb(b200nufft_plan_t plan, int nb);
int b200nufft_interp_modulated(b200nufft_plan_t plan, const b200_c64* grid, b200_c64* y, int nb, void* stream);
int b200nufft_gridding_modulated(b200nufft_plan_t plan, const b200_c64* y, b200_c64* grid, int nb, void* stream);
int b200nufft_ifft_crop_modulated(b200nufft_plan_t plan, b200_c64* grid, b200_c64* x, int nb, int mode, int combine,
                                  const b200_c64* sens, void* stream);
/* pad_fft that leaves the grid phase-modulated, the form interp_modulated reads (the forward FFT passes multiply their
 * outputs by the modulation tables; cuFFT path: one extra pass).  forward / forward_one2many use the pair
 * pad_fft_modulated -> interp_modulated (the column-sweep gather, csrc/col3d.cu) when kspace_modulated(plan) == 1.  */
int b200nufft_pad_fft_modulated(b200nufft_plan_t plan, const b200_c64* x, b200_c64* grid, int nb, int apply_sn,
                                int x_single, const b200_c64* sens, void* stream);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t b200nufft_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H */
