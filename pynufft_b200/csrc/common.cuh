// Internal declarations shared by the translation units of libb200nufft.so.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/b200nufft.h"

#define MAXD B200NUFFT_MAX_DIM
#define MAXJ B200NUFFT_MAX_J
#define MAXL B200NUFFT_MAX_L

enum {
    B200_OK = 0,
    B200_ERR_ARG = -1,
    B200_ERR_CUDA = -2,
    B200_ERR_CUFFT = -3,
    B200_ERR_UNSUPPORTED = -4,
};

void b200_set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            b200_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
            return B200_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define CUFFT_TRY(expr)                                                                      \
    do {                                                                                     \
        cufftResult _e = (expr);                                                             \
        if (_e != CUFFT_SUCCESS) {                                                           \
            b200_set_error(std::string(#expr) + ": cufft error " + std::to_string((int)_e) + \
                           " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
            return B200_ERR_CUFFT;                                                           \
        }                                                                                    \
    } while (0)

#define ARG_CHECK(cond, msg)                                                                 \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            b200_set_error(std::string("invalid argument: ") + msg);                         \
            return B200_ERR_ARG;                                                             \
        }                                                                                    \
    } while (0)

// Every C entry point runs on its plan's device (or on the device that owns its first pointer argument) and RESTORES
// the caller's current device on exit: a plan on cuda:1 must not silently switch a multi-GPU process to device 1.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != dev) {
            err = cudaSetDevice(dev);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(dev)      \
    DeviceGuard _dg(dev);   \
    CUDA_TRY(_dg.err)
static inline int device_of(const void* ptr) {      // device that owns a device pointer (current device if unknown)
    cudaPointerAttributes a;
    if (ptr && cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
    cudaGetLastError();
    int cur = 0;
    cudaGetDevice(&cur);
    return cur;
}

#define LAUNCH_CHECK()                                                                       \
    do {                                                                                     \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
        CUDA_TRY(cudaGetLastError());                                                        \
    } while (0)

// Geometry handed to kernels by value.
struct Geom {
    int ndim;
    int N[MAXD];        // image size
    int K[MAXD];        // oversampled grid size
    int J[MAXD];        // interpolator width
    int Joff[MAXD];     // prefix sum of J
    int sumJ, prodJ;
    int tile[MAXD];     // coarse bin (tile) edge per dim
    int ntile[MAXD];    // tiles per dim
    int sub[MAXD];      // sub-tile edge per dim (divides tile); the bin key is (tile, sub-tile)
    int nsub[MAXD];     // sub-tiles per tile per dim
    int nsubprod;       // sub-tiles per tile
    int snoff[MAXD];    // offset of dim d inside the concatenated sn vector
    long long Nprod, Kprod;
    long long Kstride[MAXD];  // C-order element stride of the grid
    int recw;           // words (4 B) per sample record, multiple of 4
    float2 E[MAXD][MAXJ];     // E_d[j] = exp(+i gam_d (N_d-1)/2 (j+1)), j = 0..J_d-1
    // last-dimension phase split used by the tiled kernels: E_last[j] = Fl[cc] * Gl[rel] with box column
    // cc = rel + j:  Fl[cc] = exp(+i s (cc+1)), Gl[rel] = exp(-i s rel), s = gam (N-1)/2 of the last dim
    float2 Fl[40];            // 37 columns of the 2-D single-coil box
    float2 Gl[32];
    float2 F0[24];            // the same split for dimension 0 (2-D batch kernels fold both phases into the box)
    float2 G0[16];
};

// Per-dimension float64 constants of the min-max interpolator (plan kernels only).
struct PlanConst {
    double gam[MAXD];           // 2 pi / K           (helper.py:905)
    double ratio[MAXD];         // K / N              (helper.py:1093)
    double alpha[MAXD][MAXL];   // alpha_|l|, l = 0..L
    int L[MAXD];
    double T[MAXD][MAXJ * MAXJ];// T_d row-major J x J (helper.py:1060-1083)
};

// column-sweep gridding (col3d.cu): 4 x 5 cross-section columns of first-neighbour cells swept along dim 0;
// 32-word sample records in sweep order
constexpr int COL_T1 = 4, COL_T2 = 5, COL_RECW = 32, COL_SEG = 320;
// 2-D multi-coil row sweep (sweep2d.cu): strips of SW2_CT first-neighbour columns (dim 1) swept along dim 0, coil on the
// lanes; 20-word sample records in sweep order
constexpr int SW2_CT = 3, SW2_RECW = 20, SW2_SEG = 64;
// tiled 3-D gather (interp_tiled.cu): row / plane pitch (complex elements) of the staged 13 x 21 x 21 box; the plan bakes the
// box-relative address of a sample's first neighbour into its gather record
constexpr int TILE_RP = 21, TILE_PP = 446, TILE_RECW = 24;

struct WorkItem {   // one launch unit of the tiled kernels: samples [begin, end) of one tile
    int tile;
    int begin;
    int end;
    int pad;
};

struct b200nufft_plan_s {
    int device = 0;
    Geom g;
    PlanConst* d_pc = nullptr;      // device copy
    long long M = 0;
    int batch = 1;
    // device buffers
    double* d_om = nullptr;         // (M, ndim) original order
    int* d_perm = nullptr;          // (M,)
    float* d_rec = nullptr;         // (M, recw) sorted order
    float* d_trec = nullptr;        // (M, TILE_RECW) sorted order: records of the tiled 3-D gather, ready to use
    float* d_sn = nullptr;          // (sum N)
    int* d_bin_start = nullptr;     // (n_bins + 1), n_bins = n_tiles * nsubprod
    WorkItem* d_work = nullptr;     // interp: per tile
    int n_work = 0;
    WorkItem* d_gwork = nullptr;    // gridding: per sub-tile (WorkItem.tile = bin id)
    int n_gwork = 0;
    int n_tiles = 0;
    int n_bins = 0;
    // column-sweep gridding (3-D, J = 6; col3d.cu): a second copy of the samples sorted by (column, first plane)
    bool has_col = false;
    int* d_cperm = nullptr;         // (M,) sweep-order permutation (parity export)
    float* d_crec = nullptr;        // (M, COL_RECW) records in sweep order
    float4* d_cside = nullptr;      // (M,) (P''.re, P''.im, original index, 0) in sweep order
    WorkItem* d_cwork = nullptr;    // (column, begin, end) segments of at most COL_SEG samples
    int n_cwork = 0;
    // 2-D multi-coil row sweep (2-D, J = 6; sweep2d.cu): samples sorted by (strip, first row), batch-innermost grids
    bool has_sw2 = false;
    int* d_sw_perm = nullptr;       // (M,) sweep-order permutation (parity export)
    float* d_sw_rec = nullptr;      // (M, SW2_RECW)
    WorkItem* d_sw_work = nullptr;  // (strip, begin, end) segments of at most SW2_SEG samples
    int n_sw_work = 0;
    cufftHandle fft_bi = 0;         // strided-batch cuFFT plan for the batch-innermost grid layout
    int fft_bi_nb = 0;
    bool fft_bi_valid = false;
    bool attr_sw2 = false;
    float2* d_fbi_tw = nullptr;     // fftbi.cu: twiddle tables [K0 | K1]
    int* d_fbi_rev = nullptr;       // fftbi.cu: digit-reversal tables [K0 | K1]
    int* d_ccount = nullptr;        // per-coil work counters of the persistent column kernels
    int ccount_nb = 0;
    int n_sm = 0;
    float2* d_mod = nullptr;        // modulation tables m_d[g] = exp(i s_d g), concatenated over dims (sum K)
    // bin-sorted copy of the data for the tiled gridding kernel (16-byte slots)
    float4* d_ys = nullptr;
    int ys_nb = 0;
    // sweep-ordered, phase-folded copy of the data for the column-sweep gridding kernel (8-byte slots, even coil stride)
    float2* d_ys2 = nullptr;
    int ys2_nb = 0;
    float2* d_ysb = nullptr;        // bin-sorted rows y[perm[i], :] for the 2-D batch kernels
    long long ysb_elems = 0;
    // scratch grids for the compositions
    float2* d_grid = nullptr;
    int grid_nb = 0;
    float2* d_grid2 = nullptr;      // coil-major FFT scratch of the batch-innermost 2-D path (sweep2d.cu)
    int grid2_nb = 0;
    // adjoint grid of the column-sweep path, kept zeroed BETWEEN calls: the zero-fill of the next call is issued on a
    // side stream behind the inverse FFT passes of this one (stages.cu adjoint_impl), off the caller's stream
    float2* d_gridz = nullptr;
    int gridz_nb = 0;
    bool gridz_clean = false;       // a zero-fill is in flight / done on zs; gridz_ev marks its end
    cudaStream_t zs = nullptr;
    cudaEvent_t gridz_ev = nullptr, gridz_use = nullptr;
    // host staging for the *_host entry points
    float2* d_xin = nullptr;
    float2* d_yio = nullptr;
    int io_nb = 0;
    // pipelined host entry points (stages.cu): copy streams, per-(direction, slot) staging buffers and events
    struct HostPipe {
        static constexpr int NSLOT = B200NUFFT_HOST_SLOTS;
        cudaStream_t s_in = nullptr, s_out = nullptr;
        float2* d_in[2][NSLOT] = {};      // [0 forward | 1 adjoint][slot]
        float2* d_out[2][NSLOT] = {};
        cudaEvent_t ev_in[2][NSLOT], ev_comp[2][NSLOT], ev_out[2][NSLOT];
        int nb = 0;
        bool ready = false;
    } pipe;
    // cuFFT
    cufftHandle fft = 0;
    int fft_nb = 0;
    bool fft_valid = false;
    cufftHandle fft2d = 0, fft1d = 0;   // pruned 3-D transform (stages.cu)
    bool fftp_valid = false;
    // fused pruned FFT passes (fft256.cu)
    float2* d_tw256 = nullptr;
    float2* d_xc = nullptr;         // per-coil image scratch for many2one
    int xc_nb = 0;
    bool attr_grid = false, attr_interp = false, attr_col = false;   // cudaFuncSetAttribute done on this plan's device
    int interp_variant = 0, gridding_variant = 0, fft_variant = 0;   // 0 auto, 1 generic / cuFFT, 2 tiled, 3 column sweep (interp)
    long long bytes = 0;
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// 2-D calls with an even number (>= 8) of coils run on BATCH-INNERMOST grids, k[K0][K1][nb] (the reference's own layout,
// linalg/nufft_hsa.py:225-227), with the coil on the lanes (sweep2d.cu); every other call keeps coil-major grids
static inline bool use_bi(const b200nufft_plan_s* p, int nb) {
    return p->has_sw2 && p->interp_variant != 1 && p->gridding_variant != 1 && nb >= 8 && (nb & 1) == 0;
}
// batch-innermost 2-D path with the fused FFT passes: the grid between the FFT and the sweep kernels is kept phase-
// modulated inside forward / adjoint (the passes along dim 0 apply / undo the modulation; the kernels skip it)
bool fftbi_supported(const Geom& g);
static inline bool bi_fused_mod(const b200nufft_plan_s* p, int nb) {
    static const bool off = [] { const char* e = getenv("B200NUFFT_BI_NOMOD"); return e && atoi(e) != 0; }();
    return !off && use_bi(p, nb) && p->fft_variant != 1 && p->d_mod && fftbi_supported(p->g);
}

// tiled kernels (interp_tiled.cu / grid_tiled.cu); return B200_ERR_UNSUPPORTED if geometry does not fit
// modulated: `grid` is the phase-modulated grid of the column-sweep gridding kernel (needs the plan's tables)
int interp_tiled_launch(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st, bool modulated);
int gridding_tiled_launch(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st);
bool tiled_supported(const Geom& g);
int ensure_scratch(b200nufft_plan_t p, int nb);
int ensure_scratch2(b200nufft_plan_t p, int nb);
// col3d.cu: register-resident column-sweep gridding (3-D, J = 6); the grid it produces is phase-modulated
bool col3d_supported(const Geom& g);
// zeroes the grid itself (inside its pre-pass kernel)
int col3d_gridding(b200nufft_plan_t p, const float2* y, float2* grid_mod, int nb, cudaStream_t st, bool prezeroed = false,
                   bool demodulate = false);
int col3d_demodulate(b200nufft_plan_t p, float2* grid, int nb, cudaStream_t st);
// register-resident column-sweep gather on the phase-modulated grid (needs K0 >= 8)
bool col3d_interp_supported(const Geom& g);
int col3d_interp(b200nufft_plan_t p, const float2* grid_mod, float2* y, int nb, cudaStream_t st);
int col3d_modulate(b200nufft_plan_t p, const float2* in, float2* out, int nb, cudaStream_t st);
// the gridding output is phase-modulated iff the column-sweep kernel runs (gridding variant "auto")
static inline bool gridding_modulated(const b200nufft_plan_s* p) { return p->has_col && p->gridding_variant == 0; }
// interp variant 3: the column-sweep gather everywhere (a true grid is modulated first).  "auto" (0) uses it wherever the
// grid is phase-modulated already -- forward() on the fused FFT passes, the k-space solvers -- since its planes arrive
// through a cp.async ring (237 us against 252 us for the tiled gather at configuration 3); true grids handed to the
// interp stage stay on the tiled gather, which needs no modulation pass.
static inline bool interp_uses_col(const b200nufft_plan_s* p) {
    return p->has_col && p->d_mod && p->interp_variant == 3 && col3d_interp_supported(p->g);
}
static inline bool interp_col_on_modulated(const b200nufft_plan_s* p) {
    return p->has_col && p->d_mod && (p->interp_variant == 0 || p->interp_variant == 3) && col3d_interp_supported(p->g);
}
// the gather can read that modulated grid directly (k-space solvers iterate on modulated vectors): column-sweep gather,
// or the tiled gather's MOD instantiation
static inline bool interp_takes_modulated(const b200nufft_plan_s* p) {
    return p->has_col && p->d_mod && p->interp_variant != 1 && (interp_uses_col(p) || tiled_supported(p->g));
}
// grid_modulated: `grid` is phase-modulated (only legal when interp_takes_modulated(p))
int interp_impl(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st, bool grid_modulated);
// modulated_ok: the caller accepts the phase-modulated grid (it hands it to ifft_crop_impl(..., modulated))
int gridding_impl(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st, bool modulated_ok);
// fft256.cu; `modulated`: the grid enters phase-modulated (col3d.cu)
bool fft256_supported(const Geom& g);
int fft256_forward(b200nufft_plan_t p, const float2* x, float2* grid, int nb, int apply_sn, int x_single,
                   const float2* sens, bool modulated, cudaStream_t st);
int fft256_inverse(b200nufft_plan_t p, float2* grid, float2* x, int nb, int mode, float scale, bool modulated,
                   cudaStream_t st);
int combine_coils(const float2* xc, const float2* sens, float2* s, long long N, int nb, cudaStream_t st);
// single2d.cu: 2-D kernels for fewer than 8 coils
bool single2d_supported(const Geom& g);
int single2d_interp(b200nufft_plan_t p, const float2* grid, float2* y, int nb, cudaStream_t st);
int single2d_gridding(b200nufft_plan_t p, const float2* y, float2* grid, int nb, cudaStream_t st);
// sweep2d.cu: 2-D multi-coil kernels with the coil on the lanes, register-resident row sweep, batch-innermost grids
bool sweep2d_supported(const Geom& g);
int sweep2d_interp(b200nufft_plan_t p, const float2* grid_bi, float2* y, int nb, cudaStream_t st, bool modulated = false);
int sweep2d_gridding(b200nufft_plan_t p, const float2* y, float2* grid_bi, int nb, cudaStream_t st, bool modulated = false);   // zero-fills
int sweep2d_scale_pad(b200nufft_plan_t p, const float2* x, float2* grid_bi, int nb, int apply_sn, int x_single,
                      const float2* sens, cudaStream_t st);
int sweep2d_crop_scale(b200nufft_plan_t p, const float2* grid_bi, float2* x, int nb, int mode, int combine,
                       const float2* sens, float scale, cudaStream_t st);
int sweep2d_fft(b200nufft_plan_t p, float2* grid_bi, int nb, int inverse, cudaStream_t st);
int sweep2d_pad_fft(b200nufft_plan_t p, const float2* x, float2* grid_bi, int nb, int apply_sn, int x_single,
                    const float2* sens, cudaStream_t st);
// fftbi.cu: fused, pruned FFT passes on batch-innermost grids (power-of-two Kd, 64 .. 1024)
bool fftbi_supported(const Geom& g);
int fftbi_forward(b200nufft_plan_t p, const float2* x, float2* grid_bi, int nb, int apply_sn, int x_single,
                  const float2* sens, cudaStream_t st, bool modulate = false);
int fftbi_inverse(b200nufft_plan_t p, float2* grid_bi, float2* x, int nb, int mode, float scale, cudaStream_t st,
                  bool demodulate = false);
int sweep2d_ifft_to_scratch(b200nufft_plan_t p, const float2* grid_bi, int nb, cudaStream_t st);   // coil-major result in p->d_grid2


// ---- small device helpers ---------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // conj(a) * b
    return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {  // acc += a*b
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
