// Plan: per-sample min-max interpolator (float64), bin keys, stable sort, sample records.
// Follows src/_helper/helper.py:148-210 (OMEGA_u, OMEGA_k), :606-618 (min_max), :900-907
// (nufft_offset), :1086-1117 (nufft_r) of the reference; what is stored is new (per-dimension
// REAL factors + one complex phase per sample instead of M x sum(J) complex values).
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "common.cuh"

static thread_local std::string g_err;
std::atomic<long long> g_launches{0};
void b200_set_error(const std::string& msg) { g_err = msg; }

extern "C" const char* b200nufft_last_error(void) { return g_err.c_str(); }
extern "C" int b200nufft_version(void) { return 100; }
extern "C" int64_t b200nufft_launch_count(void) { return g_launches.load(); }

// preference for plans created afterwards: 1 = also build the column-sweep gridding records where supported
// (default), 0 = tile bins only
static std::atomic<int> g_layout_pref{1};
extern "C" int b200nufft_set_layout_preference(int pref) {
    ARG_CHECK(pref == 0 || pref == 1, "layout preference must be 0 or 1");
    g_layout_pref.store(pref);
    return B200_OK;
}

// ------------------------------------------------------------------------------------------
// float64 per-(sample, dim) math
// ------------------------------------------------------------------------------------------
struct DimResult {
    long long k0;      // floor(om/gam - J/2)                       (helper.py:905-906)
    double dk;         // om/gam - k0                               (helper.py:1108)
    double c[MAXJ];    // T . r                                     (helper.py:616)
};

__device__ __forceinline__ double np_sinc(double x) {  // numpy.sinc
    double y = 3.141592653589793 * (x == 0.0 ? 1.0e-20 : x);
    return sin(y) / y;
}

__device__ __forceinline__ long long offset_k0(double om, double gam, int J, double* q_out) {
    // numpy: floor(1.0*om/gam - 1.0*J/2.0); explicit _rn ops forbid fma contraction
    double q = __ddiv_rn(om, gam);
    *q_out = q;
    return (long long)floor(__dsub_rn(q, 0.5 * (double)J));
}

__device__ void dim_math(double om, int d, const Geom& g, const PlanConst* __restrict__ pc,
                         DimResult& out) {
    const int J = g.J[d];
    const int L = pc->L[d];
    double q;
    long long k0 = offset_k0(om, pc->gam[d], J, &q);
    double dk = __dsub_rn(q, (double)k0);
    out.k0 = k0;
    out.dk = dk;
    const double ratio = pc->ratio[d];
    double r[MAXJ];
    for (int j = 0; j < J; ++j) {
        double arg = -(double)(j + 1) + dk;           // outer_sum(-arange(1,J+1), dk)
        double rr = 0.0;
        for (int l = -L; l <= L; ++l) {               // same accumulation order as nufft_r
            double a = pc->alpha[d][l < 0 ? -l : l];
            rr = rr + a * np_sinc((arg + 1.0 * l) / ratio);   // beta == 1
        }
        r[j] = rr;
    }
    const double* T = pc->T[d];
    for (int j1 = 0; j1 < J; ++j1) {
        double s = 0.0;
        for (int j2 = 0; j2 < J; ++j2) s += T[j1 * J + j2] * r[j2];
        out.c[j1] = s;
    }
}

__device__ __forceinline__ int wrap_index(long long v, int K) {  // numpy.mod for integers
    long long m = v % K;
    if (m < 0) m += K;
    return (int)m;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__global__ void k_bin_keys(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om,
                           long long M, int* __restrict__ keys, int* __restrict__ vals) {
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= M) return;
    int key = 0, skey = 0;
    for (int d = 0; d < g.ndim; ++d) {
        double q;
        long long k0 = offset_k0(om[m * g.ndim + d], pc->gam[d], g.J[d], &q);
        int ks = wrap_index(k0 + 1, g.K[d]);
        int t = ks / g.tile[d];
        key = key * g.ntile[d] + t;
        skey = skey * g.nsub[d] + (ks - t * g.tile[d]) / g.sub[d];
    }
    keys[m] = key * g.nsubprod + skey;
    vals[m] = (int)m;
}

// one thread per sorted sample: gather om through perm, write the record
// record words: [c_0[0..J0) | c_1 | c_2 | P.re P.im | kstart_0.. | perm | pad]
// trec (3-D tiled gather, may be NULL): [c_0[6] c_1[6] c_2[6] | base perm | P'.re P'.im | P''.re P''.im], where
//   base = rel0 * TILE_PP + rel1 * TILE_RP + rel2, rel = first neighbour relative to its 8 x 16 x 16 slab,
//   P'  = P * exp(-i s_2 rel2)                       (true grid: the staged box carries exp(+i s_2 (column + 1)))
//   P'' = prod_d exp(i (om N/2 - s dk - s (k0' - 1))) (phase-modulated grid)
__global__ void k_build_records(Geom g, const PlanConst* __restrict__ pc,
                                const double* __restrict__ om, const int* __restrict__ perm,
                                long long M, float* __restrict__ rec, float* __restrict__ trec) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int m = perm[i];
    float* out = rec + i * g.recw;
    double pr = 1.0, pi = 0.0;
    double ph1 = 0.0, ph2 = 0.0;
    int ksv[MAXD] = {0, 0, 0};
    for (int d = 0; d < g.ndim; ++d) {
        DimResult R;
        double o = om[(long long)m * g.ndim + d];
        dim_math(o, d, g, pc, R);
        for (int j = 0; j < g.J[d]; ++j) out[g.Joff[d] + j] = (float)R.c[j];
        if (trec) {
            const int k = wrap_index(R.k0 + 1, g.K[d]);
            ksv[d] = k;
            const double sd = pc->gam[d] * ((double)g.N[d] - 1.0) / 2.0;
            const double base_ph = o * (double)g.N[d] / 2.0 - sd * R.dk;
            ph1 += base_ph - (d == g.ndim - 1 ? sd * (double)(k & 15) : 0.0);
            ph2 += base_ph - sd * (double)(k - 1);
            for (int j = 0; j < g.J[d]; ++j) trec[i * TILE_RECW + g.Joff[d] + j] = (float)R.c[j];
        }
        // P_d = exp(i (om N/2 - gam (N-1)/2 dk))
        double s = pc->gam[d] * ((double)g.N[d] - 1.0) / 2.0;
        double ph = o * (double)g.N[d] / 2.0 - s * R.dk;
        double sn, cs;
        sincos(ph, &sn, &cs);
        double npr = pr * cs - pi * sn;
        double npi = pr * sn + pi * cs;
        pr = npr;
        pi = npi;
        reinterpret_cast<int*>(out)[g.sumJ + 2 + d] = wrap_index(R.k0 + 1, g.K[d]);
    }
    out[g.sumJ + 0] = (float)pr;
    out[g.sumJ + 1] = (float)pi;
    reinterpret_cast<int*>(out)[g.sumJ + 2 + g.ndim] = m;
    for (int w = g.sumJ + 3 + g.ndim; w < g.recw; ++w) out[w] = 0.f;
    if (trec) {
        float* t = trec + i * TILE_RECW;
        reinterpret_cast<int*>(t)[18] = (ksv[0] & 7) * TILE_PP + (ksv[1] & 15) * TILE_RP + (ksv[2] & 15);
        reinterpret_cast<int*>(t)[19] = m;
        double sn, cs;
        sincos(ph1, &sn, &cs);
        t[20] = (float)cs;
        t[21] = (float)sn;
        sincos(ph2, &sn, &cs);
        t[22] = (float)cs;
        t[23] = (float)sn;
    }
}

// ---- column-sweep gridding (col3d.cu): key = (column, first plane), 32-word records in sweep order ----
// A column is a COL_T1 x COL_T2 cross-section (dims 1, 2) of first-neighbour cells; its box is 9 rows x 10 columns.
// record words: [c1g[0..11] | c0[0..5] | p0 | run | w10[0..9] | 0 0]
//   c1g[4 g + i] = c1t[g + 3 i] (g, i = 0..2), c1t[b] = c1[b - k1rel] (zero outside the footprint): box rows g, g+3, g+6
//   c0[j]: weight of plane p0 + j (the kernels are unrolled over the window phase, so no rotation is stored)
//   run: samples from this one to the end of its (column, first plane) bin -- the kernels' counted inner loop
//   w10[c] = c2[c - k2rel] (zero outside the footprint): box columns
//   every weight of a neighbour that wraps around the periodic grid (k + j >= K_d) carries the factor
//   e^{i s_d K_d} = (-1)^(N_d - 1) of the phase-modulated grid, so that the kernels never see a sign
// side: (P''.re, P''.im, original index, 0), P'' = prod_d exp(i (om N/2 - s dk - s (k0' - 1))): the per-offset phase
// exp(i s (j+1)) of the reference coefficient (helper.py:148-162) is carried by the modulated grid.
__global__ void k_col_keys(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om, long long M,
                           int nq2, int* __restrict__ keys, int* __restrict__ vals) {
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= M) return;
    int ks[3];
    for (int d = 0; d < 3; ++d) {
        double q;
        ks[d] = wrap_index(offset_k0(om[m * 3 + d], pc->gam[d], g.J[d], &q) + 1, g.K[d]);
    }
    keys[m] = ((ks[1] / COL_T1) * nq2 + ks[2] / COL_T2) * g.K[0] + ks[0];
    vals[m] = (int)m;
}

__global__ void k_col_records(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om,
                              const int* __restrict__ perm, long long M, int nq2, const int* __restrict__ cbin,
                              float* __restrict__ rec, float4* __restrict__ side) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int m = perm[i];
    float* out = rec + i * COL_RECW;
    int* outi = reinterpret_cast<int*>(out);
    double ph = 0.0;
    int ks[3];
    for (int w = 0; w < COL_RECW; ++w) out[w] = 0.f;
    for (int d = 0; d < 3; ++d) {
        DimResult R;
        const double o = om[(long long)m * 3 + d];
        dim_math(o, d, g, pc, R);
        const int k = wrap_index(R.k0 + 1, g.K[d]);
        ks[d] = k;
        const float sgd = ((g.N[d] - 1) & 1) ? -1.f : 1.f;
        for (int j = 0; j < 6; ++j) {
            const float cj = (k + j >= g.K[d]) ? sgd * (float)R.c[j] : (float)R.c[j];
            if (d == 0) {
                out[12 + j] = cj;                      // plane k + j
            } else if (d == 1) {
                const int b = k % COL_T1 + j;          // box row 0..8
                out[4 * (b % 3) + b / 3] = cj;
            } else {
                out[20 + k % COL_T2 + j] = cj;         // box column 0..9
            }
        }
        const double s = pc->gam[d] * ((double)g.N[d] - 1.0) / 2.0;
        ph += o * (double)g.N[d] / 2.0 - s * R.dk - s * (double)(k - 1);
    }
    const int key = ((ks[1] / COL_T1) * nq2 + ks[2] / COL_T2) * g.K[0] + ks[0];
    outi[18] = ks[0];
    outi[19] = cbin[key + 1] - (int)i;
    double sn, cs;
    sincos(ph, &sn, &cs);
    side[i] = make_float4((float)cs, (float)sn, __int_as_float(m), 0.f);
}

// ---- 2-D multi-coil row sweep (sweep2d.cu): key = (strip, first row), 20-word records in sweep order ----
// A strip is SW2_CT first-neighbour columns (dim 1) wide; its box is 8 columns.
// record words: [c0[0..3] | c0[4] c0[5] p0 run | c1t[0..3] | c1t[4..7] | P''.re P''.im m 0]
//   c0[j]: weight of row p0 + j;  c1t[b] = c1[b - k1rel] (zero outside the footprint): box columns
//   run: samples from this one to the end of its (strip, first row) bin
//   weights of neighbours that wrap around the periodic grid carry (-1)^(N_d - 1) (phase-modulated grid)
//   P'' = prod_d exp(i (om N/2 - s dk - s (k0' - 1))), m = original sample index
__global__ void k_sw2_keys(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om, long long M,
                           int* __restrict__ keys, int* __restrict__ vals) {
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= M) return;
    int ks[2];
    for (int d = 0; d < 2; ++d) {
        double q;
        ks[d] = wrap_index(offset_k0(om[m * 2 + d], pc->gam[d], g.J[d], &q) + 1, g.K[d]);
    }
    keys[m] = (ks[1] / SW2_CT) * g.K[0] + ks[0];
    vals[m] = (int)m;
}

__global__ void k_sw2_records(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om,
                              const int* __restrict__ perm, long long M, const int* __restrict__ cbin,
                              float* __restrict__ rec) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int m = perm[i];
    float* out = rec + i * SW2_RECW;
    int* outi = reinterpret_cast<int*>(out);
    for (int w = 0; w < SW2_RECW; ++w) out[w] = 0.f;
    double ph = 0.0;
    int ks[2];
    for (int d = 0; d < 2; ++d) {
        DimResult R;
        const double o = om[(long long)m * 2 + d];
        dim_math(o, d, g, pc, R);
        const int k = wrap_index(R.k0 + 1, g.K[d]);
        ks[d] = k;
        const float sgd = ((g.N[d] - 1) & 1) ? -1.f : 1.f;
        for (int j = 0; j < 6; ++j) {
            const float cj = (k + j >= g.K[d]) ? sgd * (float)R.c[j] : (float)R.c[j];
            if (d == 0) out[j] = cj;                       // words 0..5
            else out[8 + k % SW2_CT + j] = cj;             // box column 0..7
        }
        const double s = pc->gam[d] * ((double)g.N[d] - 1.0) / 2.0;
        ph += o * (double)g.N[d] / 2.0 - s * R.dk - s * (double)(k - 1);
    }
    const int key = (ks[1] / SW2_CT) * g.K[0] + ks[0];
    outi[6] = ks[0];
    outi[7] = cbin[key + 1] - (int)i;
    double sn, cs;
    sincos(ph, &sn, &cs);
    out[16] = (float)cs;
    out[17] = (float)sn;
    outi[18] = m;
}

__global__ void k_bin_start(const int* __restrict__ sorted_keys, long long M, int nbins,
                             int* __restrict__ bin_start) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nbins) return;
    // lower_bound(sorted_keys, t)
    long long lo = 0, hi = M;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < t) lo = mid + 1; else hi = mid;
    }
    bin_start[t] = (int)lo;
}

// parity exports, original order, reference pELL layout (helper.py:346-393)
__global__ void k_export(Geom g, const PlanConst* __restrict__ pc, const double* __restrict__ om,
                         long long M, uint32_t* __restrict__ kindx, float2* __restrict__ udata,
                         int* __restrict__ k0out) {
    long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (int d = 0; d < g.ndim; ++d) {
        double o = om[m * g.ndim + d];
        if (udata) {
            DimResult R;
            dim_math(o, d, g, pc, R);
            double s = pc->gam[d] * ((double)g.N[d] - 1.0) / 2.0;
            for (int j = 0; j < g.J[d]; ++j) {
                double arg = -(double)(j + 1) + R.dk;
                double ph = o * (double)g.N[d] / 2.0 - s * arg;
                double sn, cs;
                sincos(ph, &sn, &cs);
                udata[m * g.sumJ + g.Joff[d] + j] = make_float2((float)(R.c[j] * cs), (float)(R.c[j] * sn));
            }
        }
        double q;
        long long k0 = offset_k0(o, pc->gam[d], g.J[d], &q);
        if (k0out) k0out[m * g.ndim + d] = (int)k0;
        if (kindx) {
            for (int j = 0; j < g.J[d]; ++j) {
                long long idx = wrap_index(k0 + j + 1, g.K[d]);
                kindx[m * g.sumJ + g.Joff[d] + j] = (uint32_t)(idx * g.Kstride[d]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// plan-time temporaries: freed on every exit path
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() { cudaFree(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

static void choose_tiles(Geom& g) {
    // tile edge per dim: the bin key of the sort contract.  3-D: 16^3 (staged box 21^3 for J=6),
    // 2-D: 32^2, 1-D: 256; never larger than K.
    // 3-D tiles are split into 8^3 sub-tiles (the gridding kernel's private boxes).
    int want = g.ndim == 3 ? 16 : (g.ndim == 2 ? 32 : 256);
    g.nsubprod = 1;
    for (int d = 0; d < g.ndim; ++d) {
        g.tile[d] = std::min(want, g.K[d]);
        g.ntile[d] = (g.K[d] + g.tile[d] - 1) / g.tile[d];
        g.sub[d] = g.tile[d];
        if (g.ndim == 3 && g.tile[d] % 8 == 0) g.sub[d] = 8;
        // 2-D: 8 x 16 sub-tiles (13 x 21 boxes of the batch-innermost multi-coil kernels, batch2d.cu)
        if (g.ndim == 2 && d == 0 && g.tile[d] % 8 == 0) g.sub[d] = 8;
        if (g.ndim == 2 && d == 1 && g.tile[d] % 16 == 0) g.sub[d] = 16;
        g.nsub[d] = g.tile[d] / g.sub[d];
        g.nsubprod *= g.nsub[d];
    }
}

extern "C" int b200nufft_plan_create(b200nufft_plan_t* out, int device, int ndim,
                                     const int32_t* Nd, const int32_t* Kd, const int32_t* Jd,
                                     int64_t M, const double* om, int batch,
                                     const double* alpha, const int32_t* alpha_len,
                                     const double* Tmat, const float* sn, void* stream) {
    ARG_CHECK(out != nullptr, "out is NULL");
    ARG_CHECK(ndim >= 1 && ndim <= MAXD, "ndim must be 1..3");
    ARG_CHECK(M >= 0 && M < (1LL << 31), "M out of range");
    ARG_CHECK(batch >= 1, "batch must be >= 1");
    ARG_CHECK(om != nullptr || M == 0, "om is NULL");
    ARG_CHECK(alpha && alpha_len && Tmat && sn, "plan constants are NULL");
    ON_DEVICE(device);
    cudaStream_t st = as_stream(stream);

    b200nufft_plan_s* p = new b200nufft_plan_s();
    p->device = device;
    p->M = M;
    p->batch = batch;
    Geom& g = p->g;
    memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    g.Nprod = 1;
    g.Kprod = 1;
    g.prodJ = 1;
    int snsum = 0;
    for (int d = 0; d < ndim; ++d) {
        if (!(Nd[d] >= 1 && Kd[d] >= Nd[d] && Jd[d] >= 1 && Jd[d] <= MAXJ && Jd[d] <= Kd[d])) {
            delete p;
            b200_set_error("invalid argument: need 1 <= Nd <= Kd, 1 <= Jd <= min(16, Kd)");
            return B200_ERR_ARG;
        }
        if (!(alpha_len[d] >= 1 && alpha_len[d] <= MAXL)) {
            delete p;
            b200_set_error("invalid argument: alpha_len out of range");
            return B200_ERR_ARG;
        }
        g.N[d] = Nd[d];
        g.K[d] = Kd[d];
        g.J[d] = Jd[d];
        g.Joff[d] = g.sumJ;
        g.sumJ += Jd[d];
        g.prodJ *= Jd[d];
        g.snoff[d] = snsum;
        snsum += Nd[d];
        g.Nprod *= Nd[d];
        g.Kprod *= Kd[d];
    }
    if (g.Kprod >= (1LL << 31)) {
        delete p;
        b200_set_error("invalid argument: prod(Kd) must be < 2^31");
        return B200_ERR_ARG;
    }
    for (int d = ndim - 1; d >= 0; --d) {
        g.Kstride[d] = (d == ndim - 1) ? 1 : g.Kstride[d + 1] * g.K[d + 1];
    }
    g.recw = ((g.sumJ + 2 + ndim + 1) + 3) & ~3;
    choose_tiles(g);
    p->n_tiles = 1;
    for (int d = 0; d < ndim; ++d) p->n_tiles *= g.ntile[d];
    p->n_bins = p->n_tiles * g.nsubprod;
    {
        const char* env = getenv("B200NUFFT_LAYOUT");
        const bool want_col = g_layout_pref.load() == 1 && !(env && strcmp(env, "tile") == 0);
        p->has_col = want_col && col3d_supported(g);
    }

    PlanConst pc;
    memset(&pc, 0, sizeof(pc));
    for (int d = 0; d < ndim; ++d) {
        pc.gam[d] = 2.0 * M_PI / ((double)g.K[d] * 1.0);
        pc.ratio[d] = 1.0 * g.K[d] / g.N[d];
        pc.L[d] = alpha_len[d] - 1;
        for (int l = 0; l < alpha_len[d]; ++l) pc.alpha[d][l] = alpha[d * MAXL + l];
        for (int i = 0; i < g.J[d] * g.J[d]; ++i) pc.T[d][i] = Tmat[d * MAXJ * MAXJ + i];
        double s = pc.gam[d] * ((double)g.N[d] - 1.0) / 2.0;
        for (int j = 0; j < g.J[d]; ++j)
            g.E[d][j] = make_float2((float)cos(s * (j + 1)), (float)sin(s * (j + 1)));
        if (d == 0) {
            for (int t = 0; t < 24; ++t) g.F0[t] = make_float2((float)cos(s * (t + 1)), (float)sin(s * (t + 1)));
            for (int t = 0; t < 16; ++t) g.G0[t] = make_float2((float)cos(s * t), (float)-sin(s * t));
        }
        if (d == ndim - 1) {
            for (int t = 0; t < 40; ++t) g.Fl[t] = make_float2((float)cos(s * (t + 1)), (float)sin(s * (t + 1)));
            for (int t = 0; t < 32; ++t) g.Gl[t] = make_float2((float)cos(s * t), (float)-sin(s * t));
        }
    }

#define PLAN_TRY(expr)                                                      \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) {                                            \
            b200_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
            b200nufft_plan_destroy(p);                                      \
            return B200_ERR_CUDA;                                           \
        }                                                                   \
    } while (0)

    const long long Mal = std::max<long long>(M, 1);
    PLAN_TRY(cudaMalloc(&p->d_pc, sizeof(PlanConst)));
    PLAN_TRY(cudaMemcpyAsync(p->d_pc, &pc, sizeof(pc), cudaMemcpyHostToDevice, st));
    PLAN_TRY(cudaMalloc(&p->d_sn, sizeof(float) * snsum));
    PLAN_TRY(cudaMemcpyAsync(p->d_sn, sn, sizeof(float) * snsum, cudaMemcpyDefault, st));
    PLAN_TRY(cudaMalloc(&p->d_om, sizeof(double) * Mal * ndim));
    if (M > 0) PLAN_TRY(cudaMemcpyAsync(p->d_om, om, sizeof(double) * M * ndim, cudaMemcpyDefault, st));
    PLAN_TRY(cudaMalloc(&p->d_perm, sizeof(int) * Mal));
    PLAN_TRY(cudaMalloc(&p->d_rec, sizeof(float) * Mal * g.recw));
    PLAN_TRY(cudaMalloc(&p->d_bin_start, sizeof(int) * (p->n_bins + 1)));
    p->bytes = sizeof(PlanConst) + sizeof(float) * snsum + sizeof(double) * Mal * ndim +
               sizeof(int) * Mal + sizeof(float) * Mal * g.recw + sizeof(int) * (p->n_bins + 1);

    std::vector<int> h_bin_start(p->n_bins + 1, 0);
    if (M > 0) {
        DevTmp t_keys, t_keys_s, t_vals, t_tmp;
        PLAN_TRY(cudaMalloc(&t_keys.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_keys_s.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_vals.p, sizeof(int) * M));
        int *d_keys = t_keys.as<int>(), *d_keys_s = t_keys_s.as<int>(), *d_vals = t_vals.as<int>();
        const int TB = 256;
        const unsigned nblk = (unsigned)((M + TB - 1) / TB);
        k_bin_keys<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, M, d_keys, d_vals);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        int end_bit = 1;
        while ((1LL << end_bit) < p->n_bins) ++end_bit;
        size_t tmp_bytes = 0;
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_perm,
                                                 (int)M, 0, end_bit, st));
        PLAN_TRY(cudaMalloc(&t_tmp.p, tmp_bytes));
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(t_tmp.p, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_perm,
                                                 (int)M, 0, end_bit, st));   // stable LSD radix sort
        k_bin_start<<<(p->n_bins + 1 + TB - 1) / TB, TB, 0, st>>>(d_keys_s, M, p->n_bins, p->d_bin_start);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        if (tiled_supported(g)) {
            PLAN_TRY(cudaMalloc(&p->d_trec, sizeof(float) * Mal * TILE_RECW));
            p->bytes += sizeof(float) * Mal * TILE_RECW;
        }
        k_build_records<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, p->d_perm, M, p->d_rec, p->d_trec);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        PLAN_TRY(cudaMemcpyAsync(h_bin_start.data(), p->d_bin_start, sizeof(int) * (p->n_bins + 1),
                                 cudaMemcpyDeviceToHost, st));
        PLAN_TRY(cudaStreamSynchronize(st));
    } else {
        PLAN_TRY(cudaMemsetAsync(p->d_bin_start, 0, sizeof(int) * (p->n_bins + 1), st));
        PLAN_TRY(cudaStreamSynchronize(st));
    }

    // work list for the tiled kernels: non-empty tiles, split into chunks, heaviest first
    {
        // interp: a tile is cut along dim 0 into nsub[0] slabs (sub-tiles with the same s0 are
        // contiguous in the sort order); WorkItem.pad = first plane of the slab inside the tile.
        const int CHUNK = 1024;
        std::vector<WorkItem> work;
        const int slabs = g.nsub[0];
        const int per_slab = g.nsubprod / slabs;
        for (int t = 0; t < p->n_tiles; ++t) {
            for (int h = 0; h < slabs; ++h) {
                int b = h_bin_start[t * g.nsubprod + h * per_slab], e = h_bin_start[t * g.nsubprod + (h + 1) * per_slab];
                for (int s = b; s < e; s += CHUNK)
                    work.push_back(WorkItem{t, s, std::min(e, s + CHUNK), h * g.sub[0]});
            }
        }
        // Items keep their spatial (bin) order, so that boxes whose halos overlap are processed close in time and
        // hit in L2; only items clearly heavier than average (k-space centre of radial scans) are moved to the
        // front, largest first.
        auto order_items = [](std::vector<WorkItem>& v) {
            if (v.empty()) return;
            long long tot = 0;
            for (const WorkItem& w : v) tot += w.end - w.begin;
            const long long heavy = (3 * (tot / (long long)v.size() + 1)) / 2;       // 1.5 x the average item
            // heavy items first, largest first (LPT); the rest stays in spatial order (stable)
            std::stable_sort(v.begin(), v.end(), [heavy](const WorkItem& a, const WorkItem& b) {
                const long long na = a.end - a.begin, nb = b.end - b.begin;
                const bool ha = na >= heavy, hb = nb >= heavy;
                if (ha != hb) return ha;
                return ha && na > nb;
            });
        };
        order_items(work);
        p->n_work = (int)work.size();
        if (p->n_work > 0) {
            PLAN_TRY(cudaMalloc(&p->d_work, sizeof(WorkItem) * work.size()));
            PLAN_TRY(cudaMemcpyAsync(p->d_work, work.data(), sizeof(WorkItem) * work.size(),
                                     cudaMemcpyHostToDevice, st));
            PLAN_TRY(cudaStreamSynchronize(st));
            p->bytes += sizeof(WorkItem) * work.size();
        }
        // gridding: one item per non-empty sub-tile bin (chunks of GCHUNK), heaviest first
        const int GCHUNK = g.ndim == 2 ? 256 : 4096;      // 2-D: short serial chains per warp (dense k-space centres)
        std::vector<WorkItem> gwork;
        for (int b = 0; b < p->n_bins; ++b) {
            int s0 = h_bin_start[b], e = h_bin_start[b + 1];
            for (int s = s0; s < e; s += GCHUNK) gwork.push_back(WorkItem{b, s, std::min(e, s + GCHUNK), 0});
        }
        order_items(gwork);
        p->n_gwork = (int)gwork.size();
        if (p->n_gwork > 0) {
            PLAN_TRY(cudaMalloc(&p->d_gwork, sizeof(WorkItem) * gwork.size()));
            PLAN_TRY(cudaMemcpyAsync(p->d_gwork, gwork.data(), sizeof(WorkItem) * gwork.size(),
                                     cudaMemcpyHostToDevice, st));
            PLAN_TRY(cudaStreamSynchronize(st));
            p->bytes += sizeof(WorkItem) * gwork.size();
        }
    }

    // ---- column-sweep gridding: second sort by (column, first plane), records, work items ----
    if (p->has_col && M > 0) {
        const int col_nq1 = (g.K[1] + COL_T1 - 1) / COL_T1, col_nq2 = (g.K[2] + COL_T2 - 1) / COL_T2;
        const int col_ncol = col_nq1 * col_nq2;
        const int n_cbins = col_ncol * g.K[0];               // bins = (column, first plane)
        PLAN_TRY(cudaMalloc(&p->d_cperm, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&p->d_crec, sizeof(float) * M * COL_RECW));
        PLAN_TRY(cudaMalloc(&p->d_cside, sizeof(float4) * M));
        p->bytes += sizeof(int) * M + sizeof(float) * M * COL_RECW + sizeof(float4) * M;
        {   // modulation tables m_d[g] = exp(i s_d g), s_d = gam_d (N_d - 1) / 2
            std::vector<float2> hm(g.K[0] + g.K[1] + g.K[2]);
            int o = 0;
            for (int d = 0; d < 3; ++d) {
                const double sd = pc.gam[d] * ((double)g.N[d] - 1.0) / 2.0;
                for (int t = 0; t < g.K[d]; ++t) hm[o + t] = make_float2((float)cos(sd * t), (float)sin(sd * t));
                o += g.K[d];
            }
            PLAN_TRY(cudaMalloc(&p->d_mod, sizeof(float2) * hm.size()));
            PLAN_TRY(cudaMemcpyAsync(p->d_mod, hm.data(), sizeof(float2) * hm.size(), cudaMemcpyHostToDevice, st));
            PLAN_TRY(cudaStreamSynchronize(st));
        }
        DevTmp t_keys, t_keys_s, t_vals, t_cbin, t_tmp;
        PLAN_TRY(cudaMalloc(&t_keys.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_keys_s.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_vals.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_cbin.p, sizeof(int) * (n_cbins + 1)));
        int *d_keys = t_keys.as<int>(), *d_keys_s = t_keys_s.as<int>(), *d_vals = t_vals.as<int>(), *d_cbin = t_cbin.as<int>();
        const int TB = 256;
        const unsigned nblk = (unsigned)((M + TB - 1) / TB);
        k_col_keys<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, M, col_nq2, d_keys, d_vals);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        int end_bit = 1;
        while ((1LL << end_bit) < n_cbins) ++end_bit;
        size_t tmp_bytes = 0;
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_cperm, (int)M, 0,
                                                 end_bit, st));
        PLAN_TRY(cudaMalloc(&t_tmp.p, tmp_bytes));
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(t_tmp.p, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_cperm, (int)M, 0,
                                                 end_bit, st));
        k_bin_start<<<(n_cbins + 1 + TB - 1) / TB, TB, 0, st>>>(d_keys_s, M, n_cbins, d_cbin);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        k_col_records<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, p->d_cperm, M, col_nq2, d_cbin, p->d_crec, p->d_cside);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        std::vector<int> h_cbin(n_cbins + 1, 0);
        PLAN_TRY(cudaMemcpyAsync(h_cbin.data(), d_cbin, sizeof(int) * (n_cbins + 1), cudaMemcpyDeviceToHost, st));
        PLAN_TRY(cudaStreamSynchronize(st));
        // every column's samples (already in plane order) are cut into segments of at most COL_SEG samples; items
        // stay in column order (q2 fastest), so that columns whose halos overlap run close in time
        std::vector<WorkItem> cw;
        int col_seg = COL_SEG;
        if (const char* e = getenv("B200NUFFT_COL_SEG")) col_seg = std::max(16, atoi(e));    // tuning knob
        const long long tail_from = M - M / 8;        // the last eighth of the samples is cut four times finer
        for (int col = 0; col < col_ncol; ++col) {
            const int b = h_cbin[(size_t)col * g.K[0]], e = h_cbin[(size_t)(col + 1) * g.K[0]];
            const int n = e - b;
            if (n <= 0) continue;
            const int seg = (long long)b >= tail_from ? col_seg / 4 : col_seg;
            const int nseg = (n + seg - 1) / seg;
            for (int sgi = 0; sgi < nseg; ++sgi) {
                const int sb = b + (int)((long long)n * sgi / nseg), se = b + (int)((long long)n * (sgi + 1) / nseg);
                if (se > sb) cw.push_back(WorkItem{col, sb, se, 0});
            }
        }
        p->n_cwork = (int)cw.size();
        PLAN_TRY(cudaMalloc(&p->d_cwork, sizeof(WorkItem) * cw.size()));
        PLAN_TRY(cudaMemcpyAsync(p->d_cwork, cw.data(), sizeof(WorkItem) * cw.size(), cudaMemcpyHostToDevice, st));
        PLAN_TRY(cudaStreamSynchronize(st));
        p->bytes += sizeof(WorkItem) * cw.size();
    } else {
        p->has_col = false;
    }
    // ---- 2-D multi-coil row sweep: sort by (strip, first row), records, work items, modulation tables ----
    p->has_sw2 = sweep2d_supported(g) && M > 0;
    if (p->has_sw2) {
        const int nq = (g.K[1] + SW2_CT - 1) / SW2_CT;
        const int n_sbins = nq * g.K[0];
        PLAN_TRY(cudaMalloc(&p->d_sw_perm, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&p->d_sw_rec, sizeof(float) * M * SW2_RECW));
        p->bytes += sizeof(int) * M + sizeof(float) * M * SW2_RECW;
        {   // modulation tables m_d[g] = exp(i s_d g), concatenated [K0 | K1]
            std::vector<float2> hm(g.K[0] + g.K[1]);
            int o = 0;
            for (int d = 0; d < 2; ++d) {
                const double sd = pc.gam[d] * ((double)g.N[d] - 1.0) / 2.0;
                for (int t = 0; t < g.K[d]; ++t) hm[o + t] = make_float2((float)cos(sd * t), (float)sin(sd * t));
                o += g.K[d];
            }
            PLAN_TRY(cudaMalloc(&p->d_mod, sizeof(float2) * hm.size()));
            PLAN_TRY(cudaMemcpyAsync(p->d_mod, hm.data(), sizeof(float2) * hm.size(), cudaMemcpyHostToDevice, st));
            PLAN_TRY(cudaStreamSynchronize(st));
        }
        DevTmp t_keys, t_keys_s, t_vals, t_cbin, t_tmp;
        PLAN_TRY(cudaMalloc(&t_keys.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_keys_s.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_vals.p, sizeof(int) * M));
        PLAN_TRY(cudaMalloc(&t_cbin.p, sizeof(int) * (n_sbins + 1)));
        int *d_keys = t_keys.as<int>(), *d_keys_s = t_keys_s.as<int>(), *d_vals = t_vals.as<int>(), *d_cbin = t_cbin.as<int>();
        const int TB = 256;
        const unsigned nblk = (unsigned)((M + TB - 1) / TB);
        k_sw2_keys<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, M, d_keys, d_vals);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        int end_bit = 1;
        while ((1LL << end_bit) < n_sbins) ++end_bit;
        size_t tmp_bytes = 0;
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_sw_perm, (int)M, 0,
                                                 end_bit, st));
        PLAN_TRY(cudaMalloc(&t_tmp.p, tmp_bytes));
        PLAN_TRY(cub::DeviceRadixSort::SortPairs(t_tmp.p, tmp_bytes, d_keys, d_keys_s, d_vals, p->d_sw_perm, (int)M, 0,
                                                 end_bit, st));
        k_bin_start<<<(n_sbins + 1 + TB - 1) / TB, TB, 0, st>>>(d_keys_s, M, n_sbins, d_cbin);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        k_sw2_records<<<nblk, TB, 0, st>>>(g, p->d_pc, p->d_om, p->d_sw_perm, M, d_cbin, p->d_sw_rec);
        g_launches++;
        PLAN_TRY(cudaGetLastError());
        std::vector<int> h_cbin(n_sbins + 1, 0);
        PLAN_TRY(cudaMemcpyAsync(h_cbin.data(), d_cbin, sizeof(int) * (n_sbins + 1), cudaMemcpyDeviceToHost, st));
        PLAN_TRY(cudaStreamSynchronize(st));
        // every strip's samples (already in row order) are cut into segments of at most SW2_SEG samples, heaviest
        // strips first (radial trajectories are dense at the centre of k-space)
        std::vector<WorkItem> sw;
        int seg = SW2_SEG;
        if (const char* e = getenv("B200NUFFT_SW2_SEG")) seg = std::max(16, atoi(e));    // tuning knob
        for (int q = 0; q < nq; ++q) {
            const int b = h_cbin[(size_t)q * g.K[0]], e = h_cbin[(size_t)(q + 1) * g.K[0]];
            const int n = e - b;
            if (n <= 0) continue;
            const int nseg = (n + seg - 1) / seg;
            for (int sgi = 0; sgi < nseg; ++sgi) {
                const int sb = b + (int)((long long)n * sgi / nseg), se = b + (int)((long long)n * (sgi + 1) / nseg);
                if (se > sb) sw.push_back(WorkItem{q, sb, se, 0});
            }
        }
        p->n_sw_work = (int)sw.size();
        PLAN_TRY(cudaMalloc(&p->d_sw_work, sizeof(WorkItem) * sw.size()));
        PLAN_TRY(cudaMemcpyAsync(p->d_sw_work, sw.data(), sizeof(WorkItem) * sw.size(), cudaMemcpyHostToDevice, st));
        PLAN_TRY(cudaStreamSynchronize(st));
        p->bytes += sizeof(WorkItem) * sw.size();
    }
#undef PLAN_TRY
    *out = p;
    return B200_OK;
}

void host_pipe_destroy(b200nufft_plan_t p);   // stages.cu

extern "C" int b200nufft_plan_destroy(b200nufft_plan_t p) {
    if (!p) return B200_OK;
    DeviceGuard dg(p->device);
    host_pipe_destroy(p);
    cudaFree(p->d_pc);
    cudaFree(p->d_om);
    cudaFree(p->d_perm);
    cudaFree(p->d_rec);
    cudaFree(p->d_trec);
    cudaFree(p->d_sn);
    cudaFree(p->d_bin_start);
    cudaFree(p->d_work);
    cudaFree(p->d_gwork);
    cudaFree(p->d_cperm);
    cudaFree(p->d_crec);
    cudaFree(p->d_cside);
    cudaFree(p->d_cwork);
    cudaFree(p->d_sw_perm);
    cudaFree(p->d_sw_rec);
    cudaFree(p->d_sw_work);
    cudaFree(p->d_mod);
    cudaFree(p->d_ccount);
    cudaFree(p->d_ys);
    cudaFree(p->d_ys2);
    cudaFree(p->d_ysb);
    cudaFree(p->d_tw256);
    cudaFree(p->d_xc);
    cudaFree(p->d_grid);
    cudaFree(p->d_grid2);
    if (p->zs) {
        cudaStreamSynchronize(p->zs);
        cudaEventDestroy(p->gridz_ev);
        cudaEventDestroy(p->gridz_use);
        cudaStreamDestroy(p->zs);
    }
    cudaFree(p->d_gridz);
    cudaFree(p->d_fbi_tw);
    cudaFree(p->d_fbi_rev);
    cudaFree(p->d_xin);
    cudaFree(p->d_yio);
    if (p->fft_valid) cufftDestroy(p->fft);
    if (p->fftp_valid) { cufftDestroy(p->fft2d); cufftDestroy(p->fft1d); }
    if (p->fft_bi_valid) cufftDestroy(p->fft_bi);
    delete p;
    return B200_OK;
}

extern "C" int b200nufft_plan_get_layout(b200nufft_plan_t p) { return p ? ((p->has_col || p->has_sw2) ? 1 : 0) : -1; }
// 1 if grids handed to / returned by the stage entry points for a call with nb coils are BATCH-INNERMOST
// (k[K0][K1][nb], the reference's Kd + (batch,) order), 0 if they are coil-major (nb contiguous Kd grids)
extern "C" int b200nufft_grid_layout(b200nufft_plan_t p, int nb) { return (p && use_bi(p, nb)) ? 1 : 0; }

static int run_export(b200nufft_plan_t p, uint32_t* kindx, float2* udata, int* k0, void* stream) {
    ARG_CHECK(p != nullptr, "plan is NULL");
    if (p->M == 0) return B200_OK;
    ON_DEVICE(p->device);
    const int TB = 128;
    k_export<<<(unsigned)((p->M + TB - 1) / TB), TB, 0, as_stream(stream)>>>(p->g, p->d_pc, p->d_om, p->M,
                                                                            kindx, udata, k0);
    LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200nufft_plan_get_kindx(b200nufft_plan_t p, uint32_t* kindx, void* stream) {
    return run_export(p, kindx, nullptr, nullptr, stream);
}
extern "C" int b200nufft_plan_get_udata(b200nufft_plan_t p, b200_c64* udata, void* stream) {
    return run_export(p, nullptr, reinterpret_cast<float2*>(udata), nullptr, stream);
}
extern "C" int b200nufft_plan_get_k0(b200nufft_plan_t p, int32_t* k0, void* stream) {
    return run_export(p, nullptr, nullptr, k0, stream);
}
extern "C" int b200nufft_plan_get_perm(b200nufft_plan_t p, int32_t* perm, void* stream) {
    ARG_CHECK(p != nullptr, "plan is NULL");
    if (p->M == 0) return B200_OK;
    CUDA_TRY(cudaMemcpyAsync(perm, p->d_perm, sizeof(int) * p->M, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return B200_OK;
}
extern "C" int b200nufft_plan_get_tile(b200nufft_plan_t p, int32_t* tile_host) {
    ARG_CHECK(p != nullptr, "plan is NULL");
    for (int d = 0; d < p->g.ndim; ++d) {
        tile_host[d] = p->g.tile[d];
        tile_host[p->g.ndim + d] = p->g.sub[d];
    }
    return B200_OK;
}
// sweep-order permutation of the column-sweep gridding records: stable argsort of
// key = (q1 * nq2 + q2) * K0 + first plane, i.e. the generic bin key with tile = (K0, T1, T2), sub-tile = (1, T1, T2)
extern "C" int b200nufft_plan_get_col_perm(b200nufft_plan_t p, int32_t* perm, int32_t* tile_host, void* stream) {
    ARG_CHECK(p != nullptr, "plan is NULL");
    ARG_CHECK(p->has_col || p->has_sw2, "plan has no sweep records");
    if (p->has_sw2) {               // 2-D: tile = (K0, 3), sub-tile = (1, 3) in the first four entries
        if (tile_host) {
            const int t[6] = {p->g.K[0], SW2_CT, 1, SW2_CT, 0, 0};
            for (int i = 0; i < 6; ++i) tile_host[i] = t[i];
        }
        if (perm && p->M > 0)
            CUDA_TRY(cudaMemcpyAsync(perm, p->d_sw_perm, sizeof(int) * p->M, cudaMemcpyDeviceToDevice, as_stream(stream)));
        return B200_OK;
    }
    if (tile_host) {
        const int t[6] = {p->g.K[0], COL_T1, COL_T2, 1, COL_T1, COL_T2};
        for (int i = 0; i < 6; ++i) tile_host[i] = t[i];
    }
    if (perm && p->M > 0)
        CUDA_TRY(cudaMemcpyAsync(perm, p->d_cperm, sizeof(int) * p->M, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return B200_OK;
}
extern "C" int64_t b200nufft_plan_bytes(b200nufft_plan_t p) { return p ? p->bytes : 0; }

extern "C" int b200nufft_set_variant(b200nufft_plan_t p, int iv, int gv) {
    ARG_CHECK(p != nullptr, "plan is NULL");
    ARG_CHECK(iv >= 0 && iv <= 3 && gv >= 0 && gv <= 2, "interp variant must be 0..3, gridding variant 0..2");
    if (iv == 3 && !(p->has_col && p->d_mod && col3d_interp_supported(p->g))) {
        b200_set_error("interp variant 3 (column-sweep gather) needs a 3-D J = 6 plan with sweep records and K0 >= 10");
        return B200_ERR_UNSUPPORTED;
    }
    p->interp_variant = iv;
    p->gridding_variant = gv;
    p->fft_variant = (iv == 1 && gv == 1) ? 1 : 0;   // 'generic' also selects the cuFFT path
    return B200_OK;
}
