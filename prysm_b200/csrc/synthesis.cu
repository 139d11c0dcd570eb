// Pupil synthesis on the device -- the step before the path: coordinate grids, circular
// masks (binary / one-sample grey edge), Jacobi polynomials by three-term recurrence, Zernike sequences and
// coefficient-weighted Zernike sums.  All streaming, HBM-bound on their outputs; the recurrences run in fp64
// registers whatever the storage precision (the reference runs them in config precision, prysm/polynomials/
// jacobi.py:147-175, zernike.py:74-181) and are driven by a plan passed in kernel-parameter space, so a launch
// reads no table from memory and is CUDA-graph capturable.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pb {

static inline int grid_for(long long n, int threads, int sm_count) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)sm_count * 16;
    return (int)std::max<long long>(1, std::min(g, cap));
}

#define PB_GRID_STRIDE(i, n) \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__device__ inline float mul_rn(float a, float b) { return __fmul_rn(a, b); }       // one rounding per operation, like numpy
__device__ inline double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// x = (ix - nx/2)*dx, y = (iy - ny/2)*dx in working precision (prysm/coordinates.py:344-378); r, t from them
template <typename R>
__global__ void xy_grid_kernel(int ny, int nx, double dx, R* __restrict__ x, R* __restrict__ y, R* __restrict__ r, R* __restrict__ t) {
    const long long n = (long long)ny * nx;
    PB_GRID_STRIDE(i, n) {
        const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
        const R xv = mul_rn((R)(ix - nx / 2), (R)dx), yv = mul_rn((R)(iy - ny / 2), (R)dx);
        if (x) x[i] = xv;
        if (y) y[i] = yv;
        if (r) r[i] = (R)hypot((double)xv, (double)yv);
        if (t) t[i] = (R)atan2((double)yv, (double)xv);
    }
}

template <typename R>
__global__ void cart_to_polar_kernel(const R* __restrict__ x, const R* __restrict__ y, long long n, R* __restrict__ r, R* __restrict__ t) {
    PB_GRID_STRIDE(i, n) {
        const double xv = x[i], yv = y[i];
        if (r) r[i] = (R)hypot(xv, yv);
        if (t) t[i] = (R)atan2(yv, xv);
    }
}

// aa_dx <= 0: out(u8) = (r - radius <= 0)            (geometry.circle, prysm/geometry.py:337-372)
// aa_dx  > 0: out(R)  = clip(0.5 - (r - radius)/aa_dx, 0, 1)   (antialias(circle_sdf), geometry.py:11-34)
template <typename R>
__global__ void circle_kernel(const R* __restrict__ r, long long n, R radius, R aa_dx, void* out) {
    PB_GRID_STRIDE(i, n) {
        const R d = r[i] - radius;
        if (aa_dx > R(0)) {
            const R c = R(0.5) - d / aa_dx;
            reinterpret_cast<R*>(out)[i] = c < R(0) ? R(0) : (c > R(1) ? R(1) : c);
        } else {
            reinterpret_cast<unsigned char*>(out)[i] = d <= R(0) ? 1 : 0;
        }
    }
}

#define PB_JAC_MAX 120
struct JacobiPlan {
    double alpha, beta;
    int nmax;
    short slot[PB_JAC_MAX + 1];       // output index of order j, or -1
    double abc[3 * PB_JAC_MAX];       // (A, B, C) of order k-1 for the step that makes P_k, k = 2..nmax at [3*(k-2)]
};

template <typename R>
__global__ void jacobi_seq_kernel(const R* __restrict__ x, long long n, const __grid_constant__ JacobiPlan pl, R* __restrict__ out) {
    PB_GRID_STRIDE(i, n) {
        const double xv = x[i];
        double Pm2 = 1.0, Pm1 = pl.alpha + 1.0 + (pl.alpha + pl.beta + 2.0) * ((xv - 1.0) / 2.0);
        if (pl.slot[0] >= 0) out[(long long)pl.slot[0] * n + i] = (R)Pm2;
        if (pl.nmax >= 1 && pl.slot[1] >= 0) out[(long long)pl.slot[1] * n + i] = (R)Pm1;
        for (int k = 2; k <= pl.nmax; ++k) {
            const double A = pl.abc[3 * (k - 2)], B = pl.abc[3 * (k - 2) + 1], C = pl.abc[3 * (k - 2) + 2];
            const double Pn = (A * xv + B) * Pm1 - C * Pm2;
            Pm2 = Pm1; Pm1 = Pn;
            if (pl.slot[k] >= 0) out[(long long)pl.slot[k] * n + i] = (R)Pn;
        }
    }
}

// Zernike plan: one group per |m| (ascending); per Jacobi order j of a group one step holding the recurrence
// coefficients that make P_j and the weights of the (at most one) cosine and sine mode of that (|m|, j).
// m = 0 modes are "cosine" modes of the |m| = 0 group (cos 0 = r^0 = 1).  Every index is warp-uniform, so the
// plan is read through the constant bank with uniform loads.
#define PB_ZERN_GROUPS 24
#define PB_ZERN_STEPS 72
struct ZernGroup { short am, njmax, first, pad; };
struct ZernStep { double A, B, C, wc, ws; };     // P_j = (A x + B) P_{j-1} - C P_{j-2};  value = P_j r^|m| (wc cos + ws sin)
struct ZernPlan {
    int ngroups;
    ZernGroup g[PB_ZERN_GROUPS];
    short kc[PB_ZERN_STEPS], ks[PB_ZERN_STEPS];  // output slots of the cosine / sine mode of a step (MODE 0), -1 = absent
    ZernStep s[PB_ZERN_STEPS];
};

// MODE 0: out[k][i] = w * Z_k   (zernike_nm_seq);  MODE 1: out[i] (+)= sum w * Z_k  (zernike_sum, w = c*norm)
// POLAR: inputs are (r, t); otherwise Cartesian (x, y) with r = hypot, cos t = x/r, sin t = y/r
template <typename R, int MODE, bool POLAR>
__global__ void zernike_kernel(const R* __restrict__ a, const R* __restrict__ b, long long n, const __grid_constant__ ZernPlan pl,
                               int accumulate, R* __restrict__ out) {
    PB_GRID_STRIDE(i, n) {
        double r, c1, s1;
        if (POLAR) {
            r = a[i];
            sincos((double)b[i], &s1, &c1);
        } else {
            const double xv = a[i], yv = b[i];
            r = hypot(xv, yv);
            c1 = r > 0.0 ? xv / r : 1.0;
            s1 = r > 0.0 ? yv / r : 0.0;
        }
        const double x2 = 2.0 * (r * r) - 1.0, hx = (x2 - 1.0) / 2.0;
        double rc = 1.0, rs = 0.0;                // r^am cos(am t), r^am sin(am t), advanced group to group
        const double rc1 = r * c1, rs1 = r * s1;
        int am_now = 0;
        double acc = 0.0;
        for (int gi = 0; gi < pl.ngroups; ++gi) {
            const ZernGroup g = pl.g[gi];
            for (; am_now < g.am; ++am_now) {     // (r e^{it})^am by repeated complex multiplication
                const double t0 = rc * rc1 - rs * rs1;
                rs = rs * rc1 + rc * rs1;
                rc = t0;
            }
            double Pm2 = 1.0, Pm1 = 1.0 + ((double)g.am + 2.0) * hx, P = 1.0;
            for (int j = 0; j <= g.njmax; ++j) {
                const int si = g.first + j;
                if (j == 1) P = Pm1;
                else if (j >= 2) {
                    P = (pl.s[si].A * x2 + pl.s[si].B) * Pm1 - pl.s[si].C * Pm2;
                    Pm2 = Pm1; Pm1 = P;
                }
                if (MODE == 1) {
                    acc += P * (pl.s[si].wc * rc + pl.s[si].ws * rs);
                } else {
                    const int kc = pl.kc[si], ks = pl.ks[si];
                    if (kc >= 0) out[(long long)kc * n + i] = (R)(pl.s[si].wc * P * rc);
                    if (ks >= 0) out[(long long)ks * n + i] = (R)(pl.s[si].ws * P * rs);
                }
            }
        }
        if (MODE == 1) out[i] = accumulate ? (R)((double)out[i] + acc) : (R)acc;
    }
}

// DLMF 18.9 coefficients as the reference writes them (prysm/polynomials/jacobi.py:13-39)
static void recurrence_abc(int n, double alpha, double beta, double* abc) {
    const double apb = alpha + beta;
    if (n == 0 && (apb == 0.0 || apb == -1.0)) {
        abc[0] = 0.5 * (alpha + beta) + 1.0; abc[1] = 0.5 * (alpha - beta); abc[2] = 1.0;
        return;
    }
    abc[0] = ((2 * n + alpha + beta + 1) * (2 * n + alpha + beta + 2)) / (2 * (n + 1) * (n + alpha + beta + 1));
    abc[1] = ((alpha * alpha - beta * beta) * (2 * n + alpha + beta + 1)) / (2 * (n + 1) * (n + alpha + beta + 1) * (2 * n + alpha + beta));
    abc[2] = ((n + alpha) * (n + beta) * (2 * n + alpha + beta + 2)) / ((n + 1) * (n + alpha + beta + 1) * (2 * n + alpha + beta));
}

struct HostTerm { int am, nj, trig, k; double scale; };   // trig 1: cosine (and m = 0), 2: sine

// Fills a plan with as many whole |m| groups as fit; duplicate (n, m) pairs beyond the first stay in `terms` for a
// later plan (a step holds one cosine and one sine weight).
static int next_plan(std::vector<HostTerm>& terms, ZernPlan& pl) {
    pl.ngroups = 0;
    int nsteps = 0;
    std::vector<HostTerm> rest;
    size_t pos = 0;
    bool full = false;
    while (pos < terms.size()) {
        size_t end = pos;
        int njmax = 0;
        while (end < terms.size() && terms[end].am == terms[pos].am) { njmax = std::max(njmax, terms[end].nj); ++end; }
        if (njmax + 1 > PB_ZERN_STEPS) return PB_ERR_UNSUPPORTED;
        if (full || pl.ngroups == PB_ZERN_GROUPS || nsteps + njmax + 1 > PB_ZERN_STEPS) {
            full = true;
            rest.insert(rest.end(), terms.begin() + pos, terms.begin() + end);
            pos = end;
            continue;
        }
        ZernGroup& g = pl.g[pl.ngroups++];
        g.am = (short)terms[pos].am; g.njmax = (short)njmax; g.first = (short)nsteps; g.pad = 0;
        for (int j = 0; j <= njmax; ++j) {
            ZernStep& st = pl.s[nsteps + j];
            double abc[3] = {0.0, 0.0, 0.0};
            if (j >= 2) recurrence_abc(j - 1, 0.0, (double)g.am, abc);
            st.A = abc[0]; st.B = abc[1]; st.C = abc[2]; st.wc = 0.0; st.ws = 0.0;
            pl.kc[nsteps + j] = -1; pl.ks[nsteps + j] = -1;
        }
        for (size_t q = pos; q < end; ++q) {
            const HostTerm& t = terms[q];
            const int si = nsteps + t.nj;
            short& slot = t.trig == 2 ? pl.ks[si] : pl.kc[si];
            if (slot >= 0) { rest.push_back(t); continue; }       // a repeated (n, m): next plan
            slot = (short)t.k;
            (t.trig == 2 ? pl.s[si].ws : pl.s[si].wc) = t.scale;
        }
        nsteps += njmax + 1;
        pos = end;
    }
    terms.swap(rest);
    return PB_OK;
}

static int build_terms(Handle* h, int k, const int* n_host, const int* m_host, const double* coefs, int norm,
                       std::vector<HostTerm>& terms) {
    terms.clear();
    for (int q = 0; q < k; ++q) {
        const int n = n_host[q], m = m_host[q], am = m < 0 ? -m : m;
        if (n < 0 || am > n || ((n - am) & 1)) return fail(h, PB_ERR_INVALID, "zernike (n, m): need n >= |m| and n - |m| even");
        if (coefs && coefs[q] == 0.0) continue;                              // zernike.py:178-179 skips zero weights
        double sc = coefs ? coefs[q] : 1.0;
        if (norm) sc *= std::sqrt((2.0 * (n + 1)) / (m == 0 ? 2.0 : 1.0));   // zernike_norm, zernike.py:25-27
        terms.push_back({am, (n - am) / 2, m < 0 ? 2 : 1, q, sc});
    }
    std::stable_sort(terms.begin(), terms.end(), [](const HostTerm& a, const HostTerm& b) {
        return a.am != b.am ? a.am < b.am : a.nj < b.nj;
    });
    return PB_OK;
}

}  // namespace pb

using namespace pb;

#define PB_HANDLE(hh)                                   \
    PB_ENTER(hh);                      \
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128"); \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int pb_xy_grid(pb_handle_t hh, int dtype, int ny, int nx, double dx, void* x, void* y, void* r, void* t, void* stream) {
    PB_HANDLE(hh);
    if (ny < 1 || nx < 1) return fail(h, PB_ERR_INVALID, "bad grid shape");
    const int g = grid_for((long long)ny * nx, 256, h->sm_count);
    if (dtype == PB_C64) xy_grid_kernel<float><<<g, 256, 0, st>>>(ny, nx, dx, (float*)x, (float*)y, (float*)r, (float*)t);
    else xy_grid_kernel<double><<<g, 256, 0, st>>>(ny, nx, dx, (double*)x, (double*)y, (double*)r, (double*)t);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_cart_to_polar(pb_handle_t hh, int dtype, const void* x, const void* y, long long count, void* r, void* t,
                                void* stream) {
    PB_HANDLE(hh);
    if (!x || !y) return fail(h, PB_ERR_INVALID, "null coordinates");
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64) cart_to_polar_kernel<float><<<g, 256, 0, st>>>((const float*)x, (const float*)y, count, (float*)r, (float*)t);
    else cart_to_polar_kernel<double><<<g, 256, 0, st>>>((const double*)x, (const double*)y, count, (double*)r, (double*)t);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_circle(pb_handle_t hh, int dtype, const void* r, long long count, double radius, double aa_dx, void* out,
                         void* stream) {
    PB_HANDLE(hh);
    if (!r || !out) return fail(h, PB_ERR_INVALID, "null pointer");
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64) circle_kernel<float><<<g, 256, 0, st>>>((const float*)r, count, (float)radius, (float)aa_dx, out);
    else circle_kernel<double><<<g, 256, 0, st>>>((const double*)r, count, radius, aa_dx, out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_jacobi_seq(pb_handle_t hh, int dtype, const void* x, long long count, int nmax, double alpha, double beta,
                             const int* slots_host, void* out, void* stream) {
    PB_HANDLE(hh);
    if (!x || !out || !slots_host || nmax < 0) return fail(h, PB_ERR_INVALID, "bad jacobi arguments");
    if (nmax > PB_JAC_MAX) return fail(h, PB_ERR_UNSUPPORTED, "jacobi order above 120");
    if (count <= 0) return PB_OK;
    JacobiPlan pl;
    pl.alpha = alpha; pl.beta = beta; pl.nmax = nmax;
    for (int j = 0; j <= PB_JAC_MAX; ++j) pl.slot[j] = j <= nmax ? (short)slots_host[j] : (short)-1;
    for (int k = 2; k <= nmax; ++k) recurrence_abc(k - 1, alpha, beta, pl.abc + 3 * (k - 2));
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64) jacobi_seq_kernel<float><<<g, 256, 0, st>>>((const float*)x, count, pl, (float*)out);
    else jacobi_seq_kernel<double><<<g, 256, 0, st>>>((const double*)x, count, pl, (double*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

template <int MODE>
static int run_zernike(Handle* h, int dtype, int polar, const void* a, const void* b, long long count,
                       std::vector<HostTerm>& terms, void* out, cudaStream_t st) {
    const int g = grid_for(count, 256, h->sm_count);
    int launches = 0;
    if (MODE == 1 && terms.empty()) {      // all weights zero: the sum is zero
        PB_CUDA(h, cudaMemsetAsync(out, 0, (size_t)count * (dtype == PB_C64 ? 4 : 8), st));
        return PB_OK;
    }
    while (!terms.empty()) {
        ZernPlan pl;
        PB_TRY(next_plan(terms, pl));
        const int acc = launches > 0;
        if (dtype == PB_C64) {
            if (polar) zernike_kernel<float, MODE, true><<<g, 256, 0, st>>>((const float*)a, (const float*)b, count, pl, acc, (float*)out);
            else zernike_kernel<float, MODE, false><<<g, 256, 0, st>>>((const float*)a, (const float*)b, count, pl, acc, (float*)out);
        } else {
            if (polar) zernike_kernel<double, MODE, true><<<g, 256, 0, st>>>((const double*)a, (const double*)b, count, pl, acc, (double*)out);
            else zernike_kernel<double, MODE, false><<<g, 256, 0, st>>>((const double*)a, (const double*)b, count, pl, acc, (double*)out);
        }
        PB_LAUNCH_CHECK(h);
        ++launches;
    }
    return PB_OK;
}

extern "C" int pb_zernike_seq(pb_handle_t hh, int dtype, int polar, const void* a, const void* b, long long count, int k,
                              const int* n_host, const int* m_host, int norm, void* out, void* stream) {
    PB_HANDLE(hh);
    if (!a || !b || !out || k < 1 || !n_host || !m_host) return fail(h, PB_ERR_INVALID, "bad zernike arguments");
    if (k > 32767) return fail(h, PB_ERR_UNSUPPORTED, "more than 32767 modes");
    if (count <= 0) return PB_OK;
    std::vector<HostTerm> terms;
    PB_TRY(build_terms(h, k, n_host, m_host, nullptr, norm, terms));
    return run_zernike<0>(h, dtype, polar, a, b, count, terms, out, st);
}

extern "C" int pb_zernike_sum(pb_handle_t hh, int dtype, int polar, const void* a, const void* b, long long count, int k,
                              const int* n_host, const int* m_host, const double* coefs_host, int norm, void* out,
                              void* stream) {
    PB_HANDLE(hh);
    if (!a || !b || !out || k < 0 || (k > 0 && (!n_host || !m_host || !coefs_host))) return fail(h, PB_ERR_INVALID, "bad zernike arguments");
    if (count <= 0) return PB_OK;
    std::vector<HostTerm> terms;
    PB_TRY(build_terms(h, k, n_host, m_host, coefs_host, norm, terms));
    return run_zernike<1>(h, dtype, polar, a, b, count, terms, out, st);
}
