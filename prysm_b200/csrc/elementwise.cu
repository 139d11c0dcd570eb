// Streaming elementwise kernels and small reductions on the path (all HBM-bound):
// phase-screen synthesis, |.|^2 (+ weighted accumulate), separable multiplies, weighted mode
// sum, OTF normalisation, first moments, angular-spectrum factor vectors, MDFT bases.
#include "common.cuh"
#include "czt.cuh"

namespace pb {

static inline int grid_for(long long n, int threads, int sm_count) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)sm_count * 16;  // grid-stride beyond a few waves
    return (int)std::max<long long>(1, std::min(g, cap));
}

template <typename R>
__global__ void phase_screen_kernel(const void* amp, int amp_kind, const R* __restrict__ opd, double kturns,
                                    long long n, cplx<R>* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        R a = R(1);
        if (amp_kind == PB_AMP_REAL) a = reinterpret_cast<const R*>(amp)[i];
        else if (amp_kind == PB_AMP_U8) a = reinterpret_cast<const unsigned char*>(amp)[i] ? R(1) : R(0);
        cplx<R> e = mk<R>(R(0), R(0));
        if (a != R(0)) {
            e = expi_turns(kturns * (double)opd[i], R(0));
            e.x *= a; e.y *= a;
        }
        out[i] = e;
    }
}

template <typename R>
__global__ void intensity_kernel(const cplx<R>* __restrict__ in, long long n, R weight, int accumulate,
                                 R* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cplx<R> v = in[i];
        R I = v.x * v.x + v.y * v.y;
        if (accumulate) out[i] += weight * I; else out[i] = weight * I;
    }
}

template <typename R>
__global__ void mul_outer_kernel(const cplx<R>* __restrict__ in, long long in_ld, int ny, int nx,
                                 const cplx<R>* __restrict__ vy, int conj_y, const cplx<R>* __restrict__ vx,
                                 int conj_x, R scale, cplx<R>* __restrict__ out, long long out_ld) {
    const long long n = (long long)ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        cplx<R> v = in[y * in_ld + x];
        if (vy) { cplx<R> m = vy[y]; v = conj_y ? cmulc(v, m) : cmul(v, m); }
        if (vx) { cplx<R> m = vx[x]; v = conj_x ? cmulc(v, m) : cmul(v, m); }
        v.x *= scale; v.y *= scale;
        out[y * out_ld + x] = v;
    }
}

// out = a (op) b, or with b == nullptr a (op) scalar; reverse swaps the operands.  op: 0 mul, 1 div, 2 add, 3 sub
template <typename R>
__global__ void binary_kernel(int op, const cplx<R>* __restrict__ a, const cplx<R>* __restrict__ b, cplx<R> sc,
                              int reverse, long long n, cplx<R>* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cplx<R> x = a[i], y = b ? b[i] : sc;
        if (reverse) { cplx<R> t = x; x = y; y = t; }
        cplx<R> r;
        if (op == 0) r = cmul(x, y);
        else if (op == 1) { const R d = y.x * y.x + y.y * y.y; r = cmulc(x, y); r.x /= d; r.y /= d; }
        else if (op == 2) r = cadd(x, y);
        else r = csub(x, y);
        out[i] = r;
    }
}

#define PB_MAX_MODES 64
template <typename R> struct ModeWeights { R w[PB_MAX_MODES]; };

template <typename R>
__global__ void weighted_sum_kernel(const R* __restrict__ modes, int k, long long n, ModeWeights<R> w, int accumulate,
                                    R* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        R acc = accumulate ? out[i] : R(0);
        for (int m = 0; m < k; ++m) acc += w.w[m] * modes[(long long)m * n + i];
        out[i] = acc;
    }
}

template <typename R>
__global__ void otf_normalize_kernel(const cplx<R>* __restrict__ D, long long n, long long centre, int which,
                                     R* __restrict__ mtf, R* __restrict__ ptf, cplx<R>* __restrict__ otf) {
    const cplx<R> c = D[centre];
    const R den = c.x * c.x + c.y * c.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cplx<R> v = D[i];
        cplx<R> q = cmulc(v, c);  // v * conj(c)
        q.x /= den; q.y /= den;
        if (which & 1) mtf[i] = hypot(q.x, q.y);
        if (which & 2) ptf[i] = atan2(q.y, q.x);
        if (which & 4) otf[i] = q;
    }
}

template <typename R>
__global__ void moments_kernel(const R* __restrict__ d, int ny, int nx, double* __restrict__ sums) {
    double s0 = 0, sy = 0, sx = 0;
    const long long n = (long long)ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        const double v = (double)d[i];
        s0 += v; sy += v * y; sx += v * x;
    }
    for (int o = 16; o; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        sy += __shfl_down_sync(0xffffffffu, sy, o);
        sx += __shfl_down_sync(0xffffffffu, sx, o);
    }
    __shared__ double sh[3][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = s0; sh[1][wid] = sy; sh[2][wid] = sx; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        s0 = lane < nw ? sh[0][lane] : 0; sy = lane < nw ? sh[1][lane] : 0; sx = lane < nw ? sh[2][lane] : 0;
        for (int o = 16; o; o >>= 1) {
            s0 += __shfl_down_sync(0xffffffffu, s0, o);
            sy += __shfl_down_sync(0xffffffffu, sy, o);
            sx += __shfl_down_sync(0xffffffffu, sx, o);
        }
        if (lane == 0) { atomicAdd(&sums[0], s0); atomicAdd(&sums[1], sy); atomicAdd(&sums[2], sx); }
    }
}

// Baliga & Cohn encircled energy: r * sum( MTF * J1(2 pi r nu) / nu ) * dnu^2 on the fftrange*df grid, nu(0) -> 1e-16
// (prysm/otf.py:319-414).  One block-reduction per radius (blockIdx.y), fp64 Bessel and accumulation.
template <typename R>
__global__ void encircled_energy_kernel(const R* __restrict__ mtf, int ny, int nx, double df, const double* __restrict__ radii,
                                        double* __restrict__ out) {
    const double r = radii[blockIdx.y];
    const long long n = (long long)ny * nx;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        const double fy = (double)(y - ny / 2) * df, fx = (double)(x - nx / 2) * df;
        double nu = hypot(fx, fy);
        if (nu == 0.0) nu = 1e-16;
        acc += (double)mtf[i] * j1(6.283185307179586476925 * r * nu) / nu;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicAdd(&out[blockIdx.y], acc * r * df * df);
    }
}

// Bluestein pieces of one CZT axis built on the device from scalars (prysm/fttools.py:372-389, 277-281):
//   b[n]    = exp(sign*i*pi*alpha*n^2),            n = -(N/2) .. ;      N entries
//   post[m] = exp(sign*i*pi*alpha*(m+shift)^2) * exp(sign*2*pi*i*xc*f[m]),  f[m] = f0 + m*df;   M entries
//   h[d]    = exp(-sign*i*pi*alpha*(d+shift)^2),   d = m0-n_last .. m_last-n0, zero-extended to K entries
// all phases in fp64 turns, reduced before the sincos.
template <typename R>
__global__ void czt_plan_kernel(int N, int M, int K, double shift, double alpha, int sign, double xc, double f0, double df,
                                cplx<R>* __restrict__ b, cplx<R>* __restrict__ post, cplx<R>* __restrict__ hk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double half = 0.5 * sign * alpha;
    if (i < N) {
        const double n = (double)(i - N / 2);
        b[i] = expi_turns(half * n * n, R(0));
    }
    if (i < M) {
        const double q = (double)(i - M / 2) + shift;
        post[i] = expi_turns(half * q * q + sign * xc * (f0 + i * df), R(0));
    }
    if (i < K) {
        const int nd = N + M - 1;
        cplx<R> v = mk<R>(R(0), R(0));
        if (i < nd) {
            const double d = (double)(i + (-(M / 2)) - (-(N / 2) + N - 1)) + shift;   // m0 - n_last + i + shift
            v = expi_turns(-half * d * d, R(0));
        }
        hk[i] = v;
    }
}

// Hadj[k] = conj(H[k]) * exp(-2*pi*i*(N-1)*k/K): the adjoint's embedding at offset N-1 as a spectral phase
template <typename R>
__global__ void czt_hadj_kernel(int N, int K, const cplx<R>* __restrict__ H, cplx<R>* __restrict__ Hadj) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const long long r = ((long long)(N - 1) * k) % K;
    const cplx<R> ph = expi_turns(-(double)r / (double)K, R(0));
    Hadj[k] = cmul(cconj(H[k]), ph);
}

// exp(-i*pi*wvl_mm*z*k^2), k = fftfreq(n, dx).  The phase is formed in fp64 for both dtypes
// (the reference rounds k to float32 first when precision=32, costing it ~1e-4 rad at C5 sizes).
template <typename R>
__global__ void as_vector_kernel(int n, double wvl_mm, double dx, double z, cplx<R>* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int fi = i < (n + 1) / 2 ? i : i - n;      // numpy fftfreq ordering
    const double kf = (double)fi / ((double)n * dx);
    const double k2 = kf * kf;                        // kept in fp64: the complex64 path then tracks the fp64 reference
    out[i] = expi_turns(-0.5 * wvl_mm * z * k2, R(0));  // -pi*w*z*k^2 rad = -(w*z*k^2)/2 turns
}

template <typename R>
__global__ void mdft_basis_kernel(const double* __restrict__ f, int m, const double* __restrict__ x, int n, int sign,
                                  cplx<R>* __restrict__ E) {
    const long long tot = (long long)m * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / n), l = (int)(i - (long long)j * n);
        E[i] = expi_turns(sign * f[j] * x[l], R(0));
    }
}

}  // namespace pb

using namespace pb;

#define PB_HANDLE(hh)                                   \
    PB_ENTER(hh);                      \
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128"); \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int pb_phase_screen(pb_handle_t hh, int dtype, const void* amp, int amp_kind, const void* opd,
                               double kscale, long long count, void* out, void* stream) {
    PB_HANDLE(hh);
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    const double kt = kscale / (2.0 * M_PI);
    if (dtype == PB_C64) phase_screen_kernel<float><<<g, 256, 0, st>>>(amp, amp_kind, (const float*)opd, kt, count, (float2*)out);
    else phase_screen_kernel<double><<<g, 256, 0, st>>>(amp, amp_kind, (const double*)opd, kt, count, (double2*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_intensity(pb_handle_t hh, int dtype, const void* in, long long count, double weight, int accumulate,
                            void* out, void* stream) {
    PB_HANDLE(hh);
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64) intensity_kernel<float><<<g, 256, 0, st>>>((const float2*)in, count, (float)weight, accumulate, (float*)out);
    else intensity_kernel<double><<<g, 256, 0, st>>>((const double2*)in, count, weight, accumulate, (double*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_binary(pb_handle_t hh, int dtype, int op, const void* a, const void* b, double s_re, double s_im,
                         int reverse, long long count, void* out, void* stream) {
    PB_HANDLE(hh);
    if (op < 0 || op > 3 || !a || !out) return fail(h, PB_ERR_INVALID, "bad binary-op arguments");
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64)
        binary_kernel<float><<<g, 256, 0, st>>>(op, (const float2*)a, (const float2*)b, make_float2((float)s_re, (float)s_im), reverse, count, (float2*)out);
    else
        binary_kernel<double><<<g, 256, 0, st>>>(op, (const double2*)a, (const double2*)b, make_double2(s_re, s_im), reverse, count, (double2*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_mul_outer(pb_handle_t hh, int dtype, const void* in, long long in_ld, int ny, int nx, const void* vy,
                            int conj_y, const void* vx, int conj_x, double scale, void* out, long long out_ld,
                            void* stream) {
    PB_HANDLE(hh);
    if (ny <= 0 || nx <= 0) return PB_OK;
    const int g = grid_for((long long)ny * nx, 256, h->sm_count);
    if (dtype == PB_C64)
        mul_outer_kernel<float><<<g, 256, 0, st>>>((const float2*)in, in_ld, ny, nx, (const float2*)vy, conj_y,
                                                   (const float2*)vx, conj_x, (float)scale, (float2*)out, out_ld);
    else
        mul_outer_kernel<double><<<g, 256, 0, st>>>((const double2*)in, in_ld, ny, nx, (const double2*)vy, conj_y,
                                                    (const double2*)vx, conj_x, scale, (double2*)out, out_ld);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_weighted_sum(pb_handle_t hh, int dtype, const void* modes, int k, long long count,
                               const double* weights_host, void* out, void* stream) {
    PB_HANDLE(hh);
    if (k < 1 || count <= 0 || !weights_host) return fail(h, PB_ERR_INVALID, "bad mode count / weights");
    const int g = grid_for(count, 256, h->sm_count);
    for (int k0 = 0; k0 < k; k0 += PB_MAX_MODES) {
        const int kk = std::min(PB_MAX_MODES, k - k0);
        if (dtype == PB_C64) {
            ModeWeights<float> w;
            for (int i = 0; i < kk; ++i) w.w[i] = (float)weights_host[k0 + i];
            weighted_sum_kernel<float><<<g, 256, 0, st>>>((const float*)modes + (long long)k0 * count, kk, count, w, k0 > 0, (float*)out);
        } else {
            ModeWeights<double> w;
            for (int i = 0; i < kk; ++i) w.w[i] = weights_host[k0 + i];
            weighted_sum_kernel<double><<<g, 256, 0, st>>>((const double*)modes + (long long)k0 * count, kk, count, w, k0 > 0, (double*)out);
        }
        PB_LAUNCH_CHECK(h);
    }
    return PB_OK;
}

extern "C" int pb_otf_normalize(pb_handle_t hh, int dtype, const void* D, int ny, int nx, int which, void* mtf,
                                void* ptf, void* otf, void* stream) {
    PB_HANDLE(hh);
    if (ny < 1 || nx < 1) return fail(h, PB_ERR_INVALID, "bad shape");
    const long long n = (long long)ny * nx, centre = (long long)(ny / 2) * nx + nx / 2;
    const int g = grid_for(n, 256, h->sm_count);
    if (dtype == PB_C64)
        otf_normalize_kernel<float><<<g, 256, 0, st>>>((const float2*)D, n, centre, which, (float*)mtf, (float*)ptf, (float2*)otf);
    else
        otf_normalize_kernel<double><<<g, 256, 0, st>>>((const double2*)D, n, centre, which, (double*)mtf, (double*)ptf, (double2*)otf);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_moments(pb_handle_t hh, int dtype, const void* data, int ny, int nx, double* sums_host,
                          void* stream) {
    PB_HANDLE(hh);
    if (ny < 1 || nx < 1 || !sums_host) return fail(h, PB_ERR_INVALID, "bad shape");
    void* d = nullptr;
    PB_TRY(ensure_scratch(h, 2, 3 * sizeof(double), &d));
    PB_CUDA(h, cudaMemsetAsync(d, 0, 3 * sizeof(double), st));
    const int g = grid_for((long long)ny * nx, 256, h->sm_count);
    if (dtype == PB_C64) moments_kernel<float><<<g, 256, 0, st>>>((const float*)data, ny, nx, (double*)d);
    else moments_kernel<double><<<g, 256, 0, st>>>((const double*)data, ny, nx, (double*)d);
    PB_LAUNCH_CHECK(h);
    PB_CUDA(h, cudaMemcpyAsync(sums_host, d, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    PB_CUDA(h, cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_encircled_energy(pb_handle_t hh, int dtype, const void* mtf, int ny, int nx, double df,
                                   const double* radii_mm_host, int nr, double* out_host, void* stream) {
    PB_HANDLE(hh);
    if (ny < 1 || nx < 1 || nr < 1 || !mtf || !radii_mm_host || !out_host) return fail(h, PB_ERR_INVALID, "bad encircled-energy arguments");
    void* d = nullptr;
    PB_TRY(ensure_scratch(h, 2, (size_t)2 * nr * sizeof(double), &d));
    double* dr = (double*)d;
    double* dout = dr + nr;
    PB_CUDA(h, cudaMemcpyAsync(dr, radii_mm_host, nr * sizeof(double), cudaMemcpyHostToDevice, st));
    PB_CUDA(h, cudaMemsetAsync(dout, 0, nr * sizeof(double), st));
    const int gx = (int)std::min<long long>(((long long)ny * nx + 255) / 256, (long long)h->sm_count * 4);
    dim3 g(gx, nr);
    if (dtype == PB_C64) encircled_energy_kernel<float><<<g, 256, 0, st>>>((const float*)mtf, ny, nx, df, dr, dout);
    else encircled_energy_kernel<double><<<g, 256, 0, st>>>((const double*)mtf, ny, nx, df, dr, dout);
    PB_LAUNCH_CHECK(h);
    PB_CUDA(h, cudaMemcpyAsync(out_host, dout, nr * sizeof(double), cudaMemcpyDeviceToHost, st));
    PB_CUDA(h, cudaStreamSynchronize(st));
    return PB_OK;
}

int pb::czt_plan_impl(Handle* h, pb_handle_t hh, int dtype, int N, int M, int K, double shift, double alpha, int sign, double xc,
                      double f0, double df, void* b, void* post, void* H, void* Hadj, void* hk, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n = std::max(K, std::max(N, M));
    if (dtype == PB_C64) czt_plan_kernel<float><<<(n + 255) / 256, 256, 0, st>>>(N, M, K, shift, alpha, sign, xc, f0, df, (float2*)b, (float2*)post, (float2*)hk);
    else czt_plan_kernel<double><<<(n + 255) / 256, 256, 0, st>>>(N, M, K, shift, alpha, sign, xc, f0, df, (double2*)b, (double2*)post, (double2*)hk);
    PB_LAUNCH_CHECK(h);
    PB_TRY(pb_fft1(hh, dtype, hk, 1, K, K, 1, K, -1, 1.0, H, K, stream));   // H = FFT_K(h)
    if (Hadj) {
        if (dtype == PB_C64) czt_hadj_kernel<float><<<(K + 255) / 256, 256, 0, st>>>(N, K, (const float2*)H, (float2*)Hadj);
        else czt_hadj_kernel<double><<<(K + 255) / 256, 256, 0, st>>>(N, K, (const double2*)H, (double2*)Hadj);
        PB_LAUNCH_CHECK(h);
    }
    return PB_OK;
}

extern "C" int pb_czt_plan(pb_handle_t hh, int dtype, int N, int M, int K, double shift, double alpha, int sign, double xc,
                           double f0, double df, void* b, void* post, void* H, void* Hadj, void* stream) {
    PB_HANDLE(hh);
    (void)st;
    if (N < 1 || M < 1 || K < N + M - 1 || (sign != 1 && sign != -1) || !b || !post || !H || !Hadj)
        return fail(h, PB_ERR_INVALID, "bad czt plan arguments");
    void* hk = nullptr;
    PB_TRY(ensure_scratch(h, 2, (size_t)K * csize(dtype), &hk));
    return czt_plan_impl(h, hh, dtype, N, M, K, shift, alpha, sign, xc, f0, df, b, post, H, Hadj, hk, stream);
}

extern "C" int pb_angular_spectrum_vectors(pb_handle_t hh, int dtype, int ky, int kx, double wvl_um, double dx_mm,
                                           double z_mm, void* ty, void* tx, void* stream) {
    PB_HANDLE(hh);
    if (ky < 1 || kx < 1) return fail(h, PB_ERR_INVALID, "bad shape");
    const double w = wvl_um / 1e3;
    if (dtype == PB_C64) {
        as_vector_kernel<float><<<(ky + 255) / 256, 256, 0, st>>>(ky, w, dx_mm, z_mm, (float2*)ty);
        as_vector_kernel<float><<<(kx + 255) / 256, 256, 0, st>>>(kx, w, dx_mm, z_mm, (float2*)tx);
    } else {
        as_vector_kernel<double><<<(ky + 255) / 256, 256, 0, st>>>(ky, w, dx_mm, z_mm, (double2*)ty);
        as_vector_kernel<double><<<(kx + 255) / 256, 256, 0, st>>>(kx, w, dx_mm, z_mm, (double2*)tx);
    }
    PB_LAUNCH_CHECK(h);
    h->launches++;
    return PB_OK;
}

extern "C" int pb_mdft_basis(pb_handle_t hh, int dtype, const double* f_host, int m, const double* x_host, int n,
                             int sign, void* E, void* stream) {
    PB_HANDLE(hh);
    if (m < 1 || n < 1 || !f_host || !x_host || (sign != 1 && sign != -1)) return fail(h, PB_ERR_INVALID, "bad basis arguments");
    void* d = nullptr;
    PB_TRY(ensure_scratch(h, 2, (size_t)(m + n) * sizeof(double), &d));
    double* df = (double*)d;
    double* dxv = df + m;
    // pageable host memory: the copies return after staging, so the caller's arrays may be freed
    PB_CUDA(h, cudaMemcpyAsync(df, f_host, m * sizeof(double), cudaMemcpyHostToDevice, st));
    PB_CUDA(h, cudaMemcpyAsync(dxv, x_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    const int g = grid_for((long long)m * n, 256, h->sm_count);
    if (dtype == PB_C64) mdft_basis_kernel<float><<<g, 256, 0, st>>>(df, m, dxv, n, sign, (float2*)E);
    else mdft_basis_kernel<double><<<g, 256, 0, st>>>(df, m, dxv, n, sign, (double2*)E);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}
