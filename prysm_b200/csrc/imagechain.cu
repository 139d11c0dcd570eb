// Image-chain consumers of the path's FFTs (prysm/convolution.py:9-32): the two elementwise
// kernels that let a real-object x real-PSF convolution run as ONE forward and ONE inverse transform.
//   z = obj + i*psf  ->  Z = FFT2(z) = O + i*H with O, H Hermitian, so conj(Z[-k]) = O[k] - i*H[k] and
//   O[k]*H[k] = (Z[k]^2 - conj(Z[-k])^2) / (4i).
// Both stream; the product reads each spectrum sample twice (k and -k) through L2.
// Rounding: the transform's error is relative to |Z|, so a PSF of unit sum packed beside an object of sum 1e5 would
// lose log2(1e5) bits.  pb_balance_scale measures s = sqrt(sum obj^2 / sum psf^2) on the device; the pack kernel
// multiplies the imaginary part by s and the product kernel divides it back out -- no host round trip.
#include "common.cuh"

namespace pb {

static inline int grid_for(long long n, int threads, int sm_count) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)sm_count * 16;
    return (int)std::max<long long>(1, std::min(g, cap));
}

template <typename R>
__global__ void pack_complex_kernel(const R* __restrict__ re, const R* __restrict__ im, const double* __restrict__ im_scale,
                                    long long n, cplx<R>* __restrict__ out) {
    const R s = im_scale ? (R)*im_scale : R(1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = mk<R>(re[i], im ? s * im[i] : R(0));
}

// sums[0] += sum a^2, sums[1] += sum b^2
template <typename R>
__global__ void norms2_kernel(const R* __restrict__ a, const R* __restrict__ b, long long n, double* __restrict__ sums) {
    double sa = 0, sb = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = a[i], y = b[i];
        sa += x * x; sb += y * y;
    }
    for (int o = 16; o; o >>= 1) { sa += __shfl_down_sync(0xffffffffu, sa, o); sb += __shfl_down_sync(0xffffffffu, sb, o); }
    __shared__ double sh[2][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = sa; sh[1][wid] = sb; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        sa = lane < nw ? sh[0][lane] : 0.0; sb = lane < nw ? sh[1][lane] : 0.0;
        for (int o = 16; o; o >>= 1) { sa += __shfl_down_sync(0xffffffffu, sa, o); sb += __shfl_down_sync(0xffffffffu, sb, o); }
        if (lane == 0) { atomicAdd(&sums[0], sa); atomicAdd(&sums[1], sb); }
    }
}

__global__ void balance_finalize_kernel(const double* __restrict__ sums, double* __restrict__ s) {
    const double a = sums[0], b = sums[1];
    *s = (a > 0.0 && b > 0.0 && isfinite(a) && isfinite(b)) ? sqrt(a / b) : 1.0;
}

template <typename R>
__global__ void packed_spectrum_product_kernel(const cplx<R>* __restrict__ Z, int ny, int nx, R scale,
                                               const double* __restrict__ im_scale, cplx<R>* __restrict__ out) {
    const long long n = (long long)ny * nx;
    if (im_scale) scale = (R)((double)scale / *im_scale);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ky = (int)(i / nx), kx = (int)(i - (long long)ky * nx);
        const int my = ky ? ny - ky : 0, mx = kx ? nx - kx : 0;
        const cplx<R> a = Z[i], b = Z[(long long)my * nx + mx];
        // a^2 - conj(b)^2
        const R pre = (a.x * a.x - a.y * a.y) - (b.x * b.x - b.y * b.y);
        const R pim = 2 * a.x * a.y + 2 * b.x * b.y;
        // divide by 4i: (p)/(4i) = (pim - i*pre)/4
        out[i] = mk<R>(scale * R(0.25) * pim, -scale * R(0.25) * pre);
    }
}

// detector sampling (prysm/detector.py:151-338): block mean / sum over fy x fx samples, its adjoint (repeat), and the
// analytic pixel / OLPF transfer functions on separable frequency vectors
template <typename R>
__global__ void bindown_kernel(const R* __restrict__ in, int oy, int ox, int fy, int fx, R scale, R* __restrict__ out) {
    const long long n = (long long)oy * ox;
    const long long in_ld = (long long)ox * fx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / ox), x = (int)(i - (long long)y * ox);
        const R* p = in + (long long)y * fy * in_ld + (long long)x * fx;
        double acc = 0.0;
        for (int a = 0; a < fy; ++a)
            for (int b = 0; b < fx; ++b) acc += (double)p[a * in_ld + b];
        out[i] = (R)(acc * (double)scale);
    }
}

template <typename R>
__global__ void tile_kernel(const R* __restrict__ in, int ny, int nx, int fy, int fx, R scale, R* __restrict__ out) {
    const long long ox = (long long)nx * fx, n = (long long)ny * fy * ox;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long y = i / ox, x = i - y * ox;
        out[i] = scale * in[(y / fy) * nx + x / fx];
    }
}

// kind 0: sinc(fx wx) sinc(fy wy) (pixel_ft);  1: cos(2 wx fx) cos(2 wy fy) (olpf_ft)
template <typename R>
__global__ void separable_tf_kernel(int kind, const R* __restrict__ fxv, const R* __restrict__ fyv, int ny, int nx, double wx,
                                    double wy, R* __restrict__ out) {
    const long long n = (long long)ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        const double ax = (double)fxv[x] * wx, ay = (double)fyv[y] * wy;
        double v;
        if (kind == 0) v = (ax == 0.0 ? 1.0 : sinpi(ax) / (M_PI * ax)) * (ay == 0.0 ? 1.0 : sinpi(ay) / (M_PI * ay));
        else v = cos(2.0 * ax) * cos(2.0 * ay);
        out[i] = (R)v;
    }
}

}  // namespace pb

using namespace pb;

#define PB_HANDLE(hh)                                   \
    PB_ENTER(hh);                      \
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128"); \
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int pb_pack_complex(pb_handle_t hh, int dtype, const void* re, const void* im, const double* im_scale_dev,
                               long long count, void* out, void* stream) {
    PB_HANDLE(hh);
    if (!re || !out) return fail(h, PB_ERR_INVALID, "null pointer");
    if (count <= 0) return PB_OK;
    const int g = grid_for(count, 256, h->sm_count);
    if (dtype == PB_C64) pack_complex_kernel<float><<<g, 256, 0, st>>>((const float*)re, (const float*)im, im_scale_dev, count, (float2*)out);
    else pack_complex_kernel<double><<<g, 256, 0, st>>>((const double*)re, (const double*)im, im_scale_dev, count, (double2*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_balance_scale(pb_handle_t hh, int dtype, const void* a, const void* b, long long count, double* s_dev,
                                void* stream) {
    PB_HANDLE(hh);
    if (!a || !b || !s_dev || count <= 0) return fail(h, PB_ERR_INVALID, "bad balance arguments");
    void* d = nullptr;
    PB_TRY(ensure_scratch(h, 2, 2 * sizeof(double), &d));
    PB_CUDA(h, cudaMemsetAsync(d, 0, 2 * sizeof(double), st));
    const int g = (int)std::min<long long>((count + 255) / 256, (long long)h->sm_count * 4);
    if (dtype == PB_C64) norms2_kernel<float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, count, (double*)d);
    else norms2_kernel<double><<<g, 256, 0, st>>>((const double*)a, (const double*)b, count, (double*)d);
    balance_finalize_kernel<<<1, 1, 0, st>>>((const double*)d, s_dev);
    PB_LAUNCH_CHECK(h);
    h->launches++;
    return PB_OK;
}

extern "C" int pb_packed_spectrum_product(pb_handle_t hh, int dtype, const void* Z, int ny, int nx, double scale,
                                          const double* im_scale_dev, void* out, void* stream) {
    PB_HANDLE(hh);
    if (!Z || !out || ny < 1 || nx < 1 || Z == out) return fail(h, PB_ERR_INVALID, "bad spectrum-product arguments (out must not alias Z)");
    const int g = grid_for((long long)ny * nx, 256, h->sm_count);
    if (dtype == PB_C64) packed_spectrum_product_kernel<float><<<g, 256, 0, st>>>((const float2*)Z, ny, nx, (float)scale, im_scale_dev, (float2*)out);
    else packed_spectrum_product_kernel<double><<<g, 256, 0, st>>>((const double2*)Z, ny, nx, scale, im_scale_dev, (double2*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_bindown(pb_handle_t hh, int dtype, const void* in, int ny, int nx, int fy, int fx, int mean, void* out,
                          void* stream) {
    PB_HANDLE(hh);
    if (!in || !out || fy < 1 || fx < 1 || ny < fy || nx < fx || ny % fy || nx % fx)
        return fail(h, PB_ERR_INVALID, "array shape must be a positive integer multiple of the binning factor");
    const int oy = ny / fy, ox = nx / fx;
    const double scale = mean ? 1.0 / ((double)fy * fx) : 1.0;
    const int g = grid_for((long long)oy * ox, 256, h->sm_count);
    if (dtype == PB_C64) bindown_kernel<float><<<g, 256, 0, st>>>((const float*)in, oy, ox, fy, fx, (float)scale, (float*)out);
    else bindown_kernel<double><<<g, 256, 0, st>>>((const double*)in, oy, ox, fy, fx, scale, (double*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_tile(pb_handle_t hh, int dtype, const void* in, int ny, int nx, int fy, int fx, double scale, void* out,
                       void* stream) {
    PB_HANDLE(hh);
    if (!in || !out || fy < 1 || fx < 1 || ny < 1 || nx < 1) return fail(h, PB_ERR_INVALID, "bad tile arguments");
    const int g = grid_for((long long)ny * fy * nx * fx, 256, h->sm_count);
    if (dtype == PB_C64) tile_kernel<float><<<g, 256, 0, st>>>((const float*)in, ny, nx, fy, fx, (float)scale, (float*)out);
    else tile_kernel<double><<<g, 256, 0, st>>>((const double*)in, ny, nx, fy, fx, scale, (double*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

extern "C" int pb_separable_tf(pb_handle_t hh, int dtype, int kind, const void* fx, const void* fy, int ny, int nx, double wx,
                               double wy, void* out, void* stream) {
    PB_HANDLE(hh);
    if (!fx || !fy || !out || ny < 1 || nx < 1 || kind < 0 || kind > 1) return fail(h, PB_ERR_INVALID, "bad transfer-function arguments");
    const int g = grid_for((long long)ny * nx, 256, h->sm_count);
    if (dtype == PB_C64) separable_tf_kernel<float><<<g, 256, 0, st>>>(kind, (const float*)fx, (const float*)fy, ny, nx, wx, wy, (float*)out);
    else separable_tf_kernel<double><<<g, 256, 0, st>>>(kind, (const double*)fx, (const double*)fy, ny, nx, wx, wy, (double*)out);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}
