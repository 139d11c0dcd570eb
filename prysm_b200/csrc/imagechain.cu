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

}  // namespace pb

using namespace pb;

#define PB_HANDLE(hh)                                   \
    Handle* h = reinterpret_cast<Handle*>(hh);          \
    if (!h) return PB_ERR_INVALID;                      \
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
