// Tuned register-resident FFT kernels for the headline shapes.
#include "fft_tuned.cuh"

namespace pb {

int try_tuned_axis_pass(Handle*, const AxisPass&, cudaStream_t) { return PB_ERR_UNSUPPORTED; }

int try_tuned_fft2(Handle*, int, const void*, int, const void*, int, double, int, int, long long, int, int, int, double,
                   int, int, void*, int, double, int, int, long long, cudaStream_t) {
    return PB_ERR_UNSUPPORTED;
}

int try_tuned_angular_spectrum(Handle*, int, const void*, int, int, int, int, const void*, const void*, const void*,
                               int, void*, int, int, cudaStream_t) {
    return PB_ERR_UNSUPPORTED;
}

}  // namespace pb
