// tcgen05 / TMEM complex GEMM for the complex64 matrix DFT.
#include "mdft_tc.cuh"

namespace pb {

int try_mdft_tc(Handle*, int, const void*, const void*, int, int, int, int, const void*, void*, double, int, int, void*,
                cudaStream_t) {
    return PB_ERR_UNSUPPORTED;
}

}  // namespace pb
