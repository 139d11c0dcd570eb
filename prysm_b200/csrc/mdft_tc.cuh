// tcgen05 / TMEM complex GEMM for the complex64 matrix DFT (see mdft_tc.cu).
#pragma once
#include "common.cuh"

namespace pb {

// PB_ERR_UNSUPPORTED (error string untouched) when the shape / dtype is not covered.
int try_mdft_tc(Handle* h, int dtype, const void* Ey, const void* Ex, int my, int ny, int mx, int nx, const void* a,
                void* out, double norm, int adjoint, int left_first, void* work, cudaStream_t st);

bool mdft_tc_shape_ok(int my, int ny, int mx, int nx);

int cgemm_simt(Handle* h, int dtype, int opA, int opB, int m, int n, int k, double alpha, const void* A,
               long long lda, const void* B, long long ldb, void* C, long long ldc, cudaStream_t st);

}  // namespace pb
