// tcgen05 / TMEM complex GEMM for the complex64 matrix DFT (see mdft_tc.cu).
#pragma once
#include "common.cuh"

namespace pb {

bool mdft_tc_shape_ok(int my, int ny, int mx, int nx);

int cgemm_simt(Handle* h, int dtype, int opA, int opB, int m, int n, int k, double alpha, const void* A,
               long long lda, const void* B, long long ldb, void* C, long long ldc, cudaStream_t st);

}  // namespace pb
