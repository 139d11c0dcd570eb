// Microbenchmark: issue / pipe throughput of scalar vs packed FP32 on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
template <int MODE>
__global__ void k(float2* out, float2 seed) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-9f, -1e-9f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }          // 2 FFMA
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                              // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                               // 2 FADD
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }                                                 // 1 FADD2
            if (MODE == 4) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; }                               // 2 FMUL
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], m); }                                                 // 1 FMUL2
            if (MODE == 6) { a[i].x = fmaf(a[i].x, a[(i + 1) & 7].y, a[(i + 3) & 7].x); a[i].y = fmaf(a[i].y, a[(i + 2) & 7].x, a[(i + 5) & 7].y); } // 3-reg FFMA
            if (MODE == 7) { a[i] = __ffma2_rn(a[i], a[(i + 1) & 7], a[(i + 3) & 7]); }                   // 3-reg FFMA2
        }
    }
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float2* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 2, threads = 512;
    k<MODE><<<blocks, threads>>>(d, make_float2(1.f, 2.f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, make_float2(1.f, 2.f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop_lane_ops = (double)blocks * threads * ITERS * 8 * 2;  // fp32 lane-ops (fma counted once)
    printf("%-12s %8.3f ms  %7.2f T lane-op/s  (%.1f per clk per SM @1.9GHz)\n", name, ms, flop_lane_ops / ms / 1e9, flop_lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 2 * 512 * sizeof(float2));
    run<0>("FFMA imm", d); run<1>("FFMA2 imm", d); run<2>("FADD", d); run<3>("FADD2", d);
    run<4>("FMUL", d); run<5>("FMUL2", d); run<6>("FFMA 3reg", d); run<7>("FFMA2 3reg", d);
    return 0;
}
