/* The headline call from plain C: a 2048x2048 complex64 pupil on the device -> its 4096x4096 focal field
 * (prysm: Wavefront.focus(efl, Q=2), prysm/propagation/fft.py:7-25).  Build:
 *     gcc -std=c99 -Iinclude examples/focus_from_c.c -Lprysm_b200/_lib -lprysm_b200 -L/usr/local/cuda/lib64 -lcudart
 * The caller owns every buffer; the library owns only the handle. */
#include <stdio.h>

#include "prysm_b200.h"

/* the two CUDA runtime calls this example needs, declared here so that it compiles without cuda_runtime.h */
extern int cudaMalloc(void** p, size_t bytes);
extern int cudaFree(void* p);
extern int cudaDeviceSynchronize(void);

int main(void) {
    const int n = 2048, k = 4096, batch = 4;
    pb_handle_t h = NULL;
    void *pupil = NULL, *psf = NULL;
    int rc = pb_create(&h, 0);
    if (rc != PB_OK) { fprintf(stderr, "pb_create failed (%d): is a CUDA device visible?\n", rc); return 1; }
    if (cudaMalloc(&pupil, (size_t)batch * n * n * 8) || cudaMalloc(&psf, (size_t)batch * k * k * 8)) return 2;
    /* ... fill `pupil` (interleaved complex64, row-major [y][x]) ... */
    rc = pb_fft2(h, PB_C64, pupil, PB_IN_COMPLEX, NULL, PB_AMP_NONE, 0.0, n, n, n, k, k, /*dir*/ -1, /*scale*/ 1.0 / k,
                 /*shift_in*/ 1, /*shift_out*/ 1, psf, PB_OUT_COMPLEX, 1.0, k, k, k, /*stream*/ NULL);
    if (rc != PB_OK) { fprintf(stderr, "pb_fft2: %s\n", pb_last_error(h)); return 3; }
    /* the same for a stack of fields in one call: they share launches */
    rc = pb_fft2_batch(h, PB_C64, pupil, PB_IN_COMPLEX, NULL, PB_AMP_NONE, 0.0, batch, (long long)n * n, 0, n, n, n, k, k, -1,
                       1.0 / k, 1, 1, psf, PB_OUT_COMPLEX, 1.0, k, k, k, (long long)k * k, NULL);
    if (rc != PB_OK) { fprintf(stderr, "pb_fft2_batch: %s\n", pb_last_error(h)); return 4; }
    cudaDeviceSynchronize();
    printf("%s: %lld kernel launches\n", pb_version(), pb_launch_count(h));
    cudaFree(pupil);
    cudaFree(psf);
    pb_destroy(h);
    return 0;
}
