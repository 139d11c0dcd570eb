/*
 * prysm_b200.h -- C ABI of the B200-native propagation engine (libprysm_b200.so).
 *
 * This is the drop-in boundary for the prysm hot path (prysm.propagation + prysm.fttools +
 * the psf/otf reductions, reference v0.22).  The reference has no FFI of its own: its plug-in
 * surface is the Python module shim in prysm/mathops.py:11-116.  Each entry point below
 * states the reference call it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All array pointers are DEVICE pointers owned
 *     by the caller (row-major, index order [y][x], interleaved complex) unless a parameter is
 *     documented as "host".  The library owns only the opaque handle (twiddle tables, scratch).
 *   - Every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  A handle is bound to one device and is not thread-safe.
 *   - Return value: PB_OK (0) or a negative pb_status; pb_last_error(h) gives the message.
 *     Nothing throws or exits across this boundary.  There is no CPU fallback.
 *   - dtype selects the working precision like prysm.conf.config.precision
 *     (prysm/conf.py:28-96): PB_C64 = complex64 fields / float32 reals, PB_C128 = complex128 /
 *     float64.
 */
#ifndef PRYSM_B200_H
#define PRYSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_handle_s* pb_handle_t;

typedef enum { PB_C64 = 0, PB_C128 = 1 } pb_dtype;

typedef enum {
    PB_OK = 0,
    PB_ERR_INVALID = -1,      /* bad argument */
    PB_ERR_UNSUPPORTED = -2,  /* shape / length outside what the kernels cover */
    PB_ERR_CUDA = -3,         /* CUDA runtime error (message has the cudaError string) */
    PB_ERR_ALLOC = -4
} pb_status;

/* how the input field of pb_fft2 is presented */
typedef enum {
    PB_IN_COMPLEX = 0,  /* in = complex (ny, nx) */
    PB_IN_REAL = 1,     /* in = real (ny, nx), imaginary part 0 (prysm/otf.py:31 takes a real PSF) */
    PB_IN_AMP_OPD = 2   /* field = amp * exp(i*kscale*opd): prysm/propagation/wavefront.py:59-79 fused in */
} pb_in_kind;

typedef enum { PB_AMP_NONE = 0, PB_AMP_REAL = 1, PB_AMP_U8 = 2 } pb_amp_kind;

typedef enum {
    PB_OUT_COMPLEX = 0,    /* complex field */
    PB_OUT_INTENSITY = 1,  /* re^2+im^2 (prysm/propagation/wavefront.py:147-151) into a real array */
    PB_OUT_ACCUMULATE = 2  /* out += weight*(re^2+im^2): the per-wavelength term of
                              prysm/polynomials/fitting.py:37 applied as it is produced */
} pb_out_kind;

/* ---- lifetime ------------------------------------------------------------------------ */
int pb_create(pb_handle_t* out, int device);
int pb_destroy(pb_handle_t h);
const char* pb_last_error(pb_handle_t h);
const char* pb_version(void);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
long long pb_launch_count(pb_handle_t h);

/* ---- centred / plain 2-D FFT with fused pad, shifts, synthesis and |.|^2 ---------------
 * out = crop( S_out( FFT2_{dir}( S_in( pad( field ) ) ) ) ) * scale
 *   field : (ny, nx), see pb_in_kind; row pitch in_ld elements
 *   pad   : centred zero pad to (ky, kx) with the pad2d offset rule ceil((k-n)/2)
 *           (prysm/fttools.py:43-100)
 *   S_in  : ifftshift if shift_in, S_out: fftshift if shift_out (all axes)
 *   dir   : -1 forward (exp(-i..)), +1 inverse; scale is applied by the caller's norm rule
 *           (ortho: 1/sqrt(ky*kx); numpy default: 1 forward, 1/(ky*kx) inverse)
 *   crop  : centred (oy, ox) window with the crop_center rule (prysm/fttools.py:103-125)
 * Replaces: propagation.focus / unfocus and their adjoints (prysm/propagation/fft.py:7-85),
 *           fft.fft2 / ifft2 call sites (prysm/otf.py:31,59), Wavefront.from_amp_and_phase +
 *           .focus + .intensity chains (prysm/propagation/wavefront.py:59-79,147-151,478-504).
 * Any ky, kx >= 1 is accepted (non powers of two run Bluestein on the same kernels). */
int pb_fft2(pb_handle_t h, int dtype,
            const void* in, int in_kind, const void* amp, int amp_kind, double kscale,
            int ny, int nx, long long in_ld,
            int ky, int kx, int dir, double scale, int shift_in, int shift_out,
            void* out, int out_kind, double weight, int oy, int ox, long long out_ld,
            void* stream);
/* pb_fft2 over `batch` independent fields in one call: field b starts at in + b*in_bs, amp + b*amp_bs (0 = one
 * amplitude shared by all) and out + b*out_bs, strides in elements of the respective array's scalar type.  On the
 * fused focus shapes the fields share launches (eight per launch pair by default), which evens out the 3.46-wave
 * column pass and the 9.2-round row pass of a single 2048^2 field; other shapes run field by field.  The reference
 * has no batch form: fft.py:7-25 is called once per wavefront. */
int pb_fft2_batch(pb_handle_t h, int dtype, const void* in, int in_kind, const void* amp, int amp_kind,
                  double kscale, int batch, long long in_bs, long long amp_bs, int ny, int nx,
                  long long in_ld, int ky, int kx, int dir, double scale, int shift_in, int shift_out,
                  void* out, int out_kind, double weight, int oy, int ox, long long out_ld,
                  long long out_bs, void* stream);

/* ---- 1-D FFT along one axis of a 2-D array, numpy fft(a, n, axis) semantics ------------
 * Zero-extends or truncates to n along `axis` (0 = y/columns, 1 = x/rows).
 * Replaces fft.fft / fft.ifft call sites in prysm/fttools.py:301-321, 519-533. */
int pb_fft1(pb_handle_t h, int dtype, const void* in, int ny, int nx, long long in_ld,
            int axis, int n, int dir, double scale, void* out, long long out_ld, void* stream);

/* ---- the same, with fused multipliers and an output window -----------------------------
 * For every line b along `axis`:  u[j] = in_line[j] * pre_e[j] * pre_b[b]   (j < n_line, zero-extended to n)
 *                                 U = DFT_n(u);   out_line[q] = U[q + out_off] * post_e[q] * post_b[b] * scale
 * for q in [0, n_out).  Any multiplier may be NULL; *_conj conjugates it.  pre_e has n_line
 * entries, post_e has n_out entries, pre_b / post_b have one entry per line.
 * One call is one Bluestein half-step of CZT (prysm/fttools.py:298-324) or one FFTDFT axis
 * (prysm/fttools.py:443-459) with the chirp / ramp multiplies and the slice fused in. */
int pb_axis_dft(pb_handle_t h, int dtype, const void* in, int ny, int nx, long long in_ld, int axis,
                int n, int dir, double scale,
                const void* pre_e, int pre_e_conj, const void* pre_b, int pre_b_conj,
                const void* post_e, int post_e_conj, const void* post_b, int post_b_conj,
                int out_off, int n_out, void* out, long long out_ld, void* stream);

/* ---- one Bluestein axis of the chirp-z transform, fused -----------------------------------
 * For every line along `axis`:  u = line * pre_e (zero-extended to K);  c = IFFT_K( FFT_K(u) * H );
 * out_line[q] = c[q + out_off] * post_e[q] * scale,  q in [0, n_out).   K must be a power of two.
 * On the tuned path the data stays in registers between the two transforms (one kernel, one read and one
 * write of the array).  Replaces the per-axis block of CZT.__call__ / adjoint, prysm/fttools.py:296-361. */
int pb_czt_axis(pb_handle_t h, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int K,
                const void* pre_e, int pre_conj, const void* H, const void* post_e, int post_conj,
                int out_off, int n_out, double scale, void* out, long long out_ld, void* stream);

/* The same axis with the modulus fused into its store: out_kind = PB_OUT_INTENSITY writes |.|^2 (a REAL array),
 * PB_OUT_ACCUMULATE adds weight * |.|^2 to it -- the last pass of one wavelength of an incoherent sum never writes its
 * complex field.  Replaces the last axis of CZT.__call__ (prysm/fttools.py:296-325) followed by Wavefront.intensity
 * (prysm/propagation/wavefront.py:147-151) and the running term of sum_of_2d_modes (prysm/polynomials/fitting.py:37). */
int pb_czt_axis_intensity(pb_handle_t h, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int K,
                          const void* pre_e, int pre_conj, const void* H, const void* post_e, int post_conj,
                          int out_off, int n_out, double scale, int out_kind, double weight, void* out,
                          long long out_ld, void* stream);

/* ---- the wavelength loop of an incoherent (polychromatic) PSF as one call -------------------------------------
 * plane(m,m) += sum_i weight_i * | CZT_i( amp * exp(i * kscale_i * opd) ) |^2  over n_units wavelengths of a square n x n
 * pupil on a common square m x m focal grid.  Replaces, per wavelength, Wavefront.from_amp_and_phase
 * (prysm/propagation/wavefront.py:59-79) -> prepare_executor(kind='czt') (propagation/dft.py:69-117, fttools.py:257-291)
 * -> focus_dft (fttools.py:296-325) -> .intensity (wavefront.py:147-151) and the running weighted sum of
 * sum_of_2d_modes (polynomials/fitting.py:37) -- the loop of docs/source/how-tos/Polychromatic Propagation.ipynb:86-98.
 * units: n_units records of 8 doubles on the HOST: kscale (phase per unit of opd), then the pb_czt_plan scalars shift,
 * alpha, xc, f0, df of the (square, centred) geometry, the executor norm, the weight.  K: Bluestein length, a power of
 * two >= n + m - 1.  The plan of unit i+1 is built on the handle's helper stream while unit i transforms.
 * work: pb_polychromatic_czt_work_bytes(dtype, n, m, K) bytes, 256-byte aligned.  plane: real (float / double). */
long long pb_polychromatic_czt_work_bytes(int dtype, int n, int m, int K);
int pb_polychromatic_czt(pb_handle_t h, int dtype, const void* amp, int amp_kind, const void* opd, int n, int m, int K,
                         int n_units, const double* units, void* work, void* plane, void* stream);

/* The Bluestein pieces of one CZT axis built on the device from scalars (no host maths, no uploads):
 * b (N), post = a*phase (M), H = FFT_K(h) (K), Hadj (K) as consumed by pb_czt_axis.  alpha = dx*dfx,
 * shift = f[M/2]/df, xc = x[N/2], f0 = f[0], df the frequency step.  prysm/fttools.py:257-291, 372-389 */
int pb_czt_plan(pb_handle_t h, int dtype, int N, int M, int K, double shift, double alpha, int sign,
                double xc, double f0, double df, void* b, void* post, void* H, void* Hadj, void* stream);

/* ---- angular spectrum ------------------------------------------------------------------
 * out(ky,kx) = ifft2( fft2( pad(in -> ky,kx) ) * TF ), TF = outer(ty, tx) when tf == NULL,
 * else the full (ky,kx) array tf; conj_tf applies conj(TF) (adjoint); the result is cropped to
 * (oy, ox) with the crop_center rule (forward: oy=ky, ox=kx i.e. stays padded).
 * Replaces prysm/propagation/angular_spectrum.py:9-79. */
int pb_angular_spectrum(pb_handle_t h, int dtype, const void* in, int ny, int nx,
                        int ky, int kx, const void* ty, const void* tx, const void* tf,
                        int conj_tf, void* out, int oy, int ox, void* stream);

/* The same with the field multiplied by a complex (ny,nx) `screen` (conj_screen: by its conjugate) inside the first
 * transform pass: one step `wf = (wf * screen).free_space(dz)` of a plane-to-plane chain without the separate
 * elementwise pass.  Replaces prysm/propagation/wavefront.py:360-383 (Wavefront.__mul__) followed by
 * prysm/propagation/angular_spectrum.py:9-42. */
int pb_angular_spectrum_screen(pb_handle_t h, int dtype, const void* in, const void* screen, int conj_screen,
                               int ny, int nx, int ky, int kx, const void* ty, const void* tx, const void* tf,
                               int conj_tf, void* out, int oy, int ox, void* stream);

/* the two separable factors exp(-i*pi*wvl_mm*z*k^2), k = fftfreq(n, dx) rounded to the working
 * precision first (prysm/propagation/angular_spectrum.py:102-113).  ty: ky values, tx: kx. */
int pb_angular_spectrum_vectors(pb_handle_t h, int dtype, int ky, int kx, double wvl_um,
                                double dx_mm, double z_mm, void* ty, void* tx, void* stream);

/* ---- matrix DFT ------------------------------------------------------------------------
 * basis E[j,l] = exp(sign*2*pi*i*f[j]*x[l]), (m, n) complex.  f, x: HOST double arrays.
 * The phase is formed and range-reduced in fp64 before the sincos.  prysm/fttools.py:187-191 */
int pb_mdft_basis(pb_handle_t h, int dtype, const double* f_host, int m, const double* x_host,
                  int n, int sign, void* E, void* stream);

/* C(m,n) = alpha * opA(A) * opB(B); op: 0 = as is, 1 = transpose, 2 = conjugate transpose,
 * 3 = conjugate.  A is (m,k) after op, B is (k,n) after op.  Complex, row-major.
 * Replaces the `@` chains of MDFT.__call__ / adjoint (prysm/fttools.py:201-228). */
int pb_cgemm(pb_handle_t h, int dtype, int opA, int opB, int m, int n, int k, double alpha,
             const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
             void* stream);

/* out(my,mx) = norm * Ey(my,ny) @ a(ny,nx) @ Ex(mx,nx)^T   (adjoint = 0), or
 * out(ny,nx) = norm * Ey^H @ g(my,mx) @ conj(Ex)            (adjoint = 1),
 * association order chosen by left_first (prysm/fttools.py:198-228).  `work` is caller scratch
 * of at least pb_mdft_work_elems(...) complex elements. */
int pb_mdft_apply(pb_handle_t h, int dtype, const void* Ey, const void* Ex, int my, int ny,
                  int mx, int nx, const void* a, void* out, double norm, int adjoint,
                  int left_first, void* work, void* stream);
long long pb_mdft_work_elems(int my, int ny, int mx, int nx, int adjoint, int left_first);

/* ---- matrix DFT on the tcgen05 tensor cores (complex64, forward) -------------------------
 * 3xTF32 split (fp32-accurate) real embedding of the complex product; see csrc/mdft_tc.cu.
 *   pb_mdft_tc_supported : 1 when (my, ny, mx, nx) is covered (multiples of 128 / 128 / 128 / 16)
 *   pb_mdft_tc_expand    : complex basis E (m,n) -> real expansions hi, lo, each (2m, 2n) fp32;
 *                          done once per executor for Ex and for Ey (prysm/fttools.py:187-191)
 *   pb_mdft_tc_apply     : out(my,mx) = norm * Ey @ a @ Ex^T   (prysm/fttools.py:201-207);
 *                          work: pb_mdft_tc_work_bytes() bytes of 16-byte aligned device scratch (the transposed
 *                          intermediate, split-K partial planes, and the partial tiles + flags of the stream-K form
 *                          that under-filled tile grids take) */
int pb_mdft_tc_supported(int my, int ny, int mx, int nx);
int pb_mdft_tc_expand(pb_handle_t h, const void* E, int m, int n, void* hi, void* lo, void* stream);
long long pb_mdft_tc_work_bytes(int my, int ny, int mx, int nx);
int pb_mdft_tc_apply(pb_handle_t h, const void* ExB_hi, const void* ExB_lo, const void* EyB_hi,
                     const void* EyB_lo, int my, int ny, int mx, int nx, const void* a, void* out,
                     double norm, void* work, void* stream);

/* ---- elementwise / reductions -----------------------------------------------------------
 * out = amp * exp(i*kscale*opd)   (amp optional).  prysm/propagation/wavefront.py:59-96 */
int pb_phase_screen(pb_handle_t h, int dtype, const void* amp, int amp_kind, const void* opd,
                    double kscale, long long count, void* out, void* stream);
/* out (+)= weight * |in|^2.  prysm/propagation/wavefront.py:147-151 */
int pb_intensity(pb_handle_t h, int dtype, const void* in, long long count, double weight,
                 int accumulate, void* out, void* stream);
/* out = a (op) b elementwise on complex arrays; b == NULL uses the scalar (s_re, s_im) instead;
 * reverse swaps the operands.  op: 0 mul, 1 div, 2 add, 3 sub.
 * Wavefront.__numerical_operation__ (prysm/propagation/wavefront.py:360-411). */
int pb_binary(pb_handle_t h, int dtype, int op, const void* a, const void* b, double s_re,
              double s_im, int reverse, long long count, void* out, void* stream);
/* out[y,x] = in[y,x] * vy[y] * vx[x] * scale (either vector may be NULL; conj flags per vector).
 * The chirp / phase-ramp multiplies of CZT and FFTDFT (prysm/fttools.py:298-324, 443-459). */
int pb_mul_outer(pb_handle_t h, int dtype, const void* in, long long in_ld, int ny, int nx,
                 const void* vy, int conj_y, const void* vx, int conj_x, double scale, void* out,
                 long long out_ld, void* stream);
/* out(n) = sum_k weights[k] * modes[k][n]; weights: HOST doubles.  prysm/polynomials/fitting.py:37 */
int pb_weighted_sum(pb_handle_t h, int dtype, const void* modes, int k, long long count,
                    const double* weights_host, void* out, void* stream);
/* centre-normalised transfer functions from the complex transform D (ny,nx):
 * n = D / D[ny/2, nx/2]; which bit0 -> mtf=|n| (real), bit1 -> ptf=angle(n) (real), bit2 -> otf=n.
 * prysm/otf.py:62-202 */
int pb_otf_normalize(pb_handle_t h, int dtype, const void* D, int ny, int nx, int which,
                     void* mtf, void* ptf, void* otf, void* stream);
/* Baliga-Cohn encircled energy of a real MTF array (ny,nx) with frequency spacing df [cy/mm] at nr radii
 * [mm] (HOST arrays in/out; synchronises the stream).  prysm/otf.py:319-414 */
int pb_encircled_energy(pb_handle_t h, int dtype, const void* mtf, int ny, int nx, double df,
                        const double* radii_mm_host, int nr, double* out_host, void* stream);
/* moments of a real array: sums_host[0..2] = sum(d), sum(d*y), sum(d*x) (HOST doubles; this call
 * synchronises the stream).  prysm/psf.py:174-203 */
int pb_moments(pb_handle_t h, int dtype, const void* data, int ny, int nx, double* sums_host,
               void* stream);


/* ---- adjoint twins and Lyot-coronagraph compositions -----------------------
 * masked multiply, the elementwise step of every composition in prysm/propagation/coronagraph.py and of
 * _adjoint_multiply (prysm/propagation/_kernels.py:30-38):
 *     out (+)= scale * (a - b) * f(m) * w
 * a: complex; b: complex or NULL; m: real / complex (m_kind) or NULL; w: real or NULL.  flags (OR of pb_mask_flags):
 * CONJ -> f = conj(m); ONE_MINUS -> f = 1 - m (Babinet complement, coronagraph.py:338); REAL_OUT -> out is a
 * REAL array receiving the real part (the real=True branch of _adjoint_multiply); ACCUMULATE -> out += ... */
typedef enum { PB_MASK_REAL = 0, PB_MASK_COMPLEX = 1 } pb_mask_kind;
typedef enum { PB_MASK_CONJ = 1, PB_MASK_ONE_MINUS = 2, PB_MASK_REAL_OUT = 4, PB_MASK_ACCUMULATE = 8 } pb_mask_flags;
int pb_mask_multiply(pb_handle_t h, int dtype, const void* a, const void* b, const void* m, int m_kind,
                     int flags, const void* w, double scale, long long count, void* out, void* stream);
/* adjoints of from_amp_and_phase (prysm/propagation/wavefront.py:172-242), kscale = 2*pi/(1e3*wavelength):
 * mode 0: out (complex) = i*kscale*imag(bar*conj(field))              -- w.r.t. phase (the reference's prefix is
 *         complex, so its result is purely imaginary; kept);
 * mode 1: out (real) = real(bar*conj(field))/|field|, 0 where field=0 -- w.r.t. amplitude, phasor from the field;
 * mode 2: out (real) = real(bar*conj(exp(i*kscale*opd)))              -- w.r.t. amplitude, phasor rebuilt from opd. */
int pb_field_adjoint(pb_handle_t h, int dtype, int mode, const void* field, const void* bar,
                     const void* opd, double kscale, long long count, void* out, void* stream);
/* real array out = re (which 0) / im (1) / angle (2) / abs (3) of a complex array: Wavefront.real / .imag / .phase
 * (prysm/propagation/wavefront.py:153-166). */
int pb_component(pb_handle_t h, int dtype, int which, const void* in, long long count, void* out,
                 void* stream);
/* out_host[0..1] = sum w * a * conj(b) (w real or NULL; fp64 accumulation; synchronises the stream).
 * The reduction of thin_lens_adjoint (prysm/propagation/wavefront.py:244-280) and of <x, y> adjoint checks. */
int pb_dot(pb_handle_t h, int dtype, const void* a, const void* b, const void* w, long long count,
           double* out_host, void* stream);
/* out_host[2j..2j+1] = sum_n modes[j][n] * bar[n] for k real modes; bar real or complex (bar_complex).
 * sum_of_2d_modes_adjoint (prysm/polynomials/fitting.py:40-57).  Synchronises the stream. */
int pb_mode_projection(pb_handle_t h, int dtype, const void* modes, int k, long long count,
                       const void* bar, int bar_complex, double* out_host, void* stream);
/* k-space gradient of mtf (which=1, bar real) / ptf (2, bar real) / otf (4, bar complex) _from_psf_adjoint given
 * the complex transform D (ny,nx), including the centre-normalisation term (prysm/otf.py:205-316); the caller
 * finishes with transform_psf_adjoint (pb_fft2, dir=+1) and takes the real part. */
int pb_otf_adjoint_seed(pb_handle_t h, int dtype, int which, const void* bar, const void* D, int ny,
                        int nx, void* data_bar, void* stream);
/* MTF-plane gradient of encircled_energy: mtf_bar = sum_r ee_bar[r]*r*J1(2 pi r nu)/nu*df^2, radii [mm] and
 * ee_bar HOST arrays (prysm/otf.py:417-471). */
int pb_encircled_energy_adjoint_seed(pb_handle_t h, int dtype, int ny, int nx, double df,
                                     const double* radii_mm_host, const double* ee_bar_host, int nr,
                                     void* mtf_bar, void* stream);
/* out = exp(i*charge*atan2(yf, xf)) on real coordinate arrays: vortex_phase_mask
 * (prysm/propagation/coronagraph.py:102-132). */
int pb_vortex_phase(pb_handle_t h, int dtype, int charge, const void* xf, const void* yf,
                    long long count, void* out, void* stream);
/* out = bilinear interpolation of the complex `map` (ny, nx) at col = (xf - cx)/dx + nx/2, row = (yf - cy)/dx + ny/2 with
 * edge replication, and `fill` (a complex array like out, or the scalar (fill_re, fill_im) when NULL) where a coordinate
 * leaves [0, n-1]: the fpm(xf, yf) callable of prepare_measured_fpm at order=1
 * (prysm/propagation/coronagraph.py:135-209; scipy.ndimage.map_coordinates(order=1, mode='nearest')). */
int pb_resample_bilinear(pb_handle_t h, int dtype, const void* map, int ny, int nx, const void* xf,
                         const void* yf, long long count, double cx, double cy, double dx, const void* fill,
                         double fill_re, double fill_im, void* out, void* stream);
/* one level of prepare_multiresolution (prysm/propagation/dft.py:155-167, 262-292): focal grids
 * xf = (ix - nx/2)*fdx + shift, yf likewise, and the hand-off window
 * taper(r; a0, b0) - taper(r; a1, b1), taper = 1 - smootherstep((r - a)/(b - a));
 * a0 < 0 -> first term 1 (coarsest level); a1 < 0 -> second term 0 (finest level).  Any output may be NULL. */
int pb_radial_window(pb_handle_t h, int dtype, int ny, int nx, double fdx, double shift, double a0,
                     double b0, double a1, double b1, void* win, void* xf, void* yf, void* stream);


/* ---- pupil synthesis: the step before the path ------------------------
 * x[iy,ix] = (ix - nx/2)*dx, y[iy,ix] = (iy - ny/2)*dx (make_xy_grid, prysm/coordinates.py:344-378) and their polar
 * form r = hypot(x, y), t = atan2(y, x) (cart_to_polar, coordinates.py:73-102), real arrays, any of them NULL. */
int pb_xy_grid(pb_handle_t h, int dtype, int ny, int nx, double dx, void* x, void* y, void* r, void* t,
               void* stream);
int pb_cart_to_polar(pb_handle_t h, int dtype, const void* x, const void* y, long long count, void* r,
                     void* t, void* stream);
/* aa_dx <= 0: out (uint8) = (r - radius <= 0): geometry.circle (prysm/geometry.py:337-372);
 * aa_dx  > 0: out (real)  = clip(0.5 - (r - radius)/aa_dx, 0, 1): antialias(circle_sdf(radius, r), aa_dx)
 *             (prysm/geometry.py:11-34), the grey-edge aperture coronagraph models need. */
int pb_circle(pb_handle_t h, int dtype, const void* r, long long count, double radius, double aa_dx,
              void* out, void* stream);
/* Jacobi polynomials P_j^(alpha,beta)(x), j = 0..nmax, by the three-term recurrence of
 * prysm/polynomials/jacobi.py:13-39, 147-175; order j is written to out[slots_host[j]] (a (n_out, count) real
 * array) when slots_host[j] >= 0.  nmax <= 120. */
int pb_jacobi_seq(pb_handle_t h, int dtype, const void* x, long long count, int nmax, double alpha,
                  double beta, const int* slots_host, void* out, void* stream);
/* Zernike polynomials Z_n^m = norm * P_{(n-|m|)/2}^{(0,|m|)}(2r^2-1) * r^|m| * {cos(m t), sin(|m| t)} for k HOST
 * (n, m) pairs, the Jacobi recurrences shared per |m| (zernike_nm_seq, prysm/polynomials/zernike.py:74-166).
 * polar != 0: a = r, b = t; polar == 0: a = x, b = y (Cartesian; zernike_sum's input form).
 * pb_zernike_seq writes the basis out (k, count); pb_zernike_sum writes out (count) = sum_k coefs[k]*Z_k without
 * materialising the basis (zernike_sum, zernike.py:169-181; zero weights are skipped there too). */
int pb_zernike_seq(pb_handle_t h, int dtype, int polar, const void* a, const void* b, long long count,
                   int k, const int* n_host, const int* m_host, int norm, void* out, void* stream);
int pb_zernike_sum(pb_handle_t h, int dtype, int polar, const void* a, const void* b, long long count,
                   int k, const int* n_host, const int* m_host, const double* coefs_host, int norm,
                   void* out, void* stream);


/* ---- image-chain consumers of the same FFTs ------------------------------
 * conv(obj, psf) of two REAL arrays (prysm/convolution.py:9-32) as one forward and one inverse transform:
 * pb_balance_scale: *s_dev (a DEVICE double) = sqrt(sum a^2 / sum b^2) (1 if either sum is 0 / not finite): the two
 *   arrays share one transform whose rounding error is relative to the larger of them, so the PSF is scaled to the
 *   object's norm for the round trip -- on the device, no host synchronisation;
 * pb_pack_complex: out = re + i*s*im (im NULL -> 0; im_scale_dev NULL -> s = 1); pb_fft2 of that gives Z = O + i*s*H;
 * pb_packed_spectrum_product: out[k] = scale/s * (Z[k]^2 - conj(Z[-k])^2)/(4i) = scale * O[k]*H[k], indices modulo
 *   (ny, nx), unshifted spectrum order; out must not alias Z. */
int pb_balance_scale(pb_handle_t h, int dtype, const void* a, const void* b, long long count, double* s_dev,
                     void* stream);
int pb_pack_complex(pb_handle_t h, int dtype, const void* re, const void* im, const double* im_scale_dev,
                    long long count, void* out, void* stream);
int pb_packed_spectrum_product(pb_handle_t h, int dtype, const void* Z, int ny, int nx, double scale,
                               const double* im_scale_dev, void* out, void* stream);

/* detector sampling of the image chain (prysm/detector.py:151-338), real arrays:
 * pb_bindown: out (ny/fy, nx/fx) = sum (mean != 0: mean) of each fy x fx block (bindown);
 * pb_tile:    out (ny*fy, nx*fx) = scale * in[y/fy, x/fx] (tile, the adjoint of bindown);
 * pb_separable_tf: out[y,x] = sinc(fx[x] wx) sinc(fy[y] wy) (kind 0, pixel_ft) or cos(2 wx fx[x]) cos(2 wy fy[y])
 *              (kind 1, olpf_ft) from the two frequency VECTORS. */
int pb_bindown(pb_handle_t h, int dtype, const void* in, int ny, int nx, int fy, int fx, int mean,
               void* out, void* stream);
int pb_tile(pb_handle_t h, int dtype, const void* in, int ny, int nx, int fy, int fx, double scale,
            void* out, void* stream);
int pb_separable_tf(pb_handle_t h, int dtype, int kind, const void* fx, const void* fy, int ny, int nx,
                    double wx, double wy, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRYSM_B200_H */
