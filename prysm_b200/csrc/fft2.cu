// 2-D operators composed from axis passes: pb_fft2, pb_fft1, pb_angular_spectrum.
// Arbitrary (non power-of-two) lengths run Bluestein's algorithm on the same pass kernels.
#include <cmath>
#include <complex>

#include "axis_pass.cuh"
#include "fft_tuned.cuh"
#include "czt.cuh"

namespace pb {

static inline int ceil_half(int d) { return (d + 1) / 2; }  // ceil(d/2) for d >= 0

// ---------------------------------------------------------------------------------------
// host-side table builders
// ---------------------------------------------------------------------------------------

static void host_fft_pow2(std::vector<std::complex<double>>& a, int sign) {
    const int n = (int)a.size();
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= n; len <<= 1) {
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double ang = sign * 2.0 * M_PI * k / len;
                std::complex<double> w(cos(ang), sin(ang));
                auto u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

int upload_table(Handle* h, const TwKey& key, const std::vector<std::complex<double>>& t, const void** out) {
    const size_t n = t.size();
    void* d = nullptr;
    PB_CUDA(h, cudaMalloc(&d, n * csize(key.dtype)));
    if (key.dtype == PB_C64) {
        std::vector<float2> f(n);
        for (size_t i = 0; i < n; ++i) f[i] = make_float2((float)t[i].real(), (float)t[i].imag());
        PB_CUDA(h, cudaMemcpy(d, f.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    } else {
        PB_CUDA(h, cudaMemcpy(d, t.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    }
    h->tables[key] = d;
    *out = d;
    return PB_OK;
}

int get_twiddles(Handle* h, int n, int dtype, const void** out) {
    TwKey key{n, dtype, 0};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = it->second; return PB_OK; }
    std::vector<std::complex<double>> t(n);
    for (int k = 0; k < n; ++k) {
        // exact octant symmetry is not needed; long double keeps the table correctly rounded
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * k / n;
        t[k] = std::complex<double>((double)cosl(ang), (double)sinl(ang));
    }
    return upload_table(h, key, t, out);
}

static int bluestein_len(int n) { int m = 1; while (m < 2 * n - 1) m <<= 1; return m; }

// chirp c[j] = exp(dir*pi*i*j^2/n), j^2 reduced mod 2n exactly
static int get_chirp(Handle* h, int n, int dtype, int dir, const void** out) {
    TwKey key{n, dtype, dir < 0 ? 1 : 2};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = it->second; return PB_OK; }
    std::vector<std::complex<double>> t(n);
    for (int j = 0; j < n; ++j) {
        const long long r = ((long long)j * j) % (2LL * n);
        const long double ang = dir * 3.14159265358979323846264338327950288L * r / n;
        t[j] = std::complex<double>((double)cosl(ang), (double)sinl(ang));
    }
    return upload_table(h, key, t, out);
}

// spectrum of the Bluestein filter g[d mod M] = conj(c[|d|]), |d| < n
static int get_chirp_filter(Handle* h, int n, int dtype, int dir, const void** out) {
    TwKey key{n, dtype, dir < 0 ? 3 : 4};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = it->second; return PB_OK; }
    const int M = bluestein_len(n);
    std::vector<std::complex<double>> g(M, {0.0, 0.0});
    for (int d = 0; d < n; ++d) {
        const long long r = ((long long)d * d) % (2LL * n);
        const long double ang = -dir * 3.14159265358979323846264338327950288L * r / n;
        std::complex<double> v((double)cosl(ang), (double)sinl(ang));
        g[d] = v;
        if (d) g[M - d] = v;
    }
    host_fft_pow2(g, -1);
    return upload_table(h, key, g, out);
}

// ---------------------------------------------------------------------------------------
// one logical DFT pass of arbitrary length p.Llog (p.L is ignored on entry)
// ---------------------------------------------------------------------------------------
static int axis_dft(Handle* h, AxisPass p, cudaStream_t st) {
    const int n = p.Llog;
    if (is_pow2(n)) {
        p.L = n;
        return launch_axis_pass(h, p, st);
    }
    // Bluestein: X[k] = c[k] * IFFT_M( FFT_M(u*c) * G )[k]
    const int M = bluestein_len(n);
    const void *chirp = nullptr, *G = nullptr;
    PB_TRY(get_chirp(h, n, p.dtype, p.dir, &chirp));
    PB_TRY(get_chirp_filter(h, n, p.dtype, p.dir, &G));
    void* tmp = nullptr;
    PB_TRY(ensure_scratch(h, 2, (size_t)p.nb * M * csize(p.dtype), &tmp));

    AxisPass a = p;  // stage 1: chirp-modulated forward FFT_M times the filter spectrum
    a.L = M; a.Llog = n; a.Llog_out = M;
    a.pre_e2 = p.pre_e; a.pre_off2 = p.pre_off; a.pre_e2_conj = p.pre_e_conj;
    a.pre_e = chirp; a.pre_off = 0; a.pre_e_conj = 0;
    if (p.pre_e2) return fail(h, PB_ERR_UNSUPPORTED, "two pre-multipliers on a non power-of-two axis");
    a.dir = -1;
    a.out = tmp; a.n_out = M; a.crop_off = 0; a.rot_out = 0;
    a.obs = p.batch_contiguous ? 1 : M;
    a.oes = p.batch_contiguous ? p.nb : 1;
    a.post_e = G; a.post_off = 0; a.post_e_conj = 0;
    a.post_e2 = nullptr; a.post_b = nullptr; a.post_mat = nullptr;
    a.scale = 1.0; a.out_kind = PB_OUT_COMPLEX;
    PB_TRY(launch_axis_pass(h, a, st));

    AxisPass b = p;  // stage 2: inverse FFT_M, keep k < n, demodulate, user's epilogue
    b.L = M; b.Llog = M; b.Llog_out = n;
    b.in = tmp; b.in_kind = PB_IN_COMPLEX; b.amp = nullptr;
    b.ibs = a.obs; b.ies = a.oes;
    b.n_in = M; b.in_off = 0; b.rot_in = 0;
    b.pre_e = nullptr; b.pre_e2 = nullptr; b.pre_b = nullptr; b.pre_mat = nullptr;
    b.dir = +1;
    if (p.post_e2) return fail(h, PB_ERR_UNSUPPORTED, "two post-multipliers on a non power-of-two axis");
    b.post_e2 = p.post_e; b.post_off2 = p.post_off; b.post_e2_conj = p.post_e_conj;
    b.post_e = chirp; b.post_off = 0; b.post_e_conj = 0;
    b.scale = p.scale / M;
    return launch_axis_pass(h, b, st);
}

// ---------------------------------------------------------------------------------------
// pb_fft2
// ---------------------------------------------------------------------------------------
struct Fft2Args {
    int dtype;
    const void* in; int in_kind; const void* amp; int amp_kind; double kturns;
    int ny, nx; long long in_ld;
    int ky, kx, dir; double scale; int shift_in, shift_out;
    void* out; int out_kind; double weight; int oy, ox; long long out_ld;
};

static int fft2_generic(Handle* h, const Fft2Args& a, cudaStream_t st) {
    void* tmp = nullptr;
    PB_TRY(ensure_scratch(h, 0, (size_t)a.ny * a.ox * csize(a.dtype), &tmp));
    AxisPass x;  // rows: only the ny populated rows are transformed (zero rows stay implicit)
    x.dtype = a.dtype;
    x.in = a.in; x.in_kind = a.in_kind; x.amp = a.amp; x.amp_kind = a.amp_kind; x.kturns = a.kturns;
    x.ibs = a.in_ld; x.ies = 1; x.nb = a.ny;
    x.Llog = a.kx; x.n_in = a.nx; x.in_off = ceil_half(a.kx - a.nx); x.rot_in = a.shift_in ? a.kx / 2 : 0;
    x.dir = a.dir;
    x.out = tmp; x.obs = a.ox; x.oes = 1;
    x.n_out = a.ox; x.crop_off = ceil_half(a.kx - a.ox); x.rot_out = a.shift_out ? a.kx / 2 : 0;
    x.batch_contiguous = 0;
    PB_TRY(axis_dft(h, x, st));

    AxisPass y;  // columns
    y.dtype = a.dtype;
    y.in = tmp; y.ibs = 1; y.ies = a.ox; y.nb = a.ox;
    y.Llog = a.ky; y.n_in = a.ny; y.in_off = ceil_half(a.ky - a.ny); y.rot_in = a.shift_in ? a.ky / 2 : 0;
    y.dir = a.dir;
    y.out = a.out; y.obs = 1; y.oes = a.out_ld;
    y.n_out = a.oy; y.crop_off = ceil_half(a.ky - a.oy); y.rot_out = a.shift_out ? a.ky / 2 : 0;
    y.scale = a.scale; y.out_kind = a.out_kind; y.weight = a.weight;
    y.batch_contiguous = 1;
    return axis_dft(h, y, st);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_fft2(pb_handle_t hh, int dtype, const void* in, int in_kind, const void* amp, int amp_kind,
                       double kscale, int ny, int nx, long long in_ld, int ky, int kx, int dir, double scale,
                       int shift_in, int shift_out, void* out, int out_kind, double weight, int oy, int ox,
                       long long out_ld, void* stream) {
    PB_ENTER(hh);
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || ky < ny || kx < nx) return fail(h, PB_ERR_INVALID, "need 1 <= n <= k on both axes");
    if (oy < 1 || ox < 1 || oy > ky || ox > kx) return fail(h, PB_ERR_INVALID, "crop window must fit the transform");
    if (dir != -1 && dir != 1) return fail(h, PB_ERR_INVALID, "dir must be -1 or +1");
    if (!in || !out) return fail(h, PB_ERR_INVALID, "null array pointer");
    if (in_kind < 0 || in_kind > 2 || out_kind < 0 || out_kind > 2) return fail(h, PB_ERR_INVALID, "bad in/out kind");
    if (in_ld < nx || out_ld < ox) return fail(h, PB_ERR_INVALID, "row pitch smaller than the row");
    Fft2Args a{dtype, in, in_kind, amp, amp_kind, kscale / (2.0 * M_PI), ny, nx, in_ld, ky, kx, dir, scale,
               shift_in, shift_out, out, out_kind, weight, oy, ox, out_ld};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = try_tuned_fft2(h, a.dtype, a.in, a.in_kind, a.amp, a.amp_kind, a.kturns, a.ny, a.nx, a.in_ld, a.ky,
                            a.kx, a.dir, a.scale, a.shift_in, a.shift_out, a.out, a.out_kind, a.weight, a.oy, a.ox,
                            a.out_ld, st);
    if (rc != PB_ERR_UNSUPPORTED) return rc;
    return fft2_generic(h, a, st);
}

extern "C" int pb_fft2_batch(pb_handle_t hh, int dtype, const void* in, int in_kind, const void* amp, int amp_kind,
                             double kscale, int batch, long long in_bs, long long amp_bs, int ny, int nx, long long in_ld,
                             int ky, int kx, int dir, double scale, int shift_in, int shift_out, void* out, int out_kind,
                             double weight, int oy, int ox, long long out_ld, long long out_bs, void* stream) {
    PB_ENTER(hh);
    if (batch < 1) return fail(h, PB_ERR_INVALID, "batch must be >= 1");
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || ky < ny || kx < nx) return fail(h, PB_ERR_INVALID, "need 1 <= n <= k on both axes");
    if (oy < 1 || ox < 1 || oy > ky || ox > kx) return fail(h, PB_ERR_INVALID, "crop window must fit the transform");
    if (dir != -1 && dir != 1) return fail(h, PB_ERR_INVALID, "dir must be -1 or +1");
    if (!in || !out) return fail(h, PB_ERR_INVALID, "null array pointer");
    if (in_kind < 0 || in_kind > 2 || out_kind < 0 || out_kind > 2) return fail(h, PB_ERR_INVALID, "bad in/out kind");
    if (in_ld < nx || out_ld < ox) return fail(h, PB_ERR_INVALID, "row pitch smaller than the row");
    if (batch > 1 && (in_bs < (long long)(ny - 1) * in_ld + nx || out_bs < (long long)(oy - 1) * out_ld + ox))
        return fail(h, PB_ERR_INVALID, "batch stride smaller than one field");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = try_tuned_fft2_batch(h, dtype, in, in_kind, amp, amp_kind, kscale / (2.0 * M_PI), batch, in_bs, amp_bs, ny, nx,
                                  in_ld, ky, kx, dir, scale, shift_in, shift_out, out, out_kind, weight, oy, ox, out_ld,
                                  out_bs, st);
    if (rc != PB_ERR_UNSUPPORTED) return rc;
    // shapes outside the fused kernels: one field at a time through the generic passes
    const size_t rsz = dtype == PB_C64 ? 4 : 8;
    const size_t in_elt = in_kind == PB_IN_COMPLEX ? 2 * rsz : rsz;
    const size_t amp_elt = amp_kind == PB_AMP_U8 ? 1 : rsz;
    const size_t out_elt = out_kind == PB_OUT_COMPLEX ? 2 * rsz : rsz;
    for (int b = 0; b < batch; ++b) {
        Fft2Args a{dtype, reinterpret_cast<const char*>(in) + (size_t)b * in_bs * in_elt, in_kind,
                   amp ? reinterpret_cast<const char*>(amp) + (size_t)b * amp_bs * amp_elt : nullptr, amp_kind,
                   kscale / (2.0 * M_PI), ny, nx, in_ld, ky, kx, dir, scale, shift_in, shift_out,
                   reinterpret_cast<char*>(out) + (size_t)b * out_bs * out_elt, out_kind, weight, oy, ox, out_ld};
        PB_TRY(fft2_generic(h, a, st));
    }
    return PB_OK;
}

extern "C" int pb_fft1(pb_handle_t hh, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int n,
                       int dir, double scale, void* out, long long out_ld, void* stream) {
    PB_ENTER(hh);
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || n < 1 || (axis != 0 && axis != 1)) return fail(h, PB_ERR_INVALID, "bad shape / axis");
    if (dir != -1 && dir != 1) return fail(h, PB_ERR_INVALID, "dir must be -1 or +1");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AxisPass p;
    p.dtype = dtype; p.in = in; p.dir = dir; p.scale = scale; p.out = out;
    p.Llog = n; p.in_off = 0; p.rot_in = 0; p.n_out = n; p.crop_off = 0; p.rot_out = 0;
    if (axis == 1) {
        p.nb = ny; p.ibs = in_ld; p.ies = 1; p.n_in = nx < n ? nx : n;
        p.obs = out_ld; p.oes = 1; p.batch_contiguous = 0;
    } else {
        p.nb = nx; p.ibs = 1; p.ies = in_ld; p.n_in = ny < n ? ny : n;
        p.obs = 1; p.oes = out_ld; p.batch_contiguous = 1;
    }
    return axis_dft(h, p, st);
}

extern "C" int pb_axis_dft(pb_handle_t hh, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int n,
                           int dir, double scale, const void* pre_e, int pre_e_conj, const void* pre_b, int pre_b_conj,
                           const void* post_e, int post_e_conj, const void* post_b, int post_b_conj, int out_off,
                           int n_out, void* out, long long out_ld, void* stream) {
    PB_ENTER(hh);
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || n < 1 || (axis != 0 && axis != 1)) return fail(h, PB_ERR_INVALID, "bad shape / axis");
    if (dir != -1 && dir != 1) return fail(h, PB_ERR_INVALID, "dir must be -1 or +1");
    if (out_off < 0 || n_out < 1 || out_off + n_out > n) return fail(h, PB_ERR_INVALID, "output window outside the transform");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AxisPass p;
    p.dtype = dtype; p.in = in; p.dir = dir; p.scale = scale; p.out = out;
    p.Llog = n; p.n_out = n_out; p.crop_off = out_off;
    p.pre_e = pre_e; p.pre_e_conj = pre_e_conj; p.pre_b = pre_b; p.pre_b_conj = pre_b_conj;
    p.post_e = post_e; p.post_e_conj = post_e_conj; p.post_off = out_off; p.post_b = post_b; p.post_b_conj = post_b_conj;
    if (axis == 1) {
        p.nb = ny; p.ibs = in_ld; p.ies = 1; p.n_in = nx < n ? nx : n;
        p.obs = out_ld; p.oes = 1; p.batch_contiguous = 0;
    } else {
        p.nb = nx; p.ibs = 1; p.ies = in_ld; p.n_in = ny < n ? ny : n;
        p.obs = 1; p.oes = out_ld; p.batch_contiguous = 1;
    }
    return axis_dft(h, p, st);
}

int pb::czt_axis_impl(Handle* h, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int K,
                         const void* pre_e, int pre_conj, const void* H, const void* post_e, int post_conj, int out_off,
                         int n_out, double scale, int out_kind, double weight, void* out, long long out_ld, void* stream) {
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || K < 1 || (axis != 0 && axis != 1) || !H) return fail(h, PB_ERR_INVALID, "bad czt arguments");
    if (out_off < 0 || n_out < 1 || out_off + n_out > K) return fail(h, PB_ERR_INVALID, "output window outside the transform");
    if (!is_pow2(K)) return fail(h, PB_ERR_INVALID, "the Bluestein length K must be a power of two");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nline = axis == 1 ? ny : nx, n_in = std::min(axis == 1 ? nx : ny, K);
    AxisPass f;  // fused: (in * pre) -> FFT_K -> * H -> IFFT_K -> window -> * post
    f.dtype = dtype; f.in = in; f.dir = -1; f.out = out; f.scale = scale / K;
    f.L = K; f.Llog = K; f.n_in = n_in; f.n_out = n_out; f.crop_off = out_off; f.nb = nline;
    f.pre_e = pre_e; f.pre_e_conj = pre_conj;
    f.post_e = H;
    f.post_e2 = post_e; f.post_e2_conj = post_conj;
    if (axis == 1) { f.ibs = in_ld; f.ies = 1; f.obs = out_ld; f.oes = 1; f.batch_contiguous = 0; }
    else { f.ibs = 1; f.ies = in_ld; f.obs = 1; f.oes = out_ld; f.batch_contiguous = 1; }
    f.roundtrip = 1;
    f.out_kind = out_kind; f.weight = weight;
    int rc = try_tuned_axis_pass(h, f, st);
    if (rc != PB_ERR_UNSUPPORTED) return rc;
    // two passes through a (lines x K) scratch
    void* tmp = nullptr;
    PB_TRY(ensure_scratch(h, 1, (size_t)nline * K * csize(dtype), &tmp));
    AxisPass a = f;
    a.roundtrip = 0; a.post_e2 = nullptr; a.out = tmp; a.n_out = K; a.crop_off = 0; a.scale = 1.0;
    a.out_kind = PB_OUT_COMPLEX; a.weight = 1.0;
    if (axis == 1) { a.obs = K; a.oes = 1; } else { a.obs = 1; a.oes = nline; }
    PB_TRY(axis_dft(h, a, st));
    AxisPass b;
    b.dtype = dtype; b.in = tmp; b.dir = +1; b.out = out; b.scale = scale / K;
    b.Llog = K; b.n_in = K; b.n_out = n_out; b.crop_off = out_off; b.nb = nline;
    b.post_e = post_e; b.post_e_conj = post_conj; b.post_off = out_off;
    b.ibs = a.obs; b.ies = a.oes; b.obs = f.obs; b.oes = f.oes; b.batch_contiguous = f.batch_contiguous;
    b.out_kind = out_kind; b.weight = weight;
    return axis_dft(h, b, st);
}

extern "C" int pb_czt_axis(pb_handle_t hh, int dtype, const void* in, int ny, int nx, long long in_ld, int axis, int K,
                           const void* pre_e, int pre_conj, const void* H, const void* post_e, int post_conj, int out_off,
                           int n_out, double scale, void* out, long long out_ld, void* stream) {
    PB_ENTER(hh);
    return czt_axis_impl(h, dtype, in, ny, nx, in_ld, axis, K, pre_e, pre_conj, H, post_e, post_conj, out_off, n_out, scale,
                         PB_OUT_COMPLEX, 1.0, out, out_ld, stream);
}

extern "C" int pb_czt_axis_intensity(pb_handle_t hh, int dtype, const void* in, int ny, int nx, long long in_ld, int axis,
                                     int K, const void* pre_e, int pre_conj, const void* H, const void* post_e,
                                     int post_conj, int out_off, int n_out, double scale, int out_kind, double weight,
                                     void* out, long long out_ld, void* stream) {
    PB_ENTER(hh);
    if (out_kind != PB_OUT_INTENSITY && out_kind != PB_OUT_ACCUMULATE) return fail(h, PB_ERR_INVALID, "out_kind must be intensity or accumulate");
    return czt_axis_impl(h, dtype, in, ny, nx, in_ld, axis, K, pre_e, pre_conj, H, post_e, post_conj, out_off, n_out, scale,
                         out_kind, weight, out, out_ld, stream);
}

static int angular_spectrum_impl(Handle* h, int dtype, const void* in, int ny, int nx, int ky, int kx, const void* ty,
                                 const void* tx, const void* tf, int conj_tf, const void* screen, int conj_screen, void* out,
                                 int oy, int ox, cudaStream_t st) {
    if (dtype != PB_C64 && dtype != PB_C128) return fail(h, PB_ERR_INVALID, "dtype must be PB_C64 or PB_C128");
    if (ny < 1 || nx < 1 || ky < ny || kx < nx || oy < 1 || ox < 1 || oy > ky || ox > kx)
        return fail(h, PB_ERR_INVALID, "bad shapes");
    if (!tf && (!ty || !tx)) return fail(h, PB_ERR_INVALID, "need tf or both ty and tx");

    const size_t cs = csize(dtype);
    void *t0 = nullptr, *t1 = nullptr;
    PB_TRY(ensure_scratch(h, 0, (size_t)ky * kx * cs, &t0));
    PB_TRY(ensure_scratch(h, 1, (size_t)ky * kx * cs, &t1));
    AxisPass p;  // rows forward (times the phase screen, if any): ny populated rows -> t0 (ny, kx)
    p.dtype = dtype; p.in = in; p.ibs = nx; p.ies = 1; p.nb = ny;
    p.pre_mat = screen; p.pmi_bs = nx; p.pmi_es = 1; p.pre_mat_conj = conj_screen;
    p.Llog = kx; p.n_in = nx; p.in_off = ceil_half(kx - nx);
    p.dir = -1; p.out = t0; p.obs = kx; p.oes = 1; p.n_out = kx;
    PB_TRY(axis_dft(h, p, st));
    if (!tf && dtype == PB_C64) {  // separable TF: forward * TF * inverse along y in ONE tuned pass, if it is covered
        AxisPass f;
        f.dtype = dtype; f.in = t0; f.ibs = 1; f.ies = kx; f.nb = kx;
        f.L = ky; f.Llog = ky; f.n_in = ny; f.in_off = ceil_half(ky - ny);
        f.dir = -1; f.out = t1; f.obs = 1; f.oes = kx; f.n_out = oy; f.crop_off = ceil_half(ky - oy); f.batch_contiguous = 1;
        f.post_e = ty; f.post_e_conj = conj_tf; f.post_b = tx; f.post_b_conj = conj_tf;
        f.roundtrip = 1;
        int frc = is_pow2(ky) ? try_tuned_axis_pass(h, f, st) : PB_ERR_UNSUPPORTED;
        if (frc == PB_OK) {
            AxisPass s2;  // rows inverse from t1 (oy, kx), keep the ox centred columns, 1/(ky*kx)
            s2.dtype = dtype; s2.in = t1; s2.ibs = kx; s2.ies = 1; s2.nb = oy;
            s2.Llog = kx; s2.n_in = kx; s2.dir = +1;
            s2.out = out; s2.obs = ox; s2.oes = 1; s2.n_out = ox; s2.crop_off = ceil_half(kx - ox);
            s2.scale = 1.0 / ((double)ky * (double)kx);
            return axis_dft(h, s2, st);
        }
        if (frc != PB_ERR_UNSUPPORTED) return frc;
    }
    AxisPass q;  // columns forward, times the transfer function -> t1 (ky, kx)
    q.dtype = dtype; q.in = t0; q.ibs = 1; q.ies = kx; q.nb = kx;
    q.Llog = ky; q.n_in = ny; q.in_off = ceil_half(ky - ny);
    q.dir = -1; q.out = t1; q.obs = 1; q.oes = kx; q.n_out = ky; q.batch_contiguous = 1;
    if (tf) { q.post_mat = tf; q.pm_bs = 1; q.pm_es = kx; q.pm_conj = conj_tf; }
    else { q.post_e = ty; q.post_e_conj = conj_tf; q.post_b = tx; q.post_b_conj = conj_tf; }
    PB_TRY(axis_dft(h, q, st));
    AxisPass r;  // columns inverse, keep the oy centred rows -> t0 (oy, kx)
    r.dtype = dtype; r.in = t1; r.ibs = 1; r.ies = kx; r.nb = kx;
    r.Llog = ky; r.n_in = ky; r.dir = +1;
    r.out = t0; r.obs = 1; r.oes = kx; r.n_out = oy; r.crop_off = ceil_half(ky - oy); r.batch_contiguous = 1;
    PB_TRY(axis_dft(h, r, st));
    AxisPass s;  // rows inverse, keep the ox centred columns, 1/(ky*kx)
    s.dtype = dtype; s.in = t0; s.ibs = kx; s.ies = 1; s.nb = oy;
    s.Llog = kx; s.n_in = kx; s.dir = +1;
    s.out = out; s.obs = ox; s.oes = 1; s.n_out = ox; s.crop_off = ceil_half(kx - ox);
    s.scale = 1.0 / ((double)ky * (double)kx);
    return axis_dft(h, s, st);
}

extern "C" int pb_angular_spectrum(pb_handle_t hh, int dtype, const void* in, int ny, int nx, int ky, int kx,
                                   const void* ty, const void* tx, const void* tf, int conj_tf, void* out, int oy,
                                   int ox, void* stream) {
    PB_ENTER(hh);
    return angular_spectrum_impl(h, dtype, in, ny, nx, ky, kx, ty, tx, tf, conj_tf, nullptr, 0, out, oy, ox,
                                 reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pb_angular_spectrum_screen(pb_handle_t hh, int dtype, const void* in, const void* screen, int conj_screen,
                                          int ny, int nx, int ky, int kx, const void* ty, const void* tx, const void* tf,
                                          int conj_tf, void* out, int oy, int ox, void* stream) {
    PB_ENTER(hh);
    if (!screen) return fail(h, PB_ERR_INVALID, "null screen");
    return angular_spectrum_impl(h, dtype, in, ny, nx, ky, kx, ty, tx, tf, conj_tf, screen, conj_screen, out, oy, ox,
                                 reinterpret_cast<cudaStream_t>(stream));
}
