// Tuned register-resident FFT kernels for the headline path: Wavefront.focus / unfocus with Q = 2
// on N x N complex64 pupils, N in {512, 1024, 2048}  (reference prysm/propagation/fft.py:7-65).
//
// Algebra.  K = 2N.  focus = fftshift(FFT2_K(ifftshift(pad(x)))).  Along one axis, with q = x
// zero-extended to K:   U[k] = (-dir*i)^k * Q[k]   (the centred pad + ifftshift is a circular
// shift by K/4), and the zero half makes the first radix-2 level trivial:
//      Q[2j]   = FFT_N(x)[j]                    ("even" half)
//      Q[2j+1] = FFT_N(x * w_K^n)[j]            ("odd" half)
// so every length-K transform of the padded data is two length-N transforms of the un-padded
// data, the shifts are sign patterns / index rotations at the store, and no zero is ever read,
// written or multiplied.
//
// Data flow (columns first so that the big K x K output is written by the row kernel in full
// 32 KB rows with 128-bit stores):
//   phase p in {even, odd} output rows:
//     focus_col_kernel : FFT_N down T adjacent columns of the pupil (x w_K^n for the odd phase)
//                        -> N x N intermediate (rows j  <->  output row 2j+p, rotated by K/2)
//     focus_row_kernel : per intermediate row, the two FFT_N (even / odd output columns) in one CTA,
//                        outputs interleaved so each thread stores 16 contiguous bytes
//   HBM traffic: pupil 8N^2 (+ a second read that mostly hits L2), output 8K^2; the N x N x 8 B
//   intermediate of each phase (32 MiB at N = 2048) is produced and consumed inside the 126 MB L2.
//
// FFT_N engine: N/16 threads per transform, 16 points per thread in registers, radix 16 x 16 x N/256,
// two shared-memory exchanges (1-in-16 padded, conflict free), twiddles from an L1-resident table.
#include <cstdlib>
#include "fft_tuned.cuh"

namespace pb {
namespace {

#define PB_SQRT1_2 0.70710678118654752440f
#define PB_C1_8 0.92387953251128675613f   // cos(pi/8)
#define PB_S1_8 0.38268343236508977173f   // sin(pi/8)

__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 fcmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by w4 = exp(-+ i pi/2):  -i forward, +i inverse
template <bool INV> __device__ __forceinline__ float2 mul_w4(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by w = (c, -+ s)
template <bool INV> __device__ __forceinline__ float2 mul_cs(float2 a, float c, float s) {
    return INV ? make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.y, c, a.x * s))
               : make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
}

template <bool INV> __device__ __forceinline__ void dft2(float2& a, float2& b) {
    float2 s = a + b, d = a - b;
    a = s; b = d;
}

template <bool INV> __device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = mul_w4<INV>(x1 - x3);
    x0 = s02 + s13; x2 = s02 - s13; x1 = d02 + d13; x3 = d02 - d13;
}

template <bool INV> __device__ __forceinline__ void dft8(float2* v) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    o1 = mul_cs<INV>(o1, PB_SQRT1_2, PB_SQRT1_2);
    o2 = mul_w4<INV>(o2);
    o3 = mul_cs<INV>(o3, -PB_SQRT1_2, PB_SQRT1_2);
    v[0] = e0 + o0; v[4] = e0 - o0;
    v[1] = e1 + o1; v[5] = e1 - o1;
    v[2] = e2 + o2; v[6] = e2 - o2;
    v[3] = e3 + o3; v[7] = e3 - o3;
}

template <bool INV> __device__ __forceinline__ void dft16(float2* v) {
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft8<INV>(e);
    dft8<INV>(o);
    o[1] = mul_cs<INV>(o[1], PB_C1_8, PB_S1_8);
    o[2] = mul_cs<INV>(o[2], PB_SQRT1_2, PB_SQRT1_2);
    o[3] = mul_cs<INV>(o[3], PB_S1_8, PB_C1_8);
    o[4] = mul_w4<INV>(o[4]);
    o[5] = mul_cs<INV>(o[5], -PB_S1_8, PB_C1_8);
    o[6] = mul_cs<INV>(o[6], -PB_SQRT1_2, PB_SQRT1_2);
    o[7] = mul_cs<INV>(o[7], -PB_C1_8, PB_S1_8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = e[i] + o[i]; v[i + 8] = e[i] - o[i]; }
}

template <int R, bool INV> __device__ __forceinline__ void dftR(float2* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else dft16<INV>(v);
}

__device__ __forceinline__ int pad16(int a) { return a + (a >> 4); }

struct SyncCta { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct SyncNamed {
    int id, count;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
};

// First two radix-16 stages of FFT_L (L >= 512).  In: v[n] = x[n*L/16 + t].  Out: the stage-2 results in S
// (1-in-16 padded, position-major) ready for the final radix-(L/256) stage:  item i in [0, 256) reads
// S[pad16(n*256 + i)], n < L/256, and after dft produces X[i + 256*k].   tw: w_TL^j table, TLS = TL / L.
template <int L, bool INV, int TLS, class Sync>
__device__ __forceinline__ void fft_two_stages(float2 (&v)[16], const int t, float2* __restrict__ S,
                                               const float2* __restrict__ tw, Sync sync) {
    constexpr int NT = L / 16;
    dft16<INV>(v);
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        float2 w = __ldg(tw + t * k * TLS);
        if (INV) w.y = -w.y;
        v[k] = fcmul(v[k], w);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) S[t * 17 + k] = v[k];
    sync();
#pragma unroll
    for (int n = 0; n < 16; ++n) v[n] = S[pad16(n * NT + t)];
    sync();
    dft16<INV>(v);
    const int m = t >> 4, a = t & 15;
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        float2 w = __ldg(tw + m * k * 16 * TLS);
        if (INV) w.y = -w.y;
        v[k] = fcmul(v[k], w);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) S[m * 272 + a + 17 * k] = v[k];
    sync();
}

template <int L> struct Geo {
    static constexpr int NT = L / 16;          // threads per transform
    static constexpr int R3 = L / 256;         // last radix
    static constexpr int SBUF = L + L / 16;    // padded exchange buffer, complex elements
};

struct FocusParams {
    // column kernel input
    const void* in; int in_kind; const void* amp; int amp_kind; double kturns; long long in_ld;
    float2* tmp;            // [phases][N][N] intermediate
    const float2* tw;       // w_K^j, K = 2N entries
    // row kernel output
    void* out; long long out_ld; int out_kind; float scale; float weight;
    int phase0;             // first phase handled by this launch (blockIdx.y adds to it)
};

// ---- column pass: FFT_N down T adjacent columns --------------------------------------------------
template <int L, bool INV, int T>
__global__ void __launch_bounds__(T * L / 16) focus_col_kernel(const FocusParams p) {
    using G = Geo<L>;
    extern __shared__ __align__(16) float2 smem[];
    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int col = blockIdx.x * T + c;
    const int phase = p.phase0 + blockIdx.y;
    float2 v[16];
    if (p.in_kind == PB_IN_COMPLEX) {
        const float2* __restrict__ src = reinterpret_cast<const float2*>(p.in) + col;
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = __ldg(src + (long long)(n * G::NT + t) * p.in_ld);
    } else {
        const float* __restrict__ opd = reinterpret_cast<const float*>(p.in) + col;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const long long off = (long long)(n * G::NT + t) * p.in_ld;
            float a = 1.0f;
            if (p.amp_kind == PB_AMP_REAL) a = __ldg(reinterpret_cast<const float*>(p.amp) + col + off);
            else if (p.amp_kind == PB_AMP_U8) a = __ldg(reinterpret_cast<const unsigned char*>(p.amp) + col + off) ? 1.0f : 0.0f;
            float2 e = make_float2(0.f, 0.f);
            if (a != 0.0f) {
                e = expi_turns(p.kturns * (double)__ldg(opd + off), 0.0f);
                e.x *= a; e.y *= a;
            }
            v[n] = e;
        }
    }
    if (phase) {  // odd output rows: modulate by w_K^n
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            float2 w = __ldg(p.tw + n * G::NT + t);
            if (INV) w.y = -w.y;
            v[n] = fcmul(v[n], w);
        }
    }
    float2* S = smem + c * (G::SBUF + 4);  // +4: skews the T buffers across banks
    fft_two_stages<L, INV, 2>(v, t, S, p.tw, SyncCta());
    constexpr int R3 = G::R3, GI = 16 / R3;
#pragma unroll
    for (int g = 0; g < GI; ++g)
#pragma unroll
        for (int n = 0; n < R3; ++n) v[g * R3 + n] = S[pad16(n * 256 + t + g * G::NT)];
#pragma unroll
    for (int g = 0; g < GI; ++g) dftR<R3, INV>(v + g * R3);
    // (-dir*i)^k with k = 2j + phase; j = t + g*NT + 256*kk has the parity of t
    const float sgn = (t & 1) ? -1.0f : 1.0f;
    float2* __restrict__ dst = p.tmp + (long long)blockIdx.y * L * L + col;
#pragma unroll
    for (int g = 0; g < GI; ++g)
#pragma unroll
        for (int kk = 0; kk < R3; ++kk) {
            const int j = t + g * G::NT + 256 * kk;
            float2 x = v[g * R3 + kk];
            if (phase) x = INV ? make_float2(x.y, -x.x) : make_float2(-x.y, x.x);  // * (-dir*i)
            dst[(long long)j * L] = make_float2(sgn * x.x, sgn * x.y);
        }
}

// ---- row pass: both half-transforms of one intermediate row, interleaved 128-bit stores -----------
template <int L, bool INV>
__global__ void __launch_bounds__(L / 8) focus_row_kernel(const FocusParams p) {
    using G = Geo<L>;
    constexpr int NT = G::NT, NTH = 2 * NT, R3 = G::R3, GI = 2048 / L;
    extern __shared__ __align__(16) float2 smem[];
    const int tid = threadIdx.x;
    const int f = tid / NT, t = tid - f * NT;
    const int r = blockIdx.x;               // intermediate row (compact)
    const int phase = p.phase0 + blockIdx.y;
    const float2* __restrict__ src = p.tmp + (long long)blockIdx.y * L * L + (long long)r * L;
    float2 v[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) v[n] = __ldg(src + n * NT + t);
    if (f) {
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            float2 w = __ldg(p.tw + n * NT + t);
            if (INV) w.y = -w.y;
            v[n] = fcmul(v[n], w);
        }
    }
    fft_two_stages<L, INV, 2>(v, t, smem + f * G::SBUF, p.tw, SyncNamed{1 + f, NT});
    __syncthreads();
    // merged last stage: thread owns items i = tid + g*NTH of BOTH halves -> adjacent outputs 2j, 2j+1
#pragma unroll
    for (int g = 0; g < GI; ++g)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int n = 0; n < R3; ++n) v[(g * 2 + h) * R3 + n] = smem[h * G::SBUF + pad16(n * 256 + tid + g * NTH)];
#pragma unroll
    for (int q = 0; q < 2 * GI; ++q) dftR<R3, INV>(v + q * R3);
    const float sgn = ((tid & 1) ? -1.0f : 1.0f) * p.scale;
    const int orow = (2 * r + phase + L) & (2 * L - 1);      // fftshift along y
    if (p.out_kind == PB_OUT_COMPLEX) {
        float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(p.out) + (long long)orow * p.out_ld);
#pragma unroll
        for (int g = 0; g < GI; ++g)
#pragma unroll
            for (int kk = 0; kk < R3; ++kk) {
                const int j = tid + g * NTH + 256 * kk;
                const float2 A = v[(g * 2 + 0) * R3 + kk], B = v[(g * 2 + 1) * R3 + kk];
                // U[2j] = (-1)^j A,  U[2j+1] = (-1)^j (-dir*i) B ; fftshift along x: pair index (j + L/2) mod L
                const float2 Bm = INV ? make_float2(B.y, -B.x) : make_float2(-B.y, B.x);
                __stcs(dst + ((j + L / 2) & (L - 1)), make_float4(sgn * A.x, sgn * A.y, sgn * Bm.x, sgn * Bm.y));
            }
    } else {
        float2* __restrict__ dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + (long long)orow * p.out_ld);
        const float s2 = p.scale * p.scale;
#pragma unroll
        for (int g = 0; g < GI; ++g)
#pragma unroll
            for (int kk = 0; kk < R3; ++kk) {
                const int j = tid + g * NTH + 256 * kk;
                const float2 A = v[(g * 2 + 0) * R3 + kk], B = v[(g * 2 + 1) * R3 + kk];
                float2 I = make_float2(s2 * (A.x * A.x + A.y * A.y), s2 * (B.x * B.x + B.y * B.y));
                float2* q = dst + ((j + L / 2) & (L - 1));
                if (p.out_kind == PB_OUT_ACCUMULATE) {
                    const float2 old = *q;
                    I = make_float2(fmaf(p.weight, I.x, old.x), fmaf(p.weight, I.y, old.y));
                }
                *q = I;
            }
    }
}

template <int L, bool INV>
int launch_focus(Handle* h, FocusParams p, cudaStream_t st) {
    using G = Geo<L>;
    constexpr int T = 4;
    const size_t smem_col = (size_t)T * (G::SBUF + 4) * sizeof(float2);
    const size_t smem_row = (size_t)2 * G::SBUF * sizeof(float2);
    static bool attr_done = false;
    if (!attr_done) {
        PB_CUDA(h, cudaFuncSetAttribute(focus_col_kernel<L, INV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
        PB_CUDA(h, cudaFuncSetAttribute(focus_row_kernel<L, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
        attr_done = true;
    }
    static const int mode = [] { const char* e = getenv("PB_FOCUS_PHASES_PER_LAUNCH"); return e ? atoi(e) : 1; }();
    if (mode == 2) {  // both phases per launch: 2 launches, 2*N*N intermediate
        p.phase0 = 0;
        focus_col_kernel<L, INV, T><<<dim3(L / T, 2), T * G::NT, smem_col, st>>>(p);
        PB_LAUNCH_CHECK(h);
        focus_row_kernel<L, INV><<<dim3(L, 2), L / 8, smem_row, st>>>(p);
        PB_LAUNCH_CHECK(h);
    } else {          // one phase at a time: 4 launches, N*N intermediate stays in L2
        for (int ph = 0; ph < 2; ++ph) {
            p.phase0 = ph;
            focus_col_kernel<L, INV, T><<<dim3(L / T, 1), T * G::NT, smem_col, st>>>(p);
            PB_LAUNCH_CHECK(h);
            focus_row_kernel<L, INV><<<dim3(L, 1), L / 8, smem_row, st>>>(p);
            PB_LAUNCH_CHECK(h);
        }
    }
    return PB_OK;
}

}  // namespace

int try_tuned_axis_pass(Handle*, const AxisPass&, cudaStream_t) { return PB_ERR_UNSUPPORTED; }

int try_tuned_fft2(Handle* h, int dtype, const void* in, int in_kind, const void* amp, int amp_kind, double kturns,
                   int ny, int nx, long long in_ld, int ky, int kx, int dir, double scale, int shift_in, int shift_out,
                   void* out, int out_kind, double weight, int oy, int ox, long long out_ld, cudaStream_t st) {
    static const bool disabled = getenv("PB_DISABLE_TUNED") != nullptr;
    if (disabled) return PB_ERR_UNSUPPORTED;
    if (dtype != PB_C64 || ny != nx || ky != 2 * ny || kx != 2 * nx || !shift_in || !shift_out || oy != ky || ox != kx)
        return PB_ERR_UNSUPPORTED;
    if (nx != 512 && nx != 1024 && nx != 2048) return PB_ERR_UNSUPPORTED;
    if (in_kind == PB_IN_REAL) return PB_ERR_UNSUPPORTED;
    if (out_kind == PB_OUT_COMPLEX ? (out_ld & 1) || ((uintptr_t)out & 15) : (out_ld & 1) || ((uintptr_t)out & 7))
        return PB_ERR_UNSUPPORTED;  // vector stores need aligned rows
    const int N = nx;
    FocusParams p;
    p.in = in; p.in_kind = in_kind; p.amp = amp; p.amp_kind = amp_kind; p.kturns = kturns; p.in_ld = in_ld;
    void* tmp = nullptr;
    PB_TRY(ensure_scratch(h, 0, (size_t)2 * N * N * sizeof(float2), &tmp));
    p.tmp = reinterpret_cast<float2*>(tmp);
    const void* tw = nullptr;
    PB_TRY(get_twiddles(h, 2 * N, PB_C64, &tw));
    p.tw = reinterpret_cast<const float2*>(tw);
    p.out = out; p.out_ld = out_ld; p.out_kind = out_kind; p.scale = (float)scale; p.weight = (float)weight;
    p.phase0 = 0;
    if (dir < 0) {
        if (N == 512) return launch_focus<512, false>(h, p, st);
        if (N == 1024) return launch_focus<1024, false>(h, p, st);
        return launch_focus<2048, false>(h, p, st);
    }
    if (N == 512) return launch_focus<512, true>(h, p, st);
    if (N == 1024) return launch_focus<1024, true>(h, p, st);
    return launch_focus<2048, true>(h, p, st);
}

int try_tuned_angular_spectrum(Handle*, int, const void*, int, int, int, int, const void*, const void*, const void*,
                               int, void*, int, int, cudaStream_t) {
    return PB_ERR_UNSUPPORTED;
}

}  // namespace pb
