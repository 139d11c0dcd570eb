// Tuned register-resident FFT kernels for the headline path: Wavefront.focus / unfocus with Q = 2
// on N x N complex64 pupils, N in {512, 1024, 2048}  (reference prysm/propagation/fft.py:7-65).
//
// Algebra.  K = 2N.  focus = fftshift(FFT2_K(ifftshift(pad(x)))).  Along one axis, with q = x
// zero-extended to K:   U[k] = (-dir*i)^k * Q[k]   (the centred pad + ifftshift is a circular
// shift by K/4), and the zero half makes the first radix-2 level trivial:
//      Q[2j]   = FFT_N(x)[j]                    ("even" half, lane A)
//      Q[2j+1] = FFT_N(x * w_K^n)[j]            ("odd" half,  lane B)
// so every length-K transform of the padded data is two length-N transforms of the un-padded
// data, the shifts are sign patterns / index rotations at the store, and no zero is ever read,
// written or multiplied.
//
// Packed arithmetic.  sm_100 issues two fp32 operations per instruction on 64-bit register pairs
// (FADD2 / FMUL2 / FFMA2).  Each thread therefore carries BOTH half-transforms of its line in
// structure-of-arrays form -- re = (re_A, re_B), im = (im_A, im_B) -- and every butterfly, twiddle
// multiply and store acts on the pair: the instruction count of the FFT is halved and the kernels
// move from issue-bound to FP32-pipe-bound.  The two lanes differ only in their inputs
// (lane B = x * w_32^n, immediates) and in the stage-1 twiddle (w_K^t folded into lane B's table entry).
//
// Data flow (columns first so that the K x K output is written by the row kernel in full rows with
// 128-bit stores):
//   focus_col_kernel : FFT_N down T adjacent columns of the pupil, both halves ->
//                      intermediate [2][N][N] (plane p, row j  <->  output row 2j+p, rotated by K/2)
//   focus_row_kernel : persistent; per intermediate row both half-transforms; thread owns outputs
//                      (2j, 2j+1) and stores them as one float4.  Rows arrive by bulk async copy
//                      (cp.async.bulk + mbarrier), one row ahead.
//   HBM traffic: pupil 8N^2 once, output 8K^2 once; the 16N^2-byte intermediate (64 MiB at N = 2048)
//   is written and re-read through the 126 MB L2.
//
// FFT_N engine: N/16 threads per line, 16 points x 2 lanes per thread in registers, radix
// 16 x 16 x N/256, two shared-memory exchanges of float4 (1-in-16 padded, conflict free).
#include <cstdlib>
#include "fft_tuned.cuh"

namespace pb {
namespace {

#define PB_SQRT1_2 0.70710678118654752440f
#define PB_C1_8 0.92387953251128675613f   // cos(pi/8)
#define PB_S1_8 0.38268343236508977173f   // sin(pi/8)

// two complex numbers (lane A, lane B), structure of arrays
struct P2 { float2 re, im; };

__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P2 operator+(P2 a, P2 b) { return {__fadd2_rn(a.re, b.re), __fadd2_rn(a.im, b.im)}; }
__device__ __forceinline__ P2 operator-(P2 a, P2 b) { return {__fadd2_rn(a.re, neg2(b.re)), __fadd2_rn(a.im, neg2(b.im))}; }
// multiply by w4 = exp(-+ i pi/2):  -i forward, +i inverse
template <bool INV> __device__ __forceinline__ P2 mul_w4(P2 a) { return INV ? P2{neg2(a.im), a.re} : P2{a.im, neg2(a.re)}; }
// multiply both lanes by w = (c, -+ s)
template <bool INV> __device__ __forceinline__ P2 mul_cs(P2 a, float c, float s) {
    const float2 C = make_float2(c, c), S = make_float2(s, s);
    if (INV) return {__ffma2_rn(a.re, C, neg2(__fmul2_rn(a.im, S))), __ffma2_rn(a.im, C, __fmul2_rn(a.re, S))};
    return {__ffma2_rn(a.re, C, __fmul2_rn(a.im, S)), __ffma2_rn(a.im, C, neg2(__fmul2_rn(a.re, S)))};
}
// per-lane twiddle: w = (wre, wim) pairs as stored in the tables (forward sign); INV conjugates
template <bool INV> __device__ __forceinline__ P2 mul_tw(P2 a, float4 w) {
    const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
    if (INV) return {__ffma2_rn(a.re, wr, __fmul2_rn(a.im, wi)), __ffma2_rn(a.im, wr, neg2(__fmul2_rn(a.re, wi)))};
    return {__ffma2_rn(a.re, wr, neg2(__fmul2_rn(a.im, wi))), __ffma2_rn(a.im, wr, __fmul2_rn(a.re, wi))};
}

template <bool INV> __device__ __forceinline__ void dft2(P2& a, P2& b) {
    P2 s = a + b, d = a - b;
    a = s; b = d;
}
template <bool INV> __device__ __forceinline__ void dft4(P2& x0, P2& x1, P2& x2, P2& x3) {
    P2 s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = mul_w4<INV>(x1 - x3);
    x0 = s02 + s13; x2 = s02 - s13; x1 = d02 + d13; x3 = d02 - d13;
}
template <bool INV> __device__ __forceinline__ void dft8(P2* v) {
    P2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    P2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
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
template <bool INV> __device__ __forceinline__ void dft16(P2* v) {
    P2 e[8], o[8];
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
template <int R, bool INV> __device__ __forceinline__ void dftR(P2* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else dft16<INV>(v);
}

__device__ __forceinline__ float4 pack(P2 a) { return make_float4(a.re.x, a.re.y, a.im.x, a.im.y); }
__device__ __forceinline__ P2 unpack(float4 a) { return {make_float2(a.x, a.y), make_float2(a.z, a.w)}; }

// lane A = x, lane B = x * w_32^n (the n-dependent part of the odd half's input ramp), compile-time n
template <bool INV, int n> __device__ __forceinline__ P2 make_lanes(float2 x) {
    // w_32^n = (cos(2 pi n/32), -sin(2 pi n/32))
    constexpr float C[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                             0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                             -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                             -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float Sn[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                              0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                              0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,
                              0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                              0.38268343236508977173f, 0.19509032201612826785f};
    const float c = C[n], s = INV ? -Sn[n] : Sn[n];          // x * (c - i s)
    float2 b;
    if (n == 0) b = x;
    else if (n == 8) b = make_float2(x.y * s, -x.x * s);     // c = 0
    else b = make_float2(fmaf(x.x, c, x.y * s), fmaf(x.y, c, -x.x * s));
    return {make_float2(x.x, b.x), make_float2(x.y, b.y)};
}

__device__ __forceinline__ int pad16(int a) { return a + (a >> 4); }

// ---- cache-policy loads / stores --------------------------------------------------------------------
// The only data with reuse inside an SM is the twiddle plan: it is pinned in L1 (evict_last) and every
// streaming access bypasses L1 allocation so that it cannot displace the plan.
__device__ __forceinline__ float4 ld_plan(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---- async-copy plumbing (cp.async.bulk = the 1-D TMA path; completion on an mbarrier) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- tensor TMA (cp.async.bulk.tensor): strided gathers of 32-byte runs into contiguous shared memory ----
__device__ __forceinline__ uint64_t l2_policy(int kind) {   // 0 evict_normal, 1 evict_first, 2 evict_last
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {   // into L2 only
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_hint(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float2* p, float2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

template <int L> struct Geo {
    static constexpr int NT = L / 16;          // threads per line
    static constexpr int R3 = L / 256;         // last radix
    static constexpr int GI = 16 / R3;         // last-stage items per thread
    static constexpr int SBUF = L + L / 16;    // padded exchange buffer, float4 elements
    // plan table layout (float4 elements):
    //   TW1[k][t], k < 16 : (w_L^(tk).re, (w_L^(tk) w_2L^t).re, w_L^(tk).im, (w_L^(tk) w_2L^t).im)
    //   TW2[k-1][m], 1 <= k < 16, m < NT/16 : (wr, wr, wi, wi), w = w_L^(16 m k)
    static constexpr int TW1 = 0, TW2 = 16 * NT, PLAN = 16 * NT + 15 * (NT / 16);
};

// First two radix-16 stages on both lanes.  In: v[n] = lanes of x[n*NT + t].  Out: stage-2 results in S
// (float4, 1-in-16 padded, position-major): final-stage item i in [0, 256) reads S[pad16(n*256 + i)],
// n < L/256, and after the radix-(L/256) dft holds X[i + 256*k].
// Twiddle sources: 0 = float4 plan in global memory (L1 evict_last), 1 = float4 plan in shared memory,
// 2 = "plain" float2 tables in global memory, the same twiddle for both lanes (k = 1..15 only).
template <int KIND> __device__ __forceinline__ float4 load_tw(const void* base, int idx) {
    if (KIND == 1) return reinterpret_cast<const float4*>(base)[idx];
    if (KIND == 0) return ld_plan(reinterpret_cast<const float4*>(base) + idx);
    float2 w;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(w.x), "=f"(w.y) : "l"(reinterpret_cast<const float2*>(base) + idx));
    return make_float4(w.x, w.x, w.y, w.y);
}

template <int L, bool INV, int KIND, class Sync1, class Sync>
__device__ __forceinline__ void fft_two_stages(P2 (&v)[16], const int t, float4* __restrict__ S,
                                               const void* __restrict__ tw1, const void* __restrict__ tw2, Sync1 sync1,
                                               Sync sync) {
    using G = Geo<L>;
    constexpr int NT = G::NT;
    dft16<INV>(v);
    if (KIND == 2) {  // tw1[(k-1)*NT + t]
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = mul_tw<INV>(v[k], load_tw<KIND>(tw1, (k - 1) * NT + t));
    } else {          // tw1[k*NT + t], k = 0 carries lane B's w_2L^t
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = mul_tw<INV>(v[k], load_tw<KIND>(tw1, k * NT + t));
    }
    {
        float4* __restrict__ d = S + t * 17;
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = pack(v[k]);
    }
    sync1();
    {
        const float4* __restrict__ s = S + pad16(t);
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = unpack(s[n * (NT + NT / 16)]);
    }
    sync();
    dft16<INV>(v);
    const int m = t >> 4, a = t & 15;
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = mul_tw<INV>(v[k], load_tw<KIND>(tw2, (k - 1) * (NT / 16) + m));
    {
        float4* __restrict__ d = S + m * 272 + a;
#pragma unroll
        for (int k = 0; k < 16; ++k) d[17 * k] = pack(v[k]);
    }
    sync();
}

// final radix-(L/256) stage inputs: afterwards v[g*R3 + n] holds the n-th input of item t + g*NT
template <int L, bool INV>
__device__ __forceinline__ void fft_last_stage_load(P2 (&v)[16], const int t, const float4* __restrict__ S) {
    using G = Geo<L>;
    const float4* __restrict__ s = S + pad16(t);
#pragma unroll
    for (int g = 0; g < G::GI; ++g)
#pragma unroll
        for (int n = 0; n < G::R3; ++n) v[g * G::R3 + n] = unpack(s[n * 272 + g * (G::NT + G::NT / 16)]);
}

#define PB_MAKE_LANES_16(v, INV, X)                                                                         \
    v[0] = make_lanes<INV, 0>(X(0));   v[1] = make_lanes<INV, 1>(X(1));   v[2] = make_lanes<INV, 2>(X(2));     \
    v[3] = make_lanes<INV, 3>(X(3));   v[4] = make_lanes<INV, 4>(X(4));   v[5] = make_lanes<INV, 5>(X(5));     \
    v[6] = make_lanes<INV, 6>(X(6));   v[7] = make_lanes<INV, 7>(X(7));   v[8] = make_lanes<INV, 8>(X(8));     \
    v[9] = make_lanes<INV, 9>(X(9));   v[10] = make_lanes<INV, 10>(X(10)); v[11] = make_lanes<INV, 11>(X(11)); \
    v[12] = make_lanes<INV, 12>(X(12)); v[13] = make_lanes<INV, 13>(X(13)); v[14] = make_lanes<INV, 14>(X(14)); \
    v[15] = make_lanes<INV, 15>(X(15));

struct SyncCta { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

struct FocusParams {
    // column kernel input
    const void* in; int in_kind; const void* amp; int amp_kind; double kturns; long long in_ld;
    float2* tmp;            // [2][N][N] intermediate (plane 0: even output rows, plane 1: odd)
    const float4* plan;     // Geo<N> plan table
    // row kernel output
    void* out; long long out_ld; int out_kind; float scale; float weight;
    int nrows;              // row kernel: batch * 2N
    // batch of independent fields in one launch pair (blockIdx.y in the column kernel; rows of all fields in the row
    // kernel's persistent loop): element strides of in / amp / out in units of their own scalar type
    long long in_bs, amp_bs, out_bs;
    // v2 pipeline
    int ntiles;             // column kernel: batch * N / T tiles, walked by a persistent grid
    int hints;              // 1: L2 eviction hints (intermediate evict_last when written / evict_first when read, streams evict_first)
};

// ---- column pass: both half-transforms down T adjacent columns -------------------------------------
template <int L, bool INV, int T>
__global__ void __launch_bounds__(T * L / 16) focus_col_kernel(const FocusParams p) {
    using G = Geo<L>;
    extern __shared__ __align__(128) float4 smem4[];
    float4* plan_s = smem4 + T * (G::SBUF + 2);                       // the twiddle plan, bulk-copied once
    uint64_t* bar = reinterpret_cast<uint64_t*>(plan_s + G::PLAN);
    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int col = blockIdx.x * T + c;
    const long long fb = blockIdx.y;                                   // field of the batch
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, G::PLAN * sizeof(float4));
        bulk_g2s(plan_s, p.plan, G::PLAN * sizeof(float4), bar);
    }
    __syncthreads();
    P2 v[16];
    float2 x[16];
    if (p.in_kind == PB_IN_COMPLEX) {
        const float2* __restrict__ src = reinterpret_cast<const float2*>(p.in) + fb * p.in_bs + col + (long long)t * p.in_ld;
        const long long step = (long long)G::NT * p.in_ld;
#pragma unroll
        for (int n = 0; n < 16; ++n) x[n] = ld_stream(src + n * step);
    } else {
        const float* __restrict__ opd = reinterpret_cast<const float*>(p.in) + fb * p.in_bs + col;
        const long long ab = fb * p.amp_bs + col;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const long long off = (long long)(n * G::NT + t) * p.in_ld;
            float a = 1.0f;
            if (p.amp_kind == PB_AMP_REAL) a = __ldg(reinterpret_cast<const float*>(p.amp) + ab + off);
            else if (p.amp_kind == PB_AMP_U8) a = __ldg(reinterpret_cast<const unsigned char*>(p.amp) + ab + off) ? 1.0f : 0.0f;
            float2 e = make_float2(0.f, 0.f);
            if (a != 0.0f) {
                e = expi_turns(p.kturns * (double)__ldg(opd + off), 0.0f);
                e.x *= a; e.y *= a;
            }
            x[n] = e;
        }
    }
#define PB_X(n) x[n]
    PB_MAKE_LANES_16(v, INV, PB_X)
#undef PB_X
    float4* S = smem4 + c * (G::SBUF + 2);  // +2: skews the T buffers across banks
    mbar_wait(bar, 0);
    fft_two_stages<L, INV, 1>(v, t, S, plan_s + G::TW1, plan_s + G::TW2, SyncCta(), SyncCta());
    fft_last_stage_load<L, INV>(v, t, S);
#pragma unroll
    for (int g = 0; g < G::GI; ++g) dftR<G::R3, INV>(v + g * G::R3);
    // (-dir*i)^k with k = 2j + lane; j = t + g*NT + 256*kk has the parity of t
    const float sgn = (t & 1) ? -1.0f : 1.0f;
    float2* __restrict__ dA = p.tmp + fb * (2LL * L * L) + col + (long long)t * L;
    float2* __restrict__ dB = dA + (long long)L * L;
#pragma unroll
    for (int g = 0; g < G::GI; ++g)
#pragma unroll
        for (int kk = 0; kk < G::R3; ++kk) {
            const P2 y = v[g * G::R3 + kk];
            const long long o = (long long)(g * G::NT + 256 * kk) * L;
            st_stream(dA + o, make_float2(sgn * y.re.x, sgn * y.im.x));
            // lane B times (-dir*i): forward (+i): (-im, re); inverse (-i): (im, -re)
            st_stream(dB + o, INV ? make_float2(sgn * y.im.y, -sgn * y.re.y) : make_float2(-sgn * y.im.y, sgn * y.re.y));
        }
}

// ---- row pass: persistent, both half-transforms of one intermediate row per iteration ---------------
template <int L, bool INV>
__global__ void __launch_bounds__(L / 16) focus_row_kernel(const FocusParams p) {
    using G = Geo<L>;
    constexpr int NT = G::NT, R3 = G::R3, GI = G::GI;
    constexpr uint32_t ROW_BYTES = L * sizeof(float2);
    extern __shared__ __align__(128) unsigned char smraw[];
    float4* S = reinterpret_cast<float4*>(smraw);                // SBUF exchange
    float2* inb = reinterpret_cast<float2*>(S + G::SBUF);        // L-element input row
    uint64_t* bar = reinterpret_cast<uint64_t*>(inb + L);        // 1 mbarrier
    const int t = threadIdx.x;
    int r = blockIdx.x;
    if (t == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        if (r < p.nrows) {
            mbar_expect_tx(bar, ROW_BYTES);
            bulk_g2s(inb, p.tmp + (long long)r * L, ROW_BYTES, bar);
        }
    }
    __syncthreads();
    const float sgn = ((t & 1) ? -1.0f : 1.0f) * p.scale;
    for (int it = 0; r < p.nrows; ++it, r += gridDim.x) {
        P2 v[16];
        mbar_wait(bar, it & 1);
        {
            const float2* __restrict__ row = inb + t;
#define PB_X(n) row[(n) * NT]
            PB_MAKE_LANES_16(v, INV, PB_X)
#undef PB_X
        }
        const int rn = r + gridDim.x;
        auto sync_and_prefetch = [&]() {
            __syncthreads();  // every thread holds its inputs in registers: the row buffer is free
            if (t == 0 && rn < p.nrows) {
                mbar_expect_tx(bar, ROW_BYTES);
                bulk_g2s(inb, p.tmp + (long long)rn * L, ROW_BYTES, bar);
            }
        };
        fft_two_stages<L, INV, 0>(v, t, S, p.plan + G::TW1, p.plan + G::TW2, sync_and_prefetch, SyncCta());
        fft_last_stage_load<L, INV>(v, t, S);
        __syncthreads();  // exchange buffer is free for the next row
#pragma unroll
        for (int g = 0; g < GI; ++g) dftR<R3, INV>(v + g * R3);
        const int fb = r / (2 * L), rf = r - fb * (2 * L);   // field of the batch, row within its [2][L] planes
        const int phs = rf / L, rr = rf - phs * L;
        const int orow = (2 * rr + phs + L) & (2 * L - 1);  // fftshift along y
        if (p.out_kind == PB_OUT_COMPLEX) {
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    // U[2j] = (-1)^j A,  U[2j+1] = (-1)^j (-dir*i) B ; fftshift along x: pair index (j + L/2) mod L
                    const float4 o = INV ? make_float4(sgn * y.re.x, sgn * y.im.x, sgn * y.im.y, -sgn * y.re.y)
                                         : make_float4(sgn * y.re.x, sgn * y.im.x, -sgn * y.im.y, sgn * y.re.y);
                    st_stream(dst + ((j + L / 2) & (L - 1)), o);
                }
        } else {
            float2* __restrict__ dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
            const float2 s2 = make_float2(p.scale * p.scale, p.scale * p.scale);
            const float2 wgt = make_float2(p.weight, p.weight);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    float2 I = __fmul2_rn(s2, __ffma2_rn(y.re, y.re, __fmul2_rn(y.im, y.im)));
                    float2* q = dst + ((j + L / 2) & (L - 1));
                    if (p.out_kind == PB_OUT_ACCUMULATE) I = __ffma2_rn(wgt, I, *q);
                    *q = I;
                }
        }
    }
}

// =====================================================================================================
// v2 pipeline: same arithmetic, different traffic.
//   * The intermediate is TILE-MAJOR: [field][plane p][j / 4][column tile = col / 4][j % 4][col % 4] complex64, i.e. the
//     4 x 4 block (4 output rows of one plane, 4 adjacent columns) is one 128-byte line.  A warp of the column kernel
//     (8 consecutive j x 4 columns) stores two full lines per instruction instead of eight 32-byte sectors in eight
//     lines -- the LSU pressure (`lg_throttle`) of the row-major layout is gone.
//   * The column kernel is persistent (one 512-thread CTA per SM walks the tiles of all fields of the launch) and takes
//     its 4-column input tile by tensor TMA (box = 4 columns x 256 rows, 32-byte runs gathered into a contiguous
//     [row][4] image that aliases the exchange buffers); the tile after next is requested as soon as the exchange
//     buffers are free, so it lands while the last butterflies and the stores of the current tile run.
//   * The row kernel gathers one intermediate row (512 runs of 32 bytes, 128 bytes apart) with two tensor-TMA boxes into
//     the same contiguous row buffer the v1 kernel fills by a bulk copy; everything after that is the v1 code.
// =====================================================================================================
template <int L> struct Geo2 {
    static constexpr int T = 4;                      // columns per tile
    static constexpr int TILES = L / T;              // tiles per field
    static constexpr int BOXR = 256;                 // rows per input box
    static constexpr int NBOX_IN = L / BOXR;
    static constexpr int BOXT = (TILES < 256 ? TILES : 256);   // tiles per row-gather box
    static constexpr int NBOX_ROW = TILES / BOXT;
};

// Column kernel of the v2 pipeline.  One CTA per SM, 16 compute warps + 1 producer warp:
//   * two independent 256-thread GROUPS (named barriers, own exchange buffers): group h computes half-transform h
//     (h = 0: FFT_N(x) -> plane 0; h = 1: FFT_N(x w_K^n) (-dir i) -> plane 1) of the CTA's 4-column tiles.  A thread's
//     two packed lanes are two ADJACENT COLUMNS, so both lanes use the same twiddle, the tile image is read with
//     128-bit shared-memory loads and the tile-major stores are float4 (two columns): four full 128-byte lines per
//     warp instruction.  The groups drift apart by up to one tile, so the exchange phases of one overlap the butterfly
//     phases of the other (the v1 column kernel ran one 16-warp line group in lock step: FP32 pipe 25 %).
//   * the 4-column x L-row input tile (64 KB) is gathered ONCE for both halves by tensor TMA (box = 4 columns x 256
//     rows) into a staging area of its own; the last of the 16 warps to have the inputs of tile i in registers (a
//     shared-memory counter) requests tile i+1 on the spot -- most of a tile time of lead (the gather of 2048 32-byte
//     runs takes ~2-3 us).
//   * one twiddle plan for both halves (float2 entries, 17 KB): half 1's stage-1 twiddles are the plan's times the
//     per-thread constant w_2L^t (its input ramp x w_K^n = x w_32^(n/NT) w_2L^t; the w_32 part is a compile-time constant).
template <int L> struct Geo4 {
    using G = Geo<L>;
    static constexpr int XBUF = G::SBUF + 4;                 // +4 float4 = 16 words: adjacent buffers land on complementary banks
    static constexpr int STAGE = 4 * XBUF;                   // float4 offset of the staging area (multiple of 8: 128-byte aligned)
    static constexpr int PLAN2 = STAGE + 2 * L;              // float4 offset of the float2 plan
    static constexpr int TW1 = 0, TW2 = 16 * G::NT, PLANLEN = 16 * G::NT + 15 * (G::NT / 16);   // float2 elements
    static constexpr size_t BAR_OFF = (size_t)PLAN2 * 16 + (size_t)((PLANLEN * 8 + 15) / 16) * 16;
    static constexpr size_t SMEM = BAR_OFF + 4 * sizeof(uint64_t);
    static constexpr int THREADS = 4 * G::NT;                // 2 groups x 2 column pairs x NT
};

struct SyncGroup {   // named barrier of one 256-thread group (ids 1 and 2; 0 is __syncthreads)
    int id, n;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
};

// the first two radix-16 stages with ONE twiddle for both lanes: float2 plan in shared memory, times u (HALF 1)
template <int L, bool INV, int HALF, class Sync>
__device__ __forceinline__ void fft_two_stages_u(P2 (&v)[16], const int t, float4* __restrict__ S, const float2* __restrict__ tw1,
                                                 const float2* __restrict__ tw2, const float2 u, Sync sync) {
    using G = Geo<L>;
    constexpr int NT = G::NT;
    dft16<INV>(v);
    if (HALF) v[0] = mul_tw<INV>(v[0], make_float4(u.x, u.x, u.y, u.y));
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        float2 w = tw1[k * NT + t];
        if (HALF) w = make_float2(fmaf(w.x, u.x, -w.y * u.y), fmaf(w.x, u.y, w.y * u.x));
        v[k] = mul_tw<INV>(v[k], make_float4(w.x, w.x, w.y, w.y));
    }
    {
        float4* __restrict__ d = S + t * 17;
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = pack(v[k]);
    }
    sync();
    {
        const float4* __restrict__ s = S + pad16(t);
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = unpack(s[n * (NT + NT / 16)]);
    }
    sync();
    dft16<INV>(v);
    const int m = t >> 4, a = t & 15;
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        const float2 w = tw2[(k - 1) * (NT / 16) + m];
        v[k] = mul_tw<INV>(v[k], make_float4(w.x, w.x, w.y, w.y));
    }
    {
        float4* __restrict__ d = S + m * 272 + a;
#pragma unroll
        for (int k = 0; k < 16; ++k) d[17 * k] = pack(v[k]);
    }
    sync();
}

template <int L, bool INV, int HALF>
__device__ __forceinline__ void focus_col4_group(const CUtensorMap& in_map, const FocusParams& p, float4* smem4, uint64_t* bar) {
    using G = Geo<L>;
    using G2 = Geo2<L>;
    using G4 = Geo4<L>;
    constexpr int NT = G::NT;
    const int lt = threadIdx.x & (2 * NT - 1);                       // thread within the group
    const int cp = lt & 1, t = lt >> 1;
    const float4* __restrict__ stage = smem4 + G4::STAGE;
    const float2* __restrict__ plan = reinterpret_cast<const float2*>(smem4 + G4::PLAN2);
    float4* S = smem4 + (HALF * 2 + cp) * G4::XBUF;
    const SyncGroup gsync{1 + HALF, 2 * NT};
    const uint64_t pol_tmp = l2_policy((p.hints & 1) ? 2 : 0), pol_in = l2_policy((p.hints & 1) ? 1 : 0);
    auto request_tile = [&](int ww) {   // one thread: gather tile ww = (field, column tile) into the staging area
        const int fb = ww / G2::TILES, tile = ww - fb * G2::TILES;
        mbar_expect_tx(bar + 1, (uint32_t)(L * 4 * sizeof(float2)));
#pragma unroll
        for (int b = 0; b < G2::NBOX_IN; ++b)
            tma_load_3d(smem4 + G4::STAGE + b * G2::BOXR * 2, &in_map, tile * 4, b * G2::BOXR, fb, bar + 1, pol_in);
        const int nn = ww + gridDim.x;   // the tile after: pull it from HBM into L2 now, so that its gather is an L2 hit
        if ((p.hints & 2) && nn < p.ntiles) {
            const int fb2 = nn / G2::TILES, tile2 = nn - fb2 * G2::TILES;
#pragma unroll
            for (int b = 0; b < G2::NBOX_IN; ++b) tma_prefetch_3d(&in_map, tile2 * 4, b * G2::BOXR, fb2);
        }
    };
    unsigned int* released = reinterpret_cast<unsigned int*>(bar + 2);   // warps that hold the inputs of the current tile, summed over tiles
    if (HALF == 0 && lt == 0) {
        mbar_expect_tx(bar, (uint32_t)(G4::PLANLEN * sizeof(float2)));
        bulk_g2s(smem4 + G4::PLAN2, p.plan, (uint32_t)(G4::PLANLEN * sizeof(float2)), bar);
        if ((int)blockIdx.x < p.ntiles) request_tile(blockIdx.x);
    }
    const float sgn = (t & 1) ? -1.0f : 1.0f;
    float2 u = make_float2(1.f, 0.f);
    if (HALF) { float sn, cs; sincospif(-(float)t / (float)L, &sn, &cs); u = make_float2(cs, sn); }   // w_2L^t = exp(-i pi t / L)
    mbar_wait(bar, 0);   // plan
    int it = 0;
    for (int w = blockIdx.x; w < p.ntiles; w += gridDim.x, ++it) {
        const int fb = w / G2::TILES, tile = w - fb * G2::TILES;
        auto sync = [&]() { gsync(); };
        P2 v[16];
        {
            float4 x[16];   // (col, col + 1) x (re, im)
            mbar_wait(bar + 1, it & 1);
            const float4* __restrict__ sg = stage + lt;   // ((n*NT + t)*2 + cp)
#pragma unroll
            for (int n = 0; n < 16; ++n) x[n] = sg[n * NT * 2];
            // lanes = the two columns; HALF 1 multiplies both by w_32^n (the n-dependent part of the input ramp)
#define PB_LANES2(n)                                                                                              \
            if (HALF == 0) v[n] = {make_float2(x[n].x, x[n].z), make_float2(x[n].y, x[n].w)};                          \
            else {                                                                                                     \
                const P2 a_ = make_lanes<INV, n>(make_float2(x[n].x, x[n].y)), b_ = make_lanes<INV, n>(make_float2(x[n].z, x[n].w)); \
                v[n] = {make_float2(a_.re.y, b_.re.y), make_float2(a_.im.y, b_.im.y)};                                 \
            }
            PB_LANES2(0) PB_LANES2(1) PB_LANES2(2) PB_LANES2(3) PB_LANES2(4) PB_LANES2(5) PB_LANES2(6) PB_LANES2(7)
            PB_LANES2(8) PB_LANES2(9) PB_LANES2(10) PB_LANES2(11) PB_LANES2(12) PB_LANES2(13) PB_LANES2(14) PB_LANES2(15)
#undef PB_LANES2
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {   // this warp holds its inputs: one of 16 releases of the staging area per tile;
                __threadfence_block();       // the LAST warp to release it requests the next tile on the spot
                const unsigned int n = atomicAdd(released, 1u) + 1u;
                __threadfence_block();
                if (n == (unsigned int)(it + 1) * (4 * NT / 32) && w + (int)gridDim.x < p.ntiles) {
                    fence_proxy_async();
                    request_tile(w + gridDim.x);
                }
            }
        }
        fft_two_stages_u<L, INV, HALF>(v, t, S, plan + G4::TW1, plan + G4::TW2, u, sync);
        fft_last_stage_load<L, INV>(v, t, S);
        sync();   // this group's exchange buffers are free for its next tile
#pragma unroll
        for (int g = 0; g < G::GI; ++g) dftR<G::R3, INV>(v + g * G::R3);
        // tile-major store: element (plane, j, tile, c) at ((plane*L/4 + j/4) * TILES + tile) * 16 + (j%4)*4 + c ; c = 2 cp
        float4* __restrict__ dst = reinterpret_cast<float4*>(
            p.tmp + fb * (2LL * L * L) + (long long)HALF * L * L + ((long long)(t >> 2) * G2::TILES + tile) * 16 + (t & 3) * 4 + 2 * cp);
#pragma unroll
        for (int g = 0; g < G::GI; ++g)
#pragma unroll
            for (int kk = 0; kk < G::R3; ++kk) {
                const P2 y = v[g * G::R3 + kk];
                const long long o = (long long)((g * NT + 256 * kk) >> 2) * (G2::TILES * 8);   // float4 units
                // (-dir*i)^k with k = 2j + HALF; j has the parity of t.  HALF 1: times (-dir*i): forward (-im, re), inverse (im, -re)
                float4 q;
                if (HALF == 0) q = make_float4(sgn * y.re.x, sgn * y.im.x, sgn * y.re.y, sgn * y.im.y);
                else q = INV ? make_float4(sgn * y.im.x, -sgn * y.re.x, sgn * y.im.y, -sgn * y.re.y)
                             : make_float4(-sgn * y.im.x, sgn * y.re.x, -sgn * y.im.y, sgn * y.re.y);
                st_hint(dst + o, q, pol_tmp);
            }
    }
}

template <int L, bool INV>
__global__ void __launch_bounds__(Geo4<L>::THREADS) focus_col4_kernel(const __grid_constant__ CUtensorMap in_map, const FocusParams p) {
    using G = Geo<L>;
    using G2 = Geo2<L>;
    using G4 = Geo4<L>;
    extern __shared__ __align__(128) float4 smem4[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(smem4) + G4::BAR_OFF);   // [0] plan [1] tile full [2] tile free
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        *reinterpret_cast<unsigned int*>(bar + 2) = 0u;   // release counter of the staging area
        mbar_fence_init();
    }
    __syncthreads();
    const int grp = threadIdx.x / (2 * G::NT);
    if (grp == 0) focus_col4_group<L, INV, 0>(in_map, p, smem4, bar);
    else focus_col4_group<L, INV, 1>(in_map, p, smem4, bar);
}

template <int L, bool INV>
__global__ void __launch_bounds__(L / 16) focus_row2_kernel(const __grid_constant__ CUtensorMap tmp_map, const FocusParams p) {
    using G = Geo<L>;
    using G2 = Geo2<L>;
    constexpr int NT = G::NT, R3 = G::R3, GI = G::GI;
    constexpr uint32_t ROW_BYTES = L * sizeof(float2);
    extern __shared__ __align__(128) unsigned char smraw[];
    float4* S = reinterpret_cast<float4*>(smraw);                // SBUF exchange
    float2* inb = reinterpret_cast<float2*>(S + G::SBUF);        // L-element input row
    uint64_t* bar = reinterpret_cast<uint64_t*>(inb + L);        // 1 mbarrier
    const int t = threadIdx.x;
    int r = blockIdx.x;
    const uint64_t pol_tmp = l2_policy((p.hints & 1) ? 1 : 0), pol_out = l2_policy((p.hints & 1) ? 1 : 0);
    auto request_row = [&](int R) {   // one thread: gather intermediate row R = (field, plane, j) into `inb`
        const int fb = R / (2 * L), rf = R - fb * (2 * L);
        const int ph = rf / L, j = rf - ph * L;
        const int grp = (fb * 2 + ph) * (L / 4) + (j >> 2);
        mbar_expect_tx(bar, ROW_BYTES);
#pragma unroll
        for (int b = 0; b < G2::NBOX_ROW; ++b) tma_load_4d(inb + b * G2::BOXT * 4, &tmp_map, 0, j & 3, b * G2::BOXT, grp, bar, pol_tmp);
    };
    if (t == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        if (r < p.nrows) request_row(r);
    }
    __syncthreads();
    const float sgn = ((t & 1) ? -1.0f : 1.0f) * p.scale;
    for (int it = 0; r < p.nrows; ++it, r += gridDim.x) {
        P2 v[16];
        mbar_wait(bar, it & 1);
        {
            const float2* __restrict__ row = inb + t;
#define PB_X(n) row[(n) * NT]
            PB_MAKE_LANES_16(v, INV, PB_X)
#undef PB_X
        }
        const int rn = r + gridDim.x;
        auto sync_and_prefetch = [&]() {
            __syncthreads();  // every thread holds its inputs in registers: the row buffer is free
            if (t == 0 && rn < p.nrows) {
                fence_proxy_async();
                request_row(rn);
            }
        };
        fft_two_stages<L, INV, 0>(v, t, S, p.plan + G::TW1, p.plan + G::TW2, sync_and_prefetch, SyncCta());
        fft_last_stage_load<L, INV>(v, t, S);
        __syncthreads();  // exchange buffer is free for the next row
#pragma unroll
        for (int g = 0; g < GI; ++g) dftR<R3, INV>(v + g * R3);
        const int fb = r / (2 * L), rf = r - fb * (2 * L);   // field of the batch, row within its [2][L] planes
        const int phs = rf / L, rr = rf - phs * L;
        const int orow = (2 * rr + phs + L) & (2 * L - 1);  // fftshift along y
        if (p.out_kind == PB_OUT_COMPLEX) {
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    const float4 o = INV ? make_float4(sgn * y.re.x, sgn * y.im.x, sgn * y.im.y, -sgn * y.re.y)
                                         : make_float4(sgn * y.re.x, sgn * y.im.x, -sgn * y.im.y, sgn * y.re.y);
                    st_hint(dst + ((j + L / 2) & (L - 1)), o, pol_out);
                }
        } else {
            float2* __restrict__ dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
            const float2 s2 = make_float2(p.scale * p.scale, p.scale * p.scale);
            const float2 wgt = make_float2(p.weight, p.weight);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    float2 I = __fmul2_rn(s2, __ffma2_rn(y.re, y.re, __fmul2_rn(y.im, y.im)));
                    float2* q = dst + ((j + L / 2) & (L - 1));
                    if (p.out_kind == PB_OUT_ACCUMULATE) I = __ffma2_rn(wgt, I, ld_stream(q));
                    st_hint(q, I, pol_out);
                }
        }
    }
}

// Row kernel, second form: ONE 512-thread CTA per SM = four independent 128-thread line groups (named barriers, own
// exchange buffer, own row buffer and mbarrier) that share the float2 twiddle plan of the column kernel in shared memory.
// Against three 128-thread CTAs with the float4 plan in L1 (focus_row2_kernel): 16 instead of 12 warps per SM, stage-1
// twiddles are 64-bit shared-memory loads (lane B's twiddle = lane A's times the per-thread constant w_2L^t, four scalar
// FP32 operations) instead of 128-bit L1 loads whose latency shows up as the kernel's top stall, and the whole unified
// array is shared memory (no L1 working set to protect).
template <int L> struct GeoR4 {
    using G = Geo<L>;
    static constexpr size_t GROUP = (size_t)G::SBUF * 16 + (size_t)L * 8;       // exchange + row buffer, bytes (multiple of 128)
    static constexpr size_t PLAN_OFF = 4 * GROUP;
    static constexpr size_t BAR_OFF = PLAN_OFF + (size_t)((Geo4<L>::PLANLEN * 8 + 15) / 16) * 16;
    static constexpr size_t SMEM = BAR_OFF + 8 * sizeof(uint64_t);
};

template <int L, bool INV>
__global__ void __launch_bounds__(4 * L / 16) focus_row4_kernel(const __grid_constant__ CUtensorMap tmp_map, const FocusParams p) {
    using G = Geo<L>;
    using G2 = Geo2<L>;
    using G4 = Geo4<L>;
    using R4 = GeoR4<L>;
    constexpr int NT = G::NT, R3 = G::R3, GI = G::GI;
    constexpr uint32_t ROW_BYTES = L * sizeof(float2);
    extern __shared__ __align__(128) unsigned char smraw[];
    const int grp = threadIdx.x / NT, t = threadIdx.x - grp * NT;
    float4* S = reinterpret_cast<float4*>(smraw + grp * R4::GROUP);      // this group's exchange buffer
    float2* inb = reinterpret_cast<float2*>(S + G::SBUF);                 // ... and L-element input row
    const float2* __restrict__ plan = reinterpret_cast<const float2*>(smraw + R4::PLAN_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + R4::BAR_OFF);   // [0] plan, [1 + grp] row of group grp
    uint64_t* bar = bars + 1 + grp;
    const uint64_t pol_tmp = l2_policy((p.hints & 1) ? 1 : 0), pol_out = l2_policy((p.hints & 1) ? 1 : 0);
    auto request_row = [&](int R) {   // one thread: gather intermediate row R = (field, plane, j) into `inb`
        const int fb = R / (2 * L), rf = R - fb * (2 * L);
        const int ph = rf / L, j = rf - ph * L;
        const int grow = (fb * 2 + ph) * (L / 4) + (j >> 2);
        mbar_expect_tx(bar, ROW_BYTES);
#pragma unroll
        for (int b = 0; b < G2::NBOX_ROW; ++b) tma_load_4d(inb + b * G2::BOXT * 4, &tmp_map, 0, j & 3, b * G2::BOXT, grow, bar, pol_tmp);
    };
    int r = blockIdx.x * 4 + grp;
    const int rstep = gridDim.x * 4;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
        mbar_expect_tx(bars, (uint32_t)(G4::PLANLEN * sizeof(float2)));
        bulk_g2s(smraw + R4::PLAN_OFF, p.plan, (uint32_t)(G4::PLANLEN * sizeof(float2)), bars);
    }
    __syncthreads();
    if (t == 0 && r < p.nrows) request_row(r);
    const SyncGroup gsync{1 + grp, NT};
    float2 u;
    { float sn, cs; sincospif(-(float)t / (float)L, &sn, &cs); u = make_float2(cs, sn); }   // w_2L^t
    const float sgn = ((t & 1) ? -1.0f : 1.0f) * p.scale;
    mbar_wait(bars, 0);
    for (int it = 0; r < p.nrows; ++it, r += rstep) {
        P2 v[16];
        mbar_wait(bar, it & 1);
        {
            const float2* __restrict__ row = inb + t;
#define PB_X(n) row[(n) * NT]
            PB_MAKE_LANES_16(v, INV, PB_X)
#undef PB_X
        }
        const int rn = r + rstep;
        // ---- stage 1
        dft16<INV>(v);
        v[0] = mul_tw<INV>(v[0], make_float4(1.0f, u.x, 0.0f, u.y));
#pragma unroll
        for (int k = 1; k < 16; ++k) {
            const float2 w = plan[G4::TW1 + k * NT + t];
            const float2 wb = make_float2(fmaf(w.x, u.x, -w.y * u.y), fmaf(w.x, u.y, w.y * u.x));
            v[k] = mul_tw<INV>(v[k], make_float4(w.x, wb.x, w.y, wb.y));
        }
        {
            float4* __restrict__ d = S + t * 17;
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k] = pack(v[k]);
        }
        gsync();   // every thread of the group holds its inputs in registers: the row buffer is free
        if (t == 0 && rn < p.nrows) {
            fence_proxy_async();
            request_row(rn);
        }
        {
            const float4* __restrict__ s4 = S + pad16(t);
#pragma unroll
            for (int n = 0; n < 16; ++n) v[n] = unpack(s4[n * (NT + NT / 16)]);
        }
        gsync();
        // ---- stage 2
        dft16<INV>(v);
        {
            const int m = t >> 4, a = t & 15;
#pragma unroll
            for (int k = 1; k < 16; ++k) {
                const float2 w = plan[G4::TW2 + (k - 1) * (NT / 16) + m];
                v[k] = mul_tw<INV>(v[k], make_float4(w.x, w.x, w.y, w.y));
            }
            float4* __restrict__ d = S + m * 272 + a;
#pragma unroll
            for (int k = 0; k < 16; ++k) d[17 * k] = pack(v[k]);
        }
        gsync();
        fft_last_stage_load<L, INV>(v, t, S);
        gsync();  // exchange buffer is free for the next row
#pragma unroll
        for (int g = 0; g < GI; ++g) dftR<R3, INV>(v + g * R3);
        const int fb = r / (2 * L), rf = r - fb * (2 * L);   // field of the batch, row within its [2][L] planes
        const int phs = rf / L, rr = rf - phs * L;
        const int orow = (2 * rr + phs + L) & (2 * L - 1);  // fftshift along y
        if (p.out_kind == PB_OUT_COMPLEX) {
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    const float4 o = INV ? make_float4(sgn * y.re.x, sgn * y.im.x, sgn * y.im.y, -sgn * y.re.y)
                                         : make_float4(sgn * y.re.x, sgn * y.im.x, -sgn * y.im.y, sgn * y.re.y);
                    st_hint(dst + ((j + L / 2) & (L - 1)), o, pol_out);
                }
        } else {
            float2* __restrict__ dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + fb * p.out_bs + (long long)orow * p.out_ld);
            const float2 s2 = make_float2(p.scale * p.scale, p.scale * p.scale);
            const float2 wgt = make_float2(p.weight, p.weight);
#pragma unroll
            for (int g = 0; g < GI; ++g)
#pragma unroll
                for (int kk = 0; kk < R3; ++kk) {
                    const int j = t + g * NT + 256 * kk;
                    const P2 y = v[g * R3 + kk];
                    float2 I = __fmul2_rn(s2, __ffma2_rn(y.re, y.re, __fmul2_rn(y.im, y.im)));
                    float2* q = dst + ((j + L / 2) & (L - 1));
                    if (p.out_kind == PB_OUT_ACCUMULATE) I = __ffma2_rn(wgt, I, ld_stream(q));
                    st_hint(q, I, pol_out);
                }
        }
    }
}

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

template <class K> int set_smem_attrs(Handle* h, K kernel, size_t smem, int ctas_per_sm) {
    if (!attr_needed(h, reinterpret_cast<const void*>(kernel))) return PB_OK;
    PB_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int pct = (int)std::min<size_t>(100, ((size_t)ctas_per_sm * (smem + 1024) * 100 + h->max_smem_optin - 1) / h->max_smem_optin);
    if (const char* e = getenv("PB_CARVEOUT_PCT")) pct = atoi(e);
    PB_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    return PB_OK;
}

// plan table, see Geo<L>
template <int L>
int get_focus_plan(Handle* h, const float4** out) {
    using G = Geo<L>;
    TwKey key{L, PB_C64, 11};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = reinterpret_cast<const float4*>(it->second); return PB_OK; }
    auto w = [&](long long num, long long den) {  // exp(-2 pi i num/den), argument reduced exactly
        num %= den;
        const double a = -2.0 * 3.14159265358979323846 * (double)num / (double)den;
        return std::complex<double>(cos(a), sin(a));
    };
    std::vector<float4> tab(G::PLAN);
    for (int k = 0; k < 16; ++k)
        for (int t = 0; t < G::NT; ++t) {
            const std::complex<double> a = w((long long)t * k, L), b = w((long long)t * (2 * k + 1), 2 * L);
            tab[G::TW1 + k * G::NT + t] = make_float4((float)a.real(), (float)b.real(), (float)a.imag(), (float)b.imag());
        }
    for (int k = 1; k < 16; ++k)
        for (int m = 0; m < G::NT / 16; ++m) {
            const std::complex<double> a = w(16LL * m * k, L);
            tab[G::TW2 + (k - 1) * (G::NT / 16) + m] = make_float4((float)a.real(), (float)a.real(), (float)a.imag(), (float)a.imag());
        }
    void* d = nullptr;
    PB_CUDA(h, cudaMalloc(&d, tab.size() * sizeof(float4)));
    PB_CUDA(h, cudaMemcpy(d, tab.data(), tab.size() * sizeof(float4), cudaMemcpyHostToDevice));
    h->tables[key] = d;
    *out = reinterpret_cast<const float4*>(d);
    return PB_OK;
}

// column-kernel plan of the v2 pipeline: float2 entries, one twiddle for both lanes (see Geo4)
template <int L>
int get_col4_plan(Handle* h, const float2** out) {
    using G4 = Geo4<L>;
    using G = Geo<L>;
    TwKey key{L, PB_C64, 13};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = reinterpret_cast<const float2*>(it->second); return PB_OK; }
    std::vector<std::complex<double>> tab(G4::PLANLEN);
    auto w = [&](long long num, long long den) {  // exp(-2 pi i num/den), argument reduced exactly
        num %= den;
        const double a = -2.0 * 3.14159265358979323846 * (double)num / (double)den;
        return std::complex<double>(cos(a), sin(a));
    };
    for (int k = 0; k < 16; ++k)
        for (int t = 0; t < G::NT; ++t) tab[G4::TW1 + k * G::NT + t] = w((long long)t * k, L);
    for (int k = 1; k < 16; ++k)
        for (int m = 0; m < G::NT / 16; ++m) tab[G4::TW2 + (k - 1) * (G::NT / 16) + m] = w(16LL * m * k, L);
    const void* d = nullptr;
    PB_TRY(upload_table(h, key, tab, &d));
    *out = reinterpret_cast<const float2*>(d);
    return PB_OK;
}

template <int L, bool INV>
int launch_focus(Handle* h, FocusParams p, int batch, cudaStream_t st) {
    using G = Geo<L>;
    constexpr int T = 4;
    const size_t smem_col = (size_t)(T * (G::SBUF + 2) + G::PLAN) * sizeof(float4) + 2 * sizeof(uint64_t);
    const size_t smem_row = (size_t)G::SBUF * sizeof(float4) + (size_t)L * sizeof(float2) + 2 * sizeof(uint64_t);
    // Resident row CTAs per SM: as many as leave >= 64 KB of the unified L1/shared array as L1 -- the 34 KB twiddle plan is
    // re-read by every line and must hit there.  (Left to itself the driver picks the 228 KB carve-out and the plan
    // streams from L2 at ~300 cycles a load.)
    const size_t unified = 256 * 1024, l1_keep = 64 * 1024;
    int row_ctas = std::max(1, (int)((unified - l1_keep) / (smem_row + 1024)));
    if (const char* e = getenv("PB_ROW_CTAS_PER_SM")) row_ctas = std::max(1, atoi(e));
    PB_TRY(set_smem_attrs(h, focus_col_kernel<L, INV, T>, smem_col, 1));
    PB_TRY(set_smem_attrs(h, focus_row_kernel<L, INV>, smem_row, row_ctas));
    PB_TRY(get_focus_plan<L>(h, &p.plan));
    p.nrows = batch * 2 * L;
    focus_col_kernel<L, INV, T><<<dim3(L / T, batch), T * G::NT, smem_col, st>>>(p);
    PB_LAUNCH_CHECK(h);
    focus_row_kernel<L, INV><<<std::min(h->sm_count * row_ctas, p.nrows), L / 16, smem_row, st>>>(p);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

// ---- TMA descriptors, cached in the handle by (base pointer, geometry) -----------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_map_c64(Handle* h, const MapKey& key, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, const CUtensorMap** out) {
    auto it = h->maps.find(key);
    if (it != h->maps.end()) { *out = &it->second; return PB_OK; }
    static EncodeTiledFn fn = nullptr;   // a driver entry point: process-wide, not per device
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* f = nullptr;
        PB_CUDA(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
        if (!f || q != cudaDriverEntryPointSuccess) return fail(h, PB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    if (h->maps.size() > 256) h->maps.clear();
    CUtensorMap m;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<void*>(key.base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, PB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    *out = &(h->maps[key] = m);
    return PB_OK;
}


template <int L, bool INV>
int launch_focus2(Handle* h, FocusParams p, int batch, cudaStream_t st) {
    using G = Geo<L>;
    using G2 = Geo2<L>;
    using G4 = Geo4<L>;
    const size_t smem_col = G4::SMEM;
    const size_t smem_row = (size_t)G::SBUF * sizeof(float4) + (size_t)L * sizeof(float2) + 2 * sizeof(uint64_t);
    auto colk = focus_col4_kernel<L, INV>;
    // resident CTAs per SM: rows -- as many as leave >= 64 KB of the unified L1/shared array as L1 (the 34 KB twiddle plan
    // is re-read by every line and must hit there; left alone the driver picks the 228 KB carve-out and the plan streams
    // from L2 at ~300 cycles a load); columns -- two 256-thread CTAs (the register file holds no more)
    static const int row_ctas_env = env_int("PB_ROW_CTAS_PER_SM", 0);
    const size_t unified = 256 * 1024, l1_keep = 64 * 1024;
    const int row_ctas = row_ctas_env > 0 ? row_ctas_env : std::max(1, (int)((unified - l1_keep) / (smem_row + 1024)));
    PB_TRY(set_smem_attrs(h, colk, smem_col, 1));
    PB_TRY(set_smem_attrs(h, focus_row2_kernel<L, INV>, smem_row, row_ctas));
    PB_TRY(get_focus_plan<L>(h, &p.plan));
    const float4* plan_row = p.plan;
    const float2* plan_col = nullptr;
    PB_TRY(get_col4_plan<L>(h, &plan_col));
    p.nrows = batch * 2 * L;
    p.ntiles = batch * G2::TILES;
    static const int hints = env_int("PB_FOCUS_L2_HINTS", 1);   // bit 0: eviction hints, bit 1: L2 prefetch of the tile after next
    p.hints = hints;
    const CUtensorMap *in_map = nullptr, *tmp_map = nullptr;
    {   // intermediate: [group = (field, plane, j/4)][tile][j%4][c]
        const cuuint64_t dims[4] = {4, 4, (cuuint64_t)G2::TILES, (cuuint64_t)batch * 2 * (L / 4)};
        const cuuint64_t str[3] = {32, 128, (cuuint64_t)G2::TILES * 128};
        const cuuint32_t box[4] = {4, 1, (cuuint32_t)G2::BOXT, 1};
        PB_TRY(get_map_c64(h, MapKey{p.tmp, 1, L, batch, 0, 0}, 4, dims, str, box, &tmp_map));
    }
    {   // input: (field, row, column) complex64; box = one 4-column run of 256 rows
        const cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)L, (cuuint64_t)batch};
        const cuuint64_t str[2] = {(cuuint64_t)p.in_ld * 8, (cuuint64_t)(batch > 1 ? p.in_bs : (long long)L * p.in_ld) * 8};
        const cuuint32_t box[3] = {4, (cuuint32_t)G2::BOXR, 1};
        PB_TRY(get_map_c64(h, MapKey{p.in, 2, L, batch, p.in_ld, batch > 1 ? p.in_bs : 0}, 3, dims, str, box, &in_map));
    }
    // column pass: one CTA per SM walks the tiles of all fields of the launch
    p.plan = reinterpret_cast<const float4*>(plan_col);
    colk<<<std::min(h->sm_count, p.ntiles), G4::THREADS, smem_col, st>>>(*in_map, p);
    PB_LAUNCH_CHECK(h);
    static const int row_version = env_int("PB_ROW_V", 4);
    if (row_version == 4) {   // four line groups per CTA, plan in shared memory
        PB_TRY(set_smem_attrs(h, focus_row4_kernel<L, INV>, GeoR4<L>::SMEM, 1));
        focus_row4_kernel<L, INV><<<std::min(h->sm_count, (p.nrows + 3) / 4), 4 * G::NT, GeoR4<L>::SMEM, st>>>(*tmp_map, p);
    } else {
        p.plan = plan_row;
        focus_row2_kernel<L, INV><<<std::min(h->sm_count * row_ctas, p.nrows), L / 16, smem_row, st>>>(*tmp_map, p);
    }
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

// =====================================================================================================
// Register-engine version of the generic axis pass (AxisPass semantics, complex64 in/out) for
// L in {1024, 2048, 4096}: two lines per thread group ride the two packed lanes.  ROWS: one CTA = 2 rows;
// COLS: one CTA = TP pairs of adjacent columns.  Used by pb_fft2 / pb_axis_dft / pb_angular_spectrum for
// every power-of-two pass the focus kernels above do not cover (CZT, FFTDFT, free space, psf -> otf).
// =====================================================================================================
__device__ __forceinline__ float2 cmul_s(float2 a, float2 b, int conj) {
    return conj ? make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y))
                : make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// DENSE: the whole line is populated and kept -- no pad window, no rotation, no crop, an even number of lines, complex
// input (free-space steps at Q = 1, plain transforms): the per-element window tests and index arithmetic of the general
// form (about as many integer instructions as there are floating-point ones) compile away.
template <int L, bool INV, bool COLS, int TP, bool RT, bool PM, bool DENSE>
__global__ void __launch_bounds__((COLS ? TP : 1) * L / 16, (RT && !COLS && L >= 4096) ? 2 : 0) axis_reg_kernel(const AxisPass p, const float2* __restrict__ tw1,
                                                                            const float2* __restrict__ tw2) {
    using G = Geo<L>;
    constexpr int NT = G::NT;
    extern __shared__ __align__(16) float4 smem4[];
    const int c = COLS ? threadIdx.x % TP : 0, t = COLS ? threadIdx.x / TP : threadIdx.x;
    const int b0 = (blockIdx.x * (COLS ? TP : 1) + c) * 2;   // lines b0 (lane A) and b0 + 1 (lane B)
    const bool hasA = DENSE || b0 < p.nb, hasB = DENSE || b0 + 1 < p.nb;
    const float2* __restrict__ in = reinterpret_cast<const float2*>(p.in);
    const float2* __restrict__ pre_e = reinterpret_cast<const float2*>(p.pre_e);
    const float2* __restrict__ pre_b = reinterpret_cast<const float2*>(p.pre_b);
    float2 pbA = make_float2(1.f, 0.f), pbB = pbA;
    if (pre_b) { if (hasA) pbA = pre_b[b0]; if (hasB) pbB = pre_b[b0 + 1]; }
    P2 v[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const int j = n * NT + t;
        int li = j;
        if (!DENSE) {
            int pp = j + p.rot_in;
            if (pp >= L) pp -= L;
            li = pp - p.in_off;
        }
        float2 xa = make_float2(0.f, 0.f), xb = xa;
        if (DENSE || (li >= 0 && li < p.n_in)) {
            const long long o = (long long)b0 * p.ibs + (long long)li * p.ies;
            if (!DENSE && p.in_kind == PB_IN_REAL) {   // real field (the PSF going to transform_psf): imaginary part 0
                const float* __restrict__ inr = reinterpret_cast<const float*>(p.in);
                if (hasA) xa.x = __ldg(inr + o);
                if (hasB) xb.x = __ldg(inr + o + p.ibs);
            } else {
                if (hasA) xa = ld_stream(in + o);
                if (hasB) xb = ld_stream(in + o + p.ibs);
            }
            if (pre_e) { const float2 w = pre_e[j - p.pre_off]; xa = cmul_s(xa, w, p.pre_e_conj); xb = cmul_s(xb, w, p.pre_e_conj); }
            if (pre_b) { xa = cmul_s(xa, pbA, p.pre_b_conj); xb = cmul_s(xb, pbB, p.pre_b_conj); }
        }
        v[n] = {make_float2(xa.x, xb.x), make_float2(xa.y, xb.y)};
    }
    if (PM) {
        // full-matrix pre-multiplier (the phase screen of a free-space chain), same access pattern as the input.  A loop of
        // its own: the 32 loads are independent of the input loads above and of each other, so they overlap, and the
        // product is one packed complex multiply per point (lane A x mA, lane B x mB).
        const float2* __restrict__ pm = reinterpret_cast<const float2*>(p.pre_mat);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const int j = n * NT + t;
            int li = j;
            if (!DENSE) {
                int pp = j + p.rot_in;
                if (pp >= L) pp -= L;
                li = pp - p.in_off;
            }
            if (DENSE || (li >= 0 && li < p.n_in)) {
                const long long om = (long long)b0 * p.pmi_bs + (long long)li * p.pmi_es;
                float2 ma = make_float2(1.f, 0.f), mb = ma;
                if (hasA) ma = ld_stream(pm + om);
                if (hasB) mb = ld_stream(pm + om + p.pmi_bs);
                const float4 w = make_float4(ma.x, mb.x, ma.y, mb.y);
                v[n] = p.pre_mat_conj ? mul_tw<true>(v[n], w) : mul_tw<false>(v[n], w);
            }
        }
    }
    // TP interleaved lines per CTA: consecutive exchange buffers are offset by 32 / TP banks (8 / TP float4 of padding), so
    // that the 8 threads of a 128-bit wavefront (8 / TP consecutive t  x  TP lines) cover all 32 banks
    float4* S = smem4 + c * (G::SBUF + 8 / (COLS ? TP : 1) % 8);
    fft_two_stages<L, INV, 2>(v, t, S, tw1, tw2, SyncCta(), SyncCta());
    fft_last_stage_load<L, INV>(v, t, S);
#pragma unroll
    for (int g = 0; g < G::GI; ++g) dftR<G::R3, INV>(v + g * G::R3);
    const float2* __restrict__ post_e = reinterpret_cast<const float2*>(p.post_e);
    const float2* __restrict__ post_b = reinterpret_cast<const float2*>(p.post_b);
    float2 qbA = make_float2(1.f, 0.f), qbB = qbA;
    if (post_b) { if (hasA) qbA = post_b[b0]; if (hasB) qbB = post_b[b0 + 1]; }
    const float scale = (float)p.scale;
    float2* __restrict__ out = reinterpret_cast<float2*>(p.out);
    if (RT) {
        // Round trip (free-space propagation along this axis): the spectrum is multiplied by the transfer
        // factors post_e[k] * post_b[line] and transformed straight back.  The forward pass leaves thread t
        // with X[t + NT*(g + GI*kk)], which IS the inverse pass's stage-1 register layout (n = g + GI*kk):
        // the data never leaves the registers between the two transforms.
        P2 w[16];
#pragma unroll
        for (int g = 0; g < G::GI; ++g)
#pragma unroll
            for (int kk = 0; kk < G::R3; ++kk) {
                const int k = t + g * NT + 256 * kk;
                const P2 y = v[g * G::R3 + kk];
                float2 ya = make_float2(y.re.x, y.im.x), yb = make_float2(y.re.y, y.im.y);
                if (post_e) { const float2 m = post_e[k - p.post_off]; ya = cmul_s(ya, m, p.post_e_conj); yb = cmul_s(yb, m, p.post_e_conj); }
                if (post_b) { ya = cmul_s(ya, qbA, p.post_b_conj); yb = cmul_s(yb, qbB, p.post_b_conj); }
                w[g + G::GI * kk] = {make_float2(ya.x, yb.x), make_float2(ya.y, yb.y)};
            }
        __syncthreads();   // all last-stage reads of S are done before the inverse pass overwrites it
        fft_two_stages<L, !INV, 2>(w, t, S, tw1, tw2, SyncCta(), SyncCta());
        fft_last_stage_load<L, !INV>(w, t, S);
#pragma unroll
        for (int g = 0; g < G::GI; ++g) dftR<G::R3, !INV>(w + g * G::R3);
#pragma unroll
        for (int g = 0; g < G::GI; ++g)
#pragma unroll
            for (int kk = 0; kk < G::R3; ++kk) {
                const int j = t + g * NT + 256 * kk;
                int q = j;
                if (!DENSE) {
                    q = j - p.crop_off + p.rot_out;
                    if (q < 0) q += L;
                    if (q >= L) q -= L;
                    if (q >= p.n_out) continue;
                }
                const P2 y = w[g * G::R3 + kk];
                float2 ya = make_float2(y.re.x, y.im.x), yb = make_float2(y.re.y, y.im.y);
                if (p.post_e2) {  // final multiplier, indexed by the output sample (the CZT's a(m)*phase(m))
                    const float2 m = reinterpret_cast<const float2*>(p.post_e2)[q];
                    ya = cmul_s(ya, m, p.post_e2_conj); yb = cmul_s(yb, m, p.post_e2_conj);
                }
                const long long o = (long long)b0 * p.obs + (long long)q * p.oes;
                if (p.out_kind == PB_OUT_COMPLEX) {
                    if (hasA) st_stream(out + o, make_float2(ya.x * scale, ya.y * scale));
                    if (hasB) st_stream(out + o + p.obs, make_float2(yb.x * scale, yb.y * scale));
                } else {   // |.|^2 (or weight * |.|^2 added to the plane): the complex field is never written
                    float* __restrict__ outr = reinterpret_cast<float*>(p.out);
                    const float s2 = scale * scale, wgt = (float)p.weight;
                    float ia = s2 * fmaf(ya.x, ya.x, ya.y * ya.y), ib = s2 * fmaf(yb.x, yb.x, yb.y * yb.y);
                    if (p.out_kind == PB_OUT_ACCUMULATE) {
                        if (hasA) ia = fmaf(wgt, ia, outr[o]);
                        if (hasB) ib = fmaf(wgt, ib, outr[o + p.obs]);
                    }
                    if (hasA) outr[o] = ia;
                    if (hasB) outr[o + p.obs] = ib;
                }
            }
        return;
    }
#pragma unroll
    for (int g = 0; g < G::GI; ++g)
#pragma unroll
        for (int kk = 0; kk < G::R3; ++kk) {
            const int k = t + g * NT + 256 * kk;
            int q = k;
            if (!DENSE) {
                q = k - p.crop_off + p.rot_out;
                if (q < 0) q += L;
                if (q >= L) q -= L;
                if (q >= p.n_out) continue;
            }
            const P2 y = v[g * G::R3 + kk];
            float2 ya = make_float2(y.re.x, y.im.x), yb = make_float2(y.re.y, y.im.y);
            if (post_e) { const float2 w = post_e[k - p.post_off]; ya = cmul_s(ya, w, p.post_e_conj); yb = cmul_s(yb, w, p.post_e_conj); }
            if (post_b) { ya = cmul_s(ya, qbA, p.post_b_conj); yb = cmul_s(yb, qbB, p.post_b_conj); }
            const long long o = (long long)b0 * p.obs + (long long)q * p.oes;
            if (hasA) st_stream(out + o, make_float2(ya.x * scale, ya.y * scale));
            if (hasB) st_stream(out + o + p.obs, make_float2(yb.x * scale, yb.y * scale));
        }
}

// plain twiddle tables: tw1[(k-1)*NT + t] = w_L^(t k), tw2[(k-1)*NT/16 + m] = w_L^(16 m k), contiguous
template <int L>
int get_plain_plan(Handle* h, const float2** tw1, const float2** tw2) {
    using G = Geo<L>;
    TwKey key{L, PB_C64, 12};
    const float2* base = nullptr;
    auto it = h->tables.find(key);
    if (it != h->tables.end()) base = reinterpret_cast<const float2*>(it->second);
    else {
        std::vector<std::complex<double>> tab(15 * G::NT + 15 * (G::NT / 16));
        auto w = [&](long long num, long long den) {
            num %= den;
            const double a = -2.0 * 3.14159265358979323846 * (double)num / (double)den;
            return std::complex<double>(cos(a), sin(a));
        };
        for (int k = 1; k < 16; ++k)
            for (int t = 0; t < G::NT; ++t) tab[(k - 1) * G::NT + t] = w((long long)t * k, L);
        for (int k = 1; k < 16; ++k)
            for (int m = 0; m < G::NT / 16; ++m) tab[15 * G::NT + (k - 1) * (G::NT / 16) + m] = w(16LL * m * k, L);
        const void* d = nullptr;
        PB_TRY(upload_table(h, key, tab, &d));
        base = reinterpret_cast<const float2*>(d);
    }
    *tw1 = base;
    *tw2 = base + 15 * G::NT;
    return PB_OK;
}

template <int L, bool INV, bool COLS, bool RT, bool PM = false, bool DENSE = false>
int launch_axis_reg(Handle* h, const AxisPass& p, cudaStream_t st) {
    using G = Geo<L>;
    constexpr int TP = COLS ? (L >= 4096 ? 2 : 2) : 1;
    const size_t smem = (size_t)TP * (G::SBUF + 8 / TP % 8) * sizeof(float4);
    if (attr_needed(h, reinterpret_cast<const void*>(axis_reg_kernel<L, INV, COLS, TP, RT, PM, DENSE>))) {
        PB_CUDA(h, cudaFuncSetAttribute(axis_reg_kernel<L, INV, COLS, TP, RT, PM, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // keep >= 64 KB of L1 for the twiddle tables (see launch_focus)
        const size_t unified = 256 * 1024, l1_keep = 64 * 1024;
        const int ctas = std::max<int>(1, (int)((unified - l1_keep) / (smem + 1024)));
        const int pct = (int)std::min<size_t>(100, ((size_t)ctas * (smem + 1024) * 100 + h->max_smem_optin - 1) / h->max_smem_optin);
        PB_CUDA(h, cudaFuncSetAttribute(axis_reg_kernel<L, INV, COLS, TP, RT, PM, DENSE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    const float2 *tw1 = nullptr, *tw2 = nullptr;
    PB_TRY(get_plain_plan<L>(h, &tw1, &tw2));
    const int lines_per_cta = 2 * TP;
    const int grid = (p.nb + lines_per_cta - 1) / lines_per_cta;
    axis_reg_kernel<L, INV, COLS, TP, RT, PM, DENSE><<<grid, TP * G::NT, smem, st>>>(p, tw1, tw2);
    PB_LAUNCH_CHECK(h);
    return PB_OK;
}

template <int L>
int dispatch_axis_reg(Handle* h, const AxisPass& p, cudaStream_t st) {
    const bool cols = p.batch_contiguous != 0;
    // the whole line populated and kept, lines in pairs, complex in: the index logic of the general kernel compiles away
    const bool dense = p.rot_in == 0 && p.in_off == 0 && p.n_in == L && p.crop_off == 0 && p.rot_out == 0 && p.n_out == L &&
                       !(p.nb & 1) && p.in_kind == PB_IN_COMPLEX && !p.post_e2 && p.pre_off == 0 && p.post_off == 0 &&
                       (!cols || !(p.nb & 3));
    if (p.pre_mat) {   // the screened first pass of a free-space step: rows, plain transform
        if (cols || p.roundtrip) return PB_ERR_UNSUPPORTED;
        if (dense) return p.dir < 0 ? launch_axis_reg<L, false, false, false, true, true>(h, p, st) : launch_axis_reg<L, true, false, false, true, true>(h, p, st);
        return p.dir < 0 ? launch_axis_reg<L, false, false, false, true>(h, p, st) : launch_axis_reg<L, true, false, false, true>(h, p, st);
    }
    if (dense && L >= 2048) {   // the free-space passes and plain transforms of the large configurations
        if (p.roundtrip) {
            if (cols) return p.dir < 0 ? launch_axis_reg<L, false, true, true, false, true>(h, p, st) : launch_axis_reg<L, true, true, true, false, true>(h, p, st);
        } else if (!cols) {
            return p.dir < 0 ? launch_axis_reg<L, false, false, false, false, true>(h, p, st) : launch_axis_reg<L, true, false, false, false, true>(h, p, st);
        } else {
            return p.dir < 0 ? launch_axis_reg<L, false, true, false, false, true>(h, p, st) : launch_axis_reg<L, true, true, false, false, true>(h, p, st);
        }
    }
    if (p.roundtrip) {
        if (p.dir < 0) return cols ? launch_axis_reg<L, false, true, true>(h, p, st) : launch_axis_reg<L, false, false, true>(h, p, st);
        return cols ? launch_axis_reg<L, true, true, true>(h, p, st) : launch_axis_reg<L, true, false, true>(h, p, st);
    }
    if (p.dir < 0) return cols ? launch_axis_reg<L, false, true, false>(h, p, st) : launch_axis_reg<L, false, false, false>(h, p, st);
    return cols ? launch_axis_reg<L, true, true, false>(h, p, st) : launch_axis_reg<L, true, false, false>(h, p, st);
}

}  // namespace

int try_tuned_axis_pass(Handle* h, const AxisPass& p, cudaStream_t st) {
    static const bool disabled = getenv("PB_DISABLE_TUNED") != nullptr || getenv("PB_DISABLE_TUNED_AXIS") != nullptr;
    if (disabled) return PB_ERR_UNSUPPORTED;
    if (p.dtype != PB_C64 || p.in_kind == PB_IN_AMP_OPD) return PB_ERR_UNSUPPORTED;
    if (p.out_kind != PB_OUT_COMPLEX && !p.roundtrip) return PB_ERR_UNSUPPORTED;   // |.|^2 stores: round-trip epilogue only
    if (p.pre_e2 || (p.post_e2 && !p.roundtrip) || p.post_mat) return PB_ERR_UNSUPPORTED;
    if (p.Llog != p.L || (p.Llog_out != 0 && p.Llog_out != p.L)) return PB_ERR_UNSUPPORTED;
    // any batch: a single line (the kernel spectrum of a CZT plan, built per wavelength) takes 23 us on the generic
    // kernel's six synchronised radix-4 stages and a fraction of that here
    switch (p.L) {
        case 1024: return dispatch_axis_reg<1024>(h, p, st);
        case 2048: return dispatch_axis_reg<2048>(h, p, st);
        case 4096: return dispatch_axis_reg<4096>(h, p, st);
        default: return PB_ERR_UNSUPPORTED;
    }
}

// fields per launch pair: more than one evens out the tails (512 column tiles on 148 SMs = 3.46 waves; 4096 rows on
// 444 persistent row CTAs = 9.2 rounds).  The intermediates of a group (64 MiB each at N = 2048) no longer fit L2
// together and spill to HBM, which has the headroom: measured 87.3 / 78.8 / 76.7 / 75.7 us per field at 1 / 2 / 4 / 8.
static int focus_fields_per_launch() {
    static const int n = [] { const char* e = getenv("PB_FOCUS_BATCH"); return e ? std::max(1, atoi(e)) : 8; }();
    return n;
}

int try_tuned_fft2_batch(Handle* h, int dtype, const void* in, int in_kind, const void* amp, int amp_kind, double kturns,
                         int batch, long long in_bs, long long amp_bs, int ny, int nx, long long in_ld, int ky, int kx, int dir,
                         double scale, int shift_in, int shift_out, void* out, int out_kind, double weight, int oy, int ox,
                         long long out_ld, long long out_bs, cudaStream_t st) {
    static const bool disabled = getenv("PB_DISABLE_TUNED") != nullptr;
    if (disabled) return PB_ERR_UNSUPPORTED;
    if (dtype != PB_C64 || ny != nx || ky != 2 * ny || kx != 2 * nx || !shift_in || !shift_out || oy != ky || ox != kx)
        return PB_ERR_UNSUPPORTED;
    if (nx != 512 && nx != 1024 && nx != 2048) return PB_ERR_UNSUPPORTED;
    if (in_kind == PB_IN_REAL) return PB_ERR_UNSUPPORTED;
    if (out_kind == PB_OUT_COMPLEX ? (out_ld & 1) || (out_bs & 1) || ((uintptr_t)out & 15) : (out_ld & 1) || (out_bs & 1) || ((uintptr_t)out & 7))
        return PB_ERR_UNSUPPORTED;  // vector stores need aligned rows
    const int N = nx;
    const int per = std::min(batch, focus_fields_per_launch());
    void* tmp = nullptr;
    PB_TRY(ensure_scratch(h, 0, (size_t)per * 2 * N * N * sizeof(float2), &tmp));
    const size_t in_elt = in_kind == PB_IN_COMPLEX ? sizeof(float2) : sizeof(float);
    const size_t amp_elt = amp_kind == PB_AMP_U8 ? 1 : sizeof(float);
    const size_t out_elt = out_kind == PB_OUT_COMPLEX ? sizeof(float2) : sizeof(float);
    for (int b0 = 0; b0 < batch; b0 += per) {
        const int nb = std::min(per, batch - b0);
        FocusParams p;
        p.in = reinterpret_cast<const char*>(in) + (size_t)b0 * in_bs * in_elt;
        p.in_kind = in_kind;
        p.amp = amp ? reinterpret_cast<const char*>(amp) + (size_t)b0 * amp_bs * amp_elt : nullptr;
        p.amp_kind = amp_kind; p.kturns = kturns; p.in_ld = in_ld;
        p.tmp = reinterpret_cast<float2*>(tmp);
        p.plan = nullptr;
        p.out = reinterpret_cast<char*>(out) + (size_t)b0 * out_bs * out_elt;
        p.out_ld = out_ld; p.out_kind = out_kind; p.scale = (float)scale; p.weight = (float)weight;
        p.nrows = 0; p.ntiles = 0; p.hints = 0;
        p.in_bs = in_bs; p.amp_bs = amp_bs; p.out_bs = out_bs;
        int rc;
        static const int version = env_int("PB_FOCUS_V", 2);
        // v2 takes its input tiles by tensor TMA: complex input, 16-byte aligned base and even pitches; everything else
        // (the fused phase-screen input, odd pitches) runs the v1 kernels
        const bool v2_ok = in_kind == PB_IN_COMPLEX && !((uintptr_t)p.in & 15) && !(in_ld & 1) && !(in_bs & 1);
        if (version >= 2 && v2_ok) {
            if (dir < 0) rc = N == 512 ? launch_focus2<512, false>(h, p, nb, st) : N == 1024 ? launch_focus2<1024, false>(h, p, nb, st) : launch_focus2<2048, false>(h, p, nb, st);
            else rc = N == 512 ? launch_focus2<512, true>(h, p, nb, st) : N == 1024 ? launch_focus2<1024, true>(h, p, nb, st) : launch_focus2<2048, true>(h, p, nb, st);
        } else if (dir < 0) rc = N == 512 ? launch_focus<512, false>(h, p, nb, st) : N == 1024 ? launch_focus<1024, false>(h, p, nb, st) : launch_focus<2048, false>(h, p, nb, st);
        else rc = N == 512 ? launch_focus<512, true>(h, p, nb, st) : N == 1024 ? launch_focus<1024, true>(h, p, nb, st) : launch_focus<2048, true>(h, p, nb, st);
        if (rc != PB_OK) return rc;
    }
    return PB_OK;
}

int try_tuned_fft2(Handle* h, int dtype, const void* in, int in_kind, const void* amp, int amp_kind, double kturns,
                   int ny, int nx, long long in_ld, int ky, int kx, int dir, double scale, int shift_in, int shift_out,
                   void* out, int out_kind, double weight, int oy, int ox, long long out_ld, cudaStream_t st) {
    return try_tuned_fft2_batch(h, dtype, in, in_kind, amp, amp_kind, kturns, 1, 0, 0, ny, nx, in_ld, ky, kx, dir, scale,
                                shift_in, shift_out, out, out_kind, weight, oy, ox, out_ld, 0, st);
}


}  // namespace pb
