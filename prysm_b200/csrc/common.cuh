// Shared declarations for libprysm_b200: handle, error plumbing, complex helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <complex>
#include <cstdio>
#include <map>
#include <set>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/prysm_b200.h"

namespace pb {

template <typename R> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };
template <typename R> using cplx = typename cplx_of<R>::type;

template <typename R> __host__ __device__ inline cplx<R> mk(R x, R y) { cplx<R> c; c.x = x; c.y = y; return c; }
template <typename C> __host__ __device__ inline C cmul(C a, C b) {
    C c; c.x = a.x * b.x - a.y * b.y; c.y = a.x * b.y + a.y * b.x; return c;
}
template <typename C> __host__ __device__ inline C cmulc(C a, C b) {  // a * conj(b)
    C c; c.x = a.x * b.x + a.y * b.y; c.y = a.y * b.x - a.x * b.y; return c;
}
template <typename C> __host__ __device__ inline C cadd(C a, C b) { C c; c.x = a.x + b.x; c.y = a.y + b.y; return c; }
template <typename C> __host__ __device__ inline C csub(C a, C b) { C c; c.x = a.x - b.x; c.y = a.y - b.y; return c; }
template <typename C> __host__ __device__ inline C cconj(C a) { a.y = -a.y; return a; }

// exp(2*pi*i*turns) with the argument reduced in fp64 before the sincos.
__device__ inline float2 expi_turns(double turns, float) {
    double fr = turns - rint(turns);
    float s, c; sincospif(2.0f * (float)fr, &s, &c);
    return make_float2(c, s);
}
__device__ inline double2 expi_turns(double turns, double) {
    double fr = turns - rint(turns);
    double s, c; sincospi(2.0 * fr, &s, &c);
    return make_double2(c, s);
}

struct TwKey {
    int n; int dtype; int kind;  // kind 0: w_n^k table; 1/2: bluestein chirp (dir -1/+1); 3/4: bluestein filter spectrum
    bool operator<(const TwKey& o) const {
        if (n != o.n) return n < o.n;
        if (dtype != o.dtype) return dtype < o.dtype;
        return kind < o.kind;
    }
};

// cache key of a TMA descriptor: everything cuTensorMapEncodeTiled was given
struct MapKey {
    const void* base; int kind; long long a, b, c, d;
    bool operator<(const MapKey& o) const {
        if (base != o.base) return base < o.base;
        if (kind != o.kind) return kind < o.kind;
        if (a != o.a) return a < o.a;
        if (b != o.b) return b < o.b;
        if (c != o.c) return c < o.c;
        return d < o.d;
    }
};

struct Handle {
    int device = 0;
    std::string err;
    long long launches = 0;
    std::map<TwKey, void*> tables;
    // scratch arenas, grown on demand (never on a warmed-up hot call)
    void* scratch[3] = {nullptr, nullptr, nullptr};
    size_t scratch_bytes[3] = {0, 0, 0};
    int sm_count = 148;
    int max_smem_optin = 0;
    // kernels whose opt-in attributes (dynamic shared memory, carve-out) were set on THIS handle's device.
    // cudaFuncSetAttribute acts on the current device's copy of the function, so the record is per handle
    // (= per device and stream), never a process-wide static.
    std::set<const void*> attr_done;
    std::map<MapKey, CUtensorMap> maps;   // TMA descriptors by (base pointer, geometry)
    // helper stream + events of the entry points that overlap small preparatory launches with the main sequence
    // (pb_polychromatic_czt: the Bluestein plan of unit i+1 beside the transforms of unit i); created on first use
    cudaStream_t side = nullptr;
    cudaEvent_t side_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

// Makes the handle's device current for the duration of an entry point and restores the caller's device on
// the way out (one process may drive several GPUs; torch's current device must not change under the caller).
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int want) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != want) switched = cudaSetDevice(want) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// NVTX range per C-ABI call: shows up in nsys / ncu --nvtx timelines, a no-op without a tool attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// First lines of every extern "C" entry point that takes a handle
#define PB_ENTER(hh)                                         \
    pb::Handle* h = reinterpret_cast<pb::Handle*>(hh);       \
    if (!h) return PB_ERR_INVALID;                           \
    pb::DeviceGuard pb_guard_(h->device);                    \
    pb::NvtxRange pb_range_(__func__)

// true the first time a kernel is seen on this handle: the caller then sets its opt-in attributes
inline bool attr_needed(Handle* h, const void* fn) { return h->attr_done.insert(fn).second; }

inline int fail(Handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define PB_CUDA(h, call)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return pb::fail(h, PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define PB_LAUNCH_CHECK(h)                                                                    \
    do {                                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess)                                                                \
            return pb::fail(h, PB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
        (h)->launches++;                                                                      \
    } while (0)

#define PB_TRY(expr)               \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != PB_OK) return rc_; \
    } while (0)

int ensure_scratch(Handle* h, int slot, size_t bytes, void** out);
// device table of w_n^k = exp(-2*pi*i*k/n), k in [0,n)
int get_twiddles(Handle* h, int n, int dtype, const void** out);
// upload a host fp64 complex table in the precision of key.dtype and cache it in the handle
int upload_table(Handle* h, const TwKey& key, const std::vector<std::complex<double>>& t, const void** out);

inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
inline size_t csize(int dtype) { return dtype == PB_C64 ? 8 : 16; }
inline size_t rsize(int dtype) { return dtype == PB_C64 ? 4 : 8; }

}  // namespace pb
