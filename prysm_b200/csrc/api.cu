// Handle lifetime, scratch arenas, error reporting.
#include "common.cuh"

namespace pb {

int ensure_scratch(Handle* h, int slot, size_t bytes, void** out) {
    if (bytes > h->scratch_bytes[slot]) {
        if (h->scratch[slot]) {
            // an earlier call may still be reading the old arena on some stream
            PB_CUDA(h, cudaDeviceSynchronize());
            PB_CUDA(h, cudaFree(h->scratch[slot]));
            h->scratch[slot] = nullptr;
            h->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8;
        if (cudaMalloc(&h->scratch[slot], want) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, PB_ERR_ALLOC, "scratch allocation of " + std::to_string(want) + " bytes failed");
        }
        h->scratch_bytes[slot] = want;
    }
    *out = h->scratch[slot];
    return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" int pb_create(pb_handle_t* out, int device) {
    if (!out) return PB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PB_ERR_CUDA;
    DeviceGuard guard(device);   // the caller's current device is restored on return
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return PB_ERR_CUDA;
    if (cudaFree(0) != cudaSuccess) return PB_ERR_CUDA;   // make sure the device's primary context exists
    Handle* h = new Handle();
    h->device = device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    *out = reinterpret_cast<pb_handle_t>(h);
    return PB_OK;
}

extern "C" int pb_destroy(pb_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h) return PB_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->tables) cudaFree(kv.second);
    for (int i = 0; i < 3; ++i) if (h->scratch[i]) cudaFree(h->scratch[i]);
    for (auto& e : h->side_ev) if (e) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    delete h;
    return PB_OK;
}

extern "C" const char* pb_last_error(pb_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    return h ? h->err.c_str() : "null handle";
}

extern "C" const char* pb_version(void) { return "prysm_b200 0.1 (sm_100a)"; }

extern "C" long long pb_launch_count(pb_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    return h ? h->launches : 0;
}
