// msda_host.h -- host-side helpers shared by the translation units of libmsda_b200.so
// (msda_api.cu: C ABI + dispatch; msda_bwd_bin.cu: the bin-major backward pass).
// Internal: nothing here is part of the C ABI (include/msda_b200.h).
#pragma once

#include <cuda_runtime.h>

#include "../../include/msda_b200.h"
#include "msda_common.cuh"

namespace msda_host {

// Status + message for msda_last_error() (thread-local, like errno).
int fail(int code, const char* fmt, ...);

// Per-kernel timing (msda_profile_*): process-wide, so that launches made on autograd's
// worker thread are seen by the thread that enabled the profile.
void prof_begin(cudaStream_t st, const char* name);
void prof_end(cudaStream_t st);

// Launch bookkeeping behind msda_last_launch_count(): process-wide counter.
void count_launch();

int num_sms();

// SURVEY.md section 8b asks for an explicit device: the ABI takes pointers, and a pointer names its device.  Every
// entry point runs under this guard: the device that owns `ptr` becomes current for the call (the reference has no
// guard and relies on torch.cuda.set_device, /root/reference/trainer.py:447), the previous one is restored on return.
// ok() is false when `ptr` is not device (or managed) memory at all.
class DeviceGuard {
public:
    explicit DeviceGuard(const void* ptr) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;
        ok_ = true;
        if (cudaGetDevice(&prev_) == cudaSuccess && prev_ != a.device) switched_ = cudaSetDevice(a.device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched_) cudaSetDevice(prev_); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
    bool ok() const { return ok_; }
private:
    int prev_ = -1;
    bool ok_ = false, switched_ = false;
};

// cudaOccupancyMaxActiveBlocksPerMultiprocessor, cached per (kernel, device, dynamic smem).
int blocks_per_sm_cached(const void* kernel, int threads, size_t dyn_smem);

template <typename K>
int persistent_grid(K kernel, int threads, long long work_items, size_t dyn_smem = 0) {
    const long long cap = (long long)num_sms() * blocks_per_sm_cached(reinterpret_cast<const void*>(kernel), threads, dyn_smem);
    long long g = work_items < cap ? work_items : cap;
    return (int)(g < 1 ? 1 : g);
}

#define MSDA_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return msda_host::fail(MSDA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

#define MSDA_LAUNCHED(name)                                                                          \
    do {                                                                                             \
        msda_host::count_launch();                                                                   \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess)                                                                       \
            return msda_host::fail(MSDA_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// msda_bwd_bin.cu: part B of the backward on the tile path (fp32 / bf16 rows): presort of oversized
// sub-bins + the shared-memory tile kernel.  vdt: MSDA_F32 or MSDA_BF16; (vec, g): lanes layout of a row.
int launch_grad_value_tile(const msda::Params& p, int vdt, int vec, int g, cudaStream_t st);

}  // namespace msda_host
