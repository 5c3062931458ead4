// Launch bookkeeping shared by every translation unit of libppde_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ppde {
extern int g_launch_count;                 // defined in cabi.cu
inline int launch_done(int n = 1) {        // call right after <<<>>>; 0 on success, else cudaError_t
    g_launch_count += n;
    return (int)cudaGetLastError();
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) function attribute: remember per DEVICE how much
// has been requested for a kernel.  `tab` is a zero-initialised static array of PPDE_MAX_DEVICES entries owned by the caller
// (one per kernel instantiation).  A race between two host threads on the first launch only repeats the driver call.
constexpr int PPDE_MAX_DEVICES = 64;
struct SmemCache { size_t v[PPDE_MAX_DEVICES]; };
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t smem, SmemCache& tab) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= PPDE_MAX_DEVICES) return cudaErrorInvalidDevice;
    if (smem > tab.v[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tab.v[dev] = smem;
    }
    return cudaSuccess;
}
inline int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
}  // namespace ppde
