// Launch bookkeeping shared by every translation unit of libppde_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace ppde {
extern int g_launch_count;                 // defined in cabi.cu
inline int launch_done(int n = 1) {        // call right after <<<>>>; 0 on success, else cudaError_t
    g_launch_count += n;
    return (int)cudaGetLastError();
}
}  // namespace ppde
