// Shared helpers for the papr_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/papr_b200.h"

namespace papr {

void set_last_cuda_error(cudaError_t e);

inline int check_launch()
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction) { set_last_cuda_error(e); return PAPR_ERR_UNSUPPORTED_DEVICE; }
    if (e != cudaSuccess) { set_last_cuda_error(e); return PAPR_ERR_CUDA; }
    return PAPR_OK;
}

#define PAPR_CUDA_TRY(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) { papr::set_last_cuda_error(_e); return PAPR_ERR_CUDA; } \
    } while (0)

// Opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute: set it once per (kernel, device).
struct SmemAttrOnce { bool done[64] = {}; };
template <typename Kernel>
inline cudaError_t ensure_dyn_smem(SmemAttrOnce &once, Kernel kernel, int bytes)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && once.done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && tracked) once.done[dev] = true;
    return e;
}

// SM count of the current device (148 on B200), queried once per device: persistent kernels size their grids with it
inline int num_sms()
{
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}
#define kNumSMs (::papr::num_sms())

// 16-byte vector reduction into global memory (sm_90+): one instruction instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace papr
