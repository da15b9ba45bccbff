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
    if (e != cudaSuccess) { set_last_cuda_error(e); return PAPR_ERR_CUDA; }
    return PAPR_OK;
}

#define PAPR_CUDA_TRY(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) { papr::set_last_cuda_error(_e); return PAPR_ERR_CUDA; } \
    } while (0)

constexpr int kNumSMs = 148;   // B200

}  // namespace papr
