// C-ABI bookkeeping entry points (version / error strings).
#include "common.cuh"
#include <string.h>

namespace papr {
static thread_local char g_last_error[256] = "no error";
void set_last_cuda_error(cudaError_t e)
{
    strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = 0;
}
}  // namespace papr

extern "C" int papr_abi_version(void) { return 1; }

extern "C" const char *papr_last_cuda_error(void) { return papr::g_last_error; }

extern "C" const char *papr_status_string(int status)
{
    switch (status) {
        case PAPR_OK: return "ok";
        case PAPR_ERR_INVALID_ARGUMENT: return "invalid argument";
        case PAPR_ERR_CUDA: return "CUDA error";
        case PAPR_ERR_UNSUPPORTED_DEVICE: return "unsupported device (needs sm_100)";
        default: return "unknown status";
    }
}
