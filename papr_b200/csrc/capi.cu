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

extern "C" int papr_abi_version(void) { return 2; }

// The library holds sm_100a code only (tcgen05 / TMEM / bulk TMA): anything else cannot run it.
extern "C" int papr_check_device(int device)
{
    int dev = device;
    if (dev < 0) PAPR_CUDA_TRY(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    PAPR_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PAPR_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return (major == 10 && minor == 0) ? PAPR_OK : PAPR_ERR_UNSUPPORTED_DEVICE;
}

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
