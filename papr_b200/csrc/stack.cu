// papr_stack_bf16: a whole MLP stack per launch (see stack_body.cuh for the kernel body and its design notes).
#include "stack_body.cuh"

namespace papr {

template <bool RELU, bool TRACE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kStkThreads, 1) stack_kernel(const __grid_constant__ StackParams p)
{
    stack_body<RELU, TRACE>(p, blockIdx.x >> 1, gridDim.x >> 1);
}

}  // namespace papr

static long long *g_stack_trace = nullptr;
// debug hook (not part of the public header): device buffer receiving clock stamps of cluster 0
extern "C" void papr_debug_stack_trace(long long *buf) { g_stack_trace = buf; }

extern "C" int papr_stack_bf16(const void *x, int K0, const papr_stack_layer *layers, int n_layers, int64_t rows, float slope,
                               void *stream)
{
    return papr_stack_bf16_ex(x, K0, layers, n_layers, rows, slope, 0, stream);
}

extern "C" int papr_stack_bf16_ex(const void *x, int K0, const papr_stack_layer *layers, int n_layers, int64_t rows, float slope,
                                  int max_ctas, void *stream)
{
    using namespace papr;
    StackParams p;
    int smem = 0;
    const int st = fill_stack_params(p, x, K0, layers, n_layers, rows, slope, &smem);
    if (st != PAPR_OK) return st;
    p.trace = g_stack_trace;
    { static int ring = -1; if (ring < 0) { const char *e = getenv("PAPR_DBG_STACK_RING"); ring = e ? atoi(e) : 0; } p.dbg_ring = ring; }
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, stack_kernel<true>, kStkMaxSmem));
    static SmemAttrOnce once1;
    PAPR_CUDA_TRY(ensure_dyn_smem(once1, stack_kernel<false>, kStkMaxSmem));
    const int64_t n_quads = (p.n_tiles + 3) / 4;
    int grid = (int)(2 * (n_quads < kNumSMs / 2 ? n_quads : kNumSMs / 2));
    { static int cap = -1; if (cap < 0) { const char *e = getenv("PAPR_DBG_STACK_GRID"); cap = e ? atoi(e) : 0; } if (cap > 0 && grid > cap) grid = cap; }
    if (max_ctas >= 2 && grid > max_ctas) grid = max_ctas & ~1;        // CTA pairs: leave the other SMs to a concurrent kernel
    if (p.trace) {       // tools/trace_stack.py: the relu kernel with its clock stamps compiled in
        static SmemAttrOnce once2;
        PAPR_CUDA_TRY(ensure_dyn_smem(once2, stack_kernel<true, true>, kStkMaxSmem));
        stack_kernel<true, true><<<grid, kStkThreads, smem, (cudaStream_t)stream>>>(p);
        return check_launch();
    }
    if (slope == 0.f) stack_kernel<true><<<grid, kStkThreads, smem, (cudaStream_t)stream>>>(p);
    else stack_kernel<false><<<grid, kStkThreads, smem, (cudaStream_t)stream>>>(p);
    return check_launch();
}
