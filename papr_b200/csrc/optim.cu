// the optimiser step of PAPR.step (reference models/model.py:439-446: one torch.optim.Adam per
// parameter group, each a handful of foreach launches) as ONE multi-tensor launch over a flat gradient / moment bucket,
// plus ONE batched launch that rebuilds every bf16 weight image the tensor-core kernels read (instead of one
// papr_pack_weight launch per layer and direction, 38 per training step).
#include "tc_common.cuh"

namespace papr {

constexpr int kAdamMaxGroups = 8;
constexpr int kAdamMaxTensors = 256;

struct AdamGroupDev { float lr, beta1, beta2, eps, wd, inv_bc1, inv_bc2_sqrt, enabled; };

struct AdamParams {
    const int64_t *offsets;     // [n_tensors + 1] element offsets of each tensor in the flat g / m / v buffers
    float *const *ptrs;         // [n_tensors] parameter storage (fp32, contiguous)
    const int32_t *group;       // [n_tensors]
    int n_tensors;
    int64_t total;
    const float *g;
    float *m, *v;
    float gscale;
    AdamGroupDev G[kAdamMaxGroups];
};

// torch.optim.Adam (single-tensor form, torch/optim/adam.py): m.lerp_(g, 1-b1); v = v*b2 + g*g*(1-b2);
// denom = sqrt(v)/sqrt(bc2) + eps; p += -(lr/bc1) * (m/denom).  weight_decay is the L2 form (g += wd*p).
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamParams p)
{
    __shared__ int64_t off_s[kAdamMaxTensors + 1];
    __shared__ float *ptr_s[kAdamMaxTensors];
    __shared__ int grp_s[kAdamMaxTensors];
    for (int i = threadIdx.x; i <= p.n_tensors; i += blockDim.x) off_s[i] = p.offsets[i];
    for (int i = threadIdx.x; i < p.n_tensors; i += blockDim.x) { ptr_s[i] = p.ptrs[i]; grp_s[i] = p.group[i]; }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int t = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
        if (i < off_s[t] || i >= off_s[t + 1]) {        // binary search for the tensor that holds element i
            int lo = 0, hi = p.n_tensors;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off_s[mid] <= i) lo = mid; else hi = mid; }
            t = lo;
        }
        const AdamGroupDev &G = p.G[grp_s[t]];
        if (G.enabled == 0.f) continue;
        float *w = ptr_s[t] + (i - off_s[t]);
        float g = p.g[i] * p.gscale;
        const float x = *w;
        if (G.wd != 0.f) g = __fmaf_rn(G.wd, x, g);
        float m = p.m[i], v = p.v[i];
        m = __fmaf_rn(g - m, 1.f - G.beta1, m);
        v = __fmaf_rn(g * g, 1.f - G.beta2, v * G.beta2);
        p.m[i] = m; p.v[i] = v;
        const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), G.inv_bc2_sqrt), G.eps);
        *w = __fmaf_rn(-(G.lr * G.inv_bc1), __fdiv_rn(m, denom), x);
    }
}

struct PackDesc {
    const float *w; int64_t ld; int32_t rows, cols, transpose, N, K, replicas; float scale; int32_t pad; int64_t rep_stride; uint8_t *img;
};

// one CTA column per descriptor (blockIdx.y), same element mapping as pack_weight_kernel
__global__ void __launch_bounds__(256) pack_batch_kernel(const PackDesc *__restrict__ descs)
{
    const PackDesc d = descs[blockIdx.y];
    const int kblk = (d.K + 63) / 64;
    const int total = kblk * d.N * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kb = i / (d.N * 64);
        const int n = (i / 64) % d.N;
        const int kk = i % 64;
        const int k = kb * 64 + kk;
        float v = 0.f;
        if (!d.transpose) { if (n < d.rows && k < d.cols) v = d.w[(size_t)n * d.ld + k]; }
        else              { if (k < d.rows && n < d.cols) v = d.w[(size_t)k * d.ld + n]; }
        const size_t off = (size_t)kb * d.N * 128 + (size_t)n * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4)) + (kk & 7) * 2;
        const __nv_bfloat16 b = __float2bfloat16(v * d.scale);
        for (int r = 0; r < d.replicas; ++r) *reinterpret_cast<__nv_bfloat16 *>(d.img + (size_t)r * d.rep_stride + off) = b;
    }
}

}  // namespace papr

extern "C" int papr_adam_step(const int64_t *offsets, float *const *param_ptrs, const int32_t *group_of, int n_tensors, int64_t total,
                              const float *grad, float *exp_avg, float *exp_avg_sq, const papr_adam_group *groups, int n_groups,
                              float grad_scale, void *stream)
{
    using namespace papr;
    if (!offsets || !param_ptrs || !group_of || !grad || !exp_avg || !exp_avg_sq || !groups) return PAPR_ERR_INVALID_ARGUMENT;
    if (n_tensors < 1 || n_tensors > kAdamMaxTensors || n_groups < 1 || n_groups > kAdamMaxGroups || total < 0) return PAPR_ERR_INVALID_ARGUMENT;
    if (total == 0) return PAPR_OK;
    AdamParams p;
    p.offsets = offsets; p.ptrs = param_ptrs; p.group = group_of; p.n_tensors = n_tensors; p.total = total;
    p.g = grad; p.m = exp_avg; p.v = exp_avg_sq; p.gscale = grad_scale;
    for (int i = 0; i < kAdamMaxGroups; ++i) p.G[i] = AdamGroupDev{0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_groups; ++i) {
        const papr_adam_group &h = groups[i];
        if (h.step < 1) { p.G[i].enabled = 0.f; continue; }
        const double bc1 = 1.0 - pow((double)h.beta1, (double)h.step), bc2 = 1.0 - pow((double)h.beta2, (double)h.step);
        p.G[i] = AdamGroupDev{h.lr, h.beta1, h.beta2, h.eps, h.weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), h.enabled ? 1.f : 0.f};
    }
    const int64_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
    adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_pack_weight_batch(const papr_pack_desc *descs_device, int n_descs, void *stream)
{
    using namespace papr;
    static_assert(sizeof(PackDesc) == sizeof(papr_pack_desc), "papr_pack_desc layout");
    if (!descs_device || n_descs < 1 || n_descs > 65535) return PAPR_ERR_INVALID_ARGUMENT;
    dim3 grid(16, n_descs);        // <= 256 x 256 elements per image: 16 blocks x 256 threads x 16 elements
    pack_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const PackDesc *)descs_device);
    return check_launch();
}
