// Between-step bookkeeping and the step before the path, on the device:
//   papr_knn            exact k-nearest-neighbour search for point growing (reference models/utils.py:9-109 add_points_knn:
//                       scipy KDTree on the host, float64 distances) -- brute force in float64, one warp per query, the
//                       k best kept sorted across the warp's lanes, ties broken by point index
//   papr_prune_compact  order-preserving stream compaction of the point tables (reference models/model.py:335-358:
//                       boolean-mask indexing of points / influence scores / features)
//   papr_generate_rays  pinhole ray generation for a patch of a view (reference dataset/utils.py:81-96 get_rays +
//                       dataset/dataset.py:19-25 origin scaling): removes the per-step H2D copy of rays
#include "common.cuh"
#include <float.h>
#include <limits.h>

namespace papr {

constexpr int kKnnWarps = 8;
constexpr int kKnnTile = 1024;

__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const float *__restrict__ points, int64_t P, const float *__restrict__ queries,
                                                             int64_t Q, int k, double *__restrict__ dist_out, int32_t *__restrict__ idx_out)
{
    __shared__ float tile[kKnnTile * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * kKnnWarps + warp;
    const bool active = q < Q;
    double qx = 0, qy = 0, qz = 0;
    if (active) { qx = queries[q * 3]; qy = queries[q * 3 + 1]; qz = queries[q * 3 + 2]; }
    double best_d = DBL_MAX;         // lane j: j-th smallest squared distance so far
    int best_i = INT_MAX;
    double kth = DBL_MAX; int kthi = INT_MAX;
    for (int64_t t0 = 0; t0 < P; t0 += kKnnTile) {
        const int n = (int)((P - t0) < kKnnTile ? (P - t0) : kKnnTile);
        __syncthreads();
        for (int i = threadIdx.x; i < n * 3; i += blockDim.x) tile[i] = points[t0 * 3 + i];
        __syncthreads();
        if (!active) continue;
        for (int base = 0; base < n; base += 32) {
            const int li = base + lane;
            double d2 = DBL_MAX; int gi = INT_MAX;
            if (li < n) {
                const double dx = qx - (double)tile[li * 3], dy = qy - (double)tile[li * 3 + 1], dz = qz - (double)tile[li * 3 + 2];
                d2 = dx * dx + dy * dy + dz * dz;
                gi = (int)(t0 + li);
            }
            unsigned mask = __ballot_sync(0xffffffffu, d2 < kth || (d2 == kth && gi < kthi));
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double nd = __shfl_sync(0xffffffffu, d2, src);
                const int ni = __shfl_sync(0xffffffffu, gi, src);
                if (!(nd < kth || (nd == kth && ni < kthi))) continue;
                const bool less = best_d < nd || (best_d == nd && best_i < ni);
                const int pos = __popc(__ballot_sync(0xffffffffu, less));
                const double pd = __shfl_up_sync(0xffffffffu, best_d, 1);
                const int pi = __shfl_up_sync(0xffffffffu, best_i, 1);
                if (lane > pos) { best_d = pd; best_i = pi; }
                else if (lane == pos) { best_d = nd; best_i = ni; }
                kth = __shfl_sync(0xffffffffu, best_d, k - 1);
                kthi = __shfl_sync(0xffffffffu, best_i, k - 1);
            }
        }
    }
    if (active && lane < k) {
        dist_out[q * k + lane] = sqrt(best_d);
        idx_out[q * k + lane] = best_i;
    }
}

__device__ __forceinline__ bool prune_keep(float v, float thresh, int keep_less) { return keep_less ? v < thresh : v > thresh; }

__global__ void __launch_bounds__(256) prune_count_kernel(const float *__restrict__ influ, int64_t P, float thresh, int keep_less, int *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int keep = (i < P && prune_keep(influ[i], thresh, keep_less)) ? 1 : 0;
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) prune_scan_kernel(int *__restrict__ counts, int nblocks, int64_t *__restrict__ total)
{
    __shared__ int part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
    int s = 0;
    for (int b = b0; b < b1; ++b) s += counts[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < 1024; ++i) { const int v = part[i]; part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int b = b0; b < b1; ++b) { const int v = counts[b]; counts[b] = run; run += v; }
}

__global__ void __launch_bounds__(256) prune_scatter_kernel(const float *__restrict__ points, const float *__restrict__ influ,
                                                            const float *__restrict__ feats, int F, int64_t P, float thresh, int keep_less,
                                                            const int *__restrict__ block_off, float *__restrict__ out_points,
                                                            float *__restrict__ out_influ, float *__restrict__ out_feats)
{
    __shared__ int warp_sums[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool keep = i < P && prune_keep(influ[i], thresh, keep_less);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    if (!keep) return;
    const int64_t dst = (int64_t)block_off[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
    out_points[dst * 3] = points[i * 3]; out_points[dst * 3 + 1] = points[i * 3 + 1]; out_points[dst * 3 + 2] = points[i * 3 + 2];
    out_influ[dst] = influ[i];
    if (feats) for (int f = 0; f < F; ++f) out_feats[dst * F + f] = feats[i * F + f];
}

// get_rays for the pixels [h0, h0+h) x [w0, w0+w) of an H x W view: pixel-centre direction (x, -y, -1) through
// linspace(0, W/focal, W+1), rotated by c2w[:3,:3], unit-normalised; origin = coord_scale * c2w[:3,3].
__global__ void __launch_bounds__(256) raygen_kernel(const float *__restrict__ c2w, int64_t n_views, int H, int W, float focal_x, float focal_y,
                                                     int h0, int w0, int h, int w, float coord_scale, float *__restrict__ rays_o,
                                                     float *__restrict__ rays_d)
{
    const int64_t per = (int64_t)h * w;
    const int64_t total = n_views * per;
    const float ex = (float)W / focal_x, ey = (float)H / focal_y;          // linspace end points
    const float sx = ex / (float)W, sy = ey / (float)H;                     // linspace steps (steps - 1 = W, H)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / per;
        const int r = (int)((i % per) / w) + h0, c = (int)(i % w) + w0;
        // torch.linspace: start + step*i in the first half, end - step*(steps-1-i) in the second
        const float lx = (c < (W + 1) / 2) ? __fmul_rn(sx, (float)c) : __fsub_rn(ex, __fmul_rn(sx, (float)(W - c)));
        const float ly = (r < (H + 1) / 2) ? __fmul_rn(sy, (float)r) : __fsub_rn(ey, __fmul_rn(sy, (float)(H - r)));
        const float x = __fadd_rn(__fsub_rn(lx, (float)((double)W / focal_x / 2)), __fdiv_rn(sx, 2.f));
        const float y = -__fadd_rn(__fsub_rn(ly, (float)((double)H / focal_y / 2)), __fdiv_rn(sy, 2.f));
        const float *m = c2w + v * 16;
        float d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            d[a] = __fadd_rn(__fadd_rn(__fmul_rn(x, m[a * 4]), __fmul_rn(y, m[a * 4 + 1])), __fmul_rn(-1.f, m[a * 4 + 2]));
        const float nrm = (float)sqrt((double)__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
        rays_d[i * 3] = __fdiv_rn(d[0], nrm); rays_d[i * 3 + 1] = __fdiv_rn(d[1], nrm); rays_d[i * 3 + 2] = __fdiv_rn(d[2], nrm);
        if (i % per == 0) {
            rays_o[v * 3] = coord_scale * m[3]; rays_o[v * 3 + 1] = coord_scale * m[7]; rays_o[v * 3 + 2] = coord_scale * m[11];
        }
    }
}

}  // namespace papr

extern "C" int papr_knn(const float *points, int64_t P, const float *queries, int64_t Q, int k, double *dist_out, int32_t *idx_out,
                        void *stream)
{
    using namespace papr;
    if (!points || !queries || !dist_out || !idx_out || P < 1 || Q < 0 || k < 1 || k > 32 || k > P) return PAPR_ERR_INVALID_ARGUMENT;
    if (P > INT_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    if (Q == 0) return PAPR_OK;
    const int64_t grid = (Q + kKnnWarps - 1) / kKnnWarps;
    knn_kernel<<<(unsigned)grid, kKnnWarps * 32, 0, (cudaStream_t)stream>>>(points, P, queries, Q, k, dist_out, idx_out);
    return check_launch();
}

extern "C" int papr_prune_compact(const float *points, const float *influ, const float *feats, int64_t P, int F, float thresh,
                                  int keep_less, float *out_points, float *out_influ, float *out_feats, int32_t *scratch,
                                  int64_t *n_kept, void *stream)
{
    using namespace papr;
    if (!points || !influ || !out_points || !out_influ || !scratch || !n_kept || P < 0 || F < 0 || (feats && !out_feats)) return PAPR_ERR_INVALID_ARGUMENT;
    if (P > (int64_t)INT_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int nblocks = (int)((P + 255) / 256);
    if (nblocks == 0) { PAPR_CUDA_TRY(cudaMemsetAsync(n_kept, 0, sizeof(int64_t), s)); return PAPR_OK; }
    prune_count_kernel<<<nblocks, 256, 0, s>>>(influ, P, thresh, keep_less, scratch);
    prune_scan_kernel<<<1, 1024, 0, s>>>(scratch, nblocks, n_kept);
    prune_scatter_kernel<<<nblocks, 256, 0, s>>>(points, influ, feats, F, P, thresh, keep_less, scratch, out_points, out_influ, out_feats);
    return check_launch();
}

extern "C" int papr_generate_rays(const float *c2w, int64_t n_views, int H, int W, float focal_x, float focal_y, int h0, int w0, int h,
                                  int w, float coord_scale, float *rays_o, float *rays_d, void *stream)
{
    using namespace papr;
    if (!c2w || !rays_o || !rays_d || n_views < 1 || H < 1 || W < 1 || h < 1 || w < 1 || h0 < 0 || w0 < 0 || h0 + h > H || w0 + w > W)
        return PAPR_ERR_INVALID_ARGUMENT;
    if (!(focal_x > 0.f) || !(focal_y > 0.f)) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t total = n_views * (int64_t)h * w;
    const int grid = (int)((total + 255) / 256 < kNumSMs * 8 ? (total + 255) / 256 : kNumSMs * 8);
    raygen_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c2w, n_views, H, W, focal_x, focal_y, h0, w0, h, w, coord_scale, rays_o, rays_d);
    return check_launch();
}
