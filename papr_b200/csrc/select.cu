// Stage a1: fused ray->point distance + top-K selection (reference models/model.py:258-283).
//
// One warp owns RPW rays of one view.  The view's points (as v = p - o, exactly rounded once per block)
// stream through shared memory in tiles; every lane evaluates one (ray, point) pair per ray per step with the
// reference's FP32 rounding sequence, so the R x P distance matrix never exists.  Each ray's K best
// (key, index) pairs live sorted across the lanes of the warp (lane j = j-th nearest); a candidate that beats
// the current K-th key is inserted with one ballot + one shuffle-up.  Ordering is (key, point index), key being
// the squared distance before the monotone sqrt, which refines the reference's ordering on sqrt(key).
//
// Exactness notes (SURVEY.md section 7 "Bit-exact top-K"):
//   * products and sums use __fmul_rn/__fadd_rn/__fsub_rn so ptxas cannot contract them into FMAs;
//   * s / den is the IEEE quotient: with r = RN(1/den), two Markstein corrections
//       q0 = s*r; q1 = fma(fma(-q0,den,s), r, q0); q = fma(fma(-q1,den,s), r, q1)
//     give RN(s/den) (q1 is faithful, then Markstein's theorem applies); tests/test_division.py checks the same
//     sequence against hardware division on 4e8 inputs;
//   * key = fma(Dz,Dz, fma(Dy,Dy, Dx*Dx)) is how torch's CPU norm kernel accumulates (pinned by make_golden.py).
#include "common.cuh"

namespace papr {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelTile = 2048;   // points per shared-memory tile (32 KB as float4)

template <int RPW>
__global__ void __launch_bounds__(kSelThreads)
select_topk_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                   const float *__restrict__ points, int64_t rays_per_view, int P, int K, float eps,
                   int32_t *__restrict__ idx_out, int blocks_per_view)
{
    __shared__ float4 tile[kSelTile];

    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;

    const float ox = rays_o[3 * view + 0], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];

    const int64_t ray0 = (int64_t)blk * (kSelWarps * RPW) + warp * RPW;   // first ray of this warp in the view
    float dx[RPW], dy[RPW], dz[RPW], den[RPW], rinv[RPW], thr[RPW], lk[RPW];
    int li[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        int64_t r = ray0 + j;
        if (r >= rays_per_view) r = rays_per_view - 1;   // duplicate the last ray; its result is not stored
        const float *d = rays_d + ((int64_t)view * rays_per_view + r) * 3;
        dx[j] = d[0]; dy[j] = d[1]; dz[j] = d[2];
        den[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx[j], dx[j]), __fmul_rn(dy[j], dy[j])),
                                     __fmul_rn(dz[j], dz[j])), eps);
        rinv[j] = __frcp_rn(den[j]);
        thr[j] = __int_as_float(0x7f800000);
        lk[j] = __int_as_float(0x7f800000);
        li[j] = -1;
    }

    for (int base = 0; base < P; base += kSelTile) {
        const int count = min(kSelTile, P - base);
        __syncthreads();
        for (int i = threadIdx.x; i < count; i += kSelThreads) {
            const float *p = points + (int64_t)(base + i) * 3;
            tile[i] = make_float4(__fsub_rn(p[0], ox), __fsub_rn(p[1], oy), __fsub_rn(p[2], oz), 0.f);
        }
        __syncthreads();

        for (int c = 0; c < count; c += 32) {
            const int pi = c + lane;
            const bool valid = pi < count;
            const float4 v = tile[valid ? pi : 0];
            const int pidx = base + pi;
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                const float s = __fadd_rn(__fadd_rn(__fmul_rn(v.x, dx[j]), __fmul_rn(v.y, dy[j])), __fmul_rn(v.z, dz[j]));
                const float q0 = __fmul_rn(s, rinv[j]);
                const float q1 = __fmaf_rn(__fmaf_rn(-q0, den[j], s), rinv[j], q0);
                const float t = __fmaf_rn(__fmaf_rn(-q1, den[j], s), rinv[j], q1);
                const float Dx = __fsub_rn(v.x, __fmul_rn(dx[j], t));
                const float Dy = __fsub_rn(v.y, __fmul_rn(dy[j], t));
                const float Dz = __fsub_rn(v.z, __fmul_rn(dz[j], t));
                float key = __fmaf_rn(Dz, Dz, __fmaf_rn(Dy, Dy, __fmul_rn(Dx, Dx)));
                if (!valid) key = __int_as_float(0x7f800000);
                unsigned m = __ballot_sync(full, key < thr[j]);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float ck = __shfl_sync(full, key, src);
                    const int ci = __shfl_sync(full, pidx, src);
                    if (ck < thr[j]) {
                        const int pos = __popc(__ballot_sync(full, lk[j] <= ck));
                        const float uk = __shfl_up_sync(full, lk[j], 1);
                        const int ui = __shfl_up_sync(full, li[j], 1);
                        if (lane == pos) { lk[j] = ck; li[j] = ci; }
                        else if (lane > pos) { lk[j] = uk; li[j] = ui; }
                        thr[j] = __shfl_sync(full, lk[j], K - 1);
                    }
                }
            }
        }
    }

#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int64_t r = ray0 + j;
        if (r < rays_per_view && lane < K)
            idx_out[((int64_t)view * rays_per_view + r) * K + lane] = li[j];
    }
}

}  // namespace papr

extern "C" int papr_select_topk(const float *rays_o, const float *rays_d, const float *points,
                                int64_t n_views, int64_t rays_per_view, int64_t P, int K, float eps,
                                int32_t *idx_out, void *stream)
{
    using namespace papr;
    if (!rays_o || !rays_d || !points || !idx_out) return PAPR_ERR_INVALID_ARGUMENT;
    if (K < 1 || K > 32 || P <= K || P > INT32_MAX || n_views < 0 || rays_per_view < 0) return PAPR_ERR_INVALID_ARGUMENT;
    if (n_views == 0 || rays_per_view == 0) return PAPR_OK;
    constexpr int RPW = 4;
    const int64_t rays_per_block = kSelWarps * RPW;
    const int64_t blocks_per_view = (rays_per_view + rays_per_block - 1) / rays_per_block;
    if (blocks_per_view * n_views > INT32_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    select_topk_kernel<RPW><<<(unsigned)(blocks_per_view * n_views), kSelThreads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, points, rays_per_view, (int)P, K, eps, idx_out, (int)blocks_per_view);
    return check_launch();
}
