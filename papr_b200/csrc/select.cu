// Stage a1: fused ray->point distance + top-K selection (reference models/model.py:258-283).
//
// One warp owns RPW rays of one view.  The view's points (as v = p - o, exactly rounded once per block)
// stream through shared memory in tiles; every lane evaluates one (ray, point) pair per ray per step, so the
// R x P distance matrix never exists.  Each ray's best (key, index) pairs live sorted across the lanes of the warp
// (lane j = j-th nearest); a candidate that beats the current last key is inserted with one ballot + one
// shuffle-up.  Final ordering is (key, point index), key being the reference's squared distance before the
// monotone sqrt, which refines the reference's ordering on sqrt(key).
//
// Exactness notes (SURVEY.md section 7 "Bit-exact top-K"):
//   * products and sums use __fmul_rn/__fadd_rn/__fsub_rn so ptxas cannot contract them into FMAs;
//   * s / den is the IEEE quotient: with r = RN(1/den), two Markstein corrections
//       q0 = s*r; q1 = fma(fma(-q0,den,s), r, q0); q = fma(fma(-q1,den,s), r, q1)
//     give RN(s/den) (q1 is faithful, then Markstein's theorem applies); tests/test_division.py checks the same
//     sequence against hardware division on 4e8 inputs;
//   * key = fma(Dz,Dz, fma(Dy,Dy, Dx*Dx)) is how torch's CPU norm kernel accumulates (pinned by make_golden.py).
#include "common.cuh"
#include <stdlib.h>

namespace papr {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelTile = 2048;   // points per shared-memory tile (32 KB as float4)

// ---------------------------------------------------------------------------------------------------------------
// Two-phase exact selection (the product path).
//   Phase 1 streams all points with a CHEAP key, a = |v x d|^2 + eps*|v|^2 (10 FP32 instructions per pair instead of
//   22), and keeps the 32 smallest per ray -- the warp is 32 lanes wide, so 32 candidates cost the same as K.
//   Phase 2 evaluates the reference-exact key (same rounding sequence as above) for those 32 candidates only and
//   ranks them by (key, index).
// The result is provably the reference's top-K whenever the K-th exact key, scaled by den, lies below
// a32 - err(a32), where a32 is the largest cheap key kept and err() bounds |cheap - exact*den| for ANY pair
// (derivation in DESIGN.md section 4: err(x) = 32 u V sqrt(x) + 128 u^2 V^2 + 8 u x, u = 2^-24, V = max|v| * |d|).
// Every point that was not kept has cheap key >= a32, hence exact key above the K-th one.  Rays that fail the test
// (exact ties at the boundary, degenerate clouds) are rescanned with the exact key -- rare, and still bit-exact.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float exact_key(float vx, float vy, float vz, float dx, float dy, float dz, float den, float rinv)
{
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(vx, dx), __fmul_rn(vy, dy)), __fmul_rn(vz, dz));
    const float q0 = __fmul_rn(s, rinv);
    const float q1 = __fmaf_rn(__fmaf_rn(-q0, den, s), rinv, q0);
    const float t = __fmaf_rn(__fmaf_rn(-q1, den, s), rinv, q1);
    const float Dx = __fsub_rn(vx, __fmul_rn(dx, t));
    const float Dy = __fsub_rn(vy, __fmul_rn(dy, t));
    const float Dz = __fsub_rn(vz, __fmul_rn(dz, t));
    return __fmaf_rn(Dz, Dz, __fmaf_rn(Dy, Dy, __fmul_rn(Dx, Dx)));
}

// sorted insert of (ck, ci) into a warp-wide list (lane j = j-th smallest); returns the new key of lane `last`
__device__ __forceinline__ float list_insert(float &lk, int &li, float ck, int ci, int lane, int last)
{
    const unsigned full = 0xffffffffu;
    const int pos = __popc(__ballot_sync(full, lk <= ck));
    const float uk = __shfl_up_sync(full, lk, 1);
    const int ui = __shfl_up_sync(full, li, 1);
    if (lane == pos) { lk = ck; li = ci; }
    else if (lane > pos) { lk = uk; li = ui; }
    return __shfl_sync(full, lk, last);
}

template <int RPW>
__global__ void __launch_bounds__(kSelThreads)
select_topk2_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                    const float *__restrict__ points, int64_t rays_per_view, int P, int K, float eps,
                    int32_t *__restrict__ idx_out, int blocks_per_view)
{
    __shared__ float4 tile[kSelTile];

    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);

    const float ox = rays_o[3 * view + 0], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];
    const int64_t ray0 = (int64_t)blk * (kSelWarps * RPW) + warp * RPW;
    float dx[RPW], dy[RPW], dz[RPW], thr[RPW], lk[RPW];
    int li[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        int64_t r = ray0 + j;
        if (r >= rays_per_view) r = rays_per_view - 1;
        const float *d = rays_d + ((int64_t)view * rays_per_view + r) * 3;
        dx[j] = d[0]; dy[j] = d[1]; dz[j] = d[2];
        thr[j] = INF; lk[j] = INF; li[j] = -1;
    }
    float wmax = 0.f;

    for (int base = 0; base < P; base += kSelTile) {
        const int count = min(kSelTile, P - base);
        __syncthreads();
        for (int i = threadIdx.x; i < count; i += kSelThreads) {
            const float *p = points + (int64_t)(base + i) * 3;
            const float vx = __fsub_rn(p[0], ox), vy = __fsub_rn(p[1], oy), vz = __fsub_rn(p[2], oz);
            tile[i] = make_float4(vx, vy, vz, fmaf(vz, vz, fmaf(vy, vy, vx * vx)));
        }
        __syncthreads();
        for (int c = 0; c < count; c += 32) {
            const int pi = c + lane;
            const bool valid = pi < count;
            const float4 v = tile[valid ? pi : 0];
            const int pidx = base + pi;
            const float ew = eps * v.w;
            wmax = fmaxf(wmax, v.w);
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                const float cx = fmaf(v.y, dz[j], -v.z * dy[j]);
                const float cy = fmaf(v.z, dx[j], -v.x * dz[j]);
                const float cz = fmaf(v.x, dy[j], -v.y * dx[j]);
                float a = fmaf(cx, cx, fmaf(cy, cy, fmaf(cz, cz, ew)));
                if (!valid) a = INF;
                unsigned m = __ballot_sync(full, a < thr[j]);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float ck = __shfl_sync(full, a, src);
                    const int ci = __shfl_sync(full, pidx, src);
                    if (ck < thr[j]) thr[j] = list_insert(lk[j], li[j], ck, ci, lane, 31);
                }
            }
        }
    }
    wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 16));
    wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 8));
    wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 4));
    wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 2));
    wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 1));

    // ---- phase 2: exact keys of the candidates, rank, safety test, (rare) exact rescan
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int64_t r = ray0 + j;
        if (r >= rays_per_view) continue;                       // warp-uniform
        const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx[j], dx[j]), __fmul_rn(dy[j], dy[j])),
                                              __fmul_rn(dz[j], dz[j])), eps);
        const float rinv = __frcp_rn(den);
        const int ci = li[j];
        float ek = INF;
        if (ci >= 0) {
            const float *p = points + (int64_t)ci * 3;
            ek = exact_key(__fsub_rn(p[0], ox), __fsub_rn(p[1], oy), __fsub_rn(p[2], oz), dx[j], dy[j], dz[j], den, rinv);
        }
        int rank = 0;
        for (int t = 0; t < 32; ++t) {
            const float ok = __shfl_sync(full, ek, t);
            const int oi = __shfl_sync(full, ci, t);
            rank += (ok < ek || (ok == ek && oi < ci)) ? 1 : 0;
        }
        // K-th exact key (rank K-1) scaled to cheap-key units, and the error-padded lower bound of everything dropped
        const unsigned who = __ballot_sync(full, rank == K - 1 && ci >= 0);
        const float eK = __shfl_sync(full, ek, who ? __ffs(who) - 1 : 0);
        const float a32 = thr[j];
        const float u = 5.9604645e-8f;
        const float V2 = wmax * den;                            // (max|v| * |d|)^2, den >= |d|^2
        const float err = 32.f * u * sqrtf(V2 * a32) + 128.f * u * u * V2 + 8.f * u * a32;
        const bool safe = (a32 == INF) || (who != 0 && a32 > 1e-8f * fmaxf(V2, 1.f) && eK * den * (1.f + 4.f * u) < a32 - err);
        if (safe) {
            if (ci >= 0 && rank < K) idx_out[((int64_t)view * rays_per_view + r) * K + rank] = ci;
        } else {
            // exact rescan of every point for this ray (same algorithm as select_topk_kernel, one ray per warp)
            float xk = INF, xt = INF;
            int xi = -1;
            for (int c = 0; c < P; c += 32) {
                const int pi = c + lane;
                float key = INF;
                if (pi < P) {
                    const float *p = points + (int64_t)pi * 3;
                    key = exact_key(__fsub_rn(p[0], ox), __fsub_rn(p[1], oy), __fsub_rn(p[2], oz), dx[j], dy[j], dz[j], den, rinv);
                }
                unsigned m = __ballot_sync(full, key < xt);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float ck = __shfl_sync(full, key, src);
                    const int cc = c + src;
                    if (ck < xt) xt = list_insert(xk, xi, ck, cc, lane, K - 1);
                }
            }
            if (lane < K) idx_out[((int64_t)view * rays_per_view + r) * K + lane] = xi;
        }
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
    static int rpw = 0;
    if (!rpw) { const char *e = getenv("PAPR_SELECT_RPW"); rpw = (e && atoi(e) == 8) ? 8 : 4; }
    const int64_t rays_per_block = kSelWarps * rpw;
    const int64_t blocks_per_view = (rays_per_view + rays_per_block - 1) / rays_per_block;
    if (blocks_per_view * n_views > INT32_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    const unsigned grid = (unsigned)(blocks_per_view * n_views);
    if (rpw == 4)
        select_topk2_kernel<4><<<grid, kSelThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, points, rays_per_view, (int)P, K, eps,
                                                                               idx_out, (int)blocks_per_view);
    else
        select_topk2_kernel<8><<<grid, kSelThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, points, rays_per_view, (int)P, K, eps,
                                                                               idx_out, (int)blocks_per_view);
    return check_launch();
}
