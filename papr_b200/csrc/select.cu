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

__device__ unsigned long long g_sel_fallbacks = 0;   // debug counter: rays that needed the exact rescan

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelTile = 2048;   // points per shared-memory tile (32 KB as float4)

// ---------------------------------------------------------------------------------------------------------------
// Two-phase exact selection (the product path).
//   Phase 1 streams all points with a CHEAP key, a = |v x d|^2 + eps*(|v|^2 - (v.d)^2 / den) (15 FP32 instructions per
//   pair instead of 22), and keeps the 32 smallest per ray -- the warp is 32 lanes wide, so 32 candidates cost the same
//   as K.  In exact arithmetic a IS the reference's squared distance times den = |d|^2 + eps: with t = (v.d)/den,
//   |v - t d|^2 den = |v x d|^2 + eps |v|^2 - eps (v.d)^2/den.  (Until the end of round 2 the last term was left out:
//   the cheap key then exceeds the exact one by up to eps |v|^2 -- 0.016 at the Caterpillar scale, where neighbouring
//   keys differ by 0.0005 -- which the 12 spare candidates of a 32-wide list absorbed in every test, but which the
//   error bound below did not cover; the grid kernel's tighter threshold exposed it.)
//   Phase 2 evaluates the reference-exact key (same rounding sequence as above) for those 32 candidates only and
//   ranks them by (key, index).
// The result is provably the reference's top-K whenever the K-th exact key, scaled by den, lies below
// a32 - err(a32), where a32 is the largest cheap key kept and err() bounds |cheap - exact*den| for ANY pair
// (err(x) = 32 u V sqrt(x) + 128 u^2 V^2 + 8 u x + 8 u eps max|v|^2, u = 2^-24, V = max|v| * |d|: the rounding of the
// cross product, of the sums, and of the eps term).
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

// A whole batch at once: the 32 keys of a batch (one per lane, +inf where there is no point) are sorted across the warp
// (bitonic network, 15 compare-exchange stages), laid against the list in reverse -- min(list[l], batch[31 - l]) is a
// bitonic sequence holding the 32 smallest of the union -- and merged in 5 more stages.  ~150 instructions whatever the
// number of newcomers, where the one-at-a-time insert costs ~14 instructions per newcomer on a chain of four dependent
// warp-wide operations: the first batches of a ray (empty or loose list: most of their points enter) go this way.
__device__ __forceinline__ void cmp_exchange(float &k, int &i, int stride, bool take_min)
{
    const float ok = __shfl_xor_sync(0xffffffffu, k, stride);
    const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
    if (take_min ? (ok < k) : (ok > k)) { k = ok; i = oi; }
}
__device__ __forceinline__ void list_merge_batch(float &lk, int &li, float bk, int bi, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
        const bool up = (lane & size) == 0 || size == 32;
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) cmp_exchange(bk, bi, stride, ((lane & stride) == 0) == up);
    }
    const float rk = __shfl_sync(0xffffffffu, bk, 31 - lane);
    const int ri = __shfl_sync(0xffffffffu, bi, 31 - lane);
    if (rk < lk) { lk = rk; li = ri; }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) cmp_exchange(lk, li, stride, (lane & stride) == 0);
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
    float dx[RPW], dy[RPW], dz[RPW], thr[RPW], lk[RPW], epd[RPW];
    int li[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        int64_t r = ray0 + j;
        if (r >= rays_per_view) r = rays_per_view - 1;
        const float *d = rays_d + ((int64_t)view * rays_per_view + r) * 3;
        dx[j] = d[0]; dy[j] = d[1]; dz[j] = d[2];
        thr[j] = INF; lk[j] = INF; li[j] = -1;
        epd[j] = eps / (fmaf(dz[j], dz[j], fmaf(dy[j], dy[j], dx[j] * dx[j])) + eps);
    }
    float wmax = 0.f;                 // max |v|^2 over the points this thread staged (reduced over the block below)

    for (int base = 0; base < P; base += kSelTile) {
        const int count = min(kSelTile, P - base);
        const int cpad = (count + 31) & ~31;
        __syncthreads();
        for (int i = threadIdx.x; i < cpad; i += kSelThreads) {
            if (i < count) {
                const float *p = points + (int64_t)(base + i) * 3;
                const float vx = __fsub_rn(p[0], ox), vy = __fsub_rn(p[1], oy), vz = __fsub_rn(p[2], oz);
                const float w = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
                wmax = fmaxf(wmax, w);
                tile[i] = make_float4(vx, vy, vz, eps * w);
            } else {
                tile[i] = make_float4(0.f, 0.f, 0.f, INF);         // padding: cheap key = +inf, never below a threshold
            }
        }
        __syncthreads();
        for (int c = 0; c < cpad; c += 32) {
            const float4 v = tile[c + lane];
            const int pidx = base + c + lane;
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                const float cx = fmaf(v.y, dz[j], -v.z * dy[j]);
                const float cy = fmaf(v.z, dx[j], -v.x * dz[j]);
                const float cz = fmaf(v.x, dy[j], -v.y * dx[j]);
                const float sd = fmaf(v.x, dx[j], fmaf(v.y, dy[j], v.z * dz[j]));
                const float a = fmaf(-epd[j], sd * sd, fmaf(cx, cx, fmaf(cy, cy, fmaf(cz, cz, v.w))));
                if (__any_sync(full, a < thr[j])) {
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
    }
    {
        __shared__ float wred[kSelWarps];
        wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 16));
        wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 8));
        wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 4));
        wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 2));
        wmax = fmaxf(wmax, __shfl_xor_sync(full, wmax, 1));
        if (lane == 0) wred[warp] = wmax;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSelWarps; ++i) wmax = fmaxf(wmax, wred[i]);
    }

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
        const float err = 32.f * u * sqrtf(V2 * a32) + 128.f * u * u * V2 + 8.f * u * a32 + 8.f * u * eps * wmax;
        const bool safe = (a32 == INF) || (who != 0 && a32 > 1e-8f * fmaxf(V2, 1.f) && eK * den * (1.f + 4.f * u) < a32 - err);
        if (safe) {
            if (ci >= 0 && rank < K) idx_out[((int64_t)view * rays_per_view + r) * K + rank] = ci;
        } else {
            // exact rescan of every point for this ray (same algorithm as select_topk_kernel, one ray per warp)
            if (lane == 0) atomicAdd(&g_sel_fallbacks, 1ull);
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

// ---------------------------------------------------------------------------------------------------------------
// Culled variant (the product path for P >= 1024): the same two-phase exact selection, but the points arrive sorted
// along a Morton curve in groups of 32 with a bounding sphere per group, and a warp skips a whole group when a
// conservative lower bound of the cheap key over (every ray of the warp) x (every point of the group) cannot beat the
// warp's current thresholds.  Skipping such a group never changes the candidate lists (an insertion needs key < thr), so
// the result is identical to the unculled scan.
//   bound: every ray direction d of the warp lies within angle rho of the warp's mean direction dc; for a sphere
//   (centre vc relative to the ray origin, radius r) the distance from any of its points to any of those lines is at
//   least  |vc x dc| cos(rho) - |vc . dc| sin(rho) - r  =: lb  (lines are unoriented, hence the absolute value), and
//   the cheap key is >= |d|^2 dist^2 >= dmin^2 lb^2.  Floating-point slack: cos(rho) is lowered / sin(rho), r raised
//   by 1e-6 relative and the comparison keeps a 0.2% margin.
// ---------------------------------------------------------------------------------------------------------------
// lexicographic (key, index) insert for lists whose candidates do not arrive in index order
__device__ __forceinline__ float list_insert_lex(float &lk, int &li, float ck, int ci, int lane, int last)
{
    const unsigned full = 0xffffffffu;
    const int pos = __popc(__ballot_sync(full, lk < ck || (lk == ck && li < ci && li >= 0)));
    const float uk = __shfl_up_sync(full, lk, 1);
    const int ui = __shfl_up_sync(full, li, 1);
    if (lane == pos) { lk = ck; li = ci; }
    else if (lane > pos) { lk = uk; li = ui; }
    return __shfl_sync(full, lk, last);
}

template <int RPW>
__global__ void __launch_bounds__(kSelThreads)
select_topk3_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                    const float *__restrict__ spts /* (P_pad,3) sorted, padded with far points */,
                    const int32_t *__restrict__ perm /* (P_pad) original index, -1 for padding */,
                    const float4 *__restrict__ spheres /* (P_pad/32) centre + radius */,
                    const float4 *__restrict__ spheres8 /* (P_pad/256) spheres around 8 consecutive groups */,
                    int64_t rays_per_view, int P_pad, int K, float eps, const float *__restrict__ pmax_ptr,
                    int32_t *__restrict__ idx_out, int blocks_per_view)
{
    __shared__ float4 tile[kSelTile];
    __shared__ float4 sph[kSelTile / 32];
    __shared__ float4 sph8[kSelTile / 256];

    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);

    const float ox = rays_o[3 * view + 0], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];
    const int64_t ray0 = (int64_t)blk * (kSelWarps * RPW) + warp * RPW;
    float dx[RPW], dy[RPW], dz[RPW], thr[RPW], lk[RPW], epd[RPW];
    int li[RPW];
    float cxs = 0.f, cys = 0.f, czs = 0.f, dmin2 = INF;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        int64_t r = ray0 + j;
        if (r >= rays_per_view) r = rays_per_view - 1;
        const float *d = rays_d + ((int64_t)view * rays_per_view + r) * 3;
        dx[j] = d[0]; dy[j] = d[1]; dz[j] = d[2];
        thr[j] = INF; lk[j] = INF; li[j] = -1;
        const float n2 = dx[j] * dx[j] + dy[j] * dy[j] + dz[j] * dz[j];
        epd[j] = eps / (n2 + eps);
        const float rn = rsqrtf(fmaxf(n2, 1e-30f));
        cxs += dx[j] * rn; cys += dy[j] * rn; czs += dz[j] * rn;
        dmin2 = fminf(dmin2, n2);
    }
    // mean direction of the warp's rays and the (padded) angle that contains all of them
    float cosr = 1.f, sinr = 1.f;
    bool can_cull;
    {
        const float n2 = cxs * cxs + cys * cys + czs * czs;
        can_cull = n2 > 1e-12f && dmin2 > 1e-20f;
        const float rn = rsqrtf(fmaxf(n2, 1e-30f));
        cxs *= rn; cys *= rn; czs *= rn;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const float n = rsqrtf(fmaxf(dx[j] * dx[j] + dy[j] * dy[j] + dz[j] * dz[j], 1e-30f));
            cosr = fminf(cosr, (cxs * dx[j] + cys * dy[j] + czs * dz[j]) * n);
        }
        cosr = cosr * (1.f - 2e-6f) - 2e-6f;
        can_cull = can_cull && cosr > 0.f;
        sinr = sqrtf(fmaxf(0.f, 1.f - cosr * cosr)) * (1.f + 2e-6f) + 2e-6f;
        dmin2 *= 0.998f;                       // the comparison margin
    }

    for (int base = 0; base < P_pad; base += kSelTile) {
        const int count = min(kSelTile, P_pad - base);      // multiple of 32
        __syncthreads();
        for (int i = threadIdx.x; i < count; i += kSelThreads) {
            const float *p = spts + (int64_t)(base + i) * 3;
            const float vx = __fsub_rn(p[0], ox), vy = __fsub_rn(p[1], oy), vz = __fsub_rn(p[2], oz);
            tile[i] = make_float4(vx, vy, vz, fmaf(vz, vz, fmaf(vy, vy, vx * vx)));
        }
        for (int i = threadIdx.x; i < count / 32; i += kSelThreads) {
            const float4 sp = spheres[base / 32 + i];
            sph[i] = make_float4(sp.x - ox, sp.y - oy, sp.z - oz, sp.w);
        }
        for (int i = threadIdx.x; i < (count + 255) / 256; i += kSelThreads) {
            const float4 sp = spheres8[base / 256 + i];
            sph8[i] = make_float4(sp.x - ox, sp.y - oy, sp.z - oz, sp.w);
        }
        __syncthreads();
        for (int c = 0; c < count; c += 32) {
            if (can_cull && (c & 255) == 0) {       // first the sphere around the next 8 groups (256 points)
                const float4 sp = sph8[c >> 8];
                const float dotc = fabsf(sp.x * cxs + sp.y * cys + sp.z * czs);
                const float kx = sp.y * czs - sp.z * cys, ky = sp.z * cxs - sp.x * czs, kz = sp.x * cys - sp.y * cxs;
                const float crn = sqrtf(kx * kx + ky * ky + kz * kz);
                const float lb = crn * cosr * (1.f - 4e-6f) - dotc * sinr - sp.w;
                float tmax = thr[0];
#pragma unroll
                for (int j = 1; j < RPW; ++j) tmax = fmaxf(tmax, thr[j]);
                if (lb > 0.f && lb * lb * dmin2 > tmax) { c += 224; continue; }     // warp-uniform: skip all 8 groups
            }
            if (can_cull) {
                const float4 sp = sph[c >> 5];
                const float dotc = fabsf(sp.x * cxs + sp.y * cys + sp.z * czs);
                const float kx = sp.y * czs - sp.z * cys, ky = sp.z * cxs - sp.x * czs, kz = sp.x * cys - sp.y * cxs;
                const float crn = sqrtf(kx * kx + ky * ky + kz * kz);
                const float lb = crn * cosr * (1.f - 4e-6f) - dotc * sinr - sp.w;
                float tmax = thr[0];
#pragma unroll
                for (int j = 1; j < RPW; ++j) tmax = fmaxf(tmax, thr[j]);
                if (lb > 0.f && lb * lb * dmin2 > tmax) continue;       // warp-uniform
            }
            const float4 v = tile[c + lane];
            const int pidx = base + c + lane;
            const float ew = eps * v.w;
#pragma unroll
            for (int j = 0; j < RPW; ++j) {
                const float cx = fmaf(v.y, dz[j], -v.z * dy[j]);
                const float cy = fmaf(v.z, dx[j], -v.x * dz[j]);
                const float cz = fmaf(v.x, dy[j], -v.y * dx[j]);
                const float sd = fmaf(v.x, dx[j], fmaf(v.y, dy[j], v.z * dz[j]));
                const float a = fmaf(-epd[j], sd * sd, fmaf(cx, cx, fmaf(cy, cy, fmaf(cz, cz, ew))));
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
    // max |v|^2 over all points of the view, conservatively: (|o| + max|p|)^2
    const float onorm = sqrtf(ox * ox + oy * oy + oz * oz);
    const float pmax = *pmax_ptr * 1.0001f;
    const float wmax = (onorm + pmax) * (onorm + pmax) * 1.0001f;

#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int64_t r = ray0 + j;
        if (r >= rays_per_view) continue;                       // warp-uniform
        const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx[j], dx[j]), __fmul_rn(dy[j], dy[j])),
                                              __fmul_rn(dz[j], dz[j])), eps);
        const float rinv = __frcp_rn(den);
        const int si = li[j];                                   // position in the sorted array
        int ci = -1;                                            // original point index
        float ek = INF;
        if (si >= 0) {
            ci = perm[si];
            if (ci >= 0) {
                const float *p = spts + (int64_t)si * 3;
                ek = exact_key(__fsub_rn(p[0], ox), __fsub_rn(p[1], oy), __fsub_rn(p[2], oz), dx[j], dy[j], dz[j], den, rinv);
            }
        }
        int rank = 0;
        for (int t = 0; t < 32; ++t) {
            const float ok = __shfl_sync(full, ek, t);
            const int oi = __shfl_sync(full, ci, t);
            rank += (ci >= 0 && oi >= 0 && (ok < ek || (ok == ek && oi < ci))) ? 1 : 0;
        }
        const unsigned who = __ballot_sync(full, rank == K - 1 && ci >= 0);
        const float eK = __shfl_sync(full, ek, who ? __ffs(who) - 1 : 0);
        const float a32 = thr[j];
        const float u = 5.9604645e-8f;
        const float V2 = wmax * den;
        const float err = 32.f * u * sqrtf(V2 * a32) + 128.f * u * u * V2 + 8.f * u * a32 + 8.f * u * eps * wmax;
        const bool finite32 = a32 < 1e30f;                      // padding points carry huge finite keys
        const bool safe = who != 0 && (!finite32 || (a32 > 1e-8f * fmaxf(V2, 1.f) && eK * den * (1.f + 4.f * u) < a32 - err));
        if (safe) {
            if (ci >= 0 && rank < K) idx_out[((int64_t)view * rays_per_view + r) * K + rank] = ci;
        } else {
            if (lane == 0) atomicAdd(&g_sel_fallbacks, 1ull);
            float xk = INF, xt = INF;
            int xi = -1;
            for (int c = 0; c < P_pad; c += 32) {
                const int pi = c + lane;
                const int oi = perm[pi];
                float key = INF;
                if (oi >= 0) {
                    const float *p = spts + (int64_t)pi * 3;
                    key = exact_key(__fsub_rn(p[0], ox), __fsub_rn(p[1], oy), __fsub_rn(p[2], oz), dx[j], dy[j], dz[j], den, rinv);
                }
                unsigned m = __ballot_sync(full, key <= xt && oi >= 0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float ck = __shfl_sync(full, key, src);
                    const int cc = __shfl_sync(full, oi, src);
                    const float lastk = __shfl_sync(full, xk, K - 1);
                    const int lasti = __shfl_sync(full, xi, K - 1);
                    if (ck < lastk || (ck == lastk && (lasti < 0 || cc < lasti))) xt = list_insert_lex(xk, xi, ck, cc, lane, K - 1);
                }
            }
            if (lane < K) idx_out[((int64_t)view * rays_per_view + r) * K + lane] = xi;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Screen-space grid variant (the product path for all but tiny clouds): the same two-phase exact selection, but a warp
// only looks at the points that can matter for its rays.
//   All rays of a view leave one origin o, so the distance from a point p (v = p - o) to the line through o with
//   direction d depends on the two DIRECTIONS only: in a camera frame (e1, e2, c) with c the view's mean ray direction,
//   write v = w3 (g1, g2, 1) and d ~ (h1, h2, 1) (gnomonic coordinates g = (v.e1, v.e2)/(v.c), h likewise).  Then
//       dist(p, line) = |v x d| / |d| = |w3| |(g,1) x (h,1)| / |(h,1)|  >=  |w3| |g - h| / sqrt(1 + |h|^2),
//   because the first two components of (g,1) x (h,1) are (g2 - h2, h1 - g1).  The host bins the points of every view
//   on a G x G grid over the gnomonic extent of the view's rays (border cells extend to infinity, points with v.c ~ 0
//   make their cell unboundable), sorts them by cell and records each cell's smallest |w3|.  A warp walks the cells in
//   square rings around the cell under its rays and skips a cell -- or stops altogether -- when
//       |d|_min^2 * zmin^2 * dist2D(cell box, box of the warp's h)^2 / (1 + max|h|^2),   less rounding slack,
//   exceeds the largest of its rays' current thresholds.  A skipped point could not have been inserted (insertion needs
//   cheap key < threshold, thresholds only fall), so the candidate lists are those of the full scan up to ties at the
//   32nd place, which the phase-2 safety test covers: the result is bit-identical to the plain kernel's.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGridViewFloats = 20;   // per view: e1(3) e2(3) c(3) gmin(2) cell(2) inv_cell(2) zmin_all wmax pad(3)

template <int RPW>
__global__ void __launch_bounds__(kSelThreads)
select_grid_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                   const float4 *__restrict__ sv /* (n_views*P): v = p - o and eps*|v|^2, sorted by cell, per view */,
                   const int32_t *__restrict__ perm /* (n_views*P) original point index */,
                   const int4 *__restrict__ cells /* (n_views*G*G): start, end, zmin bits, 0 */,
                   const float *__restrict__ views /* (n_views, kGridViewFloats) */,
                   int64_t rays_per_view, int P, int G, int K, float eps, int32_t *__restrict__ idx_out, int blocks_per_view,
                   int last /* lane whose key is a ray's threshold: K .. 31 */, int merge_min /* newcomers per batch from which the batch is merged */)
{
    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);

    const float *vp = views + (int64_t)view * kGridViewFloats;
    const float e1x = vp[0], e1y = vp[1], e1z = vp[2], e2x = vp[3], e2y = vp[4], e2z = vp[5], ccx = vp[6], ccy = vp[7], ccz = vp[8];
    const float gminx = vp[9], gminy = vp[10], cellx = vp[11], celly = vp[12], icx = vp[13], icy = vp[14];
    const float zmin_all = vp[15], wmax = vp[16];
    const int4 *vcells = cells + (int64_t)view * G * G;

    const int64_t ray0 = (int64_t)blk * (kSelWarps * RPW) + warp * RPW;
    if (ray0 >= rays_per_view) return;                          // warp-uniform; no block-wide barrier below
    float dx[RPW], dy[RPW], dz[RPW], thr[RPW], lk[RPW], epd[RPW];
    int li[RPW];
    float bx0 = INF, bx1 = -INF, by0 = INF, by1 = -INF, h2max = 0.f, dmin2 = INF, dmax2 = 0.f;
    bool can_cull = true;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        int64_t r = ray0 + j;
        if (r >= rays_per_view) r = rays_per_view - 1;
        const float *d = rays_d + ((int64_t)view * rays_per_view + r) * 3;
        dx[j] = d[0]; dy[j] = d[1]; dz[j] = d[2];
        thr[j] = INF; lk[j] = INF; li[j] = -1;
        const float n2 = dx[j] * dx[j] + dy[j] * dy[j] + dz[j] * dz[j];
        epd[j] = eps / (n2 + eps);
        const float w3 = dx[j] * ccx + dy[j] * ccy + dz[j] * ccz;
        can_cull = can_cull && (w3 * w3 > 0.04f * n2) && w3 > 0.f && n2 > 1e-30f && n2 < 1e30f;
        const float iw = 1.f / w3;
        const float hx = (dx[j] * e1x + dy[j] * e1y + dz[j] * e1z) * iw, hy = (dx[j] * e2x + dy[j] * e2y + dz[j] * e2z) * iw;
        bx0 = fminf(bx0, hx); bx1 = fmaxf(bx1, hx); by0 = fminf(by0, hy); by1 = fmaxf(by1, hy);
        h2max = fmaxf(h2max, hx * hx + hy * hy);
        dmin2 = fminf(dmin2, n2); dmax2 = fmaxf(dmax2, n2);
    }
    can_cull = can_cull && zmin_all > 0.f && bx1 - bx0 < 1e6f && by1 - by0 < 1e6f;
    // rounding slack of the gnomonic coordinates (points: host fp32, rays: above), absolute, per axis
    const float ext = fmaxf(fmaxf(fabsf(gminx), fabsf(gminx + cellx * G)), fmaxf(fabsf(gminy), fabsf(gminy + celly * G)));
    const float slack = 4e-6f * (1.f + fmaxf(ext, sqrtf(h2max)));
    const float scale = can_cull ? dmin2 / (1.f + h2max) * 0.998f : 0.f;
    const float u = 5.9604645e-8f;
    const float V2c = wmax * dmax2;                              // (max|v| |d|)^2 for the cheap key's rounding-error bound
    // lower bound of the COMPUTED cheap key of any point at gnomonic distance >= dist of a cell with smallest depth z
    auto cell_bound = [&](float z, float dist2) -> float {
        const float b = z * z * dist2 * scale;
        return b - (32.f * u * sqrtf(V2c * b) + 128.f * u * u * V2c + 8.f * u * b + 8.f * u * eps * wmax);
    };
    // The threshold of a ray is the key in lane `last` -- by default the (K + 2)-nd smallest cheap key seen, not the 32nd: every point outside lanes 0..last of
    // the list has a cheap key >= the final threshold (rejected points had key >= the threshold of their time, thresholds
    // only fall, and an entry pushed beyond lane `last` is no smaller than lane `last`), which is all phase 2's safety test
    // needs.  The tighter threshold means fewer insertions and an earlier end of the ring walk; the price is a smaller
    // margin between the K-th exact key and the threshold -- one neighbour's gap instead of twelve at K = 20 -- i.e. an
    // exact rescan where the (K+1)-st and (K+2)-nd neighbours tie with the K-th to 1e-6 (lattices, duplicates).
    int hcx = G >> 1, hcy = G >> 1;
    if (can_cull) {
        hcx = min(max((int)floorf((0.5f * (bx0 + bx1) - gminx) * icx), 0), G - 1);
        hcy = min(max((int)floorf((0.5f * (by0 + by1) - gminy) * icy), 0), G - 1);
    }
    float tmax = INF;

    for (int r = 0; r < 2 * G; ++r) {
        if (r > 0) {
            const int q = r - 1;                                 // rings 0..q have been visited: cells [hcx-q, hcx+q] x [hcy-q, hcy+q]
            const bool openL = hcx - q > 0, openR = hcx + q < G - 1, openT = hcy - q > 0, openB = hcy + q < G - 1;
            if (!(openL || openR || openT || openB)) break;      // the whole grid has been visited
            if (can_cull) {
                float lb = INF;
                if (openL) lb = fminf(lb, bx0 - (gminx + cellx * (float)(hcx - q)));
                if (openR) lb = fminf(lb, (gminx + cellx * (float)(hcx + q + 1)) - bx1);
                if (openT) lb = fminf(lb, by0 - (gminy + celly * (float)(hcy - q)));
                if (openB) lb = fminf(lb, (gminy + celly * (float)(hcy + q + 1)) - by1);
                lb = fmaxf(lb - slack, 0.f);
                if (cell_bound(zmin_all, lb * lb) > tmax) break; // nothing outside the visited square can be inserted any more
            }
        }
        const int n = r ? 8 * r : 1;
        const int x0 = hcx - r, x1 = hcx + r, y0 = hcy - r, y1 = hcy + r;
        for (int t0 = 0; t0 < n; t0 += 32) {
            const int t = t0 + lane;
            int cx = hcx, cy = hcy;
            if (r) {
                const int side = t / (2 * r), k = t - side * 2 * r;
                cx = side == 0 ? x0 + k : side == 1 ? x1 : side == 2 ? x1 - k : x0;
                cy = side == 0 ? y0 : side == 1 ? y0 + k : side == 2 ? y1 : y1 - k;
            }
            int start = 0, end = 0;
            float bnd = -INF;
            if (t < n && cx >= 0 && cx < G && cy >= 0 && cy < G) {
                const int4 m = __ldg(vcells + cy * G + cx);
                start = m.x; end = m.y;
                if (can_cull && end > start) {
                    const float z = __int_as_float(m.z);
                    const float cxa = (cx == 0) ? -INF : gminx + cellx * (float)cx, cxb = (cx == G - 1) ? INF : gminx + cellx * (float)(cx + 1);
                    const float cya = (cy == 0) ? -INF : gminy + celly * (float)cy, cyb = (cy == G - 1) ? INF : gminy + celly * (float)(cy + 1);
                    const float ddx = fmaxf(fmaxf(cxa - bx1, bx0 - cxb) - slack, 0.f);
                    const float ddy = fmaxf(fmaxf(cya - by1, by0 - cyb) - slack, 0.f);
                    bnd = cell_bound(z, ddx * ddx + ddy * ddy);
                }
            }
            unsigned todo = __ballot_sync(full, end > start && !(bnd > tmax));
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                if (__shfl_sync(full, bnd, src) > tmax) continue;            // the thresholds have fallen since the ballot
                const int cs = __shfl_sync(full, start, src), ce = __shfl_sync(full, end, src);
                for (int c = cs; c < ce; c += 32) {
                    const int pi = c + lane;
                    float4 v = make_float4(0.f, 0.f, 0.f, INF);
                    if (pi < ce) v = __ldg(sv + (int64_t)view * P + pi);
#pragma unroll
                    for (int j = 0; j < RPW; ++j) {
                        const float kx = fmaf(v.y, dz[j], -v.z * dy[j]);
                        const float ky = fmaf(v.z, dx[j], -v.x * dz[j]);
                        const float kz = fmaf(v.x, dy[j], -v.y * dx[j]);
                        const float sd = fmaf(v.x, dx[j], fmaf(v.y, dy[j], v.z * dz[j]));
                        const float a = fmaf(-epd[j], sd * sd, fmaf(kx, kx, fmaf(ky, ky, fmaf(kz, kz, v.w))));
                        unsigned m = __ballot_sync(full, a < thr[j]);
                        if (__popc(m) >= merge_min) {                        // many newcomers: sort the batch and merge
                            list_merge_batch(lk[j], li[j], a, pi < ce ? pi : -1, lane);
                            thr[j] = __shfl_sync(full, lk[j], last);
                            continue;
                        }
                        while (m) {
                            const int s2 = __ffs(m) - 1;
                            m &= m - 1;
                            const float ck = __shfl_sync(full, a, s2);
                            const int ci = c + s2;
                            if (ck < thr[j]) thr[j] = list_insert(lk[j], li[j], ck, ci, lane, last);
                        }
                    }
                }
                tmax = thr[0];
#pragma unroll
                for (int j = 1; j < RPW; ++j) tmax = fmaxf(tmax, thr[j]);
            }
        }
    }

    // ---- phase 2: exact keys of the candidates, rank by (key, original index), safety test, (rare) exact rescan
    const float4 *vsv = sv + (int64_t)view * P;
    const int32_t *vperm = perm + (int64_t)view * P;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int64_t r = ray0 + j;
        if (r >= rays_per_view) continue;                       // warp-uniform
        const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx[j], dx[j]), __fmul_rn(dy[j], dy[j])),
                                              __fmul_rn(dz[j], dz[j])), eps);
        const float rinv = __frcp_rn(den);
        const int si = li[j];                                   // position in the view's sorted array
        int ci = -1;
        float ek = INF;
        if (si >= 0) {
            ci = vperm[si];
            const float4 v = vsv[si];
            ek = exact_key(v.x, v.y, v.z, dx[j], dy[j], dz[j], den, rinv);
        }
        int rank = 0;
        for (int t = 0; t < 32; ++t) {
            const float ok = __shfl_sync(full, ek, t);
            const int oi = __shfl_sync(full, ci, t);
            rank += (ci >= 0 && oi >= 0 && (ok < ek || (ok == ek && oi < ci))) ? 1 : 0;
        }
        const unsigned who = __ballot_sync(full, rank == K - 1 && ci >= 0);
        const float eK = __shfl_sync(full, ek, who ? __ffs(who) - 1 : 0);
        const float a32 = thr[j];
        const float V2 = wmax * den;
        const float err = 32.f * u * sqrtf(V2 * a32) + 128.f * u * u * V2 + 8.f * u * a32 + 8.f * u * eps * wmax;
        const bool safe = who != 0 && ((a32 == INF) || (a32 > 1e-8f * fmaxf(V2, 1.f) && eK * den * (1.f + 4.f * u) < a32 - err));
        if (safe) {
            if (ci >= 0 && rank < K) idx_out[((int64_t)view * rays_per_view + r) * K + rank] = ci;
        } else {
            if (lane == 0) atomicAdd(&g_sel_fallbacks, 1ull);
            float xk = INF, xt = INF;
            int xi = -1;
            for (int c = 0; c < P; c += 32) {
                const int pi = c + lane;
                int oi = -1;
                float key = INF;
                if (pi < P) {
                    oi = vperm[pi];
                    const float4 v = vsv[pi];
                    key = exact_key(v.x, v.y, v.z, dx[j], dy[j], dz[j], den, rinv);
                }
                unsigned m = __ballot_sync(full, key <= xt && oi >= 0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float ck = __shfl_sync(full, key, src);
                    const int cc = __shfl_sync(full, oi, src);
                    const float lastk = __shfl_sync(full, xk, K - 1);
                    const int lasti = __shfl_sync(full, xi, K - 1);
                    if (ck < lastk || (ck == lastk && (lasti < 0 || cc < lasti))) xt = list_insert_lex(xk, xi, ck, cc, lane, K - 1);
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

extern "C" int papr_select_topk_sorted(const float *rays_o, const float *rays_d, const float *sorted_points, const int32_t *perm,
                                       const float *spheres, const float *spheres8, int64_t n_views, int64_t rays_per_view, int64_t P_pad, int64_t P,
                                       int K, float eps, const float *pmax, int32_t *idx_out, void *stream)
{
    using namespace papr;
    if (!rays_o || !rays_d || !sorted_points || !perm || !spheres || !spheres8 || !idx_out) return PAPR_ERR_INVALID_ARGUMENT;
    if (K < 1 || K > 32 || P <= K || P_pad < P || P_pad % 32 || P_pad > INT32_MAX || n_views < 0 || rays_per_view < 0 || !pmax)
        return PAPR_ERR_INVALID_ARGUMENT;
    if (n_views == 0 || rays_per_view == 0) return PAPR_OK;
    constexpr int RPW = 4;
    const int64_t rays_per_block = kSelWarps * RPW;
    const int64_t blocks_per_view = (rays_per_view + rays_per_block - 1) / rays_per_block;
    if (blocks_per_view * n_views > INT32_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    select_topk3_kernel<RPW><<<(unsigned)(blocks_per_view * n_views), kSelThreads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, sorted_points, perm, (const float4 *)spheres, (const float4 *)spheres8, rays_per_view, (int)P_pad, K, eps, pmax, idx_out,
        (int)blocks_per_view);
    return check_launch();
}

extern "C" int papr_select_topk_grid(const float *rays_o, const float *rays_d, const void *sorted_v, const int32_t *perm,
                                     const int32_t *cells, const float *view_params, int64_t n_views, int64_t rays_per_view, int64_t P,
                                     int G, int K, float eps, int32_t *idx_out, void *stream)
{
    using namespace papr;
    if (!rays_o || !rays_d || !sorted_v || !perm || !cells || !view_params || !idx_out) return PAPR_ERR_INVALID_ARGUMENT;
    if (K < 1 || K > 32 || P <= K || P > INT32_MAX || G < 1 || G > 1024 || n_views < 0 || rays_per_view < 0) return PAPR_ERR_INVALID_ARGUMENT;
    if (n_views == 0 || rays_per_view == 0) return PAPR_OK;
    constexpr int RPW = 4;
    const int64_t rays_per_block = kSelWarps * RPW;
    const int64_t blocks_per_view = (rays_per_view + rays_per_block - 1) / rays_per_block;
    if (blocks_per_view * n_views > INT32_MAX) return PAPR_ERR_INVALID_ARGUMENT;
    // threshold lane: the (K+2)-nd candidate by default; PAPR_SELECT_LAST=31 restores the widest list (A/B switch)
    int last = K + 1;
    if (const char *e = getenv("PAPR_SELECT_LAST")) last = atoi(e);
    last = last < K ? K : last;
    last = last > 31 ? 31 : last;
    int merge_min = 16;                                 // PAPR_SELECT_MERGE=33 switches the batch merge off (A/B switch)
    if (const char *e = getenv("PAPR_SELECT_MERGE")) merge_min = atoi(e);
    select_grid_kernel<RPW><<<(unsigned)(blocks_per_view * n_views), kSelThreads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, (const float4 *)sorted_v, perm, (const int4 *)cells, view_params, rays_per_view, (int)P, G, K, eps, idx_out,
        (int)blocks_per_view, last, merge_min);
    return check_launch();
}

// debug hook (not in the public header): number of rays that took the exact-rescan path since the last call
extern "C" long long papr_debug_select_fallbacks(void)
{
    unsigned long long v = 0, z = 0;
    cudaMemcpyFromSymbol(&v, papr::g_sel_fallbacks, sizeof(v));
    cudaMemcpyToSymbol(papr::g_sel_fallbacks, &z, sizeof(z));
    return (long long)v;
}
