// Stage a1, the acceleration structure of the screen-space grid selection (select_grid_kernel in select.cu): per view a
// camera frame around the mean ray direction, the gnomonic extent of the view's rays and of the points in front of the
// camera, the points binned on a G x G grid over that extent and stored cell by cell (counting sort), each cell's
// smallest |depth|.  Six small launches per call instead of ~60 torch ops (sort, searchsorted, scatter_reduce, ...):
// the structure is rebuilt every training step (the points move) and every frame (the camera moves).
//
// The selection result does not depend on anything computed here: select_grid_kernel only SKIPS cells whose conservative
// distance bound (with its own rounding slack) exceeds the current thresholds, ranks its candidates by the exact key and
// the original point index, and rescans exhaustively when its safety test fails.  So the order of the points inside a
// cell (atomic cursors below: arbitrary) and last-bit differences of the frame are free; what must hold is consistency:
// a point stored in cell (cx, cy) was binned with the SAME gmin / cell size the kernel reads from view_params, and
// sorted_v holds v = RN(p - o) and eps * |v|^2 computed exactly as the plain scan computes them.
#include "common.cuh"

namespace papr {

constexpr int kGbThreads = 256;
constexpr int kGbParts = 64;          // partial direction sums per view (fixed order: the frame is deterministic)
constexpr int kGbExt = 16;            // ints per view: hmin(2) hmax(2) pmin(2) pmax(2) max|v|^2, spare
constexpr int kGbViewFloats = 20;     // = kGridViewFloats of select.cu

// monotone float <-> int map, so float min / max become integer atomics
__device__ __forceinline__ int fkey(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

struct Frame { float e1[3], e2[3], c[3]; };
__device__ __forceinline__ Frame load_frame(const float *vp)
{
    Frame f;
#pragma unroll
    for (int i = 0; i < 3; ++i) { f.e1[i] = vp[i]; f.e2[i] = vp[3 + i]; f.c[i] = vp[6 + i]; }
    return f;
}
__device__ __forceinline__ float dot3(const float *a, float x, float y, float z) { return fmaf(a[2], z, fmaf(a[1], y, a[0] * x)); }

struct PointBin { float vx, vy, vz, vn2, gx, gy, depth; bool valid, front; };
// one point in a view's frame; the same code bins it (count pass) and stores it (scatter pass)
__device__ __forceinline__ PointBin bin_point(const float *p, float ox, float oy, float oz, const Frame &f)
{
    PointBin b;
    b.vx = __fsub_rn(p[0], ox); b.vy = __fsub_rn(p[1], oy); b.vz = __fsub_rn(p[2], oz);
    b.vn2 = fmaf(b.vz, b.vz, fmaf(b.vy, b.vy, b.vx * b.vx));            // as select_topk2_kernel stages it
    const float w1 = dot3(f.e1, b.vx, b.vy, b.vz), w2 = dot3(f.e2, b.vx, b.vy, b.vz), w3 = dot3(f.c, b.vx, b.vy, b.vz);
    const float len = sqrtf(b.vn2);
    b.valid = fabsf(w3) > 1e-3f * len;                                   // depth ~ 0: the direction has no gnomonic image
    const float iw = 1.f / (b.valid ? w3 : 1.f);
    b.gx = w1 * iw; b.gy = w2 * iw;
    b.depth = b.valid ? fabsf(w3) : 0.f;                                 // 0: the point's cell can never be skipped
    b.front = b.valid && w3 > 0.25f * len;
    return b;
}

struct GridParams { float gminx, gminy, cellx, celly, icx, icy; };
// Grid placement from the extents (every thread derives it with the same arithmetic).  The grid covers the rays AND the
// points in front of the camera, at most two ray-spans beyond the rays on each side: when the rays are a stripe of the
// frame that misses the object, a grid over the rays alone would put every point into its semi-infinite border cells.
__device__ __forceinline__ GridParams grid_params(const int *ext, int G)
{
    float hmin[2] = {fkey_inv(ext[0]), fkey_inv(ext[1])}, hmax[2] = {fkey_inv(ext[2]), fkey_inv(ext[3])};
    const float pmin[2] = {fkey_inv(ext[4]), fkey_inv(ext[5])}, pmax[2] = {fkey_inv(ext[6]), fkey_inv(ext[7])};
    if (!(hmin[0] <= hmax[0]) || !(hmin[1] <= hmax[1])) { hmin[0] = hmin[1] = -1.f; hmax[0] = hmax[1] = 1.f; }   // no ray points forward
    const float reach = 2.f * fmaxf(fmaxf(hmax[0] - hmin[0], 1e-3f), fmaxf(hmax[1] - hmin[1], 1e-3f));
    float g0[2], cell[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float lo = fmaxf(fminf(hmin[a], pmin[a]), hmin[a] - reach);
        const float hi = fminf(fmaxf(hmax[a], pmax[a]), hmax[a] + reach);
        const float span = fmaxf(hi - lo, 1e-3f);
        g0[a] = lo - 0.02f * span;
        cell[a] = span * 1.04f / (float)G;
    }
    return GridParams{g0[0], g0[1], cell[0], cell[1], 1.f / cell[0], 1.f / cell[1]};
}

__device__ __forceinline__ int cell_of(const PointBin &b, const GridParams &gp, int G)
{
    if (!b.valid) return 0;
    const float fx = fminf(fmaxf(floorf((b.gx - gp.gminx) * gp.icx), 0.f), (float)(G - 1));
    const float fy = fminf(fmaxf(floorf((b.gy - gp.gminy) * gp.icy), 0.f), (float)(G - 1));
    return (int)fy * G + (int)fx;                                        // (NaN cannot occur for a valid point; fmaxf would map it to 0)
}

// ---- 1. partial sums of the normalised ray directions ---------------------------------------------------------------
__global__ void __launch_bounds__(kGbThreads) grid_dirsum_kernel(const float *__restrict__ rays_d, int64_t R, float *__restrict__ part)
{
    const int view = blockIdx.y, b = blockIdx.x;
    const int64_t chunk = (R + kGbParts - 1) / kGbParts;
    const int64_t r0 = (int64_t)b * chunk, r1 = min(r0 + chunk, R);
    float s[3] = {0.f, 0.f, 0.f};
    for (int64_t r = r0 + threadIdx.x; r < r1; r += kGbThreads) {
        const float *d = rays_d + ((int64_t)view * R + r) * 3;
        const float x = d[0], y = d[1], z = d[2];
        const float inv = 1.f / fmaxf(sqrtf(fmaf(z, z, fmaf(y, y, x * x))), 1e-30f);
        s[0] = fmaf(x, inv, s[0]); s[1] = fmaf(y, inv, s[1]); s[2] = fmaf(z, inv, s[2]);
    }
    __shared__ float red[kGbThreads / 32][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s[0]; red[threadIdx.x >> 5][1] = s[1]; red[threadIdx.x >> 5][2] = s[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int w = 0; w < kGbThreads / 32; ++w) t += red[w][threadIdx.x];
        part[((int64_t)view * kGbParts + b) * 4 + threadIdx.x] = t;
    }
}

// ---- 2. camera frame of each view; extents and cell table reset -----------------------------------------------------
__global__ void __launch_bounds__(kGbThreads) grid_frame_kernel(const float *__restrict__ part, int64_t R, int G, float *__restrict__ views,
                                                                int *__restrict__ ext, int4 *__restrict__ cells)
{
    const int view = blockIdx.x;
    if (threadIdx.x == 0) {
        float c[3] = {0.f, 0.f, 0.f};
        for (int b = 0; b < kGbParts; ++b)
            for (int i = 0; i < 3; ++i) c[i] += part[((int64_t)view * kGbParts + b) * 4 + i];
        const float invR = 1.f / (float)(R > 0 ? R : 1);
        for (int i = 0; i < 3; ++i) c[i] *= invR;
        float n = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        if (!(n > 1e-6f)) { c[0] = 0.f; c[1] = 0.f; c[2] = 1.f; n = 1.f; }        // directions cancel: any frame will do
        for (int i = 0; i < 3; ++i) c[i] /= n;
        const float h[3] = {fabsf(c[0]) < 0.9f ? 1.f : 0.f, fabsf(c[0]) < 0.9f ? 0.f : 1.f, 0.f};
        float e1[3] = {h[1] * c[2] - h[2] * c[1], h[2] * c[0] - h[0] * c[2], h[0] * c[1] - h[1] * c[0]};      // helper x c
        const float n1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        for (int i = 0; i < 3; ++i) e1[i] /= n1;
        const float e2[3] = {c[1] * e1[2] - c[2] * e1[1], c[2] * e1[0] - c[0] * e1[2], c[0] * e1[1] - c[1] * e1[0]};   // c x e1
        float *vp = views + (int64_t)view * kGbViewFloats;
        for (int i = 0; i < 3; ++i) { vp[i] = e1[i]; vp[3 + i] = e2[i]; vp[6 + i] = c[i]; }
        for (int i = 9; i < kGbViewFloats; ++i) vp[i] = 0.f;
        int *e = ext + (int64_t)view * kGbExt;
        const int pinf = fkey(__int_as_float(0x7f800000)), ninf = fkey(__int_as_float(0xff800000));
        e[0] = e[1] = pinf; e[2] = e[3] = ninf; e[4] = e[5] = pinf; e[6] = e[7] = ninf;
        for (int i = 8; i < kGbExt; ++i) e[i] = 0;
    }
    int4 *vc = cells + (int64_t)view * G * G;
    for (int i = threadIdx.x; i < G * G; i += kGbThreads) vc[i] = make_int4(0, 0, 0x7f800000, 0);   // count 0, min depth +inf
}

__device__ __forceinline__ float warp_min(float v) { for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ float warp_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

// ---- 3. gnomonic extent of the rays (blocks [0, ray_blocks)) and of the points in front of the camera (the rest) -----
__global__ void __launch_bounds__(kGbThreads) grid_extent_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                                 const float *__restrict__ points, int64_t R, int P, int ray_blocks,
                                                                 const float *__restrict__ views, int *__restrict__ ext)
{
    const int view = blockIdx.y;
    const Frame f = load_frame(views + (int64_t)view * kGbViewFloats);
    const float INF = __int_as_float(0x7f800000);
    float lo[2] = {INF, INF}, hi[2] = {-INF, -INF}, wmax = 0.f;
    const bool on_rays = (int)blockIdx.x < ray_blocks;
    if (on_rays) {
        for (int64_t r = (int64_t)blockIdx.x * kGbThreads + threadIdx.x; r < R; r += (int64_t)ray_blocks * kGbThreads) {
            const float *d = rays_d + ((int64_t)view * R + r) * 3;
            const float x = d[0], y = d[1], z = d[2];
            const float inv = 1.f / fmaxf(sqrtf(fmaf(z, z, fmaf(y, y, x * x))), 1e-30f);
            const float w3 = dot3(f.c, x, y, z) * inv;
            if (w3 > 0.25f) {                                            // the others cannot be bounded and scan everything
                const float iw = 1.f / w3;
                const float hx = dot3(f.e1, x, y, z) * inv * iw, hy = dot3(f.e2, x, y, z) * inv * iw;
                lo[0] = fminf(lo[0], hx); hi[0] = fmaxf(hi[0], hx); lo[1] = fminf(lo[1], hy); hi[1] = fmaxf(hi[1], hy);
            }
        }
    } else {
        const float ox = rays_o[3 * view], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];
        const int nb = gridDim.x - ray_blocks;
        for (int i = (blockIdx.x - ray_blocks) * kGbThreads + threadIdx.x; i < P; i += nb * kGbThreads) {
            const PointBin b = bin_point(points + (int64_t)i * 3, ox, oy, oz, f);
            wmax = fmaxf(wmax, b.vn2);
            if (b.front) { lo[0] = fminf(lo[0], b.gx); hi[0] = fmaxf(hi[0], b.gx); lo[1] = fminf(lo[1], b.gy); hi[1] = fmaxf(hi[1], b.gy); }
        }
    }
    __shared__ float red[kGbThreads / 32][5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float a0 = warp_min(lo[0]), a1 = warp_min(lo[1]), b0 = warp_max(hi[0]), b1 = warp_max(hi[1]), wm = warp_max(wmax);
    if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = b0; red[warp][3] = b1; red[warp][4] = wm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r0 = INF, r1 = INF, r2 = -INF, r3 = -INF, r4 = 0.f;
        for (int w = 0; w < kGbThreads / 32; ++w) {
            r0 = fminf(r0, red[w][0]); r1 = fminf(r1, red[w][1]); r2 = fmaxf(r2, red[w][2]); r3 = fmaxf(r3, red[w][3]); r4 = fmaxf(r4, red[w][4]);
        }
        int *e = ext + (int64_t)view * kGbExt + (on_rays ? 0 : 4);
        atomicMin(e + 0, fkey(r0)); atomicMin(e + 1, fkey(r1)); atomicMax(e + 2, fkey(r2)); atomicMax(e + 3, fkey(r3));
        if (!on_rays) atomicMax(ext + (int64_t)view * kGbExt + 8, __float_as_int(r4));     // |v|^2 >= 0: its bits order as ints
    }
}

// ---- 4. cell of every point, points per cell, smallest depth per cell ------------------------------------------------
__global__ void __launch_bounds__(kGbThreads) grid_count_kernel(const float *__restrict__ rays_o, const float *__restrict__ points, int P, int G,
                                                                float *__restrict__ views, const int *__restrict__ ext, int4 *__restrict__ cells,
                                                                int *__restrict__ cid_ws)
{
    const int view = blockIdx.y;
    float *vp = views + (int64_t)view * kGbViewFloats;
    const Frame f = load_frame(vp);
    const GridParams gp = grid_params(ext + (int64_t)view * kGbExt, G);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        vp[9] = gp.gminx; vp[10] = gp.gminy; vp[11] = gp.cellx; vp[12] = gp.celly; vp[13] = gp.icx; vp[14] = gp.icy;
    }
    const float ox = rays_o[3 * view], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];
    int4 *vc = cells + (int64_t)view * G * G;
    for (int i = blockIdx.x * kGbThreads + threadIdx.x; i < P; i += gridDim.x * kGbThreads) {
        const PointBin b = bin_point(points + (int64_t)i * 3, ox, oy, oz, f);
        const int c = cell_of(b, gp, G);
        cid_ws[(int64_t)view * P + i] = c;
        atomicAdd(&vc[c].x, 1);
        atomicMin(&vc[c].z, __float_as_int(b.depth));                    // depth >= 0
    }
}

// ---- 5. exclusive scan of the counts -> cell ranges; min depth of the view -------------------------------------------
__global__ void __launch_bounds__(1024) grid_scan_kernel(int G, float *__restrict__ views, const int *__restrict__ ext, int4 *__restrict__ cells,
                                                         int *__restrict__ cursor)
{
    const int view = blockIdx.x, n = G * G;
    int4 *vc = cells + (int64_t)view * n;
    int *cur = cursor + (int64_t)view * n;
    __shared__ int wsum[32];
    __shared__ int carry_s;
    __shared__ float zred[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    float zmin = __int_as_float(0x7f800000);
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        int cnt = 0;
        if (i < n) { const int4 m = vc[i]; cnt = m.x; zmin = fminf(zmin, __int_as_float(m.z)); }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            wsum[lane] = w;                                              // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int start = carry + (warp ? wsum[warp - 1] : 0) + incl - cnt;
        if (i < n) { vc[i].x = start; vc[i].y = start + cnt; cur[i] = start; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    zmin = warp_min(zmin);
    if (lane == 0) zred[warp] = zmin;
    __syncthreads();
    if (threadIdx.x == 0) {
        float z = zred[0];
        for (int w = 1; w < 32; ++w) z = fminf(z, zred[w]);
        float *vp = views + (int64_t)view * kGbViewFloats;
        vp[15] = z * (1.f - 1e-6f);
        vp[16] = __int_as_float(ext[(int64_t)view * kGbExt + 8]);
    }
}

// ---- 6. counting-sort scatter: v = p - o and eps |v|^2 in cell order, original indices ------------------------------
__global__ void __launch_bounds__(kGbThreads) grid_scatter_kernel(const float *__restrict__ rays_o, const float *__restrict__ points, int P, float eps,
                                                                  const int *__restrict__ cid_ws, int *__restrict__ cursor, int cells_per_view,
                                                                  float4 *__restrict__ sv, int32_t *__restrict__ perm)
{
    const int view = blockIdx.y;
    const float ox = rays_o[3 * view], oy = rays_o[3 * view + 1], oz = rays_o[3 * view + 2];
    int *cur = cursor + (int64_t)view * cells_per_view;
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * kGbThreads; i0 < P; i0 += gridDim.x * kGbThreads) {
        const int i = i0 + threadIdx.x;
        const bool on = i < P;
        const int c = on ? cid_ws[(int64_t)view * P + i] : -1 - lane;     // idle lanes: distinct keys, no partners
        // lanes of a warp that hit the same cell take consecutive slots from ONE atomic (a border cell of a stripe render
        // can hold thousands of points)
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const int leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (on && lane == leader) base = atomicAdd(cur + c, __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (on) {
            const float *p = points + (int64_t)i * 3;
            const float vx = __fsub_rn(p[0], ox), vy = __fsub_rn(p[1], oy), vz = __fsub_rn(p[2], oz);
            const float w = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
            const int64_t pos = (int64_t)view * P + base + rank;
            sv[pos] = make_float4(vx, vy, vz, eps * w);
            perm[pos] = i;
        }
    }
}

}  // namespace papr

extern "C" int64_t papr_select_grid_workspace_bytes(int64_t n_views, int64_t P, int G)
{
    using namespace papr;
    if (n_views < 0 || P < 0 || G < 1) return 0;
    return n_views * ((int64_t)kGbParts * 4 * 4 + (int64_t)kGbExt * 4 + (int64_t)G * G * 4 + P * 4);
}

extern "C" int papr_select_grid_build(const float *rays_o, const float *rays_d, const float *points, int64_t n_views, int64_t rays_per_view,
                                      int64_t P, int G, float eps, void *sorted_v, int32_t *perm, int32_t *cells, float *view_params,
                                      void *workspace, int64_t workspace_bytes, void *stream)
{
    using namespace papr;
    if (!rays_o || !rays_d || !points || !sorted_v || !perm || !cells || !view_params || !workspace) return PAPR_ERR_INVALID_ARGUMENT;
    if (P < 1 || P > INT32_MAX || G < 1 || G > 1024 || n_views < 0 || n_views > 65535 || rays_per_view < 0) return PAPR_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < papr_select_grid_workspace_bytes(n_views, P, G)) return PAPR_ERR_INVALID_ARGUMENT;
    if (n_views == 0) return PAPR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float *part = (float *)workspace;
    int *ext = (int *)(part + n_views * kGbParts * 4);
    int *cursor = ext + n_views * kGbExt;
    int *cid_ws = cursor + n_views * (int64_t)G * G;
    const int cap = 2 * kNumSMs;
    const int ray_blocks = (int)max((int64_t)1, min((rays_per_view + 4 * kGbThreads - 1) / (4 * kGbThreads), (int64_t)cap));
    const int point_blocks = (int)max((int64_t)1, min((P + kGbThreads - 1) / kGbThreads, (int64_t)cap));
    grid_dirsum_kernel<<<dim3(kGbParts, (unsigned)n_views), kGbThreads, 0, st>>>(rays_d, rays_per_view, part);
    grid_frame_kernel<<<(unsigned)n_views, kGbThreads, 0, st>>>(part, rays_per_view, G, view_params, ext, (int4 *)cells);
    grid_extent_kernel<<<dim3(ray_blocks + point_blocks, (unsigned)n_views), kGbThreads, 0, st>>>(rays_o, rays_d, points, rays_per_view, (int)P,
                                                                                                    ray_blocks, view_params, ext);
    grid_count_kernel<<<dim3(point_blocks, (unsigned)n_views), kGbThreads, 0, st>>>(rays_o, points, (int)P, G, view_params, ext, (int4 *)cells, cid_ws);
    grid_scan_kernel<<<(unsigned)n_views, 1024, 0, st>>>(G, view_params, ext, (int4 *)cells, cursor);
    grid_scatter_kernel<<<dim3(point_blocks, (unsigned)n_views), kGbThreads, 0, st>>>(rays_o, points, (int)P, eps, cid_ws, cursor, G * G,
                                                                                       (float4 *)sorted_v, perm);
    return check_launch();
}
