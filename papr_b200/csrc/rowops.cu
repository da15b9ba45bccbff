// Per-ray CUDA-core stages of the proximity-attention path (everything that is not a dense contraction):
//   attn_prologue_fwd / _bwd : gather + ray/point geometry + positional encoding + key input LayerNorm
//                              (reference models/model.py:285-310,396-437; models/utils.py:232-257; attn.py:30-42,172-191)
//   score_blend_fwd          : key output LayerNorm folded with w_k and the query (scaled dot), ReLU, influence scores,
//                              background token, softmax, top-K renormalisation, value aggregation
//                              (attn.py:30-42,217-226; model.py:519-534)
//   blend_bwd, key_score_bwd : their backward passes (autograd of the same lines)
// One warp owns one ray (its K candidate rows); lanes sweep columns, so tile-blocked rows are read and written as
// whole 16-byte chunks and LayerNorm reductions are warp shuffles.  All math is fp32; bf16 only at the tensor-core
// operand boundary.  Row index = ray * K + k everywhere.
#include "tc_common.cuh"

namespace papr {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kMaxDk = 128;    // 9*(1+2L) <= 117 for L <= 6
constexpr int kMaxDv = 256;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct Geometry { float g[9]; };   // selected point (3), along-ray vector proj (3), perpendicular vector D (3)

// model.py:302-305 for one (ray, point): rays = d/(|d|+eps); t = (v.rays)/(rays.rays+eps); proj = rays*t; D = v-proj
__device__ __forceinline__ Geometry ray_point_geometry(const float *p, const float *o, const float *d, float eps,
                                                       float *u_out, float *den_out)
{
    const float nrm = sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]))) + eps;
    const float ux = d[0] / nrm, uy = d[1] / nrm, uz = d[2] / nrm;
    const float vx = p[0] - o[0], vy = p[1] - o[1], vz = p[2] - o[2];
    const float den = (ux * ux + uy * uy) + uz * uz + eps;
    const float t = ((vx * ux + vy * uy) + vz * uz) / den;
    Geometry r;
    r.g[0] = p[0]; r.g[1] = p[1]; r.g[2] = p[2];
    r.g[3] = ux * t; r.g[4] = uy * t; r.g[5] = uz * t;
    r.g[6] = vx - r.g[3]; r.g[7] = vy - r.g[4]; r.g[8] = vz - r.g[5];
    u_out[0] = ux; u_out[1] = uy; u_out[2] = uz; *den_out = den;
    return r;
}

// Column j of the key embedding input -> which geometry scalar, which PE slot (utils.py:232-242: per coordinate
// [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)])
struct PeCol { int src; int slot; };
__device__ __forceinline__ PeCol pe_col(int j, int S) { PeCol c; c.src = j / S; c.slot = j - c.src * S; return c; }

__device__ __forceinline__ float pe_value(float x, int slot)
{
    if (slot == 0) return x;
    const float a = ldexpf(x, (slot - 1) >> 1);
    float s, c;
    sincosf(a, &s, &c);
    return (slot & 1) ? s : c;
}
// d pe_value / dx
__device__ __forceinline__ float pe_deriv(float x, int slot)
{
    if (slot == 0) return 1.f;
    const int oct = (slot - 1) >> 1;
    const float a = ldexpf(x, oct);
    float s, c;
    sincosf(a, &s, &c);
    return ldexpf((slot & 1) ? c : -s, oct);
}

struct PrologueParams {
    const float *rays_o, *rays_d, *points, *feats, *a2, *b2;
    const int32_t *idx;
    int64_t R, rays_per_view;
    int K, L, F, dk, dv, nblk_k, nblk_v;
    float eps;
    // forward outputs
    uint8_t *kin, *vin;
    float *kin_f32, *vin_f32;         // optional fp32 taps (row-major [M,dk], [M,dv])
    // backward inputs / outputs
    const uint8_t *dkin, *dvin;
    const float *dkin_f32, *dvin_f32; // optional fp32 taps
    float *g_points, *g_feats, *g_a2, *g_b2;
};

__global__ void __launch_bounds__(kRowThreads) attn_prologue_fwd_kernel(const PrologueParams p)
{
    __shared__ float pe_s[kRowWarps][kMaxDk];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = 1 + 2 * p.L;
    const int64_t M = p.R * p.K;
    float *pe = pe_s[warp];

    PeCol cols[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) cols[m] = pe_col(min(lane + 32 * m, p.dk - 1), S);

    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        const int64_t view = ray / p.rays_per_view;
        Geometry geo;
        int pidx = 0;
        {
            float u[3], den;
            const int k = min(lane, p.K - 1);
            pidx = p.idx[ray * p.K + k];
            geo = ray_point_geometry(p.points + (size_t)pidx * 3, p.rays_o + view * 3, p.rays_d + ray * 3, p.eps, u, &den);
        }
        for (int k = 0; k < p.K; ++k) {
            const int64_t row = ray * p.K + k;
            float g[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) g[i] = __shfl_sync(0xffffffffu, geo.g[i], k);
            const int pk = __shfl_sync(0xffffffffu, pidx, k);
            float val[4], sum = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int j = lane + 32 * m;
                float x = g[0];
#pragma unroll
                for (int i = 1; i < 9; ++i) x = (cols[m].src == i) ? g[i] : x;
                val[m] = (j < p.dk) ? pe_value(x, cols[m].slot) : 0.f;
                sum += val[m];
            }
            const float mean = warp_sum(sum) / (float)p.dk;
            float sq = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int j = lane + 32 * m;
                if (j < p.dk) { const float c = val[m] - mean; sq += c * c; pe[j] = val[m]; }
            }
            const float stdv = sqrtf(warp_sum(sq) / (float)(p.dk - 1));
            const float rstd = 1.f / (stdv + p.eps);
            __syncwarp();
            // key input: LayerNorm(pe) (attn.py:39-42), zero padded to nblk_k*64 columns
            for (int c = lane; c < p.nblk_k * 8; c += 32) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = c * 8 + e;
                    f[e] = (j < p.dk) ? __ldg(p.a2 + j) * (pe[j] - mean) * rstd + __ldg(p.b2 + j) : 0.f;
                    if (p.kin_f32 && j < p.dk) p.kin_f32[row * p.dk + j] = f[e];
                }
                *reinterpret_cast<uint4 *>(p.kin + blocked_chunk_offset(row, c, p.nblk_k)) =
                    make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
            // value input: [PE(proj), PE(D), point features] (attn.py:175,187,191)
            const int dpe = 6 * S;
            for (int c = lane; c < p.nblk_v * 8; c += 32) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = c * 8 + e;
                    float t = 0.f;
                    if (j < dpe) t = pe[j + 3 * S];
                    else if (j < p.dv) t = __ldg(p.feats + (size_t)pk * p.F + (j - dpe));
                    f[e] = t;
                    if (p.vin_f32 && j < p.dv) p.vin_f32[row * p.dv + j] = t;
                }
                *reinterpret_cast<uint4 *>(p.vin + blocked_chunk_offset(row, c, p.nblk_v)) =
                    make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
            __syncwarp();
        }
    }
    // zero the padding rows [M, pad128(M)) so downstream GEMMs and column sums see zeros
    const int64_t M_pad = (M + 127) / 128 * 128;
    const int per_row = (p.nblk_k + p.nblk_v) * 8;
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < (M_pad - M) * per_row; i += (int64_t)gridDim.x * kRowThreads) {
        const int64_t row = M + i / per_row;
        const int c = (int)(i % per_row);
        if (c < p.nblk_k * 8) *reinterpret_cast<uint4 *>(p.kin + blocked_chunk_offset(row, c, p.nblk_k)) = make_uint4(0, 0, 0, 0);
        else *reinterpret_cast<uint4 *>(p.vin + blocked_chunk_offset(row, c - p.nblk_k * 8, p.nblk_v)) = make_uint4(0, 0, 0, 0);
    }
}

__device__ __forceinline__ void unpack8(const uint4 q, float *f)
{
    f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
    f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}

__global__ void __launch_bounds__(kRowThreads) attn_prologue_bwd_kernel(const PrologueParams p)
{
    __shared__ float pe_s[kRowWarps][kMaxDk];      // pe values, then per-column contributions to d(geometry)
    __shared__ float gk_s[kRowWarps][kMaxDk];      // d kin
    __shared__ float gv_s[kRowWarps][kMaxDv];      // d vin
    __shared__ float dx_s[kRowWarps][32][6];       // per candidate: d proj (3), d D (3)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = 1 + 2 * p.L;
    float *pe = pe_s[warp], *gk = gk_s[warp], *gv = gv_s[warp];

    PeCol cols[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) cols[m] = pe_col(min(lane + 32 * m, p.dk - 1), S);
    float acc_a2[4] = {0, 0, 0, 0}, acc_b2[4] = {0, 0, 0, 0};

    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        const int64_t view = ray / p.rays_per_view;
        Geometry geo;
        float u[3], den;
        int pidx;
        {
            const int k = min(lane, p.K - 1);
            pidx = p.idx[ray * p.K + k];
            geo = ray_point_geometry(p.points + (size_t)pidx * 3, p.rays_o + view * 3, p.rays_d + ray * 3, p.eps, u, &den);
        }
        for (int k = 0; k < p.K; ++k) {
            const int64_t row = ray * p.K + k;
            float g[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) g[i] = __shfl_sync(0xffffffffu, geo.g[i], k);
            const int pk = __shfl_sync(0xffffffffu, pidx, k);
            // stage the incoming gradients column-wise
            for (int c = lane; c < p.nblk_k * 8; c += 32) {
                float f[8];
                if (p.dkin_f32) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = (c * 8 + e < p.dk) ? p.dkin_f32[row * p.dk + c * 8 + e] : 0.f;
                } else unpack8(*reinterpret_cast<const uint4 *>(p.dkin + blocked_chunk_offset(row, c, p.nblk_k)), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (c * 8 + e < kMaxDk) gk[c * 8 + e] = f[e];
            }
            for (int c = lane; c < p.nblk_v * 8; c += 32) {
                float f[8];
                if (p.dvin_f32) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = (c * 8 + e < p.dv) ? p.dvin_f32[row * p.dv + c * 8 + e] : 0.f;
                } else unpack8(*reinterpret_cast<const uint4 *>(p.dvin + blocked_chunk_offset(row, c, p.nblk_v)), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) gv[c * 8 + e] = f[e];
            }
            // recompute pe + LayerNorm statistics
            float val[4], x4[4], sum = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int j = lane + 32 * m;
                float x = g[0];
#pragma unroll
                for (int i = 1; i < 9; ++i) x = (cols[m].src == i) ? g[i] : x;
                x4[m] = x;
                val[m] = (j < p.dk) ? pe_value(x, cols[m].slot) : 0.f;
                sum += val[m];
            }
            const float mean = warp_sum(sum) / (float)p.dk;
            float sq = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) { const float c = val[m] - mean; if (lane + 32 * m < p.dk) sq += c * c; }
            const float stdv = sqrtf(warp_sum(sq) / (float)(p.dk - 1));
            const float rstd = 1.f / (stdv + p.eps);
            __syncwarp();
            // LayerNorm backward: z = (pe-mean)*rstd, kin = a2*z+b2
            float z[4], gz[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int j = lane + 32 * m;
                z[m] = 0.f; gz[m] = 0.f;
                if (j < p.dk) {
                    z[m] = (val[m] - mean) * rstd;
                    const float go = gk[j];
                    acc_a2[m] += go * z[m]; acc_b2[m] += go;
                    gz[m] = go * __ldg(p.a2 + j);
                    s1 += gz[m]; s2 += gz[m] * z[m];
                }
            }
            s1 = warp_sum(s1) / (float)p.dk;
            s2 = warp_sum(s2) / ((float)(p.dk - 1) * stdv);
            // per-column contribution to d(geometry scalar): (dLN + d vin) * dPE/dx; the raw-position key features
            // are detached in the reference (model.py:405), so columns of source < 3 contribute nothing
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int j = lane + 32 * m;
                if (j < p.dk) {
                    float d = rstd * (gz[m] - s1) - z[m] * s2;
                    float contrib = 0.f;
                    if (cols[m].src >= 3) {
                        d += gv[j - 3 * S];
                        contrib = d * pe_deriv(x4[m], cols[m].slot);
                    }
                    pe[j] = contrib;
                }
            }
            __syncwarp();
            if (lane < 6) {
                float t = 0.f;
                for (int s = 0; s < S; ++s) t += pe[(3 + lane) * S + s];
                dx_s[warp][k][lane] = t;
            }
            // point-feature gradient (model.py:434-435 gather backward)
            for (int c = lane; c < p.F; c += 32) atomicAdd(p.g_feats + (size_t)pk * p.F + c, gv[6 * S + c]);
            __syncwarp();
        }
        if (lane < p.K) {
            const float *dx = dx_s[warp][lane];
            const float e0 = dx[0] - dx[3], e1 = dx[1] - dx[4], e2 = dx[2] - dx[5];
            const float s = (e0 * u[0] + e1 * u[1] + e2 * u[2]) / den;
            atomicAdd(p.g_points + (size_t)pidx * 3 + 0, dx[3] + u[0] * s);
            atomicAdd(p.g_points + (size_t)pidx * 3 + 1, dx[4] + u[1] * s);
            atomicAdd(p.g_points + (size_t)pidx * 3 + 2, dx[5] + u[2] * s);
        }
        __syncwarp();
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int j = lane + 32 * m;
        if (j < p.dk) { atomicAdd(p.g_a2 + j, acc_a2[m]); atomicAdd(p.g_b2 + j, acc_b2[m]); }
    }
}


// ------------------------------------------------------------------------------------------------ fast prologue
// Product-path variants (no fp32 taps): positional encodings come from ONE accurate sincosf per geometry scalar plus
// the double-angle recurrence (abs. error <= 2^5 * 1 ulp ~ 4e-6, far below bf16 resolution), computed by 9 lanes and
// shared through shared memory; LayerNorm affine terms live in registers; feature gradients use 16-byte vector
// reductions (red.global.add.v4.f32).
// lanes 0..8: fill pe[src*S + slot] for geometry scalar `src` = lane
__device__ __forceinline__ void pe_fill(float x, int L, float *dst)
{
    float s, c;
    sincosf(x, &s, &c);
    dst[0] = x;
    for (int i = 0; i < L; ++i) {
        dst[1 + 2 * i] = s; dst[2 + 2 * i] = c;
        const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
        s = s2; c = c2;
    }
}

// 16-byte asynchronous global -> shared copies (LDGSTS): a lane can have its whole share of a staged tile in flight at
// once without holding it in registers (a register-staged copy of 40 x 16 bytes per lane is issued in batches of a few
// loads, one DRAM latency after the other: ncu r02, 22% of the prologue backward's stall samples sat on those stores)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// One (ray, candidate) row per LANE (the default shape L = 6, F = 64: 117-wide key input, 142-wide value input).  The
// warp-per-ray kernel at the top of this file spends ~400 warp instructions per row (every column costs shuffles /
// shared-memory hops); here a lane carries its row in registers (9 sincos,
// double-angle recurrences, one statistics pass, one emitting pass with compile-time column indices): ~55 warp
// instructions per row.  A warp stages its 32 rows -- 32 x 128 B per 64-column block, already in the tile-blocked
// swizzle -- in shared memory and copies each 4 KB piece out with fully coalesced 16-byte stores.
template <int L, int F>
__global__ void __launch_bounds__(kRowThreads, 2) attn_prologue_fwd_rows_kernel(const PrologueParams p)
{
    constexpr int S = 1 + 2 * L, DK = 9 * S, DPE = 6 * S, DV = DPE + F;
    constexpr int NBK = (DK + 63) / 64, NBV = (DV + 63) / 64;
    extern __shared__ __align__(16) uint8_t rows_stage[];
    __shared__ float2 ab_s[DK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < DK; j += kRowThreads) ab_s[j] = make_float2(p.a2[j], p.b2[j]);
    __syncthreads();
    constexpr int NBS = NBK > NBV ? NBK : NBV;      // the key and the value rows are staged one after the other
    const uint32_t wst = smem_u32(rows_stage) + (uint32_t)warp * NBS * 4096u;
    const uint32_t my = wst + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)lane & 7u;
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;

    // the candidate index of the NEXT group of rows is requested a whole iteration ahead: idx -> point is otherwise a chain of
    // two dependent global loads at the top of every iteration
    const int64_t row_step = (int64_t)gridDim.x * kRowWarps * 32;
    int pidx_next = 0;
    {
        const int64_t first = ((int64_t)blockIdx.x * kRowWarps + warp) * 32 + lane;
        if (first < M) pidx_next = p.idx[first];
    }
    for (int64_t row0 = ((int64_t)blockIdx.x * kRowWarps + warp) * 32; row0 < M_pad; row0 += row_step) {
        const int64_t row = row0 + lane;
        const bool live = row < M;
        float g[9];
        const int pidx = pidx_next;
        pidx_next = 0;
        if (row + row_step < M) pidx_next = p.idx[row + row_step];
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = 0.f;
        if (live) {
            const int64_t ray = row / p.K;
            const int64_t view = ray / p.rays_per_view;
            float u[3], den;
            const Geometry geo = ray_point_geometry(p.points + (size_t)pidx * 3, p.rays_o + view * 3, p.rays_d + ray * 3, p.eps, u, &den);
#pragma unroll
            for (int i = 0; i < 9; ++i) g[i] = geo.g[i];
        }
        float s0[9], c0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) sincosf(g[i], &s0[i], &c0[i]);
        // statistics of the 117 key columns (model: mean, unbiased std)
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float s = s0[i], c = c0[i];
            sum += g[i]; sq = fmaf(g[i], g[i], sq);
#pragma unroll
            for (int o = 0; o < L; ++o) {
                sum += s + c; sq = fmaf(s, s, fmaf(c, c, sq));
                const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
                s = s2; c = c2;
            }
        }
        const float mean = sum * (1.f / DK);
        const float var = fmaxf(sq - (float)DK * mean * mean, 0.f) * (1.f / (DK - 1));
        const float rstd = live ? 1.f / (sqrtf(var) + p.eps) : 0.f;
        const float lv = live ? 1.f : 0.f;

        // emit: columns in order; eight of them make one 16-byte chunk of the row
        float kb[8], vb[8];
        auto put_k = [&](int c) {        // key chunk c <- kb
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + (uint32_t)(c >> 3) * 4096u + ((((uint32_t)c & 7u) ^ sw) << 4)),
                         "r"(pack_bf16(kb[0], kb[1])), "r"(pack_bf16(kb[2], kb[3])), "r"(pack_bf16(kb[4], kb[5])), "r"(pack_bf16(kb[6], kb[7])) : "memory");
        };
        auto put_v = [&](int c) {        // value chunk c <- vb
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + (uint32_t)(c >> 3) * 4096u + ((((uint32_t)c & 7u) ^ sw) << 4)),
                         "r"(pack_bf16(vb[0], vb[1])), "r"(pack_bf16(vb[2], vb[3])), "r"(pack_bf16(vb[4], vb[5])), "r"(pack_bf16(vb[6], vb[7])) : "memory");
        };
#pragma unroll
        for (int src = 0; src < 9; ++src) {
            float s = s0[src], c = c0[src];
#pragma unroll
            for (int slot = 0; slot < S; ++slot) {
                float val;
                if (slot == 0) val = g[src];
                else if (slot & 1) val = s;
                else {
                    val = c;
                    const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
                    s = s2; c = c2;
                }
                const int j = src * S + slot;
                const float2 ab = ab_s[j];
                kb[j & 7] = lv * fmaf((val - mean) * rstd, ab.x, ab.y);
                if ((j & 7) == 7) put_k(j >> 3);
            }
        }
        if (DK & 7) {
#pragma unroll
            for (int e = DK & 7; e < 8; ++e) kb[e] = 0.f;
            put_k(DK >> 3);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) kb[e] = 0.f;
#pragma unroll
        for (int c = (DK + 7) >> 3; c < NBK * 8; ++c) put_k(c);
        const int64_t tile = row0 >> 7;
        const uint32_t roff = (uint32_t)(row0 & 127) * 128u + (uint32_t)lane * 16u;
        auto copy_out = [&](uint8_t *base, int nb) {      // the warp's 32 rows are 4 KB contiguous inside every 16 KB block of the tile
            __syncwarp();
            for (int b = 0; b < nb; ++b) {
                uint8_t *dst = base + ((size_t)tile * nb + b) * kBlockBytes + roff;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    uint4 t;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(wst + (uint32_t)b * 4096u + (uint32_t)it * 512u + (uint32_t)lane * 16u));
                    *reinterpret_cast<uint4 *>(dst + it * 512) = t;
                }
            }
            __syncwarp();
        };
        copy_out(p.kin, NBK);
        // value side: the encodings of geometry scalars 3..8 again (unnormalised), then the point features
#pragma unroll
        for (int src = 3; src < 9; ++src) {
            float s = s0[src], c = c0[src];
#pragma unroll
            for (int slot = 0; slot < S; ++slot) {
                float val;
                if (slot == 0) val = g[src];
                else if (slot & 1) val = s;
                else {
                    val = c;
                    const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
                    s = s2; c = c2;
                }
                const int jv = (src - 3) * S + slot;
                vb[jv & 7] = lv * val;
                if ((jv & 7) == 7) put_v(jv >> 3);
            }
        }
        // point features follow the value-side encoding (model.py:396-437: cat(pe(geometry), pc_feats)).
        // A lane that fetched its own row's F floats would wait for one L2 round trip per 16-byte load (they are consumed one
        // chunk at a time and there are no registers to hold them all); instead the WARP fetches every row together: for row r
        // lane l loads features 2l, 2l+1 (+64), all 32 rows requested back to back, then each pair goes to its place in
        // the staged row as one 32-bit store (DPE is even, so a pair never straddles a 16-byte chunk).
        static_assert((DPE & 1) == 0 && (F == 64 || F == 128), "feature pairs must be 4-byte aligned in the row");
        if (DPE & 7) {                                       // the chunk shared by the last PE columns and the first features
#pragma unroll
            for (int e = DPE & 7; e < 8; ++e) vb[e] = 0.f;
            put_v(DPE >> 3);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) vb[e] = 0.f;
#pragma unroll
        for (int c = (DPE + 7) >> 3; c < NBV * 8; ++c) put_v(c);
        __syncwarp();
        {
            const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
#pragma unroll
            for (int h = 0; h < F / 64; ++h) {
                float2 q[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const int pr = __shfl_sync(0xffffffffu, pidx, r);
                    q[r] = make_float2(0.f, 0.f);
                    if ((live_mask >> r) & 1u) q[r] = __ldg(reinterpret_cast<const float2 *>(p.feats + (size_t)pr * F + h * 64) + lane);
                }
                const int jv = DPE + h * 64 + 2 * lane;      // column of the pair in the value row
                const uint32_t blk_off = (uint32_t)(jv >> 6) * 4096u, chunk = ((uint32_t)jv >> 3) & 7u, within = ((uint32_t)jv & 7u) * 2u;
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(wst + blk_off + (uint32_t)r * 128u + ((chunk ^ ((uint32_t)r & 7u)) << 4) + within),
                                 "r"(pack_bf16(q[r].x, q[r].y)) : "memory");
            }
        }
        copy_out(p.vin, NBV);
    }
}

// Backward of the prologue, one (ray, candidate) row per lane (default shape L = 6, F = 64), the counterpart of
// attn_prologue_fwd_rows_kernel.  A warp brings its 32 rows of d kin / d vin into shared memory with coalesced loads; a lane
// then walks its row's columns once for the LayerNorm statistics and once for everything else.  The chain through the
// positional encoding is linear in the column gradients, so the six geometry gradients are accumulated as four running
// sums per source (A = sum coef*gz, B = sum coef, C = sum coef*z, D = sum coef*gv) next to s1 = sum gz, s2 = sum gz*z and
// combined at the end: t = rstd*A - rstd*s1*B - s2*C + D.  The LayerNorm-affine gradients need sums over rows, not
// columns: d kin (for g_b2) and d kin * z (written back in place as bf16, for g_a2) are summed down the staged tile by
// half-warps, 16 rows each, and kept in registers until the kernel ends.
template <int L, int F>
__global__ void __launch_bounds__(128, 2) attn_prologue_bwd_rows_kernel(const PrologueParams p)
{
    constexpr int S = 1 + 2 * L, DK = 9 * S, DPE = 6 * S, DV = DPE + F;
    constexpr int NBK = (DK + 63) / 64, NBV = (DV + 63) / 64;
    constexpr int kWarps = 4;
    extern __shared__ __align__(16) uint8_t rows_stage[];
    __shared__ float a2_s[DK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < DK; j += 128) a2_s[j] = p.a2[j];
    __syncthreads();
    const uint32_t wst = smem_u32(rows_stage) + (uint32_t)warp * (NBK + NBV) * 4096u;
    const uint32_t my = wst + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)lane & 7u;
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;
    // column-sum duty: chunk cq (8 key columns) over rows [16*half, 16*half + 16) of the warp's 32
    const int cq = lane & 15, half = lane >> 4;
    float acc_a2[8], acc_b2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc_a2[e] = 0.f; acc_b2[e] = 0.f; }
    auto column_sums = [&](float *acc, int64_t row0) {
#pragma unroll 8
        for (int r = half * 16; r < half * 16 + 16; ++r) {
            if (row0 + r < M) {
                uint4 q;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                             : "r"(wst + (uint32_t)(cq >> 3) * 4096u + (uint32_t)r * 128u + ((((uint32_t)cq & 7u) ^ ((uint32_t)r & 7u)) << 4)));
                float f[8];
                unpack8(q, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += f[e];
            }
        }
    };

    // The row's candidate index, and through it the point, are two dependent global loads ahead of everything else (ncu r02
    // final: 14% of the stall samples on the first uses of the point and of the ray direction): the index is requested two
    // iterations ahead, the point and the ray direction one iteration ahead.
    const int64_t row_step = (int64_t)gridDim.x * kWarps * 32;
    const int64_t row_first = ((int64_t)blockIdx.x * kWarps + warp) * 32 + lane;
    int pidx_cur = 0, pidx_nxt = 0;
    float pt_cur[3] = {0.f, 0.f, 0.f}, rd_cur[3] = {0.f, 0.f, 1.f};
    if (row_first < M) {
        pidx_cur = p.idx[row_first];
#pragma unroll
        for (int i = 0; i < 3; ++i) { pt_cur[i] = p.points[(size_t)pidx_cur * 3 + i]; rd_cur[i] = p.rays_d[(row_first / p.K) * 3 + i]; }
    }
    if (row_first + row_step < M) pidx_nxt = p.idx[row_first + row_step];
    for (int64_t row0 = ((int64_t)blockIdx.x * kWarps + warp) * 32; row0 < M_pad; row0 += row_step) {
        const int64_t row = row0 + lane;
        const bool live = row < M;
        // stage the 32 rows: 4 KB contiguous per 64-column block
        {
            const int64_t tile = row0 >> 7;
            const uint32_t roff = (uint32_t)(row0 & 127) * 128u + (uint32_t)lane * 16u;
#pragma unroll
            for (int b = 0; b < NBK + NBV; ++b) {
                const uint8_t *src = (b < NBK ? p.dkin + ((size_t)tile * NBK + b) * kBlockBytes : p.dvin + ((size_t)tile * NBV + (b - NBK)) * kBlockBytes) + roff;
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    cp_async16(wst + (uint32_t)b * 4096u + (uint32_t)it * 512u + (uint32_t)lane * 16u, src + it * 512);
            }
        }
        // (the tile lands while the lane works out its row's geometry and encoding, which need none of it)
        float g[9], u[3] = {0.f, 0.f, 0.f}, den = 1.f;
        const int pidx = pidx_cur;
        const float pt[3] = {pt_cur[0], pt_cur[1], pt_cur[2]}, rd[3] = {rd_cur[0], rd_cur[1], rd_cur[2]};
        {                                                       // requests for the next two iterations
            const int64_t rn = row + row_step;
            pidx_cur = pidx_nxt;
            if (rn < M) {
#pragma unroll
                for (int i = 0; i < 3; ++i) { pt_cur[i] = p.points[(size_t)pidx_cur * 3 + i]; rd_cur[i] = p.rays_d[(rn / p.K) * 3 + i]; }
            }
            pidx_nxt = rn + row_step < M ? p.idx[rn + row_step] : 0;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = 0.f;
        if (live) {
            const int64_t ray = row / p.K;
            const int64_t view = ray / p.rays_per_view;
            const Geometry geo = ray_point_geometry(pt, p.rays_o + view * 3, rd, p.eps, u, &den);
#pragma unroll
            for (int i = 0; i < 9; ++i) g[i] = geo.g[i];
        }
        float s0[9], c0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) sincosf(g[i], &s0[i], &c0[i]);
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float s = s0[i], c = c0[i];
            sum += g[i]; sq = fmaf(g[i], g[i], sq);
#pragma unroll
            for (int o = 0; o < L; ++o) {
                sum += s + c; sq = fmaf(s, s, fmaf(c, c, sq));
                const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
                s = s2; c = c2;
            }
        }
        const float mean = sum * (1.f / DK);
        const float var = fmaxf(sq - (float)DK * mean * mean, 0.f) * (1.f / (DK - 1));
        const float stdv = sqrtf(var);
        const float rstd = 1.f / (stdv + p.eps);
        const float lv = live ? 1.f : 0.f;
        cp_async_wait_all();
        __syncwarp();
        column_sums(acc_b2, row0);                   // g_b2 += sum over rows of d kin
        __syncwarp();

        float S1 = 0.f, S2 = 0.f, A[6], B[6], C[6], D[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { A[i] = 0.f; B[i] = 0.f; C[i] = 0.f; D[i] = 0.f; }
        float go8[8], gv8[8], qb[8];
        auto load_chunk = [&](int blk0, int c, float *f) {
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(my + (uint32_t)(blk0 + (c >> 3)) * 4096u + ((((uint32_t)c & 7u) ^ sw) << 4)));
            unpack8(q, f);
        };
        auto put_q = [&](int c) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + (uint32_t)(c >> 3) * 4096u + ((((uint32_t)c & 7u) ^ sw) << 4)),
                         "r"(pack_bf16(qb[0], qb[1])), "r"(pack_bf16(qb[2], qb[3])), "r"(pack_bf16(qb[4], qb[5])), "r"(pack_bf16(qb[6], qb[7])) : "memory");
        };
#pragma unroll
        for (int src = 0; src < 9; ++src) {
            float s = s0[src], c = c0[src];
#pragma unroll
            for (int slot = 0; slot < S; ++slot) {
                const int j = src * S + slot;
                if ((j & 7) == 0) load_chunk(0, j >> 3, go8);
                float val, coef;
                if (slot == 0) { val = g[src]; coef = 1.f; }
                else if (slot & 1) { val = s; coef = (float)(1 << ((slot - 1) >> 1)) * c; }
                else { val = c; coef = -(float)(1 << ((slot - 1) >> 1)) * s; }
                const float z = (val - mean) * rstd;
                const float go = go8[j & 7];
                const float gz = go * a2_s[j];
                S1 += gz; S2 = fmaf(gz, z, S2);
                qb[j & 7] = lv * go * z;
                if ((j & 7) == 7) put_q(j >> 3);
                if (src >= 3) {
                    const int jv = (src - 3) * S + slot;
                    if ((jv & 7) == 0 || (src == 3 && slot == 0)) load_chunk(NBK, jv >> 3, gv8);
                    A[src - 3] = fmaf(coef, gz, A[src - 3]);
                    B[src - 3] += coef;
                    C[src - 3] = fmaf(coef, z, C[src - 3]);
                    D[src - 3] = fmaf(coef, gv8[jv & 7], D[src - 3]);
                }
                if (slot && !(slot & 1)) {
                    const float s2 = 2.f * s * c, c2 = (c - s) * (c + s);
                    s = s2; c = c2;
                }
            }
        }
        if (DK & 7) {
#pragma unroll
            for (int e = DK & 7; e < 8; ++e) qb[e] = 0.f;
            put_q(DK >> 3);
        }
        if (live) {
            const float s1 = S1 * (1.f / DK);
            const float s2 = S2 / ((float)(DK - 1) * stdv);
            float dx[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) dx[i] = rstd * A[i] - rstd * s1 * B[i] - s2 * C[i] + D[i];
            const float e0 = dx[0] - dx[3], e1 = dx[1] - dx[4], e2 = dx[2] - dx[5];
            const float sd = (e0 * u[0] + e1 * u[1] + e2 * u[2]) / den;
            atomicAdd(p.g_points + (size_t)pidx * 3 + 0, dx[3] + u[0] * sd);
            atomicAdd(p.g_points + (size_t)pidx * 3 + 1, dx[4] + u[1] * sd);
            atomicAdd(p.g_points + (size_t)pidx * 3 + 2, dx[5] + u[2] * sd);
            // d feats = the tail of d vin (columns DPE .. DV-1), sent out four at a time as they complete
            float *gf = p.g_feats + (size_t)pidx * F;
            float grp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = DPE >> 3; c <= (DV - 1) >> 3; ++c) {
                float f8[8];
                load_chunk(NBK, c, f8);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int f = c * 8 + e - DPE;
                    if (f >= 0 && f < F) {
                        grp[f & 3] = f8[e];
                        if ((f & 3) == 3) red_add_v4(gf + (f - 3), grp[0], grp[1], grp[2], grp[3]);
                    }
                }
            }
        }
        __syncwarp();
        column_sums(acc_a2, row0);                   // g_a2 += sum over rows of d kin * z (bf16 products staged in place)
        __syncwarp();
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float sa = acc_a2[e] + __shfl_xor_sync(0xffffffffu, acc_a2[e], 16);
        const float sb = acc_b2[e] + __shfl_xor_sync(0xffffffffu, acc_b2[e], 16);
        const int j = cq * 8 + e;
        if (half == 0 && j < DK) { atomicAdd(p.g_a2 + j, sa); atomicAdd(p.g_b2 + j, sb); }
    }
}

// ------------------------------------------------------------------------------------------------ score + blend
struct ScoreParams {
    const uint8_t *h5;      // blocked bf16 [M_pad, 256] key stack output (before its LayerNorm)
    const float *h5_f32;    // optional fp32 tap [M,256] used instead of h5
    const float *ua;        // [R,256]  (W_k^T q' / sqrt(d)) * a2
    const float *cprime;    // [R]
    const float *influ;     // [P]
    const int32_t *idx;     // [R,K]
    const float *v;         // fp32 [M_pad, ldv]
    int64_t R;
    int K, C, ldv, score_relu, normalize, sc_ready;
    float bkg_score, eps;
    float *fused, *attn, *sc, *stats;       // [R,C], [R,K+1], [M], [M,2]
    // backward
    const float *d_fused, *d_attn;          // [R,C], [R,K+1] or null
    uint8_t *dv;                            // blocked bf16 [M_pad, 64*ceil(C/64)]
    float *d_score, *g_influ, *g_bv;        // [M], [P], [C]
    const float *d_score_in;                // key_score_bwd input
    uint8_t *dh5;                           // blocked bf16 [M_pad,256]
    float *dh5_f32;                         // optional fp32 tap
    float *zsum, *dssum, *g_b5;             // [R,256], [R], [256]
};

__device__ __forceinline__ void load_row8(const ScoreParams &p, int64_t row, int lane, float *h)
{
    if (p.h5_f32) {
        const float4 a = *reinterpret_cast<const float4 *>(p.h5_f32 + row * 256 + lane * 8);
        const float4 b = *reinterpret_cast<const float4 *>(p.h5_f32 + row * 256 + lane * 8 + 4);
        h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
    } else {
        unpack8(*reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(row, lane, 4)), h);
    }
}

// Raw attention scores, one (ray, candidate) row per lane (bf16 h5, K >= 16): score = ua . z(h5) + c' with z the row's
// LayerNorm-normalised key-stack output (attn.py:39-42, 117, 212-226 after the key-head fold of DESIGN.md section 3).
// The warp-per-ray form spends three 5-stage shuffle reductions per row; here a lane owns its row's 256 columns (staged
// through shared memory so the tile is read with coalesced 16-byte loads) and needs none.  Sums are taken about the row's
// first element c0: mean = c0 + S/256, sum (h-mean)^2 = Q - S^2/256, ua.(h-mean) = T - (mean-c0) * sum(ua).
// Twelve warps per CTA, one CTA per SM: a warp stages 16 KB, so they fill 192 KB of shared memory (with eight the kernel was
// latency-bound -- ncu r02: 12.5% warps active, 52% of the DRAM roof; 12 warps: 2.95 -> 2.71 ms for score + blend at 800x800).
constexpr int kScoreRowsWarps = 12;
__global__ void __launch_bounds__(kScoreRowsWarps * 32, 1) score_rows_kernel(const ScoreParams p)
{
    extern __shared__ __align__(16) uint8_t rows_stage[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wst = smem_u32(rows_stage) + (uint32_t)warp * 4u * 4096u;
    const uint32_t my = wst + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)lane & 7u;
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;
    for (int64_t row0 = ((int64_t)blockIdx.x * kScoreRowsWarps + warp) * 32; row0 < M_pad; row0 += (int64_t)gridDim.x * kScoreRowsWarps * 32) {
        if (row0 >= M) break;                                  // warp-uniform: only padding rows left
        {
            const int64_t tile = row0 >> 7;
            const uint8_t *src = p.h5 + (size_t)tile * 4 * kBlockBytes + (uint32_t)(row0 & 127) * 128u + (uint32_t)lane * 16u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    cp_async16(wst + (uint32_t)b * 4096u + (uint32_t)it * 512u + (uint32_t)lane * 16u, src + (size_t)b * kBlockBytes + it * 512);
            }
        }
        const int64_t row = row0 + lane;
        const bool live = row < M;
        const int64_t ray = (live ? row : M - 1) / p.K;
        // sum(ua) of the (at most three, K >= 16) rays this warp touches: one shuffle reduction each, then pick mine
        const int64_t ray_a = row0 / p.K;
        float usum = 0.f;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int64_t rr = ray_a + t;
            float part = 0.f;
            if (rr < p.R) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(p.ua + rr * 256 + lane * 8));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(p.ua + rr * 256 + lane * 8 + 4));
                part = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
            }
            part = warp_sum(part);
            if (ray == rr) usum = part;
        }
        cp_async_wait_all();
        __syncwarp();
        const float4 *uap = reinterpret_cast<const float4 *>(p.ua + ray * 256);
        float c0 = 0.f, S = 0.f, Q = 0.f, T = 0.f;
#pragma unroll 4
        for (int c = 0; c < 32; ++c) {
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(my + (uint32_t)(c >> 3) * 4096u + ((((uint32_t)c & 7u) ^ sw) << 4)));
            float h[8];
            unpack8(q, h);
            if (c == 0) c0 = h[0];
            const float4 u0 = __ldg(uap + 2 * c), u1 = __ldg(uap + 2 * c + 1);
            const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = h[e] - c0;
                S += d; Q = fmaf(d, d, Q); T = fmaf(d, uu[e], T);
            }
        }
        if (live) {
            const float dm = S * (1.f / 256.f);                      // mean - c0
            const float sq = fmaxf(Q - S * dm, 0.f);
            const float rstd = 1.f / (sqrtf(sq * (1.f / 255.f)) + p.eps);
            const float dot = (T - dm * usum) * rstd;                // ua . z, kept for the backward kernel
            float raw = dot + p.cprime[ray];
            if (p.score_relu) raw = fmaxf(raw, 0.f);
            p.sc[row] = raw;
            *reinterpret_cast<float4 *>(p.stats + row * 4) = make_float4(c0 + dm, rstd, dot, 0.f);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kRowThreads) score_blend_fwd_kernel(const ScoreParams p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        // The value rows do not depend on the softmax: with C <= 32 (one column per lane) all K of them are requested
        // before anything else, so their HBM latency overlaps the score / softmax chain (ncu r02: 17 warps stalled on
        // long_scoreboard per issue when each row was loaded inside the accumulation loop, one at a time).
        const bool v_ahead = p.C <= 32;
        float vv[32];
        if (v_ahead) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
                vv[k] = (k < p.K && lane < p.C) ? __ldg(p.v + (ray * p.K + k) * p.ldv + lane) : 0.f;
        }
        float my_sc = 0.f;
        if (p.sc_ready) {                    // raw scores and row statistics already written by score_rows_kernel
            if (lane < p.K) my_sc = p.sc[ray * p.K + lane];
        } else {
        float ua[8];
        {
            const float4 a = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8 + 4);
            ua[0] = a.x; ua[1] = a.y; ua[2] = a.z; ua[3] = a.w; ua[4] = b.x; ua[5] = b.y; ua[6] = b.z; ua[7] = b.w;
        }
        const float cp = p.cprime[ray];
        // bf16 rows are fetched one candidate ahead (4 registers) so the HBM latency overlaps the three warp reductions
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (!p.h5_f32) nxt = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(ray * p.K, lane, 4));
        for (int k = 0; k < p.K; ++k) {
            const int64_t row = ray * p.K + k;
            float h[8], s = 0.f;
            if (p.h5_f32) load_row8(p, row, lane, h);
            else {
                const uint4 cur = nxt;
                if (k + 1 < p.K) nxt = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(row + 1, lane, 4));
                unpack8(cur, h);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) s += h[e];
            const float mean = warp_sum(s) * (1.f / 256.f);
            float sq = 0.f, dot = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float c = h[e] - mean; sq += c * c; dot += c * ua[e]; }
            sq = warp_sum(sq); dot = warp_sum(dot);
            const float rstd = 1.f / (sqrtf(sq * (1.f / 255.f)) + p.eps);
            float raw = dot * rstd + cp;
            if (p.score_relu) raw = fmaxf(raw, 0.f);
            if (lane == k) my_sc = raw;
            if (lane == 0) *reinterpret_cast<float4 *>(p.stats + row * 4) = make_float4(mean, rstd, dot * rstd, 0.f);
        }
        }
        // model.py:524-533: influence scores, background token, softmax, top-K renormalisation
        // (the background token is a warp-uniform scalar, not a lane, so K may use all 32 lanes)
        float s = -INFINITY;
        if (lane < p.K) {
            if (!p.sc_ready) p.sc[ray * p.K + lane] = my_sc;
            s = my_sc * __ldg(p.influ + p.idx[ray * p.K + lane]);
        }
        const float mx = fmaxf(warp_max(s), p.bkg_score);
        const float e = (lane < p.K) ? expf(s - mx) : 0.f;
        const float eb = expf(p.bkg_score - mx);
        const float tot = warp_sum(e) + eb;
        const float attn = e / tot;
        if (lane < p.K) p.attn[ray * (p.K + 1) + lane] = attn;
        if (lane == 0) p.attn[ray * (p.K + 1) + p.K] = eb / tot;
        const float topk = warp_sum(attn);
        const float w = p.normalize ? attn / topk : attn;
        if (v_ahead) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = fmaf(__shfl_sync(0xffffffffu, w, k), vv[k], acc);    // lanes >= K carry w = 0
            if (lane < p.C) p.fused[ray * p.C + lane] = acc;
        } else
        for (int c0 = 0; c0 < p.C; c0 += 32) {
            const int c = c0 + lane;
            float acc = 0.f;
            for (int k = 0; k < p.K; ++k) {
                const float wk = __shfl_sync(0xffffffffu, w, k);
                if (c < p.C) acc += wk * p.v[(ray * p.K + k) * p.ldv + c];
            }
            if (c < p.C) p.fused[ray * p.C + c] = acc;
        }
    }
}

__global__ void __launch_bounds__(kRowThreads) blend_bwd_kernel(const ScoreParams p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.C + 63) / 64;
    float acc_bv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc_bv[e] = 0.f;
    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        const float attn = (lane < p.K) ? p.attn[ray * (p.K + 1) + lane] : 0.f;
        const float attn_b = p.attn[ray * (p.K + 1) + p.K];          // background weight: warp-uniform, K may be 32
        const float topk = warp_sum(attn);
        const float w = p.normalize ? attn / topk : attn;
        // d w_k = d_fused . v_k   (lane k)
        float dw = 0.f;
        if (lane < p.K) {
            const float *vr = p.v + (ray * p.K + lane) * p.ldv;
            const float *df = p.d_fused + ray * p.C;
            if (((p.C | p.ldv) & 3) == 0) {
                for (int c = 0; c < p.C; c += 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(df + c));
                    const float4 b = *reinterpret_cast<const float4 *>(vr + c);
                    dw = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, dw))));
                }
            } else {
                for (int c = 0; c < p.C; ++c) dw += __ldg(df + c) * vr[c];
            }
        }
        float da;                                   // d attn
        if (p.normalize) {
            const float mix = warp_sum(lane < p.K ? dw * w : 0.f);
            da = (lane < p.K) ? (dw - mix) / topk : 0.f;
        } else da = (lane < p.K) ? dw : 0.f;
        if (p.d_attn && lane < p.K) da += p.d_attn[ray * (p.K + 1) + lane];
        const float da_b = p.d_attn ? p.d_attn[ray * (p.K + 1) + p.K] : 0.f;
        const float inner = warp_sum(attn * da) + attn_b * da_b;
        const float ds = (lane < p.K) ? attn * (da - inner) : 0.f;        // softmax backward
        if (lane < p.K) {
            const float sc = p.sc[ray * p.K + lane];
            const int pi = p.idx[ray * p.K + lane];
            const float influ = __ldg(p.influ + pi);
            atomicAdd(p.g_influ + pi, ds * sc);
            float dsc = ds * influ;
            if (p.score_relu && !(sc > 0.f)) dsc = 0.f;
            p.d_score[ray * p.K + lane] = dsc;
        }
        // d v_k = w_k * d_fused  -> tile-blocked bf16 operand of the value stack's backward (C <= 64: one block)
        {
            const int c = lane;
            const bool writer = c < nblk * 8;
            float df[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) df[e] = (writer && c * 8 + e < p.C) ? __ldg(p.d_fused + ray * p.C + c * 8 + e) : 0.f;
            for (int k = 0; k < p.K; ++k) {
                const float wk = __shfl_sync(0xffffffffu, w, k);
                if (writer) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = wk * df[e];
                    const uint4 q = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                    float r[8];
                    unpack8(q, r);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc_bv[e] += r[e];
                    *reinterpret_cast<uint4 *>(p.dv + blocked_chunk_offset(ray * p.K + k, c, nblk)) = q;
                }
            }
        }
    }
    // bias gradient of the last value layer: column sums of d v (only the first 64 columns are tracked per lane)
    if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (lane * 8 + e < p.C) atomicAdd(p.g_bv + lane * 8 + e, acc_bv[e]);
    }
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < (M_pad - M) * nblk * 8; i += (int64_t)gridDim.x * kRowThreads)
        *reinterpret_cast<uint4 *>(p.dv + blocked_chunk_offset(M + i / (nblk * 8), (int)(i % (nblk * 8)), nblk)) = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kRowThreads, 4) key_score_bwd_kernel(const ScoreParams p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc_b5[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc_b5[e] = 0.f;
    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        float ua[8], zs[8];
        {
            const float4 a = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8 + 4);
            ua[0] = a.x; ua[1] = a.y; ua[2] = a.z; ua[3] = a.w; ua[4] = b.x; ua[5] = b.y; ua[6] = b.z; ua[7] = b.w;
        }
        float uas = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { uas += ua[e]; zs[e] = 0.f; }
        const float ua_mean = warp_sum(uas) * (1.f / 256.f);
        float dss = 0.f;
        // bf16 rows are fetched three candidates ahead (12 registers): one 512-byte row in flight per warp leaves the kernel
        // latency-bound at ~2.9 TB/s (ncu: long_scoreboard 12 of 17 stall cycles per issue)
        uint4 pf0 = make_uint4(0, 0, 0, 0), pf1 = pf0, pf2 = pf0;
        if (!p.h5_f32) {
            pf0 = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(ray * p.K, lane, 4));
            if (1 < p.K) pf1 = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(ray * p.K + 1, lane, 4));
            if (2 < p.K) pf2 = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(ray * p.K + 2, lane, 4));
        }
        float ds_n = p.d_score_in[ray * p.K];
        float4 st_n = *reinterpret_cast<const float4 *>(p.stats + ray * p.K * 4);
        for (int k = 0; k < p.K; ++k) {
            const int64_t row = ray * p.K + k;
            // mean, 1/(std+eps) and dot = ua . z come from the forward kernel: no reduction over the row is needed here
            const float ds = ds_n, mean = st_n.x, rstd = st_n.y, dot = st_n.z;
            float h[8], y[8];
            if (p.h5_f32) load_row8(p, row, lane, h);
            else {
                const uint4 cur = pf0;
                pf0 = pf1; pf1 = pf2;
                if (k + 3 < p.K) pf2 = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(row + 3, lane, 4));
                unpack8(cur, h);
            }
            if (k + 1 < p.K) { ds_n = p.d_score_in[row + 1]; st_n = *reinterpret_cast<const float4 *>(p.stats + (row + 1) * 4); }
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = (h[e] - mean) * rstd;
            const float sigma = 1.f / rstd - p.eps;
            const float coef = ds * dot / (255.f * sigma);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                f[e] = rstd * ds * (ua[e] - ua_mean) - y[e] * coef;
                zs[e] += ds * y[e];
            }
            dss += ds;
            const uint4 q = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            float r[8];
            unpack8(q, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc_b5[e] += r[e];
            *reinterpret_cast<uint4 *>(p.dh5 + blocked_chunk_offset(row, lane, 4)) = q;
            if (p.dh5_f32) {
#pragma unroll
                for (int e = 0; e < 8; ++e) p.dh5_f32[row * 256 + lane * 8 + e] = f[e];
            }
        }
        *reinterpret_cast<float4 *>(p.zsum + ray * 256 + lane * 8) = make_float4(zs[0], zs[1], zs[2], zs[3]);
        *reinterpret_cast<float4 *>(p.zsum + ray * 256 + lane * 8 + 4) = make_float4(zs[4], zs[5], zs[6], zs[7]);
        if (lane == 0) p.dssum[ray] = dss;
    }
    if (p.g_b5) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(p.g_b5 + lane * 8 + e, acc_b5[e]);
    }
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < (M_pad - M) * 32; i += (int64_t)gridDim.x * kRowThreads)
        *reinterpret_cast<uint4 *>(p.dh5 + blocked_chunk_offset(M + i / 32, (int)(i % 32), 4)) = make_uint4(0, 0, 0, 0);
}


// The same for the bf16 path proper (tile-blocked h5 in, tile-blocked dh5 out, no fp32 taps, the bias gradient left to
// papr_wgrad_bias_bf16), walked BLOCK by block: for each 64-column block of the ray's K rows, eight lanes take the eight
// 16-byte chunks of a row and four such row groups sit side by side, so one warp instruction moves 512 CONTIGUOUS bytes
// (4 rows x 128 B of one block) instead of four 128-byte pieces 16 KB apart, and all of a block's row groups are
// in flight before the first is used.  The warp-per-row walk above leaves the DRAM pages it touches after 128 bytes and
// ran at 2.9 TB/s (ncu r02: 26% of the stall samples on the move that consumes the prefetched row).
__global__ void __launch_bounds__(kRowThreads, 3) key_score_bwd_blocks_kernel(const ScoreParams p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane >> 3, ch = lane & 7;
    constexpr int G = 5;                                         // row groups fetched together (K = 20: all of them)
    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        float uas;
        {
            const float4 a = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + lane * 8 + 4);
            uas = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
        }
        const float ua_mean = warp_sum(uas) * (1.f / 256.f);
        const float dsl = lane < p.K ? p.d_score_in[ray * p.K + lane] : 0.f;
        const float dss = warp_sum(dsl);
        if (lane == 0) p.dssum[ray] = dss;
        for (int b = 0; b < 4; ++b) {
            float ua8[8], zs[8];
            {
                const float4 a = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + b * 64 + ch * 8);
                const float4 c = *reinterpret_cast<const float4 *>(p.ua + ray * 256 + b * 64 + ch * 8 + 4);
                ua8[0] = a.x - ua_mean; ua8[1] = a.y - ua_mean; ua8[2] = a.z - ua_mean; ua8[3] = a.w - ua_mean;
                ua8[4] = c.x - ua_mean; ua8[5] = c.y - ua_mean; ua8[6] = c.z - ua_mean; ua8[7] = c.w - ua_mean;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) zs[e] = 0.f;
            for (int k0 = 0; k0 < p.K; k0 += 4 * G) {
                uint4 hq[G];
                float4 st[G];
                float ds[G];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int k = k0 + 4 * i + rg;
                    hq[i] = make_uint4(0, 0, 0, 0); st[i] = make_float4(0.f, 1.f, 0.f, 0.f); ds[i] = 0.f;
                    if (k < p.K) {
                        const int64_t row = ray * p.K + k;
                        hq[i] = *reinterpret_cast<const uint4 *>(p.h5 + blocked_chunk_offset(row, b * 8 + ch, 4));
                        st[i] = *reinterpret_cast<const float4 *>(p.stats + row * 4);
                        ds[i] = p.d_score_in[row];
                    }
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int k = k0 + 4 * i + rg;
                    if (k < p.K) {
                        const int64_t row = ray * p.K + k;
                        const float mean = st[i].x, rstd = st[i].y, dot = st[i].z;
                        float h[8], f[8];
                        unpack8(hq[i], h);
                        const float sigma = 1.f / rstd - p.eps;
                        const float coef = ds[i] * dot / (255.f * sigma);
                        const float ca = rstd * ds[i];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float y = (h[e] - mean) * rstd;
                            f[e] = ca * ua8[e] - y * coef;
                            zs[e] += ds[i] * y;
                        }
                        *reinterpret_cast<uint4 *>(p.dh5 + blocked_chunk_offset(row, b * 8 + ch, 4)) =
                            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                zs[e] += __shfl_xor_sync(0xffffffffu, zs[e], 8);
                zs[e] += __shfl_xor_sync(0xffffffffu, zs[e], 16);
            }
            if (rg == 0) {
                *reinterpret_cast<float4 *>(p.zsum + ray * 256 + b * 64 + ch * 8) = make_float4(zs[0], zs[1], zs[2], zs[3]);
                *reinterpret_cast<float4 *>(p.zsum + ray * 256 + b * 64 + ch * 8 + 4) = make_float4(zs[4], zs[5], zs[6], zs[7]);
            }
        }
    }
    const int64_t M = p.R * p.K, M_pad = (M + 127) / 128 * 128;
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < (M_pad - M) * 32; i += (int64_t)gridDim.x * kRowThreads)
        *reinterpret_cast<uint4 *>(p.dh5 + blocked_chunk_offset(M + i / 32, (int)(i % 32), 4)) = make_uint4(0, 0, 0, 0);
}


// ------------------------------------------------------------------------------------------------ query tail
// Per ray: z = (q5 - mean)/(std + eps) of the query stack's output (attn.py:39-42 without the affine part, which the
// host folds -- together with w_q, w_k and the key out-norm -- into one 256x256 matrix applied by papr_linear_bf16) and
// c' = w_c . z + c_const.  Warp per ray, eight columns per lane.
struct QueryTailParams {
    const float *q5;        // [R,256]
    const float *wc;        // [256]
    float c_const, eps;
    int64_t R;
    uint8_t *z;             // blocked bf16 [R_pad,256]
    float *stats, *cprime;  // [R,2], [R]
    // backward
    const float *dz, *dc;   // [R,256] (ld), [R]
    int64_t ld_dz;
    float *dq5, *g_wc, *g_cc;   // [R,256], [256] +=, [1] +=
};

__global__ void __launch_bounds__(kRowThreads) query_tail_fwd_kernel(const QueryTailParams p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) wc[e] = p.wc[lane * 8 + e];
    for (int64_t ray = (int64_t)blockIdx.x * kRowWarps + warp; ray < p.R; ray += (int64_t)gridDim.x * kRowWarps) {
        const float4 a = *reinterpret_cast<const float4 *>(p.q5 + ray * 256 + lane * 8);
        const float4 b = *reinterpret_cast<const float4 *>(p.q5 + ray * 256 + lane * 8 + 4);
        const float h[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += h[e];
        const float mean = warp_sum(s) * (1.f / 256.f);
        float sq = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float c = h[e] - mean; sq += c * c; }
        const float rstd = 1.f / (sqrtf(warp_sum(sq) * (1.f / 255.f)) + p.eps);
        float z[8], dot = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { z[e] = (h[e] - mean) * rstd; dot += z[e] * wc[e]; }
        dot = warp_sum(dot);
        *reinterpret_cast<uint4 *>(p.z + blocked_chunk_offset(ray, lane, 4)) =
            make_uint4(pack_bf16(z[0], z[1]), pack_bf16(z[2], z[3]), pack_bf16(z[4], z[5]), pack_bf16(z[6], z[7]));
        if (lane == 0) { p.stats[ray * 2] = mean; p.stats[ray * 2 + 1] = rstd; p.cprime[ray] = dot + p.c_const; }
    }
    const int64_t R_pad = (p.R + 127) / 128 * 128;
    for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < (R_pad - p.R) * 32; i += (int64_t)gridDim.x * kRowThreads)
        *reinterpret_cast<uint4 *>(p.z + blocked_chunk_offset(p.R + i / 32, (int)(i % 32), 4)) = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kRowThreads) query_tail_bwd_kernel(const QueryTailParams p)
{
    // U rays per warp iteration: their loads are all in flight before the first of the two warp reductions per ray (with one
    // ray at a time the kernel issued 10% of its slots and reached 25% of the DRAM roof, ncu r02)
    constexpr int U = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wc[8], acc_wc[8], acc_cc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { wc[e] = p.wc[lane * 8 + e]; acc_wc[e] = 0.f; }
    const int64_t stride = (int64_t)gridDim.x * kRowWarps;
    for (int64_t ray0 = (int64_t)blockIdx.x * kRowWarps + warp; ray0 < p.R; ray0 += stride * U) {
        float h[U][8], g[U][8], mean[U], rstd[U], dc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ray = ray0 + u * stride;
            const bool on = ray < p.R;
            const int64_t r = on ? ray : ray0;
            const float4 a = *reinterpret_cast<const float4 *>(p.q5 + r * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4 *>(p.q5 + r * 256 + lane * 8 + 4);
            const float4 ga = *reinterpret_cast<const float4 *>(p.dz + r * p.ld_dz + lane * 8);
            const float4 gb = *reinterpret_cast<const float4 *>(p.dz + r * p.ld_dz + lane * 8 + 4);
            h[u][0] = a.x; h[u][1] = a.y; h[u][2] = a.z; h[u][3] = a.w; h[u][4] = b.x; h[u][5] = b.y; h[u][6] = b.z; h[u][7] = b.w;
            g[u][0] = ga.x; g[u][1] = ga.y; g[u][2] = ga.z; g[u][3] = ga.w; g[u][4] = gb.x; g[u][5] = gb.y; g[u][6] = gb.z; g[u][7] = gb.w;
            mean[u] = p.stats[r * 2]; rstd[u] = p.stats[r * 2 + 1]; dc[u] = on ? p.dc[r] : 0.f;
        }
        float s1[U], s2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s1[u] = 0.f; s2[u] = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                h[u][e] = (h[u][e] - mean[u]) * rstd[u];        // z
                if (ray0 + u * stride < p.R) acc_wc[e] += dc[u] * h[u][e];
                g[u][e] += dc[u] * wc[e];                       // d z = (d ua) A + d c' * w_c
                s1[u] += g[u][e]; s2[u] += g[u][e] * h[u][e];
            }
            acc_cc += dc[u];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
                s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ray = ray0 + u * stride;
            if (ray >= p.R) continue;
            const float m1 = s1[u] * (1.f / 256.f);
            const float sigma = 1.f / rstd[u] - p.eps;
            const float m2 = s2[u] / (255.f * sigma);
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = rstd[u] * (g[u][e] - m1) - h[u][e] * m2;
            *reinterpret_cast<float4 *>(p.dq5 + ray * 256 + lane * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4 *>(p.dq5 + ray * 256 + lane * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
    }
    // one set of 257 global atomics per CTA, not per warp (they all land on the same 257 addresses)
    __shared__ float red_s[kRowWarps][257];
#pragma unroll
    for (int e = 0; e < 8; ++e) red_s[warp][lane * 8 + e] = acc_wc[e];
    if (lane == 0) red_s[warp][256] = acc_cc;
    __syncthreads();
    for (int j = threadIdx.x; j < 257; j += kRowThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) t += red_s[w][j];
        atomicAdd(j < 256 ? p.g_wc + j : p.g_cc, t);
    }
}

// ------------------------------------------------------------------------------------------------ query prologue
// Per ray: the query stack's input q = a_2 * (pe - mean) / (std + eps) + b_2 with pe = [d_c, sin(2^i d_c), cos(2^i d_c)]
// for the three components of the ray direction (models/utils.py:232-242 embed_type 1, attn.py:30-42 in-norm, unbiased
// std).  One ray per thread, everything in registers; a CTA stages its 256 rows in shared memory so that the (R, 3S) fp32
// output is written -- and, in the backward kernel, the incoming gradient read -- with coalesced accesses.  The ray
// direction is an input of the scene, so the backward kernel produces only the gradients of the affine pair:
// g_a2[j] += sum_r dq[r, j] * z[r, j],  g_b2[j] += sum_r dq[r, j].
constexpr int kQpThreads = 256;

template <int L>
__device__ __forceinline__ void query_pe_normalised(const float *d, float eps, float *z)
{
    constexpr int S = 1 + 2 * L, D = 3 * S;
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pe_fill(d[c], L, z + c * S);
#pragma unroll
        for (int j = 0; j < S; ++j) { sum += z[c * S + j]; sq = fmaf(z[c * S + j], z[c * S + j], sq); }
    }
    const float mean = sum * (1.f / D);
    const float var = fmaxf(sq - (float)D * mean * mean, 0.f) * (1.f / (D - 1));
    const float rstd = 1.f / (sqrtf(var) + eps);
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = (z[j] - mean) * rstd;
}

template <int L>
__global__ void __launch_bounds__(kQpThreads) query_prologue_fwd_kernel(const float *__restrict__ rays_d, const float *__restrict__ a2,
                                                                        const float *__restrict__ b2, int64_t R, float eps, float *__restrict__ q)
{
    constexpr int D = 3 * (1 + 2 * L);
    __shared__ float stage[kQpThreads * D];
    __shared__ float a_s[D], b_s[D];
    if (threadIdx.x < D) { a_s[threadIdx.x] = a2[threadIdx.x]; b_s[threadIdx.x] = b2[threadIdx.x]; }
    __syncthreads();
    for (int64_t r0 = (int64_t)blockIdx.x * kQpThreads; r0 < R; r0 += (int64_t)gridDim.x * kQpThreads) {
        const int64_t r = r0 + threadIdx.x;
        if (r < R) {
            const float d[3] = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
            float z[D];
            query_pe_normalised<L>(d, eps, z);
#pragma unroll
            for (int j = 0; j < D; ++j) stage[threadIdx.x * D + j] = fmaf(a_s[j], z[j], b_s[j]);      // odd row stride: no bank conflicts
        }
        __syncthreads();
        const int64_t n = (R - r0 < kQpThreads ? R - r0 : kQpThreads) * D;
        for (int64_t i = threadIdx.x; i < n; i += kQpThreads) q[r0 * D + i] = stage[i];
        __syncthreads();
    }
}

template <int L>
__global__ void __launch_bounds__(kQpThreads) query_prologue_bwd_kernel(const float *__restrict__ rays_d, const float *__restrict__ dq, int64_t R,
                                                                        float eps, float *__restrict__ g_a2, float *__restrict__ g_b2)
{
    constexpr int D = 3 * (1 + 2 * L);
    __shared__ float stage[kQpThreads * D];
    __shared__ float red[2 * D];
    if (threadIdx.x < 2 * D) red[threadIdx.x] = 0.f;
    float acc_a[D], acc_b[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { acc_a[j] = 0.f; acc_b[j] = 0.f; }
    for (int64_t r0 = (int64_t)blockIdx.x * kQpThreads; r0 < R; r0 += (int64_t)gridDim.x * kQpThreads) {
        const int64_t n = (R - r0 < kQpThreads ? R - r0 : kQpThreads) * D;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n; i += kQpThreads) stage[i] = dq[r0 * D + i];
        __syncthreads();
        const int64_t r = r0 + threadIdx.x;
        if (r < R) {
            const float d[3] = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
            float z[D];
            query_pe_normalised<L>(d, eps, z);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float g = stage[threadIdx.x * D + j];
                acc_a[j] = fmaf(g, z[j], acc_a[j]);
                acc_b[j] += g;
            }
        }
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float sa = warp_sum(acc_a[j]), sb = warp_sum(acc_b[j]);
        if (lane == 0) { atomicAdd(&red[j], sa); atomicAdd(&red[D + j], sb); }
    }
    __syncthreads();
    if (threadIdx.x < D) { atomicAdd(g_a2 + threadIdx.x, red[threadIdx.x]); atomicAdd(g_b2 + threadIdx.x, red[D + threadIdx.x]); }
}

static int row_grid(int64_t R)
{
    const int64_t blocks = (R + kRowWarps - 1) / kRowWarps;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace papr

using namespace papr;

// row-per-lane prologues exist for the shipped shapes: PE order 6 (nerfsyn) / 4 (Tanks&Temples), feature width 64 / 128 (materials.yml)
template <int L, int F>
static int launch_prologue_fwd_rows(const PrologueParams &p, cudaStream_t stream)
{
    constexpr int S = 1 + 2 * L, NBK = (9 * S + 63) / 64, NBV = (6 * S + F + 63) / 64, NBS = NBK > NBV ? NBK : NBV;
    constexpr int smem = kRowWarps * NBS * 4096;
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, attn_prologue_fwd_rows_kernel<L, F>, smem));
    const int64_t groups = ((p.R * p.K + 127) / 128 * 128 + kRowThreads - 1) / kRowThreads;
    attn_prologue_fwd_rows_kernel<L, F><<<(int)(groups < 2 * kNumSMs ? groups : 2 * kNumSMs), kRowThreads, smem, stream>>>(p);
    return check_launch();
}

template <int L, int F>
static int launch_prologue_bwd_rows(const PrologueParams &p, cudaStream_t stream)
{
    constexpr int S = 1 + 2 * L, NBK = (9 * S + 63) / 64, NBV = (6 * S + F + 63) / 64;
    constexpr int smem = 4 * (NBK + NBV) * 4096;
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, attn_prologue_bwd_rows_kernel<L, F>, smem));
    const int64_t groups = ((p.R * p.K + 127) / 128 * 128 + 127) / 128;
    attn_prologue_bwd_rows_kernel<L, F><<<(int)(groups < 2 * kNumSMs ? groups : 2 * kNumSMs), 128, smem, stream>>>(p);
    return check_launch();
}

static bool rows_shape(const PrologueParams &p, int L, int F)
{
    const int S = 1 + 2 * L;
    return p.L == L && p.F == F && p.nblk_k == (9 * S + 63) / 64 && p.nblk_v == (6 * S + F + 63) / 64;
}

static int prologue_check(int64_t R, int64_t rays_per_view, int K, int L, int F, int dk_pad, int dv_pad)
{
    if (R <= 0 || rays_per_view <= 0 || K < 1 || K > 32 || L < 0 || L > 6 || F < 0) return PAPR_ERR_INVALID_ARGUMENT;
    const int S = 1 + 2 * L, dk = 9 * S, dv = 6 * S + F;
    if (dk < 2 || dk > kMaxDk || dv > kMaxDv || dk_pad % 64 || dv_pad % 64 || dk_pad < dk || dv_pad < dv) return PAPR_ERR_INVALID_ARGUMENT;
    return PAPR_OK;
}

extern "C" int papr_attn_prologue_fwd(const float *rays_o, const float *rays_d, const float *points, const float *feats,
                                      const int32_t *idx, const float *ln_a, const float *ln_b, int64_t R,
                                      int64_t rays_per_view, int K, int L, int F, float eps, void *kin, int dk_pad,
                                      void *vin, int dv_pad, float *kin_f32, float *vin_f32, void *stream)
{
    if (!rays_o || !rays_d || !points || !idx || !ln_a || !ln_b || !kin || !vin || (F > 0 && !feats)) return PAPR_ERR_INVALID_ARGUMENT;
    int st = prologue_check(R, rays_per_view, K, L, F, dk_pad, dv_pad);
    if (st) return st;
    PrologueParams p = {};
    p.rays_o = rays_o; p.rays_d = rays_d; p.points = points; p.feats = feats; p.a2 = ln_a; p.b2 = ln_b; p.idx = idx;
    p.R = R; p.rays_per_view = rays_per_view; p.K = K; p.L = L; p.F = F; p.dk = 9 * (1 + 2 * L); p.dv = 6 * (1 + 2 * L) + F;
    p.nblk_k = dk_pad / 64; p.nblk_v = dv_pad / 64; p.eps = eps;
    p.kin = (uint8_t *)kin; p.vin = (uint8_t *)vin; p.kin_f32 = kin_f32; p.vin_f32 = vin_f32;
    if (kin_f32 || vin_f32) attn_prologue_fwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 6, 64)) return launch_prologue_fwd_rows<6, 64>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 4, 64)) return launch_prologue_fwd_rows<4, 64>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 6, 128)) return launch_prologue_fwd_rows<6, 128>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 4, 128)) return launch_prologue_fwd_rows<4, 128>(p, (cudaStream_t)stream);
    else attn_prologue_fwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);      // any other PE order / feature width
    return check_launch();
}

extern "C" int papr_attn_prologue_bwd(const float *rays_o, const float *rays_d, const float *points, const int32_t *idx,
                                      const float *ln_a, int64_t R, int64_t rays_per_view, int K, int L, int F, float eps,
                                      const void *dkin, int dk_pad, const void *dvin, int dv_pad, const float *dkin_f32,
                                      const float *dvin_f32, float *g_points, float *g_feats, float *g_ln_a,
                                      float *g_ln_b, void *stream)
{
    if (!rays_o || !rays_d || !points || !idx || !ln_a || !g_points || !g_ln_a || !g_ln_b || (F > 0 && !g_feats)) return PAPR_ERR_INVALID_ARGUMENT;
    if ((!dkin && !dkin_f32) || (!dvin && !dvin_f32)) return PAPR_ERR_INVALID_ARGUMENT;
    int st = prologue_check(R, rays_per_view, K, L, F, dk_pad, dv_pad);
    if (st) return st;
    PrologueParams p = {};
    p.rays_o = rays_o; p.rays_d = rays_d; p.points = points; p.a2 = ln_a; p.idx = idx;
    p.R = R; p.rays_per_view = rays_per_view; p.K = K; p.L = L; p.F = F; p.dk = 9 * (1 + 2 * L); p.dv = 6 * (1 + 2 * L) + F;
    p.nblk_k = dk_pad / 64; p.nblk_v = dv_pad / 64; p.eps = eps;
    p.dkin = (const uint8_t *)dkin; p.dvin = (const uint8_t *)dvin; p.dkin_f32 = dkin_f32; p.dvin_f32 = dvin_f32;
    p.g_points = g_points; p.g_feats = g_feats; p.g_a2 = g_ln_a; p.g_b2 = g_ln_b;
    if (dkin_f32 || dvin_f32) attn_prologue_bwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 6, 64)) return launch_prologue_bwd_rows<6, 64>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 4, 64)) return launch_prologue_bwd_rows<4, 64>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 6, 128)) return launch_prologue_bwd_rows<6, 128>(p, (cudaStream_t)stream);
    else if (!getenv("PAPR_PROLOGUE_GENERIC") && rows_shape(p, 4, 128)) return launch_prologue_bwd_rows<4, 128>(p, (cudaStream_t)stream);
    else attn_prologue_bwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);      // any other PE order / feature width
    return check_launch();
}

extern "C" int papr_score_blend_fwd(const void *h5, const float *h5_f32, const float *ua, const float *cprime,
                                    const float *influ, const int32_t *idx, const float *v, int64_t ldv, int64_t R, int K,
                                    int C, int score_relu, int normalize, float bkg_score, float eps, float *fused,
                                    float *attn, float *sc, float *stats, void *stream)
{
    if ((!h5 && !h5_f32) || !ua || !cprime || !influ || !idx || !v || !fused || !attn || !sc || !stats) return PAPR_ERR_INVALID_ARGUMENT;
    if (R <= 0 || K < 1 || K > 32 || C < 1 || ldv < C) return PAPR_ERR_INVALID_ARGUMENT;
    ScoreParams p = {};
    p.h5 = (const uint8_t *)h5; p.h5_f32 = h5_f32; p.ua = ua; p.cprime = cprime; p.influ = influ; p.idx = idx; p.v = v;
    p.R = R; p.K = K; p.C = C; p.ldv = (int)ldv; p.score_relu = score_relu; p.normalize = normalize;
    p.bkg_score = bkg_score; p.eps = eps; p.fused = fused; p.attn = attn; p.sc = sc; p.stats = stats;
    if (h5 && !h5_f32 && K >= 16 && !getenv("PAPR_SCORE_WARP")) {
        constexpr int smem = kScoreRowsWarps * 4 * 4096, threads = kScoreRowsWarps * 32;
        static SmemAttrOnce once;
        PAPR_CUDA_TRY(ensure_dyn_smem(once, score_rows_kernel, smem));
        const int64_t groups = (R * K + threads - 1) / threads;
        score_rows_kernel<<<(int)(groups < kNumSMs ? groups : kNumSMs), threads, smem, (cudaStream_t)stream>>>(p);
        PAPR_CUDA_TRY(cudaGetLastError());
        p.sc_ready = 1;
    }
    score_blend_fwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_blend_bwd(const float *d_fused, const float *d_attn, const float *attn, const float *sc,
                              const float *influ, const int32_t *idx, const float *v, int64_t ldv, int64_t R, int K, int C,
                              int score_relu, int normalize, void *dv_blocked, float *d_score, float *g_influ,
                              float *g_bias_v, void *stream)
{
    if (!d_fused || !attn || !sc || !influ || !idx || !v || !dv_blocked || !d_score || !g_influ || !g_bias_v) return PAPR_ERR_INVALID_ARGUMENT;
    if (R <= 0 || K < 1 || K > 32 || C < 1 || C > 64 || ldv < C) return PAPR_ERR_INVALID_ARGUMENT;
    ScoreParams p = {};
    p.d_fused = d_fused; p.d_attn = d_attn; p.attn = const_cast<float *>(attn); p.sc = const_cast<float *>(sc); p.influ = influ; p.idx = idx; p.v = v;
    p.R = R; p.K = K; p.C = C; p.ldv = (int)ldv; p.score_relu = score_relu; p.normalize = normalize;
    p.dv = (uint8_t *)dv_blocked; p.d_score = d_score; p.g_influ = g_influ; p.g_bv = g_bias_v;
    blend_bwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_key_score_bwd(const float *d_score, const void *h5, const float *h5_f32, const float *stats,
                                  const float *ua, int64_t R, int K, float eps, void *dh5_blocked, float *dh5_f32,
                                  float *zsum, float *dssum, float *g_bias5, void *stream)
{
    if (!d_score || (!h5 && !h5_f32) || !stats || !ua || !dh5_blocked || !zsum || !dssum) return PAPR_ERR_INVALID_ARGUMENT;
    if (R <= 0 || K < 1 || K > 32) return PAPR_ERR_INVALID_ARGUMENT;
    ScoreParams p = {};
    p.d_score_in = d_score; p.h5 = (const uint8_t *)h5; p.h5_f32 = h5_f32; p.stats = const_cast<float *>(stats); p.ua = ua; p.R = R; p.K = K;
    p.eps = eps; p.dh5 = (uint8_t *)dh5_blocked; p.dh5_f32 = dh5_f32; p.zsum = zsum; p.dssum = dssum; p.g_b5 = g_bias5;
    if (h5 && !h5_f32 && !dh5_f32 && !g_bias5 && !getenv("PAPR_KEY_SCORE_ROWWISE"))
        key_score_bwd_blocks_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    else
        key_score_bwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_query_prologue_fwd(const float *rays_d, const float *a2, const float *b2, int64_t R, int L, float eps,
                                       float *q, void *stream)
{
    if (!rays_d || !a2 || !b2 || !q || R <= 0) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t groups = (R + kQpThreads - 1) / kQpThreads;
    const int grid = (int)(groups < 4 * kNumSMs ? groups : 4 * kNumSMs);
    if (L == 6) query_prologue_fwd_kernel<6><<<grid, kQpThreads, 0, (cudaStream_t)stream>>>(rays_d, a2, b2, R, eps, q);
    else if (L == 4) query_prologue_fwd_kernel<4><<<grid, kQpThreads, 0, (cudaStream_t)stream>>>(rays_d, a2, b2, R, eps, q);
    else return PAPR_ERR_INVALID_ARGUMENT;
    return check_launch();
}

extern "C" int papr_query_prologue_bwd(const float *rays_d, const float *dq, int64_t R, int L, float eps, float *g_a2,
                                       float *g_b2, void *stream)
{
    if (!rays_d || !dq || !g_a2 || !g_b2 || R <= 0) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t groups = (R + kQpThreads - 1) / kQpThreads;
    const int grid = (int)(groups < 2 * kNumSMs ? groups : 2 * kNumSMs);
    if (L == 6) query_prologue_bwd_kernel<6><<<grid, kQpThreads, 0, (cudaStream_t)stream>>>(rays_d, dq, R, eps, g_a2, g_b2);
    else if (L == 4) query_prologue_bwd_kernel<4><<<grid, kQpThreads, 0, (cudaStream_t)stream>>>(rays_d, dq, R, eps, g_a2, g_b2);
    else return PAPR_ERR_INVALID_ARGUMENT;
    return check_launch();
}

extern "C" int papr_query_tail_fwd(const float *q5, const float *w_c, float c_const, float eps, int64_t R, void *z_blocked,
                                   float *stats, float *cprime, void *stream)
{
    if (!q5 || !w_c || !z_blocked || !stats || !cprime || R <= 0) return PAPR_ERR_INVALID_ARGUMENT;
    QueryTailParams p = {};
    p.q5 = q5; p.wc = w_c; p.c_const = c_const; p.eps = eps; p.R = R; p.z = (uint8_t *)z_blocked; p.stats = stats; p.cprime = cprime;
    query_tail_fwd_kernel<<<row_grid(R), kRowThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_query_tail_bwd(const float *q5, const float *stats, const float *w_c, const float *dz, int64_t ld_dz,
                                   const float *dc, float eps, int64_t R, float *dq5, float *g_wc, float *g_cconst, void *stream)
{
    if (!q5 || !stats || !w_c || !dz || !dc || !dq5 || !g_wc || !g_cconst || R <= 0 || ld_dz < 256 || ld_dz % 4) return PAPR_ERR_INVALID_ARGUMENT;
    QueryTailParams p = {};
    p.q5 = q5; p.stats = const_cast<float *>(stats); p.wc = w_c; p.dz = dz; p.ld_dz = ld_dz; p.dc = dc; p.eps = eps; p.R = R;
    p.dq5 = dq5; p.g_wc = g_wc; p.g_cc = g_cconst;
    const int grid_all = row_grid(R);           // 128 registers per thread: two CTAs per SM are resident
    query_tail_bwd_kernel<<<grid_all < 2 * kNumSMs ? grid_all : 2 * kNumSMs, kRowThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}
