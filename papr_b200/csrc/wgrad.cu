// Stage a12 building block: weight gradient of one Linear layer on the tensor cores,
//   C[a, b] += sum_m A[m, a] * B[m, b]            (dW = dZ^T X of autograd's Linear backward)
// A and B are tile-blocked bf16 activations ([rows, 64-column blocks]); because the reduction runs over ROWS, both
// are consumed as MN-major SWIZZLE_128B operands straight from the same 16 KB blocks the forward pass uses.
//
// Split-K over the grid: every CTA streams its share of 64-row half-tiles through a TMA/mbarrier ring and keeps the
// whole a_width x b_width fp32 product in TMEM (up to 2 x 256 columns = all 512) for the entire kernel; only at the
// end is it drained once and added atomically into the fp32 gradient.
#include "tc_common.cuh"
#include <stdlib.h>

namespace papr {

constexpr int kWgThreads = 256;
constexpr int kHalfBytes = kBlockBytes / 2;   // 64 rows x 128 B

struct WgradParams {
    const uint8_t *a;   // blocked bf16 [rows, a_blk*64]
    const uint8_t *b;   // blocked bf16 [rows, b_blk*64]
    float *c;           // fp32 [.., ldc], accumulated atomically
    float *a_colsum;    // [a_valid] += column sums of A over all rows (the bias gradient when A = dZ), or null
    int64_t n_units;    // half tiles
    int a_blk, b_blk, a_halves, a_used_blk, Nb, ldc, stages, transpose_out, a_valid, b_valid;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgradParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nb_used = (p.Nb + 63) >> 6;
    const int stage_bytes = (p.a_used_blk + nb_used) * kHalfBytes;
    uint8_t *ring = smem;
    uint64_t *bars = (uint64_t *)(ring + p.stages * stage_bytes);
    uint64_t *full = bars, *empty = bars + 8, *done = bars + 16;
    uint32_t *tmem_slot = (uint32_t *)(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        // a stage is free again once its MMAs have completed and -- with a_colsum -- the four summing warps have read it
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], p.a_colsum ? 5 : 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool has_work = (int64_t)blockIdx.x < p.n_units;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const int64_t tile = u >> 1;
                const size_t half_off = (size_t)(u & 1) * kHalfBytes;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                uint8_t *dst = ring + s * stage_bytes;
                for (int i = 0; i < p.a_used_blk; ++i)
                    bulk_g2s(dst + i * kHalfBytes, p.a + ((size_t)tile * p.a_blk + i) * kBlockBytes + half_off, kHalfBytes, &full[s]);
                for (int i = 0; i < nb_used; ++i)
                    bulk_g2s(dst + (p.a_used_blk + i) * kHalfBytes, p.b + ((size_t)tile * p.b_blk + i) * kBlockBytes + half_off, kHalfBytes, &full[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && has_work) {
            const uint32_t idesc = umma_idesc(128, p.Nb, true, true);
            int s = 0; uint32_t ph = 0; uint32_t first = 1;
            for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(ring + s * stage_bytes);
                const uint32_t b0 = a0 + p.a_used_blk * kHalfBytes;
                for (int ks = 0; ks < 4; ++ks) {
                    for (int h = 0; h < p.a_halves; ++h) {
                        const uint64_t ad = umma_desc(a0 + h * 2 * kHalfBytes + ks * 2048, kHalfBytes, 1024);
                        const uint64_t bd = umma_desc(b0 + ks * 2048, kHalfBytes, 1024);
                        umma_bf16(tmem_base + h * 256, ad, bd, idesc, (uint32_t)(!first || ks > 0));
                    }
                }
                first = 0;
                umma_commit(&empty[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4 && has_work) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        if (p.a_colsum) {
            // Bias gradient on the side: thread t owns columns 2t, 2t+1 of A and adds them up over the 64 rows of every
            // stage from the swizzled shared-memory image (a warp reads one 128-byte row per instruction: conflict-free).
            // The kernel is HBM-bound, these warps are otherwise idle until the drain.
            const int t = ew * 32 + lane;
            const bool live = 2 * t < p.a_used_blk * 64;
            const uint32_t blk_off = (uint32_t)(t >> 5) * kHalfBytes + (uint32_t)(lane & 3) * 4u;
            const uint32_t chunk = (uint32_t)(lane >> 2);
            float s0 = 0.f, s1 = 0.f;
            int s = 0; uint32_t ph = 0;
            for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                mbar_wait(&full[s], ph);
                if (live) {
                    const uint32_t base = smem_u32(ring + s * stage_bytes) + blk_off;
#pragma unroll 8
                    for (uint32_t r = 0; r < 64; ++r) {
                        uint32_t w2;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2) : "r"(base + r * 128u + ((chunk ^ (r & 7u)) << 4)));
                        s0 += bf16_lo(w2); s1 += bf16_hi(w2);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if (live && 2 * t < p.a_valid) atomicAdd(p.a_colsum + 2 * t, s0);
            if (live && 2 * t + 1 < p.a_valid) atomicAdd(p.a_colsum + 2 * t + 1, s1);
        }
        mbar_wait(done, 0);
        tc_fence_after();
        for (int h = 0; h < p.a_halves; ++h) {
            const int ai = h * 128 + row;
            for (int col0 = 0; col0 < p.Nb; col0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + h * 256 + col0, v);
                tmem_ld_wait();
                if (ai < p.a_valid) {
                    float *rowp = p.c + (size_t)ai * p.ldc + col0;
                    if (!p.transpose_out && col0 + 32 <= p.b_valid && (((uintptr_t)rowp) & 15) == 0) {
                        // 32 consecutive, 16-byte aligned floats of one gradient row: eight vector reductions
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_v4(rowp + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int bi = col0 + j;
                            if (bi < p.b_valid) {
                                float *dst = p.transpose_out ? p.c + (size_t)bi * p.ldc + ai : p.c + (size_t)ai * p.ldc + bi;
                                atomicAdd(dst, __uint_as_float(v[j]));
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace papr

extern "C" int papr_wgrad_bf16(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                               int a_valid, int b_valid, int transpose_out, int64_t rows, void *stream)
{
    return papr_wgrad_bf16_ex(a_blocked, a_cols, b_blocked, b_cols, c, ldc, a_valid, b_valid, transpose_out, rows, 0, stream);
}

extern "C" int papr_wgrad_bf16_ex(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                                  int a_valid, int b_valid, int transpose_out, int64_t rows, int max_ctas, void *stream)
{
    return papr_wgrad_bias_bf16(a_blocked, a_cols, b_blocked, b_cols, c, ldc, a_valid, b_valid, transpose_out, rows, max_ctas,
                                nullptr, stream);
}

extern "C" int papr_wgrad_bias_bf16(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                                    int a_valid, int b_valid, int transpose_out, int64_t rows, int max_ctas, float *a_colsum,
                                    void *stream)
{
    using namespace papr;
    if (!a_blocked || !b_blocked || !c) return PAPR_ERR_INVALID_ARGUMENT;
    if (rows <= 0 || rows % kTileRows || a_cols % 64 || b_cols % 64 || a_cols <= 0 || b_cols <= 0) return PAPR_ERR_INVALID_ARGUMENT;
    if (a_valid < 1 || a_valid > a_cols || a_valid > 256 || b_valid < 1 || b_valid > b_cols || b_valid > 256) return PAPR_ERR_INVALID_ARGUMENT;
    WgradParams p;
    p.a = (const uint8_t *)a_blocked; p.b = (const uint8_t *)b_blocked; p.c = c; p.a_colsum = a_colsum;
    p.n_units = rows / 64;
    p.a_blk = a_cols / 64; p.b_blk = b_cols / 64;
    p.a_halves = (a_valid + 127) / 128;                 // M = 128 per MMA: 1 or 2 row-halves of the product
    p.a_used_blk = p.a_halves * 2;
    if (p.a_used_blk > p.a_blk) return PAPR_ERR_INVALID_ARGUMENT;   // the a operand must be padded to a multiple of 128 columns
    p.Nb = (b_valid + 15) & ~15;
    p.ldc = (int)ldc; p.transpose_out = transpose_out; p.a_valid = a_valid; p.b_valid = b_valid;
    const int stage_bytes = (p.a_used_blk + (p.Nb + 63) / 64) * kHalfBytes;
    p.stages = (232448 - 1024 - 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    const int smem = 1024 + p.stages * stage_bytes + 1024;
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, wgrad_kernel, 232448));
    int grid = (int)(p.n_units < kNumSMs ? p.n_units : kNumSMs);
    { static int cap = -1; if (cap < 0) { const char *e = getenv("PAPR_DBG_WGRAD_GRID"); cap = e ? atoi(e) : 0; } if (cap > 0 && grid > cap) grid = cap; }
    if (max_ctas >= 1 && grid > max_ctas) grid = max_ctas;
    wgrad_kernel<<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(p);
    return check_launch();
}
