// Stage a12, fused: the whole backward of one MLP stack (dgrad of every layer + every weight gradient) in ONE launch
// (autograd of the reference models/mlp.py:47-59 loop).
//
// Why: run as two kernels (papr_stack_bf16 in the dgrad direction, then papr_wgrad_bf16 per layer), the per-layer dZ tiles
// cross HBM twice -- written by the dgrad kernel, which that makes HBM-WRITE bound (~3.9 TB/s on B200), and read again by
// the weight-gradient kernel, which is HBM-read bound.  Here they never leave L2:
//
//   producer clusters (CTA pairs, the first 2 * n_prod CTAs) run the stack kernel body (stack_body.cuh) over the tile quads
//       exactly as papr_stack_bf16 does, but the stash of every hidden layer goes to a small per-CTA ring of 64 KB slots in
//       global memory and is published with a release flag once its TMA stores have completed;
//   consumer CTAs (the rest) each own ONE layer's weight gradient for the whole launch: 256 x 256 fp32 = all 512 TMEM
//       columns, accumulated over every tile handed to them (the same split-K pipeline as wgrad.cu: TMA ring, MN-major
//       tcgen05.mma, one atomic drain at the end).  A consumer walks the producer CTAs assigned to it in production order,
//       acquires the slot's flag, streams the dZ tile from the ring (L2) and the matching layer-input tile from the forward
//       stash (HBM), and releases the slot as soon as the tile has landed in its shared memory.
//
// Every CTA of the grid is resident at once (grid <= number of SMs, one CTA per SM by shared memory), which the flag waits
// rely on; a wait that never completes traps instead of hanging.  The last layer's dZ is the kernel's input and is read
// from HBM by both sides; the input gradient of the first layer is written to HBM as before.
#include "stack_body.cuh"

namespace papr {

constexpr int kHalfBytesF = kBlockBytes / 2;   // 64 rows x 128 B

struct WgradStreamLayer {
    const uint8_t *x;       // forward input of the layer, tile-blocked bf16 [rows, x_blk * 64]
    float *c;               // fp32 gradient [n_out, ldc] (or [n_in, ldc] when swapped = transposed store)
    int x_blk, ldc, n_out, n_in;
    int swap;               // 1: A = x, B = dZ (narrow outputs), result stored transposed
    int ring_step;          // dgrad step whose stash is this layer's dZ, or -1: dZ is the kernel input
};

struct BwdFusedParams {
    StackParams sp;
    WgradStreamLayer W[kStkMaxLayers];
    int n_prod;             // producer clusters
    int n_ring_steps;       // dgrad steps that stash to the ring (= n_layers - 1)
    int dz_blk;             // 64-column blocks of the kernel input
    uint8_t cons_layer[160], cons_idx[160], cons_cnt[kStkMaxLayers];
};

// One consumer CTA: weight gradient of layer `li`, the `j`-th of `cnt` consumers of that layer.
__device__ __forceinline__ void wgrad_stream_body(const BwdFusedParams &p, const int li, const int j, const int cnt)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const WgradStreamLayer &W = p.W[li];
    const bool from_ring = W.ring_step >= 0;
    const int a_valid = W.swap ? W.n_in : W.n_out, b_valid = W.swap ? W.n_out : W.n_in;
    const int a_halves = (a_valid + 127) >> 7, a_used = 2 * a_halves;
    const int Nb = (b_valid + 15) & ~15, nb_used = (Nb + 63) >> 6;
    const int stage_bytes = (a_used + nb_used) * kHalfBytesF;
    int stages = (kStkMaxSmem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    uint8_t *ring = smem;
    uint64_t *bars = (uint64_t *)(ring + stages * stage_bytes);
    uint64_t *full = bars, *empty = bars + 8, *done = bars + 16;
    uint32_t *tmem_slot = (uint32_t *)(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_tiles = p.sp.n_tiles, n_quads = (n_tiles + 3) >> 2;
    const int n_prod = p.n_prod, RS = p.sp.rg.slots, LR = p.n_ring_steps;
    // does this consumer get any tile at all?  (its first producer CTA, first iteration, either slot)
    bool has_work = false;
    for (int c = j; c < 2 * n_prod && !has_work; c += cnt) {
        const int64_t q = c >> 1;
        if (q < n_quads && 4 * q + (c & 1) < n_tiles) has_work = true;
    }

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t i = 0; i * n_prod < n_quads; ++i) {
                for (int c = j; c < 2 * n_prod; c += cnt) {
                    const int64_t q = i * n_prod + (c >> 1);
                    if (q >= n_quads) continue;
                    for (int sl = 0; sl < 2; ++sl) {
                        const int64_t tile = 4 * q + 2 * sl + (c & 1);
                        if (tile >= n_tiles) continue;
                        const uint8_t *dz;
                        if (from_ring) {
                            const uint32_t n = (uint32_t)((i * LR + W.ring_step) * 2 + sl);
                            const size_t slot = (size_t)c * RS + n % RS;
                            flag_wait_ge(p.sp.rg.full + slot, n + 1);
                            fence_proxy_async_all();
                            dz = p.sp.rg.base + slot * kSlotBytes;
                        } else {
                            dz = p.sp.x + (size_t)tile * p.dz_blk * kBlockBytes;
                        }
                        const uint8_t *xs = W.x + (size_t)tile * W.x_blk * kBlockBytes;
                        const uint8_t *a = W.swap ? xs : dz, *b = W.swap ? dz : xs;
                        for (int h = 0; h < 2; ++h) {
                            const size_t half_off = (size_t)h * kHalfBytesF;
                            mbar_wait(&empty[s], ph ^ 1);
                            mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                            uint8_t *dst = ring + s * stage_bytes;
                            for (int k = 0; k < a_used; ++k)
                                bulk_g2s(dst + k * kHalfBytesF, a + (size_t)k * kBlockBytes + half_off, kHalfBytesF, &full[s]);
                            for (int k = 0; k < nb_used; ++k)
                                bulk_g2s(dst + (a_used + k) * kHalfBytesF, b + (size_t)k * kBlockBytes + half_off, kHalfBytesF, &full[s]);
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && has_work) {
            const uint32_t idesc = umma_idesc(128, Nb, true, true);
            int s = 0; uint32_t ph = 0; uint32_t first = 1;
            for (int64_t i = 0; i * n_prod < n_quads; ++i) {
                for (int c = j; c < 2 * n_prod; c += cnt) {
                    const int64_t q = i * n_prod + (c >> 1);
                    if (q >= n_quads) continue;
                    for (int sl = 0; sl < 2; ++sl) {
                        const int64_t tile = 4 * q + 2 * sl + (c & 1);
                        if (tile >= n_tiles) continue;
                        for (int h = 0; h < 2; ++h) {
                            mbar_wait(&full[s], ph);
                            tc_fence_after();
                            if (h == 1 && from_ring) {      // both halves of the ring slot are in shared memory: hand it back
                                const uint32_t n = (uint32_t)((i * LR + W.ring_step) * 2 + sl);
                                st_release_gpu(p.sp.rg.freed + (size_t)c * RS + n % RS, n + 1);
                            }
                            const uint32_t a0 = smem_u32(ring + s * stage_bytes);
                            const uint32_t b0 = a0 + a_used * kHalfBytesF;
                            for (int ks = 0; ks < 4; ++ks) {
                                for (int hh = 0; hh < a_halves; ++hh) {
                                    const uint64_t ad = umma_desc(a0 + hh * 2 * kHalfBytesF + ks * 2048, kHalfBytesF, 1024);
                                    const uint64_t bd = umma_desc(b0 + ks * 2048, kHalfBytesF, 1024);
                                    umma_bf16(tmem_base + hh * 256, ad, bd, idesc, (uint32_t)(!first || ks > 0));
                                }
                            }
                            first = 0;
                            umma_commit(&empty[s]);
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
            umma_commit(done);
        }
    } else if (warp >= 4 && warp < 8 && has_work) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        for (int h = 0; h < a_halves; ++h) {
            const int ai = h * 128 + row;
            for (int col0 = 0; col0 < Nb; col0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + h * 256 + col0, v);
                tmem_ld_wait();
                if (ai < a_valid) {
                    float *rowp = W.c + (size_t)ai * W.ldc + col0;
                    if (!W.swap && col0 + 32 <= b_valid && (((uintptr_t)rowp) & 15) == 0) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            red_add_v4(rowp + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int bi = col0 + e;
                            if (bi < b_valid) {
                                float *dst = W.swap ? W.c + (size_t)bi * W.ldc + ai : W.c + (size_t)ai * W.ldc + bi;
                                atomicAdd(dst, __uint_as_float(v[e]));
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kStkThreads, 1) stack_bwd_kernel(const __grid_constant__ BwdFusedParams p)
{
    const int n_prod_ctas = 2 * p.n_prod;
    if ((int)blockIdx.x < n_prod_ctas) {
        stack_body<true>(p.sp, blockIdx.x >> 1, p.n_prod);
    } else {
        const int ci = (int)blockIdx.x - n_prod_ctas;
        wgrad_stream_body(p, p.cons_layer[ci], p.cons_idx[ci], p.cons_cnt[p.cons_layer[ci]]);
    }
}

constexpr int kRingSlots = 8;

// The flag hand-over needs every CTA of the launch resident at once: ask the driver how many CTA pairs fit (a GPC with an
// odd number of SMs cannot host a pair on its last one), once per device.
static int max_coresident_ctas()
{
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (!cached[dev]) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(kNumSMs & ~1)); cfg.blockDim = dim3(kStkThreads); cfg.dynamicSmemBytes = kStkMaxSmem;
        cudaLaunchAttribute at = {};
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, stack_bwd_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = kNumSMs / 2 - 4; }
        cached[dev] = 2 * n;
    }
    return cached[dev];
}

}  // namespace papr

extern "C" int64_t papr_stack_bwd_workspace_bytes(void)
{
    using namespace papr;
    return 16384 + (int64_t)kNumSMs * kRingSlots * kSlotBytes;      // flags + one ring of slots per (possible) producer CTA
}

extern "C" int papr_stack_bwd_fused(const void *dz, int K0, const papr_stack_layer *dgrad_layers, const papr_wgrad_layer *wgrad_layers,
                                    int n_layers, int64_t rows, int producer_ctas, void *workspace, int64_t workspace_bytes, void *stream)
{
    using namespace papr;
    if (!wgrad_layers || !workspace || n_layers < 2) return PAPR_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < papr_stack_bwd_workspace_bytes() || ((uintptr_t)workspace & 1023)) return PAPR_ERR_INVALID_ARGUMENT;
    BwdFusedParams p;
    int smem = 0;
    const int st = fill_stack_params(p.sp, dz, K0, dgrad_layers, n_layers, rows, 0.f, &smem);
    if (st != PAPR_OK) return st;
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, stack_bwd_kernel, kStkMaxSmem));
    int sms = kNumSMs;
    { const int fit = max_coresident_ctas(); if (fit > 0 && fit < sms) sms = fit; }
    if (sms > 160 || sms < 2 * (n_layers + 1)) return PAPR_ERR_INVALID_ARGUMENT;
    // dgrad steps 0 .. L-2 hand their output (the dZ of forward layer L-2-step) to the consumers through the ring
    uint8_t *ws = (uint8_t *)workspace;
    static_assert(sizeof(uint32_t) * 160 * kRingSlots <= 8192, "flag arrays");
    p.sp.rg.full = (uint32_t *)ws; p.sp.rg.freed = (uint32_t *)(ws + 8192);
    p.sp.rg.base = ws + 16384;
    p.sp.rg.slots = kRingSlots;
    for (int l = 0; l < n_layers; ++l) {
        if (dgrad_layers[l].out_f32 || dgrad_layers[l].bias || dgrad_layers[l].act) return PAPR_ERR_INVALID_ARGUMENT;
        if (l < n_layers - 1) { p.sp.L[l].ring = 1; p.sp.L[l].out_blocked = p.sp.rg.base; p.sp.L[l].mode |= kModeOutBlocked; }
        else if (!dgrad_layers[l].out_blocked) return PAPR_ERR_INVALID_ARGUMENT;
    }
    p.sp.any_stash = 1;
    p.n_ring_steps = n_layers - 1;
    p.dz_blk = (K0 + 63) / 64;
    double cost[kStkMaxLayers], total = 0;
    for (int li = 0; li < n_layers; ++li) {
        const papr_wgrad_layer &h = wgrad_layers[li];
        WgradStreamLayer &w = p.W[li];
        if (!h.x_blocked || !h.gw || h.x_cols % 64 || h.n_out < 1 || h.n_in < 1 || h.n_in > h.x_cols || h.n_in > 256 || h.n_out > 256) return PAPR_ERR_INVALID_ARGUMENT;
        w.x = (const uint8_t *)h.x_blocked; w.c = h.gw; w.x_blk = h.x_cols / 64; w.ldc = (int)h.ldw; w.n_out = h.n_out; w.n_in = h.n_in;
        w.ring_step = li < n_layers - 1 ? n_layers - 2 - li : -1;
        if (li < n_layers - 1 && h.n_out != 256) return PAPR_ERR_INVALID_ARGUMENT;
        if (li == n_layers - 1 && h.n_out > K0) return PAPR_ERR_INVALID_ARGUMENT;
        w.swap = h.n_out < 128 ? 1 : 0;
        const int a_valid = w.swap ? h.n_in : h.n_out, a_blk = w.swap ? w.x_blk : (w.ring_step >= 0 ? 4 : p.dz_blk);
        const int a_halves = (a_valid + 127) / 128;
        if (2 * a_halves > a_blk) return PAPR_ERR_INVALID_ARGUMENT;      // the A operand must be padded to a multiple of 128 columns
        const int b_valid = w.swap ? h.n_out : h.n_in;
        cost[li] = a_halves * (double)((b_valid + 15) & ~15) + 64.0;      // MMA columns per 64-row unit + a fixed part
        total += cost[li];
    }
    const int64_t n_quads = (p.sp.n_tiles + 3) / 4;
    int n_prod = producer_ctas > 0 ? producer_ctas / 2 : 44;
    if (n_prod > (sms - n_layers) / 2) n_prod = (sms - n_layers) / 2;
    if (n_prod > n_quads) n_prod = (int)n_quads;
    if (n_prod < 1) n_prod = 1;
    int n_cons = sms - 2 * n_prod;
    if (n_cons > 2 * n_prod * n_layers) n_cons = 2 * n_prod * n_layers;   // at most one consumer per (producer CTA, layer)
    n_cons &= ~1;                                                        // whole clusters
    if (n_cons < n_layers) return PAPR_ERR_INVALID_ARGUMENT;
    p.n_prod = n_prod;
    // consumers per layer proportional to the layer's MMA cost (largest remainder), at least one each
    int cnt[kStkMaxLayers], assigned = 0;
    for (int li = 0; li < n_layers; ++li) { cnt[li] = (int)(cost[li] / total * n_cons); if (cnt[li] < 1) cnt[li] = 1; assigned += cnt[li]; }
    while (assigned > n_cons) { int m = 0; for (int li = 1; li < n_layers; ++li) if (cnt[li] / cost[li] > cnt[m] / cost[m]) m = li; if (cnt[m] <= 1) break; --cnt[m]; --assigned; }
    while (assigned < n_cons) { int m = 0; for (int li = 1; li < n_layers; ++li) if (cnt[li] / cost[li] < cnt[m] / cost[m]) m = li; ++cnt[m]; ++assigned; }
    if (assigned != n_cons) return PAPR_ERR_INVALID_ARGUMENT;
    // interleave the layers over the consumer CTAs so that neighbouring SMs do not all read the same stash
    {
        int given[kStkMaxLayers] = {0}, ci = 0;
        while (ci < n_cons)
            for (int li = 0; li < n_layers && ci < n_cons; ++li)
                if (given[li] < cnt[li]) { p.cons_layer[ci] = (uint8_t)li; p.cons_idx[ci] = (uint8_t)given[li]++; ++ci; }
        for (int li = 0; li < n_layers; ++li) p.cons_cnt[li] = (uint8_t)cnt[li];   // (a consumer beyond 2*n_prod simply has no work)
    }
    PAPR_CUDA_TRY(cudaMemsetAsync(workspace, 0, 16384, (cudaStream_t)stream));
    stack_bwd_kernel<<<2 * n_prod + n_cons, kStkThreads, kStkMaxSmem, (cudaStream_t)stream>>>(p);
    return check_launch();
}
