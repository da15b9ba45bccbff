// Stages a7/a12, fused (device body shared by stack.cu and stack_bwd.cu): a whole MLP stack (up to 8 Linear layers, forward or dgrad direction) per launch, with the
// 128-row activation tiles kept in shared memory from layer to layer (reference models/mlp.py:47-59 loop).
//
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one tcgen05.mma covers M = 256 rows -- 128 from each CTA -- and each
// CTA stages only HALF of every weight chunk (N/2 rows), so the weight stream from L2 costs 64 KB per tile-layer
// instead of 128 KB and the tensor pipe, not L2, becomes the limit.  Each CTA keeps two tiles in flight ("slots"):
// while the tensor core multiplies slot 1, the eight epilogue warps drain slot 0 (TMEM -> bias/activation or
// activation-derivative mask -> bf16 -> back into the slot's shared-memory tile, which is the next layer's A operand).
//
//   warp 0      producer : input tiles (TMA bulk, per slot) and this CTA's half of every weight chunk (ring of stages)
//   warp 1      MMA issuer (leader CTA only): the whole warp walks the job list so addresses stay in uniform registers,
//               one elected lane issues tcgen05.mma.cta_group::2 and commits with multicast, so the "stage free" /
//               "accumulator full" barriers fire in both CTAs
//   warp 2      TMEM allocator (both CTAs, cta_group::2), then the stash writer: TMA bulk stores of finished tiles,
//               metered so that weight loads never queue behind more than two 16 KB stores
//   warp 3      relay (peer CTA only): forwards "my TMA data landed" to the leader's barriers
//   warps 4-19  epilogue: warp (g, quad) drains 64-column group g of the current job for TMEM lane quadrant quad
// Training-time by-products are written as they appear: each layer's output tile (the next layer's input, needed by
// the weight-gradient kernel) straight from the shared-memory tile with TMA bulk stores, activation sign bits, and
// per-column sums (bias gradients) in the dgrad direction.
//
// Measured on B200 (tools/ubench_tmem.cu, tools/stack_power.py; profiles/r01_stack_kernel_study.md): the tensor pipe
// needs 2,050 cycles per job (256 rows x 256 x 256) and neither tcgen05.ld, st.shared nor TMA traffic slows it; what
// did were (a) per-instruction register->uniform broadcast loops around tcgen05.mma issued under `lane == 0` (fixed by
// elect.sync), (b) an epilogue of ~8 instructions per element (now ~2.5: bias add, one funnel shift for the sign bit,
// cvt.rn.relu.bf16x2 for activation + packing; the dgrad mask is applied to packed pairs with prmt), and (c) bulk stores
// of the stash starving the weight loads.  A sustained launch is power-capped (~990 W) on this part.
#pragma once
#include "tc_common.cuh"
#include <stdlib.h>

namespace papr {


constexpr int kStkThreads = 640;     // 4 control warps + 2 x 8 epilogue warps
constexpr int kStkMaxLayers = 8;
constexpr int kSlotBytes = 4 * kBlockBytes;      // one 128 x 256 bf16 tile
constexpr int kStkMaxSmem = 232448;
constexpr int kStageBlocks = 1;                  // 64-wide K blocks per weight-ring stage (4 MMAs each)
constexpr int kStageBytes = kStageBlocks * kBlockBytes;

constexpr uint32_t kModeOutBlocked = 1, kModeOutF32 = 2, kModeBitsOut = 4, kModeBitsIn = 8, kModeColsum = 16, kModeAct = 32, kModeBias = 64;

struct StackLayerDev {
    const uint8_t *w;          // weight image [kblk][N rows][128 B]
    const float *bias;         // [N] or null
    uint8_t *out_blocked;      // this layer's output as a tile-blocked tensor (stash / final), or null
    float *out_f32;            // fp32 row-major output, or null
    uint64_t *bits_out;        // sign bits of the pre-activation, or null
    const uint64_t *bits_in;   // activation-derivative mask bits (dgrad), or null
    float *colsum;             // [N] += column sums of the bf16 output, or null
    int64_t ld_f32;
    int64_t w_rep_stride;      // byte distance between identical copies of the weight image
    int N, kblk, k_steps, act, w_reps;
    uint32_t mode;             // kMode* flags of the fields above | N << 16: all the epilogue loop reads per job
    int ring;                  // 1: the stash of this layer goes to this CTA's slot ring (see StackRing), not to out_blocked
};

// Hand-over of stashed tiles to consumer CTAs of the same launch (stack_bwd.cu) through L2: every producer CTA owns
// `slots` 64 KB slots; its n-th ring stash (n counted per CTA over all ring layers and both tile slots) goes to slot
// n % slots once the consumer of item n - slots has released it, and is published by full[cta*slots + slot] = n + 1.
struct StackRing {
    uint8_t *base;             // [ctas][slots][4 x 16 KB]
    uint32_t *full, *freed;    // [ctas][slots], zeroed before the launch
    int slots;
};

struct StackParams {
    const uint8_t *x;          // tile-blocked input [rows, kblk0*64]
    int64_t n_tiles;
    int n_layers, kblk0, stages, any_stash, any_bits_in, store_depth, dbg_ring;
    int share_w;               // 1: one weight stream per layer, read by the jobs of both slots (a stage is released by slot 1)
    float slope;
    long long *trace;          // debug: per-job clock stamps of cluster 0 (null in production)
    StackRing rg;              // used by layers with ring == 1
    StackLayerDev L[kStkMaxLayers];
};


// ---- flags in global memory shared by the CTAs of one launch (all resident at once: grid <= number of SMs)
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all proxies, all state spaces: orders this thread's generic-proxy accesses with its async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Spin until *flag >= want.  A peer that never shows up (CTAs not co-resident, a bug) must not hang the GPU: trap after ~seconds.
__device__ __forceinline__ void flag_wait_ge(const uint32_t *flag, uint32_t want)
{
    uint32_t spins = 0;
    while (ld_acquire_gpu(flag) < want) {
        __nanosleep(40);
        if (++spins > (1u << 26)) __trap();
    }
}
template <int N> struct BulkWaitN { static __device__ __forceinline__ void run() { bulk_wait<N>(); } };
__device__ __forceinline__ void bulk_wait_keep(int newest)    // all but the `newest` most recent bulk groups are complete
{
    switch (newest) {
    case 0: bulk_wait<0>(); break;
    case 1: bulk_wait<1>(); break;
    case 2: bulk_wait<2>(); break;
    case 3: bulk_wait<3>(); break;
    default: bulk_wait<4>(); break;
    }
}

// The CTA-pair body; `cluster_id` of `n_clusters` pairs walk the tile quads (the launch may hold other CTAs besides).
template <bool RELU, bool TRACE = false>
__device__ __forceinline__ void stack_body(const StackParams &p, const int64_t cluster_id, const int64_t n_clusters)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *act = smem;                                   // 2 slots x 64 KB
    uint8_t *ring = act + 2 * kSlotBytes;                  // stages x 32 KB (this CTA's half of two 64-wide weight chunks)
    uint64_t *bars = (uint64_t *)(ring + p.stages * kStageBytes);
    uint64_t *w_full = bars, *w_empty = bars + 8, *pw_full = bars + 16;
    uint64_t *in_full = bars + 24, *pin_full = bars + 26, *in_free = bars + 28, *act_ready = bars + 30, *acc_full = bars + 32;
    uint64_t *st_ready = bars + 36, *st_done = bars + 38;
    uint32_t *tmem_slot = (uint32_t *)(bars + 34);
    float *vec_s = (float *)(bars + 40);                   // [layers][256]: biases (forward) or column sums (dgrad)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
        const int64_t n_quads = (p.n_tiles + 3) >> 2;
    const int L = p.n_layers;

    // The stash writer takes part in releasing a slot only when it has to read the LAST layer's tile out of it.  (It must
    // not arrive otherwise: with nothing to wait for it would run whole quads ahead of the epilogue warps and its early
    // arrivals would complete in_free phases that the epilogue has not reached -- a hang at scale, inference only.)
    const bool writer_on_last = p.L[L - 1].out_blocked != nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&pw_full[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&in_full[i], 1); mbar_init(&pin_full[i], 1); mbar_init(&in_free[i], writer_on_last ? 17 : 16);
            mbar_init(&st_ready[i], 16); mbar_init(&st_done[i], 1);
            mbar_init(&act_ready[i], 32); mbar_init(&acc_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(tmem_slot, 512);
    for (int i = threadIdx.x; i < L * 256; i += kStkThreads) {
        const int l = i >> 8, c = i & 255;
        vec_s[i] = (p.L[l].bias && c < p.L[l].N) ? p.L[l].bias[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (TRACE && p.trace && blockIdx.x == 0 && threadIdx.x == 0) {        // SM clock of the launch = cycles / nanoseconds
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.trace[4000] = clock64(); p.trace[4001] = (long long)ns;
    }

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0; int64_t qi = 0;
            for (int64_t q = cluster_id; q < n_quads; q += n_clusters, ++qi) {
                for (int l = 0; l < L; ++l) {
                    const StackLayerDev &Ld = p.L[l];
                    const uint32_t chunk_bytes = (uint32_t)(Ld.N >> 1) * 128u;
                    for (int s = 0; s < 2; ++s) {
                        if (l == 0) {
                            if (qi > 0) mbar_wait(&in_free[s], (uint32_t)((qi - 1) & 1));
                            int64_t tile = 4 * q + 2 * s + rank;
                            if (tile >= p.n_tiles) tile = p.n_tiles - 1;
                            mbar_arrive_expect_tx(&in_full[s], (uint32_t)(p.kblk0 * kBlockBytes));
                            for (int kb = 0; kb < p.kblk0; ++kb)
                                bulk_g2s(act + s * kSlotBytes + kb * kBlockBytes, p.x + ((size_t)tile * p.kblk0 + kb) * kBlockBytes,
                                         kBlockBytes, &in_full[s]);
                        }
                        const uint8_t *wsrc = Ld.w + (size_t)(cluster_id % Ld.w_reps) * Ld.w_rep_stride + (size_t)rank * chunk_bytes;
                        if (p.share_w && s == 1) continue;
                        for (int kc = 0; kc < Ld.kblk; kc += kStageBlocks) {
                            const int nc = min(kStageBlocks, Ld.kblk - kc);
                            mbar_wait(&w_empty[st], ph ^ 1);
                            mbar_arrive_expect_tx(&w_full[st], chunk_bytes * nc);
                            for (int c = 0; c < nc; ++c)
                                bulk_g2s(ring + st * kStageBytes + c * kBlockBytes, wsrc + (size_t)(kc + c) * Ld.N * 128, chunk_bytes, &w_full[st]);
                            if (++st == p.stages) { st = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // Stash writer: one thread moves finished activation tiles (layer outputs kept for the weight-gradient kernel) to
        // HBM with TMA bulk stores, at most `store_depth` in flight.  Issuing all four 16 KB blocks of a job at once (from
        // the epilogue groups) puts up to 2,000 cycles of store work ahead of the next weight load in the TMA unit's
        // queue and the tensor pipe starves; metered, a weight load waits behind one or two blocks at most.
        if (lane == 0) {
            uint32_t sj = 0;
            uint32_t rn = 0;                    // ring items of this CTA so far (valid or not: the consumers count the same way)
            uint32_t *pend_flag = nullptr; uint32_t pend_val = 0;      // ring item whose stores may still be in flight
            const int RS = p.rg.slots;
            for (int64_t q = cluster_id; q < n_quads; q += n_clusters) {
                for (int l = 0; l < L; ++l) {
                    uint8_t *out = p.L[l].out_blocked;
                    const bool stash = out != nullptr;
                    const bool to_ring = p.L[l].ring != 0;
                    const int ng = (p.L[l].N + 63) >> 6;
                    for (int s = 0; s < 2; ++s) {
                        if (stash) {
                            mbar_wait(&st_ready[s], sj & 1);
                            const int64_t tile = 4 * q + 2 * s + rank;
                            int issued = 0;
                            uint32_t *my_flag = nullptr;
                            if (tile < p.n_tiles) {
                                const int64_t dtile = p.dbg_ring ? tile % p.dbg_ring : tile;
                                uint8_t *dst = out + (size_t)dtile * ng * kBlockBytes;
                                if (to_ring) {
                                    const size_t slot = (size_t)blockIdx.x * RS + rn % RS;
                                    if (rn >= (uint32_t)RS) { flag_wait_ge(p.rg.freed + slot, rn - RS + 1); fence_proxy_async_all(); }
                                    dst = p.rg.base + slot * kSlotBytes;
                                    my_flag = p.rg.full + slot;
                                }
                                for (int g = 0; g < ng; ++g) {
                                    bulk_s2g(dst + (size_t)g * kBlockBytes, act + s * kSlotBytes + g * kBlockBytes, kBlockBytes);
                                    bulk_commit();
                                    if (p.store_depth > 1) bulk_wait_read<1>(); else bulk_wait_read<0>();
                                }
                                bulk_wait_read<0>();
                                issued = ng;
                            }
                            mbar_arrive(&st_done[s]);
                            if (pend_flag) {        // the previous ring item had a whole stash period to land: publish it
                                bulk_wait_keep(issued);
                                fence_proxy_async_all();
                                __threadfence();
                                st_release_gpu(pend_flag, pend_val);
                                pend_flag = nullptr;
                            }
                            if (my_flag) { pend_flag = my_flag; pend_val = rn + 1; }
                            if (to_ring) ++rn;
                        }
                        if (l == L - 1 && stash) mbar_arrive(&in_free[s]);      // the producer may refill the slot
                    }
                    if (stash) ++sj;
                }
            }
            bulk_wait<0>();
            if (pend_flag) { fence_proxy_async_all(); __threadfence(); st_release_gpu(pend_flag, pend_val); }
        }
    } else if (warp == 3) {
        if (lane == 0 && rank == 1) {       // relay: tell the leader that this CTA's TMA data has landed
            int st = 0; uint32_t ph = 0; int64_t qi = 0;
            for (int64_t q = cluster_id; q < n_quads; q += n_clusters, ++qi) {
                for (int l = 0; l < L; ++l) {
                    for (int s = 0; s < 2; ++s) {
                        if (l == 0) { mbar_wait(&in_full[s], (uint32_t)(qi & 1)); mbar_arrive_remote(&pin_full[s], 0); }
                        if (p.share_w && s == 1) continue;
                        for (int kc = 0; kc < p.L[l].kblk; kc += kStageBlocks) {
                            mbar_wait(&w_full[st], ph);
                            mbar_arrive_remote(&pw_full[st], 0);
                            if (++st == p.stages) { st = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // The whole warp walks the job list so that every address and descriptor stays warp-uniform (uniform registers,
        // no per-instruction register->uniform broadcast loops); lane 0 alone issues the MMAs and their commits.
        if (rank == 0) {
            int st = 0; uint32_t ph = 0; int64_t qi = 0;
            uint32_t jn = 0;
            const uint32_t a_base = smem_u32(act), ring_base = smem_u32(ring);
            for (int64_t q = cluster_id; q < n_quads; q += n_clusters, ++qi) {
                for (int l = 0; l < L; ++l, ++jn) {
                    const int kblk = p.L[l].kblk, k_steps = p.L[l].k_steps;
                    const uint32_t idesc = umma_idesc(256, p.L[l].N, false, false);
                    const int st_l = st; const uint32_t ph_l = ph;        // share_w: slot 1 walks the layer's weight stages again
#pragma unroll 1
                    for (int s = 0; s < 2; ++s) {
                        if (p.share_w) { st = st_l; ph = ph_l; }
                        if (l == 0) {
                            mbar_wait(&in_full[s], (uint32_t)(qi & 1));
                            mbar_wait_cluster(&pin_full[s], (uint32_t)(qi & 1));
                        }
                        const bool tr_m = TRACE && p.trace && blockIdx.x == 0 && qi < 4 && lane == 0;
                        long long t_w = 0, t_a = tr_m ? clock64() : 0;
                        if (jn > 0) mbar_wait_cluster(&act_ready[s], (jn - 1) & 1);
                        tc_fence_after();
                        if (tr_m) { p.trace[((qi * 16 + l) * 2 + s) * 8 + 0] = clock64(); p.trace[2048 + ((qi * 16 + l) * 2 + s) * 2] = clock64() - t_a; }
                        const uint32_t d = tmem_base + s * 256;
                        const uint64_t a_desc0 = umma_desc(a_base + s * kSlotBytes, 16, 1024);
                        uint32_t acc = 0;
#pragma unroll 1
                        for (int kc = 0; kc < kblk; kc += kStageBlocks) {
                            const long long t_w0 = tr_m ? clock64() : 0;
                            if (!p.share_w || s == 0) {
                                mbar_wait(&w_full[st], ph);
                                mbar_wait_cluster(&pw_full[st], ph);
                                tc_fence_after();
                            }
                            if (tr_m) t_w += clock64() - t_w0;
                            // a K step of 16 bf16 = 32 B = +2 in the descriptor's address field; a 64-wide block = 16 KB
                            const uint64_t ad = a_desc0 + (uint64_t)(kc * (kBlockBytes >> 4));
                            const uint64_t bd = umma_desc(ring_base + st * kStageBytes, 16, 1024);
                            const int nk = min(4 * kStageBlocks, k_steps - 4 * kc);
                            if (elect_one()) {
                                if (nk == 4 * kStageBlocks) {
#pragma unroll
                                    for (int k = 0; k < 4 * kStageBlocks; ++k) {
                                        const uint32_t off = (uint32_t)((k >> 2) * (kBlockBytes >> 4) + 2 * (k & 3));
                                        umma2_bf16(d, ad + off, bd + off, idesc, k ? 1u : acc);
                                    }
                                } else {
                                    for (int k = 0; k < nk; ++k) {
                                        const uint32_t off = (uint32_t)((k >> 2) * (kBlockBytes >> 4) + 2 * (k & 3));
                                        umma2_bf16(d, ad + off, bd + off, idesc, k ? 1u : acc);
                                    }
                                }
                                if (!p.share_w || s == 1) umma2_commit(&w_empty[st]);
                            }
                            __syncwarp();
                            acc = 1;
                            if (++st == p.stages) { st = 0; ph ^= 1; }
                        }
                        if (elect_one()) umma2_commit(&acc_full[s]);
                        __syncwarp();
                        if (tr_m) { p.trace[((qi * 16 + l) * 2 + s) * 8 + 1] = clock64(); p.trace[2048 + ((qi * 16 + l) * 2 + s) * 2 + 1] = t_w; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // Sixteen epilogue warps drain the jobs in issue order; warp (g, quad) owns 64-column group g of the tile for
        // the TMEM lane quadrant quad (== warp % 4), so one job is four independent 128-thread groups.
        // The loop around the drain is kept lean on purpose: at 16 warps on 4 schedulers every instruction per job costs four
        // issue slots, and the epilogue (ncu source view, r02) was issue-bound at ~520 instructions per warp and job of
        // which 180 were the drain itself -- per-layer facts come as one packed word (StackLayerDev::mode), barrier and
        // tile addresses are computed once, indices stay 32-bit, the trace stamps are compiled out of the production kernel.
        const int ew = warp - 4;
        const int g = ew >> 2, quad = ew & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int bar_id = 1 + g;
        const uint32_t blk0_s = smem_u32(act) + (uint32_t)g * kBlockBytes;     // my column block of slot 0 (shared window)
        const uint32_t row_s = (uint32_t)row * 128u, sw = (uint32_t)row & 7u;
        const uint32_t vec0_s = smem_u32(vec_s) + (uint32_t)g * 256u;          // my 64 bias / column-sum entries of layer 0
        const uint32_t acc_full_a = smem_u32(acc_full), st_done_a = smem_u32(st_done), st_ready_a = smem_u32(st_ready);
        const uint32_t in_free_a = smem_u32(in_free);
        const uint32_t act_ready_c = mapa_u32(smem_u32(act_ready), 0);         // the LEADER's barrier, cluster window
        const uint32_t taddr0 = tmem_base + lane_base + (uint32_t)g * 64u;
        const uint32_t n_tiles32 = (uint32_t)p.n_tiles;                        // rows < 2^31 * 128 is checked by the host
        uint32_t jn = 0;                 // jobs finished per slot (both slots advance in lockstep)
        uint32_t sj = 0;                 // stash jobs handed to the writer thread so far, per slot

        // activation-derivative bits of the job after the current one are fetched one job ahead (dgrad launches only)
        auto fetch_bits = [&](uint32_t tile, int l2) -> uint64_t {
            const uint32_t m2 = p.L[l2].mode;
            const uint32_t ng = ((m2 >> 16) + 63u) >> 6;
            if (!(m2 & kModeBitsIn) || tile >= n_tiles32 || (uint32_t)g >= ng) return 0;
            return __ldg(p.L[l2].bits_in + (size_t)((tile * ng + (uint32_t)g) * (uint32_t)kTileRows + (uint32_t)row));   // [tile][group][row]
        };
        uint64_t din_next = p.any_bits_in ? fetch_bits((uint32_t)(4 * cluster_id) + rank, 0) : 0;

        for (int64_t q = cluster_id; q < n_quads; q += n_clusters) {
            const uint32_t tile0 = (uint32_t)(4 * q) + rank;                  // slot 0's tile; slot 1's is tile0 + 2
            const int64_t tq = TRACE ? (q - cluster_id) / n_clusters : 0;
            for (int l = 0; l < L; ++l, ++jn) {
                const uint32_t mode = p.L[l].mode;
                const int N = (int)(mode >> 16);
                const int ngroups = (N + 63) >> 6;
                const bool last = l == L - 1;
                const bool stash = (mode & kModeOutBlocked) != 0;
                const bool to_act = !last || stash;
                const bool mine = g < ngroups;
                const bool fast = to_act && !(mode & kModeOutF32) && (g + 1) * 64 <= N && (RELU || !(mode & kModeBitsIn));
                const uint32_t vec_sa = vec0_s + (uint32_t)l * 1024u;
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const uint32_t tile = tile0 + 2u * (uint32_t)s;
                    const bool valid = tile < n_tiles32;
                    const bool tr = TRACE && p.trace && blockIdx.x == 0 && tq < 4 && ew == 0 && lane == 0;
                    const uint64_t din = din_next;
                    if (p.any_bits_in) {
                        // the next job: slot 1 of this layer, or slot 0 of the next layer (of the next quad after the last)
                        const bool wrap = s == 1 && last;
                        din_next = fetch_bits(s == 0 ? tile0 + 2u : (wrap ? tile0 + 4u * (uint32_t)n_clusters : tile0),
                                              s == 0 ? l : (wrap ? 0 : l + 1));
                    }
                    mbar_wait_a(acc_full_a + 8u * (uint32_t)s, jn & 1);
                    tc_fence_after();
                    if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 2] = clock64();
                    if (mine) {
                        const uint32_t blk_s = blk0_s + (uint32_t)s * kSlotBytes;
                        const uint32_t taddr = taddr0 + (uint32_t)s * 256u;
                        if (sj > 0 && to_act) mbar_wait_a(st_done_a + 8u * (uint32_t)s, (sj - 1) & 1);   // the writer has read this slot's last stashed tile
                        if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 7] = clock64();
                        uint64_t dout = 0;
                        if (fast) {
                            // hidden layer, all 64 columns live: TMEM -> math -> bf16 -> this row's eight 16-byte chunks
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                uint32_t v[32], w[16];
                                tmem_ld32(taddr + h * 32, v);
                                tmem_ld_wait();
                                if (mode & kModeBitsIn) {
                                    mask_pack_relu32(v, (uint32_t)(din >> (32 * h)), w);
                                } else if (RELU && (mode & kModeAct)) {
                                    uint32_t bits;
                                    if (mode & kModeBitsOut) bits = bias_relu_pack32<true>(v, vec_sa + h * 128, w);
                                    else bits = bias_relu_pack32<false>(v, vec_sa + h * 128, w);
                                    dout |= (uint64_t)bits << (32 * h);
                                } else {
                                    uint32_t bits = 0;
                                    if (mode & kModeBitsOut) epilogue_math32<EPI_BIAS_ACT_BITS, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                    else if (mode & kModeAct) epilogue_math32<EPI_BIAS_ACT, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                    else if (mode & kModeBias) epilogue_math32<EPI_BIAS, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                    dout |= (uint64_t)bits << (32 * h);
#pragma unroll
                                    for (int k = 0; k < 16; ++k) w[k] = pack_bf16(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                                }
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    sts_v4(blk_s + row_s + ((((uint32_t)(h * 4 + c)) ^ sw) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                            }
                        } else {
                            const StackLayerDev &Ld = p.L[l];
#pragma unroll 1
                            for (int h = 0; h < 2; ++h) {
                                const int col0 = g * 64 + h * 32;
                                if (col0 >= N) {
                                    if (to_act) {
#pragma unroll
                                        for (int c = 0; c < 4; ++c) sts_v4(blk_s + row_s + ((((uint32_t)(h * 4 + c)) ^ sw) << 4), 0, 0, 0, 0);
                                    }
                                    continue;
                                }
                                uint32_t v[32];
                                tmem_ld32(taddr + h * 32, v);
                                tmem_ld_wait();
                                uint32_t bits = 0;
                                const uint32_t dh = (uint32_t)(din >> (32 * h));
                                if (mode & kModeBitsIn) epilogue_math32<EPI_MASK, RELU>(v, vec_sa + h * 128, p.slope, dh, bits);
                                else if (mode & kModeBitsOut) epilogue_math32<EPI_BIAS_ACT_BITS, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                else if (mode & kModeAct) epilogue_math32<EPI_BIAS_ACT, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                else if (mode & kModeBias) epilogue_math32<EPI_BIAS, RELU>(v, vec_sa + h * 128, p.slope, 0, bits);
                                dout |= (uint64_t)bits << (32 * h);
                                if ((mode & kModeOutF32) && valid) {
                                    const int64_t grow = (int64_t)tile * kTileRows + row;
                                    float4 *dst = reinterpret_cast<float4 *>(Ld.out_f32 + grow * Ld.ld_f32 + col0);
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                                }
                                if (to_act) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        sts_v4(blk_s + row_s + ((((uint32_t)(h * 4 + c)) ^ sw) << 4),
                                               pack_bf16(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1])),
                                               pack_bf16(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                                               pack_bf16(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])),
                                               pack_bf16(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
                                }
                            }
                        }
                        if ((mode & kModeBitsOut) && valid)
                            p.L[l].bits_out[(size_t)((tile * (uint32_t)ngroups + (uint32_t)g) * (uint32_t)kTileRows + (uint32_t)row)] = dout;
                        if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 5] = clock64();
                        if (to_act) fence_proxy_async();      // generic-proxy tile writes -> visible to tcgen05.mma / TMA store
                        if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 6] = clock64();
                        if (to_act && (mode & kModeColsum)) {
                            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                            if (valid) {
                                // thread (qq, ll) adds up columns 2*ll, 2*ll+1 over rows [32*qq, 32*qq+32) of the bf16 tile
                                float *vec = vec_s + l * 256;
                                const uint32_t qq = (uint32_t)row >> 5, ll = (uint32_t)row & 31u;
                                const uint32_t cbase = blk_s + qq * 4096u + (ll & 3u) * 4u;
                                float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                                for (uint32_t r = 0; r < 32; ++r) {
                                    uint32_t w2;
                                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2) : "r"(cbase + r * 128u + (((ll >> 2) ^ (r & 7u)) << 4)));
                                    s0 += bf16_lo(w2); s1 += bf16_hi(w2);
                                }
                                atomicAdd(&vec[g * 64 + 2 * ll], s0);
                                atomicAdd(&vec[g * 64 + 2 * ll + 1], s1);
                            }
                        }
                    }
                    if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 3] = clock64();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (stash) mbar_arrive_a(st_ready_a + 8u * (uint32_t)s);
                        if (last) mbar_arrive_a(in_free_a + 8u * (uint32_t)s);
                        mbar_arrive_cluster_a(act_ready_c + 8u * (uint32_t)s);
                    }
                    if (tr) p.trace[((tq * 16 + l) * 2 + s) * 8 + 4] = clock64();
                }
                sj += stash ? 1u : 0u;
            }
        }
        asm volatile("bar.sync 5, 512;" ::: "memory");
        for (int i = threadIdx.x - 128; i < L * 256; i += 512) {
            const int l = i >> 8, c = i & 255;
            if (p.L[l].colsum && c < p.L[l].N) atomicAdd(p.L[l].colsum + c, vec_s[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (TRACE && p.trace && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        if (blockIdx.x == 0) { p.trace[4002] = clock64(); p.trace[4003] = (long long)ns; }
        if (rank == 0 && cluster_id < 80) p.trace[3900 + cluster_id] = (long long)ns;     // when each CTA pair ran out of work
    }
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

// Host side: validates the layer list of papr_stack_bf16 and fills the device parameters.  Returns a papr status.
static inline int fill_stack_params(StackParams &p, const void *x, int K0, const papr_stack_layer *layers, int n_layers, int64_t rows,
                                    float slope, int *smem_bytes)
{
    if (!x || !layers || n_layers < 1 || n_layers > kStkMaxLayers) return PAPR_ERR_INVALID_ARGUMENT;
    if (rows <= 0 || rows % kTileRows || rows / kTileRows >= (int64_t)1 << 22 || K0 < 16 || K0 > 256 || K0 % 16) return PAPR_ERR_INVALID_ARGUMENT;
    p.x = (const uint8_t *)x; p.n_tiles = rows / kTileRows; p.n_layers = n_layers; p.kblk0 = (K0 + 63) / 64; p.slope = slope;
    p.any_stash = 0;
    p.any_bits_in = 0;
    { static int depth = -1; if (depth < 0) { const char *e = getenv("PAPR_STACK_STORE_DEPTH"); depth = e ? atoi(e) : 2; } p.store_depth = depth; }
    p.trace = nullptr;
    p.dbg_ring = 0;
    { static int v = -1; if (v < 0) { const char *e = getenv("PAPR_STACK_SHARE_W"); v = e ? atoi(e) : 0; } p.share_w = v; }
    p.rg.base = nullptr; p.rg.full = nullptr; p.rg.freed = nullptr; p.rg.slots = 1;
    int K = K0;
    for (int l = 0; l < n_layers; ++l) {
        const papr_stack_layer &h = layers[l];
        const bool last = l == n_layers - 1;
        if (!h.w_image || h.N < 32 || h.N > 256 || h.N % 32) return PAPR_ERR_INVALID_ARGUMENT;
        if (!last && h.N != 256) return PAPR_ERR_INVALID_ARGUMENT;            // hidden activations fill one 128 x 256 slot
        if (last && !h.out_blocked && !h.out_f32) return PAPR_ERR_INVALID_ARGUMENT;
        if (h.out_f32 && (h.ld_f32 < h.N || h.ld_f32 % 4)) return PAPR_ERR_INVALID_ARGUMENT;
        if (h.sign_bits_in && (h.bias || h.act || h.sign_bits_out)) return PAPR_ERR_INVALID_ARGUMENT;
        if (h.sign_bits_out && !(h.bias && h.act)) return PAPR_ERR_INVALID_ARGUMENT;
        if (h.act && !h.bias) return PAPR_ERR_INVALID_ARGUMENT;
        if (h.colsum && h.bias) return PAPR_ERR_INVALID_ARGUMENT;
        StackLayerDev &d = p.L[l];
        d.w = (const uint8_t *)h.w_image; d.bias = h.bias; d.out_blocked = (uint8_t *)h.out_blocked; d.out_f32 = h.out_f32;
        d.bits_out = h.sign_bits_out; d.bits_in = h.sign_bits_in; d.colsum = h.colsum; d.ld_f32 = h.ld_f32;
        d.N = h.N; d.kblk = (K + 63) / 64; d.k_steps = K / 16; d.act = h.act; d.ring = 0;
        d.mode = (h.out_blocked ? kModeOutBlocked : 0) | (h.out_f32 ? kModeOutF32 : 0) | (h.sign_bits_out ? kModeBitsOut : 0) |
                 (h.sign_bits_in ? kModeBitsIn : 0) | (h.colsum ? kModeColsum : 0) | (h.act ? kModeAct : 0) | (h.bias ? kModeBias : 0) |
                 ((uint32_t)h.N << 16);
        if (h.sign_bits_in) p.any_bits_in = 1;
        d.w_reps = h.w_replicas > 1 ? h.w_replicas : 1; d.w_rep_stride = h.w_replica_stride;
        if (h.out_blocked) p.any_stash = 1;
        K = h.N;
    }
    const int fixed = 1024 + 2 * kSlotBytes + 512 + n_layers * 1024;
    p.stages = (kStkMaxSmem - fixed) / kStageBytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return PAPR_ERR_INVALID_ARGUMENT;
    *smem_bytes = fixed + p.stages * kStageBytes;
    return PAPR_OK;
}

}  // namespace papr
