// Thin PTX wrappers for the Blackwell (sm_100a) tensor path: mbarrier, TMA bulk copies, TMEM, tcgen05.mma.
// Also defines the activation layout every tensor-core kernel in this library shares.
#pragma once
#include "common.cuh"

namespace papr {

// ---------------------------------------------------------------------------------------------------------------
// "Tile-blocked" bf16 activation layout.  A logical [rows, cols] matrix (rows % 128 == 0, cols % 64 == 0) is stored
// as [rows/128][cols/64] blocks of 16 KB; a block holds 128 rows x 64 columns, one row per 128 bytes, with the
// eight 16-byte chunks of a row XOR-swizzled by (row & 7).  A block is byte-for-byte the shared-memory image that
// tcgen05.mma expects for a K-major SWIZZLE_128B operand (and, read the other way, for an MN-major one), so a
// single 1-D TMA bulk copy moves it between HBM and shared memory with no tensor map and no re-layout.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kBlockCols = 64;
constexpr int kBlockBytes = kTileRows * kBlockCols * 2;   // 16384

__host__ __device__ inline size_t blocked_chunk_offset(int64_t row, int chunk /* 16-byte chunk = col/8 */, int nblk)
{
    const int64_t tile = row >> 7;
    const int r = (int)(row & 127);
    const int blk = chunk >> 3, c = chunk & 7;
    return ((size_t)(tile * nblk + blk)) * kBlockBytes + (size_t)r * 128 + (size_t)((c ^ (r & 7)) << 4);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// the same on a precomputed shared-window address (hot loops: no generic -> shared conversion per call)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_saddr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITA_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEA_%=;\n\t"
        "bra WAITA_%=;\n\t"
        "DONEA_%=:\n\t}" ::"r"(bar_saddr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_saddr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
// cluster-window address of a shared-window address in CTA `cta` of the cluster (valid for the own CTA as well)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster_a(uint32_t bar_caddr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_caddr) : "memory");
}

// ------------------------------------------------------------------ TMA bulk copies (1-D)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (TMA store, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// one lane of the (fully active) warp; the compiler knows a single thread runs the guarded code, so register operands of
// tcgen05.mma / cp.async.bulk move to uniform registers without a per-instruction broadcast loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\t@P mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05.mma (bf16 x bf16 -> fp32, cta_group::1)
// Shared-memory matrix descriptor, SWIZZLE_128B (layout type 2 at bits [61,64)), descriptor version 1 at [46,48).
//   K-major operand : rows of 128 B (64 bf16 along K), 8-row groups SBO bytes apart; LBO unused (1).
//   MN-major operand: rows of 128 B (64 bf16 along M/N), 8 K-rows per 1 KB atom, next K group SBO bytes on,
//                     next 64-wide M/N group LBO bytes on.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (upper 32 bits of idescE): D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), majors at bits 15/16,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn_major, bool b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ------------------------------------------------------------------ cluster / 2-CTA (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// M = 256 across the CTA pair: each CTA supplies its 128 rows of A and half (N/2 rows) of B; issued by the leader only
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs of this thread completed
__device__ __forceinline__ void umma2_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// Epilogue flavours (compile-time so the per-element code carries no dead predicates)
enum { EPI_PLAIN = 0, EPI_BIAS = 1, EPI_BIAS_ACT = 2, EPI_BIAS_ACT_BITS = 3, EPI_MASK = 4 };

__device__ __forceinline__ float4 lds_f4(uint32_t saddr)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
    return r;
}

// 32 accumulator columns of one row: bias (from shared memory at bias_saddr), activation, sign bits / derivative mask.
// RELU: negative slope is exactly 0 (max(t,0)); otherwise leaky with 0 <= slope <= 1 (max(t, slope*t)).
// Sign bits: bit j = 1 when the pre-activation of column j has its IEEE sign bit clear (t > 0, or t == +0), gathered with
// one funnel shift per element (columns walked from 31 down to 0, each pushing its sign into the low end).
template <int EPI, bool RELU>
__device__ __forceinline__ void epilogue_math32(uint32_t (&v)[32], uint32_t bias_saddr, float slope, uint32_t din, uint32_t &dout)
{
    constexpr bool kBias = EPI == EPI_BIAS || EPI == EPI_BIAS_ACT || EPI == EPI_BIAS_ACT_BITS;
    uint32_t neg = 0;
#pragma unroll
    for (int j4 = 7; j4 >= 0; --j4) {
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (kBias) { const float4 q = lds_f4(bias_saddr + j4 * 16); b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w; }
#pragma unroll
        for (int e = 3; e >= 0; --e) {
            const int j = j4 * 4 + e;
            float t = __uint_as_float(v[j]);
            if (kBias) t += b[e];
            if (EPI == EPI_BIAS_ACT_BITS) neg = __funnelshift_l(__float_as_uint(t), neg, 1);
            if (EPI == EPI_BIAS_ACT || EPI == EPI_BIAS_ACT_BITS) t = RELU ? fmaxf(t, 0.f) : fmaxf(t, t * slope);
            if (EPI == EPI_MASK) {
                if (RELU) t = __uint_as_float(__float_as_uint(t) & (uint32_t)((int32_t)(din << (31 - j)) >> 31));
                else t = ((din >> j) & 1u) ? t : t * slope;
            }
            v[j] = __float_as_uint(t);
        }
    }
    if (EPI == EPI_BIAS_ACT_BITS) dout |= ~neg;
}

template <int EPI>
__device__ __forceinline__ void epilogue_math(uint32_t (&v)[32], const float *bias_s, int col0, float slope, uint32_t din,
                                              uint32_t &dout)
{
    const uint32_t ba = smem_u32(bias_s + col0);
    if (slope == 0.f) epilogue_math32<EPI, true>(v, ba, slope, din, dout);
    else epilogue_math32<EPI, false>(v, ba, slope, din, dout);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
// two fp32 -> packed bf16 pair (lo in bits 0-15) with negative results clamped to +0: relu costs nothing extra
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// hidden-layer forward epilogue for 32 columns, relu flavour: w[k] = bf16x2(relu(v[2k] + b[2k]), relu(v[2k+1] + b[2k+1]));
// returns the sign-clear bits of the pre-activations (see epilogue_math32) when BITS
template <bool BITS>
__device__ __forceinline__ uint32_t bias_relu_pack32(const uint32_t (&v)[32], uint32_t bias_saddr, uint32_t (&w)[16])
{
    uint32_t neg = 0;
#pragma unroll
    for (int j4 = 7; j4 >= 0; --j4) {
        const float4 q = lds_f4(bias_saddr + j4 * 16);
        const float t0 = __uint_as_float(v[4 * j4]) + q.x, t1 = __uint_as_float(v[4 * j4 + 1]) + q.y;
        const float t2 = __uint_as_float(v[4 * j4 + 2]) + q.z, t3 = __uint_as_float(v[4 * j4 + 3]) + q.w;
        if (BITS) {
            neg = __funnelshift_l(__float_as_uint(t3), neg, 1);
            neg = __funnelshift_l(__float_as_uint(t2), neg, 1);
            neg = __funnelshift_l(__float_as_uint(t1), neg, 1);
            neg = __funnelshift_l(__float_as_uint(t0), neg, 1);
        }
        w[2 * j4] = pack_bf16_relu(t0, t1);
        w[2 * j4 + 1] = pack_bf16_relu(t2, t3);
    }
    return ~neg;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// relu'(.) mask applied to 32 columns while they are packed to bf16: w[k] = bf16x2(v[2k], v[2k+1]) with each half kept
// where its bit of `din` is set and zeroed where it is clear.  Eight shifted copies of din put every bit at the top of
// some byte; prmt's sign-replicate mode (selector bit 3) then expands two of those bytes into one 16|16-bit mask word.
__device__ __forceinline__ void mask_pack_relu32(const uint32_t (&v)[32], uint32_t din, uint32_t (&w)[16])
{
    uint32_t sh[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) sh[t] = din << t;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int j0 = 2 * k, j1 = 2 * k + 1;
        const int t0 = 7 - (j0 & 7), m0 = j0 >> 3, t1 = 7 - (j1 & 7), m1 = j1 >> 3;
        const uint32_t sel = (uint32_t)((8 | m0) * 0x11) | ((uint32_t)((8 | (4 + m1)) * 0x11) << 8);
        uint32_t mask;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(mask) : "r"(sh[t0]), "r"(sh[t1]), "r"(sel));
        w[k] = pack_bf16(__uint_as_float(v[j0]), __uint_as_float(v[j1])) & mask;
    }
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

}  // namespace papr
