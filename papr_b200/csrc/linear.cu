// Stage a7/a8/a12 building block: one Linear layer of the proximity-attention MLP stacks on the 5th-gen tensor cores
// (reference models/mlp.py:53-58, models/attn.py:217-218 forward; the dgrad half of their autograd backward).
//
//   Y[M,N] = epilogue( X[M,K] * W[N,K]^T )            bf16 x bf16 -> fp32 in TMEM
//
// Persistent, warp-specialised, one CTA per SM (384 threads):
//   warp 0    TMA producer : the whole weight image once (it stays resident), then X blocks (128 rows x 64 cols,
//                            16 KB, already in the UMMA shared-memory layout) through a ring of mbarrier stages
//   warp 1    MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128, N=N, K=16 per instruction, accumulating a
//                            128 x N fp32 tile in TMEM; two accumulators so tile t+1 multiplies while t drains
//   warp 2    TMEM allocator
//   warps 4-11 epilogue    : two sets of four warps (TMEM lane quadrant = warp % 4); set s drains the 64-column groups
//                            g with (g & 1) == s: tcgen05.ld -> bias / activation / activation-derivative mask -> bf16
//                            -> swizzled staging -> TMA bulk store of the next layer's operand block; optional fp32
//                            row-major output, activation sign bits (for backward), per-column sums (bias gradients)
#include "tc_common.cuh"

namespace papr {

constexpr int kLinThreads = 384;
constexpr int kMaxSmem = 232448;   // 227 KB

struct LinearParams {
    const uint8_t *x;        // blocked bf16 [M, kblk*64]
    const uint8_t *w;        // weight image: kblk blocks of [N rows x 128 B], K-major SWIZZLE_128B
    const float *bias;       // [N] or null
    uint8_t *y_blocked;      // blocked bf16 [M, nblk_out*64] or null
    float *y_f32;            // row-major [M, ldy] or null
    uint64_t *bits_out;      // [M/128][nblk_out][128] sign bits of the pre-activation (bit j of word g: column 64g+j > 0) or null
    const uint64_t *bits_in; // dgrad: multiply column j by act'(.) read from these bits, or null
    float *colsum;           // [N] += column sums of the bf16 output (atomic), or null
    const float *addend;     // fp32 row-major [M, ld_add] added to the accumulator before bias/activation, or null
    int64_t ld_add;
    int64_t n_tiles;
    int N, kblk, k_steps, nblk_out, act, ldy, stages;
    float slope;             // negative slope of the activation (0 relu, 0.2 leakyrelu) for act and for bits_in
};

template <int EPI>
__global__ void __launch_bounds__(kLinThreads, 1) linear_kernel(const LinearParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int wbytes = p.kblk * p.N * 128;
    uint8_t *w_s = smem;
    uint8_t *ring = w_s + ((wbytes + 1023) & ~1023);
    uint8_t *stage_out = ring + p.stages * kBlockBytes;            // one 16 KB staging buffer per epilogue set
    uint64_t *bars = (uint64_t *)(stage_out + 2 * kBlockBytes);
    uint64_t *full = bars, *empty = bars + 8, *tfull = bars + 16, *tempty = bars + 18, *wbar = bars + 20;
    uint32_t *tmem_slot = (uint32_t *)(bars + 21);
    float *colsum_s = (float *)(bars + 24);   // 256 floats: column sums (dgrad) ...
    float *bias_s = colsum_s;                 // ... or the bias vector (forward); never both (checked by the caller)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 256; i += kLinThreads) {
        colsum_s[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, (uint32_t)wbytes);
            for (int kb = 0; kb < p.kblk; ++kb)
                bulk_g2s(w_s + kb * p.N * 128, p.w + (size_t)kb * p.N * 128, (uint32_t)(p.N * 128), wbar);
            int s = 0; uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.kblk; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], kBlockBytes);
                    bulk_g2s(ring + s * kBlockBytes, p.x + ((size_t)tile * p.kblk + kb) * kBlockBytes, kBlockBytes, &full[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, p.N, false, false);
            mbar_wait(wbar, 0);
            int s = 0; uint32_t ph = 0; int64_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(&tempty[acc], (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                for (int kb = 0; kb < p.kblk; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(ring + s * kBlockBytes);
                    const uint32_t b0 = smem_u32(w_s + kb * p.N * 128);
                    const int nk = min(4, p.k_steps - 4 * kb);
                    for (int k = 0; k < nk; ++k)
                        umma_bf16(d, umma_desc(a0 + k * 32, 16, 1024), umma_desc(b0 + k * 32, 16, 1024), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int set = ew >> 2;                        // which 64-column groups this warp drains
        const int quad = ew & 3;                        // TMEM lane quadrant (== warp % 4)
        const int row = quad * 32 + lane;
        const int st = (ew & 3) * 32 + lane;            // thread index within the set (0..127)
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int ngroups = (p.N + 63) >> 6;
        uint8_t *sbuf = stage_out + set * kBlockBytes;
        const int bar_id = 1 + set;
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int acc = (int)(it & 1);
            const int64_t grow = tile * kTileRows + row;
            uint64_t din[2] = {0, 0};
            if (EPI == EPI_MASK) {
                if (set < ngroups) din[0] = p.bits_in[(tile * p.nblk_out + set) * kTileRows + row];
                if (set + 2 < ngroups) din[1] = p.bits_in[(tile * p.nblk_out + set + 2) * kTileRows + row];
            }
            mbar_wait(&tfull[acc], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            for (int gi = 0; gi < 2; ++gi) {
                const int g = set + 2 * gi;
                if (g >= ngroups) break;
                const int col0 = g * 64;
                uint32_t v0[32], v1[32];
                const bool second = col0 + 32 < p.N;
                tmem_ld32(tmem_base + lane_base + acc * 256 + col0, v0);
                if (second) tmem_ld32(tmem_base + lane_base + acc * 256 + col0 + 32, v1);
                tmem_ld_wait();
                if (!second) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v1[j] = 0;
                }
                if (p.addend) {      // partial product of a split-K layer (skip connections), see papr_linear_bf16
                    const float4 *src = reinterpret_cast<const float4 *>(p.addend + grow * p.ld_add + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 a = __ldg(src + j);
                        v0[4 * j] = __float_as_uint(__uint_as_float(v0[4 * j]) + a.x);
                        v0[4 * j + 1] = __float_as_uint(__uint_as_float(v0[4 * j + 1]) + a.y);
                        v0[4 * j + 2] = __float_as_uint(__uint_as_float(v0[4 * j + 2]) + a.z);
                        v0[4 * j + 3] = __float_as_uint(__uint_as_float(v0[4 * j + 3]) + a.w);
                    }
                    if (second) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 a = __ldg(src + 8 + j);
                            v1[4 * j] = __float_as_uint(__uint_as_float(v1[4 * j]) + a.x);
                            v1[4 * j + 1] = __float_as_uint(__uint_as_float(v1[4 * j + 1]) + a.y);
                            v1[4 * j + 2] = __float_as_uint(__uint_as_float(v1[4 * j + 2]) + a.z);
                            v1[4 * j + 3] = __float_as_uint(__uint_as_float(v1[4 * j + 3]) + a.w);
                        }
                    }
                }
                uint32_t blo = 0, bhi = 0;
                epilogue_math<EPI>(v0, bias_s, col0, p.slope, (uint32_t)din[gi], blo);
                if (second) epilogue_math<EPI>(v1, bias_s, col0 + 32, p.slope, (uint32_t)(din[gi] >> 32), bhi);
                if (EPI == EPI_BIAS_ACT_BITS) p.bits_out[(tile * p.nblk_out + g) * kTileRows + row] = ((uint64_t)bhi << 32) | blo;
                if (p.y_f32) {
                    float4 *dst = reinterpret_cast<float4 *>(p.y_f32 + grow * p.ldy + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v0[4 * j]), __uint_as_float(v0[4 * j + 1]), __uint_as_float(v0[4 * j + 2]), __uint_as_float(v0[4 * j + 3]));
                    if (second) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dst[8 + j] = make_float4(__uint_as_float(v1[4 * j]), __uint_as_float(v1[4 * j + 1]), __uint_as_float(v1[4 * j + 2]), __uint_as_float(v1[4 * j + 3]));
                    }
                }
                if (p.y_blocked) {
                    uint4 q[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        q[c] = make_uint4(pack_bf16(__uint_as_float(v0[8 * c]), __uint_as_float(v0[8 * c + 1])),
                                          pack_bf16(__uint_as_float(v0[8 * c + 2]), __uint_as_float(v0[8 * c + 3])),
                                          pack_bf16(__uint_as_float(v0[8 * c + 4]), __uint_as_float(v0[8 * c + 5])),
                                          pack_bf16(__uint_as_float(v0[8 * c + 6]), __uint_as_float(v0[8 * c + 7])));
                        q[4 + c] = make_uint4(pack_bf16(__uint_as_float(v1[8 * c]), __uint_as_float(v1[8 * c + 1])),
                                              pack_bf16(__uint_as_float(v1[8 * c + 2]), __uint_as_float(v1[8 * c + 3])),
                                              pack_bf16(__uint_as_float(v1[8 * c + 4]), __uint_as_float(v1[8 * c + 5])),
                                              pack_bf16(__uint_as_float(v1[8 * c + 6]), __uint_as_float(v1[8 * c + 7])));
                    }
                    if (st == 0) bulk_wait_read<0>();      // the previous store from this staging buffer has been read
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4 *>(sbuf + row * 128 + ((c ^ (row & 7)) << 4)) = q[c];
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (st == 0) {
                        bulk_s2g(p.y_blocked + ((size_t)tile * p.nblk_out + g) * kBlockBytes, sbuf, kBlockBytes);
                        bulk_commit();
                    }
                    if (p.colsum) {
                        // thread (qq = st>>5, l = st&31): columns 2l, 2l+1 of this group over rows [32qq, 32qq+32)
                        const int qq = st >> 5, l = st & 31;
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                        for (int r = qq * 32; r < qq * 32 + 32; ++r) {
                            const uint32_t w2 = *reinterpret_cast<const uint32_t *>(sbuf + r * 128 + (((l >> 2) ^ (r & 7)) << 4) + (l & 3) * 4);
                            s0 += bf16_lo(w2); s1 += bf16_hi(w2);
                        }
                        atomicAdd(&colsum_s[col0 + 2 * l], s0);
                        atomicAdd(&colsum_s[col0 + 2 * l + 1], s1);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
        if (p.y_blocked && st == 0) bulk_wait<0>();
        if (p.colsum) {
            asm volatile("bar.sync 3, 256;" ::: "memory");
            for (int c = threadIdx.x - 128; c < p.N; c += 256) atomicAdd(p.colsum + c, colsum_s[c]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int EPI>
static int launch_linear(const LinearParams &p, int smem, cudaStream_t stream)
{
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, linear_kernel<EPI>, kMaxSmem));
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    linear_kernel<EPI><<<grid, kLinThreads, smem, stream>>>(p);
    return check_launch();
}

}  // namespace papr

extern "C" int papr_linear_bf16(const void *x, const void *w_image, const float *bias, void *y_blocked, float *y_f32,
                                int64_t ldy, uint64_t *sign_bits_out, const uint64_t *sign_bits_in, float *colsum,
                                const float *addend_f32, int64_t ld_addend, int64_t rows, int N, int K, int act,
                                float slope, void *stream)
{
    using namespace papr;
    if (!x || !w_image || (!y_blocked && !y_f32)) return PAPR_ERR_INVALID_ARGUMENT;
    if (rows <= 0 || rows % kTileRows || N < 32 || N > 256 || N % 32 || K < 16 || K > 256 || K % 16) return PAPR_ERR_INVALID_ARGUMENT;
    if (y_f32 && (ldy < N || ldy % 4)) return PAPR_ERR_INVALID_ARGUMENT;
    if (addend_f32 && (ld_addend < N || ld_addend % 4)) return PAPR_ERR_INVALID_ARGUMENT;
    if (sign_bits_in && (bias || act || sign_bits_out)) return PAPR_ERR_INVALID_ARGUMENT;   // dgrad mode is exclusive
    if (sign_bits_out && !(bias && act)) return PAPR_ERR_INVALID_ARGUMENT;
    if (act && !bias) return PAPR_ERR_INVALID_ARGUMENT;
    if (colsum && (!y_blocked || bias)) return PAPR_ERR_INVALID_ARGUMENT;
    LinearParams p;
    p.x = (const uint8_t *)x; p.w = (const uint8_t *)w_image; p.bias = bias;
    p.y_blocked = (uint8_t *)y_blocked; p.y_f32 = y_f32; p.bits_out = sign_bits_out; p.bits_in = sign_bits_in;
    p.addend = addend_f32; p.ld_add = ld_addend;
    p.colsum = colsum; p.n_tiles = rows / kTileRows; p.N = N; p.kblk = (K + 63) / 64; p.k_steps = K / 16;
    p.nblk_out = (N + 63) / 64; p.act = act; p.ldy = (int)ldy; p.slope = slope;
    const int wbytes = ((p.kblk * N * 128) + 1023) & ~1023;
    const int fixed = 1024 + wbytes + 2 * kBlockBytes + 1280;
    p.stages = (kMaxSmem - fixed) / kBlockBytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return PAPR_ERR_INVALID_ARGUMENT;
    const int smem = fixed + p.stages * kBlockBytes;
    cudaStream_t s = (cudaStream_t)stream;
    if (sign_bits_in) return launch_linear<EPI_MASK>(p, smem, s);
    if (sign_bits_out) return launch_linear<EPI_BIAS_ACT_BITS>(p, smem, s);
    if (act) return launch_linear<EPI_BIAS_ACT>(p, smem, s);
    if (bias) return launch_linear<EPI_BIAS>(p, smem, s);
    return launch_linear<EPI_PLAIN>(p, smem, s);
}
