// Stage a10 -- the SmallUNet decode (reference models/unet.py:196-258) as implicit-GEMM convolutions on the 5th-gen tensor
// cores: 3x3 / 1x1 convolutions and their data gradients (papr_conv_bf16), and their weight gradients (papr_conv_wgrad_bf16).
//
// Layout ("pixel planes").  A feature map of one image lives in a zero-padded raster: padded width Wp (a multiple of 8),
// pixel index p = y*Wp + x, G0 guard rows of zeros before and after.  Channels are split into blocks of 64; a block is a
// PLANE of S = G0 + L + G0 rows of 128 bytes (64 bf16), each row's eight 16-byte chunks XOR-swizzled by (row & 7) -- so any
// 128 consecutive rows that start at a multiple of 8 are byte-for-byte the K-major SWIZZLE_128B operand tile tcgen05.mma
// wants (and, read the other way, an MN-major one), exactly like a block of the MLP path's tile-blocked layout, and ONE
// 1-D TMA bulk copy fetches them.  A 3x3 tap (dy, dx) of output pixels [p0, p0+128) needs input pixels
// [p0 + dy*Wp + dx, ...): dy*Wp keeps the multiple-of-8 phase, dx = +-1 would not, so every map that feeds a 3x3
// convolution is stored as three copies shifted by dx = -1, 0, +1 (copy_dx[p] = X[p + dx]; 3x the write traffic of a
// tensor that is read 9 times).  The convolution is then 9 * (Cin/64) k-blocks of a plain GEMM accumulating in TMEM:
//     Y[p, :] = sum_taps sum_cb  A_tile(copy_dx, cb, rows p0 + dy*Wp) * W_image(tap, cb)^T
// and the data gradient is the same kernel with mirrored taps (sign = -1) and the weight image packed the other way.
// Zero borders (the padding of the reference's Conv2d(padding=1)) are part of the planes; the small raster kernels in
// unet_raster.cu re-establish them after every convolution.
//
// Kernel structure = papr_linear_bf16's (warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, 8 epilogue warps,
// two 256-column accumulators), except that the weight blocks stream through the ring next to the activation tiles
// (a 3x3 layer's weights are up to 2.4 MB and stay in L2).
#include "tc_common.cuh"

namespace papr {

constexpr int kConvThreads = 384;
constexpr int kConvMaxSmem = 232448;

struct ConvParams {
    const uint8_t *a;          // input planes (first used plane of copy dx = -1)
    int64_t a_copy_bytes;      // distance between the dx copies (unused when ntaps == 1)
    int64_t a_plane_bytes;     // S * 128
    int64_t a_row0;            // G0: row of pixel 0
    const uint8_t *w;          // weight image [ntaps*cbs][N][128 B]
    const float *bias;         // [N] or null
    uint8_t *y;                // output planes (copy 0 form), nblk_out of them, or null
    int64_t y_plane_bytes, y_row0;
    float *y_f32;              // fp32 row-major [n_tiles*128, ldy] or null
    int64_t ldy;
    const float *addend;       // fp32 row-major [n_tiles*128, ld_add] added to the accumulator before bias / activation, or null
    int64_t ld_add;
    int64_t n_tiles;
    int cbs, ntaps, Wp, sign, N, nblk_out, stages, b_bytes;
    float slope;
};

template <int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) conv_kernel(const ConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = kBlockBytes + p.b_bytes;
    uint8_t *ring = smem;
    uint8_t *stage_out = ring + p.stages * stage_bytes;            // one 16 KB staging buffer per epilogue set
    uint64_t *bars = (uint64_t *)(stage_out + 2 * kBlockBytes);
    uint64_t *full = bars, *empty = bars + 8, *tfull = bars + 16, *tempty = bars + 18;
    uint32_t *tmem_slot = (uint32_t *)(bars + 21);
    float *bias_s = (float *)(bars + 24);     // 256 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.ntaps * p.cbs;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 256; i += kConvThreads) bias_s[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int64_t row = p.a_row0 + tile * kTileRows;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int tap = kb / p.cbs, cb = kb - tap * p.cbs;
                    int dy = 0, dx = 0;
                    if (p.ntaps == 9) { dy = p.sign * (tap / 3 - 1); dx = p.sign * (tap % 3 - 1); }
                    const uint8_t *src = p.a + (int64_t)(p.ntaps == 9 ? dx + 1 : 0) * p.a_copy_bytes + (int64_t)cb * p.a_plane_bytes +
                                         (row + (int64_t)dy * p.Wp) * 128;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(kBlockBytes + p.N * 128));
                    bulk_g2s(ring + s * stage_bytes, src, kBlockBytes, &full[s]);
                    bulk_g2s(ring + s * stage_bytes + kBlockBytes, p.w + (size_t)kb * p.N * 128, (uint32_t)(p.N * 128), &full[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, p.N, false, false);
            int s = 0; uint32_t ph = 0; int64_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(&tempty[acc], (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(ring + s * stage_bytes);
                    const uint32_t b0 = a0 + kBlockBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
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
                if (p.addend) {      // partial sums of the split-bf16 fp32 mode (papr_b200/unet_fp32.py)
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
                uint32_t dummy = 0;
                epilogue_math<EPI>(v0, bias_s, col0, p.slope, 0u, dummy);
                if (second) epilogue_math<EPI>(v1, bias_s, col0 + 32, p.slope, 0u, dummy);
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
                if (p.y) {
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
                        bulk_s2g(p.y + (size_t)g * p.y_plane_bytes + (size_t)(p.y_row0 + tile * kTileRows) * 128, sbuf, kBlockBytes);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
        if (p.y && st == 0) bulk_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int EPI>
static int launch_conv(const ConvParams &p, int smem, cudaStream_t stream)
{
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, conv_kernel<EPI>, kConvMaxSmem));
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    conv_kernel<EPI><<<grid, kConvThreads, smem, stream>>>(p);
    return check_launch();
}

// --------------------------------------------------------------------------------------------------------------------
// Weight gradient of a convolution: C[tap][a, b] += sum_p A[p, a] * B[p + off(tap), b] with A (d output, unshifted copy) and
// B (the layer input: the dx copy of the tap, rows shifted by dy*Wp) both consumed as MN-major SWIZZLE_128B operands
// straight from the planes.  Same pipeline as papr_wgrad_bf16 (row slices streamed through a TMA ring, the whole product
// in TMEM, one atomic drain), but a CTA owns (tap, row slice): one launch covers all nine taps.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kCwThreads = 256;
constexpr int kCwHalf = kBlockBytes / 2;

struct ConvWgradParams {
    const uint8_t *a, *b;      // first used plane of each operand at row 0 of the pixel raster; b: copy dx = -1 when ntaps == 9
    int64_t a_plane_bytes, b_plane_bytes, b_copy_bytes;
    float *c;                  // fp32 [ntaps][a_valid rows, ldc] (+=, atomically), taps c_tap_stride floats apart
    int64_t c_tap_stride;
    int64_t n_units;           // 64-row units
    int a_halves, a_used_blk, nb_used, Nb, ldc, stages, a_valid, b_valid, ntaps, splits, Wp;
};

__global__ void __launch_bounds__(kCwThreads, 1) conv_wgrad_kernel(const ConvWgradParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (p.a_used_blk + p.nb_used) * kCwHalf;
    uint8_t *ring = smem;
    uint64_t *bars = (uint64_t *)(ring + p.stages * stage_bytes);
    uint64_t *full = bars, *empty = bars + 8, *done = bars + 16;
    uint32_t *tmem_slot = (uint32_t *)(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // one CTA = one tap x one slice of the rows: all nine taps of a 3x3 layer run in ONE launch, and only `splits` CTAs (not
    // the whole grid) add their partial products into a tap's gradient at the end
    const int tap = blockIdx.x / p.splits, split = blockIdx.x - tap * p.splits;
    const bool has_work = (int64_t)split < p.n_units;
    int dy = 0, dx = 0;
    if (p.ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    const uint8_t *bsrc = p.b + (int64_t)(p.ntaps == 9 ? dx + 1 : 0) * p.b_copy_bytes + (int64_t)dy * p.Wp * 128;
    float *cdst = p.c + (int64_t)tap * p.c_tap_stride;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t u = split; u < p.n_units; u += p.splits) {
                const size_t off = (size_t)u * kCwHalf;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                uint8_t *dst = ring + s * stage_bytes;
                for (int i = 0; i < p.a_used_blk; ++i) bulk_g2s(dst + i * kCwHalf, p.a + (size_t)i * p.a_plane_bytes + off, kCwHalf, &full[s]);
                for (int i = 0; i < p.nb_used; ++i)
                    bulk_g2s(dst + (p.a_used_blk + i) * kCwHalf, bsrc + (size_t)i * p.b_plane_bytes + off, kCwHalf, &full[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && has_work) {
            const uint32_t idesc = umma_idesc(128, p.Nb, true, true);
            int s = 0; uint32_t ph = 0; uint32_t first = 1;
            for (int64_t u = split; u < p.n_units; u += p.splits) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(ring + s * stage_bytes);
                const uint32_t b0 = a0 + p.a_used_blk * kCwHalf;
                for (int ks = 0; ks < 4; ++ks) {
                    for (int h = 0; h < p.a_halves; ++h) {
                        const uint64_t ad = umma_desc(a0 + h * 2 * kCwHalf + ks * 2048, kCwHalf, 1024);
                        const uint64_t bd = umma_desc(b0 + ks * 2048, kCwHalf, 1024);
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
        mbar_wait(done, 0);
        tc_fence_after();
        for (int h = 0; h < p.a_halves; ++h) {
            const int ai = h * 128 + row;
            for (int col0 = 0; col0 < p.Nb; col0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + h * 256 + col0, v);
                tmem_ld_wait();
                if (ai < p.a_valid) {
                    float *rowp = cdst + (size_t)ai * p.ldc + col0;
                    if (col0 + 32 <= p.b_valid && (((uintptr_t)rowp) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_v4(rowp + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int bi = col0 + j;
                            if (bi < p.b_valid) atomicAdd(cdst + (size_t)ai * p.ldc + bi, __uint_as_float(v[j]));
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

extern "C" int papr_conv_bf16(const void *in_planes, int64_t in_copy_bytes, int64_t in_plane_bytes, int64_t in_row0, int cbs, int ntaps,
                              int Wp, int sign, const void *w_image, const float *bias, int N, int act, float slope, void *out_planes,
                              int64_t out_plane_bytes, int64_t out_row0, float *out_f32, int64_t ld_f32, const float *addend_f32,
                              int64_t ld_addend, int64_t n_tiles, void *stream)
{
    using namespace papr;
    if (!in_planes || !w_image || (!out_planes && !out_f32)) return PAPR_ERR_INVALID_ARGUMENT;
    if (addend_f32 && (ld_addend < N || ld_addend % 4)) return PAPR_ERR_INVALID_ARGUMENT;
    if (n_tiles <= 0 || N < 32 || N > 256 || N % 32 || cbs < 1 || (ntaps != 1 && ntaps != 9) || (sign != 1 && sign != -1)) return PAPR_ERR_INVALID_ARGUMENT;
    if (ntaps == 9 && (Wp < 8 || Wp % 8 || in_row0 < Wp + 1)) return PAPR_ERR_INVALID_ARGUMENT;
    if (out_f32 && (ld_f32 < N || ld_f32 % 4)) return PAPR_ERR_INVALID_ARGUMENT;
    if (act && !bias) return PAPR_ERR_INVALID_ARGUMENT;
    ConvParams p;
    p.a = (const uint8_t *)in_planes; p.a_copy_bytes = in_copy_bytes; p.a_plane_bytes = in_plane_bytes; p.a_row0 = in_row0;
    p.w = (const uint8_t *)w_image; p.bias = bias; p.y = (uint8_t *)out_planes; p.y_plane_bytes = out_plane_bytes; p.y_row0 = out_row0;
    p.y_f32 = out_f32; p.ldy = ld_f32; p.addend = addend_f32; p.ld_add = ld_addend; p.n_tiles = n_tiles; p.cbs = cbs; p.ntaps = ntaps; p.Wp = Wp; p.sign = sign; p.N = N;
    p.nblk_out = (N + 63) / 64; p.slope = slope;
    p.b_bytes = (N * 128 + 1023) & ~1023;
    const int fixed = 1024 + 2 * kBlockBytes + 1280;
    p.stages = (kConvMaxSmem - fixed) / (kBlockBytes + p.b_bytes);
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return PAPR_ERR_INVALID_ARGUMENT;
    const int smem = fixed + p.stages * (kBlockBytes + p.b_bytes);
    cudaStream_t s = (cudaStream_t)stream;
    if (act) return launch_conv<EPI_BIAS_ACT>(p, smem, s);
    if (bias) return launch_conv<EPI_BIAS>(p, smem, s);
    return launch_conv<EPI_PLAIN>(p, smem, s);
}

extern "C" int papr_conv_wgrad_bf16(const void *a_planes, int64_t a_plane_bytes, int a_valid, const void *b_planes, int64_t b_plane_bytes,
                                    int64_t b_copy_bytes, int b_valid, int ntaps, int Wp, float *c, int64_t ldc, int64_t c_tap_stride,
                                    int64_t rows, void *stream)
{
    using namespace papr;
    if (!a_planes || !b_planes || !c) return PAPR_ERR_INVALID_ARGUMENT;
    if (rows <= 0 || rows % 64 || a_valid < 1 || a_valid > 256 || b_valid < 1 || b_valid > 256 || (ntaps != 1 && ntaps != 9)) return PAPR_ERR_INVALID_ARGUMENT;
    if (ntaps == 9 && (Wp < 8 || Wp % 8)) return PAPR_ERR_INVALID_ARGUMENT;
    ConvWgradParams p;
    p.a = (const uint8_t *)a_planes; p.b = (const uint8_t *)b_planes; p.a_plane_bytes = a_plane_bytes; p.b_plane_bytes = b_plane_bytes;
    p.b_copy_bytes = b_copy_bytes; p.ntaps = ntaps; p.Wp = Wp; p.c_tap_stride = c_tap_stride;
    p.c = c; p.n_units = rows / 64;
    p.a_halves = (a_valid + 127) / 128;                 // M = 128 per MMA: 1 or 2 row-halves of the product
    p.a_used_blk = p.a_halves * 2;                      // the caller provides ceil(a_valid/128)*2 planes (zero padded)
    p.Nb = (b_valid + 15) & ~15;
    p.nb_used = (p.Nb + 63) / 64;
    p.ldc = (int)ldc; p.a_valid = a_valid; p.b_valid = b_valid;
    const int stage_bytes = (p.a_used_blk + p.nb_used) * kCwHalf;
    p.stages = (232448 - 1024 - 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    const int smem = 1024 + p.stages * stage_bytes + 1024;
    static SmemAttrOnce once;
    PAPR_CUDA_TRY(ensure_dyn_smem(once, conv_wgrad_kernel, 232448));
    const int64_t per_tap = kNumSMs / ntaps;                        // 16 row slices per tap for 3x3, the whole grid for 1x1
    p.splits = (int)(p.n_units < per_tap ? p.n_units : per_tap);
    conv_wgrad_kernel<<<p.splits * ntaps, kCwThreads, smem, (cudaStream_t)stream>>>(p);
    return check_launch();
}
