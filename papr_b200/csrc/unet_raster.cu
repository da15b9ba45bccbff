// Raster-side kernels of the SmallUNet decode (stage a10, reference models/unet.py:196-258): everything that is not a
// contraction -- moving feature maps between the row-major world and the "pixel planes" the convolution kernels read
// (see conv.cu), 2x2 max-pooling and its backward, the pixel shuffle of the 2x2 stride-2 transposed convolution, ReLU
// masks, FiLM, bias-gradient column sums.  One thread handles one 16-byte chunk (8 channels) of one pixel.
//
// Geometry of a map of H x W pixels: padded width Wp (multiple of 8), pixel (y, x) -> raster index (y+1)*Wp + (x+1),
// plane row = G0 + raster index; copy dx of a 3-copy map holds X[p + dx] at row p, i.e. pixel p is WRITTEN to row p - dx.
// Rows that correspond to padding are never written and stay zero from the allocation.
#include "tc_common.cuh"

namespace papr {

struct Raster { int H, W, Wp; int64_t row0, plane_bytes, copy_bytes; };
constexpr int kMaxSumCols = 1024;       // widest map whose per-channel sums a raster kernel accumulates

__device__ __forceinline__ uint8_t *chunk_ptr(uint8_t *planes, const Raster &g, int copy, int cb, int64_t row, int c)
{
    return planes + (int64_t)copy * g.copy_bytes + (int64_t)cb * g.plane_bytes + row * 128 + (((int64_t)c ^ (row & 7)) << 4);
}
__device__ __forceinline__ const uint8_t *chunk_ptr(const uint8_t *planes, const Raster &g, int copy, int cb, int64_t row, int c)
{
    return planes + (int64_t)copy * g.copy_bytes + (int64_t)cb * g.plane_bytes + row * 128 + (((int64_t)c ^ (row & 7)) << 4);
}
__device__ __forceinline__ void unpack8f(const uint4 &q, float *f)
{
    f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
    f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8f(const float *f)
{
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
// write one chunk of pixel (y, x) to the `ncopies` shifted copies of a map (ncopies = 1: the unshifted copy only)
__device__ __forceinline__ void store_copies(uint8_t *planes, const Raster &g, int ncopies, int cb, int y, int x, int c, const uint4 &q)
{
    const int64_t p = g.row0 + (int64_t)(y + 1) * g.Wp + (x + 1);
    if (ncopies == 1) { *reinterpret_cast<uint4 *>(chunk_ptr(planes, g, 0, cb, p, c)) = q; return; }
#pragma unroll
    for (int d = 0; d < 3; ++d) *reinterpret_cast<uint4 *>(chunk_ptr(planes, g, d, cb, p - (d - 1), c)) = q;
}

// fp32 (H, W, C) row-major [optionally x*gamma + beta per channel] -> planes (channels zero-padded to 64*cbs)
__global__ void __launch_bounds__(256) nhwc_to_planes_kernel(const float *__restrict__ src, int64_t ld_pix, int C, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, uint8_t *__restrict__ dst, Raster g, int ncopies, int cbs)
{
    const int64_t total = (int64_t)g.H * g.W * cbs * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % cbs);
        const int64_t pix = i / (8 * cbs);
        const int y = (int)(pix / g.W), x = (int)(pix % g.W);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ch = cb * 64 + c * 8 + e;
            float v = 0.f;
            if (ch < C) {
                v = src[pix * ld_pix + ch];
                if (gamma) v = fmaf(v, gamma[ch], beta[ch]);
            }
            f[e] = v;
        }
        store_copies(dst, g, ncopies, cb, y, x, c, pack8f(f));
    }
}

// Zero everything of a freshly allocated map that the raster kernels never write: the guard rows before and after the
// raster and its one-pixel (right: up to seven-pixel) frame, per shifted copy.  The interior -- all of it -- is written by
// the kernel that produces the map, so a full memset of the buffer (0.25 ms for a 256-channel 800x800 map, 1 ms per
// training step over all maps) is 97% redundant.
// In raster order the frame is H + 1 contiguous runs of pixel rows per plane: [guard rows, line 0, left pad of line 1], then
// for every line its right pad together with the left pad of the next line, and finally [right pad of line H, line H + 1,
// guard rows].  One warp zeroes one run (a first version tested every 16-byte chunk of the buffer for "interior": two
// 64-bit divisions per chunk, 0.1 ms per map and 1.1 ms per training step for 3% of the bytes).
__global__ void __launch_bounds__(256) zero_border_kernel(uint8_t *__restrict__ planes, Raster g, int ncopies, int cbs, int64_t rows_total)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pc = blockIdx.y, cb = pc % cbs, d = pc / cbs;
    // copy d of three holds pixel row p at p - (d - 1) (store_copies); a single copy is unshifted
    const int64_t shift = g.row0 - (ncopies == 3 ? d - 1 : 0);          // raster index q lives in pixel row q + shift
    uint8_t *plane = planes + (int64_t)d * g.copy_bytes + (int64_t)cb * g.plane_bytes;
    for (int s = blockIdx.x * 8 + warp; s <= g.H; s += gridDim.x * 8) {
        int64_t r0 = s == 0 ? 0 : (int64_t)s * g.Wp + g.W + 1 + shift;
        int64_t r1 = s == g.H ? rows_total : (int64_t)(s + 1) * g.Wp + 1 + shift;
        r0 = r0 < 0 ? 0 : r0;
        r1 = r1 > rows_total ? rows_total : r1;
        uint4 *dst = reinterpret_cast<uint4 *>(plane + r0 * 128);
        const int64_t n = (r1 - r0) * 8;
        for (int64_t i = lane; i < n; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
    }
}

// planes (unshifted copy) -> fp32 (H, W, C) row-major
__global__ void __launch_bounds__(256) planes_to_nhwc_kernel(const uint8_t *__restrict__ src, Raster g, int cbs, float *__restrict__ dst, int64_t ld_pix, int C)
{
    const int64_t total = (int64_t)g.H * g.W * cbs * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % cbs);
        const int64_t pix = i / (8 * cbs);
        const int y = (int)(pix / g.W), x = (int)(pix % g.W);
        const int64_t p = g.row0 + (int64_t)(y + 1) * g.Wp + (x + 1);
        float f[8];
        unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(src, g, 0, cb, p, c)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ch = cb * 64 + c * 8 + e;
            if (ch < C) dst[pix * ld_pix + ch] = f[e];
        }
    }
}

// The work-horse between two convolutions: value = src[p] (+ add[p]) (+ unpooled gradient), masked by (mask[p] > 0),
// optionally FiLM-modulated, written to the interior of `dst` (1 or 3 copies); optional per-channel sums of what was written
// (bias gradients).  src / add / mask are unshifted planes of the same raster.
struct SpreadParams {
    const uint8_t *src; int src_cb0;
    const uint8_t *add; int add_cb0;
    const uint8_t *mask; int mask_cb0;
    const uint8_t *pool_grad; Raster pg;        // gradient w.r.t. the pooled map (half resolution), or null
    const uint8_t *pool_ref; int pool_ref_cb0;  // the map that was pooled (this raster), for the argmax
    const float *gamma, *beta;                  // FiLM on the stored value (channels of dst), or null
    uint8_t *dst; Raster dg; int dst_cb0, ncopies;
    float *colsum;                              // [64*cbs] +=, or null
    Raster g;                                   // raster of src / add / mask
    int cbs;
};

__global__ void __launch_bounds__(256) spread_kernel(const SpreadParams p)
{
    __shared__ float ssum[kMaxSumCols];     // per-block column sums (one global atomic per column and block)
    if (p.colsum) for (int i = threadIdx.x; i < 64 * p.cbs; i += blockDim.x) ssum[i] = 0.f;
    if (p.colsum) __syncthreads();
    const int64_t total = (int64_t)p.g.H * p.g.W * p.cbs * 8;
    // consecutive threads walk chunks of one pixel, then cbs, then pixels: a thread's (cb, c) is fixed when the grid stride is
    // a multiple of 8*cbs, which the launcher guarantees -- so column sums can stay in registers
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    int my_cb = -1, my_c = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % p.cbs);
        my_cb = cb; my_c = c;
        const int64_t pix = i / (8 * p.cbs);
        const int y = (int)(pix / p.g.W), x = (int)(pix % p.g.W);
        const int64_t row = p.g.row0 + (int64_t)(y + 1) * p.g.Wp + (x + 1);
        float f[8];
        unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(p.src, p.g, 0, p.src_cb0 + cb, row, c)), f);
        if (p.add) {
            float a[8];
            unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(p.add, p.g, 0, p.add_cb0 + cb, row, c)), a);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += a[e];
        }
        if (p.pool_grad) {
            // max_pool2d(2) backward: the gradient of a pooled cell goes to the FIRST maximum of its 2x2 window in
            // row-major order (torch's rule); pixels outside any complete window get nothing
            const int cy = y >> 1, cx = x >> 1;
            if (cy < p.pg.H && cx < p.pg.W) {
                float best[8], gr[8];
                int arg[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int yy = 2 * cy + (w >> 1), xx = 2 * cx + (w & 1);
                    const int64_t r2 = p.g.row0 + (int64_t)(yy + 1) * p.g.Wp + (xx + 1);
                    float v[8];
                    unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(p.pool_ref, p.g, 0, p.pool_ref_cb0 + cb, r2, c)), v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) if (w == 0 || v[e] > best[e]) { best[e] = v[e]; arg[e] = w; }
                }
                const int me = ((y & 1) << 1) | (x & 1);
                const int64_t prow = p.pg.row0 + (int64_t)(cy + 1) * p.pg.Wp + (cx + 1);
                unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(p.pool_grad, p.pg, 0, cb, prow, c)), gr);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (arg[e] == me) f[e] += gr[e];
            }
        }
        if (p.mask) {
            float m[8];
            unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(p.mask, p.g, 0, p.mask_cb0 + cb, row, c)), m);
#pragma unroll
            for (int e = 0; e < 8; ++e) if (!(m[e] > 0.f)) f[e] = 0.f;
        }
        if (p.gamma) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { const int ch = cb * 64 + c * 8 + e; f[e] = fmaf(f[e], p.gamma[ch], p.beta[ch]); }
        }
        const uint4 q = pack8f(f);
        if (p.colsum) {
            float r[8];
            unpack8f(q, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += r[e];
        }
        if (p.dst) store_copies(p.dst, p.dg, p.ncopies, p.dst_cb0 + cb, y, x, c, q);
    }
    if (p.colsum) {
        if (my_cb >= 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&ssum[my_cb * 64 + my_c * 8 + e], acc[e]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * p.cbs; i += blockDim.x) if (ssum[i] != 0.f) atomicAdd(p.colsum + i, ssum[i]);
    }
}

// 2x2 max pooling (floor): planes (unshifted copy) at H x W -> 3-copy planes at floor(H/2) x floor(W/2)
__global__ void __launch_bounds__(256) pool_kernel(const uint8_t *__restrict__ src, Raster g, int src_cb0, uint8_t *__restrict__ dst, Raster dg,
                                                   int ncopies, int cbs)
{
    const int64_t total = (int64_t)dg.H * dg.W * cbs * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % cbs);
        const int64_t pix = i / (8 * cbs);
        const int cy = (int)(pix / dg.W), cx = (int)(pix % dg.W);
        float best[8];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t row = g.row0 + (int64_t)(2 * cy + (w >> 1) + 1) * g.Wp + (2 * cx + (w & 1) + 1);
            float v[8];
            unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(src, g, 0, src_cb0 + cb, row, c)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) best[e] = (w == 0) ? v[e] : fmaxf(best[e], v[e]);
        }
        store_copies(dst, dg, ncopies, cb, cy, cx, c, pack8f(best));
    }
}

// Pixel shuffle of ConvTranspose2d(kernel 2, stride 2): low-resolution GEMM output with channels ordered (a, b, co)
// -> high-resolution map: out[(2i+a+pad_y, 2j+b+pad_x), co] = in[(i,j), (a*2+b)*cout + co] + bias[co]
__global__ void __launch_bounds__(256) convt_scatter_kernel(const uint8_t *__restrict__ src, Raster g /* low res */, int cout, const float *__restrict__ bias,
                                                           uint8_t *__restrict__ dst, Raster dg /* high res */, int dst_cb0, int ncopies, int pad_y, int pad_x)
{
    const int cbs = cout / 64;
    const int64_t total = (int64_t)g.H * 2 * g.W * 2 * cbs * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % cbs);
        const int64_t pix = i / (8 * cbs);
        const int yy = (int)(pix / (2 * g.W)), xx = (int)(pix % (2 * g.W));
        const int a = yy & 1, b = xx & 1;
        const int64_t row = g.row0 + (int64_t)((yy >> 1) + 1) * g.Wp + ((xx >> 1) + 1);
        const int ch0 = (a * 2 + b) * cout + cb * 64 + c * 8;       // first of the 8 source channels
        float f[8];
        unpack8f(*reinterpret_cast<const uint4 *>(chunk_ptr(src, g, 0, ch0 >> 6, row, (ch0 & 63) >> 3)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += bias[cb * 64 + c * 8 + e];
        const int oy = yy + pad_y, ox = xx + pad_x;
        if (oy < dg.H && ox < dg.W) store_copies(dst, dg, ncopies, dst_cb0 + cb, oy, ox, c, pack8f(f));
    }
}

// The reverse gather for the backward pass: high-resolution gradient (unshifted planes) -> low-resolution planes with
// channels (a, b, co); optional per-channel sums of the gathered values (the transposed convolution's bias gradient)
__global__ void __launch_bounds__(256) convt_gather_kernel(const uint8_t *__restrict__ src, Raster g /* high res */, int src_cb0, int cout,
                                                          uint8_t *__restrict__ dst, Raster dg /* low res */, int pad_y, int pad_x, float *__restrict__ colsum)
{
    __shared__ float ssum[kMaxSumCols];
    const int cbs = cout / 64;
    if (colsum) for (int i = threadIdx.x; i < cout; i += blockDim.x) ssum[i] = 0.f;
    if (colsum) __syncthreads();
    const int64_t total = (int64_t)dg.H * 2 * dg.W * 2 * cbs * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    int my_cb = -1, my_c = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 8), cb = (int)((i / 8) % cbs);
        my_cb = cb; my_c = c;
        const int64_t pix = i / (8 * cbs);
        const int yy = (int)(pix / (2 * dg.W)), xx = (int)(pix % (2 * dg.W));
        const int a = yy & 1, b = xx & 1;
        const int oy = yy + pad_y, ox = xx + pad_x;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (oy < g.H && ox < g.W) {
            const int64_t row = g.row0 + (int64_t)(oy + 1) * g.Wp + (ox + 1);
            q = *reinterpret_cast<const uint4 *>(chunk_ptr(src, g, 0, src_cb0 + cb, row, c));
        }
        if (colsum) {
            float r[8];
            unpack8f(q, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += r[e];
        }
        const int64_t drow = dg.row0 + (int64_t)((yy >> 1) + 1) * dg.Wp + ((xx >> 1) + 1);
        const int ch0 = (a * 2 + b) * cout + cb * 64 + c * 8;
        *reinterpret_cast<uint4 *>(chunk_ptr(dst, dg, 0, ch0 >> 6, drow, (ch0 & 63) >> 3)) = q;
    }
    if (colsum) {
        if (my_cb >= 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&ssum[my_cb * 64 + my_c * 8 + e], acc[e]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cout; i += blockDim.x) if (ssum[i] != 0.f) atomicAdd(colsum + i, ssum[i]);
    }
}

static int raster_grid(int64_t total, int cbs)
{
    // a multiple of 8*cbs threads in total, so that a thread keeps its (channel block, chunk) over the grid-stride loop
    const int64_t unit = 8 * cbs;                               // threads per pixel
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    // blocks * 256 must be a multiple of `unit`: 256 = 32 * 8, so any block count works when cbs divides 32*blocks
    while ((blocks * 256) % unit) ++blocks;
    return (int)blocks;
}

static Raster make_raster(const papr_raster &r)
{
    Raster g;
    g.H = r.H; g.W = r.W; g.Wp = r.Wp; g.row0 = r.row0; g.plane_bytes = r.plane_bytes; g.copy_bytes = r.copy_bytes;
    return g;
}

}  // namespace papr

using namespace papr;

static bool raster_ok(const papr_raster *r)
{
    return r && r->H > 0 && r->W > 0 && r->Wp >= r->W + 2 && r->Wp % 8 == 0 && r->row0 >= r->Wp + 1 && r->plane_bytes > 0;
}

extern "C" int papr_unet_pack_input(const float *src, int64_t ld_pix, int C, const float *gamma, const float *beta, void *dst_planes,
                                    const papr_raster *geom, int ncopies, int cbs, void *stream)
{
    if (!src || !dst_planes || !raster_ok(geom) || C < 1 || cbs < 1 || C > 64 * cbs || (ncopies != 1 && ncopies != 3) || ((gamma == nullptr) != (beta == nullptr)))
        return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)geom->H * geom->W * cbs * 8;
    nhwc_to_planes_kernel<<<raster_grid(total, cbs), 256, 0, (cudaStream_t)stream>>>(src, ld_pix, C, gamma, beta, (uint8_t *)dst_planes, make_raster(*geom), ncopies, cbs);
    return check_launch();
}

extern "C" int papr_unet_zero_border(void *planes, const papr_raster *geom, int ncopies, int cbs, void *stream)
{
    if (!planes || !raster_ok(geom) || cbs < 1 || (ncopies != 1 && ncopies != 3) || geom->plane_bytes % 128) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t rows_total = geom->plane_bytes / 128;
    const dim3 grid((unsigned)((geom->H + 1 + 7) / 8), (unsigned)(ncopies * cbs));
    zero_border_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t *)planes, make_raster(*geom), ncopies, cbs, rows_total);
    return check_launch();
}

extern "C" int papr_unet_unpack(const void *src_planes, const papr_raster *geom, int cbs, float *dst, int64_t ld_pix, int C, void *stream)
{
    if (!src_planes || !dst || !raster_ok(geom) || C < 1 || cbs < 1 || C > 64 * cbs) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)geom->H * geom->W * cbs * 8;
    planes_to_nhwc_kernel<<<raster_grid(total, cbs), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src_planes, make_raster(*geom), cbs, dst, ld_pix, C);
    return check_launch();
}

extern "C" int papr_unet_spread(const papr_spread_args *a, void *stream)
{
    if (!a || !a->src || !raster_ok(&a->geom) || a->cbs < 1 || (!a->dst && !a->colsum)) return PAPR_ERR_INVALID_ARGUMENT;
    if (a->colsum && 64 * a->cbs > kMaxSumCols) return PAPR_ERR_INVALID_ARGUMENT;
    if (a->dst && (!raster_ok(&a->dst_geom) || (a->ncopies != 1 && a->ncopies != 3) || a->dst_geom.H != a->geom.H || a->dst_geom.W != a->geom.W))
        return PAPR_ERR_INVALID_ARGUMENT;
    if (a->pool_grad && (!a->pool_ref || !raster_ok(&a->pool_geom))) return PAPR_ERR_INVALID_ARGUMENT;
    if ((a->gamma == nullptr) != (a->beta == nullptr)) return PAPR_ERR_INVALID_ARGUMENT;
    SpreadParams p;
    p.src = (const uint8_t *)a->src; p.src_cb0 = a->src_cb0; p.add = (const uint8_t *)a->add; p.add_cb0 = a->add_cb0;
    p.mask = (const uint8_t *)a->mask; p.mask_cb0 = a->mask_cb0;
    p.pool_grad = (const uint8_t *)a->pool_grad; p.pool_ref = (const uint8_t *)a->pool_ref; p.pool_ref_cb0 = a->pool_ref_cb0;
    if (a->pool_grad) p.pg = make_raster(a->pool_geom); else p.pg = make_raster(a->geom);
    p.gamma = a->gamma; p.beta = a->beta;
    p.dst = (uint8_t *)a->dst; p.dst_cb0 = a->dst_cb0; p.ncopies = a->ncopies;
    p.dg = a->dst ? make_raster(a->dst_geom) : make_raster(a->geom);
    p.colsum = a->colsum; p.g = make_raster(a->geom); p.cbs = a->cbs;
    const int64_t total = (int64_t)a->geom.H * a->geom.W * a->cbs * 8;
    spread_kernel<<<raster_grid(total, a->cbs), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

extern "C" int papr_unet_pool(const void *src_planes, const papr_raster *geom, int src_cb0, void *dst_planes, const papr_raster *dst_geom,
                              int ncopies, int cbs, void *stream)
{
    if (!src_planes || !dst_planes || !raster_ok(geom) || !raster_ok(dst_geom) || cbs < 1 || (ncopies != 1 && ncopies != 3)) return PAPR_ERR_INVALID_ARGUMENT;
    if (dst_geom->H != geom->H / 2 || dst_geom->W != geom->W / 2) return PAPR_ERR_INVALID_ARGUMENT;
    const int64_t total = (int64_t)dst_geom->H * dst_geom->W * cbs * 8;
    pool_kernel<<<raster_grid(total, cbs), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src_planes, make_raster(*geom), src_cb0, (uint8_t *)dst_planes,
                                                                            make_raster(*dst_geom), ncopies, cbs);
    return check_launch();
}

extern "C" int papr_unet_convt_scatter(const void *src_planes, const papr_raster *low, int cout, const float *bias, void *dst_planes,
                                       const papr_raster *high, int dst_cb0, int ncopies, int pad_y, int pad_x, void *stream)
{
    if (!src_planes || !dst_planes || !bias || !raster_ok(low) || !raster_ok(high) || cout < 64 || cout % 64 || (ncopies != 1 && ncopies != 3) || pad_y < 0 || pad_x < 0)
        return PAPR_ERR_INVALID_ARGUMENT;
    const int cbs = cout / 64;
    const int64_t total = (int64_t)low->H * 2 * low->W * 2 * cbs * 8;
    convt_scatter_kernel<<<raster_grid(total, cbs), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src_planes, make_raster(*low), cout, bias, (uint8_t *)dst_planes,
                                                                                     make_raster(*high), dst_cb0, ncopies, pad_y, pad_x);
    return check_launch();
}

extern "C" int papr_unet_convt_gather(const void *src_planes, const papr_raster *high, int src_cb0, int cout, void *dst_planes,
                                      const papr_raster *low, int pad_y, int pad_x, float *colsum, void *stream)
{
    if (!src_planes || !dst_planes || !raster_ok(low) || !raster_ok(high) || cout < 64 || cout % 64 || cout > kMaxSumCols || pad_y < 0 || pad_x < 0)
        return PAPR_ERR_INVALID_ARGUMENT;
    const int cbs = cout / 64;
    const int64_t total = (int64_t)low->H * 2 * low->W * 2 * cbs * 8;
    convt_gather_kernel<<<raster_grid(total, cbs), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src_planes, make_raster(*high), src_cb0, cout, (uint8_t *)dst_planes,
                                                                                    make_raster(*low), pad_y, pad_x, colsum);
    return check_launch();
}
