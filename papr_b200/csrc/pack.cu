// Layout helpers for the tensor-core path: weight images and fp32 <-> tile-blocked bf16 conversion.
#include "tc_common.cuh"

namespace papr {

// image[kb][n][128 B]: K-major SWIZZLE_128B block per 64 input columns; element (n,k) = scale * W[n][k] (or W[k][n]).
__global__ void pack_weight_kernel(const float *__restrict__ w, int ld, int src_rows, int src_cols, int transpose,
                                   int N, int K, float scale, __nv_bfloat16 *__restrict__ image)
{
    const int kblk = (K + 63) / 64;
    const int total = kblk * N * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kb = i / (N * 64);
        const int n = (i / 64) % N;
        const int kk = i % 64;
        const int k = kb * 64 + kk;
        float v = 0.f;
        if (!transpose) { if (n < src_rows && k < src_cols) v = w[(size_t)n * ld + k]; }
        else            { if (k < src_rows && n < src_cols) v = w[(size_t)k * ld + n]; }
        const size_t off = (size_t)kb * N * 128 + (size_t)n * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4)) + (kk & 7) * 2;
        image[off / 2] = __float2bfloat16(v * scale);
    }
}

__global__ void blocked_from_f32_kernel(const float *__restrict__ src, int64_t src_rows, int src_cols, int64_t ld,
                                        uint8_t *__restrict__ dst, int64_t rows_pad, int nblk)
{
    const int64_t total = rows_pad * nblk * 8;   // 16-byte chunks
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (nblk * 8);
        const int chunk = (int)(i % (nblk * 8));
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = chunk * 8 + j;
            f[j] = (row < src_rows && col < src_cols) ? src[row * ld + col] : 0.f;
        }
        uint4 q = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        *reinterpret_cast<uint4 *>(dst + blocked_chunk_offset(row, chunk, nblk)) = q;
    }
}

__global__ void blocked_to_f32_kernel(const uint8_t *__restrict__ src, int nblk, float *__restrict__ dst,
                                      int64_t dst_rows, int dst_cols, int64_t ld)
{
    const int nchunk = (dst_cols + 7) / 8;
    const int64_t total = dst_rows * nchunk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / nchunk;
        const int chunk = (int)(i % nchunk);
        const uint4 q = *reinterpret_cast<const uint4 *>(src + blocked_chunk_offset(row, chunk, nblk));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = chunk * 8 + j;
            if (col < dst_cols) dst[row * ld + col] = (j & 1) ? bf16_hi(w[j >> 1]) : bf16_lo(w[j >> 1]);
        }
    }
}

}  // namespace papr

extern "C" int papr_pack_weight(const float *w, int64_t ld, int src_rows, int src_cols, int transpose, int N, int K,
                                float scale, void *image, void *stream)
{
    using namespace papr;
    if (!w || !image || N < 16 || N > 256 || N % 16 || K < 16 || K > 256 || K % 16) return PAPR_ERR_INVALID_ARGUMENT;
    const int total = ((K + 63) / 64) * N * 64;
    pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (int)ld, src_rows, src_cols, transpose, N, K, scale,
                                                                             (__nv_bfloat16 *)image);
    return check_launch();
}

extern "C" int papr_blocked_from_f32(const float *src, int64_t src_rows, int src_cols, int64_t ld, void *dst,
                                     int64_t rows_pad, int cols_pad, void *stream)
{
    using namespace papr;
    if (!src || !dst || rows_pad % kTileRows || cols_pad % 64 || rows_pad < src_rows || cols_pad < src_cols) return PAPR_ERR_INVALID_ARGUMENT;
    if (rows_pad == 0) return PAPR_OK;
    const int64_t total = rows_pad * (cols_pad / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    blocked_from_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_rows, src_cols, ld, (uint8_t *)dst, rows_pad, cols_pad / 64);
    return check_launch();
}

extern "C" int papr_blocked_to_f32(const void *src, int cols_pad, float *dst, int64_t dst_rows, int dst_cols, int64_t ld,
                                   void *stream)
{
    using namespace papr;
    if (!src || !dst || cols_pad % 64 || cols_pad < dst_cols) return PAPR_ERR_INVALID_ARGUMENT;
    if (dst_rows == 0) return PAPR_OK;
    const int64_t total = dst_rows * ((dst_cols + 7) / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    blocked_to_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src, cols_pad / 64, dst, dst_rows, dst_cols, ld);
    return check_launch();
}
