/*
 * papr_b200 -- C ABI of the B200-native PAPR hot path (libpapr_b200.so).
 *
 * The reference (zvict/papr) has no FFI of its own: the path is reached through the Python methods of
 * models.PAPR.  Each entry point below replaces one stage of those methods; papr_b200/ops.py binds them with
 * ctypes and papr_b200/model.py keeps the reference's Python surface on top.
 *
 * Conventions: every pointer is a DEVICE pointer into caller-owned memory unless stated otherwise; tensors are
 * dense row-major; `stream` is a cudaStream_t passed as void*; calls only enqueue work (no synchronisation, no
 * allocation, no global state).  Return value: 0 on success, a negative papr_status on failure (launch errors are
 * reported through cudaGetLastError at enqueue time).
 */
#ifndef PAPR_B200_H
#define PAPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum papr_status {
    PAPR_OK = 0,
    PAPR_ERR_INVALID_ARGUMENT = -1, /* bad size / unsupported K, widths, null pointer */
    PAPR_ERR_CUDA = -2,             /* a CUDA runtime call failed; see papr_last_cuda_error() */
    PAPR_ERR_UNSUPPORTED_DEVICE = -3 /* not an sm_100 device */
} papr_status;

/* ABI version (bumped when a signature changes) and human-readable error text. */
int papr_abi_version(void);
const char *papr_status_string(int status);
const char *papr_last_cuda_error(void);

/*
 * Stage a1 -- replaces PAPR._calculate_global_distances / _get_points (reference models/model.py:258-283, 312-333).
 * For every ray, the K points with the smallest perpendicular distance to the ray's line, computed with the
 * reference's exact FP32 rounding sequence; the full rays x points distance matrix is never materialised.
 *   rays_o  (n_views, 3) f32        one origin per view (model.py:273)
 *   rays_d  (n_views, rays_per_view, 3) f32   directions, used as given (not re-normalised, model.py:277)
 *   points  (P, 3) f32
 *   idx_out (n_views*rays_per_view, K) i32    ordered by (distance, point index) ascending
 * Requires 1 <= K <= 32 and K < P (the caller handles the K >= P bypass of model.py:326-327).
 */
int papr_select_topk(const float *rays_o, const float *rays_d, const float *points,
                     int64_t n_views, int64_t rays_per_view, int64_t P, int K, float eps,
                     int32_t *idx_out, void *stream);

/*
 * Tensor-core building blocks (stages a7/a8/a12).  Activations use the library's "tile-blocked" bf16 layout: a logical
 * [rows, cols] matrix (rows % 128 == 0, cols % 64 == 0) stored as [rows/128][cols/64] blocks of 16 KB, each block
 * 128 rows x 64 columns with the eight 16-byte chunks of a row XOR-swizzled by (row & 7) -- the shared-memory image
 * tcgen05.mma consumes, moved by 1-D TMA bulk copies.
 */

/* fp32 row-major (src_rows x src_cols, leading dim ld) -> tile-blocked bf16 (rows_pad x cols_pad, zero padded). */
int papr_blocked_from_f32(const float *src, int64_t src_rows, int src_cols, int64_t ld, void *dst,
                          int64_t rows_pad, int cols_pad, void *stream);
/* tile-blocked bf16 (cols_pad wide) -> fp32 row-major (dst_rows x dst_cols, leading dim ld). */
int papr_blocked_to_f32(const void *src, int cols_pad, float *dst, int64_t dst_rows, int dst_cols, int64_t ld,
                        void *stream);
/*
 * torch Linear weight (src_rows x src_cols fp32, leading dim ld; reference models/mlp.py:36, attn.py:204-205) -> bf16
 * weight image for papr_linear_bf16: element (n,k) = scale * W[n][k] (transpose=0) or scale * W[k][n] (transpose=1),
 * zero padded to N x K (N,K multiples of 16, <= 256).  Image size: ceil(K/64) * N * 128 bytes.
 */
int papr_pack_weight(const float *w, int64_t ld, int src_rows, int src_cols, int transpose, int N, int K,
                     float scale, void *image, void *stream);
/*
 * One Linear layer, replaces nn.Linear + activation inside models/mlp.py:53-58 (and w_k/w_q of attn.py:217-218):
 *   Y = act(X W^T + bias)    X: tile-blocked bf16 [rows, ceil(K/64)*64];  fp32 accumulation in TMEM.
 * Outputs (any combination): y_blocked tile-blocked bf16 [rows, ceil(N/64)*64]; y_f32 fp32 row-major (ldy);
 * sign_bits_out [rows, ceil(N/64)] u64, bit j of word g = (pre-activation column 64g+j > 0);
 * colsum[N] += column sums of the bf16 output (bias gradients).
 * Backward use (dgrad): X = dZ, weight image packed with transpose=1, sign_bits_in = the forward layer's sign bits:
 * output column j is multiplied by 1 (bit set) or `slope` (bit clear), i.e. by act'(.) of relu/leakyrelu.
 * act: 0 none, 1 relu/leakyrelu with negative slope `slope`.  rows % 128 == 0; N a multiple of 32, K a multiple of 16, both <= 256.
 */
int papr_linear_bf16(const void *x, const void *w_image, const float *bias, void *y_blocked, float *y_f32,
                     int64_t ldy, uint64_t *sign_bits_out, const uint64_t *sign_bits_in, float *colsum,
                     int64_t rows, int N, int K, int act, float slope, void *stream);
/*
 * Weight gradient of a Linear layer (autograd of models/mlp.py:53-58): C[a,b] += sum_rows A[row,a] * B[row,b] with
 * A, B tile-blocked bf16 (a_cols, b_cols wide; a_cols >= 128*ceil(a_valid/128)); C fp32 (leading dim ldc), updated
 * atomically; transpose_out stores C[b][a] instead.  dW = papr_wgrad(dZ, X).
 */
int papr_wgrad_bf16(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                    int a_valid, int b_valid, int transpose_out, int64_t rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PAPR_B200_H */
