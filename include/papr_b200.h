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
/* PAPR_OK if `device` (-1: the current one) is an sm_100 part, PAPR_ERR_UNSUPPORTED_DEVICE otherwise; every launch on
 * another architecture reports the same status. */
int papr_check_device(int device);
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
 * Stage a1 with spatial culling (same result as papr_select_topk, bit for bit).  The caller provides the points sorted
 * along a space-filling curve and padded to a multiple of 32 with far-away points (coordinates ~1e18):
 *   sorted_points (P_pad,3) f32; perm (P_pad) i32 = original index of each sorted point (-1 for padding);
 *   spheres (P_pad/32, 4) f32 = centre xyz + radius of each group of 32 consecutive sorted points (radius rounded up);
 *   spheres8 (ceil(P_pad/256), 4) f32 = the same around every 8 consecutive groups (two-level culling);
 *   pmax: DEVICE pointer to one float >= max |p| over the real points.
 * idx_out holds ORIGINAL point indices ordered by (distance, original index).
 */
int papr_select_topk_sorted(const float *rays_o, const float *rays_d, const float *sorted_points, const int32_t *perm,
                            const float *spheres, const float *spheres8, int64_t n_views, int64_t rays_per_view, int64_t P_pad, int64_t P,
                            int K, float eps, const float *pmax, int32_t *idx_out, void *stream);

/*
 * Stage a1 with screen-space culling (same result as papr_select_topk, bit for bit) -- the product path.  All rays of a
 * view share one origin, so a point can only be near a ray if its DIRECTION from the origin is near the ray's: the caller
 * bins every view's points on a G x G grid in gnomonic coordinates around the view's mean ray direction and sorts them by
 * cell (papr_b200/ops.py view_grids); a warp then visits the cells in rings around its rays and stops when a conservative
 * lower bound on the distance to everything unvisited exceeds its current thresholds.
 *   sorted_v (n_views*P, 4) f32 = (p - o, eps*|p - o|^2) in cell order, per view;  perm (n_views*P) i32 original indices;
 *   cells (n_views*G*G, 4) i32 = first / one-past-last position within the view, bits of the smallest |depth| of the cell, 0;
 *   view_params (n_views, 20) f32 = e1(3) e2(3) c(3) gmin(2) cell(2) 1/cell(2) min depth, max |v|^2, 0, 0, 0.
 */
int papr_select_topk_grid(const float *rays_o, const float *rays_d, const void *sorted_v, const int32_t *perm,
                          const int32_t *cells, const float *view_params, int64_t n_views, int64_t rays_per_view, int64_t P,
                          int G, int K, float eps, int32_t *idx_out, void *stream);

/*
 * Builds the inputs of papr_select_topk_grid on the device (six small launches, no host synchronisation): the camera frame
 * around each view's mean ray direction, the grid placement over the gnomonic extent of the view's rays and of the points
 * in front of the camera, the points binned and stored cell by cell (counting sort; the order inside a cell is
 * unspecified and does not influence the selection), each cell's smallest |depth|.  Outputs as papr_select_topk_grid
 * reads them: sorted_v (n_views*P, 4) f32, perm (n_views*P) i32, cells (n_views*G*G, 4) i32, view_params (n_views, 20)
 * f32.  workspace: at least papr_select_grid_workspace_bytes(n_views, P, G) bytes of device memory (contents undefined).
 * Rebuilt whenever the points or the cameras change, i.e. before every selection (model.py:258-283 has no such
 * structure: it materialises all rays x points distances).
 */
int64_t papr_select_grid_workspace_bytes(int64_t n_views, int64_t P, int G);
int papr_select_grid_build(const float *rays_o, const float *rays_d, const float *points, int64_t n_views, int64_t rays_per_view,
                           int64_t P, int G, float eps, void *sorted_v, int32_t *perm, int32_t *cells, float *view_params,
                           void *workspace, int64_t workspace_bytes, void *stream);

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
 * sign_bits_out u64 [rows/128][ceil(N/64)][128] (tile, 64-column group, row within the tile), bit j of word g = 1 when the pre-activation of column 64g+j has its sign bit clear (> 0, or exactly +0);
 * colsum[N] += column sums of the bf16 output (bias gradients).
 * Backward use (dgrad): X = dZ, weight image packed with transpose=1, sign_bits_in = the forward layer's sign bits:
 * output column j is multiplied by 1 (bit set) or `slope` (bit clear), i.e. by act'(.) of relu/leakyrelu.
 * addend_f32 (rows, ld_addend) fp32, optional: added to X W^T before bias/activation/mask -- the other half of a layer
 * whose input is a concatenation (mlp.py:54-55 skip_layers: [h, inp] W^T = h W1^T + inp W2^T), or an extra gradient term.
 * act: 0 none, 1 relu/leakyrelu with negative slope `slope`.  rows % 128 == 0; N a multiple of 32, K a multiple of 16, both <= 256.
 */
int papr_linear_bf16(const void *x, const void *w_image, const float *bias, void *y_blocked, float *y_f32,
                     int64_t ldy, uint64_t *sign_bits_out, const uint64_t *sign_bits_in, float *colsum,
                     const float *addend_f32, int64_t ld_addend, int64_t rows, int N, int K, int act, float slope,
                     void *stream);
/*
 * Weight gradient of a Linear layer (autograd of models/mlp.py:53-58): C[a,b] += sum_rows A[row,a] * B[row,b] with
 * A, B tile-blocked bf16 (a_cols, b_cols wide; a_cols >= 128*ceil(a_valid/128)); C fp32 (leading dim ldc), updated
 * atomically; transpose_out stores C[b][a] instead.  dW = papr_wgrad(dZ, X).
 */
int papr_wgrad_bf16(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                    int a_valid, int b_valid, int transpose_out, int64_t rows, void *stream);
/* Same, on at most max_ctas SMs (0 = all), so that it can share the GPU with a papr_stack_bf16_ex launch on another
 * stream: in the backward pass the dgrad of one slice of rows (HBM-write bound) overlaps the weight gradients of the
 * previous slice (HBM-read bound). */
int papr_wgrad_bf16_ex(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                       int a_valid, int b_valid, int transpose_out, int64_t rows, int max_ctas, void *stream);
/* The same, plus the bias gradient on the side: a_colsum[a_valid] (fp32, may be NULL) += the column sums of A over all
 * rows -- with A = dZ that is autograd's `grad_bias = dZ.sum(0)` of the Linear layer (reference models/mlp.py:53), read
 * from the tiles the kernel streams anyway instead of in the dgrad epilogue. */
int papr_wgrad_bias_bf16(const void *a_blocked, int a_cols, const void *b_blocked, int b_cols, float *c, int64_t ldc,
                         int a_valid, int b_valid, int transpose_out, int64_t rows, int max_ctas, float *a_colsum, void *stream);

/*
 * Per-ray CUDA-core stages.  Rows are (ray, candidate) pairs, row = ray*K + k; M = R*K rows, padded to 128.
 * L = positional-encoding order (same for every feature; embed_type 1), S = 1+2L, dk = 9S, dv = 6S + F.
 */

/*
 * Stages a2-a6 -- replaces the gathers of PAPR._get_points/_get_kqv, PAPR._calculate_distances, posenc and the key
 * stack's input LayerNorm (reference models/model.py:285-310,330,396-437; models/utils.py:232-257; attn.py:39-42,172-191).
 *   kin (M_pad, dk_pad) tile-blocked bf16 = LayerNorm([PE(point), PE(proj), PE(D)]) with affine ln_a, ln_b (dk)
 *   vin (M_pad, dv_pad) tile-blocked bf16 = [PE(proj), PE(D), pc_feats[idx]]
 * kin_f32 (M,dk) / vin_f32 (M,dv): optional fp32 copies of the same values (parity taps), may be null.
 */
int papr_attn_prologue_fwd(const float *rays_o, const float *rays_d, const float *points, const float *feats,
                           const int32_t *idx, const float *ln_a, const float *ln_b, int64_t R,
                           int64_t rays_per_view, int K, int L, int F, float eps, void *kin, int dk_pad,
                           void *vin, int dv_pad, float *kin_f32, float *vin_f32, void *stream);
/*
 * Backward of the above (autograd of the same reference lines; the raw-position key features are detached,
 * model.py:405).  Gradients are ACCUMULATED (atomically) into g_points (P,3), g_feats (P,F), g_ln_a/g_ln_b (dk).
 * dkin/dvin: tile-blocked bf16 gradients, or fp32 row-major taps dkin_f32 (M,dk) / dvin_f32 (M,dv) if non-null.
 */
int papr_attn_prologue_bwd(const float *rays_o, const float *rays_d, const float *points, const int32_t *idx,
                           const float *ln_a, int64_t R, int64_t rays_per_view, int K, int L, int F, float eps,
                           const void *dkin, int dk_pad, const void *dvin, int dv_pad, const float *dkin_f32,
                           const float *dvin_f32, float *g_points, float *g_feats, float *g_ln_a,
                           float *g_ln_b, void *stream);
/*
 * Stages a6 (key output LayerNorm), a8 and a9 -- replaces FeedForward.outnorm + AttentionLayer + the blend of
 * PAPR.forward/evaluate (reference attn.py:39-42,117,212-226,53-54; models/model.py:519-534).
 * Uses score = (W_k^T q'/sqrt(d)) . LN(h5) + q'.b_k/sqrt(d): the caller passes ua (R,256) = (W_k^T q'/sqrt d) * a_2 and
 * cprime (R) = (W_k^T q'/sqrt d) . b_2 + q'.b_k/sqrt d, with a_2,b_2 the key outnorm affine terms.
 *   h5 (M_pad,256) tile-blocked bf16 key-stack output (or fp32 tap h5_f32 (M,256)); v (M_pad, ldv) fp32 value-stack output
 *   fused (R,C); attn (R,K+1) softmax incl. background, un-renormalised (model.py:481/529);
 *   sc (M) activated scores before influence; stats (M,4) = LayerNorm mean, 1/(std+eps), ua . z, 0 (saved for backward)
 */
int papr_score_blend_fwd(const void *h5, const float *h5_f32, const float *ua, const float *cprime,
                         const float *influ, const int32_t *idx, const float *v, int64_t ldv, int64_t R, int K,
                         int C, int score_relu, int normalize, float bkg_score, float eps, float *fused,
                         float *attn, float *sc, float *stats, void *stream);
/*
 * Backward of the blend (model.py:524-534): from d_fused (R,C) and optional d_attn (R,K+1) to
 *   dv_blocked (M_pad,64) tile-blocked bf16 (C <= 64), d_score (M) gradient of the pre-activation score,
 *   g_influ (P) += , g_bias_v (C) += column sums of dv.
 */
int papr_blend_bwd(const float *d_fused, const float *d_attn, const float *attn, const float *sc,
                   const float *influ, const int32_t *idx, const float *v, int64_t ldv, int64_t R, int K, int C,
                   int score_relu, int normalize, void *dv_blocked, float *d_score, float *g_influ,
                   float *g_bias_v, void *stream);
/*
 * Backward of the folded LayerNorm + scaled dot: d_score (M) -> dh5 (M_pad,256) tile-blocked bf16 (+ optional fp32 tap),
 * zsum (R,256) = sum_k d_score * normalised h5 (= d ua), dssum (R) = sum_k d_score (= d cprime),
 * g_bias5 (256) += column sums of dh5; may be NULL (then papr_wgrad_bias_bf16 of the last key layer produces them and the
 * bf16-only call walks the rows block by block, 512 contiguous bytes per warp instruction).
 */
int papr_key_score_bwd(const float *d_score, const void *h5, const float *h5_f32, const float *stats,
                       const float *ua, int64_t R, int K, float eps, void *dh5_blocked, float *dh5_f32,
                       float *zsum, float *dssum, float *g_bias5, void *stream);

/*
 * A whole MLP stack per launch (fused papr_linear_bf16 chain, reference models/mlp.py:47-59): the 128-row activation
 * tiles stay in shared memory between layers; CTA pairs (tcgen05 cta_group::2) split every weight chunk.  Layer l maps
 * its K_l inputs (K_0 = K0, K_l = N_{l-1}) to N_l outputs; hidden layers must have N = 256, the last one 32..256.
 * Per layer, the same options as papr_linear_bf16: bias/act (forward), sign_bits_out, sign_bits_in + colsum (dgrad),
 * out_blocked (tile-blocked copy of the layer's output: the stash the weight-gradient kernel needs, or the final
 * result) and out_f32.  `layers` is a HOST array.
 */
typedef struct papr_stack_layer {
    const void *w_image;          /* papr_pack_weight image, N x K_l */
    const float *bias;            /* [N] or NULL */
    void *out_blocked;            /* tile-blocked bf16 [rows, ceil(N/64)*64] or NULL */
    float *out_f32;               /* fp32 row-major [rows, ld_f32] or NULL */
    int64_t ld_f32;
    uint64_t *sign_bits_out;      /* [rows/128][ceil(N/64)][128] or NULL */
    const uint64_t *sign_bits_in; /* [rows/128][ceil(N/64)][128] or NULL */
    float *colsum;                /* [N] accumulated, or NULL */
    int32_t N;
    int32_t act;                  /* 0 none, 1 relu/leakyrelu(slope) */
    int32_t w_replicas;           /* >= 1: identical copies of the image, w_replica_stride bytes apart; CTA pairs read different
                                     copies so that 148 SMs streaming the same chunk do not hammer the same L2 slices */
    int32_t _pad;
    int64_t w_replica_stride;
} papr_stack_layer;

int papr_stack_bf16(const void *x, int K0, const papr_stack_layer *layers, int n_layers, int64_t rows, float slope,
                    void *stream);
/* Same, on at most max_ctas SMs (0 = all; rounded down to whole CTA pairs). */
int papr_stack_bf16_ex(const void *x, int K0, const papr_stack_layer *layers, int n_layers, int64_t rows, float slope,
                       int max_ctas, void *stream);

/*
 * Stage a12, fused -- the whole backward of one MLP stack in one launch: the dgrad of every layer (as papr_stack_bf16 run
 * on `dgrad_layers`, which are in dgrad order: the transposed image of the LAST forward layer first; sign_bits_in and
 * colsum as there) plus every weight gradient gw[i] += dZ_i^T X_i.  Replaces autograd's Linear backward over the
 * reference models/mlp.py:47-59 loop.  The per-layer dZ tiles are handed from the dgrad CTAs to the weight-gradient CTAs
 * through `workspace` (kept in L2) instead of being written to and re-read from HBM; only the LAST dgrad layer's
 * out_blocked (the gradient of the stack input) is written -- out_blocked of the others is ignored.
 *   dz: tile-blocked bf16 gradient of the stack output, K0 = its valid width (multiple of 16); hidden widths are 256.
 *   wgrad_layers[i], i in FORWARD order: x_blocked = the input of forward layer i (x_cols wide), gw (n_out, ldw) fp32.
 *   producer_ctas: SMs given to the dgrad side (0 = default split); workspace: papr_stack_bwd_workspace_bytes() bytes,
 *   1 KB aligned, reusable by the next call on the same stream.  ReLU / no activation only (slope 0).
 */
typedef struct {
    const void *x_blocked;
    int x_cols;
    float *gw;
    int64_t ldw;
    int n_out, n_in;
} papr_wgrad_layer;

int64_t papr_stack_bwd_workspace_bytes(void);
int papr_stack_bwd_fused(const void *dz, int K0, const papr_stack_layer *dgrad_layers, const papr_wgrad_layer *wgrad_layers,
                         int n_layers, int64_t rows, int producer_ctas, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * Per-ray query tail -- replaces the query stack's output LayerNorm, w_q, and the part of AttentionLayer that depends
 * only on the ray (reference attn.py:39-42, 117, 217-218, 53-54).  Everything after the normalisation is linear, so the
 * host folds the LayerNorm affine terms, w_q, w_k and the key out-norm into one 256x256 matrix A, a bias c0, a vector
 * w_c and a scalar c_const:  ua = A z + c0 (papr_linear_bf16),  c' = w_c . z + c_const (here).
 *   fwd: q5 (R,256) f32 -> z tile-blocked bf16 (R_pad,256) = (q5-mean)/(std+eps), stats (R,2), cprime (R)
 *   bwd: dz (R,256; ld_dz) = gradient w.r.t. z through A, dc (R) -> dq5 (R,256); g_wc (256) +=, g_cconst (1) +=
 */
int papr_query_tail_fwd(const float *q5, const float *w_c, float c_const, float eps, int64_t R, void *z_blocked,
                        float *stats, float *cprime, void *stream);
int papr_query_tail_bwd(const float *q5, const float *stats, const float *w_c, const float *dz, int64_t ld_dz,
                        const float *dc, float eps, int64_t R, float *dq5, float *g_wc, float *g_cconst, void *stream);

/*
 * Query-side prologue: the input of the query stack for R rays (reference models/attn.py:148-176 `embed_q` input:
 * utils.py:232-242 positional encoding of the ray direction, embed_type 1, then the in-norm of attn.py:30-42 with the
 * unbiased std), replacing ~10 elementwise / reduction launches over an (R, 3(1+2L)) tensor by one.
 *   fwd: rays_d (R,3) f32, a2 / b2 (3(1+2L)) -> q (R, 3(1+2L)) f32 = a2 * (pe - mean) / (std + eps) + b2
 *   bwd: dq (R, 3(1+2L)) f32 -> g_a2 += sum_r dq * z, g_b2 += sum_r dq   (the ray direction is data: no gradient)
 * L must be 4 or 6 (the PE orders of the shipped configs); PAPR_ERR_INVALID_ARGUMENT otherwise.
 */
int papr_query_prologue_fwd(const float *rays_d, const float *a2, const float *b2, int64_t R, int L, float eps, float *q,
                            void *stream);
int papr_query_prologue_bwd(const float *rays_d, const float *dq, int64_t R, int L, float eps, float *g_a2, float *g_b2,
                            void *stream);


/*
 * SURVEY section 8(f1) -- replaces the optimiser half of PAPR.step (reference models/model.py:439-446: one
 * torch.optim.Adam.step per parameter group) with ONE multi-tensor launch.  Gradients and both Adam moments live in
 * flat fp32 buffers of `total` elements; tensor t occupies [offsets[t], offsets[t+1]) of them, its parameter storage is
 * param_ptrs[t] (contiguous fp32) and it belongs to group group_of[t].  offsets / param_ptrs / group_of are DEVICE
 * arrays, `groups` is a HOST array (n_groups <= 8).  Update rule = torch.optim.Adam (L2 weight decay, bias correction
 * from `step`, which counts from 1); grad_scale multiplies every gradient first (1/world after a summed all-reduce).
 * A group with enabled == 0 or step < 1 is skipped.
 */
typedef struct papr_adam_group {
    float lr, beta1, beta2, eps, weight_decay;
    int32_t step;
    int32_t enabled;
} papr_adam_group;

int papr_adam_step(const int64_t *offsets, float *const *param_ptrs, const int32_t *group_of, int n_tensors, int64_t total,
                   const float *grad, float *exp_avg, float *exp_avg_sq, const papr_adam_group *groups, int n_groups,
                   float grad_scale, void *stream);

/*
 * Batched papr_pack_weight: one launch rebuilds every weight image listed in the DEVICE array `descs` (after an
 * optimiser step: all forward and transposed images of the key / value / query stacks).  Each image may be written
 * `replicas` times, rep_stride bytes apart (papr_stack_layer.w_replicas).
 */
typedef struct papr_pack_desc {
    const float *w;        /* fp32 source matrix, leading dimension ld */
    int64_t ld;
    int32_t rows, cols;    /* source shape */
    int32_t transpose;     /* element (n,k) = W[k][n] instead of W[n][k] */
    int32_t N, K;          /* padded image shape (multiples of 16, <= 256) */
    int32_t replicas;
    float scale;
    int32_t _pad;
    int64_t rep_stride;
    void *image;
} papr_pack_desc;

int papr_pack_weight_batch(const papr_pack_desc *descs, int n_descs, void *stream);

/*
 * Stage a14 -- point growing / pruning bookkeeping on the device (reference models/utils.py:9-109 add_points_knn uses a
 * host scipy KDTree; models/model.py:335-358 prune_points uses boolean-mask indexing).
 * papr_knn: for each of Q queries (Q,3) the k <= 32 nearest of P points (P,3), exact float64 Euclidean distances as the
 * KDTree computes them, ascending, ties by smaller point index: dist_out (Q,k) f64, idx_out (Q,k) i32.
 * papr_prune_compact: keeps, in order, the rows whose influence score is > thresh (keep_less = 0, the reference's
 * prune_type "<") or < thresh (keep_less = 1); writes compacted points / influ / feats (feats may be NULL) and the
 * number of kept rows to the DEVICE scalar n_kept.  scratch: ceil(P/256) int32.
 */
int papr_knn(const float *points, int64_t P, const float *queries, int64_t Q, int k, double *dist_out, int32_t *idx_out,
             void *stream);
int papr_prune_compact(const float *points, const float *influ, const float *feats, int64_t P, int F, float thresh,
                       int keep_less, float *out_points, float *out_influ, float *out_feats, int32_t *scratch,
                       int64_t *n_kept, void *stream);

/*
 * SURVEY section 8(f3) -- the step before the path: pinhole rays of the pixels [h0,h0+h) x [w0,w0+w) of an H x W view
 * (reference dataset/utils.py:81-96 get_rays, dataset/dataset.py:19-25 origin scaling) generated on the device, so that
 * a training step uploads a 4x4 pose instead of 12 bytes per ray.  c2w (n_views,4,4) f32 -> rays_o (n_views,3) =
 * coord_scale * c2w[:3,3], rays_d (n_views,h,w,3) unit norm.
 */
int papr_generate_rays(const float *c2w, int64_t n_views, int H, int W, float focal_x, float focal_y, int h0, int w0, int h,
                       int w, float coord_scale, float *rays_o, float *rays_d, void *stream);

/*
 * Stage a10 -- the SmallUNet decode (reference models/unet.py:196-258, built by models/renderer.py:21-34) on the library's
 * own tensor-core kernels instead of cuDNN.
 *
 * "Pixel planes": a feature map of one image is a zero-padded raster (padded width Wp, a multiple of 8; pixel (y,x) ->
 * raster index (y+1)*Wp + (x+1); row0 guard rows of zeros before and after); every block of 64 channels is a plane of
 * 128-byte rows (64 bf16), chunks XOR-swizzled by (row & 7), so any 128 rows that start at a multiple of 8 are a
 * tcgen05.mma operand tile fetched by one 1-D TMA bulk copy.  Maps that feed a 3x3 convolution are stored three times,
 * shifted by dx = -1, 0, +1 (copy_dx[p] = X[p + dx], copy_bytes apart), so that all nine taps are row offsets that keep
 * the swizzle phase.  papr_raster describes one map; planes of a copy are plane_bytes apart.
 */
typedef struct papr_raster {
    int32_t H, W, Wp, _pad;
    int64_t row0;          /* guard rows before pixel 0 (>= Wp + 1) */
    int64_t plane_bytes;   /* bytes of one 64-channel plane (rows * 128) */
    int64_t copy_bytes;    /* distance between the dx copies (0 for single-copy maps) */
} papr_raster;

/*
 * Convolution as implicit GEMM (unet.py:14-26 Conv2d 3x3 padding 1, :171-179 1x1; their data gradients with sign = -1 and
 * the weight image packed as W'[ci][tap*Cout + co]): for each tile of 128 raster rows
 *     Y[p, n] = act( sum_{tap, cb} in(copy dx(tap), plane cb, row p + sign*dy(tap)*Wp) . w_image[tap*cbs + cb][n] + bias[n] )
 * in_planes points at plane 0 of copy dx = -1 (ntaps = 9) or of the only copy (ntaps = 1); w_image holds ntaps*cbs blocks of
 * N x 128 B (papr_pack_weight_batch layout over K = ntaps*cbs*64).  Output: unshifted planes (out_planes, ceil(N/64) of
 * them, written as whole 128-row tiles INCLUDING padding rows -- papr_unet_spread re-establishes the zero border) and / or
 * fp32 rows out_f32[tile*128 + r][ld_f32].  addend_f32 (same row indexing, ld_addend) is added to the accumulator before bias /
 * activation: the partial sums of the fp32 parity mode (three-way bf16 split of both operands, six launches).
 * 32 <= N <= 256, N % 32 == 0.
 */
int papr_conv_bf16(const void *in_planes, int64_t in_copy_bytes, int64_t in_plane_bytes, int64_t in_row0, int cbs, int ntaps,
                   int Wp, int sign, const void *w_image, const float *bias, int N, int act, float slope, void *out_planes,
                   int64_t out_plane_bytes, int64_t out_row0, float *out_f32, int64_t ld_f32, const float *addend_f32,
                   int64_t ld_addend, int64_t n_tiles, void *stream);
/*
 * Weight gradient of a convolution (autograd of the same lines), all taps in one launch:
 *     C[tap][a][b] += sum_rows A[row][a] * B(copy dx(tap))[row + dy(tap)*Wp][b]        over `rows` raster rows (multiple of 64)
 * A = d output (unshifted planes, a_valid <= 256 channels, ceil(a_valid/128)*2 planes present), B = the layer input
 * (b_planes: plane 0 of copy dx = -1 at raster row 0 when ntaps = 9; of the only copy when ntaps = 1), b_valid <= 256.
 * C: fp32, leading dimension ldc, taps c_tap_stride floats apart, accumulated atomically.
 */
int papr_conv_wgrad_bf16(const void *a_planes, int64_t a_plane_bytes, int a_valid, const void *b_planes, int64_t b_plane_bytes,
                         int64_t b_copy_bytes, int b_valid, int ntaps, int Wp, float *c, int64_t ldc, int64_t c_tap_stride,
                         int64_t rows, void *stream);

/* fp32 (H,W,C) rows of ld_pix floats [FiLM x*gamma+beta, unet.py:213-217] -> planes (1 or 3 copies, cbs planes each). */
int papr_unet_pack_input(const float *src, int64_t ld_pix, int C, const float *gamma, const float *beta, void *dst_planes,
                         const papr_raster *geom, int ncopies, int cbs, void *stream);
/* planes (unshifted copy) -> fp32 (H,W,C). */
int papr_unet_unpack(const void *src_planes, const papr_raster *geom, int cbs, float *dst, int64_t ld_pix, int C, void *stream);
/*
 * Between two convolutions: dst(interior) = film( relu_mask( src + add + maxpool_backward(pool_grad) ) ), 1 or 3 shifted
 * copies, optional per-channel sums of what was stored (bias gradients).  All sources are unshifted planes on `geom`;
 * pool_grad lives on pool_geom (half resolution) and pool_ref is the map that was pooled (unet.py:45-53 MaxPool2d(2):
 * the gradient goes to the first maximum of each 2x2 window).
 */
typedef struct papr_spread_args {
    const void *src, *add, *mask, *pool_grad, *pool_ref;
    const float *gamma, *beta;
    void *dst;
    float *colsum;
    papr_raster geom, pool_geom, dst_geom;
    int32_t src_cb0, add_cb0, mask_cb0, pool_ref_cb0, dst_cb0, ncopies, cbs, _pad;
} papr_spread_args;
int papr_unet_spread(const papr_spread_args *args, void *stream);
/* MaxPool2d(2) (unet.py:45-53): unshifted planes at H x W -> planes at floor(H/2) x floor(W/2). */
int papr_unet_pool(const void *src_planes, const papr_raster *geom, int src_cb0, void *dst_planes, const papr_raster *dst_geom,
                   int ncopies, int cbs, void *stream);
/* Zeroes the part of a freshly allocated map that no raster kernel writes -- the guard rows and the frame of padding pixels
 * of each (shifted) copy -- in place of a memset of the whole buffer; the producer of the map writes every interior pixel.
 * (The zero padding is what nn.Conv2d(padding=1) of reference models/unet.py:20-33 adds.) */
int papr_unet_zero_border(void *planes, const papr_raster *geom, int ncopies, int cbs, void *stream);
/*
 * ConvTranspose2d(kernel 2, stride 2) (unet.py:60-76) = a 1x1 GEMM to 4*cout channels ordered (a, b, co) followed by this
 * pixel shuffle (+ bias, + the F.pad offset of unet.py:70-74); the gather is its inverse for the backward pass (colsum: the
 * bias gradient).
 */
int papr_unet_convt_scatter(const void *src_planes, const papr_raster *low, int cout, const float *bias, void *dst_planes,
                            const papr_raster *high, int dst_cb0, int ncopies, int pad_y, int pad_x, void *stream);
int papr_unet_convt_gather(const void *src_planes, const papr_raster *high, int src_cb0, int cout, void *dst_planes,
                           const papr_raster *low, int pad_y, int pad_x, float *colsum, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PAPR_B200_H */
