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

#ifdef __cplusplus
}
#endif
#endif /* PAPR_B200_H */
