/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the ray->point distance + top-K stage.
 * Never linked into or called by the product library; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it.
 *
 * Restates reference models/model.py:258-283 (PAPR._calculate_global_distances) with the
 * exact FP32 rounding sequence that PyTorch's CPU kernels use for it (contiguous tensors):
 *   v    = p - o                                   model.py:276
 *   s    = (v.x*d.x + v.y*d.y) + v.z*d.z           model.py:277  torch.sum over a size-3 dim = (a+b)+c,
 *                                                  products rounded separately (no FMA contraction)
 *   den  = ((d.x*d.x + d.y*d.y) + d.z*d.z) + eps   model.py:277
 *   t    = s / den                                 IEEE division
 *   proj = d * t ; D = v - proj                    model.py:277-278
 *   key  = fma(D.z,D.z, fma(D.y,D.y, D.x*D.x))     model.py:279  torch.norm(dim=-1) on CPU accumulates with FMA
 *   dist = (float)sqrt((double)key)
 * and model.py:281 (topk, largest=False): the K smallest, here made deterministic by ordering on
 * (key, point index).  Pinned against the real reference by oracle/make_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float pair_key(const float *o, const float *d, float den, const float *p)
{
    float vx = p[0] - o[0], vy = p[1] - o[1], vz = p[2] - o[2];
    float s = (vx * d[0] + vy * d[1]) + vz * d[2];
    float t = s / den;
    float Dx = vx - d[0] * t, Dy = vy - d[1] * t, Dz = vz - d[2] * t;
    return fmaf(Dz, Dz, fmaf(Dy, Dy, Dx * Dx));
}

static inline float ray_den(const float *d, float eps)
{
    return ((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) + eps;
}

/* dist[r*P + p] for every pair (small problems only). rays_o is per ray (R,3). */
void papr_oracle_distances(const float *rays_o, const float *rays_d, const float *points,
                           int64_t R, int64_t P, float eps, float *dist)
{
    for (int64_t r = 0; r < R; ++r) {
        const float *o = rays_o + 3 * r, *d = rays_d + 3 * r;
        float den = ray_den(d, eps);
        for (int64_t p = 0; p < P; ++p)
            dist[r * P + p] = (float)sqrt((double)pair_key(o, d, den, points + 3 * p));
    }
}

/* idx[r*K + j]: the K nearest points of ray r ordered by (key, index); kth[r] = dist of the K-th. */
void papr_oracle_topk(const float *rays_o, const float *rays_d, const float *points,
                      int64_t R, int64_t P, int K, float eps, int32_t *idx, float *kth)
{
    float *bk = (float *)malloc(sizeof(float) * K);
    int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * K);
    for (int64_t r = 0; r < R; ++r) {
        const float *o = rays_o + 3 * r, *d = rays_d + 3 * r;
        float den = ray_den(d, eps);
        int n = 0;
        for (int64_t p = 0; p < P; ++p) {
            float key = pair_key(o, d, den, points + 3 * p);
            if (n == K && !(key < bk[K - 1])) continue;   /* ties keep the smaller index */
            int j = (n < K) ? n++ : K - 1;
            while (j > 0 && bk[j - 1] > key) { bk[j] = bk[j - 1]; bi[j] = bi[j - 1]; --j; }
            bk[j] = key; bi[j] = (int32_t)p;
        }
        memcpy(idx + r * K, bi, sizeof(int32_t) * n);
        for (int j = n; j < K; ++j) idx[r * K + j] = -1;
        kth[r] = n ? (float)sqrt((double)bk[n - 1]) : 0.0f;
    }
    free(bk); free(bi);
}

/*
 * Division check used by tests/test_division.py: the CUDA select kernel replaces the IEEE
 * divide s/den by a Markstein-corrected multiply with r = RN(1/den).  This routine applies the
 * same sequence with fmaf so the test can compare it against the hardware division here.
 */
float papr_oracle_markstein_div(float s, float den)
{
    float r = 1.0f / den;
    float q0 = s * r;
    float e0 = fmaf(-q0, den, s);
    float q1 = fmaf(e0, r, q0);
    float e1 = fmaf(-q1, den, s);
    return fmaf(e1, r, q1);
}

int64_t papr_oracle_markstein_mismatches(const float *s, const float *den, int64_t n)
{
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        float a = s[i] / den[i], b = papr_oracle_markstein_div(s[i], den[i]);
        if (memcmp(&a, &b, 4) != 0 && !(a != a && b != b)) ++bad;
    }
    return bad;
}
