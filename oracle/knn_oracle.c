/*
 * knn_oracle.c — TEST INFRASTRUCTURE ONLY.  Plain-C restatement of SimpleKNN::knn
 * (KNN/simple_knn.cu:46-222, KNN = fs3dgs_benchmark/gaussian-splatting/submodules/simple-knn/): bbox with {0,0,0}
 * init, 30-bit Morton codes, stable sort, 1024-point boxes, exact 3-NN with box pruning, mean of the 3 smallest
 * squared distances.  Pinned against the reference CUDA build's output (tests/golden/knn_*.npz).
 * A brute-force variant is included to check the pruned search itself on small inputs.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BOX_SIZE 1024

static uint32_t prep_morton(uint32_t x) { /* simple_knn.cu:46-53 */
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
static uint32_t f2u_rz(float v) { /* PTX cvt.rzi.u32.f32 */
    if (v != v || v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
static void update3(const float* p, const float* q, float* knn) { /* simple_knn.cu:132-146 */
    const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    float dist = dx * dx + dy * dy + dz * dz;
    for (int j = 0; j < 3; j++)
        if (knn[j] > dist) { const float t = knn[j]; knn[j] = dist; dist = t; }
}
static float dist_box_point(const float* mn, const float* mx, const float* p) { /* simple_knn.cu:120-130 */
    float d[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++)
        if (p[k] < mn[k] || p[k] > mx[k]) d[k] = fminf(fabsf(p[k] - mn[k]), fabsf(p[k] - mx[k]));
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

void oracle_knn_mean_dist2(int P, const float* pts, float* out) {
    if (P <= 0) return;
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0}; /* init {0,0,0}: simple_knn.cu:192 (quirk Q6) */
    for (int i = 0; i < P; i++)
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], pts[3 * i + k]); mx[k] = fmaxf(mx[k], pts[3 * i + k]); }
    uint32_t* code = (uint32_t*)malloc(sizeof(uint32_t) * P);
    uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * P);
    uint32_t* code2 = (uint32_t*)malloc(sizeof(uint32_t) * P);
    uint32_t* idx2 = (uint32_t*)malloc(sizeof(uint32_t) * P);
    for (int i = 0; i < P; i++) { /* coord2Morton, simple_knn.cu:55-62 */
        uint32_t m[3];
        for (int k = 0; k < 3; k++) m[k] = prep_morton(f2u_rz(((pts[3 * i + k] - mn[k]) / (mx[k] - mn[k])) * 1023.0f));
        code[i] = m[0] | (m[1] << 1) | (m[2] << 2);
        idx[i] = (uint32_t)i;
    }
    for (int shift = 0; shift < 32; shift += 8) { /* stable LSD radix sort == cub SortPairs, simple_knn.cu:211-214 */
        size_t cnt[257] = {0};
        for (int i = 0; i < P; i++) cnt[((code[i] >> shift) & 255u) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int i = 0; i < P; i++) { const size_t dst = cnt[(code[i] >> shift) & 255u]++; code2[dst] = code[i]; idx2[dst] = idx[i]; }
        uint32_t* t = code; code = code2; code2 = t;
        t = idx; idx = idx2; idx2 = t;
    }
    const int nb = (P + BOX_SIZE - 1) / BOX_SIZE;
    float* bmin = (float*)malloc(sizeof(float) * 3 * nb);
    float* bmax = (float*)malloc(sizeof(float) * 3 * nb);
    for (int b = 0; b < nb; b++) { /* boxMinMax, simple_knn.cu:79-118 */
        for (int k = 0; k < 3; k++) { bmin[3 * b + k] = FLT_MAX; bmax[3 * b + k] = -FLT_MAX; }
        for (int i = b * BOX_SIZE; i < P && i < (b + 1) * BOX_SIZE; i++)
            for (int k = 0; k < 3; k++) {
                bmin[3 * b + k] = fminf(bmin[3 * b + k], pts[3 * idx[i] + k]);
                bmax[3 * b + k] = fmaxf(bmax[3 * b + k], pts[3 * idx[i] + k]);
            }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < P; i++) { /* boxMeanDist, simple_knn.cu:148-184 */
        const float* p = pts + 3 * (size_t)idx[i];
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        const int lo = i - 3 < 0 ? 0 : i - 3, hi = i + 3 > P - 1 ? P - 1 : i + 3;
        for (int j = lo; j <= hi; j++) if (j != i) update3(p, pts + 3 * (size_t)idx[j], best);
        const float reject = best[2];
        best[0] = best[1] = best[2] = FLT_MAX;
        for (int b = 0; b < nb; b++) {
            const float d = dist_box_point(bmin + 3 * b, bmax + 3 * b, p);
            if (d > reject || d > best[2]) continue;
            for (int j = b * BOX_SIZE; j < P && j < (b + 1) * BOX_SIZE; j++) if (j != i) update3(p, pts + 3 * (size_t)idx[j], best);
        }
        out[idx[i]] = (best[0] + best[1] + best[2]) / 3.0f;
    }
    free(code); free(idx); free(code2); free(idx2); free(bmin); free(bmax);
}

/* O(P^2) definition of the same quantity, for small P */
void oracle_knn_bruteforce(int P, const float* pts, float* out) {
    for (int i = 0; i < P; i++) {
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int j = 0; j < P; j++) if (j != i) update3(pts + 3 * (size_t)i, pts + 3 * (size_t)j, best);
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}
