/* CPU oracle helper (TEST INFRASTRUCTURE, see oracle/__init__.py): brute-force K nearest neighbours in the accumulation
 * order the CUDA kernels and pytorch3d's knn.cu define - d = a_c - b_c rounded to fp32, acc = fmaf(d, d, acc), channels
 * ascending - the K smallest in ascending order, ties to the lower index.  Restates pytorch3d.ops.knn_points as
 * lib/trainer.py:1064-1065,1182 of the reference uses it (pytorch3d/csrc/knn/knn.cu: `dist += diff * diff` under nvcc's
 * default FMA contraction; MinK keeps the first of equal distances).  Built by oracle/build_c.py with
 * gcc -O2 -ffp-contract=off -fopenmp; the numpy restatement in oracle/labeler_oracle.py is the same arithmetic and is
 * what this is checked against (tests/test_oracle_labeler.py). */
#include <math.h>
#include <stdint.h>

void knn_sq_seq(const float* q, int64_t nq, const float* r, int64_t nr, int dim, int K, int64_t* idx, float* dist) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        float best[8];
        int64_t bi[8];
        for (int k = 0; k < K; ++k) { best[k] = INFINITY; bi[k] = -1; }
        const float* a = q + i * dim;
        for (int64_t j = 0; j < nr; ++j) {
            const float* b = r + j * dim;
            float acc = 0.f;
            for (int c = 0; c < dim; ++c) {
                const float d = a[c] - b[c];
                acc = fmaf(d, d, acc);
            }
            /* insert keeping (value, index) ascending; strict '<' so that an equal later column stays behind */
            int pos = K;
            while (pos > 0 && (acc < best[pos - 1] || bi[pos - 1] < 0)) --pos;
            if (pos < K) {
                for (int k = K - 1; k > pos; --k) { best[k] = best[k - 1]; bi[k] = bi[k - 1]; }
                best[pos] = acc;
                bi[pos] = j;
            }
        }
        for (int k = 0; k < K; ++k) { idx[i * K + k] = bi[k]; dist[i * K + k] = best[k]; }
    }
}
