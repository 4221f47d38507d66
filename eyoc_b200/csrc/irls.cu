// Robust linearised 6-DoF pose (util/transform_estimation.py:89-116 est_quad_linear_robust with its helpers :56-86):
// 20 x { weighted 3n x 6 small-angle system -> normal equations -> x = inv(A^T A) A^T b -> step = Rz Ry Rx | t ->
// move the points, reweight  w = par / (|r| + par),  par halves every 5 iterations }.
// The reference runs this on CPU tensors (sgemm + LAPACK inverse, fp32).  Here one CTA does all iterations on the device:
// the 27 sums of the normal equations are accumulated in fp64 (thread -> warp shuffle -> shared memory), thread 0 solves
// the 6 x 6 system by Gauss-Jordan with partial pivoting in fp64 and composes the step, all threads move their points
// and recompute their weights in fp32 exactly as the reference formulas read.  Tolerance-level parity (the reference's
// summation order inside MKL is unknown): poses agree to ~1e-6 on the reference's own golden vector.
#include "common.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int NT = 512;

__global__ void __launch_bounds__(NT)
irls_pose_kernel(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ w_in, int n, int iterations,
                 float* __restrict__ cur, float* __restrict__ wgt, float* __restrict__ trans_out) {
    __shared__ double red[NT / 32][27];
    __shared__ float step_s[12];          // R (row-major 3x3) | t
    __shared__ double T_s[16];            // accumulated transform (row-major 4x4)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += NT) {
        cur[3 * i] = p0[3 * i]; cur[3 * i + 1] = p0[3 * i + 1]; cur[3 * i + 2] = p0[3 * i + 2];
        wgt[i] = w_in ? w_in[i] : 1.f;
    }
    if (tid < 16) T_s[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
    __syncthreads();
    float par = 1.f;
    for (int it = 0; it < iterations; ++it) {
        if (it > 0 && it % 5 == 0) par *= 0.5f;
        // ---- normal equations: upper triangle of A^T A (21) and A^T b (6)
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = 0.0;
        for (int i = tid; i < n; i += NT) {
            const double w = wgt[i];
            const double x = cur[3 * i], y = cur[3 * i + 1], z = cur[3 * i + 2];
            const double b0 = w * ((double)p1[3 * i] - x), b1 = w * ((double)p1[3 * i + 1] - y), b2 = w * ((double)p1[3 * i + 2] - z);
            // rows (scaled by w):  r0 = [0, z, -y, 1, 0, 0], r1 = [-z, 0, x, 0, 1, 0], r2 = [y, -x, 0, 0, 0, 1]
            const double r[3][6] = {{0.0, w * z, -w * y, w, 0.0, 0.0}, {-w * z, 0.0, w * x, 0.0, w, 0.0}, {w * y, -w * x, 0.0, 0.0, 0.0, w}};
            const double b[3] = {b0, b1, b2};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = a; c < 6; ++c, ++k) acc[k] += r[0][a] * r[0][c] + r[1][a] * r[1][c] + r[2][a] * r[2][c];
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] += r[0][a] * b[0] + r[1][a] * b[1] + r[2][a] * b[2];
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const double v = warp_sum_d(acc[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double M[6][7];
            double s[27];
            for (int k = 0; k < 27; ++k) {
                double v = 0.0;
                for (int w = 0; w < NT / 32; ++w) v += red[w][k];
                s[k] = v;
            }
            int k = 0;
            for (int a = 0; a < 6; ++a)
                for (int c = a; c < 6; ++c, ++k) { M[a][c] = s[k]; M[c][a] = s[k]; }
            for (int a = 0; a < 6; ++a) M[a][6] = s[21 + a];
            for (int col = 0; col < 6; ++col) {                       // Gauss-Jordan, partial pivoting
                int piv = col;
                for (int rr = col + 1; rr < 6; ++rr)
                    if (fabs(M[rr][col]) > fabs(M[piv][col])) piv = rr;
                for (int c = 0; c < 7; ++c) { const double t = M[col][c]; M[col][c] = M[piv][c]; M[piv][c] = t; }
                const double inv = 1.0 / M[col][col];                 // a singular system gives inf / nan like torch.inverse raises
                for (int c = 0; c < 7; ++c) M[col][c] *= inv;
                for (int rr = 0; rr < 6; ++rr) {
                    if (rr == col) continue;
                    const double f = M[rr][col];
                    for (int c = 0; c < 7; ++c) M[rr][c] -= f * M[col][c];
                }
            }
            const float x0 = (float)M[0][6], x1 = (float)M[1][6], x2 = (float)M[2][6];
            const float c0 = cosf(x0), s0 = sinf(x0), c1 = cosf(x1), s1 = sinf(x1), c2 = cosf(x2), s2 = sinf(x2);
            // R = Rz(x2) Ry(x1) Rx(x0)
            const float R[9] = {c2 * c1, c2 * s1 * s0 - s2 * c0, c2 * s1 * c0 + s2 * s0,
                                s2 * c1, s2 * s1 * s0 + c2 * c0, s2 * s1 * c0 - c2 * s0,
                                -s1,     c1 * s0,                c1 * c0};
            for (int q = 0; q < 9; ++q) step_s[q] = R[q];
            step_s[9] = (float)M[3][6]; step_s[10] = (float)M[4][6]; step_s[11] = (float)M[5][6];
            double Tn[16];                                            // trans = step * trans
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 4; ++c)
                    Tn[4 * a + c] = (double)R[3 * a] * T_s[c] + (double)R[3 * a + 1] * T_s[4 + c] + (double)R[3 * a + 2] * T_s[8 + c] +
                                    (c == 3 ? (double)step_s[9 + a] : 0.0);
            Tn[12] = 0.0; Tn[13] = 0.0; Tn[14] = 0.0; Tn[15] = 1.0;
            for (int q = 0; q < 16; ++q) T_s[q] = Tn[q];
        }
        __syncthreads();
        // ---- move the points, reweight
        for (int i = tid; i < n; i += NT) {
            const float x = cur[3 * i], y = cur[3 * i + 1], z = cur[3 * i + 2];
            const float nx = step_s[0] * x + step_s[1] * y + step_s[2] * z + step_s[9];
            const float ny = step_s[3] * x + step_s[4] * y + step_s[5] * z + step_s[10];
            const float nz = step_s[6] * x + step_s[7] * y + step_s[8] * z + step_s[11];
            cur[3 * i] = nx; cur[3 * i + 1] = ny; cur[3 * i + 2] = nz;
            const float dx = nx - p1[3 * i], dy = ny - p1[3 * i + 1], dz = nz - p1[3 * i + 2];
            wgt[i] = par / (sqrtf(dx * dx + dy * dy + dz * dz) + par);
        }
        __syncthreads();
    }
    if (tid < 16) trans_out[tid] = (float)T_s[tid];
}

}  // namespace

extern "C" size_t eyoc_irls_workspace_bytes(int64_t n) { return eyoc_align((size_t)n * 12) + eyoc_align((size_t)n * 4); }

extern "C" int eyoc_irls_pose(const float* pts0, const float* pts1, const float* weight, int64_t n, int iterations, float* trans_4x4,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(pts0 && pts1 && trans_4x4, "eyoc_irls_pose: null argument");
    EYOC_CHECK_ARG(n >= 1 && n < (1ll << 28) && iterations >= 0, "eyoc_irls_pose: bad n / iterations");
    if (workspace == nullptr || workspace_bytes < eyoc_irls_workspace_bytes(n)) {
        eyoc_set_error("eyoc_irls_pose: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    WsCarver c(workspace, workspace_bytes);
    float* cur = c.take<float>((size_t)n * 3);
    float* wgt = c.take<float>((size_t)n);
    irls_pose_kernel<<<1, NT, 0, stream>>>(pts0, pts1, weight, (int)n, iterations, cur, wgt, trans_4x4);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
