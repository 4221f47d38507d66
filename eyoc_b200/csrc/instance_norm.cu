// Sparse-tensor instance normalisation (sm_100a): ME.MinkowskiInstanceNorm as the reference's BasicBlockIN uses it
// (model/common.py:7-8, model/residual_block.py:60-61, the ResUNetIN2* variants of model/resunet.py:229-251).
//
// Per cloud (batch index) and channel:  y = (x - mean) / sqrt(var + eps) * weight + bias  with the biased variance of the
// cloud's rows (MinkowskiEngine's MinkowskiInstanceNormFunction: global average pooling of x, then of (x - mean)^2; restated,
// ME's source is not available here - oracle/resunet_oracle.py::_in carries the same statement).  Three passes over the rows,
// all bandwidth-bound: sum -> centred sum of squares (two-pass variance, fp64 accumulators) -> apply, with the residual add and
// the ReLU of the block (model/residual_block.py:47-51) fused into the apply pass.  Rows may be fp32 or split-half (xh_format.cuh).
//
// A CTA reduces its 256 rows in shared memory first (rows of a chunk belong to one or two clouds in the collated order), so a
// (cloud, channel) accumulator sees one fp64 atomic per CTA.  The order of those atomics is not fixed: the fp64 sums can differ
// in their last bit from run to run, the fp32 results they round to almost never do.
#include "common.cuh"
#include "xh_format.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int IN_ROWS = 256;       // rows per CTA of the statistics passes
constexpr int IN_SLOTS = 4;        // consecutive cloud ids a CTA reduces in shared memory (others go straight to global)

__device__ __forceinline__ void in_load8(const uint8_t* xh, const float* xf, long long row, int c, int col, float* y) {
    if (xh) {
        xh_load8(xh + (size_t)row * c * 4, col, y);
    } else {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(xf + row * c + col));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(xf + row * c + col + 4));
        y[0] = a0.x; y[1] = a0.y; y[2] = a0.z; y[3] = a0.w; y[4] = a1.x; y[5] = a1.y; y[6] = a1.z; y[7] = a1.w;
    }
}

// PASS 1: acc[b][ch] += x, cnt[b] += 1.   PASS 2: acc[b][ch] += (x - mean[b][ch])^2 with mean = sum / cnt in fp32 (as applied).
template <int PASS>
__global__ void __launch_bounds__(256)
in_stats_kernel(const uint8_t* __restrict__ xh, const float* __restrict__ xf, const int* __restrict__ coords, long long n, int c,
                int num_clouds, double* __restrict__ acc, int* __restrict__ cnt, const float* __restrict__ mean) {
    __shared__ double s_acc[IN_SLOTS][256];
    __shared__ int s_cnt[IN_SLOTS];
    const int G = c >> 3, g = threadIdx.x % G, rl = threadIdx.x / G, RP = 256 / G;
    const long long row0 = (long long)blockIdx.x * IN_ROWS;
    const int b_first = coords[4 * row0];
    for (int i = threadIdx.x; i < IN_SLOTS * 256; i += 256) (&s_acc[0][0])[i] = 0.0;
    if (threadIdx.x < IN_SLOTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    double a[8];
    float m[8];
    int cur = -1, rows = 0;
    auto flush = [&]() {
        if (cur < 0) return;
        const int slot = cur - b_first;
        if (slot >= 0 && slot < IN_SLOTS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[slot][g * 8 + j], a[j]);
            if (PASS == 1 && g == 0) atomicAdd(&s_cnt[slot], rows);
        } else if (cur < num_clouds) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(acc + (size_t)cur * c + g * 8 + j, a[j]);
            if (PASS == 1 && g == 0) atomicAdd(cnt + cur, rows);
        }
    };
    for (int r = rl; r < IN_ROWS; r += RP) {
        const long long row = row0 + r;
        if (row >= n) break;
        const int b = coords[4 * row];
        if (b != cur) {
            flush();
            cur = b;
            rows = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = 0.0;
            if (PASS == 2 && b >= 0 && b < num_clouds) {
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = mean[(size_t)b * c + g * 8 + j];
            }
        }
        float y[8];
        in_load8(xh, xf, row, c, g * 8, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (PASS == 1) {
                a[j] += (double)y[j];
            } else {
                const float d = __fsub_rn(y[j], m[j]);
                a[j] += (double)d * (double)d;
            }
        }
        ++rows;
    }
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < IN_SLOTS * c; i += 256) {
        const int slot = i / c, ch = i % c, b = b_first + slot;
        const double v = s_acc[slot][ch];
        if (b < num_clouds && v != 0.0) atomicAdd(acc + (size_t)b * c + ch, v);
    }
    if (PASS == 1 && threadIdx.x < IN_SLOTS) {
        const int b = b_first + threadIdx.x;
        if (b < num_clouds && s_cnt[threadIdx.x]) atomicAdd(cnt + b, s_cnt[threadIdx.x]);
    }
}

// which = 0: mean = sum / cnt (fp32).  which = 1: inv_std = 1 / sqrt(var + eps) with var = centred sum of squares / cnt.
__global__ void in_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ cnt, int num_clouds, int c, float eps,
                                   int which, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_clouds * c) return;
    const int k = cnt[i / c];
    if (k == 0) { out[i] = 0.f; return; }
    const float v = (float)(acc[i] / (double)k);
    out[i] = which == 0 ? v : __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(v, eps)));
}

// one thread per 8 channels: y = ((x - mean) * inv_std) * weight + bias (+ residual) (ReLU)
__global__ void __launch_bounds__(256)
in_apply_kernel(const uint8_t* __restrict__ xh, const float* __restrict__ xf, const int* __restrict__ coords, long long n8, int c,
                int num_clouds, const float* __restrict__ mean, const float* __restrict__ inv_std, const float* __restrict__ weight,
                const float* __restrict__ bias, const uint8_t* __restrict__ res_h, const float* __restrict__ res_f, int relu,
                uint8_t* __restrict__ out_h, float* __restrict__ out_f, int* __restrict__ range_status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const int G = c >> 3;
    const long long row = i / G;
    const int col = (int)(i - row * G) * 8;
    const int b = coords[4 * row];
    float y[8], r[8];
    in_load8(xh, xf, row, c, col, y);
    if (res_h || res_f) in_load8(res_h, res_f, row, c, col, r);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float mu = (b >= 0 && b < num_clouds) ? mean[(size_t)b * c + col + j] : 0.f;
        const float is = (b >= 0 && b < num_clouds) ? inv_std[(size_t)b * c + col + j] : 0.f;
        float v = __fmul_rn(__fsub_rn(y[j], mu), is);
        v = __fadd_rn(__fmul_rn(v, weight ? __ldg(weight + col + j) : 1.f), bias ? __ldg(bias + col + j) : 0.f);
        if (res_h || res_f) v = __fadd_rn(v, r[j]);
        if (relu) v = fmaxf(v, 0.f);
        bad |= !(fabsf(v) < 65504.f);
        y[j] = v;
    }
    if (out_h) {
        if (bad && range_status) atomicOr(range_status, 1);
        xh_store8(out_h + (size_t)row * c * 4, col, y);
    } else {
        *reinterpret_cast<float4*>(out_f + row * c + col) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(out_f + row * c + col + 4) = make_float4(y[4], y[5], y[6], y[7]);
    }
}

}  // namespace

extern "C" size_t eyoc_instance_norm_workspace_bytes(int num_clouds, int c) {
    const size_t bc = (size_t)(num_clouds > 0 ? num_clouds : 1) * (size_t)(c > 0 ? c : 1);
    return eyoc_align(bc * 8) * 2 + eyoc_align(bc * 4) * 2 + eyoc_align((size_t)(num_clouds > 0 ? num_clouds : 1) * 4) + 256;
}

extern "C" int eyoc_instance_norm(const void* x, int x_packed, const int32_t* coords, int64_t n, int c, int num_clouds,
                                  const float* weight, const float* bias, float eps, const void* residual, int residual_packed,
                                  int relu, void* out, int out_packed, int32_t* range_status, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(x && coords && out, "eyoc_instance_norm: null argument");
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31) && num_clouds >= 1, "eyoc_instance_norm: bad sizes");
    EYOC_CHECK_ARG(c >= 32 && c <= 256 && c % 32 == 0 && 256 % (c / 8) == 0, "eyoc_instance_norm: c must be 32, 64, 128 or 256 (got %d)", c);
    if (n == 0) return EYOC_OK;
    if (workspace == nullptr || workspace_bytes < eyoc_instance_norm_workspace_bytes(num_clouds, c)) {
        eyoc_set_error("eyoc_instance_norm: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    WsCarver w(workspace, workspace_bytes);
    const size_t bc = (size_t)num_clouds * c;
    double* sum = w.take<double>(bc);
    double* sq = w.take<double>(bc);
    float* mean = w.take<float>(bc);
    float* inv = w.take<float>(bc);
    int* cnt = w.take<int>(num_clouds);
    EYOC_CUDA(cudaMemsetAsync(sum, 0, (size_t)((char*)(cnt + num_clouds) - (char*)sum), stream));
    const uint8_t* xh = x_packed ? (const uint8_t*)x : nullptr;
    const float* xf = x_packed ? nullptr : (const float*)x;
    const unsigned gs = (unsigned)((n + IN_ROWS - 1) / IN_ROWS), gf = (unsigned)((bc + 255) / 256);
    in_stats_kernel<1><<<gs, 256, 0, stream>>>(xh, xf, coords, n, c, num_clouds, sum, cnt, nullptr);
    EYOC_LAUNCH_CHECK();
    in_finalize_kernel<<<gf, 256, 0, stream>>>(sum, cnt, num_clouds, c, eps, 0, mean);
    EYOC_LAUNCH_CHECK();
    in_stats_kernel<2><<<gs, 256, 0, stream>>>(xh, xf, coords, n, c, num_clouds, sq, cnt, mean);
    EYOC_LAUNCH_CHECK();
    in_finalize_kernel<<<gf, 256, 0, stream>>>(sq, cnt, num_clouds, c, eps, 1, inv);
    EYOC_LAUNCH_CHECK();
    const long long n8 = (long long)n * (c / 8);
    in_apply_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(
        xh, xf, coords, n8, c, num_clouds, mean, inv, weight, bias, (residual && residual_packed) ? (const uint8_t*)residual : nullptr,
        (residual && !residual_packed) ? (const float*)residual : nullptr, relu, out_packed ? (uint8_t*)out : nullptr,
        out_packed ? nullptr : (float*)out, range_status);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
