// Brute-force 1-nearest-neighbour over D-dimensional descriptors (sm_100a).
//
// Replaces, on the registration-inference path of the reference:
//   form 0  lib/eval.py:18-48 find_nn_gpu + lib/metrics.py:26-27 pdist('SquareL2'):
//           argmin_j sum_c (q_c - r_c)^2            (ties -> lowest j)
//   form 1  scripts/SC2_PCR/SC2_PCR.py:296-298 Matcher.match_pair:
//           argmin_j sqrt(2 - 2 * <q, r> + 1e-6)    (ties -> lowest j)
//
// One CTA owns a 128-query tile and walks a chunk of the reference set in 128-column tiles staged
// through shared memory (c-major so the 8x8 register tile is fed by conflict-free LDS.128).  The
// reference set is split across gridDim.y so that small problems still cover all 148 SMs; partial
// winners are merged with a 64-bit atomicMin on (ordered value bits << 32 | index), which realises
// "smallest value, then smallest index" exactly.  Accumulation order is fixed: one fp32 FMA per
// descriptor channel, channels ascending (oracle/matching_oracle.py knn_*_seq restates it).
#include "common.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int TQ = 128;      // queries per CTA
constexpr int TR = 128;      // reference columns per smem tile
constexpr int TC = 32;       // channels per smem tile
constexpr int NTHREADS = 256;

__device__ __forceinline__ unsigned long long pack_key(float v, int j) {
    // NaN sorts first (torch.argmin semantics), then by value; all values here are >= 0 or NaN.
    unsigned int enc = (v != v) ? 0u : (__float_as_uint(v) + 1u);
    return ((unsigned long long)enc << 32) | (unsigned int)j;
}

// Packed fp32 pairs (sm_100 fma/sub .f32x2): two independent round-to-nearest fp32 operations per instruction - the
// per-element arithmetic (and so every result bit) is that of the scalar instructions, at half the issue slots.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// distance value the reference compares, from the raw accumulator
template <int FORM>
__device__ __forceinline__ float value_of(float a) {
    if (FORM == 1) return __fsqrt_rn(__fadd_rn(__fsub_rn(2.f, __fmul_rn(2.f, a)), 1e-6f));
    return a;
}

template <int FORM>
__global__ void __launch_bounds__(NTHREADS, 2)
knn1_kernel(const float* __restrict__ Q, const float* __restrict__ R, int nq, int nr, int dim, int dimp,
            int tiles_per_split, unsigned long long* __restrict__ keys) {
    extern __shared__ float smem[];
    float* Qs = smem;                 // [dimp][TQ]
    float* Rs = smem + dimp * TQ;     // [TC][TR]
    const int b = blockIdx.z;
    Q += (size_t)b * nq * dim;
    R += (size_t)b * nr * dim;
    keys += (size_t)b * nq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * TQ;
    const bool vec_ok = (dim % 4 == 0);

    // stage the whole query tile (all channel chunks), transposed to c-major
    for (int r = warp * 32 + lane; r < TQ; r += NTHREADS) {
        const int q = q0 + r;
        for (int c = 0; c < dimp; c += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < nq) {
                if (vec_ok && c + 3 < dim) v = *reinterpret_cast<const float4*>(Q + (size_t)q * dim + c);
                else {
                    if (c + 0 < dim) v.x = Q[(size_t)q * dim + c + 0];
                    if (c + 1 < dim) v.y = Q[(size_t)q * dim + c + 1];
                    if (c + 2 < dim) v.z = Q[(size_t)q * dim + c + 2];
                    if (c + 3 < dim) v.w = Q[(size_t)q * dim + c + 3];
                }
            }
            Qs[(c + 0) * TQ + r] = v.x; Qs[(c + 1) * TQ + r] = v.y;
            Qs[(c + 2) * TQ + r] = v.z; Qs[(c + 3) * TQ + r] = v.w;
        }
    }

    float best[8];     // winner's raw accumulator (its distance value is value_of(best))
    int bestj[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = (FORM == 1) ? __int_as_float(0xff800000) : __int_as_float(0x7f800000); bestj[i] = -1; }
    int firstj = -1;

    const int tile0 = blockIdx.y * tiles_per_split;
    const int ntiles_total = (nr + TR - 1) / TR;
    const int tile1 = min(ntiles_total, tile0 + tiles_per_split);
    for (int t = tile0; t < tile1; ++t) {
        const int r0 = t * TR;
        f32x2 acc2[8][4];            // acc2[i][jp] = accumulators of columns (2 jp, 2 jp + 1)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
        for (int cc = 0; cc < dimp; cc += TC) {
            __syncthreads();
            for (int r = tid & 127; r < TR; r += 128) {
                const int j = r0 + r;
                for (int c = (tid >> 7) * 4; c < TC; c += 8) {
                    const int cg = cc + c;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < nr) {
                        if (vec_ok && cg + 3 < dim) v = *reinterpret_cast<const float4*>(R + (size_t)j * dim + cg);
                        else {
                            if (cg + 0 < dim) v.x = R[(size_t)j * dim + cg + 0];
                            if (cg + 1 < dim) v.y = R[(size_t)j * dim + cg + 1];
                            if (cg + 2 < dim) v.z = R[(size_t)j * dim + cg + 2];
                            if (cg + 3 < dim) v.w = R[(size_t)j * dim + cg + 3];
                        }
                    }
                    Rs[(c + 0) * TR + r] = v.x; Rs[(c + 1) * TR + r] = v.y;
                    Rs[(c + 2) * TR + r] = v.z; Rs[(c + 3) * TR + r] = v.w;
                }
            }
            __syncthreads();
#pragma unroll 4
            for (int c = 0; c < TC; ++c) {
                const float4 qa = *reinterpret_cast<const float4*>(Qs + (cc + c) * TQ + ty * 8);
                const float4 qb = *reinterpret_cast<const float4*>(Qs + (cc + c) * TQ + ty * 8 + 4);
                const float4 ra = *reinterpret_cast<const float4*>(Rs + c * TR + tx * 4);
                const float4 rb = *reinterpret_cast<const float4*>(Rs + c * TR + 64 + tx * 4);
                const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                const f32x2 r2[4] = {pack2(ra.x, ra.y), pack2(ra.z, ra.w), pack2(rb.x, rb.y), pack2(rb.z, rb.w)};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const f32x2 q2 = pack2(q[i], q[i]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (FORM == 0) {
                            const f32x2 d = sub2(q2, r2[j]);
                            acc2[i][j] = fma2(d, d, acc2[i][j]);
                        } else {
                            acc2[i][j] = fma2(q2, r2[j], acc2[i][j]);
                        }
                    }
                }
            }
        }
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
        // running argmin; this thread visits its columns in ascending j, so strict '<' keeps the first.
        // Fast reject on the raw accumulator: form 1's  sqrt((2 - 2 acc) + 1e-6)  is monotonically non-increasing in
        // acc (every rounding step is monotonic), so a candidate can only win when acc > the winner's acc; the exact
        // rounded value is formed and compared only then (a handful of times per row).  NaN always takes the slow path.
        bool any = false;            // does any of the 64 accumulators beat its row's current winner?  (columns past
#pragma unroll                       //  nr hold zero rows: a false alarm at worst, re-checked below)
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) any |= (FORM == 1) ? !(acc[i][j] <= best[i]) : !(acc[i][j] >= best[i]);
        if (firstj < 0) {
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                const int col = r0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < nr) firstj = col;
            }
        }
        if (any) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = r0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < nr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a = acc[i][j];
                        const bool maybe = (FORM == 1) ? !(a <= best[i]) : !(a >= best[i]);
                        if (maybe) {
                            const float v = value_of<FORM>(a), bv = value_of<FORM>(best[i]);
                            const bool take = (v < bv) || (v != v && bv == bv);
                            if (take) {
                                bestj[i] = col;
                                best[i] = (FORM == 1 && v != v) ? __int_as_float(0x7f800000) : a;    // +inf accumulator <-> NaN value
                            }
                        }
                    }
                }
            }
        }
    }
    // merge across the 16 column-threads of each query row, then across CTAs of the split
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        unsigned long long key = 0xffffffffffffffffull;
        if (bestj[i] >= 0) key = pack_key(value_of<FORM>(best[i]), bestj[i]);
        else if (firstj >= 0) key = pack_key(__int_as_float(0x7f800000), firstj);   // all +inf
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        const int q = q0 + ty * 8 + i;
        if (tx == 0 && q < nq && key != 0xffffffffffffffffull) atomicMin(keys + q, key);
    }
}

__global__ void knn1_decode_kernel(const unsigned long long* __restrict__ keys, int64_t n, int64_t* __restrict__ idx,
                                   float* __restrict__ dist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const unsigned int enc = (unsigned int)(k >> 32);
    if (idx) idx[i] = (int64_t)(unsigned int)(k & 0xffffffffu);
    if (dist) dist[i] = enc == 0u ? __int_as_float(0x7fc00000) : __uint_as_float(enc - 1u);
}

}  // namespace

extern "C" size_t eyoc_knn1_workspace_bytes(int batch, int64_t nq) {
    return eyoc_align((size_t)batch * (size_t)nq * sizeof(unsigned long long));
}

extern "C" int eyoc_knn1(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                         void* workspace, size_t workspace_bytes, int64_t* idx, float* dist, cudaStream_t stream) {
    EYOC_CHECK_ARG(q && r, "eyoc_knn1: null descriptor pointer");
    EYOC_CHECK_ARG(batch >= 1 && nq >= 0 && dim >= 1 && dim <= 256, "eyoc_knn1: bad shape batch=%d nq=%lld dim=%d", batch,
                   (long long)nq, dim);
    EYOC_CHECK_ARG(form == 0 || form == 1, "eyoc_knn1: form must be 0 (SquareL2) or 1 (sqrt(2-2ab+1e-6))");
    EYOC_CHECK_ARG(nq < (1ll << 31) && nr < (1ll << 31), "eyoc_knn1: sizes must fit int32");
    if (nq == 0) return EYOC_OK;
    if (nr <= 0) {   // torch: argmin over an empty dimension raises
        eyoc_set_error("eyoc_knn1: empty reference set (nr=%lld)", (long long)nr);
        return EYOC_ERR_DEGENERATE;
    }
    EYOC_CHECK_ARG(idx || dist, "eyoc_knn1: no output requested");
    if (workspace == nullptr || workspace_bytes < eyoc_knn1_workspace_bytes(batch, nq)) {
        eyoc_set_error("eyoc_knn1: workspace too small (%zu < %zu)", workspace_bytes, eyoc_knn1_workspace_bytes(batch, nq));
        return EYOC_ERR_WORKSPACE;
    }
    unsigned long long* keys = (unsigned long long*)workspace;
    EYOC_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)batch * nq * sizeof(unsigned long long), stream));
    const int dimp = (dim + TC - 1) / TC * TC;
    const int qtiles = (int)((nq + TQ - 1) / TQ);
    const int rtiles = (int)((nr + TR - 1) / TR);
    int nsplit = (2 * 148 + qtiles * batch - 1) / (qtiles * batch);
    nsplit = nsplit < 1 ? 1 : (nsplit > rtiles ? rtiles : nsplit);
    const int tiles_per_split = (rtiles + nsplit - 1) / nsplit;
    nsplit = (rtiles + tiles_per_split - 1) / tiles_per_split;
    const size_t smem = (size_t)(dimp * TQ + TC * TR) * sizeof(float);
    dim3 grid(qtiles, nsplit, batch);
    if (form == 0) {
        EYOC_CUDA(cudaFuncSetAttribute(knn1_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn1_kernel<0><<<grid, NTHREADS, smem, stream>>>(q, r, (int)nq, (int)nr, dim, dimp, tiles_per_split, keys);
    } else {
        EYOC_CUDA(cudaFuncSetAttribute(knn1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn1_kernel<1><<<grid, NTHREADS, smem, stream>>>(q, r, (int)nq, (int)nr, dim, dimp, tiles_per_split, keys);
    }
    EYOC_LAUNCH_CHECK();
    const int64_t n = (int64_t)batch * nq;
    knn1_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, idx, dist);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
