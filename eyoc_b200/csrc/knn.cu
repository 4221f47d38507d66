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
#include "tc_ptx.cuh"
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
            int tiles_per_split, unsigned long long* __restrict__ keys, const int* __restrict__ only_if,
            const int64_t* __restrict__ exclude) {
    extern __shared__ float smem[];
    float* Qs = smem;                 // [dimp][TQ]
    float* Rs = smem + dimp * TQ;     // [TC][TR]
    const int b = blockIdx.z;
    if (only_if != nullptr && only_if[b] == 0) return;      // fallback launch of the tensor-core path: flagged batches only
    Q += (size_t)b * nq * dim;
    R += (size_t)b * nr * dim;
    keys += (size_t)b * nq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * TQ;
    const bool vec_ok = (dim % 4 == 0);
    // second-nearest pass (eyoc_knn1_excluding): the reference column each query row must ignore
    int excl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = q0 + ty * 8 + i;
        excl[i] = (exclude != nullptr && q < nq) ? (int)exclude[(size_t)b * nq + q] : -1;
    }

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
        if (firstj < 0 && exclude == nullptr) {
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
                        const bool maybe = ((FORM == 1) ? !(a <= best[i]) : !(a >= best[i])) && col != excl[i];
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
    if (k == 0xffffffffffffffffull) {          // no admissible column (eyoc_knn1_excluding on a one-column reference set)
        if (idx) idx[i] = -1;
        if (dist) dist[i] = __int_as_float(0x7f800000);
        return;
    }
    if (idx) idx[i] = (int64_t)(unsigned int)(k & 0xffffffffu);
    if (dist) dist[i] = enc == 0u ? __int_as_float(0x7fc00000) : __uint_as_float(enc - 1u);
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core pre-filter (dim == 32): the N_q x N_r score matrix is formed on the tcgen05 tensor cores from fp16 copies of
// the descriptors - with a RIGOROUS error bound - and only the handful of columns per row that can still be the exact
// winner are re-scored with the fp32-FMA-order arithmetic of knn1_kernel above.  The result (index AND value) is the one
// knn1_kernel gives, bit for bit: the argmin over a superset of the possible winners, same accumulation order, same
// "smallest value, then lowest index" rule.
//
//   score s_j = <q, r_j>                    (form 1: the compared value sqrt(2 - 2 s + 1e-6) decreases with s)
//             = <q, r_j> - |r_j|^2 / 2      (form 0: |q - r_j|^2 = |q|^2 - 2 s_j); the norm term rides in the GEMM as
//                                            two extra K columns (q side 1, 1; r side the fp16 hi / lo parts of -|r|^2/2)
//   approximation error  |s~ - s| <= eps:   fp16 rounding of both operands 2^-9.9 |q||r| (exact products, fp32 accumulate
//                                            <= 48 terms), plus the fp32 evaluation error of the exact kernel itself;
//   tau = 2^-8 |q| max|r| + 2^-14 (|q| + max|r|)^2 + 2^-20  >=  2 (eps + delta)  with a 2x margin.
//   A column is re-scored iff s~_j >= (running row maximum of s~) - tau.  The true winner j* satisfies
//   s~_j* >= max_j s~_j - 2 eps, and so does every column whose exact value ties with it, so none of them is dropped.
// Batches with non-finite or fp16-overflowing descriptors are flagged by the preparation kernel and run through
// knn1_kernel instead (same launch sequence, no host round trip).
namespace tck {
using namespace tcp;

constexpr int TQ2 = 128;          // queries per CTA = TMEM lanes (UMMA M)
constexpr int TRT = 256;          // reference columns per accumulator tile (UMMA N)
constexpr int NS = 4;             // reference stages
constexpr int PEND = 40;          // per-row list of columns awaiting exact re-scoring (a 32-column chunk can add 32)
constexpr int ROWB = 128;         // bytes of a prepared row: 32 fp16 values | 2 fp16 norm columns | zero padding
constexpr int NEW = 8;            // epilogue warps: two per TMEM lane quadrant, each draining half of a tile's columns (the
                                  // drain - tcgen05.ld, wait, scan - is a latency chain per warp: two chains per scheduler)
constexpr int NTH = (NEW + 3) * 32;   // epilogue warps, 2 producer warps, 1 MMA warp

// fp32 rows -> prepared fp16 rows, |q| per query row, max |r| per batch, fallback flag per batch
__global__ void knn_prep_kernel(const float* __restrict__ X, long long rows_per_batch, long long rows, int is_ref, int form,
                                uint8_t* __restrict__ Xh, float* __restrict__ norm_out, unsigned int* __restrict__ rmax_bits,
                                int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int b = (int)(i / rows_per_batch);
    float x[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(X + i * 32) + c);
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
    float ss = 0.f;
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        ss = __fmaf_rn(x[c], x[c], ss);
        bad |= !(fabsf(x[c]) <= 60000.f);            // NaN, Inf or beyond the fp16 range
    }
    const float nrm = sqrtf(ss);
    bad |= !(ss <= 1.0e5f);                          // -|r|^2 / 2 must fit fp16 as well
    uint4 out[8];
    __half* h = reinterpret_cast<__half*>(out);
#pragma unroll
    for (int c = 0; c < 64; ++c) h[c] = __float2half_rn(0.f);
#pragma unroll
    for (int c = 0; c < 32; ++c) h[c] = __float2half_rn(bad ? 0.f : x[c]);
    if (form == 0 && !bad) {
        if (is_ref) {
            const float hn = -0.5f * ss;
            h[32] = __float2half_rn(hn);
            h[33] = __float2half_rn(hn - __half2float(h[32]));
        } else {
            h[32] = __float2half_rn(1.f);
            h[33] = __float2half_rn(1.f);
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(Xh + (size_t)i * ROWB)[c] = out[c];
    if (bad) atomicOr(flag + b, 1);
    else if (is_ref) {
        const unsigned int bits = __float_as_uint(nrm);          // non-negative floats order like their bit patterns
        if (bits > rmax_bits[b]) atomicMax(rmax_bits + b, bits);
    } else norm_out[i] = nrm;
    if (bad && !is_ref) norm_out[i] = 0.f;
}

template <int FORM>
__global__ void __launch_bounds__(NTH, 1)
knn_tc_kernel(const float* __restrict__ Q, const float* __restrict__ R, const uint8_t* __restrict__ Qh, const uint8_t* __restrict__ Rh,
              const float* __restrict__ qnorm, const unsigned int* __restrict__ rmax_bits, const int* __restrict__ flag, int nq,
              int nr, int tiles_per_split, unsigned long long* __restrict__ keys) {
    const int b = blockIdx.z;
    if (flag[b] != 0) return;                                    // this batch goes through knn1_kernel
    constexpr int NK = FORM == 0 ? 3 : 2;                        // K = 16 MMAs per tile: 32 channels (+ the norm columns)
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TRT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_off = ((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw);
    const uint32_t sQ = smem_u32(smem_raw) + smem_off;
    const uint32_t sR = sQ + TQ2 * ROWB;
    int* const pend = reinterpret_cast<int*>(smem_raw + smem_off + TQ2 * ROWB + NS * TRT * ROWB);
    __shared__ uint64_t bars[1 + 2 * NS + 4];
    __shared__ uint32_t tmem_base_s;
    const uint32_t q_full = smem_u32(&bars[0]), r_full = smem_u32(&bars[1]), r_empty = smem_u32(&bars[1 + NS]);
    const uint32_t acc_full = smem_u32(&bars[1 + 2 * NS]), acc_empty = smem_u32(&bars[1 + 2 * NS + 2]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * TQ2;
    const int ntiles_total = (nr + TRT - 1) / TRT;
    const int tile0 = blockIdx.y * tiles_per_split;
    const int ntl = min(ntiles_total, tile0 + tiles_per_split) - tile0;      // tiles of this CTA (>= 1 by construction)
    Q += (size_t)b * nq * 32;
    R += (size_t)b * nr * 32;
    Qh += (size_t)b * nq * ROWB;
    Rh += (size_t)b * nr * ROWB;

    if (tid == 0) {
        mbar_init(q_full, 64);
        for (int i = 0; i < NS; ++i) { mbar_init(r_full + 8 * i, 64); mbar_init(r_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, NEW * 32); }
        mbar_init_fence();
    }
    if (warp == NEW + 2) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < NEW) {
        // =========================================================== epilogue: thread = query row = TMEM lane; the two warps
        // of a lane quadrant take the lower / upper half of every tile's columns and keep independent candidates (each
        // re-scores its own exactly; the 64-bit atomicMin merges them like it merges the reference splits)
        const int quad = warp & 3, half = warp >> 2;
        constexpr int CPW = (TRT / 32) / (NEW / 4);              // 32-column chunks per warp and tile
        const int row = quad * 32 + lane;
        const int q = q0 + row;
        const bool live = q < nq;
        float qf[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) v = __ldg(reinterpret_cast<const float4*>(Q + (size_t)q * 32) + c);
            qf[4 * c] = v.x; qf[4 * c + 1] = v.y; qf[4 * c + 2] = v.z; qf[4 * c + 3] = v.w;
        }
        const float rm = __uint_as_float(rmax_bits[b]);
        const float qn = live ? qnorm[(size_t)b * nq + q] : 0.f;
        const float tau = __fmaf_rn(0.00390625f, qn * rm, __fmaf_rn(6.103515625e-5f, (qn + rm) * (qn + rm), 9.5367431640625e-7f));
        float runmax = __int_as_float(0xff800000);
        float best = 0.f;              // raw accumulator of the exact winner so far
        int bestj = -1, cnt = 0;
        int* const mypend = pend + (half * TQ2 + row) * PEND;
        // exact re-scoring of the pending columns, in the order they were found (ascending): knn1_kernel's arithmetic
        auto flush = [&]() {
            for (int p = 0; p < cnt; ++p) {
                const int col = mypend[p];
                float rr[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(R + (size_t)col * 32) + c);
                    rr[4 * c] = v.x; rr[4 * c + 1] = v.y; rr[4 * c + 2] = v.z; rr[4 * c + 3] = v.w;
                }
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (FORM == 0) {
                        const float d = __fsub_rn(qf[c], rr[c]);
                        acc = __fmaf_rn(d, d, acc);
                    } else {
                        acc = __fmaf_rn(qf[c], rr[c], acc);
                    }
                }
                bool take = bestj < 0;
                if (!take) {
                    const float v = value_of<FORM>(acc), bv = value_of<FORM>(best);
                    take = (v < bv) || (v != v && bv == bv);
                }
                if (take) { best = acc; bestj = col; }
            }
            cnt = 0;
        };
        for (int tt = 0; tt < ntl; ++tt) {
            const int ab = tt & 1;
            const int r0 = (tile0 + tt) * TRT;
            mbar_wait(acc_full + 8 * ab, (uint32_t)(tt >> 1) & 1u);
            tc_fence_after();
            const bool ragged = r0 + TRT > nr;                   // last tile: columns past nr must not raise the row maximum
#pragma unroll 1
            for (int ch = half * CPW; ch < (half + 1) * CPW; ++ch) {
                uint32_t v[32];
                const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * TRT + ch * 32);
                tmem_ld16_nowait(ta, v);
                tmem_ld16_nowait(ta + 16, v + 16);
                tmem_ld_wait();
                const int c0 = r0 + ch * 32;
                if (ragged) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i >= nr) v[i] = 0xff800000u;
                }
                float cm = __uint_as_float(v[0]);
#pragma unroll
                for (int i = 1; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(v[i]));
                runmax = fmaxf(runmax, cm);
                // form 1: a product > 1 makes sqrt(2 - 2 s + 1e-6) NaN, and torch.argmin returns the FIRST NaN column
                // whatever its s: every column whose score can exceed 1 is queued, not only those near the row maximum
                const float th = (FORM == 1 ? fminf(runmax, 1.0f) : runmax) - tau;
                if (live && cm >= th && cm > __int_as_float(0xff800000)) {      // (a chunk wholly past nr holds only -inf)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (__uint_as_float(v[i]) >= th) mypend[cnt++] = c0 + i;
                }
                if (ch == (half + 1) * CPW - 1) {                // this warp's share is drained: the MMA warp may reuse it
                    tc_fence_before();                           // once all NEW warps have arrived
                    mbar_arrive(acc_empty + 8 * ab);
                }
                if (cnt >= 8 || (ch == (half + 1) * CPW - 1 && cnt > 0)) flush();
            }
        }
        if (live && bestj >= 0) atomicMin(keys + (size_t)b * nq + q, pack_key(value_of<FORM>(best), bestj));
    } else if (warp < NEW + 2) {
        // =========================================================== producers: prepared rows -> SWIZZLE_128B operand tiles
        const int w4 = warp - NEW;
        const int c = lane & 7, rsub = lane >> 3;
        // the query tile: 128 rows, 64 per producer warp
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int p = w4 * 64 + qq * 4 + rsub;
            const int q = q0 + p;
            const uint32_t dst = sQ + (uint32_t)p * 128u + (uint32_t)((c ^ (p & 7)) << 4);
            cp_async16_or_zero(dst, Qh + (size_t)max(min(q, nq - 1), 0) * ROWB + c * 16, q < nq ? 0 : -1);
        }
        // Hand-over by the book: a thread's copies are complete (cp.async.wait_group) and made visible to the async proxy
        // the MMA reads through (fence.proxy.async) BEFORE the thread arrives on the stage's barrier.  The wait is deferred
        // by one tile so the copies of tile t + 1 are in flight while tile t lands.
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        mbar_arrive(q_full);
        for (int tt = 0; tt <= ntl; ++tt) {
            if (tt < ntl) {
                const uint32_t s = (uint32_t)(tt % NS);
                const int r0 = (tile0 + tt) * TRT;
                mbar_wait(r_empty + 8 * s, ((uint32_t)(tt / NS) & 1u) ^ 1u);
                const uint32_t base = sR + s * (uint32_t)(TRT * ROWB);
#pragma unroll 8
                for (int qq = 0; qq < 32; ++qq) {
                    const int p = w4 * 128 + qq * 4 + rsub;
                    const int j = r0 + p;
                    const uint32_t dst = base + (uint32_t)p * 128u + (uint32_t)((c ^ (p & 7)) << 4);
                    cp_async16_or_zero(dst, Rh + (size_t)min(j, nr - 1) * ROWB + c * 16, j < nr ? 0 : -1);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (tt > 0) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");      // tile tt - 1 has landed
                fence_proxy_async();
                mbar_arrive(r_full + 8 * (uint32_t)((tt - 1) % NS));
            }
        }
    } else {
        // =========================================================== MMA issuer
        if (lane == 0) {
            mbar_wait(q_full, 0);
            for (int tt = 0; tt < ntl; ++tt) {
                const uint32_t s = (uint32_t)(tt % NS), ab = (uint32_t)(tt & 1);
                mbar_wait(r_full + 8 * s, (uint32_t)(tt / NS) & 1u);
                mbar_wait(acc_empty + 8 * ab, ((uint32_t)(tt >> 1) & 1u) ^ 1u);
                fence_proxy_async();
                tc_fence_after();
                const uint32_t d = tmem_base + ab * TRT;
                const uint32_t rs = sR + s * (uint32_t)(TRT * ROWB);
#pragma unroll
                for (int j = 0; j < NK; ++j)
                    umma_f16(d, make_desc_sw128(sQ + j * 32), make_desc_sw128(rs + j * 32), IDESC, j > 0 ? 1u : 0u);
                umma_commit(r_empty + 8 * s);
                umma_commit(acc_full + 8 * ab);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == NEW + 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tck

// the fp32-FMA kernel over all batches (only_if == NULL) or over the batches whose only_if[b] != 0
static int launch_ffma(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                       unsigned long long* keys, const int* only_if, cudaStream_t stream, const int64_t* exclude = nullptr) {
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
        knn1_kernel<0><<<grid, NTHREADS, smem, stream>>>(q, r, (int)nq, (int)nr, dim, dimp, tiles_per_split, keys, only_if, exclude);
    } else {
        EYOC_CUDA(cudaFuncSetAttribute(knn1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn1_kernel<1><<<grid, NTHREADS, smem, stream>>>(q, r, (int)nq, (int)nr, dim, dimp, tiles_per_split, keys, only_if, exclude);
    }
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

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
    {
        const int rc = launch_ffma(q, r, batch, nq, nr, dim, form, keys, nullptr, stream);
        if (rc != EYOC_OK) return rc;
    }
    const int64_t n = (int64_t)batch * nq;
    knn1_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, idx, dist);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

// Nearest neighbour EXCLUDING one reference column per query (exclude [batch, nq] int64): the second pass of a K = 2
// search (pytorch3d.ops.knn_points(K=2) in lib/trainer.py:1064-1065: squared distances, ascending, ties by lowest index).
// Rows whose reference set has nothing but the excluded column get idx = -1 / dist = +inf.
extern "C" int eyoc_knn1_excluding(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                                   const int64_t* exclude, void* workspace, size_t workspace_bytes, int64_t* idx, float* dist,
                                   cudaStream_t stream) {
    EYOC_CHECK_ARG(q && r && exclude, "eyoc_knn1_excluding: null pointer");
    EYOC_CHECK_ARG(batch >= 1 && nq >= 0 && dim >= 1 && dim <= 256, "eyoc_knn1_excluding: bad shape batch=%d nq=%lld dim=%d", batch,
                   (long long)nq, dim);
    EYOC_CHECK_ARG(form == 0 || form == 1, "eyoc_knn1_excluding: form must be 0 or 1");
    EYOC_CHECK_ARG(nq < (1ll << 31) && nr < (1ll << 31) && nr >= 1, "eyoc_knn1_excluding: bad sizes");
    EYOC_CHECK_ARG(idx || dist, "eyoc_knn1_excluding: no output requested");
    if (nq == 0) return EYOC_OK;
    if (workspace == nullptr || workspace_bytes < eyoc_knn1_workspace_bytes(batch, nq)) {
        eyoc_set_error("eyoc_knn1_excluding: workspace too small (%zu < %zu)", workspace_bytes, eyoc_knn1_workspace_bytes(batch, nq));
        return EYOC_ERR_WORKSPACE;
    }
    unsigned long long* keys = (unsigned long long*)workspace;
    EYOC_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)batch * nq * sizeof(unsigned long long), stream));
    const int rc = launch_ffma(q, r, batch, nq, nr, dim, form, keys, nullptr, stream, exclude);
    if (rc != EYOC_OK) return rc;
    const int64_t n = (int64_t)batch * nq;
    knn1_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, idx, dist);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_knn1_tc_supported(int dim) { return dim == 32 ? 1 : 0; }

// keys | flag, rmax per batch | |q| per query | prepared query rows | prepared reference rows
extern "C" size_t eyoc_knn1_tc_workspace_bytes(int batch, int64_t nq, int64_t nr) {
    return eyoc_align((size_t)batch * (size_t)nq * 8) + eyoc_align((size_t)batch * 8) + eyoc_align((size_t)batch * (size_t)nq * 4) +
           eyoc_align((size_t)batch * (size_t)nq * tck::ROWB) + eyoc_align((size_t)batch * (size_t)nr * tck::ROWB);
}

extern "C" int eyoc_knn1_tc(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                            void* workspace, size_t workspace_bytes, int64_t* idx, float* dist, cudaStream_t stream) {
    EYOC_CHECK_ARG(q && r, "eyoc_knn1_tc: null descriptor pointer");
    EYOC_CHECK_ARG(dim == 32, "eyoc_knn1_tc: the tensor-core path takes 32-channel descriptors (dim=%d): use eyoc_knn1", dim);
    EYOC_CHECK_ARG(batch >= 1 && nq >= 0, "eyoc_knn1_tc: bad shape batch=%d nq=%lld", batch, (long long)nq);
    EYOC_CHECK_ARG(form == 0 || form == 1, "eyoc_knn1_tc: form must be 0 (SquareL2) or 1 (sqrt(2-2ab+1e-6))");
    EYOC_CHECK_ARG(nq < (1ll << 31) && nr < (1ll << 31), "eyoc_knn1_tc: sizes must fit int32");
    if (nq == 0) return EYOC_OK;
    if (nr <= 0) {
        eyoc_set_error("eyoc_knn1_tc: empty reference set (nr=%lld)", (long long)nr);
        return EYOC_ERR_DEGENERATE;
    }
    EYOC_CHECK_ARG(idx || dist, "eyoc_knn1_tc: no output requested");
    if (workspace == nullptr || workspace_bytes < eyoc_knn1_tc_workspace_bytes(batch, nq, nr)) {
        eyoc_set_error("eyoc_knn1_tc: workspace too small (%zu < %zu)", workspace_bytes, eyoc_knn1_tc_workspace_bytes(batch, nq, nr));
        return EYOC_ERR_WORKSPACE;
    }
    WsCarver cv(workspace, workspace_bytes);
    unsigned long long* keys = cv.take<unsigned long long>((size_t)batch * nq);
    int* flag = cv.take<int>((size_t)batch * 2);
    unsigned int* rmax = (unsigned int*)(flag + batch);
    float* qnorm = cv.take<float>((size_t)batch * nq);
    uint8_t* qh = cv.take<uint8_t>((size_t)batch * nq * tck::ROWB);
    uint8_t* rh = cv.take<uint8_t>((size_t)batch * nr * tck::ROWB);
    EYOC_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)batch * nq * sizeof(unsigned long long), stream));
    EYOC_CUDA(cudaMemsetAsync(flag, 0, (size_t)batch * 8, stream));
    const long long rows_q = (long long)batch * nq, rows_r = (long long)batch * nr;
    tck::knn_prep_kernel<<<(unsigned)((rows_q + 127) / 128), 128, 0, stream>>>(q, nq, rows_q, 0, form, qh, qnorm, rmax, flag);
    EYOC_LAUNCH_CHECK();
    tck::knn_prep_kernel<<<(unsigned)((rows_r + 127) / 128), 128, 0, stream>>>(r, nr, rows_r, 1, form, rh, nullptr, rmax, flag);
    EYOC_LAUNCH_CHECK();
    const int qtiles = (int)((nq + tck::TQ2 - 1) / tck::TQ2);
    const int rtiles = (int)((nr + tck::TRT - 1) / tck::TRT);
    int nsplit = (148 + qtiles * batch - 1) / (qtiles * batch);
    nsplit = nsplit < 1 ? 1 : (nsplit > rtiles ? rtiles : nsplit);
    const int tiles_per_split = (rtiles + nsplit - 1) / nsplit;
    nsplit = (rtiles + tiles_per_split - 1) / tiles_per_split;
    const size_t smem = (size_t)tck::TQ2 * tck::ROWB + (size_t)tck::NS * tck::TRT * tck::ROWB + (size_t)(tck::NEW / 4) * tck::TQ2 * tck::PEND * 4 + 1024;
    dim3 grid(qtiles, nsplit, batch);
    if (form == 0) {
        EYOC_CUDA(cudaFuncSetAttribute(tck::knn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tck::knn_tc_kernel<0><<<grid, tck::NTH, smem, stream>>>(q, r, qh, rh, qnorm, rmax, flag, (int)nq, (int)nr, tiles_per_split, keys);
    } else {
        EYOC_CUDA(cudaFuncSetAttribute(tck::knn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tck::knn_tc_kernel<1><<<grid, tck::NTH, smem, stream>>>(q, r, qh, rh, qnorm, rmax, flag, (int)nq, (int)nr, tiles_per_split, keys);
    }
    EYOC_LAUNCH_CHECK();
    {   // flagged batches (non-finite / out-of-range descriptors): the fp32-FMA kernel, gated per batch on the device
        const int rc = launch_ffma(q, r, batch, nq, nr, dim, form, keys, flag, stream);
        if (rc != EYOC_OK) return rc;
    }
    const int64_t n = (int64_t)batch * nq;
    knn1_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, idx, dist);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
