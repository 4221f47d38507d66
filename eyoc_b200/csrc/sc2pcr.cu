// SC2-PCR rigid-transform estimator, batched over independent pairs (sm_100a).
//
// Replaces scripts/SC2_PCR/SC2_PCR.py:307-384 (+ :33-59, :61-168, :170-196, :238-278, :409-411) and
// scripts/SC2_PCR/common.py:7-45 of the reference.  Nothing N x N is ever materialised in fp32:
//   * first-order compatibility is kept as two BIT matrices (hard: cross<d, tight: cross<d/2), 1 bit/entry;
//   * the soft SC matrix of the leading-eigenvector power iteration is nonzero exactly where `hard` is
//     set, so every mat-vec walks the set bits of a row and recomputes the fp32 entry from coordinates;
//   * the second-order measure SC2 = (tight[seeds] @ tight) * hard[seeds] (an 8000-deep 0/1 GEMM in the
//     reference) is AND+POPC over bit rows, evaluated only where hard[seed] is set, and reduced on the
//     fly to the stable top-k1 of each seed row.
// Arithmetic that decides a discrete outcome (thresholded distances) is written with explicit
// round-to-nearest intrinsics in the order torch-CPU evaluates it (tests/test_oracle_arith.py pins that
// order), so the bit matrices, seed lists and top-k index sets are bit-identical to the oracle.
// Tie rule everywhere: descending value, lowest index first (oracle SC2Config.stable_ties).
#include "common.cuh"
#include <cub/device/device_segmented_radix_sort.cuh>
#include <math.h>
#include "../../include/eyoc_b200.h"

namespace {

struct __align__(16) Pt {
    float sx, sy, sz, pad0, tx, ty, tz, pad1;
};

constexpr int MAXK = 32;          // k1, k2 <= 32
constexpr int CT = 1024;          // column tile of the N x N sweeps

// ------------------------------------------------------------------------------------------ packing
__global__ void pack_points_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int64_t total,
                                   Pt* __restrict__ P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Pt p;
    p.sx = src[i * 3 + 0]; p.sy = src[i * 3 + 1]; p.sz = src[i * 3 + 2]; p.pad0 = 0.f;
    p.tx = tgt[i * 3 + 0]; p.ty = tgt[i * 3 + 1]; p.tz = tgt[i * 3 + 2]; p.pad1 = 0.f;
    P[i] = p;
}

__device__ __forceinline__ Pt load_pt(const Pt* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    Pt r;
    r.sx = a.x; r.sy = a.y; r.sz = a.z; r.pad0 = 0.f; r.tx = b.x; r.ty = b.y; r.tz = b.z; r.pad1 = 0.f;
    return r;
}

// cross_dist = | ||s_i - s_j|| - ||t_i - t_j|| |   (SC2_PCR.py:333-335; torch.norm == sequential FMA)
__device__ __forceinline__ float cross_dist(const Pt& a, const Pt& b) {
    const float ds = dist3_fma(a.sx, a.sy, a.sz, b.sx, b.sy, b.sz);
    const float dt = dist3_fma(a.tx, a.ty, a.tz, b.tx, b.ty, b.tz);
    return fabsf(__fsub_rn(ds, dt));
}

// ------------------------------------------------------------------------------- first-order bit rows
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 32 x 32 bit-matrix transpose across the lanes of a warp: lane r holds row r (bit c = M[r][c]) and receives column r
// (bit c = M[c][r]).  Five exchange stages instead of one ballot per column.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y & ~m) >> s)) : ((x & m) | ((y & m) << s));
    }
    return x;
}

// grid (ceil(n/32), batch); lane = row, the 8 warps split the 32-bit words of each 1024-column tile.
// cross_dist(i, j) == cross_dist(j, i) bit for bit ((a-b)^2 == (b-a)^2), so only the 32x32 bit blocks on and above the
// diagonal are evaluated; the mirrored block is the bit transpose (shuffle butterfly).
// Three bit matrices: hard (cross < d), tight (cross < d/2), near (||s_i - s_j|| < R, the NMS neighbourhood of pick_seeds,
// SC2_PCR.py:50: "dist >= R" <=> "sum of squares >= s0", see sqrt_threshold).
// The two correctly rounded square roots of cross_dist are only needed near a threshold: MUFU approximations (relative
// error <= 2^-23, PTX ISA) first, and the exact evaluation for the whole warp step only when some lane's approximate value
// lies within the error margin of d or d/2.  The margin 2^-20 (ds + dt) is > 4x the worst-case difference between the
// approximate and the exactly rounded |ds - dt| ((2^-23 + 2^-24)(ds + dt) + 2^-24 |ds - dt|), so the bits are those of the
// exact evaluation (tests/test_sc2pcr_gpu.py::test_first_order_bits_bit_exact).
__global__ void __launch_bounds__(256)
first_order_bits_kernel(const Pt* __restrict__ P, int n, int W, float d_thre, float d_half, float near_s0,
                        uint32_t* __restrict__ hard, uint32_t* __restrict__ tight, uint32_t* __restrict__ near) {
    __shared__ Pt tile[CT];
    const int b = blockIdx.y;
    P += (size_t)b * n;
    hard += (size_t)b * n * W;
    tight += (size_t)b * n * W;
    near += (size_t)b * n * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int I = blockIdx.x;                        // 32-row block == word index of these rows
    const int i = I * 32 + lane;
    const bool row_ok = i < n;
    Pt me = load_pt(P + min(i, n - 1));
    for (int c0 = (I * 32 / CT) * CT; c0 < n; c0 += CT) {
        __syncthreads();
        for (int t = threadIdx.x; t < CT; t += 256) tile[t] = load_pt(P + min(c0 + t, n - 1));
        __syncthreads();
        for (int wl = warp; wl < (CT / 32); wl += 8) {
            const int J = (c0 >> 5) + wl;            // word (32-column block) index
            if (J * 32 >= n) break;
            if (J < I) continue;                      // below the diagonal: written by the mirrored block
            uint32_t hb = 0, tb = 0, nb = 0;
#pragma unroll 4
            for (int bit = 0; bit < 32; ++bit) {
                const int j = J * 32 + bit;
                const Pt q = tile[wl * 32 + bit];
                float dx = __fsub_rn(me.sx, q.sx), dy = __fsub_rn(me.sy, q.sy), dz = __fsub_rn(me.sz, q.sz);
                const float ss = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));      // dist3_fma's sum of squares
                dx = __fsub_rn(me.tx, q.tx); dy = __fsub_rn(me.ty, q.ty); dz = __fsub_rn(me.tz, q.tz);
                const float tt = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float da = sqrt_approx(ss), db = sqrt_approx(tt);
                const float ca = fabsf(da - db), mg = (da + db) * 9.5367431640625e-07f;      // 2^-20
                bool h = ca < d_thre, t = ca < d_half;
                const bool unsure = !(fabsf(ca - d_thre) > mg) || !(fabsf(ca - d_half) > mg);   // also true for NaN / Inf
                if (__any_sync(0xffffffffu, unsure)) {
                    const float c = fabsf(__fsub_rn(__fsqrt_rn(ss), __fsqrt_rn(tt)));
                    h = c < d_thre;
                    t = c < d_half;
                }
                const bool ok = row_ok && j < n;
                hb |= (uint32_t)(ok && h) << bit;
                tb |= (uint32_t)(ok && t) << bit;
                nb |= (uint32_t)(ok && !(ss >= near_s0)) << bit;
            }
            if (row_ok) {
                hard[(size_t)i * W + J] = hb;
                tight[(size_t)i * W + J] = tb;
                near[(size_t)i * W + J] = nb;
            }
            if (J != I) {                             // warp-uniform
                const uint32_t mh = transpose32(hb, lane), mt = transpose32(tb, lane), mn = transpose32(nb, lane);
                const int jrow = J * 32 + lane;       // row J*32+lane, columns I*32 .. I*32+31
                if (jrow < n) {
                    hard[(size_t)jrow * W + I] = mh;
                    tight[(size_t)jrow * W + I] = mt;
                    near[(size_t)jrow * W + I] = mn;
                }
            }
        }
    }
}

// ------------------------------------------------------------------- leading eigenvector (power iteration)
// SC2_PCR.py:179-190.  The soft SC matrix  SC_ij = clamp(1 - cross^2 / d^2, 0)  (SC2_PCR.py:341) is non-zero exactly
// where `hard` is set and does not change over the <= 20 iterations, so it is evaluated ONCE into a CSR image
// (uint16 column + fp32 value per set bit, rows in bit order) and every iteration is a sparse mat-vec over it.
// Pairs whose hard matrix is denser than the caller-sized CSR capacity keep the recompute-from-coordinates path.
struct Csr {
    uint32_t* rowptr;    // [batch, n + 1]
    uint16_t* cols;      // [batch, cap]
    float* vals;         // [batch, cap]
    int* ok;             // [batch] 1 = CSR image valid
    size_t cap;
};

// grid (ceil(n / 8), batch): warp per row -> number of set bits
__global__ void __launch_bounds__(256)
csr_count_kernel(const uint32_t* __restrict__ hard, int n, int W, uint32_t* __restrict__ rowptr) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint32_t* row = hard + ((size_t)b * n + i) * W;
    int c = 0;
    for (int w = lane; w < W; w += 32) c += __popc(row[w]);
    c = warp_sum_i(c);
    if (lane == 0) rowptr[(size_t)b * (n + 1) + i + 1] = (uint32_t)c;
}

// grid (batch): in-place inclusive scan of rowptr[1..n] (rowptr[0] = 0); ok = total fits the capacity
__global__ void __launch_bounds__(1024)
csr_scan_kernel(uint32_t* __restrict__ rowptr, int n, size_t cap, int* __restrict__ ok) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    uint32_t* r = rowptr + (size_t)blockIdx.x * (n + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_s = 0; r[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        uint32_t x = i < n ? r[i + 1] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = x + (warp ? wsum[warp - 1] : 0u) + carry;
        if (i < n) r[i + 1] = incl;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) ok[blockIdx.x] = (size_t)r[n] <= cap ? 1 : 0;
}

// grid (ceil(n / 8), batch): warp per row writes (column, SC value) for every set bit, in bit order
__global__ void __launch_bounds__(256)
csr_fill_kernel(const Pt* __restrict__ P, const uint32_t* __restrict__ hard, int n, int W, float d_sq, Csr csr) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n || !csr.ok[b]) return;
    P += (size_t)b * n;
    const uint32_t* row = hard + ((size_t)b * n + i) * W;
    uint16_t* cols = csr.cols + (size_t)b * csr.cap;
    float* vals = csr.vals + (size_t)b * csr.cap;
    uint32_t base = csr.rowptr[(size_t)b * (n + 1) + i];
    const Pt me = load_pt(P + i);
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        uint32_t m = w < W ? row[w] : 0u;
        const int c = __popc(m);
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        uint32_t pos = base + (uint32_t)(x - c);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int j = w * 32 + bit;
            const float cd = cross_dist(me, load_pt(P + j));
            cols[pos] = (uint16_t)j;
            vals[pos] = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(cd, cd), d_sq)), 0.f);
            ++pos;
        }
        base += (uint32_t)__shfl_sync(0xffffffffu, x, 31);
    }
}

// One launch per iteration; warp per row.  The last CTA of each pair normalises, applies the torch.allclose
// stopping rule and publishes the iteration count; later launches exit at once when done.
struct PowerState {
    int* done;          // [batch]
    int* iters;         // [batch]
    unsigned int* tickets;   // [batch * (num_iterations + 1)]
};

__global__ void __launch_bounds__(256)
power_step_kernel(const Pt* __restrict__ P, const uint32_t* __restrict__ hard, int n, int W, float d_sq, int t,
                  int num_iterations, float* __restrict__ vbuf, float* __restrict__ u, float* __restrict__ conf,
                  PowerState st, Csr csr) {
    const int b = blockIdx.y;
    if (st.done[b]) return;
    P += (size_t)b * n;
    hard += (size_t)b * n * W;
    const size_t batch_stride = (size_t)gridDim.y * n;
    const float* vprev = vbuf + (size_t)((t + 1) & 1) * batch_stride + (size_t)b * n;
    float* vnext = vbuf + (size_t)(t & 1) * batch_stride + (size_t)b * n;
    u += (size_t)b * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i < n) {
        float acc = 0.f;
        if (csr.ok[b]) {
            const uint16_t* cols = csr.cols + (size_t)b * csr.cap;
            const float* vals = csr.vals + (size_t)b * csr.cap;
            const uint32_t s = csr.rowptr[(size_t)b * (n + 1) + i], e = csr.rowptr[(size_t)b * (n + 1) + i + 1];
            if (t == 1) {
                for (uint32_t p = s + lane; p < e; p += 32) acc = __fmaf_rn(__ldg(vals + p), 1.0f, acc);
            } else {
                uint32_t p = s + lane;
                for (; p + 96 < e; p += 128) {               // 4 independent gathers in flight per lane
                    const float a0 = __ldg(vals + p), a1 = __ldg(vals + p + 32), a2 = __ldg(vals + p + 64), a3 = __ldg(vals + p + 96);
                    const float v0 = __ldg(vprev + __ldg(cols + p)), v1 = __ldg(vprev + __ldg(cols + p + 32));
                    const float v2 = __ldg(vprev + __ldg(cols + p + 64)), v3 = __ldg(vprev + __ldg(cols + p + 96));
                    acc = __fmaf_rn(a0, v0, acc); acc = __fmaf_rn(a1, v1, acc);
                    acc = __fmaf_rn(a2, v2, acc); acc = __fmaf_rn(a3, v3, acc);
                }
                for (; p < e; p += 32) acc = __fmaf_rn(__ldg(vals + p), __ldg(vprev + __ldg(cols + p)), acc);
            }
        } else {
            const Pt me = load_pt(P + i);
            for (int w = lane; w < W; w += 32) {
                uint32_t m = hard[(size_t)i * W + w];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int j = w * 32 + bit;
                    const float c = cross_dist(me, load_pt(P + j));
                    const float sc = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(c, c), d_sq)), 0.f);
                    const float vj = (t == 1) ? 1.0f : __ldg(vprev + j);
                    acc = __fmaf_rn(sc, vj, acc);
                }
            }
        }
    acc = warp_sum(acc);
        if (lane == 0) u[i] = acc;
    }
    // ---- last CTA of this pair: normalise + allclose
    __shared__ bool is_last;
    __shared__ double red[8];
    __shared__ int red_i[8];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(st.tickets + (size_t)b * (num_iterations + 1) + t, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double ss = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) {
        const float x = __ldcg(u + k);
        ss += (double)x * (double)x;
    }
    ss = warp_sum_d(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += red[k];
    const float denom = __fadd_rn((float)sqrt(tot), 1e-6f);
    int notclose = 0;
    for (int k = threadIdx.x; k < n; k += 256) {
        const float v = __fdiv_rn(__ldcg(u + k), denom);
        const float vp = (t == 1) ? 1.0f : vprev[k];
        const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, vp)));
        if (!(fabsf(__fsub_rn(v, vp)) <= allowed)) notclose = 1;
        vnext[k] = v;
    }
    notclose = warp_sum_i(notclose);
    if (lane == 0) red_i[warp] = notclose;
    __syncthreads();
    int nc = 0;
    for (int k = 0; k < 8; ++k) nc += red_i[k];
    if (nc == 0 || t == num_iterations) {
        for (int k = threadIdx.x; k < n; k += 256) conf[(size_t)b * n + k] = vnext[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            st.iters[b] = t;
            __threadfence();
            st.done[b] = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------- pick_seeds
// SC2_PCR.py:47-51: i survives iff for all j: score_i >= score_j or ||s_i - s_j|| >= R.  The neighbourhood test is the
// `near` bit row written by first_order_bits_kernel, so this is a sparse scan: warp per row, grid (ceil(n / 8), batch).
__global__ void __launch_bounds__(256)
nms_bits_kernel(const uint32_t* __restrict__ near, const float* __restrict__ conf, int n, int W, float* __restrict__ scores) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    conf += (size_t)b * n;
    const uint32_t* row = near + ((size_t)b * n + i) * W;
    const float ci = conf[i];
    int suppressed = 0;
    for (int w = lane; w < W; w += 32) {
        uint32_t m = __ldg(row + w);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            suppressed |= !(ci >= conf[w * 32 + bit]);
        }
    }
    suppressed = __any_sync(0xffffffffu, suppressed);
    if (lane == 0) scores[(size_t)b * n + i] = suppressed ? __fmul_rn(ci, 0.0f) : ci;
}

// The same rule on a caller-supplied dense distance matrix (drop-in Matcher.pick_seeds(dists, scores, R, max_num)):
// warp per row, coalesced row read.
__global__ void __launch_bounds__(256)
nms_dense_kernel(const float* __restrict__ dists, const float* __restrict__ conf, int n, float R, float* __restrict__ scores) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    conf += (size_t)b * n;
    const float* row = dists + ((size_t)b * n + i) * n;
    const float ci = conf[i];
    int suppressed = 0;
    for (int j = lane; j < n; j += 32) suppressed |= (!(ci >= conf[j])) && (!(__ldg(row + j) >= R));
    suppressed = __any_sync(0xffffffffu, suppressed);
    if (lane == 0) scores[(size_t)b * n + i] = suppressed ? __fmul_rn(ci, 0.0f) : ci;
}

// SC2_PCR.py:53-57 argsort(descending) -> first S: a stable segmented radix sort (descending value, ties keep the
// ascending index order they start in).  Scores are >= 0 or NaN; NaN is keyed above everything, as torch sorts it.
__global__ void seed_key_kernel(const float* __restrict__ scores, int64_t total, int n, uint32_t* __restrict__ keys,
                                int32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = scores[i];
    keys[i] = (v != v) ? 0xffffffffu : (__float_as_uint(v) & 0x7fffffffu);     // -0.0 -> 0
    idx[i] = (int32_t)(i % n);
}
__global__ void seed_take_kernel(const int32_t* __restrict__ sorted_idx, int n, int S, int batch, int32_t* __restrict__ seeds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * S) return;
    seeds[i] = sorted_idx[(size_t)(i / S) * n + (i % S)];
}

// ------------------------------------------------------------------------ per-seed consensus (cal_seed_trans)
// One CTA per seed.  SC2_PCR.py:84-134: stable top-k1 of the seed's SC2 row, local hard consensus -> top-k2,
// soft 20x20 measure, power iteration (every iterate is stored; the stopping rule is global over all seeds
// of the pair, SC2_PCR.py:186, so it is resolved by the next kernel from the per-iteration counters).
struct SeedArgs {
    const Pt* P;
    const uint32_t* hard;
    const uint32_t* tight;
    const int32_t* seeds;
    const float* sc2_dense;   // hook: caller's dense SC2 rows [batch, S, n] (drop-in Matcher.cal_seed_trans), or null
    int* status;              // bit 0: a dense SC2 value is not an integer in [0, 65535]
    int n, W, S, k1, k2, num_iterations;
    float d_thre, d_sq;
    int32_t* topk1;       // [batch, S, k1]
    int32_t* topk2;       // [batch, S, k2]
    float* local_v;       // [batch, S, num_iterations, MAXK]
    int* local_notclose;  // [batch, num_iterations + 1]
};

__global__ void __launch_bounds__(256)
seed_consensus_kernel(SeedArgs a) {
    extern __shared__ uint32_t sm[];
    const int n = a.n, W = a.W;
    uint32_t* trow = sm;                 // [W]
    uint32_t* hrow = sm + W;             // [W]
    uint32_t* nzmap = sm + 2 * W;        // [W] columns with a non-zero SC2 value
    uint32_t* keys = sm + 3 * W;         // [n]  (count << 16) | (65535 - j)
    __shared__ int ncand, nnz, nwin;
    __shared__ unsigned int hist[256], wtot[8], sel_bin, sel_rem;
    __shared__ uint32_t win[MAXK];
    __shared__ int idx1[MAXK], idx2[MAXK], fine[MAXK], lval[MAXK];
    __shared__ uint32_t lhard[MAXK];
    __shared__ float ls[MAXK][3], lt[MAXK][3];
    __shared__ float M[MAXK][MAXK + 1];

    const int b = blockIdx.y, s = blockIdx.x;
    const Pt* P = a.P + (size_t)b * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k1 = a.k1, k2 = a.k2;

    if (tid == 0) { ncand = 0; nnz = 0; nwin = 0; }
    for (int w = tid; w < W; w += 256) nzmap[w] = 0;
    __syncthreads();
    int nc;
    if (a.sc2_dense == nullptr) {
        const uint32_t* tight = a.tight + (size_t)b * n * W;
        const int seed = a.seeds[(size_t)b * a.S + s];
        for (int w = tid; w < W; w += 256) {
            trow[w] = tight[(size_t)seed * W + w];
            hrow[w] = a.hard[((size_t)b * n + seed) * W + w];
        }
        __syncthreads();
        // 1. compact the columns where hard[seed] is set
        for (int w = tid; w < W; w += 256) {
            uint32_t m = hrow[w];
            const int c = __popc(m);
            if (c) {
                int pos = atomicAdd(&ncand, c);
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    keys[pos++] = 65535u - (uint32_t)(w * 32 + bit);
                }
            }
        }
        __syncthreads();
        nc = ncand;
        // 2. SC2[seed, j] = popc(tight[seed] & tight[j]) for those columns (warp per column)
        int nz = 0;
        for (int c = warp; c < nc; c += 8) {
            const uint32_t jj = 65535u - keys[c];
            const uint32_t* row = tight + (size_t)jj * W;
            int cnt = 0;
            for (int w = lane; w < W; w += 32) cnt += __popc(trow[w] & __ldg(row + w));
            cnt = warp_sum_i(cnt);
            if (lane == 0) {
                if (cnt > 0) {
                    keys[c] = ((uint32_t)cnt << 16) | (65535u - jj);
                    atomicOr(&nzmap[jj >> 5], 1u << (jj & 31));
                    ++nz;
                } else {
                    keys[c] = 0u;
                }
            }
        }
        if (lane == 0 && nz) atomicAdd(&nnz, nz);
    } else {
        // dense rows handed in by the caller: every column is a candidate; the values are the integer counts of SC2_PCR.py:363
        const float* row = a.sc2_dense + ((size_t)b * a.S + s) * n;
        nc = n;
        int nz = 0, bad = 0;
        for (int j = tid; j < n; j += 256) {
            const float v = __ldg(row + j);
            const int cnt = (int)v;
            if (!(v >= 0.f && v <= 65535.f) || (float)cnt != v) bad = 1;
            if (cnt > 0 && !bad) {
                keys[j] = ((uint32_t)cnt << 16) | (65535u - (uint32_t)j);
                atomicOr(&nzmap[j >> 5], 1u << (j & 31));
                ++nz;
            } else {
                keys[j] = 0u;
            }
        }
        if (nz) atomicAdd(&nnz, nz);
        if (bad) atomicOr(a.status, 1);
    }
    __syncthreads();
    // 3. stable top-k1.  Keys are distinct ((count << 16) | (65535 - j)): an MSB-first radix select finds the k1-th
    //    largest key in four 8-bit passes over the candidates, the <= k1 keys at or above it are rank-sorted.
    //    Zero keys (SC2 == 0) are left to the tie rule below.
    const int want = min(nnz, k1);
    if (want > 0) {
        uint32_t prefix = 0, pmask = 0;
        unsigned int remaining = (unsigned int)want;
        if (nnz > k1) {
#pragma unroll 1
            for (int shift = 24; shift >= 0; shift -= 8) {
                hist[tid] = 0;
                __syncthreads();
                for (int c = tid; c < nc; c += 256) {
                    const uint32_t key = keys[c];
                    if (key != 0u && (key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                // bin t is the one where the count of keys in higher bins first reaches `remaining`
                const unsigned int h = hist[tid];
                unsigned int incl = h;                      // inclusive suffix sum inside the warp (towards higher bins)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int y = __shfl_down_sync(0xffffffffu, incl, o);
                    if (lane + o < 32) incl += y;
                }
                if (lane == 0) wtot[warp] = incl;
                __syncthreads();
                unsigned int above = incl - h;
                for (int w2 = warp + 1; w2 < 8; ++w2) above += wtot[w2];
                if (above < remaining && remaining <= above + h) { sel_bin = (unsigned int)tid; sel_rem = remaining - above; }
                __syncthreads();
                prefix |= sel_bin << shift;
                pmask |= 255u << shift;
                remaining = sel_rem;
            }
        } else {
            prefix = 1u;                                    // every non-zero key wins
        }
        for (int c = tid; c < nc; c += 256) {
            const uint32_t key = keys[c];
            if (key != 0u && key >= prefix) win[atomicAdd(&nwin, 1)] = key;
        }
        __syncthreads();
        if (tid < nwin) {
            const uint32_t key = win[tid];
            int rank = 0;
            for (int c2 = 0; c2 < nwin; ++c2) rank += win[c2] > key;
            idx1[rank] = 65535 - (int)(key & 0xffffu);
        }
    }
    __syncthreads();
    const int filled = want;
    if (filled < k1 && tid == 0) {
        // remaining entries of the row are exactly 0: ties resolve to the lowest indices
        int r = filled;
        for (int w = 0; w < W && r < k1; ++w) {
            uint32_t m = ~nzmap[w];
            while (m && r < k1) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int j = w * 32 + bit;
                if (j < n) idx1[r++] = j;
            }
        }
    }
    __syncthreads();
    if (tid < k1) {
        const int j = idx1[tid];
        a.topk1[((size_t)b * a.S + s) * k1 + tid] = j;
        const Pt p = load_pt(P + j);
        ls[tid][0] = p.sx; ls[tid][1] = p.sy; ls[tid][2] = p.sz;
        lt[tid][0] = p.tx; lt[tid][1] = p.ty; lt[tid][2] = p.tz;
        lhard[tid] = 0;
    }
    __syncthreads();
    // 4. local hard compatibility among the k1 (SC2_PCR.py:94-100; ((a-b)**2).sum(-1)**0.5 form)
    for (int e = tid; e < k1 * k1; e += 256) {
        const int p = e / k1, q = e % k1;
        const float ds = dist3_sum(ls[p][0], ls[p][1], ls[p][2], ls[q][0], ls[q][1], ls[q][2]);
        const float dt = dist3_sum(lt[p][0], lt[p][1], lt[p][2], lt[q][0], lt[q][1], lt[q][2]);
        if (fabsf(__fsub_rn(ds, dt)) < a.d_thre) atomicOr(&lhard[p], 1u << q);
    }
    __syncthreads();
    if (tid < k1) {   // local_SC2[q] = sum_p hard[0][p] * hard[p][q]
        int v = 0;
        const uint32_t r0 = lhard[0];
        for (int p = 0; p < k1; ++p) v += ((r0 >> p) & 1u) & ((lhard[p] >> tid) & 1u);
        lval[tid] = v;
    }
    __syncthreads();
    if (tid < k1) {   // stable descending rank -> first k2 (SC2_PCR.py:105-106)
        int rank = 0;
        const int v = lval[tid];
        for (int p = 0; p < k1; ++p) rank += (lval[p] > v) || (lval[p] == v && p < tid);
        if (rank < k2) fine[rank] = tid;
    }
    __syncthreads();
    if (tid < k2) {
        idx2[tid] = idx1[fine[tid]];
        a.topk2[((size_t)b * a.S + s) * k2 + tid] = idx2[tid];
    }
    __syncthreads();
    // 5. soft measure on the k2 (SC2_PCR.py:117-131), diagonal zeroed
    for (int e = tid; e < k2 * k2; e += 256) {
        const int p = e / k2, q = e % k2;
        const int fp = fine[p], fq = fine[q];
        const float ds = dist3_sum(ls[fp][0], ls[fp][1], ls[fp][2], ls[fq][0], ls[fq][1], ls[fq][2]);
        const float dt = dist3_sum(lt[fp][0], lt[fp][1], lt[fp][2], lt[fq][0], lt[fq][1], lt[fq][2]);
        const float c = fabsf(__fsub_rn(ds, dt));
        const float v = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(c, c), a.d_sq)), 0.f);
        M[p][q] = (p == q) ? 0.f : v;
    }
    __syncthreads();
    // 6. power iteration on the k2 x k2 matrix by one warp; every iterate is kept
    if (warp == 0) {
        float v = 1.0f;
        float* out = a.local_v + ((size_t)b * a.S + s) * a.num_iterations * MAXK;
        for (int t = 1; t <= a.num_iterations; ++t) {
            float uu = 0.f;
            for (int q = 0; q < k2; ++q) {
                const float vq = __shfl_sync(0xffffffffu, v, q);
                if (lane < k2) uu = __fmaf_rn(M[lane][q], vq, uu);
            }
            const float ss = warp_sum(lane < k2 ? __fmul_rn(uu, uu) : 0.f);
            const float vn = __fdiv_rn(uu, __fadd_rn(__fsqrt_rn(ss), 1e-6f));
            const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, v)));
            const bool close = (lane >= k2) || (fabsf(__fsub_rn(vn, v)) <= allowed);
            if (!__all_sync(0xffffffffu, close) && lane == 0)
                atomicAdd(a.local_notclose + (size_t)b * (a.num_iterations + 1) + t, 1);
            if (lane < k2) out[(size_t)(t - 1) * MAXK + lane] = vn;
            v = vn;
        }
    }
}

// ------------------------------------------------------------------------------------ 3x3 SVD / Kabsch (fp64)
// H V = U S by one-sided Jacobi; R = V diag(1,1,det(V U^T)) U^T (common.py:36-42).
__device__ void kabsch_from_moments(double sw, const double* sa, const double* sb, const double* sab, float* T16) {
    const double den = sw + 1e-6;
    double ca[3], cb[3];
    for (int r = 0; r < 3; ++r) { ca[r] = sa[r] / den; cb[r] = sb[r] / den; }
    double A[3][3], V[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            A[r][c] = sab[r * 3 + c] - ca[r] * sb[c] - sa[r] * cb[c] + sw * ca[r] * cb[c];
            V[r][c] = (r == c) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < 3; ++r) {
                    alpha += A[r][p] * A[r][p];
                    beta += A[r][q] * A[r][q];
                    gamma += A[r][p] * A[r][q];
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
                for (int r = 0; r < 3; ++r) {
                    const double ap = A[r][p], aq = A[r][q];
                    A[r][p] = c * ap - sn * aq;
                    A[r][q] = sn * ap + c * aq;
                    const double vp = V[r][p], vq = V[r][q];
                    V[r][p] = c * vp - sn * vq;
                    V[r][q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sg[3];
    for (int c = 0; c < 3; ++c) sg[c] = sqrt(A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c]);
    int ord[3] = {0, 1, 2};
    for (int x = 0; x < 2; ++x)
        for (int y = 0; y < 2 - x; ++y)
            if (sg[ord[y]] < sg[ord[y + 1]]) { const int tmp = ord[y]; ord[y] = ord[y + 1]; ord[y + 1] = tmp; }
    double U[3][3], Vs[3][3];
    const double tiny = 1e-13 * (sg[ord[0]] > 0 ? sg[ord[0]] : 1.0);
    int rank = 0;
    for (int c = 0; c < 3; ++c) {
        const int o = ord[c];
        for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][o];
        if (sg[o] > tiny && sg[o] > 0) {
            for (int r = 0; r < 3; ++r) U[r][c] = A[r][o] / sg[o];
            rank = c + 1;
        }
    }
    if (rank == 0) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) U[r][c] = (r == c) ? 1.0 : 0.0; rank = 3; }
    if (rank == 1) {   // any unit vector orthogonal to u0
        int m = 0;
        if (fabs(U[1][0]) < fabs(U[m][0])) m = 1;
        if (fabs(U[2][0]) < fabs(U[m][0])) m = 2;
        double e[3] = {0, 0, 0};
        e[m] = 1.0;
        const double d = U[m][0];
        double nn = 0;
        for (int r = 0; r < 3; ++r) { U[r][1] = e[r] - d * U[r][0]; nn += U[r][1] * U[r][1]; }
        nn = sqrt(nn);
        for (int r = 0; r < 3; ++r) U[r][1] /= nn;
        rank = 2;
    }
    if (rank == 2) {
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
    // det(V U^T) = det(V) det(U)
    auto det3 = [](double X[3][3]) {
        return X[0][0] * (X[1][1] * X[2][2] - X[1][2] * X[2][1]) - X[0][1] * (X[1][0] * X[2][2] - X[1][2] * X[2][0]) +
               X[0][2] * (X[1][0] * X[2][1] - X[1][1] * X[2][0]);
    };
    const double dd = det3(Vs) * det3(U) >= 0 ? 1.0 : -1.0;
    double R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = Vs[r][0] * U[c][0] + Vs[r][1] * U[c][1] + dd * Vs[r][2] * U[c][2];
    for (int r = 0; r < 3; ++r) {
        const double tr = cb[r] - (R[r][0] * ca[0] + R[r][1] * ca[1] + R[r][2] * ca[2]);
        T16[r * 4 + 0] = (float)R[r][0]; T16[r * 4 + 1] = (float)R[r][1]; T16[r * 4 + 2] = (float)R[r][2];
        T16[r * 4 + 3] = (float)tr;
    }
    T16[12] = 0.f; T16[13] = 0.f; T16[14] = 0.f; T16[15] = 1.f;
}

// R p + t the way SE3.transform / the einsum evaluate it: 3-term dot (sequential FMA) then + t
__device__ __forceinline__ void apply_T(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = __fadd_rn(__fmaf_rn(T[2], z, __fmaf_rn(T[1], y, __fmul_rn(T[0], x))), T[3]);
    oy = __fadd_rn(__fmaf_rn(T[6], z, __fmaf_rn(T[5], y, __fmul_rn(T[4], x))), T[7]);
    oz = __fadd_rn(__fmaf_rn(T[10], z, __fmaf_rn(T[9], y, __fmul_rn(T[8], x))), T[11]);
}

// ------------------------------------------------------------------------ per-seed hypothesis + fitness
// SC2_PCR.py:132-161: resolve the global stopping iteration, normalise weights, weighted Kabsch, count
// correspondences within inlier_threshold.
struct FitArgs {
    const Pt* P;
    const int32_t* topk2;
    const float* local_v;
    const int* local_notclose;
    int n, S, k2, num_iterations;
    float inlier_threshold;
    float* seed_weights;   // [batch, S, MAXK]
    float* seed_trans;     // [batch, S, 16]
    float* fitness;        // [batch, S]
    int* local_iters;      // [batch]
};

// One THREAD per seed: weights, fp64 moments of the k2 points, Kabsch.  (One CTA per seed left 127 threads waiting
// behind a 20-30 k-cycle serial fp64 Jacobi; 32 seeds per warp run those chains side by side.)
__global__ void __launch_bounds__(128)
seed_kabsch_kernel(FitArgs a, int batch) {
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= batch * a.S) return;
    const int b = g / a.S, s = g % a.S;
    const Pt* P = a.P + (size_t)b * a.n;
    int Tstop = a.num_iterations;
    for (int t = 1; t <= a.num_iterations; ++t)
        if (a.local_notclose[(size_t)b * (a.num_iterations + 1) + t] == 0) { Tstop = t; break; }
    if (s == 0) a.local_iters[b] = Tstop;
    const float* vv = a.local_v + (((size_t)b * a.S + s) * a.num_iterations + (Tstop - 1)) * MAXK;
    const int k2 = a.k2;
    float sum = 0.f;                                           // torch.sum over k2 values, sequential
    for (int q = 0; q < k2; ++q) {
        float w = vv[q];
        if (w < 0.f) w = 0.f;                                  // common.py:20 (threshold 0)
        sum = __fadd_rn(sum, w);
    }
    const float den = __fadd_rn(sum, 1e-6f);
    double m[16];
    for (int k = 0; k < 16; ++k) m[k] = 0.0;
    float* wout = a.seed_weights + ((size_t)b * a.S + s) * MAXK;
    for (int q = 0; q < MAXK; ++q) {
        float w = 0.f;
        if (q < k2) {
            w = vv[q];
            if (w < 0.f) w = 0.f;
            w = __fdiv_rn(w, den);
            const Pt p = load_pt(P + a.topk2[((size_t)b * a.S + s) * k2 + q]);
            const double wd = w, A3[3] = {p.sx, p.sy, p.sz}, B3[3] = {p.tx, p.ty, p.tz};
            m[0] += wd;
            for (int r = 0; r < 3; ++r) { m[1 + r] += wd * A3[r]; m[4 + r] += wd * B3[r]; }
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += wd * A3[r] * B3[c];
        }
        wout[q] = w;
    }
    float T[16];
    kabsch_from_moments(m[0], m + 1, m + 4, m + 7, T);
    float* out = a.seed_trans + ((size_t)b * a.S + s) * 16;
    for (int k = 0; k < 16; ++k) out[k] = T[k];
}

// Inlier counts of the seed hypotheses (SC2_PCR.py:150-158): 8 seeds per CTA share every point load.
constexpr int FS = 8;
__global__ void __launch_bounds__(256)
seed_fitness_kernel(FitArgs a) {
    __shared__ float T[FS][12];
    __shared__ int cnt[8][FS];
    const int b = blockIdx.y, s0 = blockIdx.x * FS;
    const Pt* P = a.P + (size_t)b * a.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < FS * 12) {
        const int s = min(s0 + tid / 12, a.S - 1);
        T[tid / 12][tid % 12] = a.seed_trans[((size_t)b * a.S + s) * 16 + tid % 12];
    }
    __syncthreads();
    int c[FS];
#pragma unroll
    for (int k = 0; k < FS; ++k) c[k] = 0;
    for (int j = tid; j < a.n; j += 256) {
        const Pt p = load_pt(P + j);
#pragma unroll
        for (int k = 0; k < FS; ++k) {
            float x, y, z;
            apply_T(T[k], p.sx, p.sy, p.sz, x, y, z);
            c[k] += dist3_fma(x, y, z, p.tx, p.ty, p.tz) < a.inlier_threshold;
        }
    }
#pragma unroll
    for (int k = 0; k < FS; ++k) {
        const int v = warp_sum_i(c[k]);
        if (lane == 0) cnt[warp][k] = v;
    }
    __syncthreads();
    if (tid < FS && s0 + tid < a.S) {
        int v = 0;
        for (int w = 0; w < 8; ++w) v += cnt[w][tid];
        a.fitness[(size_t)b * a.S + s0 + tid] = (float)v;
    }
}

// ------------------------------------------------------------------------ best seed + post_refinement + labels
struct RefineArgs {
    const Pt* P;
    const float* seed_trans;
    const float* fitness;
    const float* initial_trans;   // hook (may be null)
    int n, S, refine_iterations;
    float refine_threshold, inlier_threshold;
    float* trans;      // [batch, 16]
    float* labels;     // [batch, n] or null
    int* best_seed;    // [batch]
    int* refine_counts;   // [batch, refine_iterations + 1]: [0] = number of re-fits, then inlier counts
    float* initial_out;   // [batch, 16]
};

__global__ void __launch_bounds__(512)
refine_kernel(RefineArgs a) {
    __shared__ float T[16];
    __shared__ double red[16][17];
    __shared__ unsigned long long kred[16];
    __shared__ int stop;
    const int b = blockIdx.x;
    const Pt* P = a.P + (size_t)b * a.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.initial_trans) {
        if (tid < 16) T[tid] = a.initial_trans[(size_t)b * 16 + tid];
        if (tid == 0) a.best_seed[b] = -1;
    } else {
        // fitness.argmax (first maximum): key = (fitness bits << 32) | (~index)
        unsigned long long key = 0;
        for (int s = tid; s < a.S; s += 512) {
            const float f = a.fitness[(size_t)b * a.S + s];
            const unsigned long long k = ((unsigned long long)__float_as_uint(f) << 32) | (unsigned int)(0x7fffffff - s);
            key = k > key ? k : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane == 0) kred[warp] = key;
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < 16; ++k) key = kred[k] > key ? kred[k] : key;
            const int best = 0x7fffffff - (int)(key & 0xffffffffu);
            a.best_seed[b] = best;
            for (int k = 0; k < 16; ++k) T[k] = a.seed_trans[((size_t)b * a.S + best) * 16 + k];
        }
    }
    if (tid == 0) stop = 0;
    __syncthreads();
    if (tid < 16) a.initial_out[(size_t)b * 16 + tid] = T[tid];
    const float thr = a.refine_threshold;
    double prev = 0.0;
    int nfit = 0;
    for (int it = 0; it < a.refine_iterations; ++it) {
        double m[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) m[k] = 0.0;
        for (int j = tid; j < a.n; j += 512) {
            const Pt p = load_pt(P + j);
            float x, y, z;
            apply_T(T, p.sx, p.sy, p.sz, x, y, z);
            const float L2 = dist3_fma(x, y, z, p.tx, p.ty, p.tz);
            if (L2 < thr) {
                const float q = __fdiv_rn(L2, thr);
                const float w = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fmul_rn(q, q)));   // SC2_PCR.py:275
                const double wd = w, A3[3] = {p.sx, p.sy, p.sz}, B3[3] = {p.tx, p.ty, p.tz};
                m[0] += wd;
                for (int r = 0; r < 3; ++r) { m[1 + r] += wd * A3[r]; m[4 + r] += wd * B3[r]; }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += wd * A3[r] * B3[c];
                m[16] += 1.0;
            }
        }
#pragma unroll
        for (int k = 0; k < 17; ++k) m[k] = warp_sum_d(m[k]);
        __syncthreads();   // T fully consumed by every thread before thread 0 overwrites it
        if (lane == 0)
            for (int k = 0; k < 17; ++k) red[warp][k] = m[k];
        __syncthreads();
        if (tid == 0) {
            double tot[17];
            for (int k = 0; k < 17; ++k) {
                tot[k] = 0.0;
                for (int w = 0; w < 16; ++w) tot[k] += red[w][k];
            }
            const double cntd = tot[16];
            if (fabs(cntd - prev) < 1.0) {          // SC2_PCR.py:266
                stop = 1;
            } else {
                prev = cntd;
                a.refine_counts[(size_t)b * (a.refine_iterations + 1) + 1 + nfit] = (int)cntd;
                ++nfit;
                kabsch_from_moments(tot[0], tot + 1, tot + 4, tot + 7, T);
            }
        }
        __syncthreads();
        if (stop) break;
    }
    if (tid == 0) a.refine_counts[(size_t)b * (a.refine_iterations + 1)] = nfit;
    if (tid < 16) a.trans[(size_t)b * 16 + tid] = T[tid];
    if (a.labels) {   // SC2_PCR.py:409-411
        for (int j = tid; j < a.n; j += 512) {
            const Pt p = load_pt(P + j);
            float x, y, z;
            apply_T(T, p.sx, p.sy, p.sz, x, y, z);
            a.labels[(size_t)b * a.n + j] = dist3_sum(x, y, z, p.tx, p.ty, p.tz) < a.inlier_threshold ? 1.0f : 0.0f;
        }
    }
}

// ------------------------------------------------------------------------ standalone weighted Kabsch
__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ w, int n, float wthr,
              float* __restrict__ T) {
    __shared__ double red[4][16];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    A += (size_t)b * n * 3;
    B += (size_t)b * n * 3;
    if (w) w += (size_t)b * n;
    double m[16];
    for (int k = 0; k < 16; ++k) m[k] = 0.0;
    for (int j = tid; j < n; j += 128) {
        float wf = 1.0f;
        if (w) {
            wf = w[j];
            if (wf < wthr) { wf = 0.f; w[j] = 0.f; }      // common.py:20 mutates the caller's tensor
        }
        const double wd = wf, A3[3] = {A[j * 3], A[j * 3 + 1], A[j * 3 + 2]}, B3[3] = {B[j * 3], B[j * 3 + 1], B[j * 3 + 2]};
        m[0] += wd;
        for (int r = 0; r < 3; ++r) { m[1 + r] += wd * A3[r]; m[4 + r] += wd * B3[r]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += wd * A3[r] * B3[c];
    }
    for (int k = 0; k < 16; ++k) m[k] = warp_sum_d(m[k]);
    if (lane == 0)
        for (int k = 0; k < 16; ++k) red[warp][k] = m[k];
    __syncthreads();
    if (tid == 0) {
        double tot[16];
        for (int k = 0; k < 16; ++k) tot[k] = red[0][k] + red[1][k] + red[2][k] + red[3][k];
        float T16[16];
        kabsch_from_moments(tot[0], tot + 1, tot + 4, tot + 7, T16);
        for (int k = 0; k < 16; ++k) T[(size_t)b * 16 + k] = T16[k];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host

namespace {
__global__ void segment_offsets_kernel(int* offs, int batch, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= batch) offs[i] = i * n;
}
// smallest fp32 s with sqrtf(s) >= R (sqrtf is correctly rounded, like __fsqrt_rn): "dist >= R" == "dist^2 >= s"
float sqrt_threshold(float R) {
    if (!(R > 0.f)) return 0.f;
    float s = R * R;
    while (sqrtf(s) >= R) s = nextafterf(s, 0.f);
    while (!(sqrtf(s) >= R)) s = nextafterf(s, INFINITY);
    return s;
}
void effective_k(const eyoc_sc2_cfg* cfg, int n, int* k1, int* k2) {
    *k1 = cfg->k1;
    *k2 = cfg->k2;
    if (*k1 > n) { *k1 = 4; *k2 = 4; }     // SC2_PCR.py:76-78
}
__global__ void seed_take64_kernel(const int32_t* __restrict__ sorted_idx, int n, int S, int batch, int64_t* __restrict__ seeds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * S) return;
    seeds[i] = sorted_idx[(size_t)(i / S) * n + (i % S)];
}

// SC2_PCR.py:53-57: argsort(scores, descending) -> first S per pair (stable: ties keep ascending index order)
int rank_seeds(char* ws, const eyoc_sc2_layout& L, const float* scores, int batch, int n, int S, int32_t* seeds32, int64_t* seeds64,
               cudaStream_t stream) {
    const int64_t total = (int64_t)batch * n;
    uint32_t* skeys = (uint32_t*)(ws + L.sort_keys);
    int32_t* sidx = (int32_t*)(ws + L.sort_idx);
    int* soffs = (int*)(ws + L.sort_offsets);
    seed_key_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(scores, total, n, skeys, sidx);
    EYOC_LAUNCH_CHECK();
    segment_offsets_kernel<<<(batch + 256) / 256, 256, 0, stream>>>(soffs, batch, n);
    EYOC_LAUNCH_CHECK();
    size_t temp = L.sort_temp_bytes;
    EYOC_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(ws + L.sort_temp, temp, skeys, skeys + total, sidx, sidx + total,
                                                                 (int)total, batch, soffs, soffs + 1, 0, 32, stream));
    g_eyoc_launches += 4;
    if (seeds32) seed_take_kernel<<<(unsigned)((batch * S + 255) / 256), 256, 0, stream>>>(sidx + total, n, S, batch, seeds32);
    else seed_take64_kernel<<<(unsigned)((batch * S + 255) / 256), 256, 0, stream>>>(sidx + total, n, S, batch, seeds64);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
}  // namespace

extern "C" int eyoc_sc2pcr_layout(int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg, eyoc_sc2_layout* L) {
    EYOC_CHECK_ARG(cfg && L, "eyoc_sc2pcr_layout: null argument");
    EYOC_CHECK_ARG(batch >= 1 && n >= 1 && num_seeds >= 1, "eyoc_sc2pcr_layout: bad sizes batch=%d n=%d seeds=%d", batch, n, num_seeds);
    int k1, k2;
    effective_k(cfg, n, &k1, &k2);
    const size_t B = batch, N = n, S = num_seeds, I = cfg->num_iterations;
    const size_t W = (N + 31) / 32;
    WsCarver c(nullptr, 0);
    L->points = c.off; c.take<Pt>(B * N);
    L->hard_bits = c.off; c.take<uint32_t>(B * N * W);
    L->tight_bits = c.off; c.take<uint32_t>(B * N * W);
    L->near_bits = c.off; c.take<uint32_t>(B * N * W);
    L->vbuf = c.off; c.take<float>(2 * B * N);
    L->u = c.off; c.take<float>(B * N);
    L->confidence = c.off; c.take<float>(B * N);
    L->scores = c.off; c.take<float>(B * N);
    L->seeds = c.off; c.take<int32_t>(B * S);
    L->topk1 = c.off; c.take<int32_t>(B * S * k1);
    L->topk2 = c.off; c.take<int32_t>(B * S * k2);
    L->local_v = c.off; c.take<float>(B * S * I * MAXK);
    L->seed_weights = c.off; c.take<float>(B * S * MAXK);
    L->seed_trans = c.off; c.take<float>(B * S * 16);
    // CSR image of the soft SC matrix: capacity = 1/8 of the dense matrix (denser pairs recompute from coordinates)
    const size_t cap = N * N / 8 > 64 * N ? N * N / 8 : (N * N < 64 * N ? N * N : 64 * N);
    L->csr_capacity = cap;
    L->csr_rowptr = c.off; c.take<uint32_t>(B * (N + 1));
    L->csr_cols = c.off; c.take<uint16_t>(B * cap);
    L->csr_vals = c.off; c.take<float>(B * cap);
    {   // seed ranking: keys / indices in and out + cub temp storage
        size_t temp = 0;
        cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                           (const int32_t*)nullptr, (int32_t*)nullptr, (int)(B * N), (int)B,
                                                           (const int*)nullptr, (const int*)nullptr);
        L->sort_keys = c.off; c.take<uint32_t>(2 * B * N);
        L->sort_idx = c.off; c.take<int32_t>(2 * B * N);
        L->sort_offsets = c.off; c.take<int>(B + 1);
        L->sort_temp = c.off; c.take<char>(temp);
        L->sort_temp_bytes = temp;
    }
    // ---- small zero-initialised control block (one memset)
    L->counters = c.off; c.take<unsigned int>(B * (I + 1) + 3 * B);   // tickets | done | csr_ok
    L->global_iters = c.off; c.take<int>(B);
    L->local_notclose = c.off; c.take<int>(B * (I + 1) + B);          // notclose | local_iters
    L->best_seed = c.off; c.take<int>(B);
    L->refine_counts = c.off; c.take<int>(B * (cfg->refine_iterations + 1) + B * 16);   // counts | initial_trans (float)
    L->status = c.off; c.take<int>(4);
    L->total = c.off;
    L->words_per_row = (int)W;
    L->k1 = k1;
    L->k2 = k2;
    L->num_seeds = num_seeds;
    return EYOC_OK;
}

extern "C" size_t eyoc_sc2pcr_workspace_bytes(int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg) {
    eyoc_sc2_layout L;
    if (eyoc_sc2pcr_layout(batch, n, num_seeds, cfg, &L) != EYOC_OK) return 0;
    return L.total;
}

extern "C" int eyoc_sc2pcr(const float* src, const float* tgt, int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg,
                           const eyoc_sc2_hooks* hooks, void* workspace, size_t workspace_bytes, float* trans,
                           float* fitness, float* labels, cudaStream_t stream) {
    EYOC_CHECK_ARG(src && tgt && cfg && trans, "eyoc_sc2pcr: null argument");
    EYOC_CHECK_ARG(batch >= 1, "eyoc_sc2pcr: batch must be >= 1 (got %d)", batch);
    if (n < 4 || num_seeds < 1) {
        // the reference fails inside torch (argmax over an empty seed set / gather out of range)
        eyoc_set_error("eyoc_sc2pcr: degenerate input: %d correspondences, %d seeds", n, num_seeds);
        return EYOC_ERR_DEGENERATE;
    }
    EYOC_CHECK_ARG(n <= 65535, "eyoc_sc2pcr: n=%d exceeds 65535 (the reference truncates to max_points=8000 before this call)", n);
    EYOC_CHECK_ARG(num_seeds <= n, "eyoc_sc2pcr: num_seeds %d > n %d", num_seeds, n);
    EYOC_CHECK_ARG(cfg->num_iterations >= 1 && cfg->num_iterations <= 64, "eyoc_sc2pcr: num_iterations out of range");
    EYOC_CHECK_ARG(cfg->refine_iterations >= 0 && cfg->refine_iterations <= 64, "eyoc_sc2pcr: refine_iterations out of range");
    eyoc_sc2_layout L;
    int rc = eyoc_sc2pcr_layout(batch, n, num_seeds, cfg, &L);
    if (rc != EYOC_OK) return rc;
    EYOC_CHECK_ARG(L.k1 >= 1 && L.k1 <= MAXK && L.k2 >= 1 && L.k2 <= L.k1, "eyoc_sc2pcr: need 1 <= k2 <= k1 <= %d (got %d, %d)", MAXK, L.k1, L.k2);
    if (workspace == nullptr || workspace_bytes < L.total) {
        eyoc_set_error("eyoc_sc2pcr: workspace too small (%zu < %zu)", workspace_bytes, L.total);
        return EYOC_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    const int W = L.words_per_row, S = num_seeds, I = cfg->num_iterations;
    Pt* P = (Pt*)(ws + L.points);
    uint32_t* hard = (uint32_t*)(ws + L.hard_bits);
    uint32_t* tight = (uint32_t*)(ws + L.tight_bits);
    uint32_t* near = (uint32_t*)(ws + L.near_bits);
    int* status = (int*)(ws + L.status);
    float* vbuf = (float*)(ws + L.vbuf);
    float* u = (float*)(ws + L.u);
    float* conf = (float*)(ws + L.confidence);
    float* scores = (float*)(ws + L.scores);
    int32_t* seeds = (int32_t*)(ws + L.seeds);
    int32_t* topk1 = (int32_t*)(ws + L.topk1);
    int32_t* topk2 = (int32_t*)(ws + L.topk2);
    float* local_v = (float*)(ws + L.local_v);
    float* seed_weights = (float*)(ws + L.seed_weights);
    float* seed_trans = (float*)(ws + L.seed_trans);
    unsigned int* tickets = (unsigned int*)(ws + L.counters);
    int* done = (int*)(tickets + (size_t)batch * (I + 1));
    int* csr_ok = done + batch;
    Csr csr{(uint32_t*)(ws + L.csr_rowptr), (uint16_t*)(ws + L.csr_cols), (float*)(ws + L.csr_vals), csr_ok, L.csr_capacity};
    int* global_iters = (int*)(ws + L.global_iters);
    int* local_notclose = (int*)(ws + L.local_notclose);
    int* local_iters = local_notclose + (size_t)batch * (I + 1);
    int* best_seed = (int*)(ws + L.best_seed);
    int* refine_counts = (int*)(ws + L.refine_counts);
    float* initial_out = (float*)(refine_counts + (size_t)batch * (cfg->refine_iterations + 1));

    EYOC_CUDA(cudaMemsetAsync(ws + L.counters, 0, L.total - L.counters, stream));
    const int64_t total = (int64_t)batch * n;
    pack_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, tgt, total, P);
    EYOC_LAUNCH_CHECK();
    const bool skip_seed_stage = hooks && hooks->initial_trans;
    const float* sc2_dense = hooks ? hooks->sc2_dense : nullptr;
    EYOC_CHECK_ARG(!sc2_dense || (hooks->seeds && !hooks->initial_trans), "eyoc_sc2pcr: the sc2_dense hook needs the seeds hook");
    if (!skip_seed_stage) {
        if (!sc2_dense) {       // dense SC2 rows + seeds from the caller: the bit matrices are not needed
            first_order_bits_kernel<<<dim3((n + 31) / 32, batch), 256, 0, stream>>>(P, n, W, cfg->d_thre, cfg->d_thre_half,
                                                                                  sqrt_threshold(cfg->nms_radius), hard, tight, near);
            EYOC_LAUNCH_CHECK();
        }
        const float* conf_use = conf;
        if (hooks && hooks->confidence) {
            conf_use = hooks->confidence;
        } else if (!(hooks && hooks->seeds)) {
            PowerState st{done, global_iters, tickets};
            csr_count_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(hard, n, W, csr.rowptr);
            EYOC_LAUNCH_CHECK();
            csr_scan_kernel<<<batch, 1024, 0, stream>>>(csr.rowptr, n, csr.cap, csr.ok);
            EYOC_LAUNCH_CHECK();
            csr_fill_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(P, hard, n, W, cfg->d_thre_sq, csr);
            EYOC_LAUNCH_CHECK();
            for (int t = 1; t <= I; ++t) {
                power_step_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(P, hard, n, W, cfg->d_thre_sq, t, I, vbuf, u, conf, st, csr);
                EYOC_LAUNCH_CHECK();
            }
        }
        const int32_t* seeds_use = seeds;
        if (hooks && hooks->seeds) {
            seeds_use = hooks->seeds;
        } else {
            nms_bits_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(near, conf_use, n, W, scores);
            EYOC_LAUNCH_CHECK();
            rc = rank_seeds(ws, L, scores, batch, n, S, seeds, nullptr, stream);
            if (rc != EYOC_OK) return rc;
        }
        SeedArgs sa{P, hard, tight, seeds_use, sc2_dense, status, n, W, S, L.k1, L.k2, I, cfg->d_thre, cfg->d_thre_sq, topk1, topk2,
                    local_v, local_notclose};
        const size_t smem = (size_t)(3 * W + n) * sizeof(uint32_t);
        EYOC_CUDA(cudaFuncSetAttribute(seed_consensus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seed_consensus_kernel<<<dim3(S, batch), 256, smem, stream>>>(sa);
        EYOC_LAUNCH_CHECK();
        FitArgs fa{P, topk2, local_v, local_notclose, n, S, L.k2, I, cfg->inlier_threshold, seed_weights, seed_trans,
                   fitness ? fitness : scores /* scratch */, local_iters};
        if (!fitness) {
            EYOC_CHECK_ARG(S <= n, "unreachable");
        }
        seed_kabsch_kernel<<<(unsigned)((batch * S + 127) / 128), 128, 0, stream>>>(fa, batch);
        EYOC_LAUNCH_CHECK();
        seed_fitness_kernel<<<dim3((S + FS - 1) / FS, batch), 256, 0, stream>>>(fa);
        EYOC_LAUNCH_CHECK();
    }
    RefineArgs ra{P, seed_trans, fitness ? fitness : scores, skip_seed_stage ? hooks->initial_trans : nullptr, n, S,
                  cfg->refine_iterations, cfg->refine_threshold, cfg->inlier_threshold, trans, labels, best_seed,
                  refine_counts, initial_out};
    refine_kernel<<<batch, 512, 0, stream>>>(ra);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_kabsch_batched(const float* A, const float* B, float* w, int batch, int n, float weight_threshold,
                                   float* T, cudaStream_t stream) {
    EYOC_CHECK_ARG(A && B && T, "eyoc_kabsch_batched: null argument");
    EYOC_CHECK_ARG(batch >= 0 && n >= 0, "eyoc_kabsch_batched: bad sizes");
    if (batch == 0) return EYOC_OK;
    kabsch_kernel<<<batch, 128, 0, stream>>>(A, B, w, n, weight_threshold, T);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Stage entry points on DENSE caller tensors: the public stage methods of the reference's Matcher take (and return)
// dense [bs, n, n] / [bs, S, n] tensors (SC2_PCR.py:33-59, :170-196); the fused estimator above never forms them.
namespace {
struct PickLayout { size_t scores, keys, idx, offs, temp, temp_bytes, total; };
PickLayout pick_layout(int batch, int n) {
    PickLayout L;
    const size_t B = batch, N = n;
    WsCarver c(nullptr, 0);
    L.scores = c.off; c.take<float>(B * N);
    size_t temp = 0;
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                       (const int32_t*)nullptr, (int32_t*)nullptr, (int)(B * N), (int)B,
                                                       (const int*)nullptr, (const int*)nullptr);
    L.keys = c.off; c.take<uint32_t>(2 * B * N);
    L.idx = c.off; c.take<int32_t>(2 * B * N);
    L.offs = c.off; c.take<int>(B + 1);
    L.temp = c.off; c.take<char>(temp);
    L.temp_bytes = temp;
    L.total = c.off;
    return L;
}

// u = M v (warp per row, coalesced), grid (ceil(n / 8), batch)
__global__ void __launch_bounds__(256)
dense_matvec_kernel(const float* __restrict__ M, const float* __restrict__ vbuf, int n, int t, const int* __restrict__ done,
                    float* __restrict__ u) {
    if (*done) return;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* row = M + ((size_t)b * n + i) * n;
    const float* v = vbuf + ((size_t)((t + 1) & 1) * gridDim.y + b) * n;
    float acc = 0.f;
    if (t == 1) for (int j = lane; j < n; j += 32) acc += __ldg(row + j);                 // v0 = ones
    else for (int j = lane; j < n; j += 32) acc = __fmaf_rn(__ldg(row + j), v[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) u[(size_t)b * n + i] = acc;
}
// v = u / (||u|| + 1e-6); not-close count of torch.allclose(v, v_prev) (SC2_PCR.py:184-186), grid (batch)
__global__ void __launch_bounds__(256)
dense_norm_kernel(const float* __restrict__ u, float* __restrict__ vbuf, int n, int t, const int* __restrict__ done,
                  int* __restrict__ notclose, float* __restrict__ out) {
    if (*done) return;
    __shared__ double red[8];
    __shared__ int red_i[8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u += (size_t)b * n;
    const float* vprev = vbuf + ((size_t)((t + 1) & 1) * gridDim.x + b) * n;
    float* vnext = vbuf + ((size_t)(t & 1) * gridDim.x + b) * n;
    double ss = 0.0;
    for (int k = threadIdx.x; k < n; k += 256) ss += (double)u[k] * (double)u[k];
    ss = warp_sum_d(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += red[k];
    const float denom = __fadd_rn((float)sqrt(tot), 1e-6f);
    int nc = 0;
    for (int k = threadIdx.x; k < n; k += 256) {
        const float v = __fdiv_rn(u[k], denom);
        const float vp = (t == 1) ? 1.0f : vprev[k];
        const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, vp)));
        if (!(fabsf(__fsub_rn(v, vp)) <= allowed)) nc = 1;
        vnext[k] = v;
        out[(size_t)b * n + k] = v;
    }
    nc = warp_sum_i(nc);
    if (lane == 0) red_i[warp] = nc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        for (int k = 0; k < 8; ++k) c += red_i[k];
        if (c) atomicAdd(notclose + t, 1);
    }
}
// the stopping rule is global over the whole batch (one torch.allclose over [bs, n, 1])
__global__ void dense_stop_kernel(const int* __restrict__ notclose, int t, int* __restrict__ done, int* __restrict__ iters) {
    if (*done) return;
    *iters = t;
    if (notclose[t] == 0) *done = 1;
}
}  // namespace

extern "C" size_t eyoc_pick_seeds_workspace_bytes(int batch, int n) {
    if (batch < 1 || n < 1) return 0;
    return pick_layout(batch, n).total;
}

extern "C" int eyoc_pick_seeds_dense(const float* dists, const float* scores, int batch, int n, float R, int max_num,
                                     int64_t* seeds, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(dists && scores && seeds, "eyoc_pick_seeds_dense: null argument");
    EYOC_CHECK_ARG(batch >= 1 && n >= 1 && max_num >= 0 && max_num <= n && n <= 65535, "eyoc_pick_seeds_dense: bad sizes batch=%d n=%d max_num=%d", batch, n, max_num);
    const PickLayout P = pick_layout(batch, n);
    if (workspace == nullptr || workspace_bytes < P.total) {
        eyoc_set_error("eyoc_pick_seeds_dense: workspace too small (%zu < %zu)", workspace_bytes, P.total);
        return EYOC_ERR_WORKSPACE;
    }
    if (max_num == 0) return EYOC_OK;
    char* ws = (char*)workspace;
    float* sc = (float*)(ws + P.scores);
    nms_dense_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(dists, scores, n, R, sc);
    EYOC_LAUNCH_CHECK();
    eyoc_sc2_layout L;
    memset(&L, 0, sizeof(L));
    L.sort_keys = P.keys; L.sort_idx = P.idx; L.sort_offsets = P.offs; L.sort_temp = P.temp; L.sort_temp_bytes = P.temp_bytes;
    return rank_seeds(ws, L, sc, batch, n, max_num, nullptr, seeds, stream);
}

extern "C" size_t eyoc_power_iteration_workspace_bytes(int batch, int n, int num_iterations) {
    if (batch < 1 || n < 1 || num_iterations < 1) return 0;
    return eyoc_align((size_t)3 * batch * n * sizeof(float)) + eyoc_align((size_t)(num_iterations + 4) * sizeof(int));
}

extern "C" int eyoc_power_iteration_dense(const float* M, int batch, int n, int num_iterations, float* v_out, int* iters_out,
                                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(M && v_out, "eyoc_power_iteration_dense: null argument");
    EYOC_CHECK_ARG(batch >= 1 && n >= 1 && num_iterations >= 1 && num_iterations <= 1024, "eyoc_power_iteration_dense: bad sizes");
    const size_t need = eyoc_power_iteration_workspace_bytes(batch, n, num_iterations);
    if (workspace == nullptr || workspace_bytes < need) {
        eyoc_set_error("eyoc_power_iteration_dense: workspace too small (%zu < %zu)", workspace_bytes, need);
        return EYOC_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    float* vbuf = (float*)ws;                                     // [2, batch, n]
    float* u = vbuf + (size_t)2 * batch * n;
    int* ctl = (int*)(ws + eyoc_align((size_t)3 * batch * n * sizeof(float)));     // done | iters | pad | notclose[1..I]
    EYOC_CUDA(cudaMemsetAsync(ctl, 0, (size_t)(num_iterations + 4) * sizeof(int), stream));
    int* done = ctl, *iters = ctl + 1, *notclose = ctl + 2;
    for (int t = 1; t <= num_iterations; ++t) {
        dense_matvec_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(M, vbuf, n, t, done, u);
        EYOC_LAUNCH_CHECK();
        dense_norm_kernel<<<batch, 256, 0, stream>>>(u, vbuf, n, t, done, notclose, v_out);
        EYOC_LAUNCH_CHECK();
        dense_stop_kernel<<<1, 1, 0, stream>>>(notclose, t, done, iters);
        EYOC_LAUNCH_CHECK();
    }
    if (iters_out) EYOC_CUDA(cudaMemcpyAsync(iters_out, iters, sizeof(int), cudaMemcpyDeviceToDevice, stream));
    return EYOC_OK;
}
