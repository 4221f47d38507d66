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

// Packed fp32 pairs (sm_100 add / sub / mul / fma .f32x2): two independent round-to-nearest fp32 operations per instruction,
// every result bit that of the scalar instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// First-order compatibility as three bit matrices (row stride W words, W a multiple of 4):
//   hard (cross < d), tight (cross < d/2), near (||s_i - s_j|| < R: the NMS neighbourhood of pick_seeds, SC2_PCR.py:50,
//   "dist >= R" <=> "sum of squares >= s0", see sqrt_threshold).
// cross_dist(i, j) == cross_dist(j, i) bit for bit ((a-b)^2 == (b-a)^2), so only macro tiles on and above the diagonal are
// evaluated and the mirrored words are the bit transposes (shuffle butterfly).
// Work decomposition: grid (ceil(n / 512), ceil(n / 128), batch), 4 warps; warp w owns the 128 x 128 macro tile
// (A = blockIdx.y, B = 4 blockIdx.x + w), i.e. 4 x 4 blocks of 32 rows (lane = row) x 32 columns (one word).  Every store
// is then a 16-byte piece of a row (4 consecutive words) - the one-word-per-row stores of a block-at-a-time scheme cost
// a 32-byte sector for 4 useful bytes and made the LSU, not the ALU, the limiter.
// Arithmetic: sums of squares on packed f32x2 (two columns per instruction).  The two correctly rounded square roots of
// cross_dist are only needed near a threshold: MUFU approximations (relative error <= 2^-23, PTX ISA) first, and the exact
// evaluation for the whole warp step only when some lane's approximate value lies within the error margin of d or d/2.
// The margin 2^-20 (ds + dt) is > 3x the worst-case difference between the approximate and the exactly rounded |ds - dt|
// ((2^-23 + 2^-24)(ds + dt) + 2^-24 |ds - dt|), so the bits are those of the exact evaluation
// (tests/test_sc2pcr_gpu.py::test_first_order_bits_bit_exact, ..._near_threshold_stress).
constexpr int FO_COLS = 512;

// exact classification of one entry (kept out of line: it runs for a handful of warp steps per million, and inlining it
// into every unrolled copy of the inner loop made the kernel instruction-fetch bound).  Bit 31 of the three returned words
// = hard / tight / near, the form the funnel-shift inserts of the inner loop take.
__device__ __noinline__ uint3 first_order_exact(float ss, float tt, float d_thre, float d_half, float near_s0) {
    const float c = fabsf(__fsub_rn(__fsqrt_rn(ss), __fsqrt_rn(tt)));
    return make_uint3((c < d_thre) ? 0x80000000u : 0u, (c < d_half) ? 0x80000000u : 0u, !(ss >= near_s0) ? 0x80000000u : 0u);
}

__global__ void __launch_bounds__(128)
first_order_bits_kernel(const Pt* __restrict__ P, int n, int W, float d_thre, float d_half, float near_s0,
                        uint32_t* __restrict__ hard, uint32_t* __restrict__ tight, uint32_t* __restrict__ near) {
    __shared__ __align__(16) float cs[6][FO_COLS];          // column points, structure of arrays: sx sy sz tx ty tz
    __shared__ uint32_t blk[4][3][16][32];                  // per warp: the 16 block results (hard | tight | near) x lane
    const int A = blockIdx.y;
    if ((int)blockIdx.x * 4 + 3 < A) return;              // the whole CTA lies below the diagonal
    const int b = blockIdx.z;
    P += (size_t)b * n;
    hard += (size_t)b * n * W;
    tight += (size_t)b * n * W;
    near += (size_t)b * n * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c_base = blockIdx.x * FO_COLS;
    for (int t = threadIdx.x; t < FO_COLS; t += 128) {
        const Pt q = load_pt(P + min(c_base + t, n - 1));
        cs[0][t] = q.sx; cs[1][t] = q.sy; cs[2][t] = q.sz; cs[3][t] = q.tx; cs[4][t] = q.ty; cs[5][t] = q.tz;
    }
    __syncthreads();
    const int B = blockIdx.x * 4 + warp;
    if (B < A || B * 128 >= n) return;
    const float* col = &cs[0][warp * 128];
    uint32_t (*res)[16][32] = blk[warp];
#pragma unroll 1
    for (int ri = 0; ri < 4; ++ri) {
        const int i = A * 128 + ri * 32 + lane;
        const bool row_ok = i < n;
        const Pt me = load_pt(P + min(i, n - 1));
        const f32x2 mx = pack2(me.sx, me.sx), my = pack2(me.sy, me.sy), mz = pack2(me.sz, me.sz);
        const f32x2 ux = pack2(me.tx, me.tx), uy = pack2(me.ty, me.ty), uz = pack2(me.tz, me.tz);
#pragma unroll 1
        for (int cj = (A == B ? ri : 0); cj < 4; ++cj) {   // diagonal macro tile: blocks below its diagonal are mirror images
            uint32_t hb = 0, tb = 0, nb = 0;
            const int j0 = B * 128 + cj * 32;
            // Columns are visited from 31 down to 0 and every result enters its word at bit 0 with ONE funnel shift
            // (word = word << 1 | sign bit): for finite values `x < thr` is the sign of x - thr, which the margin test
            // needs anyway.  Non-finite values make `unsure` true and go through the exact routine.
#pragma unroll 4
            for (int bp = 15; bp >= 0; --bp) {
                const int c = cj * 32 + 2 * bp;
                f32x2 dx = sub2(mx, *reinterpret_cast<const f32x2*>(col + c));
                f32x2 dy = sub2(my, *reinterpret_cast<const f32x2*>(col + FO_COLS + c));
                f32x2 dz = sub2(mz, *reinterpret_cast<const f32x2*>(col + 2 * FO_COLS + c));
                const f32x2 ss2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));                 // dist3_fma's sum of squares
                dx = sub2(ux, *reinterpret_cast<const f32x2*>(col + 3 * FO_COLS + c));
                dy = sub2(uy, *reinterpret_cast<const f32x2*>(col + 4 * FO_COLS + c));
                dz = sub2(uz, *reinterpret_cast<const f32x2*>(col + 5 * FO_COLS + c));
                const f32x2 tt2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                float ss[2], tt[2];
                unpack2(ss2, ss[0], ss[1]);
                unpack2(tt2, tt[0], tt[1]);
                uint32_t wh[2], wt[2], wn[2];                  // bit 31 = the result
                bool unsure = false;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float da = sqrt_approx(ss[e]), db = sqrt_approx(tt[e]);
                    const float ca = fabsf(da - db), mg = (da + db) * 9.5367431640625e-07f;      // 2^-20
                    const float r1 = ca - d_thre, r2 = ca - d_half, r3 = ss[e] - near_s0;
                    wh[e] = __float_as_uint(r1);
                    wt[e] = __float_as_uint(r2);
                    wn[e] = __float_as_uint(r3);
                    unsure |= !(fabsf(r1) > mg) || !(fabsf(r2) > mg) || !(fabsf(r3) <= 3.0e38f);   // last: NaN / Inf sums
                }
                if (__any_sync(0xffffffffu, unsure)) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint3 x = first_order_exact(ss[e], tt[e], d_thre, d_half, near_s0);
                        wh[e] = x.x; wt[e] = x.y; wn[e] = x.z;
                    }
                }
#pragma unroll
                for (int e = 1; e >= 0; --e) {
                    hb = __funnelshift_l(wh[e], hb, 1);
                    tb = __funnelshift_l(wt[e], tb, 1);
                    nb = __funnelshift_l(wn[e], nb, 1);
                }
            }
            {   // rows / columns beyond n: cleared per word, not per entry
                const int valid = n - j0;                        // columns of this word that exist
                const uint32_t cm = !row_ok || valid <= 0 ? 0u : (valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u));
                hb &= cm; tb &= cm; nb &= cm;
            }
            res[0][ri * 4 + cj][lane] = hb;
            res[1][ri * 4 + cj][lane] = tb;
            res[2][ri * 4 + cj][lane] = nb;
        }
    }
    __syncwarp();
    if (A == B) {
#pragma unroll 1
        for (int m = 0; m < 3; ++m)
#pragma unroll 1
            for (int ri = 1; ri < 4; ++ri)
                for (int cj = 0; cj < ri; ++cj) res[m][ri * 4 + cj][lane] = transpose32(res[m][cj * 4 + ri][lane], lane);
        __syncwarp();
    }
    uint32_t* const mats[3] = {hard, tight, near};
    // rows of macro row A, words 4 B .. 4 B + 3
#pragma unroll 1
    for (int ri = 0; ri < 4; ++ri) {
        const int i = A * 128 + ri * 32 + lane;
        if (i < n) {
            const size_t o = (size_t)i * W + 4 * B;
#pragma unroll
            for (int m = 0; m < 3; ++m)
                *reinterpret_cast<uint4*>(mats[m] + o) =
                    make_uint4(res[m][ri * 4][lane], res[m][ri * 4 + 1][lane], res[m][ri * 4 + 2][lane], res[m][ri * 4 + 3][lane]);
        }
    }
    if (A == B) return;
    // mirrored: rows of macro row B, words 4 A .. 4 A + 3
#pragma unroll 1
    for (int cj = 0; cj < 4; ++cj) {
        const int jrow = B * 128 + cj * 32 + lane;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const uint4 v = make_uint4(transpose32(res[m][cj][lane], lane), transpose32(res[m][4 + cj][lane], lane),
                                       transpose32(res[m][8 + cj][lane], lane), transpose32(res[m][12 + cj][lane], lane));
            if (jrow < n) *reinterpret_cast<uint4*>(mats[m] + (size_t)jrow * W + 4 * A) = v;
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

// grid (ceil(n / 8), batch): warp per row writes (column, SC value) for every set bit, in bit order.  Two passes, so that the
// expensive part is balanced over the lanes: (1) the set bits of the row become its column list (a lane owns a word, a warp scan
// gives the positions), (2) the lanes stride over that list and evaluate one SC value each - instead of every lane looping
// over the bits of its own word while the others wait.
__global__ void __launch_bounds__(256)
csr_fill_kernel(const Pt* __restrict__ P, const uint32_t* __restrict__ hard, int n, int W, float d_sq, Csr csr) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n || !csr.ok[b]) return;
    P += (size_t)b * n;
    const uint32_t* row = hard + ((size_t)b * n + i) * W;
    uint16_t* cols = csr.cols + (size_t)b * csr.cap;
    float* vals = csr.vals + (size_t)b * csr.cap;
    const uint32_t begin = csr.rowptr[(size_t)b * (n + 1) + i], end = csr.rowptr[(size_t)b * (n + 1) + i + 1];
    uint32_t base = begin;
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        uint32_t m = w < W ? __ldg(row + w) : 0u;
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
            cols[pos++] = (uint16_t)(w * 32 + bit);
        }
        base += (uint32_t)__shfl_sync(0xffffffffu, x, 31);
    }
    __syncwarp();                                   // the warp's column list is visible to all of its lanes
    const Pt me = load_pt(P + i);
    for (uint32_t p = begin + lane; p < end; p += 32) {
        const int j = cols[p];
        const float cd = cross_dist(me, load_pt(P + j));
        vals[p] = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(cd, cd), d_sq)), 0.f);
    }
}

// The single-pass form of csr_fill_kernel (every lane walks the bits of its own word and evaluates the SC value on the spot):
// kept as the reference the two-pass kernel is tested against bit for bit (eyoc_debug_sc2_reference_kernels bit 0).
__global__ void __launch_bounds__(256)
csr_fill_serial_kernel(const Pt* __restrict__ P, const uint32_t* __restrict__ hard, int n, int W, float d_sq, Csr csr) {
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

int g_sc2_reference_kernels = 0;    // eyoc_debug_sc2_reference_kernels: bit 0 = csr_fill_serial_kernel, bit 1 = seed_fitness_shared_kernel

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

// The whole power iteration of one pair in ONE launch: a cluster of PF_C CTAs per pair, each owning n / PF_C rows.
// A CTA keeps the iterate v (all n entries) and as much of its CSR slice as fits in shared memory (the rest streams from
// L2), computes u = M v for its rows with 8 lanes per row, and pushes every u_i into all PF_C CTAs' shared memory
// (distributed shared memory stores); after one cluster barrier each CTA normalises the full u and applies the
// torch.allclose stopping rule redundantly, so all CTAs of the cluster take the same decision and no global round trip
// or relaunch separates two iterations.  The per-row sums, the norm and the division are evaluated in exactly the order
// of power_step_kernel (lane l of its warp = accumulator l / 8 of lane l % 8 here; the xor butterfly is commutative), so
// the two kernels return the same bits.  u is double-buffered: a CTA reaches its pushes of iteration t + 2 only after the
// barrier of t + 1, which every CTA passes after it has finished reading the buffer of iteration t.
constexpr int PF_C = 8;
constexpr int PF_NT = 512;
int g_power_fused = 1;              // eyoc_debug_sc2_power_fused: 0 = one power_step_kernel launch per iteration

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int PF_SMEM_MAX = 227 * 1024;

struct PfArgs {
    const Pt* P;
    const uint32_t* hard;
    int n, W;
    float d_sq;
    int num_iterations;
    float* conf;
    int* done;
    int* iters;
    Csr csr;
    int rows_per;       // ceil(n / PF_C)
    int n_pad;          // n rounded up to 4
    int cache_cap;      // CSR entries of the slice kept in shared memory
};

__device__ __forceinline__ float pf_bits_lane(const Pt* P, const uint32_t* row, const Pt& me, int W, float d_sq, int vlane,
                                              const float* v) {
    float acc = 0.f;
    for (int w = vlane; w < W; w += 32) {
        uint32_t m = __ldg(row + w);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int j = w * 32 + bit;
            const float c = cross_dist(me, load_pt(P + j));
            const float sc = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(c, c), d_sq)), 0.f);
            acc = __fmaf_rn(sc, v[j], acc);
        }
    }
    return acc;
}

__global__ void __cluster_dims__(PF_C, 1, 1) __launch_bounds__(PF_NT, 1)
power_fused_kernel(const PfArgs a) {
    extern __shared__ __align__(16) unsigned char pf_smem[];
    const int b = blockIdx.y, n = a.n;
    const unsigned rank = cluster_ctarank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* v = reinterpret_cast<float*>(pf_smem);
    float* ubuf = v + a.n_pad;                                  // [2][n_pad]
    uint32_t* rp = reinterpret_cast<uint32_t*>(ubuf + 2 * a.n_pad);          // [rows_per + 1]
    float* cvals = reinterpret_cast<float*>(rp + ((a.rows_per + 1 + 3) & ~3));
    uint16_t* ccols = reinterpret_cast<uint16_t*>(cvals + a.cache_cap);
    __shared__ double red[8];
    __shared__ int cached_rows_s;

    const Pt* P = a.P + (size_t)b * n;
    const uint32_t* hard = a.hard + (size_t)b * n * a.W;
    const int r0 = min(n, (int)rank * a.rows_per), r1 = min(n, r0 + a.rows_per);
    const bool ok = a.csr.ok[b] != 0;
    const uint16_t* gcols = a.csr.cols + (size_t)b * a.csr.cap;
    const float* gvals = a.csr.vals + (size_t)b * a.csr.cap;

    for (int k = tid; k < n; k += PF_NT) v[k] = 1.0f;
    if (tid == 0) cached_rows_s = 0;
    if (ok)
        for (int k = tid; k <= r1 - r0; k += PF_NT) rp[k] = a.csr.rowptr[(size_t)b * (n + 1) + r0 + k];
    __syncthreads();
    uint32_t s0 = 0;
    if (ok) {
        s0 = rp[0];
        for (int k = tid + 1; k <= r1 - r0; k += PF_NT)
            if (rp[k] - s0 <= (uint32_t)a.cache_cap) atomicMax(&cached_rows_s, k);
    }
    __syncthreads();
    const int cached_rows = cached_rows_s;                     // rows [r0, r0 + cached_rows) live in shared memory
    if (ok) {
        const uint32_t cnt = rp[cached_rows] - s0;
        for (uint32_t k = tid; k < cnt; k += PF_NT) {
            cvals[k] = __ldg(gvals + s0 + k);
            ccols[k] = __ldg(gcols + s0 + k);
        }
    }
    cluster_sync_all();                                         // every CTA of the cluster is resident before remote stores

    const int q = lane & 7, grp = lane >> 3;
    uint32_t u_remote;                                          // shared::cluster address of ubuf in CTA q
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(u_remote) : "r"((uint32_t)__cvta_generic_to_shared(ubuf)), "r"(q));
    for (int t = 1; t <= a.num_iterations; ++t) {
        const int cur = t & 1;
        for (int lb = warp * 4; lb < r1 - r0; lb += (PF_NT / 32) * 4) {          // warp-uniform trip count (full-mask shuffles)
            const int li = lb + grp, i = r0 + li;
            const bool live = li < r1 - r0;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (!live) {
            } else if (ok) {
                const uint32_t s = rp[li], e = rp[li + 1];
                if (li < cached_rows) {
                    for (uint32_t p = s - s0 + q, pe = e - s0; p < pe; p += 32) {
                        a0 = __fmaf_rn(cvals[p], v[ccols[p]], a0);
                        if (p + 8 < pe) a1 = __fmaf_rn(cvals[p + 8], v[ccols[p + 8]], a1);
                        if (p + 16 < pe) a2 = __fmaf_rn(cvals[p + 16], v[ccols[p + 16]], a2);
                        if (p + 24 < pe) a3 = __fmaf_rn(cvals[p + 24], v[ccols[p + 24]], a3);
                    }
                } else {
                    for (uint32_t p = s + q; p < e; p += 32) {
                        a0 = __fmaf_rn(__ldg(gvals + p), v[__ldg(gcols + p)], a0);
                        if (p + 8 < e) a1 = __fmaf_rn(__ldg(gvals + p + 8), v[__ldg(gcols + p + 8)], a1);
                        if (p + 16 < e) a2 = __fmaf_rn(__ldg(gvals + p + 16), v[__ldg(gcols + p + 16)], a2);
                        if (p + 24 < e) a3 = __fmaf_rn(__ldg(gvals + p + 24), v[__ldg(gcols + p + 24)], a3);
                    }
                }
            } else {
                const Pt me = load_pt(P + i);
                const uint32_t* row = hard + (size_t)i * a.W;
                a0 = pf_bits_lane(P, row, me, a.W, a.d_sq, q, v);
                a1 = pf_bits_lane(P, row, me, a.W, a.d_sq, q + 8, v);
                a2 = pf_bits_lane(P, row, me, a.W, a.d_sq, q + 16, v);
                a3 = pf_bits_lane(P, row, me, a.W, a.d_sq, q + 24, v);
            }
            float acc = __fadd_rn(__fadd_rn(a0, a2), __fadd_rn(a1, a3));       // butterfly offsets 16, 8
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (live) asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(u_remote + ((uint32_t)cur * a.n_pad + (uint32_t)i) * 4u), "f"(acc) : "memory");
        }
        __syncwarp();
        cluster_sync_all();
        const float* u = ubuf + (size_t)cur * a.n_pad;
        if (tid < 256) {
            double ss = 0.0;
            for (int k = tid; k < n; k += 256) {
                const float x = u[k];
                ss += (double)x * (double)x;
            }
            ss = warp_sum_d(ss);
            if (lane == 0) red[warp] = ss;
        }
        __syncthreads();
        double tot = 0.0;
        for (int k = 0; k < 8; ++k) tot += red[k];
        const float denom = __fadd_rn((float)sqrt(tot), 1e-6f);
        int notclose = 0;
        for (int k = tid; k < n; k += PF_NT) {
            const float x = __fdiv_rn(u[k], denom);
            const float vp = v[k];
            const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, vp)));
            if (!(fabsf(__fsub_rn(x, vp)) <= allowed)) notclose = 1;
            v[k] = x;
        }
        const int nc = __syncthreads_or(notclose);
        if (nc == 0 || t == a.num_iterations) {
            for (int k = r0 + tid; k < r1; k += PF_NT) a.conf[(size_t)b * n + k] = v[k];
            if (rank == 0 && tid == 0) { a.iters[b] = t; a.done[b] = 1; }
            break;
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
    int keycap;           // candidate capacity of the first-tier launch (shared-memory keys)
    int* big;             // [2 + batch * S]: number of deferred seeds | work cursor | their flat indices b * S + s
};

// A seed's candidate columns are the set bits of its `hard` row - a few dozen on LiDAR data (0.4 % of the matrix), n at
// most.  The kernel is bound by per-CTA latency chains, not by traffic, so what matters is how many seeds an SM works on at
// once: 128-thread CTAs whose candidate keys take KEYCAP words of shared memory (16 CTAs per SM) handle every seed with
// <= KEYCAP candidates; the rare denser ones are deferred to a second launch sized for n candidates.
constexpr int SC_NT = 128;
constexpr int SC_KEYCAP = 1024;

// returns false (nothing written) when the seed has more candidates than `keycap`
__device__ __forceinline__ bool seed_consensus_body(const SeedArgs& a, const int b, const int s, uint32_t* sm, const int keycap) {
    const int n = a.n, W = a.W;
    uint32_t* trow = sm;                 // [W]
    uint32_t* hrow = sm + W;             // [W]
    uint32_t* nzmap = sm + 2 * W;        // [W] columns with a non-zero SC2 value
    uint32_t* keys = sm + 3 * W;         // [keycap]  (count << 16) | (65535 - j)
    __shared__ int ncand, npos, nnz, nwin, ntw;
    __shared__ unsigned int hist[256], sel_bin, sel_rem;
    __shared__ uint32_t win[MAXK];
    __shared__ int idx1[MAXK], idx2[MAXK], fine[MAXK], lval[MAXK];
    __shared__ uint32_t lhard[MAXK];
    __shared__ float ls[MAXK][3], lt[MAXK][3];
    __shared__ float M[MAXK][MAXK + 1];

    const Pt* P = a.P + (size_t)b * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k1 = a.k1, k2 = a.k2;

    if (tid == 0) { ncand = 0; npos = 0; nnz = 0; nwin = 0; ntw = 0; }
    for (int w = tid; w < W; w += SC_NT) nzmap[w] = 0;
    __syncthreads();
    int nc;
    if (a.sc2_dense == nullptr) {
        const uint32_t* tight = a.tight + (size_t)b * n * W;
        const int seed = a.seeds[(size_t)b * a.S + s];
        int cnt = 0;
        for (int w = tid; w < W; w += SC_NT) {
            const uint32_t h = a.hard[((size_t)b * n + seed) * W + w];
            trow[w] = tight[(size_t)seed * W + w];
            hrow[w] = h;
            cnt += __popc(h);
        }
        cnt = warp_sum_i(cnt);
        if (lane == 0 && cnt) atomicAdd(&ncand, cnt);
        __syncthreads();
        nc = ncand;
        if (nc > keycap) return false;                    // CTA-uniform
        // 1. compact the columns where hard[seed] is set
        for (int w = tid; w < W; w += SC_NT) {
            uint32_t m = hrow[w];
            const int c = __popc(m);
            if (c) {
                int pos = atomicAdd(&npos, c);
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    keys[pos++] = 65535u - (uint32_t)(w * 32 + bit);
                }
            }
        }
        __syncthreads();
        // 2. SC2[seed, j] = popc(tight[seed] & tight[j]) for those columns.  tight[seed] is as sparse as hard[seed] (a few
        //    dozen bits on LiDAR data): its non-zero words are compacted (hrow is free again) and, when there are few of
        //    them, every THREAD takes a column and walks that list - ~6 instructions per (column, word) and no shuffle
        //    reduction, against ~80 instructions per column for a warp sweeping all W words.  Dense seed rows keep the
        //    coalesced warp-per-column sweep (a thread-per-column walk would fetch a 32-byte sector per word).
        uint32_t* twi = hrow;                            // word indices of the non-zero words of trow
        for (int w = tid; w < W; w += SC_NT)
            if (trow[w]) twi[atomicAdd(&ntw, 1)] = (uint32_t)w;
        __syncthreads();
        const int nzw = ntw;
        int nz = 0;
        if (nzw <= 64) {
            for (int c = tid; c < nc; c += SC_NT) {
                const uint32_t jj = 65535u - keys[c];
                const uint32_t* row = tight + (size_t)jj * W;
                int cn = 0;
#pragma unroll 4
                for (int e = 0; e < nzw; ++e) {
                    const uint32_t w = twi[e];
                    cn += __popc(trow[w] & __ldg(row + w));
                }
                if (cn > 0) {
                    keys[c] = ((uint32_t)cn << 16) | (65535u - jj);
                    atomicOr(&nzmap[jj >> 5], 1u << (jj & 31));
                    ++nz;
                } else {
                    keys[c] = 0u;
                }
            }
            nz = warp_sum_i(nz);
        } else {
            for (int c = warp; c < nc; c += SC_NT / 32) {
                const uint32_t jj = 65535u - keys[c];
                const uint32_t* row = tight + (size_t)jj * W;
                int cn = 0;
                for (int w = lane; w < W; w += 32) cn += __popc(trow[w] & __ldg(row + w));
                cn = warp_sum_i(cn);
                if (lane == 0) {
                    if (cn > 0) {
                        keys[c] = ((uint32_t)cn << 16) | (65535u - jj);
                        atomicOr(&nzmap[jj >> 5], 1u << (jj & 31));
                        ++nz;
                    } else {
                        keys[c] = 0u;
                    }
                }
            }
        }
        if (lane == 0 && nz) atomicAdd(&nnz, nz);
    } else {
        // dense rows handed in by the caller: every column is a candidate; the values are the integer counts of SC2_PCR.py:363
        if (n > keycap) return false;
        const float* row = a.sc2_dense + ((size_t)b * a.S + s) * n;
        nc = n;
        int nz = 0, bad = 0;
        for (int j = tid; j < n; j += SC_NT) {
            const float v = __ldg(row + j);
            const int cn = (int)v;
            if (!(v >= 0.f && v <= 65535.f) || (float)cn != v) bad = 1;
            if (cn > 0 && !bad) {
                keys[j] = ((uint32_t)cn << 16) | (65535u - (uint32_t)j);
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
                for (int i = tid; i < 256; i += SC_NT) hist[i] = 0;
                __syncthreads();
                for (int c = tid; c < nc; c += SC_NT) {
                    const uint32_t key = keys[c];
                    if (key != 0u && (key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (warp == 0) {          // the bin where the count of keys in higher bins first reaches `remaining`
                    unsigned int h[8], tot = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) { h[k] = hist[8 * lane + k]; tot += h[k]; }
                    unsigned int incl = tot;              // inclusive suffix sum over the lanes (towards higher bins)
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int y = __shfl_down_sync(0xffffffffu, incl, o);
                        if (lane + o < 32) incl += y;
                    }
                    unsigned int above = incl - tot;
#pragma unroll
                    for (int k = 7; k >= 0; --k) {
                        if (above < remaining && remaining <= above + h[k]) { sel_bin = (unsigned int)(8 * lane + k); sel_rem = remaining - above; }
                        above += h[k];
                    }
                }
                __syncthreads();
                prefix |= sel_bin << shift;
                pmask |= 255u << shift;
                remaining = sel_rem;
            }
        } else {
            prefix = 1u;                                    // every non-zero key wins
        }
        for (int c = tid; c < nc; c += SC_NT) {
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
    for (int e = tid; e < k1 * k1; e += SC_NT) {       // symmetric bit for bit ((a-b)^2 == (b-a)^2): upper triangle only
        const int p = e / k1, q = e % k1;
        if (q < p) continue;
        const float ds = dist3_sum(ls[p][0], ls[p][1], ls[p][2], ls[q][0], ls[q][1], ls[q][2]);
        const float dt = dist3_sum(lt[p][0], lt[p][1], lt[p][2], lt[q][0], lt[q][1], lt[q][2]);
        if (fabsf(__fsub_rn(ds, dt)) < a.d_thre) {
            atomicOr(&lhard[p], 1u << q);
            if (q != p) atomicOr(&lhard[q], 1u << p);
        }
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
    // 5. soft measure on the k2 (SC2_PCR.py:117-131), diagonal zeroed
    for (int e = tid; e < k2 * k2; e += SC_NT) {
        const int p = e / k2, q = e % k2;
        if (q < p) continue;
        if (q == p) { M[p][p] = 0.f; continue; }
        const int fp = fine[p], fq = fine[q];
        const float ds = dist3_sum(ls[fp][0], ls[fp][1], ls[fp][2], ls[fq][0], ls[fq][1], ls[fq][2]);
        const float dt = dist3_sum(lt[fp][0], lt[fp][1], lt[fp][2], lt[fq][0], lt[fq][1], lt[fq][2]);
        const float c = fabsf(__fsub_rn(ds, dt));
        const float v = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(c, c), a.d_sq)), 0.f);
        M[p][q] = v;
        M[q][p] = v;
    }
    __syncthreads();
    // 6. power iteration on the k2 x k2 matrix by one warp, the lane's matrix row in registers; every iterate is kept
    if (warp == 0) {
        float mrow[MAXK];
#pragma unroll
        for (int q = 0; q < MAXK; ++q) mrow[q] = (lane < k2 && q < k2) ? M[lane][q] : 0.f;
        float v = 1.0f;
        float* out = a.local_v + ((size_t)b * a.S + s) * a.num_iterations * MAXK;
        int* notclose = a.local_notclose + (size_t)b * (a.num_iterations + 1);
        for (int t = 1; t <= a.num_iterations; ++t) {
            float uu = 0.f;
#pragma unroll
            for (int q = 0; q < MAXK; ++q) {
                const float vq = __shfl_sync(0xffffffffu, v, q);
                if (q < k2) uu = __fmaf_rn(mrow[q], vq, uu);              // rows >= k2 hold zeros
            }
            const float ss = warp_sum(lane < k2 ? __fmul_rn(uu, uu) : 0.f);
            const float vn = __fdiv_rn(uu, __fadd_rn(__fsqrt_rn(ss), 1e-6f));
            const float allowed = __fadd_rn(1e-8f, fabsf(__fmul_rn(1e-5f, v)));
            const bool close = (lane >= k2) || (fabsf(__fsub_rn(vn, v)) <= allowed);
            if (!__all_sync(0xffffffffu, close) && lane == 0) atomicAdd(notclose + t, 1);
            if (lane < k2) out[(size_t)(t - 1) * MAXK + lane] = vn;
            v = vn;
        }
    }
    return true;
}

__global__ void __launch_bounds__(SC_NT)
seed_consensus_kernel(SeedArgs a) {
    extern __shared__ uint32_t sm[];
    const int b = blockIdx.y, s = blockIdx.x;
    if (!seed_consensus_body(a, b, s, sm, a.keycap) && threadIdx.x == 0) a.big[2 + atomicAdd(a.big, 1)] = b * a.S + s;
}

// second tier: the deferred seeds (more than SC_KEYCAP candidates), keys sized for n; persistent CTAs walk the list
__global__ void __launch_bounds__(SC_NT)
seed_consensus_big_kernel(SeedArgs a) {
    extern __shared__ uint32_t sm[];
    __shared__ int item;
    const int total = a.big[0];
    for (;;) {
        __syncthreads();                      // the previous seed's shared state is no longer read
        if (threadIdx.x == 0) item = atomicAdd(a.big + 1, 1);
        __syncthreads();
        const int it = item;
        if (it >= total) break;
        const int g = a.big[2 + it];
        seed_consensus_body(a, g / a.S, g % a.S, sm, a.n);
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
    float inlier_s0;       // sqrt_threshold(inlier_threshold)
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

// Inlier counts of the seed hypotheses (SC2_PCR.py:150-158): FS seeds per CTA share every point load; their transforms live in
// registers (the loop is bound by the fp32 pipe: 20 dependent-rounding operations per point and seed, nothing else to issue).
constexpr int FS = 4;
__global__ void __launch_bounds__(256)
seed_fitness_kernel(FitArgs a) {
    __shared__ int cnt[8][FS];
    const int b = blockIdx.y, s0 = blockIdx.x * FS;
    const Pt* P = a.P + (size_t)b * a.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float T[FS][12];
#pragma unroll
    for (int k = 0; k < FS; ++k) {
        const float4* src = reinterpret_cast<const float4*>(a.seed_trans + ((size_t)b * a.S + min(s0 + k, a.S - 1)) * 16);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float4 v = __ldg(src + q);
            T[k][4 * q] = v.x; T[k][4 * q + 1] = v.y; T[k][4 * q + 2] = v.z; T[k][4 * q + 3] = v.w;
        }
    }
    int c[FS];
#pragma unroll
    for (int k = 0; k < FS; ++k) c[k] = 0;
    for (int j = tid; j < a.n; j += 256) {
        const Pt p = load_pt(P + j);
#pragma unroll
        for (int k = 0; k < FS; ++k) {
            float x, y, z;
            apply_T(T[k], p.sx, p.sy, p.sz, x, y, z);
            // torch.norm(.) < thr  <=>  sum of squares < s0: sqrt_rn is monotonic and s0 (host: sqrt_threshold) is the
            // smallest fp32 whose correctly rounded root reaches thr - the root itself is never formed
            const float dx = __fsub_rn(x, p.tx), dy = __fsub_rn(y, p.ty), dz = __fsub_rn(z, p.tz);
            c[k] += !(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) >= a.inlier_s0);
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

// The previous form of seed_fitness_kernel (8 seeds per CTA, transforms in shared memory): the reference the register version
// is tested against (eyoc_debug_sc2_reference_kernels bit 1).
__global__ void __launch_bounds__(256)
seed_fitness_shared_kernel(FitArgs a) {
    constexpr int FS8 = 8;
    __shared__ float T[FS8][12];
    __shared__ int cnt[8][FS8];
    const int b = blockIdx.y, s0 = blockIdx.x * FS8;
    const Pt* P = a.P + (size_t)b * a.n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < FS8 * 12) {
        const int s = min(s0 + tid / 12, a.S - 1);
        T[tid / 12][tid % 12] = a.seed_trans[((size_t)b * a.S + s) * 16 + tid % 12];
    }
    __syncthreads();
    int c[FS8];
#pragma unroll
    for (int k = 0; k < FS8; ++k) c[k] = 0;
    for (int j = tid; j < a.n; j += 256) {
        const Pt p = load_pt(P + j);
#pragma unroll
        for (int k = 0; k < FS8; ++k) {
            float x, y, z;
            apply_T(T[k], p.sx, p.sy, p.sz, x, y, z);
            const float dx = __fsub_rn(x, p.tx), dy = __fsub_rn(y, p.ty), dz = __fsub_rn(z, p.tz);
            c[k] += !(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) >= a.inlier_s0);
        }
    }
#pragma unroll
    for (int k = 0; k < FS8; ++k) {
        const int v = warp_sum_i(c[k]);
        if (lane == 0) cnt[warp][k] = v;
    }
    __syncthreads();
    if (tid < FS8 && s0 + tid < a.S) {
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
    const size_t W = ((N + 31) / 32 + 3) / 4 * 4;          // words per bit row, padded to 16 bytes (first_order_bits_kernel)
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
    L->big = c.off; c.take<int>(2 + B * S);          // deferred seeds of seed_consensus (count | cursor | list)
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
            first_order_bits_kernel<<<dim3((n + FO_COLS - 1) / FO_COLS, (n + 127) / 128, batch), 128, 0, stream>>>(
                P, n, W, cfg->d_thre, cfg->d_thre_half, sqrt_threshold(cfg->nms_radius), hard, tight, near);
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
            if (g_sc2_reference_kernels & 1) csr_fill_serial_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(P, hard, n, W, cfg->d_thre_sq, csr);
            else csr_fill_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(P, hard, n, W, cfg->d_thre_sq, csr);
            EYOC_LAUNCH_CHECK();
            const int rows_per = (n + PF_C - 1) / PF_C, n_pad = (n + 3) & ~3;
            const size_t pf_fixed = (size_t)3 * n_pad * 4 + (size_t)((rows_per + 1 + 3) & ~3) * 4;
            if (g_power_fused && pf_fixed + 6 * 1024 <= (size_t)PF_SMEM_MAX - 256) {
                const int cache_cap = (int)(((size_t)PF_SMEM_MAX - 256 - pf_fixed) / 6) & ~7;
                const size_t smem = pf_fixed + (size_t)cache_cap * 6;
                PfArgs pa{P, hard, n, W, cfg->d_thre_sq, I, conf, done, global_iters, csr, rows_per, n_pad, cache_cap};
                EYOC_CUDA(cudaFuncSetAttribute(power_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                power_fused_kernel<<<dim3(PF_C, batch), PF_NT, smem, stream>>>(pa);
                EYOC_LAUNCH_CHECK();
            } else {
                for (int t = 1; t <= I; ++t) {
                    power_step_kernel<<<dim3((n + 7) / 8, batch), 256, 0, stream>>>(P, hard, n, W, cfg->d_thre_sq, t, I, vbuf, u, conf, st, csr);
                    EYOC_LAUNCH_CHECK();
                }
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
        const int keycap = sc2_dense ? n : (n < SC_KEYCAP ? n : SC_KEYCAP);
        SeedArgs sa{P, hard, tight, seeds_use, sc2_dense, status, n, W, S, L.k1, L.k2, I, cfg->d_thre, cfg->d_thre_sq, topk1, topk2,
                    local_v, local_notclose, keycap, (int*)(ws + L.big)};
        const size_t smem = (size_t)(3 * W + keycap) * sizeof(uint32_t);
        EYOC_CUDA(cudaFuncSetAttribute(seed_consensus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        seed_consensus_kernel<<<dim3(S, batch), SC_NT, smem, stream>>>(sa);
        EYOC_LAUNCH_CHECK();
        if (keycap < n) {                      // seeds with more than keycap candidates (none on typical data): exits at once
            const size_t smem_big = (size_t)(3 * W + n) * sizeof(uint32_t);
            EYOC_CUDA(cudaFuncSetAttribute(seed_consensus_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
            seed_consensus_big_kernel<<<296, SC_NT, smem_big, stream>>>(sa);
            EYOC_LAUNCH_CHECK();
        }
        FitArgs fa{P, topk2, local_v, local_notclose, n, S, L.k2, I, cfg->inlier_threshold, sqrt_threshold(cfg->inlier_threshold), seed_weights, seed_trans,
                   fitness ? fitness : scores /* scratch */, local_iters};
        if (!fitness) {
            EYOC_CHECK_ARG(S <= n, "unreachable");
        }
        seed_kabsch_kernel<<<(unsigned)((batch * S + 127) / 128), 128, 0, stream>>>(fa, batch);
        EYOC_LAUNCH_CHECK();
        if (g_sc2_reference_kernels & 2) seed_fitness_shared_kernel<<<dim3((S + 7) / 8, batch), 256, 0, stream>>>(fa);
        else seed_fitness_kernel<<<dim3((S + FS - 1) / FS, batch), 256, 0, stream>>>(fa);
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

extern "C" int eyoc_debug_sc2_power_fused(int on) {
    g_power_fused = on ? 1 : 0;
    return EYOC_OK;
}

extern "C" int eyoc_debug_sc2_reference_kernels(int mask) {
    g_sc2_reference_kernels = mask & 3;
    return EYOC_OK;
}
