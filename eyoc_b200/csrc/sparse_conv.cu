// Generalized sparse convolution, output-stationary gather-GEMM with fused epilogue (sm_100a).
//
// Replaces ME.MinkowskiConvolution / MinkowskiConvolutionTranspose (+ the MinkowskiBatchNorm, MEF.relu,
// residual add, ME.cat and the final L2 normalisation that follow them) in the reference's
// model/resunet.py:142-193 and model/residual_block.py:37-53:
//     out[o, :] = epilogue( sum_k  in[nbr[k, o], :] @ W[k] )
// The accumulation order per output element is fixed: kernel offsets k ascending (MinkowskiEngine's
// order), input channels ascending, one fp32 FMA each.  Epilogue, in order: per-channel affine (folded
// eval-mode BatchNorm, or bias), residual add, ReLU, row L2 normalisation (no epsilon, resunet.py:189).
// A concatenated input (ME.cat, resunet.py:168,175,182) is read from its two sources in place.
//
// v1 data path: one CTA owns 128 output rows x BN output channels.  For each kernel offset with at least
// one neighbour in the tile, the gathered input rows are staged c-major in shared memory in 32-channel
// chunks next to the matching W[k] slab and multiplied with an (8|4)x4 register tile per thread.
#include "common.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int BM = 128;   // output rows per CTA
constexpr int BK = 32;    // input channels per smem chunk
constexpr int NT = 256;

struct ConvArgs {
    const float* in0; int c0;
    const float* in1; int c1;
    const int32_t* nbr;       // [K, n_out] or null (identity, K == 1)
    const int32_t* row_perm;  // [n_out] or null
    const float* weight;      // [K, c0 + c1, cout]
    const float* scale;       // [cout] or null
    const float* shift;       // [cout] or null
    const float* residual;    // [n_out, cout] or null
    float* out;               // [n_out, cout]
    int K, n_out, cout, relu, l2norm;
};

template <int BN, int TM>
__global__ void __launch_bounds__(NT, 2)
sparse_conv_tiled_kernel(ConvArgs a) {
    static_assert((BN / 4) * (BM / TM) == NT, "thread tiling must cover the CTA tile");
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Ws[BK][BN];
    __shared__ int rows[BM];
    __shared__ int idx[BM];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / 4), ty = tid / (BN / 4);
    const int row0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int cin = a.c0 + a.c1;

    if (tid < BM) {
        const int r = row0 + tid;
        rows[tid] = r < a.n_out ? (a.row_perm ? a.row_perm[r] : r) : -1;
    }
    __syncthreads();

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < a.K; ++k) {
        int valid = 0;
        if (tid < BM) {
            const int r = rows[tid];
            int v = -1;
            if (r >= 0) v = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_out + r) : r;
            idx[tid] = v;
            valid = v >= 0;
        }
        if (!__syncthreads_or(valid)) continue;      // no neighbour at this offset anywhere in the tile
        for (int cc = 0; cc < cin; cc += BK) {
            const float* src = cc < a.c0 ? a.in0 : a.in1;
            const int cs = cc < a.c0 ? a.c0 : a.c1;
            const int co = cc < a.c0 ? cc : cc - a.c0;
            // gather: lane -> row (conflict-free c-major stores), 4 channels per load
            {
                const int r = tid & (BM - 1);
                const int v = idx[r];
#pragma unroll
                for (int c4 = (tid / BM) * 4; c4 < BK; c4 += (NT / BM) * 4) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (v >= 0) x = __ldg(reinterpret_cast<const float4*>(src + (size_t)v * cs + co + c4));
                    As[c4 + 0][r] = x.x; As[c4 + 1][r] = x.y; As[c4 + 2][r] = x.z; As[c4 + 3][r] = x.w;
                }
            }
            // weight slab W[k][cc:cc+BK][n0:n0+BN]
            for (int e = tid; e < BK * BN / 4; e += NT) {
                const int c = e / (BN / 4), q = e % (BN / 4);
                *reinterpret_cast<float4*>(&Ws[c][q * 4]) =
                    __ldg(reinterpret_cast<const float4*>(a.weight + ((size_t)k * cin + cc + c) * a.cout + n0 + q * 4));
            }
            __syncthreads();
#pragma unroll 8
            for (int c = 0; c < BK; ++c) {
                float av[TM];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[c][ty * TM + i]);
                    av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
                }
                const float4 bv = *reinterpret_cast<const float4*>(&Ws[c][tx * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = __fmaf_rn(av[i], bv.x, acc[i][0]);
                    acc[i][1] = __fmaf_rn(av[i], bv.y, acc[i][1]);
                    acc[i][2] = __fmaf_rn(av[i], bv.z, acc[i][2]);
                    acc[i][3] = __fmaf_rn(av[i], bv.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
    }
    // ---- fused epilogue
    const int col = n0 + tx * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.scale) sc = __ldg(reinterpret_cast<const float4*>(a.scale + col));
    if (a.shift) sh = __ldg(reinterpret_cast<const float4*>(a.shift + col));
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = rows[ty * TM + i];
        float4 y = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (a.scale) { y.x = __fmaf_rn(y.x, sc.x, sh.x); y.y = __fmaf_rn(y.y, sc.y, sh.y); y.z = __fmaf_rn(y.z, sc.z, sh.z); y.w = __fmaf_rn(y.w, sc.w, sh.w); }
        else if (a.shift) { y.x += sh.x; y.y += sh.y; y.z += sh.z; y.w += sh.w; }
        if (a.residual && r >= 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)r * a.cout + col));
            y.x += q.x; y.y += q.y; y.z += q.z; y.w += q.w;
        }
        if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        if (a.l2norm) {   // BN == cout: the BN/4 threads of a row are adjacent lanes
            float ss = y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
#pragma unroll
            for (int o = BN / 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float nrm = sqrtf(ss);
            y.x = __fdiv_rn(y.x, nrm); y.y = __fdiv_rn(y.y, nrm); y.z = __fdiv_rn(y.z, nrm); y.w = __fdiv_rn(y.w, nrm);
        }
        if (r >= 0) *reinterpret_cast<float4*>(a.out + (size_t)r * a.cout + col) = y;
    }
}

// Any channel counts (conv1: 1 -> 32 with 125 offsets): one warp per output row, lane = output channel.
__global__ void __launch_bounds__(NT)
sparse_conv_generic_kernel(ConvArgs a) {
    const int lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (r0 >= a.n_out) return;
    const int r = a.row_perm ? a.row_perm[r0] : r0;
    const int cin = a.c0 + a.c1;
    for (int n0 = 0; n0 < a.cout; n0 += 32) {
        const int col = n0 + lane;
        const bool on = col < a.cout;
        float acc = 0.f;
        for (int k = 0; k < a.K; ++k) {
            const int v = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_out + r) : r;
            if (v < 0) continue;
            const float* w = a.weight + (size_t)k * cin * a.cout + col;
            for (int c = 0; c < cin; ++c) {
                const float x = c < a.c0 ? __ldg(a.in0 + (size_t)v * a.c0 + c) : __ldg(a.in1 + (size_t)v * a.c1 + (c - a.c0));
                if (on) acc = __fmaf_rn(x, __ldg(w + (size_t)c * a.cout), acc);
            }
        }
        float y = acc;
        if (on) {
            if (a.scale) y = __fmaf_rn(y, a.scale[col], a.shift ? a.shift[col] : 0.f);
            else if (a.shift) y += a.shift[col];
            if (a.residual) y += a.residual[(size_t)r * a.cout + col];
            if (a.relu) y = fmaxf(y, 0.f);
        }
        if (a.l2norm) {   // cout <= 32 checked on the host
            const float ss = warp_sum(on ? y * y : 0.f);
            y = __fdiv_rn(y, sqrtf(ss));
        }
        if (on) a.out[(size_t)r * a.cout + col] = y;
    }
}

// First layer of the network (conv1: C_in = 1, K = 125, C_out = 32; model/resunet.py:31-37).  lane = output row, so
// the K neighbour-table reads are coalesced; W (K x 32 floats) sits in shared memory and is read by broadcast; each
// thread keeps its 32 output channels in registers and writes one full 128-byte row.
template <int COUT>
__global__ void __launch_bounds__(128)
sparse_conv_cin1_kernel(ConvArgs a) {
    extern __shared__ float wsm[];                    // [K][COUT]
    for (int e = threadIdx.x; e < a.K * COUT; e += 128) wsm[e] = a.weight[e];
    __syncthreads();
    const int r0 = blockIdx.x * 128 + threadIdx.x;
    if (r0 >= a.n_out) return;
    const int r = a.row_perm ? a.row_perm[r0] : r0;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    // The two dependent loads per offset (table entry, then the neighbour's feature) are issued five offsets at a time
    // so that their latencies overlap; the accumulation order (k ascending, missing neighbours skipped) is unchanged.
    constexpr int U = 5;
    for (int k0 = 0; k0 < a.K; k0 += U) {
        int v[U];
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = k0 + u < a.K ? (a.nbr ? __ldg(a.nbr + (size_t)(k0 + u) * a.n_out + r) : r) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = v[u] >= 0 ? __ldg(a.in0 + v[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v[u] < 0) continue;
            const float4* w4 = reinterpret_cast<const float4*>(wsm + (k0 + u) * COUT);
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) {
                const float4 w = w4[c];
                acc[4 * c + 0] = __fmaf_rn(x[u], w.x, acc[4 * c + 0]);
                acc[4 * c + 1] = __fmaf_rn(x[u], w.y, acc[4 * c + 1]);
                acc[4 * c + 2] = __fmaf_rn(x[u], w.z, acc[4 * c + 2]);
                acc[4 * c + 3] = __fmaf_rn(x[u], w.w, acc[4 * c + 3]);
            }
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        float y = acc[c];
        if (a.scale) y = __fmaf_rn(y, __ldg(a.scale + c), a.shift ? __ldg(a.shift + c) : 0.f);
        else if (a.shift) y += __ldg(a.shift + c);
        if (a.residual) y += a.residual[(size_t)r * COUT + c];
        if (a.relu) y = fmaxf(y, 0.f);
        ss = __fmaf_rn(y, y, ss);
        acc[c] = y;
    }
    const float nrm = a.l2norm ? sqrtf(ss) : 1.f;
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
        float4 y = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        if (a.l2norm) { y.x = __fdiv_rn(y.x, nrm); y.y = __fdiv_rn(y.y, nrm); y.z = __fdiv_rn(y.z, nrm); y.w = __fdiv_rn(y.w, nrm); }
        *reinterpret_cast<float4*>(a.out + (size_t)r * COUT + c) = y;
    }
}

}  // namespace

extern "C" int eyoc_sparse_conv(const float* in0, int c0, const float* in1, int c1, const int32_t* nbr, int K, int64_t n_out,
                                const int32_t* row_perm, const float* weight, const float* scale, const float* shift,
                                const float* residual, int relu, int l2norm, float* out, int cout, cudaStream_t stream) {
    EYOC_CHECK_ARG(in0 && weight && out, "eyoc_sparse_conv: null argument");
    EYOC_CHECK_ARG(c0 >= 1 && c1 >= 0 && cout >= 1 && K >= 1, "eyoc_sparse_conv: bad channel / kernel counts");
    EYOC_CHECK_ARG((in1 != nullptr) == (c1 > 0), "eyoc_sparse_conv: in1 and c1 must be given together");
    EYOC_CHECK_ARG(nbr || K == 1, "eyoc_sparse_conv: a neighbour table is required when K > 1");
    EYOC_CHECK_ARG(n_out >= 0 && n_out < (1ll << 31), "eyoc_sparse_conv: bad n_out");
    if (n_out == 0) return EYOC_OK;
    ConvArgs a{in0, c0, in1, c1, nbr, row_perm, weight, scale, shift, residual, out, K, (int)n_out, cout, relu, l2norm};
    const int cin = c0 + c1;
    const bool tiled = (cin % BK == 0) && (c0 % BK == 0) && (cout % 32 == 0) && (!l2norm || cout == 32 || cout == 64);
    if (tiled) {
        const unsigned gx = (unsigned)((n_out + BM - 1) / BM);
        if (cout % 64 == 0) {
            if (l2norm) EYOC_CHECK_ARG(cout == 64, "eyoc_sparse_conv: l2norm needs cout <= 64");
            sparse_conv_tiled_kernel<64, 8><<<dim3(gx, cout / 64), NT, 0, stream>>>(a);
        } else {
            if (l2norm) EYOC_CHECK_ARG(cout == 32, "eyoc_sparse_conv: l2norm needs the row in one tile");
            sparse_conv_tiled_kernel<32, 4><<<dim3(gx, cout / 32), NT, 0, stream>>>(a);
        }
    } else if (cin == 1 && cout == 32 && (size_t)K * 32 * 4 <= 48 * 1024) {
        sparse_conv_cin1_kernel<32><<<(unsigned)((n_out + 127) / 128), 128, (size_t)K * 32 * 4, stream>>>(a);
    } else {
        EYOC_CHECK_ARG(!l2norm || cout <= 32, "eyoc_sparse_conv: l2norm on the generic path needs cout <= 32");
        sparse_conv_generic_kernel<<<(unsigned)((n_out + NT / 32 - 1) / (NT / 32)), NT, 0, stream>>>(a);
    }
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
