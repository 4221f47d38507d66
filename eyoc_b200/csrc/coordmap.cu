// Coordinate maps and kernel maps for generalized sparse convolution (sm_100a).
//
// Replaces what MinkowskiEngine's coordinate manager does underneath the reference's
// ME.SparseTensor(F, coordinates=C) (scripts/test_kitti.py:143-147) and every
// ME.MinkowskiConvolution / MinkowskiConvolutionTranspose in model/resunet.py:31-140:
//   * an open-addressing hash of the (batch, x, y, z) int32 coordinates (64-bit packed key -> row index),
//   * the stride-2 coordinate sets  unique(floor(c / 2ts) * 2ts)  in first-occurrence order,
//   * neighbour tables  nbr[k, o] = row of  c_o + off_k * step  in the input map (or -1),
//     off_k enumerating {-r..r}^3 with x fastest (k = ix + K*(iy + K*iz)).
// All of it is integer work bounded by HBM/L2 latency; lookups are one probe sequence per (k, o) with
// coalesced writes along o.
#include "common.cuh"
#include "xh_format.cuh"
#include <cub/device/device_radix_sort.cuh>
#include "../../include/eyoc_b200.h"

namespace {

constexpr unsigned long long EMPTY = 0xffffffffffffffffull;

__device__ __forceinline__ bool in_range16(int v) { return v >= -32768 && v <= 32767; }

__device__ __forceinline__ unsigned long long pack4(int b, int x, int y, int z) {
    return ((unsigned long long)(unsigned int)(b & 0xffff) << 48) | ((unsigned long long)(unsigned int)((x + 32768) & 0xffff) << 32) |
           ((unsigned long long)(unsigned int)((y + 32768) & 0xffff) << 16) | (unsigned long long)(unsigned int)((z + 32768) & 0xffff);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

// insert key; value = min(row) over duplicates.  Returns the slot.
__device__ __forceinline__ long long hash_insert_min(unsigned long long* keys, int* vals, long long cap, unsigned long long key, int row) {
    long long slot = (long long)(mix64(key) & (unsigned long long)(cap - 1));
    while (true) {
        const unsigned long long prev = atomicCAS(keys + slot, EMPTY, key);
        if (prev == EMPTY || prev == key) {
            atomicMin(vals + slot, row);
            return slot;
        }
        slot = (slot + 1) & (cap - 1);
    }
}

__device__ __forceinline__ int hash_lookup(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long cap,
                                           unsigned long long key) {
    long long slot = (long long)(mix64(key) & (unsigned long long)(cap - 1));
    while (true) {
        const unsigned long long k = keys[slot];
        if (k == key) return vals[slot];
        if (k == EMPTY) return -1;
        slot = (slot + 1) & (cap - 1);
    }
}

// Lookups of one x-row of kernel offsets (KS keys) with all home-slot loads in flight together: the map kernels are
// bound by the latency of dependent random accesses (hash -> key -> value), so a thread that issues its KS independent
// key loads back to back, then its value loads, hides most of it.  Collisions fall back to linear probing.
template <int KS>
__device__ __forceinline__ void hash_lookup_row(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long cap,
                                                const unsigned long long (&key)[KS], const bool (&act)[KS], int (&v)[KS]) {
    long long slot[KS];
    unsigned long long got[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        slot[j] = (long long)(mix64(key[j]) & (unsigned long long)(cap - 1));
        got[j] = EMPTY;
    }
#pragma unroll
    for (int j = 0; j < KS; ++j)
        if (act[j]) got[j] = keys[slot[j]];
#pragma unroll
    for (int j = 0; j < KS; ++j) v[j] = -1;
#pragma unroll
    for (int j = 0; j < KS; ++j)
        if (act[j] && got[j] == key[j]) v[j] = vals[slot[j]];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        if (act[j] && got[j] != key[j] && got[j] != EMPTY) {          // home slot taken by another key: probe on
            long long sl = (slot[j] + 1) & (cap - 1);
            while (true) {
                const unsigned long long k = keys[sl];
                if (k == key[j]) { v[j] = vals[sl]; break; }
                if (k == EMPTY) break;
                sl = (sl + 1) & (cap - 1);
            }
        }
    }
}

__global__ void hash_clear_kernel(unsigned long long* keys, int* vals, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) { keys[i] = EMPTY; vals[i] = 0x7fffffff; }
}

// status[0] |= 1: coordinate out of the 16-bit packed range; |= 2: duplicate coordinate
__global__ void hash_build_kernel(const int* __restrict__ coords, int n, unsigned long long* keys, int* vals, long long cap,
                                  int* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    if (!(c.x >= 0 && c.x <= 65535 && in_range16(c.y) && in_range16(c.z) && in_range16(c.w))) { atomicOr(status, 1); return; }
    hash_insert_min(keys, vals, cap, pack4(c.x, c.y, c.z, c.w), i);
    // status[1] = largest batch index: atomics only when the running maximum actually moves
    if (c.x > *((volatile int*)(status + 1))) atomicMax(status + 1, c.x);
}

__global__ void hash_check_unique_kernel(const int* __restrict__ coords, int n, const unsigned long long* keys, const int* vals,
                                         long long cap, int* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    if (hash_lookup(keys, vals, cap, pack4(c.x, c.y, c.z, c.w)) != i) atomicOr(status, 2);
}

__device__ __forceinline__ int floor_to(int v, int ts) {   // floor(v / ts) * ts for negative v too
    int q = v / ts;
    if ((v % ts) != 0 && v < 0) --q;
    return q * ts;
}

// pass 1 of the stride-2 coordinate set: coarse key -> min fine row
__global__ void down_insert_kernel(const int* __restrict__ coords, int n, int ts_out, unsigned long long* keys, int* vals, long long cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    hash_insert_min(keys, vals, cap, pack4(c.x, floor_to(c.y, ts_out), floor_to(c.z, ts_out), floor_to(c.w, ts_out)), i);
}

// pass 2: flag first occurrences
__global__ void down_flag_kernel(const int* __restrict__ coords, int n, int ts_out, const unsigned long long* keys, const int* vals,
                                 long long cap, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    const unsigned long long key = pack4(c.x, floor_to(c.y, ts_out), floor_to(c.z, ts_out), floor_to(c.w, ts_out));
    flag[i] = hash_lookup(keys, vals, cap, key) == i ? 1 : 0;
}

// ---- exclusive scan (three small kernels; n is a few million at most)
constexpr int SCAN_B = 1024;
__global__ void scan_block_kernel(const int* __restrict__ in, int n, int* __restrict__ out, int* __restrict__ block_sums) {
    __shared__ int wsum[32];
    const int i = blockIdx.x * SCAN_B + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = i < n ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    const int base = warp > 0 ? wsum[warp - 1] : 0;
    if (i < n) out[i] = base + x - v;
    if (threadIdx.x == SCAN_B - 1) block_sums[blockIdx.x] = base + x;
}
__global__ void scan_sums_kernel(int* block_sums, int nblocks, int* total) {   // single CTA, sequential over chunks
    __shared__ int carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < nblocks; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const int base = carry + (warp > 0 ? wsum[warp - 1] : 0);
        if (i < nblocks) block_sums[i] = base + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 3: emit coarse coordinates in first-occurrence order and point the hash at the coarse row
__global__ void down_emit_kernel(const int* __restrict__ coords, int n, int ts_out, const int* __restrict__ flag,
                                 const int* __restrict__ pos, const int* __restrict__ block_sums, unsigned long long* keys, int* vals,
                                 long long cap, int* __restrict__ coords_out, int* __restrict__ sel_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    const int4 cc = make_int4(c.x, floor_to(c.y, ts_out), floor_to(c.z, ts_out), floor_to(c.w, ts_out));
    const int row = pos[i] + block_sums[i / SCAN_B];
    reinterpret_cast<int4*>(coords_out)[row] = cc;
    if (sel_out) sel_out[row] = i;
    const unsigned long long key = pack4(cc.x, cc.y, cc.z, cc.w);
    long long slot = (long long)(mix64(key) & (unsigned long long)(cap - 1));
    while (keys[slot] != key) slot = (slot + 1) & (cap - 1);
    vals[slot] = row;
}

// nbr[k, o] = row in the input map of  c_o + off_k * step.  One thread handles a whole x-row of kernel offsets
// (blockIdx.y = iy + KS * iz) with its KS lookups in flight together.
template <int KS>
__global__ void kernel_map_kernel(const int* __restrict__ out_coords, int n_out, const unsigned long long* __restrict__ keys,
                                  const int* __restrict__ vals, long long cap, int step, int* __restrict__ nbr) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int rj = blockIdx.y;
    if (o >= n_out) return;
    constexpr int r = (KS - 1) / 2;
    const int iy = rj % KS - r, iz = rj / KS - r;
    const int4 c = reinterpret_cast<const int4*>(out_coords)[o];
    const int y = c.z + iy * step, z = c.w + iz * step;
    const bool yz_ok = in_range16(y) && in_range16(z);
    unsigned long long key[KS];
    bool act[KS];
    int v[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const int x = c.y + (j - r) * step;
        act[j] = yz_ok && in_range16(x);
        key[j] = pack4(c.x, x, y, z);
    }
    hash_lookup_row<KS>(keys, vals, cap, key, act, v);
#pragma unroll
    for (int j = 0; j < KS; ++j) nbr[(size_t)(rj * KS + j) * n_out + o] = v[j];
}

// Voxelisation (lib/data_loaders.py:936-979): q = floor(xyz / voxel_size) evaluated as torch / numpy do in fp32 (true
// division, then floor), batch column from the optional per-point cloud index.  status bit 0: outside the 16-bit range.
__global__ void quantize_kernel(const float* __restrict__ xyz, const int* __restrict__ cloud, int n, float voxel_size,
                                int* __restrict__ q, int* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fx = floorf(__fdiv_rn(xyz[3 * (size_t)i], voxel_size)), fy = floorf(__fdiv_rn(xyz[3 * (size_t)i + 1], voxel_size)),
                fz = floorf(__fdiv_rn(xyz[3 * (size_t)i + 2], voxel_size));
    const int b = cloud ? cloud[i] : 0;
    const bool ok = fx >= -32768.f && fx <= 32767.f && fy >= -32768.f && fy <= 32767.f && fz >= -32768.f && fz <= 32767.f &&
                    b >= 0 && b <= 65535;
    if (!ok) atomicOr(status, 1);
    reinterpret_cast<int4*>(q)[i] = ok ? make_int4(b, (int)fx, (int)fy, (int)fz) : make_int4(0, 0, 0, 0);
}

// Stride-1 map of a coordinate set onto itself: offsets come in mirrored pairs (k, K^3-1-k) and i = nbr[k, o] <=>
// o = nbr[K^3-1-k, i], so only the first half of the offsets is probed; every hit is written twice (the table is
// pre-filled with -1 and the centre offset is the identity).
template <int KS>
__global__ void kernel_map_sym_kernel(const int* __restrict__ coords, int n, const unsigned long long* __restrict__ keys,
                                      const int* __restrict__ vals, long long cap, int step, int* __restrict__ nbr) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int rj = blockIdx.y;                      // x-row of offsets: k = rj * KS + j, rows 0 .. (K^3 / 2) / KS
    if (o >= n) return;
    constexpr int K3 = KS * KS * KS, r = (KS - 1) / 2;
    const int iy = rj % KS - r, iz = rj / KS - r;
    const int4 c = reinterpret_cast<const int4*>(coords)[o];
    const int y = c.z + iy * step, z = c.w + iz * step;
    const bool yz_ok = in_range16(y) && in_range16(z);
    unsigned long long key[KS];
    bool act[KS];
    int v[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const int x = c.y + (j - r) * step;
        act[j] = rj * KS + j < K3 / 2 && yz_ok && in_range16(x);        // first half of the offsets only
        key[j] = pack4(c.x, x, y, z);
    }
    hash_lookup_row<KS>(keys, vals, cap, key, act, v);
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const int k = rj * KS + j;
        if (k == K3 / 2) nbr[(size_t)k * n + o] = o;                     // the centre offset is the identity
        else if (k < K3 / 2 && v[j] >= 0) {
            nbr[(size_t)k * n + o] = v[j];
            nbr[(size_t)(K3 - 1 - k) * n + v[j]] = o;
        }
    }
}

// Transposed (fine <- coarse) map from the forward strided map between the same two levels:
// f = down[k, c]  <=>  c = up[k, f]  (same k, MinkowskiEngine's no-flip convention).
__global__ void kernel_map_transpose_kernel(const int* __restrict__ down, int n_coarse, int n_fine, int* __restrict__ up) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t k = blockIdx.y;
    if (c >= n_coarse) return;
    const int f = __ldg(down + k * n_coarse + c);
    if (f >= 0) up[k * n_fine + f] = c;
}

// rows of one map grouped by the parity class of (x, y, z) / ts  (8 classes): perm sorted by class,
// used to make the valid kernel offsets of a transposed stride-2 convolution CTA-uniform.
__global__ void parity_class_kernel(const int* __restrict__ coords, int n, int ts, int* __restrict__ cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4*>(coords)[i];
    cls[i] = (((c.y / ts) & 1)) | (((c.z / ts) & 1) << 1) | (((c.w / ts) & 1) << 2);
}

// Tile-order key of an output row: (cloud group << 27) | bit mask of the kernel offsets that have a neighbour.
// Rows sorted by this key put rows with the same neighbour pattern into the same 128-row tile, so most
// (tile, offset) items of the tensor-core convolution are either skipped or densely filled, while a group of
// clouds (the L2 working set of the gather) stays contiguous.
__global__ void tile_key_kernel(const int* __restrict__ nbr, int K, int n_out, const int* __restrict__ coords, int group,
                                unsigned long long* __restrict__ keys, int* __restrict__ iota, unsigned int* __restrict__ hist) {
    __shared__ unsigned int h_s[32];
    if (threadIdx.x < 32) h_s[threadIdx.x] = 0;
    __syncthreads();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int m = 0;
    if (o < n_out) {
        for (int k = 0; k < K; ++k) m |= (unsigned int)(__ldg(nbr + (size_t)k * n_out + o) >= 0) << k;
        const int b = coords[4 * (size_t)o];
        keys[o] = ((unsigned long long)(unsigned int)(b / group) << 27) | m;
        iota[o] = o;
    }
    // how many rows have a neighbour at each offset (decides the significance of the offsets in the sort key)
    for (int k = 0; k < K; ++k) {
        const unsigned int c = __popc(__ballot_sync(0xffffffffu, (m >> k) & 1u));
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&h_s[k], c);
    }
    __syncthreads();
    if (threadIdx.x < K && h_s[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h_s[threadIdx.x]);
}

// pos[k] = bit position of offset k in the sort key: offsets ordered by (count ascending, k ascending), the RAREST offset in
// the most significant bit.  Rows then cluster first by the offsets few rows have, and the 256-row tiles cut from the sorted
// order execute ~6 % fewer (tile, offset) items than with the offsets in natural significance (tools/tile_fill.py).
__global__ void tile_bit_order_kernel(const unsigned int* __restrict__ hist, int K, int* __restrict__ pos) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < K; ++k) {
        int rank = 0;
        for (int j = 0; j < K; ++j) rank += (hist[j] < hist[k]) || (hist[j] == hist[k] && j < k);
        pos[k] = K - 1 - rank;
    }
}

__global__ void tile_key_remap_kernel(unsigned long long* __restrict__ keys, int n_out, int K, const int* __restrict__ pos) {
    __shared__ int p_s[32];
    if (threadIdx.x < 32) p_s[threadIdx.x] = threadIdx.x < K ? pos[threadIdx.x] : 0;
    __syncthreads();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const unsigned long long key = keys[o];
    const unsigned int m = (unsigned int)(key & 0x7ffffffull);
    unsigned int r = 0;
    for (int k = 0; k < K; ++k) r |= ((m >> k) & 1u) << p_s[k];
    keys[o] = (key & ~0x7ffffffull) | r;
}

// nbr_tiled[k, i] = nbr[k, perm[i]]: grid (tiles of 256 rows, K).  The tile's "offset k has a neighbour" bit - what
// eyoc_tile_masks computes from the tiled table afterwards - is one block-wide OR of values already in registers.
__global__ void __launch_bounds__(256)
permute_columns_kernel(const int* __restrict__ nbr, int n_out, const int* __restrict__ perm, int* __restrict__ out,
                       unsigned int* __restrict__ masks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t k = blockIdx.y;
    int v = -1;
    if (i < n_out) {
        v = __ldg(nbr + k * n_out + __ldg(perm + i));
        out[k * n_out + i] = v;
    }
    if (masks) {
        const int any = __syncthreads_or(v >= 0);
        if (threadIdx.x == 0 && any) atomicOr(masks + blockIdx.x, 1u << blockIdx.y);
    }
}

// ------------------------------------------------------------------------------------------------- stem convolution
// The network's first layer (model/resunet.py:31-37: conv1, 1 -> 32 channels, 5^3 offsets, stride 1) without a 125-column
// neighbour table.  Probing the coordinate hash for all 124 offsets is ~85 % misses on surface-like clouds, so an
// OCCUPANCY table is built first: one 64-bit mask per 4x4x4 block of voxels (bit = x&3 | (y&3)<<2 | (z&3)<<4).  A KS^3
// neighbourhood (KS <= 5) spans at most 2x2x2 blocks, so a row reads eight masks (shared with its neighbours through L1),
// extracts the KS occupancy bits of every (dy, dz) line with two shifts, and probes the coordinate hash only where a
// voxel exists (every such probe hits).  The convolution accumulates in ascending k exactly like sparse_conv_cin1_kernel,
// the BN affine / ReLU epilogue and the split-half packing are applied in registers, and the 3^3 sub-cube of the
// neighbourhood leaves as the level's 3^3 neighbour table (identical to eyoc_kernel_map_self's).
__global__ void block_clear_kernel(ulonglong2* __restrict__ slots, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) slots[i] = make_ulonglong2(EMPTY, 0ull);
}

__global__ void block_insert_kernel(const int* __restrict__ coords, int n, ulonglong2* slots, long long cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int4 c = make_int4(0, 0, 0, 0);
    bool valid = i < n;
    if (valid) {
        c = reinterpret_cast<const int4*>(coords)[i];
        valid = c.x >= 0 && c.x <= 65535 && in_range16(c.y) && in_range16(c.z) && in_range16(c.w);
    }
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned long long key = pack4(c.x, c.y >> 2, c.z >> 2, c.w >> 2);
    const int bit = (c.y & 3) | ((c.z & 3) << 2) | ((c.w & 3) << 4);
    // rows of one warp that fall into the same block (neighbours in scan / sort order usually do) insert once
    const unsigned peers = __match_any_sync(active, key);
    const unsigned lo = __reduce_or_sync(peers, bit < 32 ? 1u << bit : 0u);
    const unsigned hi = __reduce_or_sync(peers, bit >= 32 ? 1u << (bit - 32) : 0u);
    if ((int)(threadIdx.x & 31) != __ffs(peers) - 1) return;
    const unsigned long long mask = ((unsigned long long)hi << 32) | lo;
    long long slot = (long long)(mix64(key) & (unsigned long long)(cap - 1));
    while (true) {
        const unsigned long long prev = atomicCAS(&slots[slot].x, EMPTY, key);
        if (prev == EMPTY || prev == key) {
            atomicOr(&slots[slot].y, mask);
            return;
        }
        slot = (slot + 1) & (cap - 1);
    }
}

struct StemArgs {
    const int* coords;
    int n;
    const unsigned long long* keys;
    const int* vals;
    long long cap;
    const ulonglong2* blocks;
    long long bcap;
    const float* in;        // [n] the single input channel; null = every value is 1.0
    const float* weight;    // [KS^3, 32]
    const float* scale;
    const float* shift;
    int relu;
    float* out;             // [n, 32] fp32, or
    uint8_t* out_xh;        // [n] split-half rows of 32 channels (128 bytes)
    int* range_status;
    int* nbr3;              // [27, n] or null
};

// Phase A of the stem kernels: the 2x2x2 block masks around the row (staged in ms[8][32], one column per lane) -> the
// KS^3-bit occupancy of its neighbourhood, bit k = ix + KS (iy + KS iz) in (occ_lo, occ_hi).
template <int KS>
__device__ __forceinline__ void stem_occupancy(const StemArgs& a, const int4 c, int lane, unsigned long long* ms,
                                               unsigned long long& occ_lo, unsigned long long& occ_hi) {
    constexpr int r = (KS - 1) / 2;
    const int bx0 = (c.y - r) >> 2, by0 = (c.z - r) >> 2, bz0 = (c.w - r) >> 2;
    const int bx1 = (c.y + r) >> 2, by1 = (c.z + r) >> 2, bz1 = (c.w + r) >> 2;
    unsigned long long bkey[8];
    ulonglong2 got[8];
    long long bslot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        bkey[j] = pack4(c.x, (j & 1) ? bx1 : bx0, (j & 2) ? by1 : by0, (j & 4) ? bz1 : bz0);
        bslot[j] = (long long)(mix64(bkey[j]) & (unsigned long long)(a.bcap - 1));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) got[j] = __ldg(a.blocks + bslot[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        unsigned long long m = 0ull;
        if (got[j].x == bkey[j]) m = got[j].y;
        else if (got[j].x != EMPTY) {
            long long sl = (bslot[j] + 1) & (a.bcap - 1);
            while (true) {
                const ulonglong2 g2 = __ldg(a.blocks + sl);
                if (g2.x == bkey[j]) { m = g2.y; break; }
                if (g2.x == EMPTY) break;
                sl = (sl + 1) & (a.bcap - 1);
            }
        }
        ms[j * 32 + lane] = m;
    }
    const int sx = (c.y - r) & 3;                              // x offset of the window inside block bx0
#pragma unroll
    for (int g = 0; g < KS * KS; ++g) {
        const int iy = g % KS - r, iz = g / KS - r;
        const int yy = c.z + iy, zz = c.w + iz;
        const int jy = (yy >> 2) - by0, jz = (zz >> 2) - bz0;                   // 0 or 1
        const int pos = ((yy & 3) << 2) | ((zz & 3) << 4);
        const unsigned long long m0 = ms[(jz * 4 + jy * 2) * 32 + lane], m1 = ms[(jz * 4 + jy * 2 + 1) * 32 + lane];
        const unsigned line = (unsigned)((m0 >> pos) & 0xFull) | ((unsigned)((m1 >> pos) & 0xFull) << 4);   // 8 voxels along x
        const unsigned long long bits = (line >> sx) & ((1u << KS) - 1u);
        const int bp = g * KS;                                 // compile-time after unrolling
        if (bp < 64) {
            occ_lo |= bits << bp;
            if (bp + KS > 64) occ_hi |= bits >> (64 - bp);
        } else {
            occ_hi |= bits << (bp - 64);
        }
    }
}

// One warp per 32 rows, 8 warps per CTA.
//   A (lane = row): the 2x2x2 block masks -> the row's KS^3-bit occupancy.
//   B: the set bits of the warp's rows become one list of (row, k) entries, rows in order, k ascending within a row.
//   C (lane = entry, 4 entries per lane in flight): hash probe (always a hit) -> neighbour row -> its input value; entries of
//      the 3^3 sub-cube also go to the warp's [27][32] slice of the neighbour table.
//   D (half-warp = row, lane = 2 output channels): each row's entries are accumulated in list order (= k ascending, the order
//      of sparse_conv_cin1_kernel), then BN affine / ReLU / store (fp32 row, or the split-half row: 64 B of hi, 64 B of lo').
// Only neighbours that exist cost instructions: ~18 of 124 on LiDAR surfaces, where a thread-per-row loop over all offsets
// pays for every offset that ANY of its warp's 32 rows has.
constexpr int STEM_NT = 256;
constexpr int STEM_ECAP = 960;          // list entries per pass (a row has <= 125: at least 7 rows per pass)
constexpr int STEM_WARP_BYTES = STEM_ECAP * 8 + 27 * 32 * 4 + 32 * 16 + 36 * 4;

template <int KS>
__global__ void __launch_bounds__(STEM_NT)
stem_conv_kernel(const StemArgs a) {
    constexpr int K3 = KS * KS * KS, r = (KS - 1) / 2, CO = 32, CENTRE = K3 / 2;
    static_assert(K3 <= 128, "entry encoding: 7 bits of k");
    extern __shared__ __align__(16) unsigned char stem_smem[];
    float* w_s = reinterpret_cast<float*>(stem_smem);                               // [K3][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* mine = stem_smem + (size_t)K3 * CO * 4 + (size_t)warp * STEM_WARP_BYTES;
    uint2* rec = reinterpret_cast<uint2*>(mine);                                     // [ECAP] B: (., row << 7 | k)  C: (x, k * 32)
    unsigned long long* ms = reinterpret_cast<unsigned long long*>(mine);            // [8][32] block masks (phase A only)
    int* n3 = reinterpret_cast<int*>(mine + STEM_ECAP * 8);                          // [27][32]
    int4* cs = reinterpret_cast<int4*>(mine + STEM_ECAP * 8 + 27 * 32 * 4);          // [32] coordinates of the warp's rows
    int* roff = reinterpret_cast<int*>(mine + STEM_ECAP * 8 + 27 * 32 * 4 + 32 * 16);    // [33] entry offsets of the rows
    for (int e = tid; e < K3 * CO; e += STEM_NT) w_s[e] = a.weight[e];
    __syncthreads();
    const int o_base = (blockIdx.x * (STEM_NT / 32) + warp) * 32;
    if (o_base >= a.n) return;                                   // whole warp out of range (no CTA barrier below)
    const int o = o_base + lane;
    const bool valid = o < a.n;
    // ---------------------------------------------------------------- A: occupancy of the row's neighbourhood
    unsigned long long occ_lo = 0ull, occ_hi = 0ull;
    int4 c = make_int4(0, 0, 0, 0);
    if (valid) c = reinterpret_cast<const int4*>(a.coords)[o];
    cs[lane] = c;
#pragma unroll
    for (int k3 = 0; k3 < 27; ++k3) n3[k3 * 32 + lane] = -1;
    if (valid) stem_occupancy<KS>(a, c, lane, ms, occ_lo, occ_hi);
    __syncwarp();                                                  // ms (aliased by rec) is dead from here on
    // ---------------------------------------------------------------- B..D in passes of <= STEM_ECAP entries
    const int cnt = __popcll(occ_lo) + __popcll(occ_hi);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    const int half = lane >> 4, l2 = (lane & 15) * 2;              // phase D: half-warp = row, lane = channels l2, l2 + 1
    const float sc0 = a.scale ? __ldg(a.scale + l2) : 1.f, sc1 = a.scale ? __ldg(a.scale + l2 + 1) : 1.f;
    const float sh0 = a.shift ? __ldg(a.shift + l2) : 0.f, sh1 = a.shift ? __ldg(a.shift + l2 + 1) : 0.f;
    bool bad = false;
    int r_begin = 0;
    while (r_begin < 32) {
        const int base = r_begin ? __shfl_sync(0xffffffffu, incl, r_begin - 1) : 0;
        const unsigned fits = __ballot_sync(0xffffffffu, incl - base <= STEM_ECAP);
        const int r_end = 32 - __clz(fits);                        // incl is monotone: rows [r_begin, r_end) fit (r_end > r_begin)
        const int n_ent = __shfl_sync(0xffffffffu, incl, r_end - 1) - base;
        if (lane >= r_begin && lane < r_end) {
            int e = incl - cnt - base;
            roff[lane] = e;
            unsigned long long m = occ_lo;
            while (m) { const int k = __ffsll((long long)m) - 1; m &= m - 1; rec[e++].y = (unsigned)((lane << 7) | k); }
            m = occ_hi;
            while (m) { const int k = 64 + __ffsll((long long)m) - 1; m &= m - 1; rec[e++].y = (unsigned)((lane << 7) | k); }
        }
        if (lane == 0) roff[r_end] = n_ent;
        __syncwarp();
        // C: probes
        for (int e0 = 0; e0 < n_ent; e0 += 128) {
            unsigned long long key[4];
            bool act[4];
            int v[4], kk[4], rr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 32 + lane;
                const unsigned en = e < n_ent ? rec[e].y : 0u;
                kk[u] = en & 127;
                rr[u] = en >> 7;
                const int4 cc = cs[rr[u]];
                const int ix = kk[u] % KS - r, iy = (kk[u] / KS) % KS - r, iz = kk[u] / (KS * KS) - r;
                // with an all-ones input only the 3^3 sub-cube (the neighbour table) needs the neighbour's row
                const bool need_row = a.in != nullptr || (a.nbr3 && ix >= -1 && ix <= 1 && iy >= -1 && iy <= 1 && iz >= -1 && iz <= 1);
                act[u] = e < n_ent && kk[u] != CENTRE && need_row;
                key[u] = pack4(cc.x, cc.y + ix, cc.z + iy, cc.w + iz);
            }
            hash_lookup_row<4>(a.keys, a.vals, a.cap, key, act, v);
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk[u] == CENTRE) v[u] = o_base + rr[u];
                x[u] = 1.0f;
                if (a.in) x[u] = (e0 + u * 32 + lane < n_ent && v[u] >= 0) ? __ldg(a.in + v[u]) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 32 + lane;
                if (e < n_ent) {
                    rec[e] = make_uint2(__float_as_uint(x[u]), (unsigned)(kk[u] * CO));
                    const int ix = kk[u] % KS - r, iy = (kk[u] / KS) % KS - r, iz = kk[u] / (KS * KS) - r;
                    if (ix >= -1 && ix <= 1 && iy >= -1 && iy <= 1 && iz >= -1 && iz <= 1)
                        n3[((ix + 1) + 3 * (iy + 1) + 9 * (iz + 1)) * 32 + rr[u]] = v[u];
                }
            }
        }
        __syncwarp();
        // D: two rows at a time, one per half-warp
        for (int rw0 = r_begin; rw0 < r_end; rw0 += 2) {
            const int rw = rw0 + half;
            const int orow = o_base + rw;
            const bool live = rw < r_end && orow < a.n;
            int e = live ? roff[rw] : 0;
            const int e1 = live ? roff[rw + 1] : 0;
            float acc0 = 0.f, acc1 = 0.f;
            for (; e + 4 <= e1; e += 4) {
                const uint2 q0 = rec[e], q1 = rec[e + 1], q2 = rec[e + 2], q3 = rec[e + 3];
                const float2 w0 = *reinterpret_cast<const float2*>(w_s + q0.y + l2), w1 = *reinterpret_cast<const float2*>(w_s + q1.y + l2);
                const float2 w2 = *reinterpret_cast<const float2*>(w_s + q2.y + l2), w3 = *reinterpret_cast<const float2*>(w_s + q3.y + l2);
                acc0 = __fmaf_rn(__uint_as_float(q0.x), w0.x, acc0); acc1 = __fmaf_rn(__uint_as_float(q0.x), w0.y, acc1);
                acc0 = __fmaf_rn(__uint_as_float(q1.x), w1.x, acc0); acc1 = __fmaf_rn(__uint_as_float(q1.x), w1.y, acc1);
                acc0 = __fmaf_rn(__uint_as_float(q2.x), w2.x, acc0); acc1 = __fmaf_rn(__uint_as_float(q2.x), w2.y, acc1);
                acc0 = __fmaf_rn(__uint_as_float(q3.x), w3.x, acc0); acc1 = __fmaf_rn(__uint_as_float(q3.x), w3.y, acc1);
            }
            for (; e < e1; ++e) {
                const uint2 q0 = rec[e];
                const float2 w0 = *reinterpret_cast<const float2*>(w_s + q0.y + l2);
                acc0 = __fmaf_rn(__uint_as_float(q0.x), w0.x, acc0); acc1 = __fmaf_rn(__uint_as_float(q0.x), w0.y, acc1);
            }
            if (live) {
                float y0 = acc0, y1 = acc1;
                if (a.scale) { y0 = __fmaf_rn(y0, sc0, sh0); y1 = __fmaf_rn(y1, sc1, sh1); }
                else if (a.shift) { y0 += sh0; y1 += sh1; }
                if (a.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                bad |= !(fabsf(y0) < 65504.f) || !(fabsf(y1) < 65504.f);
                if (a.out_xh) {
                    __half h0, g0, h1, g1;
                    xh_split(y0, h0, g0);
                    xh_split(y1, h1, g1);
                    __half2* row = reinterpret_cast<__half2*>(a.out_xh + (size_t)orow * CO * 4);
                    row[lane & 15] = __halves2half2(h0, h1);
                    row[16 + (lane & 15)] = __halves2half2(g0, g1);
                } else {
                    *reinterpret_cast<float2*>(a.out + (size_t)orow * CO + l2) = make_float2(y0, y1);
                }
            }
        }
        __syncwarp();
        r_begin = r_end;
    }
    if (a.out_xh && a.range_status && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.range_status, 1);
    if (a.nbr3 && valid) {
#pragma unroll
        for (int k3 = 0; k3 < 27; ++k3) a.nbr3[(size_t)k3 * a.n + o] = n3[k3 * 32 + lane];
    }
}

// All-ones input (StemArgs.in == NULL, the reference's occupancy-only features): the convolution is a sum of the weight rows
// of the occupied offsets, straight off the occupancy bits - no entry list, no value gather - and only the 27 offsets of the
// 3^3 neighbour table are probed (lane = row, 9 probes in flight).  Rows are accumulated by groups of 8 lanes, four channels
// per lane, in ascending k:  acc = fma(1.0f, w, acc)  ==  acc + w  rounded once, as in the general kernel.
template <int KS>
__global__ void __launch_bounds__(STEM_NT, 3)
stem_ones_kernel(const StemArgs a) {
    constexpr int K3 = KS * KS * KS, r = (KS - 1) / 2, CO = 32;
    extern __shared__ __align__(16) unsigned char stem_smem[];
    float* w_s = reinterpret_cast<float*>(stem_smem);                               // [K3][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* ms = reinterpret_cast<unsigned long long*>(stem_smem + (size_t)K3 * CO * 4) + (size_t)warp * (8 * 32 + 64);
    unsigned long long* occ_s = ms + 8 * 32;                                         // [32][2] occupancy of the warp's rows
    for (int e = tid; e < K3 * CO; e += STEM_NT) w_s[e] = a.weight[e];
    __syncthreads();
    const int o_base = (blockIdx.x * (STEM_NT / 32) + warp) * 32;
    if (o_base >= a.n) return;
    const int o = o_base + lane;
    const bool valid = o < a.n;
    unsigned long long occ_lo = 0ull, occ_hi = 0ull;
    int4 c = make_int4(0, 0, 0, 0);
    if (valid) {
        c = reinterpret_cast<const int4*>(a.coords)[o];
        stem_occupancy<KS>(a, c, lane, ms, occ_lo, occ_hi);
    }
    occ_s[lane * 2] = occ_lo;
    occ_s[lane * 2 + 1] = occ_hi;
    // ---- the 3^3 neighbour table: z-slabs of 9 probes
    if (a.nbr3 && valid) {
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
            unsigned long long key[9];
            bool act[9];
            int v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int ix = j % 3 - 1, iy = j / 3 - 1, iz = s3 - 1;
                const int k = (ix + r) + KS * ((iy + r) + KS * (iz + r));           // compile-time
                const bool occ = k < 64 ? (occ_lo >> (k & 63)) & 1ull : (occ_hi >> ((k - 64) & 63)) & 1ull;
                act[j] = occ && !(ix == 0 && iy == 0 && iz == 0);
                key[j] = pack4(c.x, c.y + ix, c.z + iy, c.w + iz);
            }
            hash_lookup_row<9>(a.keys, a.vals, a.cap, key, act, v);
            if (s3 == 1) v[4] = o;
#pragma unroll
            for (int j = 0; j < 9; ++j) a.nbr3[(size_t)(s3 * 9 + j) * a.n + o] = v[j];
        }
    }
    __syncwarp();
    // ---- convolution: 8 lanes = one row (4 rows of the warp at a time), lane = channels l4 .. l4 + 3; the occupancy is walked
    //      as 32-bit words (little endian halves of occ_lo / occ_hi), ascending k
    const unsigned int* occ32 = reinterpret_cast<const unsigned int*>(occ_s);       // [32][4]
    const int sub = lane >> 3, l4 = (lane & 7) * 4;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = a.scale ? __ldg(a.scale + l4 + j) : 1.f;
        sh[j] = a.shift ? __ldg(a.shift + l4 + j) : 0.f;
    }
    bool bad = false;
    for (int rw0 = 0; rw0 < 32; rw0 += 4) {
        const int rw = rw0 + sub, orow = o_base + rw;
        if (orow >= a.n) continue;
        float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int wd = 0; wd < (K3 + 31) / 32; ++wd) {
            unsigned int m = occ32[rw * 4 + wd];
            while (m) {
                const int k = wd * 32 + __ffs((int)m) - 1;
                m &= m - 1;
                const float4 w = *reinterpret_cast<const float4*>(w_s + k * CO + l4);
                y[0] = __fadd_rn(y[0], w.x);
                y[1] = __fadd_rn(y[1], w.y);
                y[2] = __fadd_rn(y[2], w.z);
                y[3] = __fadd_rn(y[3], w.w);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (a.scale) y[j] = __fmaf_rn(y[j], sc[j], sh[j]);
            else if (a.shift) y[j] += sh[j];
            if (a.relu) y[j] = fmaxf(y[j], 0.f);
            bad |= !(fabsf(y[j]) < 65504.f);
        }
        if (a.out_xh) {
            __half h[4], g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xh_split(y[j], h[j], g[j]);
            uint2 hv, gv;
            *reinterpret_cast<__half2*>(&hv.x) = __halves2half2(h[0], h[1]);
            *reinterpret_cast<__half2*>(&hv.y) = __halves2half2(h[2], h[3]);
            *reinterpret_cast<__half2*>(&gv.x) = __halves2half2(g[0], g[1]);
            *reinterpret_cast<__half2*>(&gv.y) = __halves2half2(g[2], g[3]);
            uint8_t* row = a.out_xh + (size_t)orow * CO * 4;
            *reinterpret_cast<uint2*>(row + l4 * 2) = hv;
            *reinterpret_cast<uint2*>(row + 64 + l4 * 2) = gv;
        } else {
            *reinterpret_cast<float4*>(a.out + (size_t)orow * CO + l4) = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
    if (a.out_xh && a.range_status && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.range_status, 1);
}

}  // namespace

extern "C" size_t eyoc_tile_order_workspace_bytes(int64_t n_out) {
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)n_out, 0, 43);
    return 2 * eyoc_align((size_t)n_out * 8) + eyoc_align((size_t)n_out * 4) + eyoc_align(temp) + 1024;
}

extern "C" int eyoc_tile_order(const int32_t* nbr, int K, int64_t n_out, const int32_t* out_coords, int group_clouds, int max_batch,
                               int32_t* row_perm, int32_t* nbr_tiled, uint32_t* tile_masks, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
    EYOC_CHECK_ARG(nbr && out_coords && row_perm && nbr_tiled, "eyoc_tile_order: null argument");
    EYOC_CHECK_ARG(K >= 1 && K <= 27 && group_clouds >= 1 && max_batch >= 0 && max_batch <= 65535, "eyoc_tile_order: bad K / group / batch");
    EYOC_CHECK_ARG(n_out >= 0 && n_out < (1ll << 31), "eyoc_tile_order: bad n_out");
    if (n_out == 0) return EYOC_OK;
    if (workspace == nullptr || workspace_bytes < eyoc_tile_order_workspace_bytes(n_out)) {
        eyoc_set_error("eyoc_tile_order: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    WsCarver c(workspace, workspace_bytes);
    unsigned long long* keys = c.take<unsigned long long>(n_out);
    unsigned long long* keys2 = c.take<unsigned long long>(n_out);
    int* iota = c.take<int>(n_out);
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, keys, keys2, iota, row_perm, (int)n_out, 0, 43);
    void* tmp = c.take<char>(temp);
    const unsigned g = (unsigned)((n_out + 255) / 256);
    unsigned int* hist = c.take<unsigned int>(32);
    int* pos = c.take<int>(32);
    EYOC_CUDA(cudaMemsetAsync(hist, 0, 32 * sizeof(unsigned int), stream));
    tile_key_kernel<<<g, 256, 0, stream>>>(nbr, K, (int)n_out, out_coords, group_clouds, keys, iota, hist);
    EYOC_LAUNCH_CHECK();
    tile_bit_order_kernel<<<1, 32, 0, stream>>>(hist, K, pos);
    EYOC_LAUNCH_CHECK();
    tile_key_remap_kernel<<<g, 256, 0, stream>>>(keys, (int)n_out, K, pos);
    EYOC_LAUNCH_CHECK();
    int gbits = 0;
    while ((max_batch / group_clouds) >> gbits) ++gbits;
    EYOC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, temp, keys, keys2, iota, row_perm, (int)n_out, 0, 27 + gbits, stream));
    g_eyoc_launches += 4;
    if (tile_masks) EYOC_CUDA(cudaMemsetAsync(tile_masks, 0, (size_t)g * sizeof(uint32_t), stream));
    permute_columns_kernel<<<dim3(g, K), 256, 0, stream>>>(nbr, (int)n_out, row_perm, nbr_tiled, tile_masks);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals, int64_t capacity,
                               int32_t* status, cudaStream_t stream) {
    EYOC_CHECK_ARG(coords && table_keys && table_vals && status, "eyoc_hash_build: null argument");
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31), "eyoc_hash_build: bad n");
    EYOC_CHECK_ARG(capacity >= 2 * n && capacity >= 2 && (capacity & (capacity - 1)) == 0, "eyoc_hash_build: capacity must be a power of two >= 2n");
    hash_clear_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>((unsigned long long*)table_keys, table_vals, capacity);
    EYOC_LAUNCH_CHECK();
    if (n == 0) return EYOC_OK;
    const unsigned g = (unsigned)((n + 255) / 256);
    hash_build_kernel<<<g, 256, 0, stream>>>(coords, (int)n, (unsigned long long*)table_keys, table_vals, capacity, status);
    EYOC_LAUNCH_CHECK();
    hash_check_unique_kernel<<<g, 256, 0, stream>>>(coords, (int)n, (const unsigned long long*)table_keys, table_vals, capacity, status);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" size_t eyoc_downsample_workspace_bytes(int64_t n) {
    const size_t nb = (size_t)((n + SCAN_B - 1) / SCAN_B);
    return eyoc_align((size_t)n * 4) * 2 + eyoc_align((nb + 1) * 4);
}

extern "C" int eyoc_coords_downsample(const int32_t* coords, int64_t n, int ts_out, uint64_t* table_keys, int32_t* table_vals,
                                      int64_t capacity, int32_t* coords_out, int32_t* n_out, void* workspace, size_t workspace_bytes,
                                      cudaStream_t stream) {
    EYOC_CHECK_ARG(coords && table_keys && table_vals && coords_out && n_out, "eyoc_coords_downsample: null argument");
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31) && ts_out >= 2, "eyoc_coords_downsample: bad n / stride");
    EYOC_CHECK_ARG(capacity >= 2 * n && capacity >= 2 && (capacity & (capacity - 1)) == 0, "eyoc_coords_downsample: capacity must be a power of two >= 2n");
    if (workspace == nullptr || workspace_bytes < eyoc_downsample_workspace_bytes(n)) {
        eyoc_set_error("eyoc_coords_downsample: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    hash_clear_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>((unsigned long long*)table_keys, table_vals, capacity);
    EYOC_LAUNCH_CHECK();
    if (n == 0) { EYOC_CUDA(cudaMemsetAsync(n_out, 0, 4, stream)); return EYOC_OK; }
    WsCarver c(workspace, workspace_bytes);
    int* flag = c.take<int>(n);
    int* pos = c.take<int>(n);
    const int nb = (int)((n + SCAN_B - 1) / SCAN_B);
    int* sums = c.take<int>(nb + 1);
    const unsigned g = (unsigned)((n + 255) / 256);
    down_insert_kernel<<<g, 256, 0, stream>>>(coords, (int)n, ts_out, (unsigned long long*)table_keys, table_vals, capacity);
    EYOC_LAUNCH_CHECK();
    down_flag_kernel<<<g, 256, 0, stream>>>(coords, (int)n, ts_out, (const unsigned long long*)table_keys, table_vals, capacity, flag);
    EYOC_LAUNCH_CHECK();
    scan_block_kernel<<<nb, SCAN_B, 0, stream>>>(flag, (int)n, pos, sums);
    EYOC_LAUNCH_CHECK();
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, nb, n_out);
    EYOC_LAUNCH_CHECK();
    down_emit_kernel<<<g, 256, 0, stream>>>(coords, (int)n, ts_out, flag, pos, sums, (unsigned long long*)table_keys, table_vals, capacity,
                                            coords_out, nullptr);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" size_t eyoc_voxelize_workspace_bytes(int64_t n) {
    return eyoc_align((size_t)n * 16) + eyoc_downsample_workspace_bytes(n);
}

extern "C" int eyoc_voxelize(const float* xyz, const int32_t* cloud, int64_t n, float voxel_size, uint64_t* table_keys,
                             int32_t* table_vals, int64_t capacity, int32_t* coords_out, int32_t* sel_out, int32_t* n_out,
                             int32_t* status, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(xyz && table_keys && table_vals && coords_out && sel_out && n_out && status, "eyoc_voxelize: null argument");
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31) && voxel_size > 0.f, "eyoc_voxelize: bad n / voxel size");
    EYOC_CHECK_ARG(capacity >= 2 * n && capacity >= 2 && (capacity & (capacity - 1)) == 0, "eyoc_voxelize: capacity must be a power of two >= 2n");
    if (workspace == nullptr || workspace_bytes < eyoc_voxelize_workspace_bytes(n)) {
        eyoc_set_error("eyoc_voxelize: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    hash_clear_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>((unsigned long long*)table_keys, table_vals, capacity);
    EYOC_LAUNCH_CHECK();
    if (n == 0) { EYOC_CUDA(cudaMemsetAsync(n_out, 0, 4, stream)); return EYOC_OK; }
    WsCarver c(workspace, workspace_bytes);
    int* q = c.take<int>(4 * n);
    int* flag = c.take<int>(n);
    int* pos = c.take<int>(n);
    const int nb = (int)((n + SCAN_B - 1) / SCAN_B);
    int* sums = c.take<int>(nb + 1);
    const unsigned g = (unsigned)((n + 255) / 256);
    quantize_kernel<<<g, 256, 0, stream>>>(xyz, cloud, (int)n, voxel_size, q, status);
    EYOC_LAUNCH_CHECK();
    down_insert_kernel<<<g, 256, 0, stream>>>(q, (int)n, 1, (unsigned long long*)table_keys, table_vals, capacity);
    EYOC_LAUNCH_CHECK();
    down_flag_kernel<<<g, 256, 0, stream>>>(q, (int)n, 1, (const unsigned long long*)table_keys, table_vals, capacity, flag);
    EYOC_LAUNCH_CHECK();
    scan_block_kernel<<<nb, SCAN_B, 0, stream>>>(flag, (int)n, pos, sums);
    EYOC_LAUNCH_CHECK();
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, nb, n_out);
    EYOC_LAUNCH_CHECK();
    down_emit_kernel<<<g, 256, 0, stream>>>(q, (int)n, 1, flag, pos, sums, (unsigned long long*)table_keys, table_vals, capacity, coords_out,
                                            sel_out);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_kernel_map(const int32_t* out_coords, int64_t n_out, const uint64_t* in_table_keys, const int32_t* in_table_vals,
                               int64_t capacity, int ksize, int step, int32_t* nbr, cudaStream_t stream) {
    EYOC_CHECK_ARG(out_coords && in_table_keys && in_table_vals && nbr, "eyoc_kernel_map: null argument");
    EYOC_CHECK_ARG(ksize >= 1 && (ksize & 1) && ksize <= 7, "eyoc_kernel_map: kernel size must be odd and <= 7 (got %d)", ksize);
    EYOC_CHECK_ARG(n_out >= 0 && n_out < (1ll << 31), "eyoc_kernel_map: bad n_out");
    if (n_out == 0) return EYOC_OK;
    dim3 grid((unsigned)((n_out + 255) / 256), ksize * ksize);
    const unsigned long long* tk = (const unsigned long long*)in_table_keys;
    switch (ksize) {
        case 1: kernel_map_kernel<1><<<grid, 256, 0, stream>>>(out_coords, (int)n_out, tk, in_table_vals, capacity, step, nbr); break;
        case 3: kernel_map_kernel<3><<<grid, 256, 0, stream>>>(out_coords, (int)n_out, tk, in_table_vals, capacity, step, nbr); break;
        case 5: kernel_map_kernel<5><<<grid, 256, 0, stream>>>(out_coords, (int)n_out, tk, in_table_vals, capacity, step, nbr); break;
        default: kernel_map_kernel<7><<<grid, 256, 0, stream>>>(out_coords, (int)n_out, tk, in_table_vals, capacity, step, nbr); break;
    }
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_kernel_map_self(const int32_t* coords, int64_t n, const uint64_t* table_keys, const int32_t* table_vals,
                                    int64_t capacity, int ksize, int step, int32_t* nbr, cudaStream_t stream) {
    EYOC_CHECK_ARG(coords && table_keys && table_vals && nbr, "eyoc_kernel_map_self: null argument");
    EYOC_CHECK_ARG(ksize >= 1 && (ksize & 1) && ksize <= 7, "eyoc_kernel_map_self: kernel size must be odd and <= 7 (got %d)", ksize);
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31), "eyoc_kernel_map_self: bad n");
    if (n == 0) return EYOC_OK;
    const int K3 = ksize * ksize * ksize;
    EYOC_CUDA(cudaMemsetAsync(nbr, 0xff, (size_t)K3 * n * sizeof(int32_t), stream));
    dim3 grid((unsigned)((n + 255) / 256), (K3 / 2) / ksize + 1);
    const unsigned long long* tk = (const unsigned long long*)table_keys;
    switch (ksize) {
        case 1: kernel_map_sym_kernel<1><<<grid, 256, 0, stream>>>(coords, (int)n, tk, table_vals, capacity, step, nbr); break;
        case 3: kernel_map_sym_kernel<3><<<grid, 256, 0, stream>>>(coords, (int)n, tk, table_vals, capacity, step, nbr); break;
        case 5: kernel_map_sym_kernel<5><<<grid, 256, 0, stream>>>(coords, (int)n, tk, table_vals, capacity, step, nbr); break;
        default: kernel_map_sym_kernel<7><<<grid, 256, 0, stream>>>(coords, (int)n, tk, table_vals, capacity, step, nbr); break;
    }
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_kernel_map_transpose(const int32_t* nbr_down, int64_t n_coarse, int64_t n_fine, int K, int32_t* nbr_up,
                                         cudaStream_t stream) {
    EYOC_CHECK_ARG(nbr_down && nbr_up && K >= 1, "eyoc_kernel_map_transpose: bad argument");
    EYOC_CHECK_ARG(n_coarse >= 0 && n_fine >= 0 && n_coarse < (1ll << 31) && n_fine < (1ll << 31), "eyoc_kernel_map_transpose: bad sizes");
    if (n_fine == 0) return EYOC_OK;
    EYOC_CUDA(cudaMemsetAsync(nbr_up, 0xff, (size_t)K * n_fine * sizeof(int32_t), stream));
    if (n_coarse == 0) return EYOC_OK;
    kernel_map_transpose_kernel<<<dim3((unsigned)((n_coarse + 255) / 256), K), 256, 0, stream>>>(nbr_down, (int)n_coarse, (int)n_fine, nbr_up);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_parity_class(const int32_t* coords, int64_t n, int ts, int32_t* cls, cudaStream_t stream) {
    EYOC_CHECK_ARG(coords && cls && ts >= 1, "eyoc_parity_class: bad argument");
    if (n == 0) return EYOC_OK;
    parity_class_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(coords, (int)n, ts, cls);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" size_t eyoc_stem_conv_workspace_bytes(int64_t capacity) { return (size_t)capacity * 16 + 256; }

extern "C" int eyoc_stem_conv(const int32_t* coords, int64_t n, const uint64_t* table_keys, const int32_t* table_vals, int64_t capacity,
                              int ksize, const float* in, const float* weight, const float* scale, const float* shift, int relu,
                              void* out, int out_packed, int32_t* range_status, int32_t* nbr3, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
    EYOC_CHECK_ARG(coords && table_keys && table_vals && weight && out, "eyoc_stem_conv: null argument");
    EYOC_CHECK_ARG(ksize == 3 || ksize == 5, "eyoc_stem_conv: kernel size must be 3 or 5 (got %d)", ksize);
    EYOC_CHECK_ARG(n >= 0 && n < (1ll << 31), "eyoc_stem_conv: bad n");
    EYOC_CHECK_ARG(capacity >= 2 * n && capacity >= 2 && (capacity & (capacity - 1)) == 0, "eyoc_stem_conv: capacity must be a power of two >= 2n");
    if (n == 0) return EYOC_OK;
    if (workspace == nullptr || workspace_bytes < eyoc_stem_conv_workspace_bytes(capacity)) {
        eyoc_set_error("eyoc_stem_conv: workspace too small");
        return EYOC_ERR_WORKSPACE;
    }
    ulonglong2* blocks = (ulonglong2*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    block_clear_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>(blocks, capacity);
    EYOC_LAUNCH_CHECK();
    block_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(coords, (int)n, blocks, capacity);
    EYOC_LAUNCH_CHECK();
    StemArgs a{coords, (int)n, (const unsigned long long*)table_keys, table_vals, capacity, blocks, capacity, in, weight, scale, shift,
               relu, out_packed ? nullptr : (float*)out, out_packed ? (uint8_t*)out : nullptr, range_status, nbr3};
    const unsigned grid = (unsigned)((n + STEM_NT - 1) / STEM_NT);
    if (in == nullptr) {
        const size_t smem = (size_t)ksize * ksize * ksize * 32 * sizeof(float) + (size_t)(STEM_NT / 32) * (8 * 32 + 64) * 8;
        if (ksize == 3) {
            EYOC_CUDA(cudaFuncSetAttribute(stem_ones_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            stem_ones_kernel<3><<<grid, STEM_NT, smem, stream>>>(a);
        } else {
            EYOC_CUDA(cudaFuncSetAttribute(stem_ones_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            stem_ones_kernel<5><<<grid, STEM_NT, smem, stream>>>(a);
        }
    } else {
        const size_t smem = (size_t)ksize * ksize * ksize * 32 * sizeof(float) + (size_t)(STEM_NT / 32) * STEM_WARP_BYTES;
        if (ksize == 3) {
            EYOC_CUDA(cudaFuncSetAttribute(stem_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            stem_conv_kernel<3><<<grid, STEM_NT, smem, stream>>>(a);
        } else {
            EYOC_CUDA(cudaFuncSetAttribute(stem_conv_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            stem_conv_kernel<5><<<grid, STEM_NT, smem, stream>>>(a);
        }
    }
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
