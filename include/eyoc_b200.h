/* eyoc_b200 — C ABI of the B200-native EYOC registration-inference hot path.
 *
 * The reference (liuQuan98/EYOC) has no FFI on this path: everything is Python calling
 * MinkowskiEngine's pybind backend and torch ATen.  This header is therefore the boundary a
 * maintainer binds INSTEAD of those two libraries (ctypes stub: eyoc_b200/_C.py; see
 * INTEGRATION.md).  Each entry point cites the reference interface it replaces
 * (paths relative to the reference tree).
 *
 * Conventions
 *   - return 0 = OK; -1 bad argument; -2 CUDA error; -3 workspace too small; -4 degenerate input.
 *     eyoc_last_error() returns a thread-local message for the last non-zero return.
 *   - every tensor pointer is CALLER-OWNED DEVICE memory, contiguous row-major; the library
 *     never allocates, frees or retains device memory: scratch comes from the caller through
 *     (workspace, workspace_bytes) sized by the matching *_workspace_bytes() function.
 *   - all work is stream-ordered on `stream`; no call synchronises the device or the host.
 *   - no torch / C++ types cross the boundary.
 */
#ifndef EYOC_B200_H
#define EYOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* eyoc_stream_t; /* == cudaStream_t */

int eyoc_version(void);
const char* eyoc_last_error(void);
int eyoc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
unsigned long long eyoc_launch_count(void);

/* Row gather out[i, :] = src[idx[i], :] (src [n, c] fp32, idx [m] int64 -> out [m, c]): the index compositions of
 * scripts/test_kitti.py:36-42,69-73 (find_corr, random_sample) and scripts/SC2_PCR/SC2_PCR.py:290-305 (match_pair) for a
 * whole block of pairs in one launch. */
int eyoc_gather_rows(const float* src, const int64_t* idx, int64_t m, int c, float* out, eyoc_stream_t stream);

/* ---------------------------------------------------------------- nearest neighbour matching
 * Replaces lib/eval.py:18-48 find_nn_gpu (+ lib/metrics.py:26-27 pdist 'SquareL2')  [form 0]
 * and the matching core of scripts/SC2_PCR/SC2_PCR.py:296-298 Matcher.match_pair       [form 1].
 * q [batch, nq, dim], r [batch, nr, dim] fp32 -> idx [batch, nq] int64 (ties -> lowest index,
 * NaN distance wins like torch.argmin), dist [batch, nq] fp32 (either output may be NULL). */
size_t eyoc_knn1_workspace_bytes(int batch, int64_t nq);
int eyoc_knn1(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
              void* workspace, size_t workspace_bytes, int64_t* idx, float* dist, eyoc_stream_t stream);
/* Nearest neighbour ignoring one reference column per query (exclude [batch, nq] int64): second pass of the K = 2 search
 * of lib/trainer.py:1064-1065 (pytorch3d.ops.knn_points: squared L2, ascending, ties by lowest index).  A query whose
 * reference set holds nothing but the excluded column gets idx -1 / dist +inf.  Workspace: eyoc_knn1_workspace_bytes. */
int eyoc_knn1_excluding(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                        const int64_t* exclude, void* workspace, size_t workspace_bytes, int64_t* idx, float* dist,
                        eyoc_stream_t stream);
/* The same operator for 32-channel descriptors with a tensor-core pre-filter: fp16 scores on tcgen05 with a rigorous
 * error bound select the few columns per row that can be the exact winner; those are re-scored in eyoc_knn1's fp32-FMA
 * order, so idx / dist are bit-identical to eyoc_knn1's.  Batches holding non-finite or fp16-overflowing values are
 * routed to the fp32-FMA kernel on the device. */
int eyoc_knn1_tc_supported(int dim);
size_t eyoc_knn1_tc_workspace_bytes(int batch, int64_t nq, int64_t nr);
int eyoc_knn1_tc(const float* q, const float* r, int batch, int64_t nq, int64_t nr, int dim, int form,
                 void* workspace, size_t workspace_bytes, int64_t* idx, float* dist, eyoc_stream_t stream);

/* ---------------------------------------------------------------- SC2-PCR estimator
 * Replaces scripts/SC2_PCR/SC2_PCR.py:307-384 Matcher.SC2_PCR (first/second-order spatial
 * compatibility, :170-196 power iteration, :33-59 pick_seeds, :61-168 cal_seed_trans,
 * :238-278 post_refinement), the label pass of :409-411 Matcher.estimator, and
 * scripts/SC2_PCR/common.py:7-45 rigid_transform_3d (weighted Kabsch; 3x3 SVD on device).
 * Batched over `batch` independent pairs (the reference asserts batch == 1, SC2_PCR.py:44,249).
 * Thresholds are passed already rounded to fp32 exactly as torch rounds Python scalars. */
typedef struct {
    float inlier_threshold;  /* config_KITTI.json:6   final labels + per-seed fitness            */
    float d_thre;            /* :7  first-order compatibility   cross < d_thre                    */
    float d_thre_half;       /* fp32(d_thre / 2)  tight compatibility, SC2_PCR.py:357             */
    float d_thre_sq;         /* fp32(d_thre ** 2) soft measure denominator, SC2_PCR.py:341        */
    float nms_radius;        /* :14 */
    float refine_threshold;  /* 1.2 unless inlier_threshold == 0.10, SC2_PCR.py:254-257           */
    int num_iterations;      /* :2  power-iteration cap                                           */
    int k1, k2;              /* :4-5 (the k1 > n fallback to 4/4, SC2_PCR.py:76-78, is applied inside) */
    int refine_iterations;   /* 20, SC2_PCR.py:374 */
} eyoc_sc2_cfg;

/* Optional stage inputs: when non-NULL the corresponding stage is skipped and these device arrays are
 * used instead (stage-parity tests feed the oracle's upstream tensors through them). */
typedef struct {
    const float* confidence;      /* [batch, n]      skip the leading-eigenvector stage        */
    const int32_t* seeds;         /* [batch, S]      skip NMS + ranking                        */
    const float* initial_trans;   /* [batch, 4, 4]   skip the whole seed stage                 */
    const float* sc2_dense;       /* [batch, S, n]   caller's second-order measure (integer-valued, SC2_PCR.py:363) instead
                                     of the bit-matrix one; needs `seeds` (drop-in Matcher.cal_seed_trans)             */
} eyoc_sc2_hooks;

/* Byte offsets of the intermediate buffers inside the workspace (for tests and diagnostics). */
typedef struct {
    size_t points, hard_bits, tight_bits, vbuf, u, confidence, scores, seeds, topk1, topk2, local_v,
        seed_weights, seed_trans, counters, global_iters, local_notclose, best_seed, refine_counts, total,
        csr_rowptr, csr_cols, csr_vals, csr_capacity, sort_keys, sort_idx, sort_offsets, sort_temp, sort_temp_bytes,
        near_bits, status, big;
    int words_per_row, k1, k2, num_seeds;
} eyoc_sc2_layout;

int eyoc_sc2pcr_layout(int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg, eyoc_sc2_layout* out);
size_t eyoc_sc2pcr_workspace_bytes(int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg);
/* src, tgt [batch, n, 3] fp32 putative correspondences  ->
 * trans [batch, 4, 4], fitness [batch, num_seeds] (may be NULL), labels [batch, n] fp32 0/1 (may be NULL). */
int eyoc_sc2pcr(const float* src, const float* tgt, int batch, int n, int num_seeds, const eyoc_sc2_cfg* cfg,
                const eyoc_sc2_hooks* hooks, void* workspace, size_t workspace_bytes, float* trans, float* fitness,
                float* labels, eyoc_stream_t stream);

/* Stage entry points on DENSE caller tensors (the reference's public Matcher stage methods take dense matrices; the fused
 * estimator above never forms them).
 * scripts/SC2_PCR/SC2_PCR.py:33-59 pick_seeds: dists [batch, n, n], scores [batch, n] -> seeds [batch, max_num] int64
 * (NMS: i survives iff for all j score_i >= score_j or dists_ij >= R; then descending score, ties by lowest index). */
size_t eyoc_pick_seeds_workspace_bytes(int batch, int n);
int eyoc_pick_seeds_dense(const float* dists, const float* scores, int batch, int n, float R, int max_num, int64_t* seeds,
                          void* workspace, size_t workspace_bytes, eyoc_stream_t stream);
/* scripts/SC2_PCR/SC2_PCR.py:179-190 cal_leading_eigenvector(method='power'): M [batch, n, n] -> v [batch, n];
 * v <- M v / (||M v|| + 1e-6) from ones, at most num_iterations times, stopping when torch.allclose(v, v_prev) holds over
 * the WHOLE batch (one allclose call in the reference).  iters_out (device int, may be NULL) = iterations run. */
size_t eyoc_power_iteration_workspace_bytes(int batch, int n, int num_iterations);
int eyoc_power_iteration_dense(const float* M, int batch, int n, int num_iterations, float* v_out, int* iters_out,
                               void* workspace, size_t workspace_bytes, eyoc_stream_t stream);

/* scripts/SC2_PCR/common.py:7-45 rigid_transform_3d: A, B [batch, n, 3], w [batch, n] (NULL = ones;
 * negative weights are zeroed IN PLACE like common.py:20) -> T [batch, 4, 4]. */
int eyoc_kabsch_batched(const float* A, const float* B, float* w, int batch, int n, float weight_threshold,
                        float* T, eyoc_stream_t stream);

/* ---------------------------------------------------------------- coordinate maps / kernel maps
 * Replace MinkowskiEngine's coordinate manager under ME.SparseTensor(F, coordinates=C)
 * (scripts/test_kitti.py:143-147) and under every ME.MinkowskiConvolution[Transpose] of
 * model/resunet.py:31-140.  coords are [n, 4] int32 (batch, x, y, z), 16-bit range per field.
 * Hash table: open addressing, capacity a power of two >= 2n; keys[capacity] u64, vals[capacity] i32.
 * status (device int32[2], caller-zeroed): status[0] bit 0 = coordinate out of range, bit 1 = duplicate coordinate;
 * status[1] = largest batch index seen. */
int eyoc_hash_build(const int32_t* coords, int64_t n, uint64_t* table_keys, int32_t* table_vals, int64_t capacity,
                    int32_t* status, eyoc_stream_t stream);
/* Stride-2 coordinate set unique(floor(c / ts_out) * ts_out) in first-occurrence order (+ its hash table).
 * coords_out has room for n rows; *n_out (device) receives the row count. */
size_t eyoc_downsample_workspace_bytes(int64_t n);
int eyoc_coords_downsample(const int32_t* coords, int64_t n, int ts_out, uint64_t* table_keys, int32_t* table_vals,
                           int64_t capacity, int32_t* coords_out, int32_t* n_out, void* workspace, size_t workspace_bytes,
                           eyoc_stream_t stream);
/* Voxelisation + collate on the device (lib/data_loaders.py:936-979, :31-85): q = floor(xyz / voxel_size) in fp32,
 * first occurrence of every (cloud, q) kept, in ascending point order - what ME.utils.sparse_quantize(return_index=True)
 * followed by floor(xyz[sel] / voxel).int() gives.  xyz [n, 3]; cloud [n] = batch index per point or NULL (all 0);
 * coords_out [n, 4] / sel_out [n]: the first *n_out rows are valid; the hash table is left pointing at the new rows
 * (it IS the stride-1 table eyoc_hash_build would make).  status as in eyoc_hash_build (bit 0 = out of range). */
size_t eyoc_voxelize_workspace_bytes(int64_t n);
int eyoc_voxelize(const float* xyz, const int32_t* cloud, int64_t n, float voxel_size, uint64_t* table_keys,
                  int32_t* table_vals, int64_t capacity, int32_t* coords_out, int32_t* sel_out, int32_t* n_out,
                  int32_t* status, void* workspace, size_t workspace_bytes, eyoc_stream_t stream);
/* nbr[k, o] = row in the input map of (c_o + off_k * step), else -1;  k = ix + K*(iy + K*iz), off = (ix,iy,iz) - (K-1)/2.
 * step = +tensor_stride_in for a forward convolution, -tensor_stride_out for a transposed one. */
int eyoc_kernel_map(const int32_t* out_coords, int64_t n_out, const uint64_t* in_table_keys, const int32_t* in_table_vals,
                    int64_t capacity, int ksize, int step, int32_t* nbr, eyoc_stream_t stream);
/* The same table for a stride-1 map of a coordinate set onto itself (coords = the set that built the table, step = its
 * tensor stride): mirrored offsets are filled from one probe (i = nbr[k, o] <=> o = nbr[K^3-1-k, i]). */
int eyoc_kernel_map_self(const int32_t* coords, int64_t n, const uint64_t* table_keys, const int32_t* table_vals,
                         int64_t capacity, int ksize, int step, int32_t* nbr, eyoc_stream_t stream);
/* Transposed-convolution table from the forward strided one between the same levels:
 * nbr_up[k, f] = c  <=>  nbr_down[k, c] = f  (nbr_down [K, n_coarse] -> nbr_up [K, n_fine], -1 elsewhere). */
int eyoc_kernel_map_transpose(const int32_t* nbr_down, int64_t n_coarse, int64_t n_fine, int K, int32_t* nbr_up,
                              eyoc_stream_t stream);
/* Tile order for the tensor-core convolution: row_perm = output rows stably sorted by (cloud / group_clouds, bit mask
 * of the kernel offsets that have a neighbour - the offsets ordered by how many rows have them, the rarest in the most
 * significant bit, ties by offset index), nbr_tiled[k, i] = nbr[k, row_perm[i]].  Rows with the same neighbour
 * pattern share 128-row tiles (dense or skipped (tile, offset) items) while each group of clouds stays contiguous
 * (L2 working set of the gather).  K <= 27; max_batch = largest batch index (sizes the sort key).  tile_masks (optional,
 * [ceil(n_out / 256)]) receives what eyoc_tile_masks computes from nbr_tiled, out of the same pass. */
size_t eyoc_tile_order_workspace_bytes(int64_t n_out);
int eyoc_tile_order(const int32_t* nbr, int K, int64_t n_out, const int32_t* out_coords, int group_clouds, int max_batch,
                    int32_t* row_perm, int32_t* nbr_tiled, uint32_t* tile_masks, void* workspace, size_t workspace_bytes,
                    eyoc_stream_t stream);
/* cls[i] = parity class (3 bits) of coords[i] / ts: groups the rows of a transposed stride-2 convolution by
 * their set of admissible kernel offsets. */
int eyoc_parity_class(const int32_t* coords, int64_t n, int ts, int32_t* cls, eyoc_stream_t stream);

/* ---------------------------------------------------------------- sparse convolution + fused epilogue
 * Replaces ME.MinkowskiConvolution / MinkowskiConvolutionTranspose forward (model/resunet.py:31-140) fused with
 * what follows it in model/resunet.py:142-193 and model/residual_block.py:37-53:
 *   out[o] = l2norm?( relu?( (sum_k [in0 | in1][nbr[k, o]] @ W[k]) * scale + shift + residual[o] ) )
 * in0 [n_in, c0], in1 [n_in, c1] or NULL (fused ME.cat), weight [K, c0 + c1, cout], scale/shift [cout] or NULL
 * (folded eval BatchNorm; shift alone = bias), residual [n_out, cout] or NULL, nbr NULL = identity (K == 1),
 * row_perm [n_out] or NULL = order in which output rows are tiled (results do not depend on it). */
int eyoc_sparse_conv(const float* in0, int c0, const float* in1, int c1, const int32_t* nbr, int K, int64_t n_out,
                     const int32_t* row_perm, const float* weight, const float* scale, const float* shift,
                     const float* residual, int relu, int l2norm, float* out, int cout, eyoc_stream_t stream);

/* Tensor-core data path of the same operator: tcgen05.mma kind::tf32, issued transposed (weights = M side, 256 gathered
 * rows = N side), tf32 hi/lo split of both operands for fp32-level accuracy, accumulators in TMEM.
 * Weights are first split and laid out once as shared-memory images (one TMA bulk copy per slab):
 * weight [K, cin, cout] -> wt_img, eyoc_conv_weight_image_floats(K, cin, cout) floats.
 * Supported when eyoc_sparse_conv_tc_supported() returns 1 (cin, c0 multiples of 32; cout in {32,64,128,256}; K <= 27;
 * l2norm only for cout <= 128).
 * nbr_tiled != 0: nbr's columns are already in tile order (nbr_tiled[k, i] = nbr[k, row_perm[i]]), so the table is
 * read coalesced; row_perm then only says where each tile row is written. */
size_t eyoc_conv_weight_image_floats(int K, int cin, int cout);
/* Measurement aid (tools/conv_ablate.py): switch off parts of the tensor-core kernel (bit 0 MMAs, bit 1 gather +
 * x_lo pass, bit 2 weight-slab copies) to time the others alone.  0 = normal operation; results are invalid otherwise. */
int eyoc_debug_conv_ablate(int flags);
/* bit 3 of the flags: the first 1024 CTAs record clock64 at {start, work list built, main loop done, accumulators
 * complete, epilogue done} and their item count; this copies the [1024][6] table to the host. */
int eyoc_debug_conv_times(long long* host_out_1024x6);
int eyoc_conv_split_weights(const float* weight, int K, int cin, int cout, float* wt_img, eyoc_stream_t stream);
int eyoc_sparse_conv_tc_supported(int c0, int c1, int cout, int K, int l2norm);
int eyoc_sparse_conv_tc(const float* in0, int c0, const float* in1, int c1, const int32_t* nbr, int K, int64_t n_out,
                        const int32_t* row_perm, int nbr_tiled, const float* wt_img, const float* scale,
                        const float* shift, const float* residual, int relu, int l2norm, float* out, int cout,
                        eyoc_stream_t stream);

/* ME.MinkowskiInstanceNorm as the reference's BasicBlockIN uses it (model/common.py:7-8, model/residual_block.py:60-61; the
 * ResUNetIN2* variants of model/resunet.py:229-251): per cloud (coords[:, 0], 0 .. num_clouds - 1) and channel,
 * y = (x - mean) / sqrt(var + eps) * weight + bias with the biased variance of the cloud's rows; `residual` (optional) is added
 * and ReLU applied afterwards (model/residual_block.py:47-51).  x / residual / out: fp32 [n, c] rows or split-half rows
 * (*_packed); c in {32, 64, 128, 256}; weight / bias [c] or NULL.  Three bandwidth-bound passes, fp64 statistics. */
size_t eyoc_instance_norm_workspace_bytes(int num_clouds, int c);
int eyoc_instance_norm(const void* x, int x_packed, const int32_t* coords, int64_t n, int c, int num_clouds, const float* weight,
                       const float* bias, float eps, const void* residual, int residual_packed, int relu, void* out,
                       int out_packed, int32_t* range_status, void* workspace, size_t workspace_bytes, eyoc_stream_t stream);

/* The network's first convolution (reference model/resunet.py:31-37 `conv1`: ME.MinkowskiConvolution(1, 32, kernel_size=5,
 * stride=1) + model/resunet.py:38 `norm1` + ReLU of :147-150) fused with its own neighbour search on the stride-1 coordinate
 * set: `in` [n] is the single input channel (NULL = every value is 1.0, the occupancy-only input the reference feeds:
 * lib/data_loaders.py / scripts/test_kitti.py build `feats = ones(N, 1)`; only the 27 offsets of nbr3 are then probed), `weight` [ksize^3, 32] (k = ix + ks (iy + ks iz), as eyoc_kernel_map), ksize 3 or 5;
 * coords / table_* / capacity are eyoc_hash_build's.  out: [n, 32] fp32 rows, or split-half rows (128 bytes) when out_packed
 * (range_status as eyoc_xh_pack).  nbr3 (optional) [27, n] receives the 3^3 neighbour table of the same coordinate set,
 * identical to eyoc_kernel_map_self(ksize = 3).  Same accumulation order (k ascending) and epilogue as eyoc_sparse_conv with
 * the full table.  workspace: eyoc_stem_conv_workspace_bytes(capacity) (an occupancy table, one 64-bit mask per 4x4x4 block). */
size_t eyoc_stem_conv_workspace_bytes(int64_t capacity);
int eyoc_stem_conv(const int32_t* coords, int64_t n, const uint64_t* table_keys, const int32_t* table_vals, int64_t capacity,
                   int ksize, const float* in, const float* weight, const float* scale, const float* shift, int relu, void* out,
                   int out_packed, int32_t* range_status, int32_t* nbr3, void* workspace, size_t workspace_bytes,
                   eyoc_stream_t stream);

/* fp16 hi/lo split data path of the same operator (csrc/sparse_conv_h.cu): tcgen05.mma kind::f16, activations kept in
 * HBM in the SPLIT-HALF format - a row of c channels (c % 32 == 0) is c / 32 chunks of 128 bytes, each 32 fp16 "hi"
 * values followed by 32 fp16 "lo'" values with x = hi + lo' * 2^-11 (|x| < 65504; 22 significant bits, 4 c bytes per
 * row like fp32) - so a gathered row chunk is the tensor-core operand as it lies in memory.
 * in0 / in1 / residual (when residual_packed) / out (when out_packed) are split-half; otherwise fp32 rows.
 * Weights: wt_img = eyoc_convh_split_weights(weight * wscale) with wscale a power of two chosen by the caller so that
 * max |w| * wscale lies in [2^13, 2^14); acc_scale = 1 / wscale undoes it (exactly) in the epilogue.
 * tile_masks [ceil(n_out / 256)] (eyoc_tile_masks of the tiled table) or NULL (each CTA then derives its own).
 * range_status (device int32, may be NULL; also on eyoc_xh_pack): bit 0 is OR-ed in when a value written in the split-half
 * format is not representable (|x| >= 65504 or not finite) where the fp32 reference would carry on: the caller re-runs in
 * the tf32 / fp32 data path (eyoc_b200.model.ResUNet2.forward does).
 * counters: 8 bytes of caller-owned device scratch PER LAUNCH (the persistent grid's tile-pair hand-out counters; cleared
 * here in stream order, so concurrent launches on any number of streams never share state).
 * Supported shapes as reported by eyoc_sparse_conv_h_supported (those of the tf32 path). */
size_t eyoc_convh_weight_image_halves(int K, int cin, int cout);
int eyoc_convh_split_weights(const float* weight, int K, int cin, int cout, float wscale, void* wt_img, eyoc_stream_t stream);
int eyoc_xh_pack(const float* x, int64_t n, int c, void* xh, int32_t* range_status, eyoc_stream_t stream);
int eyoc_xh_unpack(const void* xh, int64_t n, int c, float* x, eyoc_stream_t stream);
/* out = in * scale + shift per channel, split-half in and out: a stand-alone eval ME.MinkowskiBatchNorm between two blocks
 * (model/resunet.py:404-408, the ResUNetExpanded variants).  In-place (out == in) is allowed. */
int eyoc_xh_affine(const void* in, int64_t n, int c, const float* scale, const float* shift, void* out, int32_t* range_status,
                   eyoc_stream_t stream);
int eyoc_tile_masks(const int32_t* nbr_tiled, int K, int64_t n_out, uint32_t* masks, eyoc_stream_t stream);
int eyoc_sparse_conv_h_supported(int c0, int c1, int cout, int K, int l2norm);
int eyoc_sparse_conv_h(const void* in0, int c0, const void* in1, int c1, const int32_t* nbr, int K, int64_t n_out,
                       const int32_t* row_perm, int nbr_tiled, const uint32_t* tile_masks, const void* wt_img,
                       float acc_scale, const float* scale, const float* shift, const void* residual, int residual_packed,
                       int relu, int l2norm, void* out, int out_packed, int cout, int32_t* range_status, uint32_t* counters,
                       eyoc_stream_t stream);
/* Test aid of eyoc_sc2pcr: 1 (default) = the leading-eigenvector power iteration of a pair runs in one launch on a cluster of
 * 8 CTAs; 0 = one launch per iteration (the path taken anyway when n is too large for the cluster's shared memory).  Both
 * return the same bits. */
int eyoc_debug_sc2_power_fused(int on);
/* Test aid of eyoc_sc2pcr: run the kernels that csr_fill_kernel (bit 0) and seed_fitness_kernel (bit 1) replaced; the results must
 * not change by a bit. */
int eyoc_debug_sc2_reference_kernels(int mask);
/* Test aid: cap gridDim.x of the persistent grid (0 = one CTA per SM), so that small inputs walk many tile pairs per CTA. */
int eyoc_debug_convh_grid_cap(int max_ctas);
/* Tuning aid: 1 = one MMA-issuing thread in the 128-channel instantiation (default), 2 = one per accumulator tile. */
int eyoc_debug_convh_wide_issuers(int n);
/* Measurement aids of the split-half kernel, as eyoc_debug_conv_ablate / eyoc_debug_conv_times above. */
int eyoc_debug_convh_ablate(int flags);
int eyoc_debug_convh_times(long long* host_out_1024x6);
/* bit 4 of the flags: CTAs 200..203 record clock64 per work item (first 96) at {producer: empty-wait start, end, arrival;
 * MMA thread: full-wait start, end, after commit}. */
int eyoc_debug_convh_trace(long long* host_out_4x96x6);

/* Development probe of the TMA row gather (cp.async.bulk.tensor.2d tile::gather4; csrc/gather4_probe.cu): `ctas` CTAs each gather
 * 256 rows of X [n_rows, 64] fp16 by index into a SWIZZLE_128B stage `reps` times; out = CTA 0's last stage (32 KB),
 * cycles [ctas] = clock64 spent.  tools/gather4_probe.py checks the layout and prints the rate. */
int eyoc_debug_gather4_probe(const void* X, int64_t n_rows, const int32_t* idx, int n_idx, int ctas, int reps, int box_rows, void* out,
                             long long* cycles, eyoc_stream_t stream);

/* ---------------------------------------------------------------- robust linearised pose (validation path)
 * util/transform_estimation.py:89-116 est_quad_linear_robust: `iterations` (reference: 20) rounds of a weighted small-angle
 * 6-DoF solve from pts0 [n, 3] to pts1 [n, 3] with the reference's reweighting (w = par / (|r| + par), par halved every
 * five rounds); weight [n] or NULL (ones).  trans_4x4 [16] row-major maps pts0 onto pts1.  One CTA; normal equations and
 * the 6 x 6 solve in fp64. */
size_t eyoc_irls_workspace_bytes(int64_t n);
int eyoc_irls_pose(const float* pts0, const float* pts1, const float* weight, int64_t n, int iterations, float* trans_4x4,
                   void* workspace, size_t workspace_bytes, eyoc_stream_t stream);

/* ---------------------------------------------------------------- host-side index planning (no device work)
 * The six numpy draws the reference makes per pair on the global legacy RandomState (scripts/test_kitti.py:33-34,
 * :69-71 twice, scripts/SC2_PCR/SC2_PCR.py:288-289), restated on the raw MT19937 state (key[624], pos as returned by
 * np.random.get_state()) so that the stream advances exactly as under numpy.  Outputs are GLOBAL row indices into the
 * concatenated clouds (row_offsets[2p], row_offsets[2p+1] = first row of cloud 0 / 1 of pair p):
 * fc0, fc1 [num_pairs, subsample_size] (may both be NULL to skip find_corr), src, tgt [num_pairs, num_node]. */
int eyoc_plan_draws(uint32_t* mt_key624, int32_t* mt_pos, int num_pairs, const int64_t* n0, const int64_t* n1,
                    const int64_t* row_offsets, int subsample_size, int num_sample, int num_node, int64_t* fc0,
                    int64_t* fc1, int64_t* src, int64_t* tgt);

#ifdef __cplusplus
}
#endif
#endif /* EYOC_B200_H */
