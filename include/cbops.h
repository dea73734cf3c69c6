/*
 * cbops.h — C ABI of libcbops.so, the B200-native (sm_100a) point-cloud operator stack that drops
 * in behind the operator API of LiyaoTang/contrastBoundary.
 *
 * Conventions (all entry points):
 *   - plain C types only: raw DEVICE pointers, ints, a cudaStream_t passed as `void *stream`
 *     (NULL = legacy default stream).  No torch / ATen types.
 *   - the CALLER allocates every output and every workspace (the reference does the same:
 *     pytorch/lib/pointops/functions/pointops.py:21-22,40-41,57).  The library never allocates
 *     device memory and never synchronises the stream.
 *   - return value: 0 on success, negative CB_E* on failure; cb_last_error_string() gives detail.
 *     (The reference launchers return void and never check cudaGetLastError.)
 *   - `offset` / `new_offset` are int32 CUMULATIVE scene ends as in the reference
 *     (pytorch/util/s3dis.py:117-126); `*_batches` / `*_lens` (TF-side ops) are per-scene LENGTHS
 *     (tensorflow/ops/tf_custom_ops/tf_neighbors/tf_batch_neighbors.cpp:8-14).
 *   - floats are fp32, indices int32, row-major contiguous.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference
 * repo root).  INTEGRATION.md shows the reference-side binding for each.
 */
#ifndef CBOPS_H_
#define CBOPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK 0
#define CB_EINVAL (-1)     /* bad argument (NULL pointer, negative size, unsupported K, ...) */
#define CB_EWORKSPACE (-2) /* workspace too small */
#define CB_ECUDA (-3)      /* a CUDA launch / runtime call failed */
#define CB_EUNSUPPORTED (-4)

#define CB_KNN_MAX_NSAMPLE 1024 /* same ceiling as the reference's best_dist[1024] (knnquery_cuda_kernel.cu:89-90) */

int cb_version(void);
const char *cb_last_error_string(void);
unsigned long long cb_launch_count(void); /* kernels launched by this library so far (host-side counter) */

/* ------------------------------------------------------------------------------------------------
 * a1  K-nearest-neighbour query            replaces knnquery_cuda_launcher
 *     pytorch/lib/pointops/src/knnquery/knnquery_cuda_kernel.h:10-17 (kernel .cu:65-119)
 *
 * Uniform-grid search; results are bit-identical to the reference's brute-force heap kernel:
 * idx (m,nsample) int32, dist2 (m,nsample) f32 = squared distance t=dy*dy; t=fma(dx,dx,t);
 * t=fma(dz,dz,t) (the reference's SASS), ascending, short scenes padded with (scene_start, 1e10).
 * Queries whose result depends on the reference's heap mechanics (exact d2 ties) are re-run
 * through an exact replay of that heap.
 *   n, b            number of support points / scenes (the reference launcher infers them)
 *   new_xyz         may equal xyz (self query)
 *   sqrt_dist       bit 0: write sqrtf(dist2) instead (what pointops.py:43 returns); bit 1 ("set semantics"): the caller
 *                   uses the result as a SET — ties inside it may come out in any order (no heap replay for them)
 *   workspace       >= cb_knn_workspace_bytes(n, m, b) bytes, 256-byte aligned
 * ---------------------------------------------------------------------------------------------- */
size_t cb_knn_workspace_bytes(int n, int m, int b);
int cb_knn_query(int m, int nsample, const float *xyz, int n, const float *new_xyz, const int *offset,
                 const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, void *workspace,
                 size_t workspace_bytes, void *stream);

/* Split form: build the search grid of a support set once, query it many times (different
 * query sets / K).  `grid` is the workspace filled by cb_grid_build; nsample_hint tunes the cell
 * size (use the K you will query most). */
int cb_grid_build(const float *xyz, int n, const int *offset, int b, int nsample_hint, void *grid,
                  size_t grid_bytes, void *stream);
int cb_knn_query_grid(int m, int nsample, const float *xyz, int n, const float *new_xyz, const int *offset,
                      const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, void *grid,
                      size_t grid_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a1+a3 fused KNN + neighbour-feature gather (the north-star kernel; reference = knnquery then
 *     feat[idx.long()] in queryandgroup, pytorch/lib/pointops/functions/pointops.py:79-100)
 * grouped (m,nsample,c) = feat[idx]; also writes idx and dist2 as cb_knn_query does.
 * ---------------------------------------------------------------------------------------------- */
int cb_knn_gather(int m, int nsample, int c, const float *xyz, int n, const float *new_xyz,
                  const float *feat, const int *offset, const int *new_offset, int b, int *idx, float *dist2,
                  float *grouped, void *workspace, size_t workspace_bytes, void *stream);
/* split form on a grid built by cb_grid_build (c % 4 == 0, 16-byte aligned feat/grouped, nsample <= 256) */
int cb_knn_gather_grid(int m, int nsample, int c, const float *xyz, int n, const float *new_xyz,
                       const float *feat, const int *offset, const int *new_offset, int b, int *idx, float *dist2,
                       float *grouped, void *grid, size_t grid_bytes, void *stream);
float cb_knn_set_occupancy(float factor);       /* tuning knob: target points per occupied grid cell = factor * K (default 0.45) */
int cb_knn_gather_set_spin_ns(int ns);          /* tuning knob: nanosleep back-off of search warps on a full hand-off queue (default 0) */
int cb_knn_gather_set_chunk_bytes(int bytes);   /* tuning knob: bytes per TMA chunk (default 8192); returns the value in use */
int cb_knn_gather_set_mode(int mode);   /* tuning knob: -1 (default) = 7 for rows up to 256 B, else 3;
                                           3 = 7 search warps + 1 TMA copy warp per CTA (6-slab ring);
                                           0 = every warp searches and copies (TMA); 1,2 = other ring depths; 4 = register copy;
                                           5 = 6 search warps + a loader warp + a storer warp (same speed as 3);
                                           6,7 = 7 / 6 search warps + 1 / 2 load-store-unit copy warps (no TMA, no staging) */
int cb_knn_gather_set_l2_hint(int bits); /* tuning knob: L2 policy of the TMA copies; bit 0 = output stores evict_first,
                                           bit 1 = feature-row loads evict_last (default 3) */

/* ------------------------------------------------------------------------------------------------
 * a2  farthest point sampling             replaces furthestsampling_cuda_launcher
 *     pytorch/lib/pointops/src/sampling/sampling_cuda_kernel.h (kernel .cu:14-171)
 * idx (new_offset[b-1]) int32, bit-identical to the reference including its tie rule
 * (block of opt_n_threads(n_max) strided threads + lower-slot-wins tree).
 * n_max = max scene length (host-known, as in the reference launcher); tmp (n) f32 scratch is the
 * running min-distance buffer the reference also takes — on return it holds the same values.
 * ---------------------------------------------------------------------------------------------- */
int cb_furthest_sampling(int b, int n_max, const float *xyz, const int *offset, const int *new_offset,
                         float *tmp, int *idx, void *stream);
/* same results; with a workspace (>= cb_knn_workspace_bytes(n, 0, b), 256-byte aligned) scenes of
 * 8192 < n_max <= 86016 points use the bucket-pruned kernels (grid-sorted supports, O(n/j) work in iteration j
 * instead of O(n)): a thread-block cluster per scene with points and min-distances resident in the cluster's
 * shared memory and a distributed-shared-memory all-to-all arg-max (default), or the single-CTA variant. */
int cb_furthest_sampling_ws(int b, int n_max, const float *xyz, int n, const int *offset, const int *new_offset,
                            float *tmp, int *idx, void *workspace, size_t workspace_bytes, void *stream);
/* developer knob: mode 0 cluster bucket kernel, 8 CTAs x 8 warps (default) | 1 single-CTA bucket kernel |
 * 2 / 3 / 4 cluster kernel with 4x4 / 8x4 / 4x8 (CTAs x warps);
 * ws_min: scenes up to this many points use the register-resident kernels.  Returns the mode in force. */
int cb_fps_set_mode(int mode, int ws_min);

/* ------------------------------------------------------------------------------------------------
 * a3  grouping                            replaces grouping_{forward,backward}_cuda_launcher
 *     pytorch/lib/pointops/src/grouping/grouping_cuda_kernel.h (.cu:5-25)
 * forward: output(m,nsample,c) = input[idx];  backward: grad_input(n,c) += scatter(grad_output)
 * (grad_input must be zero-filled by the caller, as pointops.py:72 does).
 * ---------------------------------------------------------------------------------------------- */
int cb_grouping_forward(int m, int nsample, int c, const float *input, const int *idx, float *output, void *stream);
int cb_grouping_backward(int m, int nsample, int c, const float *grad_output, const int *idx, float *grad_input,
                         void *stream);

/* a4  subtraction   subtraction_{forward,backward}_cuda_launcher (subtraction_cuda_kernel.cu:5-30) */
int cb_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2, const int *idx,
                           float *output, void *stream);
int cb_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output,
                            float *grad_input1, float *grad_input2, void *stream);

/* a4  aggregation   aggregation_{forward,backward}_cuda_launcher (aggregation_cuda_kernel.h:14-21, .cu:5-39)
 * output(n,c) = sum_k (input[idx[n,k],c] + position[n,k,c]) * weight[n,k,c % w_c]  (output overwritten). */
int cb_aggregation_forward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                           const float *weight, const int *idx, float *output, void *stream);
int cb_aggregation_backward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                            const float *weight, const int *idx, const float *grad_output, float *grad_input,
                            float *grad_position, float *grad_weight, void *stream);

/* a6  interpolation  interpolation_{forward,backward}_cuda_launcher (interpolation_cuda_kernel.cu:5-33)
 * output(n,c) = sum_i input[idx[n,i],c] * weight[n,i]  (output overwritten). */
int cb_interpolation_forward(int n, int c, int k, const float *input, const int *idx, const float *weight,
                             float *output, void *stream);
int cb_interpolation_backward(int n, int c, int k, const float *grad_output, const int *idx, const float *weight,
                              float *grad_input, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a4  fused PointTransformer local aggregation (vector self-attention over the K neighbours)
 *     replaces PointTransformerLayer.forward, pytorch/model/blocks.py:31-44 — and with it the
 *     subtraction / aggregation kernels it subsumes (pointops.py:103-161) and ~25 torch kernels.
 * Given x_q, x_k, x_v = linear_{q,k,v}(x) (n,c), neighbour idx (n,k) and rel = p[idx]-p (cb_pt_rel):
 *     out = sum_k (x_v[idx] + pr) * softmax_k(linear_w(x_k[idx] - x_q + pr)),  pr = linear_p(rel)
 * with training-mode BatchNorm statistics computed in-kernel (running stats updated, momentum/eps as
 * torch.nn.BatchNorm1d).  CbPtLayer holds DEVICE pointers to the layer's parameters in the layout of
 * the reference's state_dict (linear weights row-major [out][in]).
 * Buffers (caller-allocated): w2buf, abuf (n,k,c/8) f32 and bnbuf (cb_pt_bnbuf_floats(c)) are saved
 * for backward; stats = cb_pt_stats_doubles(c) doubles of scratch.  c in {32,64,128,256,512}, k <= 32.
 * ld = row stride (floats) of x_q / x_k / x_v and of their gradients: c for separate tensors, 3c when they are the
 * three column blocks of one fused (n,3c) projection.
 * ---------------------------------------------------------------------------------------------- */
typedef struct CbPtLayer {
    const float *w1, *b1;                                   /* linear_p.0  (3,3),(3)   */
    const float *bn1_weight, *bn1_bias; float *bn1_running_mean, *bn1_running_var;   /* linear_p.1 */
    const float *w2, *b2;                                   /* linear_p.3  (c,3),(c)   */
    const float *bn2_weight, *bn2_bias; float *bn2_running_mean, *bn2_running_var;   /* linear_w.0 */
    const float *w3, *b3;                                   /* linear_w.2  (c/8,c),(c/8) */
    const float *bn3_weight, *bn3_bias; float *bn3_running_mean, *bn3_running_var;   /* linear_w.3 */
    const float *w4, *b4;                                   /* linear_w.5  (c/8,c/8),(c/8) */
    float momentum, eps;
    int training;
} CbPtLayer;

size_t cb_pt_bnbuf_floats(int c);
size_t cb_pt_stats_doubles(int c);
/* rel (n,k,3) = p[idx] - p[n] and its 9 moments (3 sums, 6 second moments) — once per level */
int cb_pt_rel(int n, int k, const float *p, const int *idx, float *rel, double *moments, void *stream);
int cb_pt_layer_forward(int n, int k, int c, int ld, const CbPtLayer *L, const float *rel, const double *moments,
                        const int *idx, const float *xq, const float *xk, const float *xv, float *out,
                        float *w2buf, float *abuf, float *bnbuf, double *stats, float *w0buf, void *stream);
/* w0buf: optional (n,k,c) output (training only): the pre-BatchNorm activation w0 = x_k[idx] - x_q + pr, kept so that
 * cb_pt_layer_backward can run its two dense contractions as tensor-core GEMMs over streamed operands; NULL = not kept. */

/* ------------------------------------------------------------------------------------------------
 * a12  radius neighbours        replaces batch_nanoflann_neighbors / op BatchOrderedNeighbors
 *      tensorflow/ops/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:213-336, tf_batch_neighbors.cpp:8-120
 * Distances in nanoflann's arithmetic (no fma, nanoflann.hpp:432-440), strict d2 < r^2, rows ascending,
 * padded with ns (the shadow index).  q_offset / s_offset are CUMULATIVE ends (the python facade converts
 * the reference's per-scene lengths).  Two steps because the row width is data dependent:
 *   cb_radius_count: builds the support grid in `workspace`, writes counts[nq] and *max_count (device int);
 *   cb_radius_fill : rows of `width` nearest (width = max count for the reference's behaviour, or a
 *                    smaller neighbourhood limit: datasets/base.py:762 crops columns right afterwards).
 * Tie order inside a row is unspecified in the reference (std::sort on distance only).  width <= 256.
 * ---------------------------------------------------------------------------------------------- */
int cb_radius_count(int nq, const float *queries, int ns, const float *supports, const int *q_offset,
                    const int *s_offset, int b, float radius, int *counts, int *max_count, void *workspace,
                    size_t workspace_bytes, void *stream);
int cb_radius_fill(int nq, int width, const float *queries, int ns, const float *supports, const int *q_offset,
                   const int *s_offset, int b, float radius, int *neighbors, void *workspace, size_t workspace_bytes,
                   void *stream);

/* ------------------------------------------------------------------------------------------------
 * a11  grid subsampling         replaces batch_grid_subsampling (op BatchGridSubsampling) and the CPython
 *      grid_subsampling.compute   tf_subsampling/grid_subsampling/grid_subsampling.cpp:6-162,
 *                                 cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106
 * cb_grid_subsample_cells (device): voxel barycentres (+ feature means), bit-exact (fp32 sums in arrival
 *   order), in voxel-key order, plus each voxel's key (scene << 44 | key) and first input index.
 * cb_unordered_map_order (HOST function on host arrays): the permutation that puts the voxels in the
 *   reference's output order, i.e. the iteration order of its std::unordered_map<size_t, ...>.
 * cb_grid_subsample_permute (device): apply it.
 * ---------------------------------------------------------------------------------------------- */
size_t cb_grid_subsample_workspace_bytes(int n, int b, int fdim);
int cb_grid_subsample_cells(const float *xyz, int n, const int *offset, int b, float dl, const float *feat, int fdim,
                            float *cells_xyz, float *cells_feat, unsigned long long *cells_key, int *cells_first_idx,
                            int *point_cell /* (n) voxel (key order) of every input point, may be NULL */,
                            int *ncells_dev, void *workspace, size_t workspace_bytes, void *stream);
int cb_label_vote_host(const int *labels, int count);   /* HOST: tie rule of the reference's label vote */
int cb_unordered_map_order(const unsigned long long *keys, const int *first_idx, int ncells, int b, int *perm,
                           int *scene_counts);
int cb_grid_subsample_permute(int ncells, int fdim, const int *perm, const float *in_xyz, const float *in_feat,
                              float *out_xyz, float *out_feat, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a9  fused contrastive-boundary loss of one stage        replaces ContrastHead.point_contrast
 *     pytorch/model/heads.py:185-246 (+ dist_l2 :116-119, posmask_cnt :145-149, contrast_softnn
 *     :151-165) and the sub-scene label propagation of pytorch/model/basic_operators.py:9-50.
 * cb_cbl_classes: cls[i] = target[i] (label_idx NULL) or the arg-max class among the kr nearest
 *   full-resolution labels target[label_idx[i,:]] (first maximum).  target is int64 as in torch.
 * cb_cbl_forward : sums[0] += sum of loss_i over boundary points, sums[1] += their count (sums zeroed by the caller);
 *   idx (m,K) int32 with column 0 = the point itself (dropped as heads.py:196); d in {32,64,72}; K <= 65.
 * cb_cbl_backward: grad_feat (m,d), zero-filled by the caller, += scale[0] * d(sum loss_i)/d feat.
 * ---------------------------------------------------------------------------------------------- */
/* test-time boundary / plain masks of get_boundary_mask (pytorch/model/basic_operators.py:69-97; tool/test.py:392-428):
 * labels (n) int64 (negative = invalid), neighbor_idx (n,kr) int32, valid_mask (n) bytes or NULL; any output may be NULL.
 * bound_cnt[i] = #valid neighbours with another label, bound = cnt > 0, plain = all neighbours invalid-or-equal (& valid). */
int cb_boundary_mask(long long n, int kr, const long long *labels, const int *neighbor_idx, const unsigned char *valid_mask,
                     int *bound_cnt, unsigned char *bound, unsigned char *plain, void *stream);
int cb_cbl_classes(int m, int kr, int ncls, const int *label_idx, const long long *target, int *cls, void *stream);
int cb_cbl_forward(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                   float *sums, void *stream);
int cb_cbl_backward(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                    const float *scale, float *grad_feat, void *stream);
/* a14  TF flavour of the same loss (contrast_head, tensorflow/models/heads/head.py:462-807 with softnn/l2):
 * neighbours with idx >= n_valid are shadow entries of the radius search and count neither as positive nor
 * negative (solve_samples_mask :641-662); flavour 1 uses dist = sqrt(max(s, 1e-12)) (calc_dist :183-185)
 * instead of sqrt(s + 1e-12); the ratio is pos / (pos + neg) over valid neighbours (:750-771). */
int cb_cbl_forward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                      float *sums, int n_valid, int flavour, void *stream);
int cb_cbl_backward_ex(int m, int K, int D, const float *feat, const int *idx, const int *cls, float temperature,
                       const float *scale, float *grad_feat, int n_valid, int flavour, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a13  AdaptiveWeight ("ConvNet") local aggregation        tensorflow/models/local_aggregation_operators.py:316-500
 *      with config/s3dis/adapt.yaml:19-26 (input 'dp', one FC, shared_channels 1, mean reduction, no softmax):
 *      out[n,c] = sum_k (W[c,:].((support[idx]-query[n])/radius) + b[c]) * feat[idx[n,k],c] / (cnt[n] + 1e-5)
 * shadow neighbours (idx >= n0) contribute zero; cnt = #(idx < max(idx)) as the reference counts (:466-470);
 * pad_num (1 device int) receives max(idx) and is reused by the backward.  grad_feat / grad_W / grad_b must be
 * zero-filled by the caller.
 * ---------------------------------------------------------------------------------------------- */
int cb_adaptive_weight_forward(int n, int k, int c, int n0, const float *query_pts, const float *support_pts,
                               const int *idx, const float *feat, const float *W, const float *bias, float radius,
                               int *pad_num, float *out, void *stream);
int cb_adaptive_weight_backward(int n, int k, int c, int n0, const float *query_pts, const float *support_pts,
                                const int *idx, const float *feat, const float *W, const float *bias, float radius,
                                const int *pad_num, const float *grad_out, float *grad_feat, float *grad_W,
                                float *grad_b, void *stream);

/* ------------------------------------------------------------------------------------------------
 * a5  fused TransitionDown (stride > 1)          replaces the body of TransitionDown.forward, pytorch/model/blocks.py:69-73
 *     out[m,c'] = max_k relu(bn( Wxyz (p[idx]-p_new) + z[idx] )),  z = x Wf^T  (the bias-free Linear(3+c -> c') applied
 *     BEFORE the gather; W = [Wxyz | Wf]).  rel (m,k,3) from cb_td_rel.  argk (m,c') uint8 and bnbuf (4c' floats) are
 *     saved for backward; stats = 2c' doubles of scratch.  Backward: grad_z (n,c') and grad_wxyz (c',3) must be
 *     zero-filled by the caller; grad_bn_* are overwritten; scratch >= 2c' doubles + 3c' floats, 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int cb_td_rel(int m, int k, const float *p_support, const float *p_query, const int *idx, float *rel, void *stream);
int cb_td_forward(int m, int k, int c, const float *rel, const int *idx, const float *z, const float *wxyz,
                  const float *bn_weight, const float *bn_bias, float *running_mean, float *running_var, float momentum,
                  float eps, int training, float *out, unsigned char *argk, float *bnbuf, double *stats, void *stream);
int cb_td_backward(int m, int k, int c, const float *rel, const int *idx, const float *z, const float *wxyz,
                   const float *bn_weight, int training, const float *bnbuf, const float *out, const unsigned char *argk,
                   const float *grad_out, float *grad_z, float *grad_wxyz, float *grad_bn_weight, float *grad_bn_bias,
                   float *scratch, void *stream);

/* ------------------------------------------------------------------------------------------------
 * tall-skinny FP32 linear layers of the per-point MLPs (nn.Linear calls of blocks.py:33,72,76,108,
 * 127-131 and the heads' MLPs): Y = X W^T + b ; dX = G W ; dW = G^T X, db = sum G (dW/db overwritten).
 * ---------------------------------------------------------------------------------------------- */
int cb_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, void *stream);
int cb_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, void *stream);
int cb_linear_wgrad(int n, int ci, int co, const float *X, const float *G, float *dW, float *db, void *stream);
/* fused BatchNorm1d (+ residual) (+ ReLU) over an (n, c) matrix — relu(bn(linear(x))) and relu(bn3(.) + identity) of
 * PointTransformerBlock / TransitionDown / TransitionUp / the head MLPs (pytorch/model/blocks.py:76,104-108,125-133).
 * forward: y = act((x - mean) * invstd * gamma + beta [+ residual]); training: batch statistics (biased variance), running
 * statistics updated with the unbiased one (nn.BatchNorm1d semantics); bnbuf (4c floats: scale, shift, mean, invstd) and
 * stats (2c doubles, scratch) are caller-allocated; bnbuf is kept for the backward.
 * backward: grad_x, grad_residual (optional, = grad_y masked by the ReLU), grad_gamma, grad_beta; sums: 2c doubles scratch.
 * training: bit 0 = batch statistics; bit 1 = stats / sums is a PERSISTENT accumulator of 2c + 1 doubles (the last one holds
 * an int ticket) that is all-zero on entry and left all-zero on exit (the last block of the apply kernel clears it): no
 * memset is issued.  Without bit 1 the buffer is plain scratch and is cleared by a cudaMemsetAsync first. */
int cb_bn_act_forward(long long n, int c, const float *x, const float *residual, const float *gamma, const float *beta,
                      float *running_mean, float *running_var, float momentum, float eps, int training, int relu, float *y,
                      float *bnbuf, double *stats, void *stream);
int cb_bn_act_backward(long long n, int c, const float *x, const float *y, const float *gamma, const float *bnbuf, int training,
                       int relu, const float *grad_y, float *grad_x, float *grad_residual, float *grad_gamma, float *grad_beta,
                       double *sums, void *stream);

/* 1 (default): the three calls above run on the tensor cores with 3xTF32 error compensation (hi/lo operand split,
 * FP32 accumulate; tc_gemm.cu) and are HBM-bound; 0: exact-FP32 SIMT kernels (and the ci*co <= 16384 limit of the SIMT
 * wgrad).  Returns the setting in force. */
int cb_linear_set_tensor_cores(int on);

/* tall-skinny linear layers on the 5th-generation tensor cores (umma_linear.cu): tcgen05.mma kind::tf32 issued by one thread
 * per CTA, operands staged in shared memory as TF32 hi + lo (3xTF32: FP32 parity, 1e-5 vs float64), accumulator in tensor
 * memory (TMEM), tcgen05.ld epilogue.  cb_linear_forward / cb_linear_dgrad route here when the shape allows
 * (ci % 8 == 0 and co % 16 == 0, resp. co % 8 == 0 and ci % 16 == 0; 16-byte aligned rows); cb_linear_set_umma(0) returns
 * to the mma.sync kernels.  Same math as the nn.Linear calls of pytorch/model/blocks.py:33,72,76,108,127-131. */
int cb_umma_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, void *stream);
int cb_umma_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, void *stream);
int cb_linear_set_umma(int on);
int cb_linear_set_umma_version(int v);   /* 2 (default): persistent warp-specialised pipeline (TMA-less producer warps, MMA
                                            issuer, epilogue warps, double-buffered TMEM accumulators); 1: one tile per CTA */
int cb_grid_set_fused(int mode);   /* 1 (default): the search-grid build is ONE cooperative kernel; 0: ten small kernels; 2: fused unless capturing */

/* backward of cb_pt_layer_forward.  grad_xk / grad_xv (n,c) and grad_params must be ZERO-FILLED by the
 * caller (scatter / accumulation targets); grad_xq is overwritten.  grad_params layout (floats):
 * [dW1 9][db1 3][dbn1_w 3][dbn1_b 3][dW2 3c][db2 c][dbn2_w c][dbn2_b c][dW3 c*c/8][db3 c/8][dbn3_w c/8]
 * [dbn3_b c/8][dW4 (c/8)^2][db4 c/8].  scratch: cb_pt_bwd_scratch_floats(n,k,c) floats, 16-byte aligned. */
size_t cb_pt_bwd_scratch_floats(int n, int k, int c);
int cb_pt_layer_backward(int n, int k, int c, int ld, const CbPtLayer *L, const float *rel, const int *idx,
                         const float *xq, const float *xk, const float *xv, const float *w2buf,
                         const float *abuf, const float *bnbuf, const float *grad_out, float *grad_xq,
                         float *grad_xk, float *grad_xv, float *grad_params, float *scratch, const float *w0buf,
                         void *stream);
/* w0buf: the buffer cb_pt_layer_forward filled, or NULL (FP32 SIMT kernels that re-gather and recompute). */
/* 1 (default): the (n*k) x c x c/8 contraction of the layer (linear_w[2], blocks.py:40) runs on the tensor cores
 * (mma.sync m16n8k8 TF32 with 3xTF32 error compensation, ptlayer_mma.cu); 0: FP32 SIMT kernels.  Returns the setting. */
int cb_pt_set_tensor_cores(int on);

/* ------------------------------------------------------------------------------------------------
 * config 3 (ConvNet = AdaptiveWeight ResNet + CBL, TF tree): the device operators around cb_adaptive_weight_* (convnet_ops.cu)
 * cb_ind_max_pool_*      ind_max_pool, tensorflow/models/basic_operators.py:155-172 (strided-bottleneck shortcut,
 *                        models/backbone/resnet.py:268): out[i,c] = max_k [x ; colmin][inds[i,k], c]; colmin (c) = the
 *                        column minima of x (the reference's shadow row); arg (n2,c) uint8 saved for the backward
 *                        (255 = shadow row won); grad_x zero-filled by the caller.  c % 4 == 0, k < 255.
 * cb_label_vote_idx      hard sub-scene label = arg-max (first maximum) of the histogram of target[label_idx[i,:]], entries
 *                        outside [0, n_valid) skipped (head.py:25-49 with reduction 'max', :117-131)
 * cb_label_vote_radius   the same vote over ALL supports within `radius` of the query (head.py:158-176 runs a radius
 *                        search with r_sample[i-1] for stages >= 2 and gathers thousands of labels per coarse point; here
 *                        the histogram is built inside the search).  workspace >= cb_knn_workspace_bytes(ns, 0, b).
 * ---------------------------------------------------------------------------------------------- */
int cb_ind_max_pool_forward(int n2, int k, int c, int n1, const float *x, const float *colmin, const int *inds, float *out,
                            unsigned char *arg, void *stream);
int cb_ind_max_pool_backward(int n2, int k, int c, const int *inds, const unsigned char *arg, const float *grad_out,
                             float *grad_x, void *stream);
int cb_label_vote_idx(int m, int kr, int ncls, int n_valid, const int *label_idx, const long long *target, int *cls,
                      void *stream);
int cb_label_vote_radius(int nq, const float *queries, int ns, const float *supports, const int *q_offset,
                         const int *s_offset, int b, float radius, int ncls, const long long *target, int *cls,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * f1  the step BEFORE the path, on the device (dataprep.cu): voxelize + data_prepare + collate
 *     pytorch/util/voxelize.py:4-16,38-56   pytorch/util/data_util.py:45-92   pytorch/util/s3dis.py:94-130
 * coord / feat are float32 or float64 device arrays (coord_is_f64); the arithmetic (shift, coord / voxel_size, floor,
 * squared crop distance) runs in that dtype, as NumPy's does; outputs are float32 / int64 like the torch tensors the
 * reference builds.  Bit-exact: voxel keys (FNV64-1A), occupied voxels, counts, voxel order, crop distances, both
 * min-shifts, feat / 255.  Defined modulo (reference = NumPy global RNG / unstable sort): the point kept per voxel,
 * the order among equal crop distances, the shuffle permutation.  See dataprep.cu for every argument.
 * cb_voxelize      = voxelize(coord, voxel_size, 'fnv', mode=1) -> idx_sort (n), count (first *nvox entries), *nvox
 * cb_data_prepare  = data_prepare of ONE cloud, appended to a batch buffer at the DEVICE-side row offset row_offset[0];
 *                    row_offset[1] = row_offset[0] + kept count (chain the calls of a batch over a (B+1)-long array
 *                    starting at 0: its tail is the collate `offset`).  No device->host read anywhere.
 * ---------------------------------------------------------------------------------------------- */
size_t cb_data_prepare_workspace_bytes(int n);
int cb_voxelize(const void *coord, int coord_is_f64, int n, double voxel_size, int *idx_sort, int *count, int *nvox,
                unsigned long long *keys_sorted, void *workspace, size_t workspace_bytes, void *stream);
int cb_data_prepare(const void *coord, const void *feat, int coord_is_f64, int fdim, const long long *label, int n,
                    double voxel_size, int voxel_max, int pick_mode, int centre_mode, int shuffle, unsigned long long seed,
                    float feat_div, int *row_offset, int out_capacity, float *out_coord, float *out_feat,
                    long long *out_label, int *out_index, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * The reference's OWN native ABI (compat.cu): the ten `extern "C"` launchers its pybind glue calls,
 * with the reference's exact names and argument lists, so that the unmodified glue
 * (pytorch/lib/pointops/src/<op>/<op>_cuda.cpp + pointops_api.cpp) links against libcbops.so in place of
 * its own .cu objects.  Declarations: knnquery_cuda_kernel.h:10-17, sampling_cuda_kernel.h,
 * grouping_cuda_kernel.h, interpolation_cuda_kernel.h, subtraction_cuda_kernel.h,
 * aggregation_cuda_kernel.h:14-21.  They launch on the stream given to cb_compat_set_stream (default NULL =
 * the legacy default stream the reference launches on); knnquery reads the scene / support counts back
 * from the device offsets (one stream synchronise per call) and uses a library-owned grow-only workspace.
 * Being `void`, they report failures on stderr + cb_last_error_string().
 * ---------------------------------------------------------------------------------------------- */
void cb_compat_set_stream(void *stream);
void knnquery_cuda_launcher(int m, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                            const int *new_offset, int *idx, float *dist2);
void furthestsampling_cuda_launcher(int b, int n, const float *xyz, const int *offset, const int *new_offset,
                                    float *tmp, int *idx);
void grouping_forward_cuda_launcher(int m, int nsample, int c, const float *input, const int *idx, float *output);
void grouping_backward_cuda_launcher(int m, int nsample, int c, const float *grad_output, const int *idx,
                                     float *grad_input);
void interpolation_forward_cuda_launcher(int n, int c, int k, const float *input, const int *idx, const float *weight,
                                         float *output);
void interpolation_backward_cuda_launcher(int n, int c, int k, const float *grad_output, const int *idx,
                                          const float *weight, float *grad_input);
void subtraction_forward_cuda_launcher(int n, int nsample, int c, const float *input1, const float *input2,
                                       const int *idx, float *output);
void subtraction_backward_cuda_launcher(int n, int nsample, int c, const int *idx, const float *grad_output,
                                        float *grad_input1, float *grad_input2);
void aggregation_forward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                       const float *weight, const int *idx, float *output);
void aggregation_backward_cuda_launcher(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                        const float *weight, const int *idx, const float *grad_output,
                                        float *grad_input, float *grad_position, float *grad_weight);

/* developer hook (tools/debug_fps.py): copies the cycle / bucket counters of the cluster FPS kernel to out8 (touched buckets,
 * iterations, cycles of warp 0: total / refresh / exchange, max buckets per warp-iteration, ...), clears them, and arms
 * (enable = 1) or disarms (0) the counting. */
int cb_debug_fps(int enable, unsigned long long *out8);

/* tuning knob: 1 (default) = kernels that support it are launched with programmatic stream serialization (their set-up
 * overlaps the tail of the previous kernel; they wait for it with griddepcontrol.wait before touching memory); 0 = ordinary
 * launches.  Returns the setting in force. */
int cb_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------
 * segmentation cross-entropy      replaces nn.CrossEntropyLoss(ignore_index)(output, target) of
 *                                 pytorch/model/pointtransformer_seg.py:15-25 (mean over the rows whose target != ignore_index)
 * logits (n, c) float32 row-major, target (n) int64; acc: 3 doubles of scratch kept for the backward (sum, count, ticket);
 * loss: 1 float.  backward: grad_logits = (softmax - onehot) * grad_loss[0] / count, zero rows for ignored targets.
 * ------------------------------------------------------------------------------------------------ */
int cb_cross_entropy_forward(int n, int c, const float *logits, const long long *target, long long ignore_index, double *acc,
                             float *loss, void *stream);
int cb_cross_entropy_backward(int n, int c, const float *logits, const long long *target, long long ignore_index,
                              const double *acc, const float *grad_loss, float *grad_logits, void *stream);

/* ------------------------------------------------------------------------------------------------
 * optimiser step, all tensors in one launch    replaces torch.optim.SGD(...).step() of pytorch/tool/train.py:154,324
 * d = g + weight_decay * p;  m = first_step ? d : momentum * m + d;  p -= lr * m
 * g: the packed gradient (total floats); off: ntensors + 1 prefix offsets into g (device, int64);
 * pp / mp: device arrays of ntensors device pointers (parameter / momentum buffer of tensor t, off[t+1] - off[t] floats)
 * ------------------------------------------------------------------------------------------------ */
int cb_sgd_momentum_step(long long total, int ntensors, const long long *off, float *const *pp, float *const *mp,
                         const float *g, float lr, float momentum, float weight_decay, int first_step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CBOPS_H_ */
