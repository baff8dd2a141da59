/*
 * cpfn_b200.h -- C ABI of libcpfn_b200.so, the sm_100a implementation of the
 * CPFN data-parallel hot path (PointNet++ grouping backbone + SPFN weighted
 * total-least-squares fitters).
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name ends in _host;
 *   - tensors are dense, row-major, float32 / int32 exactly as the reference
 *     pybind module takes them (PointNet2/pointnet2_ops/cuda_ops/src/
 *     bindings.cpp:6-19, checks in include/utils.h:5-25);
 *   - outputs and workspaces are caller-allocated; the library never
 *     allocates device memory, never synchronises and never touches torch;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL is
 *     the legacy default stream) and the call returns immediately;
 *   - the return value is 0 on success or a negative CPFN_E* code.  A failed
 *     launch is REPORTED (the reference calls exit(-1), include/
 *     cuda_utils.h:30-39; this library does not);
 *   - element offsets are computed in 64 bits (the reference overflows int
 *     past 2^31 elements per tensor).
 *
 * Paths in the "replaces" notes are relative to the reference tree.
 */
#ifndef CPFN_B200_H_
#define CPFN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPFN_OK 0
#define CPFN_EINVAL (-1)   /* bad size / null pointer / unsupported shape */
#define CPFN_ELAUNCH (-2)  /* CUDA reported a launch error (see cpfn_last_cuda_error) */
#define CPFN_EWORKSPACE (-3) /* workspace missing or too small */

#if defined(__GNUC__)
#define CPFN_API __attribute__((visibility("default")))
#else
#define CPFN_API
#endif

typedef void *cpfn_stream_t;

/* Library identity / diagnostics. */
CPFN_API int cpfn_version(void);                       /* 100 * major + minor */
CPFN_API const char *cpfn_error_string(int code);
CPFN_API const char *cpfn_last_cuda_error(void);       /* cudaGetErrorString of the last failure on this thread */
CPFN_API int cpfn_sm_count(void);                      /* SMs of the current device (<0 on error) */

/* ---------------------------------------------------------------------------
 * The nine pointnet2 ops (drop-in for the pybind module `cuda_ops`).
 * ------------------------------------------------------------------------- */

/* Furthest point sampling split over several launches: rounds j_begin .. j_end-1 of an nsamples-round sampling
 * (sample j is chosen in round j; round 0 = the first point).  idx [B,nsamples] / new_xyz [B,nsamples,3] receive
 * the samples of these rounds; `state` f32 [B,N] carries the running minimum distances from one launch to the next
 * (required unless the call covers all rounds).  Bit-identical to one launch.  Why: a sampling is a chain of
 * nsamples dependent rounds on a fraction of the SMs; cutting it lets the ball query and the set-abstraction MLP of
 * the centroids already chosen run on the other SMs while the later rounds are still being sampled
 * (fused.pointnet2_forward).  smem_floor_bytes > 0 raises the sampling CTAs' shared-memory request to that many
 * bytes so that the co-running kernels' CTAs do not fit beside them.  Only for the cluster kernel's domain:
 * cpfn_fps_rounds_supported(B, N) != 0 (2048 <= N <= 16384, B clouds x >= 2 CTAs fit the GPU). */
CPFN_API int cpfn_fps_rounds_supported(int B, int N);
CPFN_API int cpfn_furthest_point_sampling_rounds(const float *xyz, int B, int N, int nsamples, int j_begin, int j_end,
                                                 int32_t *idx, float *new_xyz, float *state, size_t state_bytes,
                                                 size_t smem_floor_bytes, cpfn_stream_t stream);

/* Diagnostic: with the environment variable CPFN_FPS_PROFILE set, the cluster FPS kernel sums, over its rounds, the
 * cycles thread 0 of CTA 0 spends in {distance update, warp arg-max, key push, waiting for the cluster's keys,
 * cluster arg-max, centroid look-up}; this copies the six sums of the last such launch to the host (synchronises). */
CPFN_API int cpfn_debug_fps_profile(long long *cycles6);

/* Diagnostic: with CPFN_CHAIN_PROFILE=1 every cpfn_mlp_chain launch sums, over its CTAs, the cycles its MMA thread,
 * its weight producer and its first worker warp spend waiting / working (16 sums per launch, slot meaning in
 * csrc/mlp_chain.cu); copies the sums of the launches since the last reset (at most 64, in launch order) to `out`
 * (16 values each), optionally resets, returns the number of launches copied.  Synchronises the device. */
CPFN_API int cpfn_debug_chain_profile(unsigned long long *out, int max_launches, int reset);

/* Furthest point sampling.  Replaces farthest_point_sampling
 * (src/sampling.cpp:65-86, kernel src/sampling_gpu.cu:63-159).
 * xyz [B,N,3] f32 -> idx [B,nsamples] i32.  Bit-exact with the reference,
 * including its tie-break (shared-memory tree order of a block of
 * opt_n_threads(N) threads) and its skip of points with |p|^2 <= 1e-3.
 * `workspace` is needed only when cpfn_fps_workspace_bytes(B,N) > 0. */
CPFN_API size_t cpfn_fps_workspace_bytes(int B, int N);
CPFN_API int cpfn_furthest_point_sampling(const float *xyz, int B, int N, int nsamples,
                                 int32_t *idx, void *workspace,
                                 size_t workspace_bytes, cpfn_stream_t stream);
/* The same sampling that also writes the sampled centroids new_xyz f32 [B,nsamples,3] = xyz[b, idx[b,s], :]
 * (the select_point_subset that follows FPS in pointset_abstraction.py:49-50) from inside the kernel; NULL = skip. */
CPFN_API int cpfn_furthest_point_sampling_xyz(const float *xyz, int B, int N, int nsamples, int32_t *idx,
                                              float *new_xyz, void *workspace, size_t workspace_bytes,
                                              cpfn_stream_t stream);

/* Ball query.  Replaces ball_query (src/ball_query.cpp:8-32, kernel
 * src/ball_query_gpu.cu:9-44).  new_xyz [B,S,3], xyz [B,N,3] -> idx
 * [B,S,nsample] i32: the first `nsample` points (ascending index) with
 * d^2 < radius*radius, padded with the first hit; all zeros when no hit. */
CPFN_API int cpfn_ball_query(const float *new_xyz, const float *xyz, int B, int N, int S,
                    float radius, int nsample, int32_t *idx,
                    cpfn_stream_t stream);

/* Ball query on a uniform grid (cells >= radius, built per call in `workspace`): bit-identical output,
 * a query only tests the points of its 27 neighbouring cells; hits are collected in an index bitmap so
 * the "first nsample in index order" rule of the reference holds whatever order the grid yields them.
 * Falls back to cpfn_ball_query for N < 2048 (the scan is faster there) or N > 32768. */
CPFN_API size_t cpfn_ball_query_grid_workspace_bytes(int B, int N);
CPFN_API int cpfn_ball_query_grid(const float *new_xyz, const float *xyz, int B, int N, int S,
                                  float radius, int nsample, int32_t *idx, void *workspace,
                                  size_t workspace_bytes, cpfn_stream_t stream);
/* The two halves of cpfn_ball_query_grid: the grid depends on (xyz, radius) only, so it can be built on another
 * stream while the queries (the SA layer's FPS centroids) are still being computed; the query half then needs
 * the same workspace.  Outside the grid kernel's range (N < 2048, N > 32768) build is a no-op and query
 * falls back to cpfn_ball_query. */
CPFN_API int cpfn_ball_query_grid_build(const float *xyz, int B, int N, float radius, void *workspace,
                                        size_t workspace_bytes, cpfn_stream_t stream);
CPFN_API int cpfn_ball_query_grid_query(const float *new_xyz, const float *xyz, int B, int N, int S, float radius,
                                        int nsample, int32_t *idx, void *workspace, size_t workspace_bytes,
                                        cpfn_stream_t stream);
/* The same query for centroids s_begin .. s_begin+s_count-1 of every cloud only (new_xyz and idx keep their full
 * [B,S,.] shapes): lets the ball query follow a furthest point sampling that is still running
 * (cpfn_furthest_point_sampling_rounds).  Grid kernel only (2048 <= N <= 32768, radius > 0). */
CPFN_API int cpfn_ball_query_grid_query_range(const float *new_xyz, const float *xyz, int B, int N, int S, int s_begin,
                                              int s_count, float radius, int nsample, int32_t *idx, void *workspace,
                                              size_t workspace_bytes, cpfn_stream_t stream);

/* Gather / group (+ gradients).  Replace gather_points(_grad)
 * (src/sampling.cpp:15-64, src/sampling_gpu.cu:8-53) and group_points(_grad)
 * (src/group_points.cpp:12-60, src/group_points_gpu.cu:8-74).
 * points [B,C,N]; idx [B,M] or [B,S,K]; out [B,C,M] or [B,C,S,K].
 * The *_grad entry points OVERWRITE grad_points [B,C,N] (zero + scatter-add). */
CPFN_API int cpfn_gather_points(const float *points, const int32_t *idx, int B, int C,
                       int N, int M, float *out, cpfn_stream_t stream);
CPFN_API int cpfn_gather_points_grad(const float *grad_out, const int32_t *idx, int B,
                            int C, int N, int M, float *grad_points,
                            cpfn_stream_t stream);
CPFN_API int cpfn_group_points(const float *points, const int32_t *idx, int B, int C,
                      int N, int S, int K, float *out, cpfn_stream_t stream);
CPFN_API int cpfn_group_points_grad(const float *grad_out, const int32_t *idx, int B,
                           int C, int N, int S, int K, float *grad_points,
                           cpfn_stream_t stream);

/* Three nearest neighbours.  Replaces three_nn (src/interpolate.cpp:14-40,
 * kernel src/interpolate_gpu.cu:9-59).  unknown [B,n,3], known [B,m,3] ->
 * dist2 [B,n,3] f32 (SQUARED distances), idx [B,n,3] i32; ties go to the
 * lower index; m < 3 leaves index 0 / dist2 = +inf in the unused slots. */
CPFN_API int cpfn_three_nn(const float *unknown, const float *known, int B, int n, int m,
                  float *dist2, int32_t *idx, cpfn_stream_t stream);

/* Three-point weighted sum (+ gradient).  Replace three_weighted_sum(_grad)
 * (src/interpolate.cpp:42-99, kernels src/interpolate_gpu.cu:72-143).
 * points [B,C,M], idx/weight [B,n,3] -> out [B,C,n];
 * grad: grad_out [B,C,n] -> grad_points [B,C,M] (overwritten). */
CPFN_API int cpfn_three_weighted_sum(const float *points, const int32_t *idx,
                            const float *weight, int B, int C, int M, int n,
                            float *out, cpfn_stream_t stream);
CPFN_API int cpfn_three_weighted_sum_grad(const float *grad_out, const int32_t *idx,
                                 const float *weight, int B, int C, int n,
                                 int M, float *grad_points,
                                 cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * SPFN weighted total-least-squares primitive fitters.
 * ------------------------------------------------------------------------- */

/* Plane, sphere, cylinder and cone parameters of every (cloud, instance slot).
 * Replaces SPFN/losses_implementation.py:255-278 (compute_parameters) and the four
 * fitters it dispatches to: plane_fitter.py:9-17, sphere_fitter.py:9-19,
 * cylinder_fitter.py:10-28, cone_fitter.py:12-36 (and through them
 * SPFN/geometry_utils.py:8-27,74-84,121-142,209-223 and differentiable_tls.py:200-209).
 * P [B,N,3] points, W [B,N,K] soft or one-hot memberships, X [B,N,3] unit normals.
 * `out` is a struct-of-arrays block of 22*B*K floats (BK = B*K):
 *   [ 0BK) plane_normal [B,K,3]        [ 3BK) plane_center [B,K]
 *   [ 4BK) sphere_center [B,K,3]       [ 7BK) sphere_radius_squared [B,K]
 *   [ 8BK) cylinder_axis [B,K,3]       [11BK) cylinder_center [B,K,3]
 *   [14BK) cylinder_radius_squared     [15BK) cone_apex [B,K,3]
 *   [18BK) cone_axis [B,K,3]           [21BK) cone_half_angle [B,K]
 * plane_normal and cylinder_axis are defined up to sign (as the reference's SVD);
 * here the component of largest magnitude is made positive.  K <= 256.
 * Four kernels (two passes over W, two per-slot solves) are enqueued on `stream`; no host synchronisation. */
CPFN_API size_t cpfn_fit_workspace_bytes(int B, int N, int K);
CPFN_API int cpfn_fit_primitives(const float *P, const float *W, const float *X, int B,
                                 int N, int K, float *out, void *workspace,
                                 size_t workspace_bytes, cpfn_stream_t stream);

/* Training path of the fitters.  Raw weighted moments up to third order, fp64 accumulation:
 *   M[b,k,f] = sum_n Wt[b,n,k] * psi_f(P[b,n], X[b,n]),  f = 0..31:
 *   [1 | p (3) | p p^T xx,xy,xz,yy,yz,zz | p p p xxx,xxy,xxz,xyy,xyz,xzz,yyy,yyz,yzz,zzz | x (3) | x x^T (6) | x (p.x) (3)]
 * (everything SPFN/differentiable_tls.py:200-209 and SPFN/geometry_utils.py:74-84,121-142,209-223 sum over
 * the points, without tiling P / X K times), and the backward of that linear map:
 *   dWt[b,n,k] = sum_f psi_f dM[b,k,f],  dX[b,n,:] = sum_k Wt[b,n,k] sum_f dM[b,k,f] dpsi_f/dx.
 * M, dM: [B,K,32] doubles.  dWt / dX may be NULL to skip them.  P receives no gradient (as in the reference). */
CPFN_API size_t cpfn_moments_workspace_bytes(int B, int N, int K);
CPFN_API int cpfn_weighted_moments(const float *P, const float *X, const float *Wt, int B, int N, int K,
                                   double *M, void *workspace, size_t workspace_bytes, cpfn_stream_t stream);
CPFN_API int cpfn_weighted_moments_grad(const float *P, const float *X, const float *Wt, const double *dM,
                                        int B, int N, int K, float *dWt, float *dX, cpfn_stream_t stream);

/* Batched tiny linear algebra of the differentiable fitters, fp64, one thread per matrix (csrc/small_linalg.cu).
 * Replaces the torch.svd of Custom_svd_v_colum (SPFN/differentiable_tls.py:123-143; symmetric operands, so the
 * singular vectors are the eigenvectors) and the torch.solve of guarded_matrix_solve_ls
 * (SPFN/geometry_utils.py:131-141, normal equations) on [B*K, D, D]-sized batches.
 *   cpfn_sym_eigh_small: A [n, D, D] symmetric, D = 2 or 3 -> lam [n, D] ascending, Q [n, D, D] eigenvectors in the
 *                        columns (Q may be NULL: eigenvalues only)
 *   cpfn_small_solve:    A [n, D, D], b [n, D], D <= 3 -> x [n, D] with A x = b (transpose != 0: A^T x = b),
 *                        Gaussian elimination with partial pivoting */
CPFN_API int cpfn_sym_eigh_small(const double *A, long long n, int D, double *lam, double *Q, cpfn_stream_t stream);
CPFN_API int cpfn_small_solve(const double *A, const double *b, long long n, int D, int transpose, double *x,
                              cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused shared-MLP chains (set abstraction, feature propagation, heads) on the
 * tcgen05 tensor cores.  Inference only: BatchNorm is folded into the weights.
 * ------------------------------------------------------------------------- */
#define CPFN_MLP_MAX_LAYERS 6
#define CPFN_MLP_IN_DENSE 0   /* rows of a_src [cols, a_ch] */
#define CPFN_MLP_IN_GROUP 1   /* [a_src[b, idx] (a_ch) | xyz[b, idx] - centers[b, col / group_k] (3)] */
#define CPFN_MLP_IN_INTERP 2  /* [a_src[col] (a_ch, skip) | sum_q nn_w[col,q] * b_src[b, idx[col,q]] (b_ch)] */
#define CPFN_MLP_OUT_ROWS 0   /* out [cols, ldo]: every column's last-layer output */
#define CPFN_MLP_OUT_POOL 1   /* out [cols / pool_g, ldo]: max over each group of pool_g columns */

typedef struct {
  int32_t cin, cout;          /* real channel counts; layer l+1 cin == layer l cout */
  int32_t relu;               /* ReLU after the bias */
  int32_t bias_per_cloud;     /* bias is [B, 128*ceil(cout/128)] instead of [128*ceil(cout/128)] */
  const float *bias;          /* BN-folded bias, zero padded to a multiple of 128 */
  const uint32_t *mask_bits;  /* optional dropout keep-mask after ReLU, one bit per (column, channel): word
                               * [col * ceil(cout/32) + ch/32], bit ch%32 (cpfn_dropout_mask_bits); cout <= 128 */
  float mask_scale;           /* kept values are multiplied by this (1 / keep probability), dropped ones are 0 */
  float *out_cm;              /* optional channel-major copy [B, cout, cols_per_cloud] of this layer's output */
} cpfn_mlp_layer_t;

typedef struct {
  int32_t n_layers;
  cpfn_mlp_layer_t layers[CPFN_MLP_MAX_LAYERS];
  const void *weights;        /* device: the layers' cpfn_mlp_pack_weights_host images, concatenated */
  size_t weight_bytes;
  int32_t tile_cols;          /* 128, 64 or 32 columns per CTA tile */
  int32_t in_mode;            /* CPFN_MLP_IN_* */
  int32_t B, cols_per_cloud;  /* columns = B * cols_per_cloud (S*K grouped samples, or N points) */
  const float *a_src; int32_t a_ch; int32_t a_rows;   /* point-major [B, a_rows, a_ch] (DENSE / INTERP: rows = columns) */
  const int32_t *idx;         /* GROUP: [B,S,K] ball-query result; INTERP: [B,N,3] three_nn indices */
  const float *xyz;           /* GROUP: [B, a_rows, 3] */
  const float *centers;       /* GROUP: [B, S, 3] */
  int32_t group_k;            /* GROUP: K (columns per centre) */
  const float *b_src; int32_t b_ch; int32_t b_rows;   /* INTERP: [B, b_rows, b_ch] */
  const float *nn_w;          /* INTERP: [B,N,3] interpolation weights */
  int32_t out_mode;           /* CPFN_MLP_OUT_* */
  float *out; int32_t ldo; int32_t pool_g;
  int32_t split_cout;         /* 1: single-layer chain, tile_cols 64/32: one CTA per (column tile, 128-channel chunk) */
  /* GROUP mode with a_ch == 0 (set abstraction on bare positions): optional fp32 first layer
   * relu(l0_w [l0_cout,3] * (xyz[idx] - centre) + l0_b) evaluated on the CUDA cores while the tile is
   * built; layers[0].cin must then equal l0_cout. */
  const float *l0_w; const float *l0_b; int32_t l0_cout;
  /* INTERP mode: when non-NULL, relu(interpolated + in_bias[b_ch]) is what enters layer 0.  A layer in front
   * of the interpolation is linear in the rows it interpolates, W (sum_j w_j f_j) = sum_j w_j (W f_j): the caller
   * applies that layer's W once per SOURCE row and its bias + ReLU here, per interpolated row
   * (pointset_feature_propagation.py:36-51 for FP3, whose input is the interpolation alone). */
  const float *in_bias;
  /* OUT_POOL with atomic merging needs the pooled output zero-filled; non-zero = the caller has already done
   * that (e.g. off the critical path, on another stream), the library skips its own cudaMemsetAsync. */
  int32_t out_prezeroed;
  /* Column window: when win_cols > 0 the launch processes columns [win_off, win_off + win_cols) of EVERY cloud only
   * (both multiples of tile_cols and, for OUT_POOL, of pool_g; OUT_POOL then needs out_prezeroed).  All other fields
   * keep describing the whole arrays.  Lets a set-abstraction MLP start on the centroids already sampled. */
  int32_t win_cols, win_off;
  /* > 0: upper bound on the number of (persistent) CTAs of the launch, for a chain that shares the GPU with another
   * kernel which must keep its SMs (the CTAs loop over the tiles, so any number works). */
  int32_t max_ctas;
  /* GROUP mode with features (a_ch > 0), tile_cols 128: layer 0's three position columns kept OUT of the tensor-core
   * operand.  xyz_w = [128 * ceil(cout0 / 128)][4] fp32 rows (wx, wy, wz, bias) of layer 0; layers[0].cin is then a_ch
   * (no position columns, no K atom of padding for them) and layer 0's epilogue adds
   * wx * dx + wy * dy + wz * dz + bias in fp32, (dx, dy, dz) = xyz[idx] - centre as in pointset_abstraction.py:62-63.
   * layers[0].bias is ignored.  NULL: the position columns are the last three input channels of layer 0. */
  const float *xyz_w;
} cpfn_mlp_chain_t;

/* Replaces the conv+BN+ReLU(+max) chains of pointset_abstraction.py:61-74,
 * pointset_feature_propagation.py:36-51 and pn2_network.py:60-68 (see csrc/mlp_chain.cu). */
CPFN_API size_t cpfn_mlp_packed_bytes(int cout, int cin);
CPFN_API int cpfn_mlp_pack_weights_host(const float *W_host, int cout, int cin, void *packed_host);
CPFN_API int cpfn_mlp_chain(const cpfn_mlp_chain_t *chain, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Small fused helpers of the inference forward.
 * ------------------------------------------------------------------------- */

/* three_nn followed by the inverse-distance weights of the FP layer: replaces
 * three_nn + sqrt (modules/geometry_utils.py:184) + 1/(d+1e-8) normalised
 * (pointset_feature_propagation.py:38-42).  weight, idx: [B,n,3]. */
CPFN_API int cpfn_three_nn_weights(const float *unknown, const float *known, int B, int n, int m,
                                   float *weight, int32_t *idx, cpfn_stream_t stream);
/* The same for queries given as (x, y, z, index bits) float4 records in a spatially sorted order: the first B * n
 * records of a cpfn_ball_query_grid_build workspace of the SAME cloud (cell order).  Thread j answers record j and
 * writes weight / idx at the record's own index, so the result equals cpfn_three_nn_weights on the original order
 * while a warp's queries share their candidate cells.  384 <= m <= 2048 (the shared-memory grid of the known
 * cloud); other sizes: CPFN_EINVAL (use cpfn_three_nn_weights). */
CPFN_API int cpfn_three_nn_weights_sorted(const void *sorted_queries, const float *known, int B, int n, int m,
                                          float *weight, int32_t *idx, cpfn_stream_t stream);

/* new_xyz[b,s,:] = xyz[b, idx[b,s], :] (select_point_subset on positions,
 * pointset_abstraction.py:50).  xyz [B,N,3], idx [B,S] -> out [B,S,3]. */
CPFN_API int cpfn_gather_xyz(const float *xyz, const int32_t *idx, int B, int N, int S, float *out,
                             cpfn_stream_t stream);

/* out[r, co] = bias[co] + sum_k W[co,k] x[r,k] in fp32 (few rows): the per-cloud constant part of
 * the first FP layer when pos2 is None (pointset_feature_propagation.py:33-34: feats2 repeated over
 * all points contributes the same vector to every point of a cloud).  Columns cout .. ldo-1 of every
 * output row are set to zero (the chains read biases zero padded to a multiple of 128). */
CPFN_API int cpfn_linear_rows(const float *x, const float *W, const float *bias, int rows, int cin,
                              int cout, int ldo, float *out, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Segmentation glue of the losses / metrics.
 * ------------------------------------------------------------------------- */

/* One pass over the memberships: S[b,g,k] = sum over the points with label g of W[b,n,k] (g < G), colsum[b,k] = sum over
 * ALL points, count[b,g] = number of points with label g, n_gt[b] = max label + 1 -- everything
 * SPFN/losses_implementation.py:11-30 (hungarian_matching) and :77-89 (compute_miou_loss) need from W and I_gt
 * ([B,N] int32 or int64, -1 = no label).  K, G <= 64.  Deterministic (no floating-point atomics). */
CPFN_API size_t cpfn_seg_workspace_bytes(int B, int N, int K, int G);
CPFN_API int cpfn_label_membership_sums(const float *W, const void *labels, int labels_are_int64, int B, int N, int K,
                                        int G, float *S, float *colsum, float *count, int32_t *n_gt, void *workspace,
                                        size_t workspace_bytes, cpfn_stream_t stream);

/* hungarian_matching on the device: IoU cost S / clamp(count + colsum - S, 1e-10) for the n_gt[b] ground-truth rows,
 * maximum-weight assignment with scipy.optimize.linear_sum_assignment's algorithm and tie-breaking (one warp per
 * sample, K <= 32).  matching int64 [B,K] (entries >= n_gt[b] are 0, as in the reference), mask u8 [B,K] | NULL
 * (metric_implementation's second result). */
CPFN_API int cpfn_hungarian_matching(const float *S, const float *colsum, const float *count, const int32_t *n_gt, int B,
                                     int K, int G, long long *matching, unsigned char *mask, cpfn_stream_t stream);

/* Stream-ordered zero fill (cudaMemsetAsync): the pooled output of a set-abstraction chain that runs as several
 * column windows is cleared once by the caller (cpfn_mlp_chain_t.out_prezeroed). */
CPFN_API int cpfn_zero_fill(void *dst, size_t bytes, cpfn_stream_t stream);

/* LocalSPFN input normalisation (Dataset/dataloaders.py:249-253): out[b, i] = (P[idx[b, i]] - mean_b) / max_i |P[idx[b, i]]
 * - mean_b|, mean_b = the patch's mean point (fp64 accumulation in a fixed order, rounded to fp32: the result of a
 * patch does not depend on the number of patches in the call).  points f32 [Ng,3], patch_idx int32 or int64 [nb,Np]
 * -> out f32 [nb,Np,3]. */
CPFN_API int cpfn_normalise_patches(const float *points, long long Ng, const void *patch_idx, int idx_is_int64,
                                    int nb, int Np, float *out, cpfn_stream_t stream);

/* Dropout keep-mask of F.dropout(x, p) for a channel-major x [B, C, N] (pn2_network.py:63, always on), as ONE BIT
 * per element instead of the fp32 mask tensor: bits[(b*N + n) * ceil(C/32) + c/32] bit c%32 = keep.  The random
 * stream is torch's own for that call: Philox4x32-10 keyed by `seed`, thread t of torch's launch (`torch_threads`
 * threads in all) owns the 4 consecutive elements 4*(t + k*torch_threads) .. +3 in its k-th iteration and draws
 * them from counter (offset/4 + k, subsequence t); an element is kept iff uint32 * 2^-32 + 2^-33 < keep_prob
 * (curand_uniform).  `rng_state` is a DEVICE pair {seed, offset} (uint64 each) so that a captured CUDA graph
 * draws a fresh mask at every replay: cpfn_rng_set (a one-thread kernel) writes it, stream-ordered. */
CPFN_API int cpfn_rng_set(unsigned long long *rng_state, unsigned long long seed, unsigned long long offset,
                          cpfn_stream_t stream);
CPFN_API int cpfn_dropout_mask_bits(const unsigned long long *rng_state, int B, int C, int N, float keep_prob,
                                    long long torch_threads, uint32_t *bits, cpfn_stream_t stream);

/* X = normalize(heads[:, x_off:x_off+3]), W = softmax(heads[:, w_off:w_off+K])
 * (Utils/training_utils.py:141-142); optionally inst = argmax_k W (hard_W_encoding's argmax) and
 * type = argmax of the n_types type logits at t_off (what evaluation_globalSPFN.py:97-110 saves).
 * heads [rows, ld] -> X [rows,3], W [rows,K], inst / type int32 [rows] (NULL to skip); K <= 64. */
CPFN_API int cpfn_spfn_post(const float *heads, long long rows, int ld, int x_off, int t_off, int n_types,
                            int w_off, int K, float *X, float *W, int32_t *inst, int32_t *type,
                            cpfn_stream_t stream);

/* cpfn_spfn_post for the patch-sharded cascade: additionally writes the type logits as T_out
 * [rows, n_types] (NULL to skip), and sends the X / W / T_out rows of local patch j (rows_per_patch rows each, a
 * multiple of 256) to the slab of patch j * patch_stride + patch_offset of the destination arrays.  X, W and T_out
 * may be PEER-MAPPED device pointers (the receive buffers of the rank that runs the merge, reached over NVLink): the
 * kernel that produces the per-point outputs is then also the exchange step -- no separate all-gather.  inst / type
 * stay in local row order.  rows_per_patch = 0: plain cpfn_spfn_post. */
CPFN_API int cpfn_spfn_post_scatter(const float *heads, long long rows, int ld, int x_off, int t_off, int n_types,
                                    int w_off, int K, float *X, float *W, float *T_out, int32_t *inst, int32_t *type,
                                    int rows_per_patch, int patch_stride, int patch_offset, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Patch extraction: the k nearest high-resolution points of every seed.
 * Replaces, per seed, the numpy block of Utils/sampling_utils.py:9-13 and
 * Preprocessing/preprocessing_sampling_patch.py:36-40:
 *     distances = np.linalg.norm(seed - gt_points_hr, axis=1)
 *     patch_indices = np.argsort(distances)[:k]; patch_distances = np.sort(distances)[:k]
 * hr_xyz [N,3], seeds_xyz [S,3] (device) -> out_idx int32 [S,k] ordered by (distance, index)
 * -- the stable argsort order; numpy's default sort leaves ties unordered --, out_dist f32 [S,k]
 * (np.sort(distances)[:k], bit-identical; NULL to skip), out_radius f32 [S] = out_dist[:,k-1]
 * (np.max(patch_distances), the pool-pruning radius of :16; NULL to skip).  1 <= k <= min(N, 16384).
 * Exact radix select, no full sort; all S seeds share the launches.
 * ------------------------------------------------------------------------- */
CPFN_API size_t cpfn_extract_patches_workspace_bytes(int N, int S, int k);
CPFN_API int cpfn_extract_patches(const float *hr_xyz, int N, const float *seeds_xyz, int S, int k,
                                  int32_t *out_idx, float *out_dist, float *out_radius, void *workspace,
                                  size_t workspace_bytes, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Patch -> object merging: Utils/merging_utils.py and the fusion block of
 * evaluation_localSPFN.py:99-130, without the dense [N_global, nb*Kl+Kg] matrix.
 * W [nb,Np,Kl] patch memberships, patch_idx int32 [nb,Np] global index of every patch point (unique inside
 * a patch), S [Ng,Kg] object-level labels as float, M = nb*Kl + Kg.
 * ------------------------------------------------------------------------- */

/* inv int32 [nb,Ng]: inv[b,p] = position of global point p in patch b, or -1. */
CPFN_API size_t cpfn_merge_inverse_bytes(int nb, int Ng);
CPFN_API int cpfn_merge_inverse_index(const int32_t *patch_idx, int nb, int Np, int Ng, int32_t *inv,
                                      cpfn_stream_t stream);

/* similarity_soft (merging_utils.py:6-15): out [M,M] = A^T A, fp64 accumulation; Kl, Kg <= 32. */
CPFN_API size_t cpfn_merge_similarity_workspace_bytes(int nb, int Kl, int Kg);
CPFN_API int cpfn_merge_similarity(const float *W, const int32_t *patch_idx, const float *S, const int32_t *inv,
                                   int nb, int Np, int Kl, int Ng, int Kg, float *out, void *workspace,
                                   size_t workspace_bytes, cpfn_stream_t stream);

/* heuristic_merging (merging_utils.py:17-33), HOST arrays: pairs int64 [P,2], penalty f64 [P],
 * patch_id int64 [n_nodes] -> segment_id int64 [n_nodes].  Same result as the reference's
 * repeated-argmax loop in one pass over the stably sorted pairs. */
CPFN_API int cpfn_heuristic_merging_host(const int64_t *pairs, const double *penalty, int64_t n_pairs,
                                         const int64_t *patch_id, int64_t n_nodes, int64_t *segment_id);

/* The pair list of run_heuristic_solver (merging_utils.py:37-39: similarity > threshold, i < j, np.where order)
 * taken straight from the HOST matrix similarity f64 [n_nodes,n_nodes], then the same greedy merge. */
CPFN_API int cpfn_merge_solve_host(const double *similarity, int64_t n_nodes, double threshold,
                                   const int64_t *patch_id, int64_t *segment_id);
/* The same for the float32 matrix similarity_soft returns (what the reference passes): no conversion, 32-bit sort keys. */
CPFN_API int cpfn_merge_solve_host_f32(const float *similarity, int64_t n_nodes, float threshold,
                                       const int64_t *patch_id, int64_t *segment_id);

/* The same solve on the DEVICE, straight from the similarity matrix cpfn_merge_similarity left there (f32 [M,M],
 * M = nb*Kl + Kg <= 4096, nb + 1 <= 64 patches): pair keys -> radix sort -> one-CTA greedy pass -> label replacement
 * of the empty slots and np.unique (merging_utils.py:35-44 + :17-33) without the device->host copy of the matrix and
 * the host loop.  labels int32 [M] (consecutive, in np.unique's order), label_weight f32 [M] (entry l = 1 / (members
 * of label l + 1e-10), what get_point_final divides by; 0 beyond the last label), n_labels int32 [1] (device),
 * segments int32 [M] | NULL (heuristic_merging's raw segment ids, before the replacement). */
CPFN_API size_t cpfn_merge_solve_workspace_bytes(int nb, int Kl, int Kg);
CPFN_API int cpfn_merge_solve(const float *similarity, int nb, int Kl, int Kg, float threshold, int32_t *labels,
                              float *label_weight, int32_t *n_labels, int32_t *segments, void *workspace,
                              size_t workspace_bytes, cpfn_stream_t stream);

/* Fused evaluation_localSPFN.py:103-111 + get_point_final (merging_utils.py:46-50): labels int32 [M] in
 * [0,L), label_weight f32 [L] = 1/(members+1e-10); out [Ng,L].  Points inside a patch drop the object block. */
CPFN_API int cpfn_merge_point_labels(const float *W, const float *S, const int32_t *inv, const int32_t *labels,
                                     const float *label_weight, int nb, int Np, int Kl, int Ng, int Kg, int L,
                                     float *out, cpfn_stream_t stream);

/* get_point_final on a dense A [Ng,M] (the reference's own call signature). */
CPFN_API int cpfn_merge_dense_labels(const float *A, const int32_t *labels, const float *label_weight, int Ng, int M,
                                     int L, float *out, cpfn_stream_t stream);

/* Merged normals / types (evaluation_localSPFN.py:113-130): X [nb,Np,3], T [nb,Np,n_types] summed over the
 * patches containing a point (ascending patch order), uncovered points take obj_normals [Ng,3] /
 * obj_types [Ng,n_types]; normals re-normalised, types averaged.  n_types <= 8. */
CPFN_API int cpfn_merge_normals_types(const float *X, const float *T, const int32_t *inv, const float *obj_normals,
                                      const float *obj_types, int nb, int Np, int Ng, int n_types, float *out_normals,
                                      float *out_types, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Point-to-primitive residues (SURVEY 8a row a14 + the residue part of 8f row f3).
 * Parameter tensors as in the dictionary of compute_parameters: [B,Kp,3] / [B,Kp], contiguous; a NULL
 * pointer is allowed for a type that is not requested.  Class ids: 0 plane, 1 sphere, 2 cylinder, 3 cone.
 * ------------------------------------------------------------------------- */
typedef struct cpfn_primitive_params {
  const float *plane_normal, *plane_center;
  const float *sphere_center, *sphere_radius_squared;
  const float *cylinder_axis, *cylinder_center, *cylinder_radius_squared;
  const float *cone_apex, *cone_axis, *cone_half_angle;
} cpfn_primitive_params_t;

/* compute_residue_loss (SPFN/losses_implementation.py:351-387) with *_fitter.compute_residue_single:
 * matching int32 [B,K] selects the parameter slot of every primitive; points: the primitive's n_pts points at
 * points + b*stride_b + k*stride_k (floats; stride_k = 0 shares one cloud between all primitives, as
 * compute_P_coverage does); class_ids[T] = the requested types in output order (T <= 4).
 * per_point [B,K,n_pts,T] (NULL to skip), mean [B,K,T] = torch.mean(residue_per_point, dim=2) (NULL to skip). */
CPFN_API int cpfn_primitive_residues(const cpfn_primitive_params_t *params, const int32_t *matching, const float *points,
                                     long long stride_b, long long stride_k, int B, int Kp, int K, int n_pts,
                                     const int *class_ids, int T, float *per_point, float *mean, cpfn_stream_t stream);

/* compute_P_coverage (SPFN/metric_implementation.py:409-415) without the [B,K,N,T] intermediate:
 * count[b,e] = #points of P [B,N,3] whose smallest sqrt_safe(residue) over the K primitives (each with its own
 * class prim_class int32 [B,K]) is below epsilons_host[e] (HOST array, n_eps <= 4).  Coverage = count / N. */
CPFN_API int cpfn_p_coverage(const cpfn_primitive_params_t *params, const int32_t *matching, const int32_t *prim_class,
                             const float *P, int B, int Kp, int K, int N, const float *epsilons_host, int n_eps,
                             float *count, cpfn_stream_t stream);

/* ---------------------------------------------------------------------------
 * Farthest point sampling of ONE large cloud with the semantics of the reference's preprocessing
 * (Preprocessing/preprocessing_sampling_lowres.py:14-42) -- not those of the pointnet2 op:
 * true distances, first-index tie-break, seeds / labels.
 *   labels == NULL : furthest_point_sampling(points, seeds, n_out): the seeds (int32 [n_seeds], may be NULL) only
 *                    start with distance 0; out = the n_out successive arg-max indices.
 *   labels != NULL : furthest_point_sampling_per_label(points, labels): out[0] = start_index (the reference draws it
 *                    with np.random.randint), after every pick the points of the picked label drop out; n_out = the
 *                    number of distinct labels.
 * points f32 [N,3], N <= 8 * 1024 * SM count; cooperative launch over the whole GPU.
 * ------------------------------------------------------------------------- */
CPFN_API size_t cpfn_fps_dense_workspace_bytes(void);
CPFN_API int cpfn_fps_dense(const float *points, int N, const int32_t *labels, const int32_t *seeds, int n_seeds,
                            int start_index, int n_out, int32_t *out, void *workspace, size_t workspace_bytes,
                            cpfn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CPFN_B200_H_ */
