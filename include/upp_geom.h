/*
 * upp_geom.h -- C ABI of libupp_geom.so: the B200 (sm_100a) implementation of the UPP
 * point-geometry hot path (farthest-point sampling, kNN patch grouping, Chamfer fwd/bwd).
 *
 * This is the drop-in boundary.  Every entry point
 *   - is extern "C", takes raw DEVICE pointers + sizes + an explicit stream
 *     (upp_stream_t == cudaStream_t), no torch types;
 *   - never allocates, never synchronises the host, never keeps state between calls
 *     (safe to capture in a CUDA graph); scratch, where needed, is passed in;
 *   - returns 0 (UPP_OK), a negative UPP_ERR_* for argument errors, or the positive
 *     cudaError_t the launch produced.  Nothing is printed, nothing calls exit()
 *     (the reference prints and carries on: extensions/chamfer_dist/chamfer.cu:166-169,224-227);
 *   - all tensors are dense row-major ("contiguous") fp32 unless stated.
 *
 * Each prototype cites the reference interface it replaces (paths relative to the
 * reference repo root, zhoujiahuan1991/ICCV2025-UPP).  INTEGRATION.md shows the Python
 * (ctypes) binding a maintainer adds on the reference side.
 */
#ifndef UPP_GEOM_H_
#define UPP_GEOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* upp_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define UPP_API __attribute__((visibility("default")))
#else
#define UPP_API
#endif

#define UPP_OK 0
#define UPP_ERR_INVALID_ARG (-1) /* null pointer, negative size, k > N, ... */
#define UPP_ERR_UNSUPPORTED (-2) /* shape outside what the kernels cover (see each call) */
#define UPP_ERR_WORKSPACE (-3)   /* workspace too small / missing */

#define UPP_VERSION 100 /* 0.1.0 */

/* Library version (UPP_VERSION of the build). */
UPP_API int upp_version(void);

/* Human-readable text for a return code of any function below (static storage). */
UPP_API const char* upp_error_string(int rc);

/* ---------------------------------------------------------------------------------------
 * Farthest point sampling.
 * Replaces pointnet2_ops.pointnet2_utils.furthest_point_sample(xyz, npoint)
 *   call sites: utils/misc.py:18, tools/runner_module.py:151,450 (third-party CUDA op;
 *   upstream pointnet2_ops/_ext-src/src/sampling_gpu.cu furthest_point_sampling_kernel).
 * xyz (B,N,3) f32 -> idx_out (B,M) int32.  First sample is index 0; points with
 * x^2+y^2+z^2 <= 1e-3 are never selected (upstream quirk, kept); ties resolve to the LOWEST
 * point index (BASELINE.json north_star; upstream resolves exact ties thread-major).
 * M > N is allowed (as upstream): once every point has been taken the arg-max keeps
 * returning already-selected points.
 * centers_out: optional (nullable) (B,M,3) f32 receiving xyz[b, idx[b,j], :] -- fuses the
 *   gather_operation + two transposes of utils/misc.py:19.
 * workspace: only needed when upp_fps_workspace_bytes(B,N,M) > 0 (N beyond the
 *   register-resident limit); may be NULL otherwise.
 */
UPP_API size_t upp_fps_workspace_bytes(int B, int N, int M);
UPP_API int upp_fps_f32(const float* xyz, int B, int N, int M, int32_t* idx_out, float* centers_out,
                void* workspace, size_t workspace_bytes, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Channel-first gather and its gradient.
 * Replaces pointnet2_ops.pointnet2_utils.gather_operation(features, idx) fwd / bwd
 *   call sites: utils/misc.py:19, tools/runner_module.py:153,455
 *   (upstream sampling_gpu.cu gather_points_kernel / gather_points_grad_kernel).
 * features (B,C,N) f32, idx (B,M) int32 in [0,N) -> out (B,C,M): out[b,c,j] = features[b,c,idx[b,j]].
 * Gradient: grad_features (B,C,N) is OVERWRITTEN with the scatter-add of grad_out (B,C,M)
 *   (the callee zero-fills; duplicate indices accumulate).
 */
UPP_API int upp_gather_f32(const float* features, const int32_t* idx, int B, int C, int N, int M,
                   float* out, upp_stream_t stream);
UPP_API int upp_gather_grad_f32(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                        float* grad_features, upp_stream_t stream);

/* Gradient of the row-major coordinate gather fused into upp_fps_f32 (centers_out):
 *   grad (B,N,C) is OVERWRITTEN with the scatter-add of grad_rows (B,M,C) by idx (B,M) int32.
 * Equals gather_operation's backward (upstream gather_points_grad_kernel) applied to the transposed
 * tensors of utils/misc.py:19, without the two transpose copies. */
UPP_API int upp_rows_scatter_add_f32(const float* grad_rows, const int32_t* idx, int B, int N, int M, int C,
                             float* grad, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Exact brute-force k nearest neighbours.
 * Replaces knn_cuda.KNN(k, transpose_mode=True).forward(ref, query)
 *   constructed models/Point_MAE_unify.py:56, called :69 (third-party KNN_CUDA 0.2,
 *   knn.cu cuComputeDistanceGlobal / cuInsertionSort / cuParallelSqrt, one cloud per
 *   Python-loop iteration; here one launch for the batch).
 * ref (B,N,3), query (B,Q,3) -> dist_out (B,Q,k) f32 EUCLIDEAN (sqrt applied), ascending;
 * idx_out (B,Q,k) int64, 0-based; equal distances keep the lower ref index first.
 * dist_out may be NULL (Group discards it).  Requires 1 <= k <= N (upstream reads out of
 * bounds when k > N; rejected here with UPP_ERR_INVALID_ARG).
 */
UPP_API int upp_knn_f32(const float* ref, const float* query, int B, int N, int Q, int k,
                float* dist_out, int64_t* idx_out, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Chamfer distance forward.
 * Replaces chamfer.forward(xyz1, xyz2) = chamfer_cuda_forward
 *   extensions/chamfer_dist/chamfer_cuda.cpp:36-39, chamfer.cu:147-171 (kernel :15-145);
 *   consumed by extensions/chamfer_dist/__init__.py:16.
 * xyz1 (B,N,3), xyz2 (B,M,3) -> dist1 (B,N), dist2 (B,M) f32 SQUARED nearest distances and
 * idx1 (B,N), idx2 (B,M) int32 arg-mins (lowest index on ties).  N == 0 or M == 0 yields
 * zero-filled outputs, as the reference's torch::zeros does.
 * partial_sums: optional (nullable) 4 floats, OVERWRITTEN with
 *   { sum dist1, sum dist2, sum sqrt(dist1), sum sqrt(dist2) } over the whole call -- the
 *   send buffer of the one NCCL all-reduce the batch-sharded loss needs (utils/dist_utils.py:41-48).
 * workspace: optional scratch of upp_chamfer_fwd_workspace_bytes(B,N,M) bytes (contents irrelevant
 *   on entry, clobbered on exit).  With it the single-pass kernel runs (each pairwise distance
 *   computed once for both directions); with NULL the two-direction kernel runs.  Results are
 *   identical either way.
 */
UPP_API size_t upp_chamfer_fwd_workspace_bytes(int B, int N, int M);
UPP_API int upp_chamfer_fwd_f32(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                        float* dist2, int32_t* idx1, int32_t* idx2, float* partial_sums,
                        void* workspace, size_t workspace_bytes, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Chamfer forward of a batch SHARDED across the GPUs of one node, fused with the all-reduce of its sums.
 * Replaces, for the batch-sharded loss, chamfer.forward + torch.mean + dist_utils.reduce_tensor
 *   (extensions/chamfer_dist/__init__.py:28-84, utils/dist_utils.py:41-48, tools/runner_pretask.py:241).
 * Same outputs as upp_chamfer_fwd_f32 for this rank's B clouds, but global_sums (4 floats) receives
 *   { sum dist1, sum dist2, sum sqrt(dist1), sum sqrt(dist2) } over ALL ranks' clouds: the kernel that finishes
 * the local sums stores them into every peer's exchange buffer over NVLink (CUDA IPC-mapped peer memory), waits
 * for the peers' contributions in its own buffer and adds them in rank order -- bit-identical on every rank, no
 * NCCL launch, no host synchronisation, CUDA-graph capturable.  All ranks must issue their calls in the same order
 * (as with any collective).  Each rank provides:
 *   slots[r]  device pointer (valid in THIS process) to rank r's exchange buffer of upp_peer_exchange_bytes(world)
 *             bytes, zero-filled once before the first call;
 *   seq       this rank's call counter in device memory, zero before the first call.
 * Requires the workspace (single-pass path) and max(N, M) >= 128; otherwise UPP_ERR_UNSUPPORTED (use
 * upp_chamfer_fwd_f32 + an NCCL all-reduce of its partial sums).  B == 0 (this rank's shard is empty) is valid: the rank
 * contributes zeros to the exchange and receives the global sums like everybody else.
 */
#define UPP_MAX_PEERS 16
typedef struct upp_peer_exchange {
  float* slots[UPP_MAX_PEERS];
  int rank;
  int world;
  unsigned int* seq;
  int defer; /* 0: the call returns the global sums; 1: the call only SENDS (global_sums receives this rank's local
                sums) and upp_peer_allreduce_finish_f32, launched later on the same stream, waits and adds */
  unsigned int* status; /* nullable: host-visible (pinned, mapped) word, zero before the first call.  A wait for a peer
                that runs out stores the failed call's sequence number here (system scope) besides poisoning the sums with
                NaN, so the host can raise instead of training on NaN silently. */
  long long timeout_cycles; /* SM clocks a wait for a peer may last; 0 = default 2^38 (about 2.3 minutes at 1.965 GHz:
                rank skew of seconds -- checkpointing, evaluation, a data-loader stall -- is normal and is waited out) */
} upp_peer_exchange;
#define upp_peer_exchange_bytes(world) ((size_t)2 * (size_t)(world) * 8 * sizeof(float))
UPP_API int upp_chamfer_fwd_sharded_f32(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                                float* dist2, int32_t* idx1, int32_t* idx2, float* global_sums,
                                void* workspace, size_t workspace_bytes, const upp_peer_exchange* peers,
                                upp_stream_t stream);

/* The exchange on its own: SUM all-reduce of 4 floats over the mapped peer buffers, rank order, bit-identical on every
 * rank; one tiny launch, honours peers->defer like the sharded forward.  local4 == NULL contributes zeros: this is what a
 * rank whose shard is EMPTY issues in place of upp_chamfer_fwd_sharded_f32 (which does it itself for B == 0), so that the
 * ranks' call sequences stay aligned.  Mirrors utils/dist_utils.py:41-48 for a 16-byte payload. */
UPP_API int upp_peer_allreduce_f32(const upp_peer_exchange* peers, const float* local4, float* global4,
                                   upp_stream_t stream);

/* Second half of a deferred exchange (peers->defer was 1 in the last upp_chamfer_fwd_sharded_f32 call on this
 * stream): waits for every rank's contribution of that call and writes the rank-ordered sum to global_sums (4 floats).
 * Launch it where the loss VALUE is consumed (end of the step): a kernel that waits on its peers early in the step
 * stalls the work queued behind it on the same hardware queue; by the end of the step the peers have delivered.
 * Exactly one finish per deferred call, before the next sharded call. */
UPP_API int upp_peer_allreduce_finish_f32(const upp_peer_exchange* peers, float* global_sums, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Chamfer distance backward.
 * Replaces chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2) = chamfer_cuda_backward
 *   extensions/chamfer_dist/chamfer_cuda.cpp:36-39, chamfer.cu:203-229 (kernel :173-201);
 *   consumed by extensions/chamfer_dist/__init__.py:24.
 * grad_xyz1 (B,N,3), grad_xyz2 (B,M,3) are OVERWRITTEN with
 *   grad_xyz1[b,j] = 2 g1[b,j] (x1_j - x2_idx1[j]) - sum_{k: idx2[k]==j} 2 g2[b,k] (x2_k - x1_j)
 * and symmetrically for grad_xyz2; inf*0 = NaN appears exactly where the reference produces it.
 * grad_dist1 (B,N), grad_dist2 (B,M) must be dense (the binding makes them so; the reference
 * silently assumes it, chamfer.cu:217).
 * One launch, no memset, no float atomics: each output row is produced by one thread that adds its partner terms in
 * ascending point order, so the gradients are bit-identical run to run (the reference's atomicAdd order is not).
 */
UPP_API int upp_chamfer_bwd_f32(const float* xyz1, const float* xyz2, const int32_t* idx1,
                        const int32_t* idx2, const float* grad_dist1, const float* grad_dist2,
                        int B, int N, int M, float* grad_xyz1, float* grad_xyz2,
                        upp_stream_t stream);

/* Chamfer backward + gradient statistics (BASELINE.json north_star: "all-reduce the scalar Chamfer loss and its gradient
 * statistics"; the quantity tools/runner_module.py:204 clip_grad_norm_ needs of the coordinate gradients).
 * Same gradients as upp_chamfer_bwd_f32 (same kernel), plus sqnorm_out[0..1] = { sum ||grad_xyz1||^2, sum ||grad_xyz2||^2 }
 * over the call, summed in a fixed order (deterministic).  With peers != NULL (batch sharded over the GPUs of one node)
 * the two sums are all-reduced over NVLink peer memory inside the same kernel, exactly like the forward's loss sums:
 * every rank receives the sums over the GLOBAL batch, bit-identical; all ranks must call in the same order.
 * workspace: upp_chamfer_bwd_stats_workspace_bytes(B,N,M) bytes of scratch (contents irrelevant on entry). */
UPP_API size_t upp_chamfer_bwd_stats_workspace_bytes(int B, int N, int M);
UPP_API int upp_chamfer_bwd_stats_f32(const float* xyz1, const float* xyz2, const int32_t* idx1,
                              const int32_t* idx2, const float* grad_dist1, const float* grad_dist2,
                              int B, int N, int M, float* grad_xyz1, float* grad_xyz2, float* sqnorm_out,
                              void* workspace, size_t workspace_bytes, const upp_peer_exchange* peers,
                              upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused Group divider (opt-in fast path; the unchanged call site composes the ops above).
 * Replaces the body of Group.forward  models/Point_MAE_unify.py:58-92:
 *   misc.fps -> KNN -> index gather -> subtract centre.
 * xyz (B,N,3) -> neighborhood (B,G,k,3) (already centre-subtracted), center (B,G,3),
 * idx (B,G,k) int64 LOCAL ref indices (the gather_idx=True convention; the binding adds b*N
 * for the flat convention), center_idx (B,G) int32.  idx / center_idx may be NULL.
 * Same workspace rule as upp_fps_f32 (query with upp_fps_workspace_bytes(B,N,G)).
 */
UPP_API int upp_group_f32(const float* xyz, int B, int N, int G, int k, float* neighborhood,
                  float* center, int64_t* idx, int32_t* center_idx, void* workspace,
                  size_t workspace_bytes, upp_stream_t stream);

/* Backward of the fused Group w.r.t. xyz:
 *   grad_xyz[b, idx[b,g,j]] += grad_nb[b,g,j];  grad_xyz[b, center_idx[b,g]] += grad_center[b,g] - sum_j grad_nb[b,g,j]
 * grad_xyz (B,N,3) is OVERWRITTEN.  grad_center may be NULL (treated as zero). */
UPP_API int upp_group_bwd_f32(const float* grad_nb, const float* grad_center, const int64_t* idx,
                      const int32_t* center_idx, int B, int N, int G, int k, float* grad_xyz,
                      upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K nearest neighbours in the pytorch3d convention.
 * Replaces pytorch3d.ops.knn_points(p1, p2, K=K, return_nn=True) (third-party, un-pinned; call site
 *   models/Point_MAE_pretask_dev.py:680: K = 4 clean neighbours of every noise point).
 * p1 (B,N1,3) queries, p2 (B,N2,3) references -> dist2_out (B,N1,K) f32 SQUARED distances ascending
 * (nullable), idx_out (B,N1,K) int64 into p2, nn_out (B,N1,K,3) = p2[idx] (nullable).  Equal distances
 * keep the lower index first.  d = fma(dz,dz,fma(dy,dy,dx*dx)).  Requires 1 <= K <= min(N2, 32).
 */
UPP_API int upp_knn_points_f32(const float* p1, const float* p2, int B, int N1, int N2, int K,
                       float* dist2_out, int64_t* idx_out, float* nn_out, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Viewpoint crop of a batch of clouds.
 * Replaces the per-cloud body of misc.seprate_point_cloud  utils/misc.py:232-239
 *   (torch.norm(center - points) -> torch.argsort -> idx[:num_crop] / idx[num_crop:] gathers), called once per training
 *   step from tools/runner_module.py:131, tools/runner_pretask.py:179, tools/runner_unify_seg.py:212.
 * xyz (B,n,3), viewpoints (B,3) -> crop_out (B,num_crop,3): the num_crop points nearest to the cloud's viewpoint, nearest
 * first; input_out: the remaining points in ascending distance (B,n-num_crop,3), or -- padding_zeros != 0 -- the cloud in
 * its original order with the cropped rows multiplied by zero (B,n,3).  order_out (nullable, (B,n) int32): the full
 * ascending order.  Equal distances keep the lower point index first (a stable sort).  Distance: sqrt_rn of
 * fma(dz,dz,fma(dy,dy,dx*dx)), d* = viewpoint - point.  One launch for the batch; n <= 8192 (UPP_ERR_UNSUPPORTED beyond).
 */
UPP_API int upp_crop_split_f32(const float* xyz, const float* viewpoints, int B, int n, int num_crop, int padding_zeros,
                       float* crop_out, float* input_out, int32_t* order_out, upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * k-nearest inverse-distance feature interpolation.
 * Replaces the pure-torch body of
 *   propagate(xyz1, xyz2, points1, points2, de_neighbors, dist_e)   models/Point_MAE_unify.py:22-48
 *   PointNetFeaturePropagation.forward (interpolation part)         models/Point_MAE_unify_segment.py:289-313,
 *                                                                   models/Point_MAE_pretask_dev.py:437-461
 * i.e. square_distance (models/modules.py:13-32, expanded form) -> full sort -> first k ->
 * 1/(d+eps) weights normalised -> index_points gather -> weighted sum.
 * xyz1 (B,N,3) targets, xyz2 (B,S,3) sources, feat2 (B,S,C) channel-last source features ->
 *   out (B,N,C) = (base ? base : 0) + alpha * sum_j weight_j * feat2[idx_j]
 *   (propagate: base = points1, alpha = 0.3, eps = 1e-8; feature propagation: base NULL, alpha 1, eps 1e-4)
 * idx (B,N,k) int32, weight (B,N,k), dist (B,N,k, nullable): the selection, saved for backward;
 * neighbours ascending by (distance, source index).  Requires 1 <= k <= min(S, 32).
 */
UPP_API int upp_interp_fwd_f32(const float* xyz1, const float* xyz2, const float* feat2, const float* base,
                       float alpha, float eps, int B, int N, int S, int C, int k, float* out,
                       int32_t* idx, float* weight, float* dist, upp_stream_t stream);

/* The two halves of upp_interp_fwd_f32 as separate calls, so that a caller can keep a selection and pay for it once:
 * the SA-units of the Rectification Prompter call propagate six times per forward on IDENTICAL geometry
 * (models/Point_MAE_pretask_dev.py:298), and every PointNetFeaturePropagation call of one forward shares its xyz pair.
 *   upp_interp_select_f32  xyz1, xyz2 -> idx (B,N,k) int32, weight (B,N,k), dist (B,N,k, nullable): selection + weights only
 *   upp_interp_blend_f32   out (B,N,C) = (base ? base : 0) + alpha * sum_j weight_j * feat2[idx_j] from a saved selection
 * select followed by blend produces bit-identical output to upp_interp_fwd_f32 (same kernels / same arithmetic order);
 * upp_interp_bwd_f32 takes the same saved selection. */
UPP_API int upp_interp_select_f32(const float* xyz1, const float* xyz2, float eps, int B, int N, int S, int k,
                          int32_t* idx, float* weight, float* dist, upp_stream_t stream);
UPP_API int upp_interp_blend_f32(const float* feat2, const float* base, float alpha, const int32_t* idx,
                         const float* weight, int B, int N, int S, int C, int k, float* out, upp_stream_t stream);

/* Backward of the interpolation (deterministic, no atomics).
 * grad_feat2 (B,S,C) is OVERWRITTEN with alpha * sum_{(n,j): idx = s} weight * grad_out[b,n,:].
 * Coordinate gradients (through the weights) are produced when gd_workspace (B*N*k floats) is given:
 * grad_xyz1 (B,N,3) and grad_xyz2 (B,S,3) are then OVERWRITTEN (either may be NULL) and dist, feat2,
 * xyz1, xyz2 must be the forward's; with gd_workspace == NULL those five pointers are ignored.
 * (The gradient w.r.t. base is grad_out itself.)
 * workspace (nullable): upp_interp_bwd_workspace_bytes(B,N,S,C,k) bytes of 16-byte-aligned scratch; when given
 * (and the size query is non-zero: C a multiple of 128, S <= 128, k <= 8) large problems take the streamed
 * feature-gradient kernel, which reads grad_out once instead of k times.  Results are the same either way. */
UPP_API size_t upp_interp_bwd_workspace_bytes(int B, int N, int S, int C, int k);
UPP_API int upp_interp_bwd_f32(const float* grad_out, const int32_t* idx, const float* weight, const float* dist,
                       const float* feat2, const float* xyz1, const float* xyz2, float alpha, float eps,
                       int B, int N, int S, int C, int k, float* grad_feat2, float* grad_xyz1,
                       float* grad_xyz2, float* gd_workspace, void* workspace, size_t workspace_bytes,
                       upp_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Introspection used by bench.py / tests: number of kernel launches the library has issued
 * since load (monotonic, process-wide, relaxed atomic). */
UPP_API unsigned long long upp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* UPP_GEOM_H_ */
