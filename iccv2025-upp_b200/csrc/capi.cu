// capi.cu -- the extern "C" surface of libupp_geom.so (include/upp_geom.h): argument checks and
// dispatch only.  No allocation, no host synchronisation, no state besides the launch counter.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace upp {

static std::atomic<unsigned long long> g_launches{0};
__global__ void tmap_probe_kernel() {}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: the library keeps linking against
// cudart only (no libcuda at build time, and none needed on the CPU-only build box).
int make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<encode_fn>(fn);
  }();
  if (encode == nullptr) return UPP_ERR_UNSUPPORTED;
  const cuuint64_t gdim[2] = {inner, rows};
  const cuuint64_t gstride[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = CUDA_ERROR_UNKNOWN;
  for (int attempt = 0; attempt < 2 && r != CUDA_SUCCESS; ++attempt) {
    if (attempt == 1) {
      // A thread that has made no runtime call needing a context yet (autograd's worker thread when the caching
      // allocator served every allocation from its pool) has no current context for the driver call above: observed
      // as ONE fallback launch on the first backward of bench.py's module-level step.  A function-attribute query
      // binds the primary context to the thread and is legal during stream capture.
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, tmap_probe_kernel) != cudaSuccess) break;
    }
    r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? UPP_OK : UPP_ERR_INVALID_ARG;
}

const char* tuning_env(const char* name) {
  const char* on = getenv("UPP_TUNING");
  if (on == nullptr || on[0] != '1' || on[1] != '\0') return nullptr;
  return getenv(name);
}
int tuning_env_int(const char* name, int dflt) {
  const char* s = tuning_env(name);
  return s ? atoi(s) : dflt;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }

int fps_launch(const float*, int, int, int, int32_t*, float*, void*, size_t, cudaStream_t);
size_t fps_workspace_bytes(int, int);
int knn_launch(const float*, const float*, int, int, int, int, float*, int64_t*, float*, cudaStream_t);
int chamfer_fwd_launch(const float*, const float*, int, int, int, float*, float*, int32_t*, int32_t*,
                       float*, void*, size_t, const upp_peer_exchange*, cudaStream_t);
size_t chamfer_fwd_workspace_bytes(int, int, int);
int peer_finish_launch(const upp_peer_exchange*, float*, cudaStream_t);
int chamfer_bwd_launch(const float*, const float*, const int32_t*, const int32_t*, const float*,
                       const float*, int, int, int, float*, float*, float*, void*, size_t, const upp_peer_exchange*,
                       cudaStream_t);
size_t chamfer_bwd_stats_workspace_bytes(int, int, int);
int peer_allreduce_launch(const upp_peer_exchange*, const float*, float*, cudaStream_t);
int gather_launch(const float*, const int32_t*, int, int, int, int, float*, cudaStream_t);
int gather_grad_launch(const float*, const int32_t*, int, int, int, int, float*, cudaStream_t);
int group_gather_launch(const float*, const float*, const int64_t*, int, int, int, int, float*,
                        cudaStream_t);
int rows_scatter_add_launch(const float*, const int32_t*, int, int, int, int, float*, cudaStream_t);
int group_bwd_launch(const float*, const float*, const int64_t*, const int32_t*, int, int, int, int,
                     float*, cudaStream_t);
int crop_split_launch(const float*, const float*, int, int, int, int, float*, float*, int32_t*, cudaStream_t);
int group_fused_launch(const float*, int, int, int, int, float*, float*, int64_t*, int32_t*, cudaStream_t);

int knn_points_launch(const float*, const float*, int, int, int, int, float*, int64_t*, float*, cudaStream_t);
int interp_fwd_launch(const float*, const float*, const float*, const float*, float, float, int, int, int, int,
                      int, float*, int32_t*, float*, float*, cudaStream_t);
int interp_blend_launch(const float*, const float*, float, const int32_t*, const float*, int, int, int, int, int, float*,
                        cudaStream_t);
int interp_bwd_launch(const float*, const int32_t*, const float*, const float*, const float*, const float*,
                      const float*, float, float, int, int, int, int, int, float*, float*, float*, float*,
                      void*, size_t, cudaStream_t);
size_t interp_bwd_workspace_bytes(int, int, int, int, int);

}  // namespace upp

using namespace upp;

#define UPP_REQUIRE(cond) \
  do {                    \
    if (!(cond)) return UPP_ERR_INVALID_ARG; \
  } while (0)

static inline cudaStream_t S(upp_stream_t s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int upp_version(void) { return UPP_VERSION; }

const char* upp_error_string(int rc) {
  switch (rc) {
    case UPP_OK: return "ok";
    case UPP_ERR_INVALID_ARG: return "upp_geom: invalid argument (null pointer, negative size, or k outside [1, N])";
    case UPP_ERR_UNSUPPORTED: return "upp_geom: shape not supported by the sm_100a kernels";
    case UPP_ERR_WORKSPACE: return "upp_geom: workspace missing or too small (see upp_fps_workspace_bytes)";
    default: return rc > 0 ? cudaGetErrorString(static_cast<cudaError_t>(rc)) : "upp_geom: unknown error";
  }
}

unsigned long long upp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t upp_fps_workspace_bytes(int B, int N, int M) {
  (void)M;
  if (B <= 0 || N <= 0) return 0;
  return fps_workspace_bytes(B, N);
}

int upp_fps_f32(const float* xyz, int B, int N, int M, int32_t* idx_out, float* centers_out,
                void* workspace, size_t workspace_bytes, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  if (B == 0 || M == 0) return UPP_OK;
  UPP_REQUIRE(N >= 1 && xyz != nullptr && idx_out != nullptr);
  return fps_launch(xyz, B, N, M, idx_out, centers_out, workspace, workspace_bytes, S(stream));
}

int upp_gather_f32(const float* features, const int32_t* idx, int B, int C, int N, int M, float* out,
                   upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0);
  if (static_cast<size_t>(B) * C * M == 0) return UPP_OK;
  UPP_REQUIRE(N >= 1 && features && idx && out);
  return gather_launch(features, idx, B, C, N, M, out, S(stream));
}

int upp_gather_grad_f32(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                        float* grad_features, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0);
  if (static_cast<size_t>(B) * C * N == 0) return UPP_OK;
  UPP_REQUIRE(grad_features != nullptr);
  UPP_REQUIRE(static_cast<size_t>(B) * C * M == 0 || (grad_out && idx));
  return gather_grad_launch(grad_out, idx, B, C, N, M, grad_features, S(stream));
}

int upp_rows_scatter_add_f32(const float* grad_rows, const int32_t* idx, int B, int N, int M, int C,
                             float* grad, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && C >= 0);
  if (static_cast<size_t>(B) * N * C == 0) return UPP_OK;
  UPP_REQUIRE(grad != nullptr);
  UPP_REQUIRE(static_cast<size_t>(B) * M * C == 0 || (grad_rows && idx));
  return rows_scatter_add_launch(grad_rows, idx, B, N, M, C, grad, S(stream));
}

int upp_knn_f32(const float* ref, const float* query, int B, int N, int Q, int k, float* dist_out,
                int64_t* idx_out, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && Q >= 0);
  UPP_REQUIRE(k >= 1 && k <= N);
  if (B == 0 || Q == 0) return UPP_OK;
  UPP_REQUIRE(ref && query && idx_out);
  return knn_launch(ref, query, B, N, Q, k, dist_out, idx_out, nullptr, S(stream));
}

size_t upp_chamfer_fwd_workspace_bytes(int B, int N, int M) { return chamfer_fwd_workspace_bytes(B, N, M); }

int upp_chamfer_fwd_f32(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                        float* dist2, int32_t* idx1, int32_t* idx2, float* partial_sums,
                        void* workspace, size_t workspace_bytes, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  UPP_REQUIRE(n1 == 0 || (dist1 && idx1));
  UPP_REQUIRE(n2 == 0 || (dist2 && idx2));
  if (N == 0 || M == 0 || B == 0) {  // reference: outputs stay torch::zeros (chamfer.cu:152-157)
    cudaError_t e = cudaSuccess;
    if (n1) { e = cudaMemsetAsync(dist1, 0, n1 * 4, S(stream)); if (e == cudaSuccess) e = cudaMemsetAsync(idx1, 0, n1 * 4, S(stream)); }
    if (n2 && e == cudaSuccess) { e = cudaMemsetAsync(dist2, 0, n2 * 4, S(stream)); if (e == cudaSuccess) e = cudaMemsetAsync(idx2, 0, n2 * 4, S(stream)); }
    if (partial_sums && e == cudaSuccess) e = cudaMemsetAsync(partial_sums, 0, 16, S(stream));
    return static_cast<int>(e);
  }
  UPP_REQUIRE(xyz1 && xyz2);
  return chamfer_fwd_launch(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, partial_sums, workspace,
                            workspace_bytes, nullptr, S(stream));
}

int upp_chamfer_fwd_sharded_f32(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                                float* dist2, int32_t* idx1, int32_t* idx2, float* global_sums, void* workspace,
                                size_t workspace_bytes, const upp_peer_exchange* peers, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 1 && M >= 1);
  UPP_REQUIRE(global_sums && peers);
  UPP_REQUIRE(peers->world >= 1 && peers->world <= UPP_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
  UPP_REQUIRE(peers->world == 1 || peers->seq != nullptr);
  for (int r = 0; r < peers->world && peers->world > 1; ++r) UPP_REQUIRE(peers->slots[r] != nullptr);
  if (B == 0) {  // an empty shard still takes part in the exchange (zeros), so the ranks' call sequences stay aligned
    if (peers->world == 1) return static_cast<int>(cudaMemsetAsync(global_sums, 0, 16, S(stream)));
    return peer_allreduce_launch(peers, nullptr, global_sums, S(stream));
  }
  UPP_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2);
  return chamfer_fwd_launch(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, global_sums, workspace, workspace_bytes,
                            peers, S(stream));
}

int upp_peer_allreduce_finish_f32(const upp_peer_exchange* peers, float* global_sums, upp_stream_t stream) {
  UPP_REQUIRE(peers && global_sums);
  UPP_REQUIRE(peers->world >= 2 && peers->world <= UPP_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
  UPP_REQUIRE(peers->seq != nullptr);
  for (int r = 0; r < peers->world; ++r) UPP_REQUIRE(peers->slots[r] != nullptr);
  return peer_finish_launch(peers, global_sums, S(stream));
}

int upp_chamfer_bwd_f32(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                        const float* grad_dist1, const float* grad_dist2, int B, int N, int M,
                        float* grad_xyz1, float* grad_xyz2, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && M >= 0);
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  UPP_REQUIRE(n1 == 0 || grad_xyz1);
  UPP_REQUIRE(n2 == 0 || grad_xyz2);
  if (N == 0 || M == 0 || B == 0) {
    cudaError_t e = cudaSuccess;
    if (n1) e = cudaMemsetAsync(grad_xyz1, 0, n1 * 12, S(stream));
    if (n2 && e == cudaSuccess) e = cudaMemsetAsync(grad_xyz2, 0, n2 * 12, S(stream));
    return static_cast<int>(e);
  }
  UPP_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && grad_dist1 && grad_dist2);
  return chamfer_bwd_launch(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2, B, N, M, grad_xyz1,
                            grad_xyz2, nullptr, nullptr, 0, nullptr, S(stream));
}

size_t upp_chamfer_bwd_stats_workspace_bytes(int B, int N, int M) { return chamfer_bwd_stats_workspace_bytes(B, N, M); }

int upp_chamfer_bwd_stats_f32(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                              const float* grad_dist1, const float* grad_dist2, int B, int N, int M,
                              float* grad_xyz1, float* grad_xyz2, float* sqnorm_out, void* workspace,
                              size_t workspace_bytes, const upp_peer_exchange* peers, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && sqnorm_out != nullptr);
  if (peers != nullptr) {
    UPP_REQUIRE(peers->world >= 1 && peers->world <= UPP_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
    UPP_REQUIRE(peers->world == 1 || peers->seq != nullptr);
    for (int r = 0; r < peers->world && peers->world > 1; ++r) UPP_REQUIRE(peers->slots[r] != nullptr);
  }
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  UPP_REQUIRE(n1 == 0 || grad_xyz1);
  UPP_REQUIRE(n2 == 0 || grad_xyz2);
  if (N == 0 || M == 0 || B == 0) {  // no pairs: zero gradients, zero statistics -- but still one exchange when sharded
    cudaError_t e = cudaSuccess;
    if (n1) e = cudaMemsetAsync(grad_xyz1, 0, n1 * 12, S(stream));
    if (n2 && e == cudaSuccess) e = cudaMemsetAsync(grad_xyz2, 0, n2 * 12, S(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
    if (peers != nullptr && peers->world > 1) return peer_allreduce_launch(peers, nullptr, sqnorm_out, S(stream));  // 4 floats
    return static_cast<int>(cudaMemsetAsync(sqnorm_out, 0, 8, S(stream)));
  }
  UPP_REQUIRE(xyz1 && xyz2 && idx1 && idx2 && grad_dist1 && grad_dist2);
  return chamfer_bwd_launch(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2, B, N, M, grad_xyz1, grad_xyz2, sqnorm_out,
                            workspace, workspace_bytes, peers, S(stream));
}

int upp_peer_allreduce_f32(const upp_peer_exchange* peers, const float* local4, float* global4, upp_stream_t stream) {
  UPP_REQUIRE(peers && global4);
  UPP_REQUIRE(peers->world >= 2 && peers->world <= UPP_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
  UPP_REQUIRE(peers->seq != nullptr);
  for (int r = 0; r < peers->world; ++r) UPP_REQUIRE(peers->slots[r] != nullptr);
  return peer_allreduce_launch(peers, local4, global4, S(stream));
}

int upp_group_f32(const float* xyz, int B, int N, int G, int k, float* neighborhood, float* center,
                  int64_t* idx, int32_t* center_idx, void* workspace, size_t workspace_bytes,
                  upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && G >= 0);
  UPP_REQUIRE(k >= 1 && k <= N);
  if (B == 0 || G == 0) return UPP_OK;
  UPP_REQUIRE(xyz && neighborhood && center && idx && center_idx);
  // one launch: the kNN of centre j runs behind the FPS chain that is still producing centres j+1... (group.cu)
  int rc = group_fused_launch(xyz, B, N, G, k, neighborhood, center, idx, center_idx, S(stream));
  if (rc == UPP_OK) return UPP_OK;
  if (rc > 0) (void)cudaGetLastError();  // a cluster that cannot be placed: compose the two launches below
  rc = fps_launch(xyz, B, N, G, center_idx, center, workspace, workspace_bytes, S(stream));
  if (rc != UPP_OK) return rc;
  // kNN with the fused epilogue: neighbourhood = xyz[idx] - center, no separate gather launch
  return knn_launch(xyz, center, B, N, G, k, nullptr, idx, neighborhood, S(stream));
}

int upp_group_bwd_f32(const float* grad_nb, const float* grad_center, const int64_t* idx,
                      const int32_t* center_idx, int B, int N, int G, int k, float* grad_xyz,
                      upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && G >= 0 && k >= 0);
  if (static_cast<size_t>(B) * N == 0) return UPP_OK;
  UPP_REQUIRE(grad_xyz != nullptr);
  UPP_REQUIRE(static_cast<size_t>(B) * G == 0 || (grad_nb && idx && center_idx));
  return group_bwd_launch(grad_nb, grad_center, idx, center_idx, B, N, G, k, grad_xyz, S(stream));
}

int upp_crop_split_f32(const float* xyz, const float* viewpoints, int B, int n, int num_crop, int padding_zeros,
                       float* crop_out, float* input_out, int32_t* order_out, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && n >= 0 && num_crop >= 0 && num_crop <= n);
  if (B == 0 || n == 0) return UPP_OK;
  UPP_REQUIRE(xyz && viewpoints);
  UPP_REQUIRE(num_crop == 0 || crop_out);
  UPP_REQUIRE((padding_zeros ? n : n - num_crop) == 0 || input_out);
  return crop_split_launch(xyz, viewpoints, B, n, num_crop, padding_zeros ? 1 : 0, crop_out, input_out, order_out, S(stream));
}

int upp_knn_points_f32(const float* p1, const float* p2, int B, int N1, int N2, int K, float* dist2_out,
                       int64_t* idx_out, float* nn_out, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N1 >= 0 && N2 >= 0);
  UPP_REQUIRE(K >= 1 && K <= N2 && K <= 32);
  if (B == 0 || N1 == 0) return UPP_OK;
  UPP_REQUIRE(p1 && p2 && idx_out);
  return knn_points_launch(p1, p2, B, N1, N2, K, dist2_out, idx_out, nn_out, S(stream));
}

int upp_interp_fwd_f32(const float* xyz1, const float* xyz2, const float* feat2, const float* base, float alpha,
                       float eps, int B, int N, int S, int C, int k, float* out, int32_t* idx, float* weight,
                       float* dist, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && S >= 0 && C >= 0);
  UPP_REQUIRE(k >= 1 && k <= S && k <= 32);
  if (B == 0 || N == 0) return UPP_OK;
  UPP_REQUIRE(xyz1 && xyz2 && idx && weight);
  UPP_REQUIRE(C == 0 || (feat2 && out));
  return interp_fwd_launch(xyz1, xyz2, feat2, base, alpha, eps, B, N, S, C, k, out, idx, weight, dist,
                           static_cast<cudaStream_t>(stream));  // (the size S shadows the helper here)
}

int upp_interp_select_f32(const float* xyz1, const float* xyz2, float eps, int B, int N, int S, int k, int32_t* idx,
                          float* weight, float* dist, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && S >= 0);
  UPP_REQUIRE(k >= 1 && k <= S && k <= 32);
  if (B == 0 || N == 0) return UPP_OK;
  UPP_REQUIRE(xyz1 && xyz2 && idx && weight);
  return interp_fwd_launch(xyz1, xyz2, nullptr, nullptr, 1.0f, eps, B, N, S, 0, k, nullptr, idx, weight, dist,
                           static_cast<cudaStream_t>(stream));
}

int upp_interp_blend_f32(const float* feat2, const float* base, float alpha, const int32_t* idx, const float* weight,
                         int B, int N, int S, int C, int k, float* out, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && S >= 1 && C >= 0);
  UPP_REQUIRE(k >= 1 && k <= S && k <= 32);
  if (B == 0 || N == 0 || C == 0) return UPP_OK;
  UPP_REQUIRE(feat2 && idx && weight && out);
  return interp_blend_launch(feat2, base, alpha, idx, weight, B, N, S, C, k, out, static_cast<cudaStream_t>(stream));
}

size_t upp_interp_bwd_workspace_bytes(int B, int N, int S, int C, int k) {
  return interp_bwd_workspace_bytes(B, N, S, C, k);
}

int upp_interp_bwd_f32(const float* grad_out, const int32_t* idx, const float* weight, const float* dist,
                       const float* feat2, const float* xyz1, const float* xyz2, float alpha, float eps, int B,
                       int N, int S, int C, int k, float* grad_feat2, float* grad_xyz1, float* grad_xyz2,
                       float* gd_workspace, void* workspace, size_t workspace_bytes, upp_stream_t stream) {
  UPP_REQUIRE(B >= 0 && N >= 0 && S >= 0 && C >= 0);
  UPP_REQUIRE(k >= 1 && k <= 32);
  if (B == 0 || S == 0) return UPP_OK;
  UPP_REQUIRE(grad_feat2 != nullptr || C == 0);
  UPP_REQUIRE(N == 0 || (grad_out && idx && weight));
  if (gd_workspace) UPP_REQUIRE(N == 0 || (dist && feat2 && xyz1 && xyz2));
  return interp_bwd_launch(grad_out, idx, weight, dist, feat2, xyz1, xyz2, alpha, eps, B, N, S, C, k, grad_feat2,
                           grad_xyz1, grad_xyz2, gd_workspace, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

}  // extern "C"
