// gather.cu -- channel-first gather (+grad) and the Group divider's neighbourhood gather (+grad).
//
// gather_points / gather_points_grad replace pointnet2_ops gather_operation (reference call site
// utils/misc.py:19).  group_gather fuses Group.forward's index gather + centre subtraction
// (models/Point_MAE_unify.py:72-88); group_bwd is its scatter-add gradient.
// Forward gathers: one thread per output element, coalesced on the output side.  Backward scatter-adds: ordered
// gathers (scatter.cuh) -- deterministic, no float atomics, no memset.
#include "scatter.cuh"

namespace upp {

__global__ void __launch_bounds__(256)
    gather_points_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, int C, int N,
                         int M, size_t total, float* __restrict__ out) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % M);
    const size_t bc = i / M;  // b*C + c
    const size_t b = bc / C;
    out[i] = __ldg(feat + bc * N + __ldg(idx + b * M + j));
  }
}

// Gradient of the channel-first gather (upstream gather_points_grad_kernel: float atomicAdd): ordered gather, see
// scatter.cuh.  Destinations are the N points of a cloud, entries the M sampled indices; blockIdx.z walks the channels
// three at a time.
struct GatherGradOp {
  const float* gout;    // (B,C,M)
  const int32_t* idx;   // (B,M)
  float* gfeat;         // (B,C,N)
  int b, c0, C, N, M;
  __device__ __forceinline__ int entries() const { return M; }
  __device__ __forceinline__ int dst(int e) const { return __ldg(idx + static_cast<size_t>(b) * M + e); }
  __device__ __forceinline__ void fetch(int e, int, float (&v)[3]) const {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) v[c] = __ldg(gout + (static_cast<size_t>(b) * C + c0 + c) * M + e);
  }
  __device__ __forceinline__ void init(int, float (&)[3]) const {}
  __device__ __forceinline__ void store(int j, const float (&a)[3]) const {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) gfeat[(static_cast<size_t>(b) * C + c0 + c) * N + j] = a[c];
  }
};

template <int Q>
__global__ void __launch_bounds__(256)
    gather_points_grad_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx, int C, int N, int M,
                              float* __restrict__ gfeat) {
  extern __shared__ int s_dst[];
  GatherGradOp op{gout, idx, gfeat, static_cast<int>(blockIdx.y), static_cast<int>(blockIdx.z) * 3, C, N, M};
  ordered_scatter_cta<Q>(op, N, s_dst);
}

// neighborhood[b,g,j,:] = xyz[b, idx[b,g,j], :] - center[b,g,:]   (one thread per neighbour)
__global__ void __launch_bounds__(256)
    group_gather_kernel(const float* __restrict__ xyz, const float* __restrict__ center,
                        const int64_t* __restrict__ idx, int N, int G, int k, size_t total,
                        float* __restrict__ nb) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t bg = i / k;  // b*G + g
    const size_t b = bg / G;
    const float* p = xyz + (b * N + static_cast<size_t>(idx[i])) * 3;
    const float* c = center + bg * 3;
    nb[3 * i + 0] = __ldg(p + 0) - __ldg(c + 0);
    nb[3 * i + 1] = __ldg(p + 1) - __ldg(c + 1);
    nb[3 * i + 2] = __ldg(p + 2) - __ldg(c + 2);
  }
}

// grad_xyz[b, idx[b,g,j]] += gnb[b,g,j]; grad_xyz[b, cidx[b,g]] += gcenter[b,g] - sum_j gnb[b,g,j]
// Ordered gather (scatter.cuh): per cloud the entry list is the G*k neighbour slots in (g, j) order followed by the G
// centre slots in g order; a centre entry's value is formed by the lane that holds it (k sequential adds in j order).
struct GroupBwdOp {
  const float* gnb;       // (B,G,k,3)
  const float* gcenter;   // (B,G,3) or null
  const int64_t* idx;     // (B,G,k)
  const int32_t* cidx;    // (B,G)
  float* gxyz;            // (B,N,3)
  int b, N, G, k;
  __device__ __forceinline__ int entries() const { return G * k + G; }
  __device__ __forceinline__ int dst(int e) const {
    const int nk = G * k;
    return e < nk ? static_cast<int>(idx[static_cast<size_t>(b) * nk + e]) : __ldg(cidx + static_cast<size_t>(b) * G + (e - nk));
  }
  __device__ __forceinline__ void fetch(int e, int, float (&v)[3]) const {
    const int nk = G * k;
    if (e < nk) {
      const float* p = gnb + (static_cast<size_t>(b) * nk + e) * 3;
      v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2);
    } else {
      const int g = e - nk;
      const float* p = gnb + (static_cast<size_t>(b) * G + g) * k * 3;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int j0 = 0; j0 < k; j0 += 16) {  // 48 independent loads in flight, then the adds in j order
        float t[16][3];
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
          for (int c = 0; c < 3; ++c) t[u][c] = j0 + u < k ? __ldg(p + 3 * (j0 + u) + c) : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (j0 + u < k) { sx = __fadd_rn(sx, t[u][0]); sy = __fadd_rn(sy, t[u][1]); sz = __fadd_rn(sz, t[u][2]); }
      }
      v[0] = -sx; v[1] = -sy; v[2] = -sz;
      if (gcenter) {
        const float* c = gcenter + (static_cast<size_t>(b) * G + g) * 3;
        v[0] = __fadd_rn(v[0], __ldg(c)); v[1] = __fadd_rn(v[1], __ldg(c + 1)); v[2] = __fadd_rn(v[2], __ldg(c + 2));
      }
    }
  }
  __device__ __forceinline__ void init(int, float (&)[3]) const {}
  __device__ __forceinline__ void store(int j, const float (&a)[3]) const {
    float* o = gxyz + (static_cast<size_t>(b) * N + j) * 3;
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  }
};

template <int Q>
__global__ void __launch_bounds__(256)
    group_bwd_kernel(const float* __restrict__ gnb, const float* __restrict__ gcenter,
                     const int64_t* __restrict__ idx, const int32_t* __restrict__ cidx, int N, int G,
                     int k, float* __restrict__ gxyz) {
  extern __shared__ int s_dst[];
  GroupBwdOp op{gnb, gcenter, idx, cidx, gxyz, static_cast<int>(blockIdx.y), N, G, k};
  ordered_scatter_cta<Q>(op, N, s_dst);
}

// grad[b, idx[b,j], :] += rows[b,j,:]  (row-major / channel-last; the gradient of the coordinate gather fused
// into upp_fps_f32's centers_out, i.e. of utils/misc.py:19 without its two transposes).  Ordered gather, channels
// three at a time (blockIdx.z).
struct RowsScatterOp {
  const float* rows;     // (B,M,C)
  const int32_t* idx;    // (B,M)
  float* grad;           // (B,N,C)
  int b, c0, C, N, M;
  __device__ __forceinline__ int entries() const { return M; }
  __device__ __forceinline__ int dst(int e) const { return __ldg(idx + static_cast<size_t>(b) * M + e); }
  __device__ __forceinline__ void fetch(int e, int, float (&v)[3]) const {
    const float* p = rows + (static_cast<size_t>(b) * M + e) * C + c0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) v[c] = __ldg(p + c);
  }
  __device__ __forceinline__ void init(int, float (&)[3]) const {}
  __device__ __forceinline__ void store(int j, const float (&a)[3]) const {
    float* o = grad + (static_cast<size_t>(b) * N + j) * C + c0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c0 + c < C) o[c] = a[c];
  }
};

template <int Q>
__global__ void __launch_bounds__(256)
    rows_scatter_add_kernel(const float* __restrict__ rows, const int32_t* __restrict__ idx, int N, int M, int C,
                            float* __restrict__ grad) {
  extern __shared__ int s_dst[];
  RowsScatterOp op{rows, idx, grad, static_cast<int>(blockIdx.y), static_cast<int>(blockIdx.z) * 3, C, N, M};
  ordered_scatter_cta<Q>(op, N, s_dst);
}

static int grid_for(size_t total, int per_block) {
  const size_t want = (total + per_block - 1) / per_block;
  const size_t cap = 148 * 16;
  return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

int gather_launch(const float* feat, const int32_t* idx, int B, int C, int N, int M, float* out,
                  cudaStream_t st) {
  const size_t total = static_cast<size_t>(B) * C * M;
  if (total == 0) return UPP_OK;
  gather_points_kernel<<<grid_for(total, 256), 256, 0, st>>>(feat, idx, C, N, M, total, out);
  count_launch();
  return launch_status();
}

int gather_grad_launch(const float* gout, const int32_t* idx, int B, int C, int N, int M,
                       float* gfeat, cudaStream_t st) {
  if (static_cast<size_t>(B) * C * N == 0) return UPP_OK;
  // every destination is written exactly once (zero where nothing lands): no memset, no atomics
  const int q = scatter_pick_q(B * ((C + 2) / 3), N);
  const ScatterGrid g = scatter_grid(N, M, q);
  cudaError_t le = q == 1 ? launch_pdl(gather_points_grad_kernel<1>, dim3(g.blocks, B, (C + 2) / 3), dim3(g.warps * 32), g.smem, st, gout, idx, C, N, M, gfeat)
                          : launch_pdl(gather_points_grad_kernel<2>, dim3(g.blocks, B, (C + 2) / 3), dim3(g.warps * 32), g.smem, st, gout, idx, C, N, M, gfeat);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return launch_status();
}

int rows_scatter_add_launch(const float* rows, const int32_t* idx, int B, int N, int M, int C, float* grad,
                            cudaStream_t st) {
  if (static_cast<size_t>(B) * N * C == 0) return UPP_OK;
  const int q = scatter_pick_q(B * ((C + 2) / 3), N);
  const ScatterGrid g = scatter_grid(N, M, q);
  cudaError_t le = q == 1 ? launch_pdl(rows_scatter_add_kernel<1>, dim3(g.blocks, B, (C + 2) / 3), dim3(g.warps * 32), g.smem, st, rows, idx, N, M, C, grad)
                          : launch_pdl(rows_scatter_add_kernel<2>, dim3(g.blocks, B, (C + 2) / 3), dim3(g.warps * 32), g.smem, st, rows, idx, N, M, C, grad);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return launch_status();
}

int group_gather_launch(const float* xyz, const float* center, const int64_t* idx, int B, int N,
                        int G, int k, float* nb, cudaStream_t st) {
  const size_t total = static_cast<size_t>(B) * G * k;
  if (total == 0) return UPP_OK;
  group_gather_kernel<<<grid_for(total, 256), 256, 0, st>>>(xyz, center, idx, N, G, k, total, nb);
  count_launch();
  return launch_status();
}

int group_bwd_launch(const float* gnb, const float* gcenter, const int64_t* idx, const int32_t* cidx,
                     int B, int N, int G, int k, float* gxyz, cudaStream_t st) {
  if (static_cast<size_t>(B) * N == 0) return UPP_OK;
  const int q = scatter_pick_q(B, N);
  const ScatterGrid g = scatter_grid(N, G * k + G, q);
  cudaError_t le = q == 1 ? launch_pdl(group_bwd_kernel<1>, dim3(g.blocks, B), dim3(g.warps * 32), g.smem, st, gnb, gcenter, idx, cidx, N, G, k, gxyz)
                          : launch_pdl(group_bwd_kernel<2>, dim3(g.blocks, B), dim3(g.warps * 32), g.smem, st, gnb, gcenter, idx, cidx, N, G, k, gxyz);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return launch_status();
}

}  // namespace upp
