// gather.cu -- channel-first gather (+grad) and the Group divider's neighbourhood gather (+grad).
//
// gather_points / gather_points_grad replace pointnet2_ops gather_operation (reference call site
// utils/misc.py:19).  group_gather fuses Group.forward's index gather + centre subtraction
// (models/Point_MAE_unify.py:72-88); group_bwd is its scatter-add gradient.
// All HBM-bound copies: one thread per output element, coalesced on the output side.
#include "common.cuh"

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

__global__ void __launch_bounds__(256)
    gather_points_grad_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx, int C,
                              int N, int M, size_t total, float* __restrict__ gfeat) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % M);
    const size_t bc = i / M;
    const size_t b = bc / C;
    atomicAdd(gfeat + bc * N + __ldg(idx + b * M + j), __ldg(gout + i));
  }
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
// One warp per group: the k neighbour gradients are summed by shuffle for the centre term.
__global__ void __launch_bounds__(256)
    group_bwd_kernel(const float* __restrict__ gnb, const float* __restrict__ gcenter,
                     const int64_t* __restrict__ idx, const int32_t* __restrict__ cidx, int N, int G,
                     int k, size_t groups, float* __restrict__ gxyz) {
  const int lane = threadIdx.x & 31;
  for (size_t bg = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
       bg < groups; bg += static_cast<size_t>(gridDim.x) * (blockDim.x >> 5)) {
    const size_t b = bg / G;
    float* gb = gxyz + b * N * 3;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int j = lane; j < k; j += 32) {
      const size_t i = bg * k + j;
      const float vx = gnb[3 * i], vy = gnb[3 * i + 1], vz = gnb[3 * i + 2];
      float* dst = gb + static_cast<size_t>(idx[i]) * 3;
      atomicAdd(dst, vx); atomicAdd(dst + 1, vy); atomicAdd(dst + 2, vz);
      sx += vx; sy += vy; sz += vz;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) {
      float cx = -sx, cy = -sy, cz = -sz;
      if (gcenter) { cx += gcenter[3 * bg]; cy += gcenter[3 * bg + 1]; cz += gcenter[3 * bg + 2]; }
      float* dst = gb + static_cast<size_t>(cidx[bg]) * 3;
      atomicAdd(dst, cx); atomicAdd(dst + 1, cy); atomicAdd(dst + 2, cz);
    }
  }
}

// grad[b, idx[b,j], :] += rows[b,j,:]  (row-major / channel-last; the gradient of the coordinate gather fused
// into upp_fps_f32's centers_out, i.e. of utils/misc.py:19 without its two transposes)
__global__ void __launch_bounds__(256)
    rows_scatter_add_kernel(const float* __restrict__ rows, const int32_t* __restrict__ idx, int N, int M, int C,
                            size_t total, float* __restrict__ grad) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t bj = i / C;  // b*M + j
    const size_t b = bj / M;
    atomicAdd(grad + (b * N + __ldg(idx + bj)) * C + c, __ldg(rows + i));
  }
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
  const size_t total = static_cast<size_t>(B) * C * M;
  cudaError_t e = cudaMemsetAsync(gfeat, 0, static_cast<size_t>(B) * C * N * sizeof(float), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (total == 0) return UPP_OK;
  gather_points_grad_kernel<<<grid_for(total, 256), 256, 0, st>>>(gout, idx, C, N, M, total, gfeat);
  count_launch();
  return launch_status();
}

int rows_scatter_add_launch(const float* rows, const int32_t* idx, int B, int N, int M, int C, float* grad,
                            cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(grad, 0, static_cast<size_t>(B) * N * C * sizeof(float), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t total = static_cast<size_t>(B) * M * C;
  if (total == 0) return UPP_OK;
  rows_scatter_add_kernel<<<grid_for(total, 256), 256, 0, st>>>(rows, idx, N, M, C, total, grad);
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
  cudaError_t e = cudaMemsetAsync(gxyz, 0, static_cast<size_t>(B) * N * 3 * sizeof(float), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t groups = static_cast<size_t>(B) * G;
  if (groups == 0) return UPP_OK;
  group_bwd_kernel<<<grid_for(groups, 8), 256, 0, st>>>(gnb, gcenter, idx, cidx, N, G, k, groups, gxyz);
  count_launch();
  return launch_status();
}

}  // namespace upp
