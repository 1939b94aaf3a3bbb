// chamfer.cu -- Chamfer distance forward / backward for sm_100a.
//
// Replaces extensions/chamfer_dist/chamfer.cu: chamfer_dist_kernel (:15-145, launched twice with a
// fixed <<<(32,16),512>>> grid, one query per thread, one LDS per 1.3 pairs) and
// chamfer_dist_grad_kernel (:173-201, <<<(1,16),256>>> = 16 CTAs with the batch loop serial
// inside and six float atomics per point).
//
// Forward here: ONE launch covers both directions (blockIdx.z) and the whole batch (blockIdx.y).
// The reference cloud is staged in 2048-point tiles by TMA bulk copy; each thread owns R queries
// in registers and reads the tile as broadcast LDS.128 (3 loads feed 4 refs x R queries), so the
// inner loop is FP32-pipe bound: 3 FADD + FMUL + 2 FFMA + compare/select per pair.
// d = fma(dz,dz,fma(dx,dx,dy*dy)), d* = ref - query; strict '<' in ascending ref order gives the
// lowest index on ties (== reference).  The grid is sized from the problem, not fixed.
//
// Backward here: pass 1 WRITES each point's own term (no atomics, no memset), pass 2 scatters the
// partner term with float RED.ADD -- half the reference's atomics, all B*(N+M) points in flight.
#include <math.h>

#include "common.cuh"

namespace upp {

constexpr int kChTile = 2048;  // reference points per shared-memory tile (24 KB)

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
    chamfer_fwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int M,
                       float* __restrict__ dist1, float* __restrict__ dist2,
                       int32_t* __restrict__ idx1, int32_t* __restrict__ idx2) {
  __shared__ __align__(16) float s_ref[kChTile * 3];
  __shared__ __align__(8) uint64_t s_bar;
  const int dir = blockIdx.z;
  const int b = blockIdx.y;
  const int nq = dir == 0 ? N : M;  // queries
  const int nr = dir == 0 ? M : N;  // references
  const int q0 = blockIdx.x * (THREADS * R);
  if (q0 >= nq) return;  // whole CTA leaves together (before any barrier)
  const float* qp = (dir == 0 ? xyz1 : xyz2) + static_cast<size_t>(b) * nq * 3;
  const float* rp = (dir == 0 ? xyz2 : xyz1) + static_cast<size_t>(b) * nr * 3;
  float* dout = (dir == 0 ? dist1 : dist2) + static_cast<size_t>(b) * nq;
  int32_t* iout = (dir == 0 ? idx1 : idx2) + static_cast<size_t>(b) * nq;
  const int t = threadIdx.x;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;

  float qx[R], qy[R], qz[R], best[R];
  int bi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = q0 + r * THREADS + t;
    const int jj = j < nq ? j : nq - 1;  // clamp: computes a duplicate, never stored
    qx[r] = __ldg(qp + 3 * jj);
    qy[r] = __ldg(qp + 3 * jj + 1);
    qz[r] = __ldg(qp + 3 * jj + 2);
    best[r] = __int_as_float(0x7f800000);
    bi[r] = 0;
  }

  for (int base = 0; base < nr; base += kChTile) {
    const int tile = min(kChTile, nr - base);
    const int tile4 = (tile + 3) & ~3;
    if (base > 0) __syncthreads();
    // pad the last group of four with NaN: (NaN - q)^2 = NaN never compares '<'
    for (int i = tile * 3 + t; i < tile4 * 3; i += THREADS) s_ref[i] = __int_as_float(0x7fc00000);
    stage_points(s_ref, rp + static_cast<size_t>(base) * 3, tile, &s_bar, parity);

    const float4* s4 = reinterpret_cast<const float4*>(s_ref);
#pragma unroll 2
    for (int k = 0; k < tile4; k += 4) {
      const float4 a = s4[(k >> 2) * 3 + 0];  // x0 y0 z0 x1
      const float4 c = s4[(k >> 2) * 3 + 1];  // y1 z1 x2 y2
      const float4 e = s4[(k >> 2) * 3 + 2];  // z2 x3 y3 z3
      const int kk = base + k;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float d;
        d = dist_yxz(a.x - qx[r], a.y - qy[r], a.z - qz[r]);
        if (d < best[r]) { best[r] = d; bi[r] = kk; }
        d = dist_yxz(a.w - qx[r], c.x - qy[r], c.y - qz[r]);
        if (d < best[r]) { best[r] = d; bi[r] = kk + 1; }
        d = dist_yxz(c.z - qx[r], c.w - qy[r], e.x - qz[r]);
        if (d < best[r]) { best[r] = d; bi[r] = kk + 2; }
        d = dist_yxz(e.y - qx[r], e.z - qy[r], e.w - qz[r]);
        if (d < best[r]) { best[r] = d; bi[r] = kk + 3; }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = q0 + r * THREADS + t;
    if (j < nq) {
      dout[j] = best[r];
      iout[j] = bi[r];
    }
  }
}

// Deterministic whole-call sums { sum d1, sum d2, sum sqrt d1, sum sqrt d2 }: one CTA, fixed
// strided order + fixed tree, so the value fed to the NCCL all-reduce is run-to-run identical.
__global__ void __launch_bounds__(1024)
    chamfer_sums_kernel(const float* __restrict__ dist1, size_t n1, const float* __restrict__ dist2,
                        size_t n2, float* __restrict__ sums) {
  __shared__ float s_part[4][32];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (size_t i = threadIdx.x; i < n1; i += 1024) {
    const float d = dist1[i];
    a0 += d;
    a2 += __fsqrt_rn(d);
  }
  for (size_t i = threadIdx.x; i < n2; i += 1024) {
    const float d = dist2[i];
    a1 += d;
    a3 += __fsqrt_rn(d);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_part[0][warp] = a0; s_part[1][warp] = a1; s_part[2][warp] = a2; s_part[3][warp] = a3; }
  __syncthreads();
  if (warp < 4) {
    const float v = warp_sum(s_part[warp][lane]);
    if (lane == 0) sums[warp] = v;
  }
}

// Backward pass 1: own term, plain stores.  grad_a[j] = 2 g[j] (a_j - b_idx[j]) for both clouds.
__global__ void __launch_bounds__(256)
    chamfer_bwd_own_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                           const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                           const float* __restrict__ g1, const float* __restrict__ g2, int B, int N,
                           int M, float* __restrict__ gx1, float* __restrict__ gx2) {
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const bool first = i < n1;
    const size_t p = first ? i : i - n1;
    const int na = first ? N : M, nb = first ? M : N;
    const size_t b = p / na;
    const float* a = (first ? xyz1 : xyz2) + p * 3;
    const int j2 = (first ? idx1 : idx2)[p];
    const float* o = (first ? xyz2 : xyz1) + (b * nb + j2) * 3;
    const float g = __fmul_rn((first ? g1 : g2)[p], 2.0f);
    float* out = (first ? gx1 : gx2) + p * 3;
    out[0] = __fmul_rn(g, a[0] - o[0]);
    out[1] = __fmul_rn(g, a[1] - o[1]);
    out[2] = __fmul_rn(g, a[2] - o[2]);
  }
}

// Backward pass 2: partner term, scattered with RED.ADD.F32.  grad_b[idx[j]] += -(2 g[j] (a_j - b_idx[j])).
__global__ void __launch_bounds__(256)
    chamfer_bwd_scatter_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                               const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                               const float* __restrict__ g1, const float* __restrict__ g2, int B,
                               int N, int M, float* __restrict__ gx1, float* __restrict__ gx2) {
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const bool first = i < n1;
    const size_t p = first ? i : i - n1;
    const int na = first ? N : M, nb = first ? M : N;
    const size_t b = p / na;
    const float* a = (first ? xyz1 : xyz2) + p * 3;
    const int j2 = (first ? idx1 : idx2)[p];
    const size_t op = (b * nb + j2) * 3;
    const float* o = (first ? xyz2 : xyz1) + op;
    const float g = __fmul_rn((first ? g1 : g2)[p], 2.0f);
    float* out = (first ? gx2 : gx1) + op;
    atomicAdd(out + 0, -__fmul_rn(g, a[0] - o[0]));
    atomicAdd(out + 1, -__fmul_rn(g, a[1] - o[1]));
    atomicAdd(out + 2, -__fmul_rn(g, a[2] - o[2]));
  }
}

int chamfer_fwd_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* sums, cudaStream_t st) {
  constexpr int R = 4, THREADS = 128;
  const int per = R * THREADS;
  const int tiles = (max(N, M) + per - 1) / per;
  dim3 grid(tiles, B, 2);
  chamfer_fwd_kernel<R, THREADS><<<grid, THREADS, 0, st>>>(xyz1, xyz2, N, M, dist1, dist2, idx1, idx2);
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK || sums == nullptr) return rc;
  chamfer_sums_kernel<<<1, 1024, 0, st>>>(dist1, static_cast<size_t>(B) * N, dist2,
                                          static_cast<size_t>(B) * M, sums);
  count_launch();
  return launch_status();
}

int chamfer_bwd_launch(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                       const float* g1, const float* g2, int B, int N, int M, float* gx1, float* gx2,
                       cudaStream_t st) {
  const size_t total = static_cast<size_t>(B) * (static_cast<size_t>(N) + M);
  const size_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want > 148 * 16 ? 148 * 16 : want);
  chamfer_bwd_own_kernel<<<blocks, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, g1, g2, B, N, M, gx1, gx2);
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  chamfer_bwd_scatter_kernel<<<blocks, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, g1, g2, B, N, M, gx1, gx2);
  count_launch();
  return launch_status();
}

}  // namespace upp
