// knn.cu -- exact brute-force kNN for sm_100a, one launch for the whole batch.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward (reference call site
// models/Point_MAE_unify.py:56,69).  Upstream (KNN_CUDA 0.2) runs a Python loop over the
// batch and, per cloud, materialises the N x Q distance matrix in global memory, then sorts
// each column with ONE THREAD per query (stride-Q global accesses), then a sqrt kernel.
//
// Here: the cloud's reference points are staged once per CTA into shared memory by a TMA bulk
// copy; one WARP owns one query.  Lanes stride over the references (conflict-free LDS), the
// running top-32 lives in registers as a lane-distributed sorted list (lane i holds the i-th
// best; the first k are the answer).  A chunk of 32 distances is filtered against the current
// k-th best with one ballot; each survivor is admitted with one shuffle-up (every lane decides
// from its own and its left neighbour's entry).  Candidates arrive in ascending index order, so
// "entries <= candidate stay in front" reproduces upstream's stable insertion (equal distances
// keep the lower index first).
// No distance matrix, no global scratch.  d = fma(dz,dz,fma(dy,dy,dx*dx)) with d* = ref-query,
// sorted on the squared value, sqrt (IEEE rn) applied to the k survivors, int64 0-based indices.
#include <float.h>

#include "common.cuh"

namespace upp {

constexpr int kKnnWarps = 8;           // queries in flight per CTA
constexpr int kKnnTile = 2048;         // reference points staged per pass (24 KB)

// k <= 32 : lane-distributed sorted list.
__global__ void __launch_bounds__(kKnnWarps * kWarp)
    knn_warp_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int Q,
                    int k, float* __restrict__ dist_out, int64_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) float s_ref[];  // min(N, kKnnTile) * 3 floats
  __shared__ __align__(8) uint64_t s_bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int q = blockIdx.x * kKnnWarps + warp;
  const bool active = q < Q;
  const float* rb = ref + static_cast<size_t>(b) * N * 3;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;

  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float* qp = query + (static_cast<size_t>(b) * Q + q) * 3;
    qx = __ldg(qp);
    qy = __ldg(qp + 1);
    qz = __ldg(qp + 2);
  }
  float ld = FLT_MAX;  // lane i: squared distance of the i-th best so far (+inf sentinel below)
  int li = -1;
  ld = __int_as_float(0x7f800000);
  float thr = ld;  // squared distance of the current k-th best (uniform across the warp)

  for (int base = 0; base < N; base += kKnnTile) {
    const int tile = min(kKnnTile, N - base);
    if (base > 0) __syncthreads();  // everyone done reading the previous tile
    stage_points(s_ref, rb + static_cast<size_t>(base) * 3, tile, &s_bar, parity);
    if (!active) continue;
    for (int c0 = 0; c0 < tile; c0 += kWarp) {
      const int c = c0 + lane;
      float d = __int_as_float(0x7f800000);
      if (c < tile) d = dist_xyz_acc(s_ref[3 * c] - qx, s_ref[3 * c + 1] - qy, s_ref[3 * c + 2] - qz);
      unsigned m = __ballot_sync(0xffffffffu, d < thr);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        // Admit candidate (cd, ci).  Candidates arrive in ascending index order, so it goes behind
        // every entry with distance <= cd.  Each lane decides from its own entry and its left
        // neighbour's (one shuffle-up, no ballot/popc on the dependent chain): entries > cd shift
        // right by one, the first of them is replaced by the candidate, lane 31's entry falls off.
        const float cd = __shfl_sync(0xffffffffu, d, src);
        const int ci = base + c0 + src;
        const float ud = __shfl_up_sync(0xffffffffu, ld, 1);
        const int ui = __shfl_up_sync(0xffffffffu, li, 1);
        const bool mine_gt = ld > cd;
        const bool left_gt = (lane > 0) && (ud > cd);
        li = mine_gt ? (left_gt ? ui : ci) : li;
        ld = mine_gt ? (left_gt ? ud : cd) : ld;
      }
      thr = __shfl_sync(0xffffffffu, ld, k - 1);
    }
  }
  if (active && lane < k) {
    const size_t o = (static_cast<size_t>(b) * Q + q) * k + lane;
    if (dist_out) dist_out[o] = __fsqrt_rn(ld);
    idx_out[o] = static_cast<int64_t>(li);
  }
}

// k > 32 : selection by repeated extraction (no storage): round r finds the smallest
// (distance, index) key strictly greater than the previous round's.  O(k*N/32) per query;
// the reference never needs it (k in {8,16,32}), it keeps the ABI total.
__global__ void __launch_bounds__(kKnnWarps * kWarp)
    knn_extract_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N,
                       int Q, int k, float* __restrict__ dist_out, int64_t* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int q = blockIdx.x * kKnnWarps + warp;
  if (q >= Q) return;
  const float* rb = ref + static_cast<size_t>(b) * N * 3;
  const float* qp = query + (static_cast<size_t>(b) * Q + q) * 3;
  const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
  unsigned long long prev = 0ull;
  bool have_prev = false;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = ~0ull;
    for (int c = lane; c < N; c += kWarp) {
      const float d = dist_xyz_acc(__ldg(rb + 3 * c) - qx, __ldg(rb + 3 * c + 1) - qy,
                                   __ldg(rb + 3 * c + 2) - qz);
      const unsigned long long key =
          (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned>(c);
      if ((!have_prev || key > prev) && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    prev = best;
    have_prev = true;
    if (lane == 0) {
      const size_t o = (static_cast<size_t>(b) * Q + q) * k + r;
      if (dist_out) dist_out[o] = __fsqrt_rn(__uint_as_float(static_cast<unsigned>(best >> 32)));
      idx_out[o] = static_cast<int64_t>(best & 0xffffffffull);
    }
  }
}

int knn_launch(const float* ref, const float* query, int B, int N, int Q, int k, float* dist_out,
               int64_t* idx_out, cudaStream_t st) {
  dim3 grid((Q + kKnnWarps - 1) / kKnnWarps, B);
  if (k <= kWarp) {
    const size_t smem = static_cast<size_t>(min(N, kKnnTile)) * 3 * sizeof(float);
    knn_warp_kernel<<<grid, kKnnWarps * kWarp, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out);
  } else {
    knn_extract_kernel<<<grid, kKnnWarps * kWarp, 0, st>>>(ref, query, N, Q, k, dist_out, idx_out);
  }
  count_launch();
  return launch_status();
}

}  // namespace upp
