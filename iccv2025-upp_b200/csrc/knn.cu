// knn.cu -- exact brute-force kNN for sm_100a, one launch for the whole batch.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward (reference call site
// models/Point_MAE_unify.py:56,69).  Upstream (KNN_CUDA 0.2) runs a Python loop over the
// batch and, per cloud, materialises the N x Q distance matrix in global memory, then sorts
// each column with ONE THREAD per query (stride-Q global accesses), then a sqrt kernel.
//
// Here (knn_warp_kernel + topk.cuh): the cloud's reference points are staged once per CTA into shared
// memory by a TMA bulk copy; one WARP owns one query and works through the references in blocks
// of 1024 (32 per lane), all of a lane's distances held in registers.
//   1. per block, every lane computes its 32 distances (independent FMA chains) and its minimum;
//   2. first block: the 32 lane minima -- real (distance, index) pairs -- are sorted across the
//      warp by a 15-stage shuffle bitonic network and become the initial top-32 list, so the
//      k-th-best threshold is tight from the start (instead of 32 serial insertions from +inf);
//   3. each register slot is filtered against the current k-th best with one ballot; survivors are
//      admitted with one shuffle-up (every lane decides from its own and its left neighbour's
//      entry, no ballot/popc on the dependent chain).
// Order is the total order (squared distance, index): ascending distance, equal distances keep the
// lower reference index first == upstream's stable insertion sort.  d = fma(dz,dz,fma(dy,dy,dx*dx))
// with d* = ref - query; sqrt (IEEE rn) applied to the k survivors; int64 0-based indices.
// Optional fused epilogue (Group divider): neighbourhood[b,q,j,:] = ref[idx] - query, written by the
// lane that holds the j-th neighbour -- Group.forward's gather + centre subtraction
// (models/Point_MAE_unify.py:72-88) without another launch.
#include <float.h>

#include "topk.cuh"

namespace upp {

// k <= 32.  The per-warp selection itself lives in topk.cuh (shared with interp.cu).
// RAW selects the pytorch3d.ops.knn_points convention: SQUARED distances and, with GATHER, the neighbours
// themselves (no centre subtraction).
template <bool GATHER, bool RAW, int SLOTS>
__global__ void __launch_bounds__(kKnnWarps * kWarp)
    knn_warp_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N, int Q,
                    int k, float* __restrict__ dist_out, int64_t* __restrict__ idx_out,
                    float* __restrict__ nb_out) {
  extern __shared__ __align__(16) float s_ref[];  // min(N, kKnnTile) * 3 floats
  __shared__ __align__(8) uint64_t s_bar;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
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
  DistDirect dist;
  dist.set(qx, qy, qz);
  float ld;  // lane i: squared distance of the i-th nearest
  int li;    //         and its reference index
  if (SLOTS <= 8 && k <= 8) {  // the whole cloud is one block and few neighbours are wanted: warp arg-min rounds
    stage_points(s_ref, rb, N, &s_bar, parity);
    ld = __int_as_float(0x7f800000);
    li = 0;
    if (active) warp_topk_small<DistDirect, SLOTS>(s_ref, N, k, dist, ld, li);
  } else {
    warp_topk_scan<DistDirect, SLOTS>(rb, N, k, dist, active, s_ref, &s_bar, parity, ld, li);
  }
  if (active && lane < k) {
    const size_t o = (static_cast<size_t>(b) * Q + q) * k + lane;
    if (dist_out) dist_out[o] = RAW ? ld : __fsqrt_rn(ld);
    idx_out[o] = static_cast<int64_t>(li);
    if (GATHER) {
      const float* p = rb + static_cast<size_t>(li) * 3;
      nb_out[3 * o + 0] = RAW ? __ldg(p) : __ldg(p) - qx;
      nb_out[3 * o + 1] = RAW ? __ldg(p + 1) : __ldg(p + 1) - qy;
      nb_out[3 * o + 2] = RAW ? __ldg(p + 2) : __ldg(p + 2) - qz;
    }
  }
}

// (Measured and dropped, profiles/r01e_experiment_knn_split_warps.txt: 2 or 4 warps per query, each scanning a
//  contiguous part of the cloud and merging the sorted 32-lists with bitonic merges, bit-identical results --
//  64 q x 1024 at B = 32 went from 15.4 to 21.0 / 21.5 us, 128 q x 2048 from 29.7 to 42 / 52 us: the cost is the
//  per-warp seed sort and the serial insertion chain, which parts do not shorten, not the scan of the references.)

// k > 32 : selection by repeated extraction (no storage): round r finds the smallest
// (distance, index) key strictly greater than the previous round's.  O(k*N/32) per query;
// the reference never needs it (k in {8,16,32}), it keeps the ABI total.
template <bool GATHER>
__global__ void __launch_bounds__(kKnnWarps * kWarp)
    knn_extract_kernel(const float* __restrict__ ref, const float* __restrict__ query, int N,
                       int Q, int k, float* __restrict__ dist_out, int64_t* __restrict__ idx_out,
                       float* __restrict__ nb_out) {
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
      const unsigned c = static_cast<unsigned>(best & 0xffffffffull);
      if (dist_out) dist_out[o] = __fsqrt_rn(__uint_as_float(static_cast<unsigned>(best >> 32)));
      idx_out[o] = static_cast<int64_t>(c);
      if (GATHER) {
        nb_out[3 * o + 0] = __ldg(rb + 3 * c) - qx;
        nb_out[3 * o + 1] = __ldg(rb + 3 * c + 1) - qy;
        nb_out[3 * o + 2] = __ldg(rb + 3 * c + 2) - qz;
      }
    }
  }
}

template <bool GATHER, bool RAW>
static void launch_knn_warp(dim3 grid, size_t smem, cudaStream_t st, const float* ref, const float* query, int N,
                            int Q, int k, float* dist_out, int64_t* idx_out, float* nb_out) {
  const int threads = kKnnWarps * kWarp;
  // slots per lane sized to the cloud: short unrolled loops for the small second-level groupings
  if (N <= 128) knn_warp_kernel<GATHER, RAW, 4><<<grid, threads, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
  else if (N <= 256) knn_warp_kernel<GATHER, RAW, 8><<<grid, threads, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
  else knn_warp_kernel<GATHER, RAW, 32><<<grid, threads, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
}

// nb_out != nullptr fuses the Group gather (neighbourhood = ref[idx] - query).
int knn_launch(const float* ref, const float* query, int B, int N, int Q, int k, float* dist_out,
               int64_t* idx_out, float* nb_out, cudaStream_t st) {
  dim3 grid((Q + kKnnWarps - 1) / kKnnWarps, B);
  const int threads = kKnnWarps * kWarp;
  if (k <= kWarp) {
    const size_t smem = static_cast<size_t>(min(N, kKnnTile)) * 3 * sizeof(float);
    if (nb_out) launch_knn_warp<true, false>(grid, smem, st, ref, query, N, Q, k, dist_out, idx_out, nb_out);
    else launch_knn_warp<false, false>(grid, smem, st, ref, query, N, Q, k, dist_out, idx_out, nullptr);
  } else {
    if (nb_out) knn_extract_kernel<true><<<grid, threads, 0, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
    else knn_extract_kernel<false><<<grid, threads, 0, st>>>(ref, query, N, Q, k, dist_out, idx_out, nullptr);
  }
  count_launch();
  return launch_status();
}

// pytorch3d.ops.knn_points(p1, p2, K, return_nn) convention (k <= 32): squared distances ascending,
// int64 indices into p2, optionally the gathered neighbours p2[idx] (B,N1,K,3).
int knn_points_launch(const float* p1, const float* p2, int B, int N1, int N2, int k, float* dist2_out,
                      int64_t* idx_out, float* nn_out, cudaStream_t st) {
  dim3 grid((N1 + kKnnWarps - 1) / kKnnWarps, B);
  const size_t smem = static_cast<size_t>(min(N2, kKnnTile)) * 3 * sizeof(float);
  if (nn_out) launch_knn_warp<true, true>(grid, smem, st, p2, p1, N2, N1, k, dist2_out, idx_out, nn_out);
  else launch_knn_warp<false, true>(grid, smem, st, p2, p1, N2, N1, k, dist2_out, idx_out, nullptr);
  count_launch();
  return launch_status();
}

}  // namespace upp
