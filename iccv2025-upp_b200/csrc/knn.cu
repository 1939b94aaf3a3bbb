// knn.cu -- exact brute-force kNN for sm_100a, one launch for the whole batch.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward (reference call site
// models/Point_MAE_unify.py:56,69).  Upstream (KNN_CUDA 0.2) runs a Python loop over the
// batch and, per cloud, materialises the N x Q distance matrix in global memory, then sorts
// each column with ONE THREAD per query (stride-Q global accesses), then a sqrt kernel.
//
// Here (knn_warp_kernel): the cloud's reference points are staged once per CTA into shared
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

#include "common.cuh"

namespace upp {

constexpr int kKnnWarps = 8;     // queries in flight per CTA
constexpr int kKnnTile = 2048;   // reference points staged per pass (24 KB)
constexpr int kKnnSlots = 32;    // distances per lane per block (block = 1024 refs)

__device__ __forceinline__ bool key_less(float da, int ia, float db, int ib) {
  return da < db || (da == db && ia < ib);
}

// Bitonic sort of one (d, i) pair per lane, ascending by (d, i) over lanes 0..31, via shuffles.
__device__ __forceinline__ void warp_bitonic_sort(float& d, int& i, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
      const bool ascending = ((lane & size) == 0);  // direction of the merge this lane is in
      const bool lower = ((lane & stride) == 0);    // lower lane of the compared pair
      const bool keep_min = (lower == ascending);
      const bool take = keep_min ? key_less(od, oi, d, i) : key_less(d, i, od, oi);
      if (take) { d = od; i = oi; }
    }
  }
}

// k <= 32.
template <bool GATHER>
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
  const float kInf = __int_as_float(0x7f800000);
  float ld = kInf;      // lane i: squared distance of the i-th best so far
  int li = 0x7fffffff;  //         and its reference index
  float thr_d = kInf;   // current k-th best (warp-uniform)
  int thr_i = 0x7fffffff;
  bool seeded = false;

  for (int base = 0; base < N; base += kKnnTile) {
    const int tile = min(kKnnTile, N - base);
    if (base > 0) __syncthreads();  // everyone done reading the previous tile
    stage_points(s_ref, rb + static_cast<size_t>(base) * 3, tile, &s_bar, parity);
    if (!active) continue;
    for (int blk = 0; blk < tile; blk += kKnnSlots * kWarp) {
      // ---- 1. distances of this block into registers; slot s <-> ref index blk + s*32 + lane ----
      float d[kKnnSlots];
      float lmin = kInf;
      int lmin_s = 0;
#pragma unroll
      for (int s = 0; s < kKnnSlots; ++s) {
        const int c = blk + s * kWarp + lane;
        d[s] = kInf;
        if (blk + s * kWarp < tile) {  // warp-uniform guard, then the per-lane tail
          if (c < tile) d[s] = dist_xyz_acc(s_ref[3 * c] - qx, s_ref[3 * c + 1] - qy, s_ref[3 * c + 2] - qz);
          if (d[s] < lmin) { lmin = d[s]; lmin_s = s; }  // strict '<': lowest index among equal minima
        }
      }
      // ---- 2. seed the list with the sorted lane minima (first block only) ----
      if (!seeded) {
        seeded = true;
        ld = lmin;
        li = lmin < kInf ? base + blk + lmin_s * kWarp + lane : 0x7fffffff;
        warp_bitonic_sort(ld, li, lane);
#pragma unroll
        for (int s = 0; s < kKnnSlots; ++s)
          if (s == lmin_s) d[s] = kInf;  // consumed
        thr_d = __shfl_sync(0xffffffffu, ld, k - 1);
        thr_i = __shfl_sync(0xffffffffu, li, k - 1);
      }
      // ---- 3. stream the register slots through the threshold filter ----
#pragma unroll
      for (int s = 0; s < kKnnSlots; ++s) {
        if (blk + s * kWarp < tile) {  // warp-uniform
          const int myi = base + blk + s * kWarp + lane;
          unsigned m = __ballot_sync(0xffffffffu, key_less(d[s], myi, thr_d, thr_i));
          if (m != 0) {
            while (m) {
              const int src = __ffs(m) - 1;
              m &= m - 1;
              const float cd = __shfl_sync(0xffffffffu, d[s], src);
              const int ci = base + blk + s * kWarp + src;
              const float ud = __shfl_up_sync(0xffffffffu, ld, 1);
              const int ui = __shfl_up_sync(0xffffffffu, li, 1);
              // entries greater than the candidate shift right by one; the first of them is replaced
              const bool mine_gt = key_less(cd, ci, ld, li);
              const bool left_gt = (lane > 0) && key_less(cd, ci, ud, ui);
              li = mine_gt ? (left_gt ? ui : ci) : li;
              ld = mine_gt ? (left_gt ? ud : cd) : ld;
            }
            thr_d = __shfl_sync(0xffffffffu, ld, k - 1);
            thr_i = __shfl_sync(0xffffffffu, li, k - 1);
          }
        }
      }
    }
  }
  if (active && lane < k) {
    const size_t o = (static_cast<size_t>(b) * Q + q) * k + lane;
    if (dist_out) dist_out[o] = __fsqrt_rn(ld);
    idx_out[o] = static_cast<int64_t>(li);
    if (GATHER) {
      const float* p = rb + static_cast<size_t>(li) * 3;
      nb_out[3 * o + 0] = __ldg(p) - qx;
      nb_out[3 * o + 1] = __ldg(p + 1) - qy;
      nb_out[3 * o + 2] = __ldg(p + 2) - qz;
    }
  }
}

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

// nb_out != nullptr fuses the Group gather (neighbourhood = ref[idx] - query).
int knn_launch(const float* ref, const float* query, int B, int N, int Q, int k, float* dist_out,
               int64_t* idx_out, float* nb_out, cudaStream_t st) {
  dim3 grid((Q + kKnnWarps - 1) / kKnnWarps, B);
  const int threads = kKnnWarps * kWarp;
  if (k <= kWarp) {
    const size_t smem = static_cast<size_t>(min(N, kKnnTile)) * 3 * sizeof(float);
    if (nb_out) knn_warp_kernel<true><<<grid, threads, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
    else knn_warp_kernel<false><<<grid, threads, smem, st>>>(ref, query, N, Q, k, dist_out, idx_out, nullptr);
  } else {
    if (nb_out) knn_extract_kernel<true><<<grid, threads, 0, st>>>(ref, query, N, Q, k, dist_out, idx_out, nb_out);
    else knn_extract_kernel<false><<<grid, threads, 0, st>>>(ref, query, N, Q, k, dist_out, idx_out, nullptr);
  }
  count_launch();
  return launch_status();
}

}  // namespace upp
