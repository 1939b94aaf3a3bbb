// topk.cuh -- warp-level exact top-k (k <= 32) of one query against a reference cloud staged in
// shared memory; shared by knn.cu (knn_cuda.KNN / Group) and interp.cu (propagate / feature
// propagation).  One WARP owns one query and works through the references in blocks of 1024
// (32 per lane), all of a lane's distances held in registers:
//   1. per block, every lane computes its 32 distances (independent FMA chains) and its minimum;
//   2. first block: the 32 lane minima -- real (distance, index) pairs -- are sorted across the
//      warp by a 15-stage shuffle bitonic network and become the initial top-32 list, so the
//      k-th-best threshold is tight from the start (instead of 32 serial insertions from +inf);
//   3. each register slot is filtered against the current k-th best with one ballot; survivors are
//      admitted with one shuffle-up (every lane decides from its own and its left neighbour's
//      entry, no ballot/popc on the dependent chain).
// Order is the total order (distance, index): ascending distance, equal distances keep the lower
// reference index first (== a stable sort by distance).
#pragma once
#include "common.cuh"

namespace upp {

constexpr int kKnnWarps = 8;     // queries in flight per CTA
constexpr int kKnnTile = 2048;   // reference points staged per pass (24 KB)
constexpr int kKnnSlots = 32;    // distances per lane per block (block = 1024 refs)

// ---- distance forms ---------------------------------------------------------------------
// KNN_CUDA 0.2 knn.cu (cuComputeDistanceGlobal): ssd += (r - q)^2 over x, y, z.
struct DistDirect {
  static constexpr bool kNonNegative = true;  // sums of squares: (distance, index) compares as one unsigned 64-bit key
  float qx, qy, qz;
  __device__ __forceinline__ void set(float x, float y, float z) { qx = x; qy = y; qz = z; }
  __device__ __forceinline__ float operator()(float rx, float ry, float rz) const {
    return dist_xyz_acc(rx - qx, ry - qy, rz - qz);
  }
};
// models/modules.py:13-32 square_distance(src, dst): -2 * (src . dst) + sum(src^2) + sum(dst^2),
// in that association (the form propagate / PointNetFeaturePropagation sort on).  It can be slightly
// negative for coincident points, exactly like the reference's.
struct DistExpanded {
  static constexpr bool kNonNegative = false;  // can be slightly negative: float compare
  float qx, qy, qz, s1;
  __device__ __forceinline__ void set(float x, float y, float z) {
    qx = x; qy = y; qz = z;
    s1 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  }
  __device__ __forceinline__ float operator()(float rx, float ry, float rz) const {
    const float dot = __fmaf_rn(qz, rz, __fmaf_rn(qy, ry, __fmul_rn(qx, rx)));
    const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), s1), s2);
  }
};

__device__ __forceinline__ bool key_less(float da, int ia, float db, int ib) {
  return da < db || (da == db && ia < ib);
}
// The same order for NON-NEGATIVE distances (no NaN) and non-negative indices, as one 64-bit unsigned compare of
// (distance bits : index) -- two ISETP instead of FSETP + FSETP + ISETP + PLOP3, on the serial insertion chain.
template <bool BITS>
__device__ __forceinline__ bool key_less_t(float da, int ia, float db, int ib) {
  if constexpr (BITS) {
    const unsigned long long ka = (static_cast<unsigned long long>(__float_as_uint(da)) << 32) | static_cast<unsigned>(ia);
    const unsigned long long kb = (static_cast<unsigned long long>(__float_as_uint(db)) << 32) | static_cast<unsigned>(ib);
    return ka < kb;
  } else {
    return key_less(da, ia, db, ib);
  }
}

// Bitonic sort of one (d, i) pair per lane, ascending by (d, i) over lanes 0..31, via shuffles.
template <bool BITS = false>
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
      const bool take = keep_min ? key_less_t<BITS>(od, oi, d, i) : key_less_t<BITS>(d, i, od, oi);
      if (take) { d = od; i = oi; }
    }
  }
}

// Running selection of one warp: lane i holds the i-th best so far, plus the warp-uniform k-th best.
struct TopkState {
  float ld;
  int li;
  float thr_d;
  int thr_i;
  bool seeded;
  __device__ __forceinline__ void init() {
    ld = thr_d = __int_as_float(0x7f800000);
    li = thr_i = 0x7fffffff;
    seeded = false;
  }
};

// One staged tile (s_ref holds `tile` references whose first has cloud index `base`) through the selection.
// SLOTS = distances per lane per block (block = 32 * SLOTS references): 32 for large clouds; 4 / 8 keep the
// unrolled loops short when the whole cloud is at most 128 / 256 points.
template <class Dist, int SLOTS>
__device__ __forceinline__ void warp_topk_tile(const float* s_ref, int tile, int base, int k, const Dist& dist,
                                               TopkState& st) {
  const int lane = threadIdx.x & 31;
  const float kInf = __int_as_float(0x7f800000);
  for (int blk = 0; blk < tile; blk += SLOTS * kWarp) {
    // ---- 1. distances of this block into registers; slot s <-> ref index blk + s*32 + lane ----
    float d[SLOTS];
    float lmin = kInf;
    int lmin_s = 0;
    const bool full = blk + SLOTS * kWarp <= tile;  // warp-uniform: every slot of the block is in range
    if (full) {  // no guards, one base address + immediates (the guards and 3*c products are ~15 % of a query's instructions)
      const float* pr = s_ref + 3 * (blk + lane);
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        d[s] = dist(pr[3 * kWarp * s], pr[3 * kWarp * s + 1], pr[3 * kWarp * s + 2]);
        if (d[s] < lmin) { lmin = d[s]; lmin_s = s; }  // strict '<': lowest index among equal minima
      }
    } else {
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int c = blk + s * kWarp + lane;
        d[s] = kInf;
        if (blk + s * kWarp < tile) {  // warp-uniform guard, then the per-lane tail
          if (c < tile) d[s] = dist(s_ref[3 * c], s_ref[3 * c + 1], s_ref[3 * c + 2]);
          if (d[s] < lmin) { lmin = d[s]; lmin_s = s; }
        }
      }
    }
    // ---- 2. seed the list with the sorted lane minima (first block only) ----
    // (Seeding with the 32 smallest of the 128 per-lane GROUP minima -- four row sorts + three top-32 merges, ~4x
    //  fewer survivors below -- was measured on B200 and lost: kNN 64q x 1024 went from 15.4 to 19.5 us.)
    // (Tightening the seed bound by counting -- a six-probe bisection over the sorted seeds for the smallest seed
    //  distance with >= k elements strictly below it, REDUX.ADD per probe -- cuts the insertions per 1024-reference
    //  query from ~100 to ~13 but costs as many instructions as it saves: 15.4 us either way, 29.2 -> 31.7 us at
    //  B = 128.  The per-query instruction budget (2400) is spread over control flow, not concentrated in insertions.)
    // (Seeding with the 32 smallest of the 64 per-lane TWO smallest -- one more row sort + one bitonic merge, ~12 insertions
    //  per 1024 references instead of ~100 -- measured in round 2: Group(64,32) 20.6 -> 20.8 us, Group(32,16) on 1096
    //  points 18.6 -> 20.5 us, interpolation forward 63.9 -> 66.8 us; only C4's throughput-bound Group(64,32) at B = 128
    //  gained, 37.0 -> 34.9 us.  With k = 16 the 16th of the 32 lane minima is tight already.  Left out.)
    if (!st.seeded) {
      st.seeded = true;
      st.ld = lmin;
      st.li = lmin < kInf ? base + blk + lmin_s * kWarp + lane : 0x7fffffff;
      warp_bitonic_sort<Dist::kNonNegative>(st.ld, st.li, lane);
#pragma unroll
      for (int s = 0; s < SLOTS; ++s)
        if (s == lmin_s) d[s] = kInf;  // consumed
      st.thr_d = __shfl_sync(0xffffffffu, st.ld, k - 1);
      st.thr_i = __shfl_sync(0xffffffffu, st.li, k - 1);
    }
    // ---- 3. stream the register slots through the threshold filter ----
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (full || blk + s * kWarp < tile) {  // warp-uniform
        const int myi = base + blk + s * kWarp + lane;
        unsigned m = __ballot_sync(0xffffffffu, key_less_t<Dist::kNonNegative>(d[s], myi, st.thr_d, st.thr_i));
        if (m != 0) {
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(0xffffffffu, d[s], src);
            const int ci = base + blk + s * kWarp + src;
            const float ud = __shfl_up_sync(0xffffffffu, st.ld, 1);
            const int ui = __shfl_up_sync(0xffffffffu, st.li, 1);
            // entries greater than the candidate shift right by one; the first of them is replaced
            const bool mine_gt = key_less_t<Dist::kNonNegative>(cd, ci, st.ld, st.li);
            const bool left_gt = (lane > 0) && key_less_t<Dist::kNonNegative>(cd, ci, ud, ui);
            st.li = mine_gt ? (left_gt ? ui : ci) : st.li;
            st.ld = mine_gt ? (left_gt ? ud : cd) : st.ld;
          }
          st.thr_d = __shfl_sync(0xffffffffu, st.ld, k - 1);
          st.thr_i = __shfl_sync(0xffffffffu, st.li, k - 1);
        }
      }
    }
  }
}

// Small problems (the whole cloud in ONE block of 32 * SLOTS references, few neighbours): k rounds of a warp
// arg-min instead of the sort-and-stream machinery -- per round two REDUX.MIN (value on a monotone unsigned image
// of the float, then lowest index among the lanes that hold it) and the winning lane retires its slot; ~25
// instructions per neighbour against ~400 for seeding + streaming.  Same total order (distance, index).
__device__ __forceinline__ unsigned float_to_ordered(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <class Dist, int SLOTS>
__device__ __forceinline__ void warp_topk_small(const float* s_ref, int n, int k, const Dist& dist, float& ld,
                                                int& li) {
  const int lane = threadIdx.x & 31;
  const float kInf = __int_as_float(0x7f800000);
  float d[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int c = s * kWarp + lane;
    d[s] = kInf;
    if (s * kWarp < n && c < n) d[s] = dist(s_ref[3 * c], s_ref[3 * c + 1], s_ref[3 * c + 2]);
  }
  ld = kInf;
  li = 0x7fffffff;
  for (int r = 0; r < k; ++r) {
    float lmin = d[0];
    int ls = 0;
#pragma unroll
    for (int s = 1; s < SLOTS; ++s)
      if (d[s] < lmin) { lmin = d[s]; ls = s; }  // strict '<': lowest slot == lowest index within the lane
    const unsigned key = float_to_ordered(lmin);
    const unsigned wmin = redux_min_u32(key);
    const unsigned cand = key == wmin ? static_cast<unsigned>(ls * kWarp + lane) : 0xffffffffu;
    const unsigned widx = redux_min_u32(cand);
    if (lane == r) {
      ld = ordered_to_float(wmin);
      li = static_cast<int>(widx);
    }
    if (cand == widx) {
#pragma unroll
      for (int s = 0; s < SLOTS; ++s)
        if (s == ls) d[s] = kInf;  // retired
    }
  }
}

// Scans the N references of one cloud (rb, AoS) for the k (<= 32) nearest to this warp's query.
// Every thread of the CTA must call it (the tile staging uses __syncthreads); warps with
// active == false only help staging.  On return lane j < k holds the j-th nearest (ld, li).
template <class Dist, int SLOTS = kKnnSlots>
__device__ __forceinline__ void warp_topk_scan(const float* __restrict__ rb, int N, int k, const Dist& dist,
                                               bool active, float* s_ref, uint64_t* s_bar, unsigned& parity,
                                               float& ld, int& li) {
  TopkState st;
  st.init();
  for (int base = 0; base < N; base += kKnnTile) {
    const int tile = min(kKnnTile, N - base);
    if (base > 0) __syncthreads();  // everyone done reading the previous tile
    stage_points(s_ref, rb + static_cast<size_t>(base) * 3, tile, s_bar, parity);
    if (active) warp_topk_tile<Dist, SLOTS>(s_ref, tile, base, k, dist, st);
  }
  ld = st.ld;
  li = st.li;
}

}  // namespace upp
