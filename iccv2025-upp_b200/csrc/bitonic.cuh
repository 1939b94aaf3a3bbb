// bitonic.cuh -- CTA-wide bitonic sort of up to 8192 keys staged in shared memory (crop.cu: 64-bit (distance, index) keys;
// fps_pruned.cu: 32-bit (Morton code, index) keys).  A thread owns KPT consecutive keys, a warp 32 * KPT: every
// compare-exchange at distance < KPT stays in registers, distances KPT ... 16 * KPT are lane-xor shuffles, and only the
// distances >= 32 * KPT go through shared memory with a CTA barrier (8192 keys, KPT = 8: 21 barrier-separated passes for
// the 91 stages of the network).  Keys must be distinct (they carry the element index in their low bits), so a
// compare-exchange is one compare + one select per element.
#pragma once
#include "common.cuh"

namespace upp {

template <class Key, int KPT>
__device__ __forceinline__ void bitonic_warp_pass(Key (&v)[KPT], int size, int e0, int lane) {
  // distances min(size / 2, 16 * KPT) ... 1 of the merge step `size`
  for (int stride = min(size >> 1, 16 * KPT); stride >= KPT; stride >>= 1) {
    const int lx = stride / KPT;
    const bool keep_min = ((lane & lx) == 0) == ((e0 & size) == 0);
#pragma unroll
    for (int r = 0; r < KPT; ++r) {
      const Key o = __shfl_xor_sync(0xffffffffu, v[r], lx);
      v[r] = ((v[r] < o) == keep_min) ? v[r] : o;
    }
  }
#pragma unroll
  for (int stride = KPT / 2; stride >= 1; stride >>= 1) {
    if (stride <= (size >> 1)) {
#pragma unroll
      for (int r = 0; r < KPT; ++r) {
        if ((r & stride) == 0) {
          const bool up = ((e0 + r) & size) == 0;
          const Key x = v[r], y = v[r | stride];
          const bool keep = (x < y) == up;
          v[r] = keep ? x : y;
          v[r | stride] = keep ? y : x;
        }
      }
    }
  }
}

// Sorts s_key[0 .. npow2) ascending.  npow2: a power of two, a multiple of 32 * KPT, at most KPT * blockDim.x.
// Every thread of the CTA must call it; ends with a __syncthreads().
template <class Key, int KPT>
__device__ __forceinline__ void bitonic_sort_cta(Key* s_key, int npow2) {
  const int t = threadIdx.x, lane = t & 31, threads = blockDim.x;
  const int e0 = t * KPT;
  const bool owner = e0 < npow2;  // whole warps
  Key v[KPT];
  if (owner) {
#pragma unroll
    for (int r = 0; r < KPT; ++r) v[r] = s_key[e0 + r];
    for (int size = 2; size <= 32 * KPT; size <<= 1) bitonic_warp_pass<Key, KPT>(v, size, e0, lane);
#pragma unroll
    for (int r = 0; r < KPT; ++r) s_key[e0 + r] = v[r];
  }
  __syncthreads();
  for (int size = 64 * KPT; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride >= 32 * KPT; stride >>= 1) {
      for (int i = t; i < (npow2 >> 1); i += threads) {
        const int lo_i = 2 * i - (i & (stride - 1));
        const int hi_i = lo_i + stride;
        const Key x = s_key[lo_i], y = s_key[hi_i];
        const bool keep = (x < y) == ((lo_i & size) == 0);
        s_key[lo_i] = keep ? x : y;
        s_key[hi_i] = keep ? y : x;
      }
      __syncthreads();
    }
    if (owner) {
#pragma unroll
      for (int r = 0; r < KPT; ++r) v[r] = s_key[e0 + r];
      bitonic_warp_pass<Key, KPT>(v, size, e0, lane);
#pragma unroll
      for (int r = 0; r < KPT; ++r) s_key[e0 + r] = v[r];
    }
    __syncthreads();
  }
}

}  // namespace upp
