// scatter.cuh -- deterministic scatter-add, done as an ordered GATHER (no float atomics, no memset).
//
// Every backward on this path is "grad[dst[e]] += value(e)" over a per-cloud list of entries e -- the partner term of
// the Chamfer gradient (reference chamfer.cu:173-201: six float atomicAdd per point), the Group divider's index gather,
// the coordinate gather fused into FPS, pointnet2_ops' gather_points_grad_kernel.  With float atomics the summation
// order, and so the low bits of the gradient, change from run to run (SURVEY.md 5 asks for that to go).  Here a lane
// OWNS its destinations: a warp covers 32 * Q consecutive destinations of one cloud (lane l owns j0 + l + 32 q, its
// accumulators live in registers), the cloud's destination list is staged once per CTA in shared memory, and every
// warp scans the whole list in ASCENDING entry order, 32 entries per step: one ballot finds the entries that land in
// the warp's range, the lanes that hold them fetch their values (independent loads), and the values are handed to the
// owning lanes by shuffle, lowest entry first.  Per destination the additions therefore happen in ascending entry
// order whatever the hardware does: bit-identical run to run, one launch, every output element written exactly once.
// Cost: (N / 32Q) * (L / 32) scan steps per cloud, a few instructions each -- less than the memset + atomics it replaces
// at the sizes on this path (N, L <= 8192), and bounded for any distribution (all entries on one destination just
// serialise L shuffles in one warp).
#pragma once
#include "common.cuh"

namespace upp {

constexpr int kScatterTile = 8192;  // list entries staged per pass (32 KB)

// Op interface (all __device__):
//   int  entries() const                         list length L of this cloud
//   int  dst(int e) const                        destination of entry e (0 <= dst < N)
//   void fetch(int e, float (&v)[3]) const       value of entry e (executed by the lane that holds e)
//   void init(int j, float (&a)[3]) const        starting value of destination j (its own term, or zero)
//   void store(int j, const float (&a)[3]) const final value of destination j
// block_base: index of this op's first CTA along blockIdx.x (several ops can share one grid).
// Returns this lane's sum of squares of the values it stored (0 for lanes without destinations): the gradient-statistics
// users reduce it, everybody else ignores it.
template <int Q, class Op>
__device__ __forceinline__ float ordered_scatter_cta(const Op& op, int N, int* s_dst, int block_base = 0) {
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int nwarps = blockDim.x >> 5;
  const int j0 = ((static_cast<int>(blockIdx.x) - block_base) * nwarps + warp) * (32 * Q);
  const bool warp_live = j0 < N;  // warp-uniform; dead warps still help staging
  float acc[Q][3];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = j0 + lane + 32 * q;
    acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    if (warp_live && j < N) op.init(j, acc[q]);
  }
  const int L = op.entries();
  for (int base = 0; base < L; base += kScatterTile) {
    const int tile = min(kScatterTile, L - base);
    if (base > 0) __syncthreads();
    for (int e = threadIdx.x; e < tile; e += blockDim.x) s_dst[e] = op.dst(base + e);
    __syncthreads();
    if (!warp_live) continue;
    for (int e0 = 0; e0 < tile; e0 += 32) {
      const int e = e0 + lane;
      const int rel = e < tile ? s_dst[e] - j0 : -1;
      const bool hit = static_cast<unsigned>(rel) < static_cast<unsigned>(32 * Q);
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m == 0u) continue;  // warp-uniform
      float v[3] = {0.f, 0.f, 0.f};
      if (hit) op.fetch(base + e, v);
      while (m) {  // ascending entry order
        const int s = __ffs(m) - 1;
        m &= m - 1;
        const int r = __shfl_sync(0xffffffffu, rel, s);
        const float vx = __shfl_sync(0xffffffffu, v[0], s);
        const float vy = __shfl_sync(0xffffffffu, v[1], s);
        const float vz = __shfl_sync(0xffffffffu, v[2], s);
        if ((r & 31) == lane) {
#pragma unroll
          for (int q = 0; q < Q; ++q)
            if ((r >> 5) == q) {
              acc[q][0] = __fadd_rn(acc[q][0], vx);
              acc[q][1] = __fadd_rn(acc[q][1], vy);
              acc[q][2] = __fadd_rn(acc[q][2], vz);
            }
        }
      }
    }
  }
  float sq = 0.f;
  if (!warp_live) return sq;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = j0 + lane + 32 * q;
    if (j < N) {
      op.store(j, acc[q]);
      sq = __fmaf_rn(acc[q][2], acc[q][2], __fmaf_rn(acc[q][1], acc[q][1], __fmaf_rn(acc[q][0], acc[q][0], sq)));
    }
  }
  return sq;
}

// launch geometry shared by the users: warps per CTA and CTAs per cloud for N destinations, Q per lane
struct ScatterGrid {
  int warps, blocks;
  size_t smem;
};
inline ScatterGrid scatter_grid(int N, int L, int Q) {
  const int need = (N + 32 * Q - 1) / (32 * Q);  // warps per cloud
  ScatterGrid g;
  g.warps = need < 8 ? (need < 1 ? 1 : need) : 8;
  g.blocks = (need + g.warps - 1) / g.warps;
  if (g.blocks < 1) g.blocks = 1;
  g.smem = static_cast<size_t>(L < kScatterTile ? (L < 1 ? 1 : L) : kScatterTile) * sizeof(int);
  return g;
}

}  // namespace upp
