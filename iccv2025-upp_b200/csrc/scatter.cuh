// scatter.cuh -- deterministic scatter-add, done as an ordered GATHER (no float atomics, no memset).
//
// Every backward on this path is "grad[dst[e]] += value(e)" over a per-cloud list of entries e -- the partner term of
// the Chamfer gradient (reference chamfer.cu:173-201: six float atomicAdd per point), the Group divider's index gather,
// the coordinate gather fused into FPS, pointnet2_ops' gather_points_grad_kernel.  With float atomics the summation
// order, and so the low bits of the gradient, change from run to run (SURVEY.md 5 asks for that to go).  Here a lane
// OWNS its destinations: a warp covers 32 * Q consecutive destinations of one cloud (lane l owns j0 + l + 32 q, its
// accumulators live in registers), the cloud's destination list is staged once per CTA in shared memory, and every
// warp scans the whole list in ASCENDING entry order, 32 entries per step: one ballot finds the entries that land in
// the warp's range; they are queued in order, their values fetched 32 at a time and handed to the owning lanes
// through shared memory, lowest entry first.  Per destination the additions therefore happen in ascending entry
// order whatever the hardware does: bit-identical run to run, one launch, every output element written exactly once.
// Cost: (N / 32Q) * (L / 32) scan steps per cloud, a few instructions each -- less than the memset + atomics it replaces
// at the sizes on this path (N, L <= 8192), and bounded for any distribution (all entries on one destination just
// serialise L shuffles in one warp).
#pragma once
#include "common.cuh"

namespace upp {

constexpr int kScatterTile = 8192;   // list entries staged per pass (32 KB)
constexpr int kScatterWarpBytes = 64 * 8 + 32 * 16;  // per-warp hit queue (64 x {entry, rel}) + value stage (32 x float4)

// Op interface (all __device__):
//   int  entries() const                                  list length L of this cloud
//   int  dst(int e) const                                 destination of entry e (0 <= dst < N)
//   void fetch(int e, int dst, float (&v)[3]) const       value of entry e (executed by one lane per entry)
//   void init(int j, float (&a)[3]) const                 starting value of destination j (its own term, or zero)
//   void store(int j, const float (&a)[3]) const          final value of destination j
//
// A warp scans the staged list 32 entries per step and only COLLECTS its hits ({entry, destination - j0}, compacted in
// entry order into a 64-slot queue in shared memory: ballot + popc, no global access on the scan).  Whenever 32 hits are
// queued they are flushed together: lane i fetches the value of hit i (32 independent loads in flight -- one global
// latency per 32 hits, not per scan step), parks {rel, value} in the warp's stage, and every owning lane walks the 32
// staged hits in order (broadcast LDS.128) adding the ones addressed to it.  Ascending entry order per destination.
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
  // dynamic shared memory: [list tile: kScatterTile ints at most][per warp: queue, stage]
  const int L = op.entries();
  const int list_cap = ((L < kScatterTile ? L : kScatterTile) + 3) & ~3;  // ints, 16-byte granules (scatter_grid agrees)
  char* wbase = reinterpret_cast<char*>(s_dst) + static_cast<size_t>(list_cap) * sizeof(int) + static_cast<size_t>(warp) * kScatterWarpBytes;
  int2* s_q = reinterpret_cast<int2*>(wbase);
  float4* s_v = reinterpret_cast<float4*>(wbase + 64 * 8);
  const unsigned lanes_below = (1u << lane) - 1u;
  float acc[Q][3];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = j0 + lane + 32 * q;
    acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    if (warp_live && j < N) op.init(j, acc[q]);
  }
  int count = 0;  // queued hits (warp-uniform)
  auto flush = [&](int n) {  // the first n (<= 32) queued hits
    float v[3] = {0.f, 0.f, 0.f};
    int rel = -1;
    if (lane < n) {
      const int2 h = s_q[lane];
      rel = h.y;
      op.fetch(h.x, j0 + rel, v);
    }
    s_v[lane] = make_float4(__int_as_float(rel), v[0], v[1], v[2]);
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const float4 h = s_v[i];  // broadcast
      const int r = __float_as_int(h.x);
      const bool mine = (r & 31) == lane && r >= 0;
#pragma unroll
      for (int q = 0; q < Q; ++q) {  // selects, not branches: the accumulators stay in registers
        const bool on = mine && (r >> 5) == q;
        acc[q][0] = on ? __fadd_rn(acc[q][0], h.y) : acc[q][0];
        acc[q][1] = on ? __fadd_rn(acc[q][1], h.z) : acc[q][1];
        acc[q][2] = on ? __fadd_rn(acc[q][2], h.w) : acc[q][2];
      }
    }
    // the hits behind the flushed ones move to the front of the queue
    int2 keep = make_int2(0, 0);
    if (lane + 32 < count) keep = s_q[lane + 32];
    __syncwarp();
    if (lane + 32 < count) s_q[lane] = keep;
    count = count > 32 ? count - 32 : 0;
    __syncwarp();
  };
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
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m == 0u) continue;  // warp-uniform
      if (hit) s_q[count + __popc(m & lanes_below)] = make_int2(base + e, rel);
      count += __popc(m);
      __syncwarp();
      if (count >= 32) flush(32);
    }
  }
  float sq = 0.f;
  if (!warp_live) return sq;
  if (count > 0) flush(count);
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
  const int list_cap = ((L < kScatterTile ? (L < 0 ? 0 : L) : kScatterTile) + 3) & ~3;
  g.smem = static_cast<size_t>(list_cap) * sizeof(int) + static_cast<size_t>(g.warps) * kScatterWarpBytes;
  return g;
}

}  // namespace upp
