// scatter.cuh -- deterministic scatter-add, done as an ordered GATHER (no float atomics, no memset).
//
// Every backward on this path is "grad[dst[e]] += value(e)" over a per-cloud list of entries e -- the partner term of
// the Chamfer gradient (reference chamfer.cu:173-201: six float atomicAdd per point), the Group divider's index gather,
// the coordinate gather fused into FPS, pointnet2_ops' gather_points_grad_kernel.  With float atomics the summation
// order, and so the low bits of the gradient, change from run to run (SURVEY.md 5 asks for that to go).  Here a lane
// OWNS its destinations: a warp covers 32 * Q consecutive destinations of one cloud (their accumulators live in the
// warp's private slice of shared memory), the cloud's destination list is staged once per CTA in shared memory, and every
// warp scans the whole list in ASCENDING entry order, 32 entries per step: one ballot finds the entries that land in
// the warp's range; they are queued in order, their values fetched 32 at a time and added to the destinations'
// accumulators (the warp's slice of shared memory), lowest entry first.  Per destination the additions therefore happen in ascending entry
// order whatever the hardware does: bit-identical run to run, one launch, every output element written exactly once.
// Cost: (N / 32Q) * (L / 32) scan steps per cloud, a few instructions each -- less than the memset + atomics it replaces
// at the sizes on this path (N, L <= 8192), and bounded for any distribution (all entries on one destination just
// serialise L shuffles in one warp).
#pragma once
#include "common.cuh"

namespace upp {

constexpr int kScatterTile = 8192;   // list entries staged per pass (32 KB)
constexpr int kScatterMaxQ = 4;
// per-warp scratch: hit queue (64 x {entry, rel}) + the warp's accumulators (32 * Q destinations x 3 floats)
__host__ __device__ constexpr int scatter_warp_bytes(int Q) { return 64 * 8 + 32 * Q * 12; }

// Op interface (all __device__):
//   int  entries() const                                  list length L of this cloud
//   int  dst(int e) const                                 destination of entry e (0 <= dst < N)
//   void fetch(int e, int dst, float (&v)[3]) const       value of entry e (executed by one lane per entry)
//   void init(int j, float (&a)[3]) const                 starting value of destination j (its own term, or zero)
//   void store(int j, const float (&a)[3]) const          final value of destination j
//
// A warp scans the staged list 128 entries per step and only COLLECTS its hits ({entry, destination - j0}, compacted in
// entry order into a 64-slot queue in shared memory: ballot + popc, no global access on the scan).  Whenever 32 hits are
// queued they are flushed together: lane i fetches the value of hit i (32 independent loads in flight -- one global
// latency per 32 hits, not per scan step) and adds it to the destination's accumulator in the warp's shared-memory
// slice.  Hits of one flush that share a destination are ranked by MATCH.ANY + popc and added in rank order, one
// rank per round (typically every destination is distinct: one round).  Ascending entry order per destination.
// block_base: index of this op's first CTA along blockIdx.x (several ops can share one grid).
// Returns this lane's sum of squares of the values it stored (0 for lanes without destinations): the gradient-statistics
// users reduce it, everybody else ignores it.
template <int Q, class Op>
__device__ __forceinline__ float ordered_scatter_cta(const Op& op, int N, int* s_dst, int block_base = 0) {
  static_assert(Q >= 1 && Q <= kScatterMaxQ, "destinations per lane");
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int nwarps = blockDim.x >> 5;
  const int j0 = ((static_cast<int>(blockIdx.x) - block_base) * nwarps + warp) * (32 * Q);
  const bool warp_live = j0 < N;  // warp-uniform; dead warps still help staging
  // dynamic shared memory: [list tile][per warp: queue, accumulators]
  const int L = op.entries();
  const int list_cap = ((L < kScatterTile ? L : kScatterTile) + 3) & ~3;  // ints, 16-byte granules (scatter_grid agrees)
  char* wbase = reinterpret_cast<char*>(s_dst) + static_cast<size_t>(list_cap) * sizeof(int) +
                static_cast<size_t>(warp) * scatter_warp_bytes(Q);
  int2* s_q = reinterpret_cast<int2*>(wbase);
  float* s_acc = reinterpret_cast<float*>(wbase + 64 * 8);  // [32 * Q][3]: stride 3 words, conflict-free across lanes
  const unsigned lanes_below = (1u << lane) - 1u;
  pdl_trigger();  // (programmatic dependent launch, common.cuh: the gradient chain is a run of short dependent kernels)
  pdl_wait();     // everything read below may come from the kernel in front of this one
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = j0 + lane + 32 * q;
    float a[3] = {0.f, 0.f, 0.f};
    if (warp_live && j < N) op.init(j, a);
    s_acc[(lane + 32 * q) * 3 + 0] = a[0];
    s_acc[(lane + 32 * q) * 3 + 1] = a[1];
    s_acc[(lane + 32 * q) * 3 + 2] = a[2];
  }
  __syncwarp();
  int count = 0;  // queued hits (warp-uniform)
  // The values of a flushed batch are CONSUMED one flush later: the 32 loads issued here stay in flight while the scan
  // goes on (it needs shared memory only), so their latency is hidden behind the next scan steps instead of stalling
  // the warp once per batch.
  float pv[3] = {0.f, 0.f, 0.f};  // pending batch: this lane's fetched value ...
  int prel = -1 - lane;           // ... and its destination (negative: none; distinct keys so that nothing matches)
  bool pending = false;           // warp-uniform
  auto consume = [&]() {
    if (!pending) return;
    const unsigned same = __match_any_sync(0xffffffffu, prel);
    const int rank = __popc(same & lanes_below);          // position among the hits of this batch with my destination
    const int rounds = redux_max_s32(rank) + 1;
    for (int r = 0; r < rounds; ++r) {
      if (prel >= 0 && rank == r) {
        float* a = s_acc + prel * 3;
        a[0] = __fadd_rn(a[0], pv[0]);
        a[1] = __fadd_rn(a[1], pv[1]);
        a[2] = __fadd_rn(a[2], pv[2]);
      }
      __syncwarp();
    }
    pending = false;
  };
  auto flush = [&](int n) {  // the first n (<= 32) queued hits
    consume();
    pv[0] = pv[1] = pv[2] = 0.f;
    prel = -1 - lane;
    if (lane < n) {
      const int2 h = s_q[lane];
      prel = h.y;
      op.fetch(h.x, j0 + prel, pv);
    }
    pending = true;
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
    for (int e = threadIdx.x; e < tile; e += 8 * blockDim.x) {  // 8 independent loads in flight per thread
      int d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[u] = e + u * blockDim.x < tile ? op.dst(base + e + u * blockDim.x) : 0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (e + u * blockDim.x < tile) s_dst[e + u * blockDim.x] = d[u];
    }
    __syncthreads();
    if (!warp_live) continue;
    for (int e0 = 0; e0 < tile; e0 += 128) {
      int rel[4];
      unsigned m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + 32 * u + lane;
        rel[u] = e < tile ? s_dst[e] - j0 : -1;
        m[u] = __ballot_sync(0xffffffffu, static_cast<unsigned>(rel[u]) < static_cast<unsigned>(32 * Q));
      }
      if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;  // warp-uniform
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (m[u] == 0u) continue;
        if ((m[u] >> lane) & 1u) s_q[count + __popc(m[u] & lanes_below)] = make_int2(base + e0 + 32 * u + lane, rel[u]);
        count += __popc(m[u]);
        __syncwarp();
        if (count >= 32) flush(32);
      }
    }
  }
  float sq = 0.f;
  if (!warp_live) return sq;
  if (count > 0) flush(count);
  consume();
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = j0 + lane + 32 * q;
    if (j < N) {
      const float a[3] = {s_acc[(lane + 32 * q) * 3], s_acc[(lane + 32 * q) * 3 + 1], s_acc[(lane + 32 * q) * 3 + 2]};
      op.store(j, a);
      sq = __fmaf_rn(a[2], a[2], __fmaf_rn(a[1], a[1], __fmaf_rn(a[0], a[0], sq)));
    }
  }
  return sq;
}

// launch geometry shared by the users: warps per CTA and CTAs per cloud for N destinations, Q per lane
struct ScatterGrid {
  int warps, blocks;
  size_t smem;
};
inline size_t scatter_smem(int L, int warps, int Q) {
  const int list_cap = ((L < kScatterTile ? (L < 0 ? 0 : L) : kScatterTile) + 3) & ~3;
  return static_cast<size_t>(list_cap) * sizeof(int) + static_cast<size_t>(warps) * scatter_warp_bytes(Q);
}
// destinations per lane: 2 when the batch fills the GPU anyway (half the scan work), 1 when warps are scarce (latency)
inline int scatter_pick_q(int B, int N) { return static_cast<long>(B) * ((N + 63) / 64) >= 148L * 16 ? 2 : 1; }
inline ScatterGrid scatter_grid(int N, int L, int Q) {
  const int need = (N + 32 * Q - 1) / (32 * Q);  // warps per cloud
  ScatterGrid g;
  g.warps = need < 8 ? (need < 1 ? 1 : need) : (Q == 1 ? 4 : 8);
  g.blocks = (need + g.warps - 1) / g.warps;
  if (g.blocks < 1) g.blocks = 1;
  g.smem = scatter_smem(L, g.warps, Q);
  return g;
}

}  // namespace upp
