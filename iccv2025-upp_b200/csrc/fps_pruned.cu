// fps_pruned.cu -- farthest point sampling of LARGE clouds (2048 < N <= 8192) with exact spatial pruning.
//
// STATUS: measured experiment, NOT on the default path (fps.cu keeps it behind UPP_FPS_PRUNED=1).  It is exact (every
// FPS parity test runs through it) and it removes 85 % of the distance updates, but the round it leaves is a chain of
// dependent warp-wide reductions -- bucket maximum + arg (2 REDUX), warp winner (2), block winner (2), ~60-80 cycles
// each with the move out of the uniform register file -- and lands at 0.50 us per round whatever N and whatever the
// warp count (4 / 8 / 16 measured: 638 / 613 / 578 us for B128 8192 -> 1024 against 555 us plain and 318 us for the
// cluster kernel at B16).  Also tried: lane-local candidates (bucket-level reductions off the dependent chain) + one 64-bit
// shared-memory atomicMax per warp in place of the block stage -- 642-696 us, slower still.  The bound is not yet
// understood (issue slots 47 %, barrier wait 33 % of the samples: the slowest warp of a round, the one with two or three
// buckets to update); it needs a per-round timeline, not more guesses.
//
// Same operator as fps.cu (pointnet2_ops furthest_point_sample, reference call sites utils/misc.py:18,
// tools/runner_module.py:310: the 8192 -> 1024 resampling of every ShapeNet55 batch), same results bit for bit.  What
// changes is the work per round.  In the plain kernels every round updates the running min-distance of EVERY point:
// 8 * N flop and, at N = 8192, ~0.54 us of FMA / ALU pipe time on one SM (profiles/r02_fps8k.txt: issue 44 %, FMA pipe
// 38 %, ALU 35 %) -- 1023 rounds = 555 us for C4's FPS whatever the batch; spreading a cloud over a cluster of SMs
// (fps_cluster_kernel) trades that for a 215-cycle DSMEM exchange per round (0.30 us).  But a point's min-distance only
// changes when the new centre is closer than its current value, and after a few dozen centres that is true for a small
// neighbourhood of the new centre only.  So:
//   * the cloud is sorted once along a Z-order (Morton) curve inside the CTA (bitonic.cuh, 32-bit keys = 18-bit code :
//     13-bit index) and cut into BUCKETS of 128 consecutive points, each with its bounding box, its current maximum
//     min-distance and the point that holds it;
//   * per round a warp tests its (<= 4) buckets lane-parallel: if the squared distance from the new centre to the
//     bucket's box exceeds the bucket's maximum min-distance, NO point of the bucket can change -- the bucket is skipped,
//     its cached (maximum, arg-max) stays valid.  Measured on C4's clouds: 83 % of all bucket-rounds are skipped (94 %
//     after the first 64 centres);
//   * a bucket that cannot be skipped is updated by the whole warp (4 points per lane from shared memory: the same
//     fma(dz,dz,fma(dx,dx,dy*dy)) and fminf as everywhere), its maximum and arg-max recomputed by two REDUX;
//   * the round's arg-max is a reduction over bucket maxima: two REDUX per warp, one barrier, two REDUX over the warps.
// Skipped points would have been left unchanged by the full update (fminf(md, d) == md), so min-distances, maxima and
// selections are IDENTICAL to the unpruned kernels; the skip test carries a 2^-18 relative safety margin against the
// rounding of the box distance.  Ties resolve to the lowest ORIGINAL point index (the sort permutes positions, so the
// arg-max key carries the original index above the sorted position).  Upstream semantics kept: start at index 0,
// temp = 1e10, points with x^2+y^2+z^2 <= 1e-3 never selected, M > N allowed.
#include <limits.h>
#include <stdio.h>

#include <type_traits>

#include "bitonic.cuh"
#include "fps_round.cuh"

namespace upp {

constexpr int kPrBucket = 128;           // points per bucket: one warp pass of 4 points per lane
constexpr int kPrMaxN = 8192;

__device__ __forceinline__ unsigned morton18(unsigned x, unsigned y, unsigned z) {  // 6 bits per axis, x lowest
  unsigned c = 0;
#pragma unroll
  for (int b = 0; b < 6; ++b) c |= (((x >> b) & 1u) << (3 * b)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b + 2));
  return c;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// NW warps per cloud: every warp pays the round's fixed part (box tests, two reduction stages: ~75 instructions), so few
// fat warps beat many thin ones -- profiles/r02_fps_pruned.txt: with 16 warps the SM issued 1900 warp instructions per
// round, two thirds of them that fixed part.  A warp owns 64 / NW buckets and updates them two at a time (two
// independent LDS -> FMA -> REDUX chains in flight).
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
    fps_pruned_kernel(const float* __restrict__ xyz, int N, int NP, int npow2, int M, int32_t* __restrict__ idx_out,
                      float* __restrict__ centers_out) {
  constexpr int kPrWarps = NW, kPrThreads = NW * 32, kPrKpt = kPrMaxN / kPrThreads;
  // dynamic shared memory: sorted cloud (3 * NP floats), then min-distances aliasing the sort keys (max(NP, npow2) words),
  // then the arg-max key of every sorted position: (original index << 13) | position (NP words)
  extern __shared__ __align__(16) float s_pts[];
  float* s_md = s_pts + 3 * NP;
  unsigned* s_key = reinterpret_cast<unsigned*>(s_md);
  unsigned* s_pkey = reinterpret_cast<unsigned*>(s_md + (NP > npow2 ? NP : npow2));
  __shared__ float s_red[6][kPrWarps];
  __shared__ int2 s_slot[2][kPrWarps];

  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  // ---- 1. bounding box of the cloud -> 6-bit cell per axis -> (Morton code : index) keys ----
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = t; i < N; i += kPrThreads) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(p + 3 * i + a);
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = warp_min_f(lo[a]);
    hi[a] = warp_max_f(hi[a]);
    if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  float scale[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float l = s_red[a][0], h = s_red[3 + a][0];
    for (int w = 1; w < kPrWarps; ++w) { l = fminf(l, s_red[a][w]); h = fmaxf(h, s_red[3 + a][w]); }
    lo[a] = l;
    scale[a] = h > l ? 63.99f / (h - l) : 0.f;  // (any monotone cell map will do: the sort only shapes the buckets)
  }
  for (int i = t; i < npow2; i += kPrThreads) {
    unsigned key = 0xffffffffu;  // padding sorts last
    if (i < N) {
      unsigned q[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float f = (__ldg(p + 3 * i + a) - lo[a]) * scale[a];
        q[a] = static_cast<unsigned>(fminf(fmaxf(f, 0.f), 63.f));
      }
      key = (morton18(q[0], q[1], q[2]) << 13) | static_cast<unsigned>(i);
    }
    s_key[i] = key;
  }
  __syncthreads();
  bitonic_sort_cta<unsigned, kPrKpt>(s_key, npow2);

  // ---- 2. sorted cloud, original indices, initial min-distances (the keys' words are reused: read, then written) ----
  for (int pos = t; pos < NP; pos += kPrThreads) {
    const unsigned key = pos < npow2 ? s_key[pos] : 0xffffffffu;
    float x = 0.f, y = 0.f, z = 0.f, md = kOutOfRange;
    unsigned orig = 0;
    if (pos < N) {
      const int i = static_cast<int>(key & 8191u);
      x = __ldg(p + 3 * i); y = __ldg(p + 3 * i + 1); z = __ldg(p + 3 * i + 2);
      md = fps_initial_md(x, y, z);
      orig = static_cast<unsigned>(i);
    }
    s_pts[3 * pos] = x; s_pts[3 * pos + 1] = y; s_pts[3 * pos + 2] = z;
    s_pkey[pos] = (orig << 13) | static_cast<unsigned>(pos);
    s_md[pos] = md;  // (same word as s_key[pos]: this thread has read it; NP <= npow2, the keys beyond NP are dead)
  }
  __syncthreads();

  // ---- 3. buckets: warp w owns buckets t * NW + w (t < bpw) -- INTERLEAVED, because the buckets a new centre touches are
  //         neighbours on the Z-curve: with contiguous ownership they would all be updated by the same warp, one after
  //         the other, while the other 15 wait at the barrier.  Lane t keeps bucket t's box / maximum / arg-max ----
  const int nbk = NP / kPrBucket;
  const int bpw = (nbk + kPrWarps - 1) / kPrWarps;
  float blo[3] = {0.f, 0.f, 0.f}, bhi[3] = {0.f, 0.f, 0.f};
  int bmax = INT_MAX;          // order-preserving bits of the bucket's largest min-distance; INT_MAX: not computed yet
  unsigned bkey = 0xffffffffu; // (original index << 13) | sorted position of the point that holds it
  const int my_bucket = lane * kPrWarps + warp;
  const bool owns = lane < bpw && my_bucket < nbk;
  for (int tb = 0; tb < bpw; ++tb) {  // boxes (points past N are excluded: their slots repeat the bucket's first point)
    const int bk = tb * kPrWarps + warp;
    if (bk >= nbk) break;  // warp-uniform
    float l3[3] = {3.4e38f, 3.4e38f, 3.4e38f}, h3[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int pos = bk * kPrBucket + lane * 4 + r;
      if (pos < N) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          l3[a] = fminf(l3[a], s_pts[3 * pos + a]);
          h3[a] = fmaxf(h3[a], s_pts[3 * pos + a]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      l3[a] = warp_min_f(l3[a]);
      h3[a] = warp_max_f(h3[a]);
      if (lane == tb) { blo[a] = l3[a]; bhi[a] = h3[a]; }
    }
  }

  // first centre: original index 0 -- its coordinates straight from global memory
  float cx = __ldg(p), cy = __ldg(p + 1), cz = __ldg(p + 2);
  if (t == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }

  for (int j = 1; j < M; ++j) {
    // ---- a. which of my buckets can change?  lane-parallel box test ----
    bool need = false;
    if (owns) {
      const float ex = fmaxf(fmaxf(blo[0] - cx, cx - bhi[0]), 0.f);
      const float ey = fmaxf(fmaxf(blo[1] - cy, cy - bhi[1]), 0.f);
      const float ez = fmaxf(fmaxf(blo[2] - cz, cz - bhi[2]), 0.f);
      const float dlb = __fmul_rn(__fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey))), 0.99999619f);  // 1 - 2^-18
      // skip only when every point's distance provably exceeds the bucket's largest min-distance (negative maxima:
      // nothing selectable left in the bucket, nothing to update either)
      need = bmax >= 0 && !(dlb > __int_as_float(bmax));  // (INT_MAX reads as NaN: the comparison fails, the bucket is updated)
    }
    unsigned todo = __ballot_sync(0xffffffffu, need);
    // ---- b. update the buckets that can, two at a time ----
    auto load_bucket = [&](int tb, float4& a0, float4& a1, float4& a2, float4& m, uint4& pk, int& base) {
      base = (tb * kPrWarps + warp) * kPrBucket + lane * 4;
      a0 = *reinterpret_cast<const float4*>(s_pts + 3 * base);      // x0 y0 z0 x1
      a1 = *reinterpret_cast<const float4*>(s_pts + 3 * base + 4);  // y1 z1 x2 y2
      a2 = *reinterpret_cast<const float4*>(s_pts + 3 * base + 8);  // z2 x3 y3 z3
      m = *reinterpret_cast<const float4*>(s_md + base);
      pk = *reinterpret_cast<const uint4*>(s_pkey + base);
    };
    auto update_bucket = [&](const float4& a0, const float4& a1, const float4& a2, float4& m, const uint4& pk, int base,
                             int& lmax, unsigned& lkey) {
      m.x = fminf(m.x, dist_yxz(a0.x - cx, a0.y - cy, a0.z - cz));
      m.y = fminf(m.y, dist_yxz(a0.w - cx, a1.x - cy, a1.y - cz));
      m.z = fminf(m.z, dist_yxz(a1.z - cx, a1.w - cy, a2.x - cz));
      m.w = fminf(m.w, dist_yxz(a2.y - cx, a2.z - cy, a2.w - cz));
      *reinterpret_cast<float4*>(s_md + base) = m;
      const int k0 = __float_as_int(m.x), k1 = __float_as_int(m.y), k2 = __float_as_int(m.z), k3 = __float_as_int(m.w);
      lmax = max(max(k0, k1), max(k2, k3));
      lkey = min(min(k0 == lmax ? pk.x : 0xffffffffu, k1 == lmax ? pk.y : 0xffffffffu),
                 min(k2 == lmax ? pk.z : 0xffffffffu, k3 == lmax ? pk.w : 0xffffffffu));
    };
    while (todo) {
      const int tb0 = __ffs(todo) - 1;
      todo &= todo - 1;
      if (todo) {  // warp-uniform: two buckets in flight
        const int tb1 = __ffs(todo) - 1;
        todo &= todo - 1;
        float4 p0, p1, p2, pm, q0, q1, q2, qm;
        uint4 pk, qk;
        int pb, qb, plmax, qlmax;
        unsigned plkey, qlkey;
        load_bucket(tb0, p0, p1, p2, pm, pk, pb);
        load_bucket(tb1, q0, q1, q2, qm, qk, qb);
        update_bucket(p0, p1, p2, pm, pk, pb, plmax, plkey);
        update_bucket(q0, q1, q2, qm, qk, qb, qlmax, qlkey);
        const int pw = redux_max_s32(plmax), qw = redux_max_s32(qlmax);
        const unsigned pwk = redux_min_u32(plmax == pw ? plkey : 0xffffffffu);
        const unsigned qwk = redux_min_u32(qlmax == qw ? qlkey : 0xffffffffu);
        if (lane == tb0) { bmax = pw; bkey = pwk; }
        if (lane == tb1) { bmax = qw; bkey = qwk; }
      } else {
        float4 p0, p1, p2, pm;
        uint4 pk;
        int pb, plmax;
        unsigned plkey;
        load_bucket(tb0, p0, p1, p2, pm, pk, pb);
        update_bucket(p0, p1, p2, pm, pk, pb, plmax, plkey);
        const int pw = redux_max_s32(plmax);
        const unsigned pwk = redux_min_u32(plmax == pw ? plkey : 0xffffffffu);
        if (lane == tb0) { bmax = pw; bkey = pwk; }
      }
    }
    // ---- c. warp winner over its buckets, block winner over the warps ----
    const int v = owns ? bmax : INT_MIN;
    const int wbest = redux_max_s32(v);
    const unsigned wk = redux_min_u32(v == wbest ? bkey : 0xffffffffu);
    int2* slot = s_slot[j & 1];
    if (lane == 0) slot[warp] = make_int2(wbest, static_cast<int>(wk));
    __syncthreads();
    const int2 s = lane < kPrWarps ? slot[lane] : make_int2(INT_MIN, -1);
    const int cbest = redux_max_s32(s.x);
    const unsigned ck = redux_min_u32(s.x == cbest ? static_cast<unsigned>(s.y) : 0xffffffffu);
    const int pos = static_cast<int>(ck & 8191u);
    cx = s_pts[3 * pos];
    cy = s_pts[3 * pos + 1];
    cz = s_pts[3 * pos + 2];
    if (t == 0) {
      out[j] = static_cast<int32_t>(ck >> 13);
      if (cen) { cen[3 * j] = cx; cen[3 * j + 1] = cy; cen[3 * j + 2] = cz; }
    }
  }
}

template <int NW>
static int launch_pruned(const float* xyz, int B, int N, int M, int32_t* idx, float* centers, cudaStream_t st) {
  const int NP = (N + kPrBucket - 1) / kPrBucket * kPrBucket;
  int npow2 = 32 * (kPrMaxN / (NW * 32));  // one warp-sorted block at least
  while (npow2 < N) npow2 <<= 1;
  const size_t words = static_cast<size_t>(NP > npow2 ? NP : npow2);
  const size_t smem = static_cast<size_t>(NP) * 12 + words * 4 + static_cast<size_t>(NP) * 4 + 16;
  cudaError_t e = cudaFuncSetAttribute(fps_pruned_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  fps_pruned_kernel<NW><<<B, NW * 32, smem, st>>>(xyz, N, NP, npow2, M, idx, centers);
  count_launch();
  return launch_status();
}


// ---------------------------------------------------------------------------------------------------------------------
// fps_rows_kernel -- the third cut of the pruned round (UPP_FPS_PRUNED=2; an EXPERIMENT like the kernel above: exact, every
// FPS parity test runs through it, NOT on the default path).  It was built to answer "where does a pruned round go?" with a
// per-phase timeline (clock64 build: make EXTRA=-DUPP_ROWS_PROFILE, scripts/prof_rows.py) and ncu source counters
// (profiles/r02_fps_rows.txt).  Design:
//   * the sorted cloud lives in REGISTERS exactly as in fps_blk_kernel (packed fp32x2 coordinates, running min-distances),
//     but ownership is by ROWS: row r of warp w is the 64 consecutive sorted points [(w * RPW + r) * 64, +64), two per lane.
//     A row is a spatial bucket (Z-curve neighbours) with its box in lane r; four rows (a NIBBLE) share one bound, the
//     nibble's largest min-distance;
//   * a round starts with the lane-parallel box test (one ballot -> a warp-uniform mask of rows that can change), then
//     updates the nibbles that hold such a row behind warp-uniform branches (four interleaved chains; a row that did not
//     need it comes out unchanged).  A warp with an empty mask copies its previous record and goes straight to the barrier;
//   * the arg-max costs ONE REDUX per active warp: every lane keeps a 4-ary max tree over its rows, position-independent --
//     ties are not ordered by position here (the sort permutes the indices) but DETECTED (two equal children on the
//     winning lane's tree descent, two lanes in the ballot, two warps within 8 ulp) and sent to a slow path that takes the
//     lowest ORIGINAL index over all points holding the maximum.  Real clouds never take it; lattices and duplicates do;
//   * the winning lane posts {key, position | tie flag} AND the point's coordinates (its descent ends at a compile-time
//     register) and enters (key & ~7) | warp and (key & ~7) | (7 - warp) in two words with native 32-bit shared-memory
//     atomic maxima: after the barrier one LDS.64 names the winning warp (both words agree unless two warps are within
//     8 ulp: exact path), one LDS.128 of its record is the next centre;
//   * the nibble bounds are reduced (one REDUX each) with the update and consumed after the barrier.
// MEASURED (B200, profiles/r02_fps_rows.txt, scripts/gpu_quick_fps.py): the pruning works -- 2 of 16 rows per warp and
// round are due, 45 % of the warp-rounds have nothing to do -- and the kernel is still SLOWER than the plain one:
// 0.65 us per round at N = 8192 (plain 0.54, the cluster kernel 0.31), 0.40 us at N = 1228 where the update is free
// (plain 0.15).  The timeline says why.  A round is not bound by distance work and not by the four or five long-latency
// steps one counts on paper; it is bound by the NUMBER OF DEPENDENT INSTRUCTIONS on the busiest warp's path, each worth
// 5-6 cycles (stall reason "wait", the fixed dependent-issue latency, 36 % of all samples; branch resolution another
// 10 %): box test 25, nibble update 39 (the only part with instruction-level parallelism), tree + REDUX + vote 10, tree
// descent with tie detection ~25, record + atomics 10, resolve 20, bounds 8, bookkeeping 25: ~160 instructions = 800
// cycles.  fps_blk_kernel's round at N = 1228 is 109 instructions of which 75 are the 5-way interleaved update: 35
// dependent ones, 293 cycles.  Any pruning scheme that adds more than ~30 dependent instructions to the round loses what
// it saves, on every cloud size this library sees; the three versions of this file (six REDUX, lane candidates, rows)
// added 60 to 120.
// What would pay instead (simulated, not built): amortising the chain over several selections.  With per-warp (best,
// runner-up bound) records the runner-up of a round is provably the next selection whenever it is the unique maximum of
// the other warps' bests and the winner's bound and lies farther from the winner than its own min-distance; on the C4
// clouds that accepts 1.82 selections per round with one look-ahead and 2.66 with three (8 Morton regions) -- but every
// extra candidate costs another REDUX + vote + record load (~150 cycles) in the resolve, which eats the gain at today's
// round length.
template <int P2>
struct RowTree {
  static constexpr int NG = P2 / 4;  // nibbles of four rows: the update / bound granularity and the tree's inner level
  static_assert(P2 % 4 == 0 && NG >= 1 && NG <= 4, "4, 8, 12 or 16 rows per warp");
  int l0[P2], nib[NG];  // l0[r] = max of row r's two keys (order-preserving float bits), nib[g] = max of rows 4g .. 4g+3

  template <int G>
  __device__ __forceinline__ void fold() { nib[G] = max(max(l0[4 * G], l0[4 * G + 1]), max(l0[4 * G + 2], l0[4 * G + 3])); }
  __device__ __forceinline__ int top() const {
    int v = nib[0];
#pragma unroll
    for (int g = 1; g < NG; ++g) v = max(v, nib[g]);
    return v;
  }
  __device__ __forceinline__ int build() {
    fold<0>();
    if constexpr (NG > 1) fold<1>();
    if constexpr (NG > 2) fold<2>();
    if constexpr (NG > 3) fold<3>();
    return top();
  }
  // A SLOT (2 * row + half) whose key equals `best` (the root) and its coordinates (every leaf is a compile-time
  // register).  `tie` is raised when two children of a visited node both hold `best`: any second slot holding the
  // maximum shares such a node with the one returned.
  template <int R>
  __device__ __forceinline__ int leaf(int best, bool& tie, const float (&md)[2 * P2], const f32x2 (&X)[P2],
                                      const f32x2 (&Y)[P2], const f32x2 (&Z)[P2], float4& rec) const {
    const bool e0 = __float_as_int(md[2 * R]) == best, e1 = __float_as_int(md[2 * R + 1]) == best;
    tie = tie || (e0 && e1);
    float x0, x1, y0, y1, z0, z1;
    unpack2(X[R], x0, x1); unpack2(Y[R], y0, y1); unpack2(Z[R], z0, z1);
    rec.x = e0 ? x0 : x1; rec.y = e0 ? y0 : y1; rec.z = e0 ? z0 : z1;
    return 2 * R + (e0 ? 0 : 1);
  }
  template <int G>
  __device__ __forceinline__ int in_nibble(int best, bool& tie, const float (&md)[2 * P2], const f32x2 (&X)[P2],
                                           const f32x2 (&Y)[P2], const f32x2 (&Z)[P2], float4& rec) const {
    const bool e0 = l0[4 * G] == best, e1 = l0[4 * G + 1] == best, e2 = l0[4 * G + 2] == best, e3 = l0[4 * G + 3] == best;
    tie = tie || (e0 && (e1 || e2 || e3)) || (e1 && (e2 || e3)) || (e2 && e3);
    if (e0) return leaf<4 * G>(best, tie, md, X, Y, Z, rec);
    if (e1) return leaf<4 * G + 1>(best, tie, md, X, Y, Z, rec);
    if (e2) return leaf<4 * G + 2>(best, tie, md, X, Y, Z, rec);
    return leaf<4 * G + 3>(best, tie, md, X, Y, Z, rec);
  }
  __device__ __forceinline__ int find(int best, bool& tie, const float (&md)[2 * P2], const f32x2 (&X)[P2],
                                      const f32x2 (&Y)[P2], const f32x2 (&Z)[P2], float4& rec) const {
    if constexpr (NG == 1) {
      return in_nibble<0>(best, tie, md, X, Y, Z, rec);
    } else {
      int holders = 0;
#pragma unroll
      for (int g = 0; g < NG; ++g) holders += nib[g] == best ? 1 : 0;
      tie = tie || holders > 1;
      if (nib[0] == best) return in_nibble<0>(best, tie, md, X, Y, Z, rec);
      if constexpr (NG == 2) {
        return in_nibble<1>(best, tie, md, X, Y, Z, rec);
      } else {
        if (nib[1] == best) return in_nibble<1>(best, tie, md, X, Y, Z, rec);
        if constexpr (NG == 3) {
          return in_nibble<2>(best, tie, md, X, Y, Z, rec);
        } else {
          if (nib[2] == best) return in_nibble<2>(best, tie, md, X, Y, Z, rec);
          return in_nibble<3>(best, tie, md, X, Y, Z, rec);
        }
      }
    }
  }
};

__device__ __forceinline__ void red_max_shared_s32(int* addr, int v) {  // one native RED, no compiler-side warp aggregation
  asm volatile("red.relaxed.cta.shared::cta.max.s32 [%0], %1;" ::"r"(smem_u32(addr)), "r"(v) : "memory");
}

#ifdef UPP_ROWS_PROFILE  // debug build (make EXTRA=-DUPP_ROWS_PROFILE): per-phase cycle sums of block 0, printed per warp
#define UPP_PROF(i_) { const long long c_ = clock64(); prof[i_] += c_ - tlast; tlast = c_; }
#else
#define UPP_PROF(i_)
#endif

constexpr int kRowWarps = 8, kRowThreads = kRowWarps * 32, kRowKpt = kPrMaxN / kRowThreads;  // 32 sort keys per thread

// a warp's candidate of a round: 32 bytes, {key, position | tie flag << 31, -, -} {x, y, z, position | tie flag << 31}
struct __align__(16) RowRec {
  int key, pf, pad0, pad1;
  float x, y, z;
  int pf2;
};

template <int P2>
__global__ void __launch_bounds__(kRowThreads, 1)
    fps_rows_kernel(const float* __restrict__ xyz, int N, int RPW, int npow2, int M, int32_t* __restrict__ idx_out,
                    float* __restrict__ centers_out) {
  constexpr int NW = kRowWarps;
  const int CAP = NW * RPW * 64;  // sorted slots (>= N); slot (w, r, lane, h) = (w * RPW + r) * 64 + 2 * lane + h
  extern __shared__ __align__(16) float s_pts[];                      // 3 * CAP floats: the sorted cloud
  unsigned* s_key = reinterpret_cast<unsigned*>(s_pts + 3 * CAP);    // npow2 sort keys
  unsigned short* s_orig = reinterpret_cast<unsigned short*>(s_key + npow2);  // CAP original indices
  __shared__ float s_red[6][NW];
  __shared__ RowRec s_rec[2][NW];
  __shared__ __align__(8) int2 s_max[3];
  __shared__ unsigned s_tie[NW];

  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  // ---- 1. Z-order sort of the cloud (as fps_pruned_kernel) ----
  {
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = t; i < N; i += kRowThreads) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float v = __ldg(p + 3 * i + a);
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = warp_min_f(lo[a]);
      hi[a] = warp_max_f(hi[a]);
      if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float l = s_red[a][0], h = s_red[3 + a][0];
      for (int w = 1; w < NW; ++w) { l = fminf(l, s_red[a][w]); h = fmaxf(h, s_red[3 + a][w]); }
      lo[a] = l;
      scale[a] = h > l ? 63.99f / (h - l) : 0.f;
    }
    for (int i = t; i < npow2; i += kRowThreads) {
      unsigned key = 0xffffffffu;  // padding sorts last
      if (i < N) {
        unsigned q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float f = (__ldg(p + 3 * i + a) - lo[a]) * scale[a];
          q[a] = static_cast<unsigned>(fminf(fmaxf(f, 0.f), 63.f));
        }
        key = (morton18(q[0], q[1], q[2]) << 13) | static_cast<unsigned>(i);
      }
      s_key[i] = key;
    }
    __syncthreads();
    bitonic_sort_cta<unsigned, kRowKpt>(s_key, npow2);
    for (int pos = t; pos < CAP; pos += kRowThreads) {
      float x = 0.f, y = 0.f, z = 0.f;
      unsigned orig = 0;
      if (pos < N) {
        orig = s_key[pos] & 8191u;
        x = __ldg(p + 3 * orig); y = __ldg(p + 3 * orig + 1); z = __ldg(p + 3 * orig + 2);
      }
      s_pts[3 * pos] = x; s_pts[3 * pos + 1] = y; s_pts[3 * pos + 2] = z;
      s_orig[pos] = static_cast<unsigned short>(orig);
    }
    __syncthreads();
  }

  // ---- 2. registers: my two points of every row, their min-distances, the row boxes / bounds (lane r: row r) ----
  f32x2 X[P2], Y[P2], Z[P2];
  float md[2 * P2];
  RowTree<P2> tr;
  float blo[3] = {0.f, 0.f, 0.f}, bhi[3] = {0.f, 0.f, 0.f};
  // lane r: row r's largest min-distance times (1 + 2^-17), rounded up -- the safety margin of the skip test against the
  // rounding of the box distance; -1: nothing left to update (all selected) or nothing selectable
  float thr = -1.0f;
  const int slot0 = warp * RPW * 64 + 2 * lane;
#pragma unroll
  for (int r = 0; r < P2; ++r) {
    float c[2][3];
    float l3[3] = {3.4e38f, 3.4e38f, 3.4e38f}, h3[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pos = slot0 + r * 64 + h;
      if (r < RPW && pos < N) {
        c[h][0] = s_pts[3 * pos]; c[h][1] = s_pts[3 * pos + 1]; c[h][2] = s_pts[3 * pos + 2];
        md[2 * r + h] = fps_initial_md(c[h][0], c[h][1], c[h][2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) { l3[a] = fminf(l3[a], c[h][a]); h3[a] = fmaxf(h3[a], c[h][a]); }
      } else {
        c[h][0] = c[h][1] = c[h][2] = 0.f;
        md[2 * r + h] = kOutOfRange;
      }
    }
    X[r] = pack2(c[0][0], c[1][0]);
    Y[r] = pack2(c[0][1], c[1][1]);
    Z[r] = pack2(c[0][2], c[1][2]);
    tr.l0[r] = max(__float_as_int(md[2 * r]), __float_as_int(md[2 * r + 1]));
    if (r < RPW) {  // warp-uniform
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        l3[a] = warp_min_f(l3[a]);
        h3[a] = warp_max_f(h3[a]);
        if (lane == r) { blo[a] = l3[a]; bhi[a] = h3[a]; }
      }
    }
  }
  {  // a nibble's rows share one bound: the nibble's largest min-distance, with the margin of the skip test
    const int root = tr.build();
    (void)root;
#pragma unroll
    for (int g = 0; g < RowTree<P2>::NG; ++g) {
      const int v = redux_max_s32(tr.nib[g]);
      if ((lane >> 2) == g) thr = v > 0 ? __fmul_ru(__int_as_float(v), 1.0000076294f) : -1.0f;
    }
    if (lane >= RPW) {  // no such row: a box nothing is ever near
#pragma unroll
      for (int a = 0; a < 3; ++a) { blo[a] = 3.0e18f; bhi[a] = -3.0e18f; }
    }
  }

  const unsigned lanes_below = (1u << lane) - 1u;
  // the warp's candidate: one REDUX over the lanes' tree roots; the lowest lane holding it walks its tree down, writes
  // the warp's record and enters the key in the round's two arg-max words (see the resolve step below)
  auto warp_post = [&](int best, RowRec* recs, int2* mx) {
    const int wbest = redux_max_s32(best);
    const unsigned winners = __ballot_sync(0xffffffffu, best == wbest);
    if (best == wbest && (winners & lanes_below) == 0u) {
      red_max_shared_s32(&mx->x, (wbest & ~7) | warp);
      red_max_shared_s32(&mx->y, (wbest & ~7) | (7 - warp));
      bool tie = (winners & (winners - 1u)) != 0u;  // a second lane holds it too
      float4 rc;
      const int sl = tr.find(wbest, tie, md, X, Y, Z, rc);
      int pf = slot0 + (sl >> 1) * 64 + (sl & 1);
      if (tie) pf |= static_cast<int>(0x80000000u);
      int4* dst = reinterpret_cast<int4*>(&recs[warp]);
      dst[0] = make_int4(wbest, pf, 0, 0);
      dst[1] = make_int4(__float_as_int(rc.x), __float_as_int(rc.y), __float_as_int(rc.z), pf);
    }
  };
  if (t < 3) s_max[t] = make_int2(INT_MIN, INT_MIN);
  __syncthreads();
  warp_post(tr.build(), s_rec[0], &s_max[0]);

  // first centre: original index 0 -- its coordinates straight from global memory
  float cx = __ldg(p), cy = __ldg(p + 1), cz = __ldg(p + 2);
  if (t == 0) out[0] = 0;
  __syncthreads();

  const int my_nibble = lane >> 2;
  int nbound[RowTree<P2>::NG];
#pragma unroll
  for (int g = 0; g < RowTree<P2>::NG; ++g) nbound[g] = -1;
#ifdef UPP_ROWS_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
  int nactive = 0, nrows = 0, nties = 0, nslow = 0;
#endif
  int cur = 1;  // j % 3: the round's arg-max words
  for (int j = 1; j < M; ++j) {
    // ---- a. which of my rows can change?  lane r tests row r (lanes >= RPW hold a box at infinity: never).  A row is skipped
    //         only when every point's distance provably exceeds the row's largest min-distance: thr carries the margin ----
    const float ex = fmaxf(fmaxf(blo[0] - cx, cx - bhi[0]), 0.f);
    const float ey = fmaxf(fmaxf(blo[1] - cy, cy - bhi[1]), 0.f);
    const float ez = fmaxf(fmaxf(blo[2] - cz, cz - bhi[2]), 0.f);
    const float dlb = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
    const unsigned mask = __ballot_sync(0xffffffffu, !(dlb > thr));
    RowRec* recs = s_rec[j & 1];
    int2* mx = &s_max[cur];
    UPP_PROF(0)
#ifdef UPP_ROWS_PROFILE
    nactive += mask != 0u; nrows += __popc(mask);
#endif
    if (mask != 0u) {  // warp-uniform
      const f32x2 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
      // Update granularity: NIBBLES of four rows, their chains interleaved (a lone row is a 6-deep dependent chain; rows
      // that did not need it come out unchanged).  Rows due in the same round are Z-curve neighbours: ~1.2 nibbles per
      // active warp.  One code path per nibble keeps the loop inside the instruction cache (the per-row / per-group /
      // all-rows variants of the first cut stalled on instruction fetch at every branch target).
      auto update_nibble = [&](auto gc) {
        constexpr int R0 = decltype(gc)::value, NR = 4;
        f32x2 D[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) D[r] = sub2(Y[R0 + r], CY);
#pragma unroll
        for (int r = 0; r < NR; ++r) D[r] = mul2(D[r], D[r]);
#pragma unroll
        for (int r = 0; r < NR; ++r) { const f32x2 dx = sub2(X[R0 + r], CX); D[r] = fma2(dx, dx, D[r]); }
#pragma unroll
        for (int r = 0; r < NR; ++r) { const f32x2 dz = sub2(Z[R0 + r], CZ); D[r] = fma2(dz, dz, D[r]); }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          float d0, d1;
          unpack2(D[r], d0, d1);
          md[2 * (R0 + r)] = fminf(md[2 * (R0 + r)], d0);
          md[2 * (R0 + r) + 1] = fminf(md[2 * (R0 + r) + 1], d1);
          tr.l0[R0 + r] = max(__float_as_int(md[2 * (R0 + r)]), __float_as_int(md[2 * (R0 + r) + 1]));
        }
        tr.template fold<R0 / 4>();
        nbound[R0 / 4] = redux_max_s32(tr.nib[R0 / 4]);  // the nibble's new bound: consumed after the barrier
      };
#define UPP_NIBBLE(G_) \
  if constexpr ((G_) < P2) if ((mask >> (G_)) & 0xfu) update_nibble(std::integral_constant<int, (G_)>{});
      UPP_NIBBLE(0) UPP_NIBBLE(4) UPP_NIBBLE(8) UPP_NIBBLE(12)
#undef UPP_NIBBLE
      const int best = tr.top();
      UPP_PROF(1)
      warp_post(best, recs, mx);
      UPP_PROF(2)
    } else if (lane < 8) {  // nothing of mine changed: the previous record stands (8 lanes x 4 bytes; lane 0 holds the key)
      const int wd = reinterpret_cast<const int*>(&s_rec[(j & 1) ^ 1][warp])[lane];
      reinterpret_cast<int*>(&recs[warp])[lane] = wd;
      if (lane == 0) {
        red_max_shared_s32(&mx->x, (wd & ~7) | warp);
        red_max_shared_s32(&mx->y, (wd & ~7) | (7 - warp));
      }
    }
    __syncthreads();
    UPP_PROF(3)
    // ---- b. block winner.  The posting lanes entered (key & ~7) | warp and (key & ~7) | (7 - warp) with two native
    //         32-bit shared-memory atomic maxima: when both name the same warp it alone holds the largest key prefix, hence
    //         the largest key -- one LDS.64, one LDS.128 of its record, and the next centre is known.  Two warps within
    //         8 ulp of each other (or a flagged record) take the exact path ----
    const int2 ab = *mx;
    const int wa = ab.x & 7, wb = 7 - (ab.y & 7);
    int4 win = reinterpret_cast<const int4*>(&recs[wa])[1];  // {x, y, z, position | tie flag}
    if (t == 64) s_max[cur == 0 ? 2 : cur - 1] = make_int2(INT_MIN, INT_MIN);  // (j + 2) % 3: last read in round j - 1
    cur = cur == 2 ? 0 : cur + 1;
    // ---- c. the bounds of the nibbles touched: their reductions were issued with the update, the results land here.
    //         A nibble's four rows share one bound (its largest min-distance) but keep their own boxes ----
    if (mask != 0u) {
#define UPP_REFRESH(G_)                                                                              \
  if constexpr ((G_) < P2) if (((mask >> (G_)) & 0xfu) && my_nibble == (G_) / 4)                      \
    thr = nbound[(G_) / 4] > 0 ? __fmul_ru(__int_as_float(nbound[(G_) / 4]), 1.0000076294f) : -1.0f;
      UPP_REFRESH(0) UPP_REFRESH(4) UPP_REFRESH(8) UPP_REFRESH(12)
#undef UPP_REFRESH
    }
    UPP_PROF(4)
    int kbest = 0;
    bool tied = win.w < 0;
    if (wa != wb || tied) {  // block-uniform, rare: resolve on the exact keys
#ifdef UPP_ROWS_PROFILE
      ++nslow;
#endif
      const int kw = lane < NW ? recs[lane].key : INT_MIN;
      kbest = redux_max_s32(kw);
      const unsigned holders = __ballot_sync(0xffffffffu, kw == kbest);
      win = reinterpret_cast<const int4*>(&recs[__ffs(holders) - 1])[1];
      tied = (holders & (holders - 1u)) != 0u || win.w < 0;
    }
    cx = __int_as_float(win.x);
    cy = __int_as_float(win.y);
    cz = __int_as_float(win.z);
    int pos = win.w & 8191;
    if (tied) {  // block-uniform: several points hold the maximum -> the lowest ORIGINAL index among all of them
#ifdef UPP_ROWS_PROFILE
      ++nties;
#endif
      unsigned cand = 0xffffffffu;
#pragma unroll
      for (int s = 0; s < 2 * P2; ++s) {
        if (__float_as_int(md[s]) == kbest) {
          const int ps = slot0 + (s >> 1) * 64 + (s & 1);
          cand = min(cand, (static_cast<unsigned>(s_orig[ps]) << 13) | static_cast<unsigned>(ps));
        }
      }
      cand = redux_min_u32(cand);
      if (lane == 0) s_tie[warp] = cand;
      __syncthreads();
      unsigned m = s_tie[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) m = min(m, s_tie[w]);
      pos = static_cast<int>(m & 8191u);
      cx = s_pts[3 * pos];
      cy = s_pts[3 * pos + 1];
      cz = s_pts[3 * pos + 2];
    }
    if (t == 0) out[j] = pos;  // sorted position for now: a store with nothing to wait for (translated after the loop)
    UPP_PROF(5)
  }
#ifdef UPP_ROWS_PROFILE
  if (b == 0 && lane == 0)
    printf("rows prof warp %d: box %lld update %lld post %lld barrier %lld resolve+refresh %lld tail %lld | active rounds %d rows %d slow %d ties %d of %d rounds\n",
           warp, prof[0] / (M - 1), prof[1] / (M - 1), prof[2] / (M - 1), prof[3] / (M - 1), prof[4] / (M - 1), prof[5] / (M - 1), nactive, nrows, nslow, nties, M - 1);
#endif
  // sorted positions -> original indices, and the centres (the fused gather of utils/misc.py:19), by the whole CTA
  __syncthreads();  // thread 0's out[] stores are visible to the block
  for (int j = 1 + t; j < M; j += kRowThreads) {
    const int pos = out[j];
    out[j] = static_cast<int32_t>(s_orig[pos]);
    if (cen) { cen[3 * j] = s_pts[3 * pos]; cen[3 * j + 1] = s_pts[3 * pos + 1]; cen[3 * j + 2] = s_pts[3 * pos + 2]; }
  }
  if (cen && t == 0) { cen[0] = __ldg(p); cen[1] = __ldg(p + 1); cen[2] = __ldg(p + 2); }
}

template <int P2>
static int launch_rows(const float* xyz, int B, int N, int M, int32_t* idx, float* centers, cudaStream_t st) {
  const int nrows = (N + 63) / 64;
  const int rpw = (nrows + kRowWarps - 1) / kRowWarps;
  const int cap = kRowWarps * rpw * 64;
  int npow2 = 32 * kRowKpt;  // one warp-sorted block at least
  while (npow2 < N) npow2 <<= 1;
  size_t smem = static_cast<size_t>(cap) * 12 + static_cast<size_t>(npow2) * 4 + static_cast<size_t>(cap) * 2 + 16;
  if (B <= 148 && smem < 229376) smem = 229376;  // as every FPS chain kernel: keep throughput CTAs off this SM (fps.cu)
  cudaError_t e = cudaFuncSetAttribute(fps_rows_kernel<P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  fps_rows_kernel<P2><<<B, kRowThreads, smem, st>>>(xyz, N, rpw, npow2, M, idx, centers);
  count_launch();
  return launch_status();
}

int fps_rows_launch(const float* xyz, int B, int N, int M, int32_t* idx, float* centers, cudaStream_t st) {
  if (N > kPrMaxN || N < 64) return UPP_ERR_UNSUPPORTED;
  const int rpw = ((N + 63) / 64 + kRowWarps - 1) / kRowWarps;
  if (rpw <= 4) return launch_rows<4>(xyz, B, N, M, idx, centers, st);
  if (rpw <= 8) return launch_rows<8>(xyz, B, N, M, idx, centers, st);
  if (rpw <= 12) return launch_rows<12>(xyz, B, N, M, idx, centers, st);
  return launch_rows<16>(xyz, B, N, M, idx, centers, st);
}

// UPP_OK when the pruned kernel took the call, UPP_ERR_UNSUPPORTED when the shape is outside it.
int fps_pruned_launch(const float* xyz, int B, int N, int M, int32_t* idx, float* centers, cudaStream_t st) {
  if (N > kPrMaxN || N < 2 * kPrBucket) return UPP_ERR_UNSUPPORTED;
  const int nw = tuning_env_int("UPP_FPS_PRUNED_NW", 8);
  if (nw == 4) return launch_pruned<4>(xyz, B, N, M, idx, centers, st);
  if (nw == 16) return launch_pruned<16>(xyz, B, N, M, idx, centers, st);
  return launch_pruned<8>(xyz, B, N, M, idx, centers, st);
}

}  // namespace upp
