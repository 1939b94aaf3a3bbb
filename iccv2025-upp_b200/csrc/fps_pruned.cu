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

// UPP_OK when the pruned kernel took the call, UPP_ERR_UNSUPPORTED when the shape is outside it.
int fps_pruned_launch(const float* xyz, int B, int N, int M, int32_t* idx, float* centers, cudaStream_t st) {
  if (N > kPrMaxN || N < 2 * kPrBucket) return UPP_ERR_UNSUPPORTED;
  const int nw = tuning_env_int("UPP_FPS_PRUNED_NW", 8);
  if (nw == 4) return launch_pruned<4>(xyz, B, N, M, idx, centers, st);
  if (nw == 16) return launch_pruned<16>(xyz, B, N, M, idx, centers, st);
  return launch_pruned<8>(xyz, B, N, M, idx, centers, st);
}

}  // namespace upp
