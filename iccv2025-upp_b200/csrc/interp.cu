// interp.cu -- k-nearest inverse-distance feature interpolation for sm_100a.
//
// Replaces the body of the reference's pure-torch
//   propagate()                         models/Point_MAE_unify.py:22-48      (k = de_neighbors, eps 1e-8, x0.3 + points1)
//   PointNetFeaturePropagation.forward  models/Point_MAE_unify_segment.py:289-313, models/Point_MAE_pretask_dev.py:437-461
//                                                                            (k = interpolate_neighbors, eps 1e-4)
// which build the full (B,N,S) square_distance matrix, SORT every row, slice k columns, materialise
// index_points(points2, idx) as a (B,N,k,C) tensor, multiply by the weights and sum.
//
// Here, small / narrow problems take ONE forward launch: the S source points of a cloud are staged once per
// CTA by TMA bulk copy, one warp per target point selects its k nearest sources with the register-resident
// warp top-k of topk.cuh on the reference's own distance form (square_distance: -2 a.b + |a|^2 + |b|^2),
// turns them into weights (1/(d+eps), normalised) and immediately gathers the k feature rows (float4,
// coalesced over channels) into the output row -- no distance matrix, no sort, no (B,N,k,C) tensor.
// Backward is deterministic and atomic-free: a target-side kernel (one warp per target) produces
// d loss / d dist and grad_xyz1; a source-side kernel (one CTA per source point) scans the cloud's
// (N,k) index list in order and accumulates grad_points2 / grad_xyz2 rows in registers.
// Wide features (the seg propagation: 2048 <- 128 sources, 1152 channels) are HBM-bound and get their own
// kernels further down: selection + shared-memory blend over a 2-D TMA tile (forward), per-tile CSR + a
// TMA pipeline that streams grad_out exactly once (feature gradient).  Same results, see the dispatch in
// interp_fwd_launch / interp_bwd_launch.
#include "topk.cuh"

namespace upp {

// out[b,n,:] = (base ? base[b,n,:] : 0) + alpha * sum_j w_j feat2[b, idx_j, :]
// Gather: channels in super-blocks of 8 x 128 (one float4 per lane and block, 8 independent LDG.128 in flight per
// neighbour), neighbours four at a time with their weight / row pointer shuffled once per super-block; products and
// sums as packed fp32x2 (written mul.f32x2 then add.f32x2; ptxas contracts each pair to one FFMA2 -- SASS-verified -- so a
// term is rounded once where torch's mul-then-sum rounds twice: the parity tests hold the output to 1e-5 relative).
constexpr int kInterpCB = 8;  // float4 channel blocks per super-block (1024 channels)

// CB x 128 channels starting at cs for one target: neighbours four at a time (weight / row pointer shuffled once),
// CB independent LDG.128 in flight per neighbour.  GUARD: lanes past the last channel are masked (remainder block).
template <int CB, bool GUARD>
__device__ __forceinline__ void interp_gather_blocks(const float* __restrict__ fb, const float* __restrict__ brow,
                                                     float* __restrict__ orow, int C, int cs, int k, float w, int li,
                                                     float alpha) {
  const int lane = threadIdx.x & 31;
  const int c0 = cs + lane * 4;
  f32x2 acc[CB][2];
#pragma unroll
  for (int u = 0; u < CB; ++u) acc[u][0] = acc[u][1] = pack2(0.f, 0.f);
  for (int j0 = 0; j0 < k; j0 += 4) {  // warp-uniform: every lane joins the shuffles
    float wj[4];
    const float* fj[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int jj = min(j0 + v, k - 1);
      wj[v] = __shfl_sync(0xffffffffu, w, jj);
      fj[v] = fb + static_cast<size_t>(__shfl_sync(0xffffffffu, li, jj)) * C + c0;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (j0 + v < k) {  // warp-uniform
        const f32x2 W2 = pack2(wj[v], wj[v]);
#pragma unroll
        for (int u = 0; u < CB; ++u) {
          if (!GUARD || c0 + u * 128 < C) {
            const ulonglong2 f = __ldg(reinterpret_cast<const ulonglong2*>(fj[v] + u * 128));
            acc[u][0] = add2(acc[u][0], mul2(f.x, W2));
            acc[u][1] = add2(acc[u][1], mul2(f.y, W2));
          }
        }
      }
    }
  }
  const f32x2 A2 = pack2(alpha, alpha);
#pragma unroll
  for (int u = 0; u < CB; ++u) {
    const int c = c0 + u * 128;
    if (!GUARD || c < C) {
      ulonglong2 o;
      o.x = mul2(A2, acc[u][0]);
      o.y = mul2(A2, acc[u][1]);
      if (brow) {
        const ulonglong2 p = __ldg(reinterpret_cast<const ulonglong2*>(brow + c));
        o.x = add2(p.x, o.x);
        o.y = add2(p.y, o.y);
      }
      *reinterpret_cast<ulonglong2*>(orow + c) = o;
    }
  }
}

// `tpw` targets per warp: a CTA serves 8 * tpw consecutive targets of one cloud from ONE staging of the sources
// (single-tile clouds, S <= 2048; larger source clouds run tpw = 1 through the tile loop) -- the per-CTA start-up
// (barrier init, TMA round trip) otherwise dominates a warp whose own work is ~1 us.
template <bool VEC4, int SLOTS>
__global__ void __launch_bounds__(kKnnWarps * kWarp, 3)
    interp_fwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                      const float* __restrict__ feat2, const float* __restrict__ base, float alpha, float eps,
                      int N, int S, int C, int k, int tpw, float* __restrict__ out, int32_t* __restrict__ idx_out,
                      float* __restrict__ w_out, float* __restrict__ d_out) {
  extern __shared__ __align__(16) float s_ref[];  // min(S, kKnnTile) * 3 floats
  __shared__ __align__(8) uint64_t s_bar;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int b = blockIdx.y;
  const float* rb = xyz2 + static_cast<size_t>(b) * S * 3;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;
  const bool single_tile = S <= kKnnTile;
  if (single_tile) stage_points(s_ref, rb, S, &s_bar, parity);

  for (int it = 0; it < tpw; ++it) {
  const int n = (blockIdx.x * tpw + it) * kKnnWarps + warp;
  const bool active = n < N;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float* qp = xyz1 + (static_cast<size_t>(b) * N + n) * 3;
    qx = __ldg(qp);
    qy = __ldg(qp + 1);
    qz = __ldg(qp + 2);
  }
  DistExpanded dist;
  dist.set(qx, qy, qz);
  float ld;
  int li;
  if (single_tile) {
    if (SLOTS <= 8 && k <= 8) {  // one block of sources, few neighbours: k warp arg-min rounds
      ld = __int_as_float(0x7f800000);
      li = 0;
      if (active) warp_topk_small<DistExpanded, SLOTS>(s_ref, S, k, dist, ld, li);
    } else {
      TopkState st;
      st.init();
      if (active) warp_topk_tile<DistExpanded, SLOTS>(s_ref, S, 0, k, dist, st);
      ld = st.ld;
      li = st.li;
    }
  } else {
    warp_topk_scan<DistExpanded, SLOTS>(rb, S, k, dist, active, s_ref, &s_bar, parity, ld, li);
  }
  if (!active) continue;

  // weights: dist_recip = 1/(d + eps); weight = dist_recip / sum(dist_recip)  (sum in neighbour order)
  const float r = lane < k ? __fdiv_rn(1.0f, __fadd_rn(ld, eps)) : 0.f;
  float norm = 0.f;
  for (int j = 0; j < k; ++j) norm = __fadd_rn(norm, __shfl_sync(0xffffffffu, r, j));
  const float w = __fdiv_rn(r, norm);
  const size_t row = static_cast<size_t>(b) * N + n;
  if (lane < k) {
    idx_out[row * k + lane] = li;
    w_out[row * k + lane] = w;
    if (d_out) d_out[row * k + lane] = ld;
  }
  if (lane >= k) li = 0;  // only lanes < k are shuffled from; keep the address arithmetic in range

  const float* fb = feat2 + static_cast<size_t>(b) * S * C;
  float* orow = out + row * C;
  const float* brow = base ? base + row * C : nullptr;
  if (VEC4) {
    // full super-blocks of 8 x 128 channels, then the remainder in blocks of 128 (lane guard only there)
    int cs = 0;
    for (; cs + kInterpCB * 128 <= C; cs += kInterpCB * 128)
      interp_gather_blocks<kInterpCB, false>(fb, brow, orow, C, cs, k, w, li, alpha);
    for (; cs < C; cs += 128) interp_gather_blocks<1, true>(fb, brow, orow, C, cs, k, w, li, alpha);
  } else {
    for (int c0 = 0; c0 < C; c0 += 32) {  // warp-uniform trip count (shuffles inside)
      const int c = c0 + lane;
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        const float wj = __shfl_sync(0xffffffffu, w, j);
        const int ij = __shfl_sync(0xffffffffu, li, j);
        if (c < C) acc = __fadd_rn(acc, __fmul_rn(__ldg(fb + static_cast<size_t>(ij) * C + c), wj));
      }
      if (c < C) {
        float o = __fmul_rn(alpha, acc);
        if (brow) o = __fadd_rn(__ldg(brow + c), o);
        orow[c] = o;
      }
    }
  }
  }  // targets of this warp
}

// ---------------------------------------------------------------------------------------------
// Wide features (seg feature propagation: 2048 <- 128 sources, 1152 channels): TWO launches.
// The one-launch kernel above gathers the k source rows from L2 per target and is bound by that
// latency (ncu: 52 % long-scoreboard stalls, 3.0 TB/s of output).  Here
//   1. the selection runs alone: interp_select_kernel (k <= 4: one THREAD per target scanning the
//      staged sources with a branch-free sorted insertion -- the warp-per-target selection costs
//      ~700 issue cycles per target for k = 3, 40 us at the seg shape) or interp_fwd_kernel with
//      C = 0; both write idx / weight / dist only;
//   2. interp_blend_kernel streams the output: a CTA owns (cloud, 128-channel chunk, span of
//      targets), stages its S x 128-channel slice of the source features ONCE in shared memory (one
//      2-D TMA tile load); every warp owns a run of the span's targets, turns their (idx, weight)
//      pairs into {row offset, weight} records in its own slice of shared memory 16 targets at a time
//      (the next batch is in flight in registers while this one is blended) and blends target after
//      target from shared memory: one float4 per lane, per neighbour one broadcast LDS.64 + one
//      LDS.128, one streaming 512-byte store per target.  The only long-latency traffic left is the
//      output write, which is what bounds the op: 50 us for 302 MB at the seg shape, memset 45 us
//      (insensitive to 8 / 16 warps per CTA, to CTA-wide or warp-private batches and to the output
//      stride -- a one-chunk layout with fully linear rows takes the same time).
// Arithmetic is the fused kernel's (products and sums in neighbour order -- ptxas contracts each
// mul.f32x2 + add.f32x2 pair to FFMA2 in both kernels, SASS-verified -- then alpha, then base): bit-identical output.
// ---- explicit shared-window loads (32-bit addresses; see interp_bwd_stream_kernel) ----
__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ ulonglong2 lds_v2u64(uint32_t a) {
  ulonglong2 v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
  return v;
}

// (Programmatic dependent launch of the blend after the selection, and of the streamed backward after the CSR
//  kernel -- griddepcontrol.wait after the prologue / feature tile load -- was built and measured: 68.1 vs 68.6 us
//  forward, 65.0 vs 64.5 us backward, C5 step 0.1956 vs 0.1951 ms.  No gain; ordinary launches kept.)
constexpr int kBlendCh = 128;
constexpr int kBlendMaxS = 192;   // 96 KB of staged features: two CTAs per SM (three up to S = 136)

// Selection for k = K <= 4 without a warp per target: FOUR threads per target (adjacent lanes), each scanning one
// contiguous quarter of the staged sources (single tile, S <= 1024) with a branch-free sorted insertion, then two
// shuffle-xor merge rounds.  Sources are visited in index order, a candidate only moves ahead of strictly larger
// distances, and on equal distances a merge keeps the entry of the lower quarter first: equal distances keep the
// lower source index, the order of the warp kernels (a stable sort by distance).  Distances are DistExpanded's,
// operation for operation (|b|^2 is computed once per source while staging).  Quarter q lives at s4[q * (Q + 1) ...]:
// the one-slot skew keeps the four broadcast LDS.128 of a warp in different banks.
template <int K>
__device__ __forceinline__ void topk_insert(float (&bd)[K], int (&bi)[K], float d, int s, bool ties_first) {
  bool lt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) lt[j] = ties_first ? d <= bd[j] : d < bd[j];
#pragma unroll
  for (int j = K - 1; j >= 0; --j) {
    if (j > 0 && lt[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
    else if (lt[j]) { bd[j] = d; bi[j] = s; }
  }
}

constexpr int kSelThreads = 128;

// TPT = threads per target (1, 2 or 4): the sources are cut into TPT contiguous parts.
template <int K, int TPT>
__global__ void __launch_bounds__(kSelThreads)
    interp_select_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, float eps, int N, int S,
                         int32_t* __restrict__ idx_out, float* __restrict__ w_out, float* __restrict__ d_out) {
  extern __shared__ __align__(16) float s_sel[];  // [TPT x (Q + 1) x {x, y, z, |b|^2}] then the AoS staging area (3 S floats)
  __shared__ __align__(8) uint64_t s_bar;
  const int Q = (S + TPT - 1) / TPT;
  float4* s4 = reinterpret_cast<float4*>(s_sel);
  float* s_ref = s_sel + 4 * TPT * (Q + 1);
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;
  stage_points(s_ref, xyz2 + static_cast<size_t>(b) * S * 3, S, &s_bar, parity);
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float x = s_ref[3 * s], y = s_ref[3 * s + 1], z = s_ref[3 * s + 2];
    const int q = s / Q;
    s4[q * (Q + 1) + (s - q * Q)] =
        make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  }
  __syncthreads();
  const int q = threadIdx.x % TPT;
  const int n = blockIdx.x * (kSelThreads / TPT) + threadIdx.x / TPT;
  const bool active = n < N;
  const size_t row = static_cast<size_t>(b) * N + (active ? n : N - 1);
  const float qx = __ldg(xyz1 + row * 3), qy = __ldg(xyz1 + row * 3 + 1), qz = __ldg(xyz1 + row * 3 + 2);
  const float s1 = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
  float bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { bd[j] = __int_as_float(0x7f800000); bi[j] = 0; }
  const int sbeg = q * Q;
  const int cnt = min(Q, S - sbeg);  // <= 0 for an empty quarter
  const float4* sq = s4 + q * (Q + 1);
#pragma unroll 4
  for (int j = 0; j < cnt; ++j) {
    const float4 r = sq[j];
    const float dot = __fmaf_rn(qz, r.z, __fmaf_rn(qy, r.y, __fmul_rn(qx, r.x)));
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), s1), r.w);
    topk_insert<K>(bd, bi, d, sbeg + j, false);
  }
#pragma unroll
  for (int x = 1; x < TPT; x <<= 1) {  // merge with the partner part's sorted list; the lower part wins ties
    float pd[K];
    int pi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      pd[j] = __shfl_xor_sync(0xffffffffu, bd[j], x);
      pi[j] = __shfl_xor_sync(0xffffffffu, bi[j], x);
    }
    const bool partner_lower = (q & x) != 0;
    // a lower-indexed partner's entries go BEFORE equal distances: inserted last to first, so that two of its own
    // entries with equal distance keep their (index) order; a higher-indexed partner's go after equals, first to last
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int jj = partner_lower ? K - 1 - j : j;
      float pdj = pd[0];
      int pij = pi[0];
#pragma unroll
      for (int u = 1; u < K; ++u)
        if (u == jj) { pdj = pd[u]; pij = pi[u]; }
      topk_insert<K>(bd, bi, pdj, pij, partner_lower);
    }
  }
  if (!active || q != 0) return;
  // weights: dist_recip = 1/(d + eps); weight = dist_recip / sum(dist_recip)  (sum in neighbour order)
  float r[K], norm = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    r[j] = __fdiv_rn(1.0f, __fadd_rn(bd[j], eps));
    norm = __fadd_rn(norm, r[j]);
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    idx_out[row * K + j] = bi[j];
    w_out[row * K + j] = __fdiv_rn(r[j], norm);
    if (d_out) d_out[row * K + j] = bd[j];
  }
}

// K = 0: any k <= 32 (neighbour loop not unrolled, no register prefetch of the next record batch).
// NWB warps per CTA (8 or 16: the staged feature slice caps the CTAs per SM at three, so warps per CTA set the occupancy).
// A warp owns a CONTIGUOUS run of the CTA's targets and keeps its own record batches (TW targets at a time) in a
// private slice of shared memory: no CTA barrier after the feature tile has landed (with CTA-wide batches the two
// barriers per batch were 18 % of the stall samples at 16 warps).
__host__ __device__ constexpr int blend_tw(int K) { return K == 0 ? 4 : (K <= 4 ? 16 : 8); }  // targets per warp batch

template <int K, int NWB>
__global__ void __launch_bounds__(NWB * kWarp)
    interp_blend_kernel(const __grid_constant__ CUtensorMap fmap, const float* __restrict__ base, float alpha,
                        const int32_t* __restrict__ idx, const float* __restrict__ weight, int N, int S, int C,
                        int krt, int span, float* __restrict__ out) {
  constexpr int TW = blend_tw(K);
  // bytes of one staged feature row: 512 for 128-channel chunks, 4 C when the whole (narrow, C < 128) row is one chunk
  const unsigned rowb = static_cast<unsigned>(min(C, kBlendCh)) * 4u;
  extern __shared__ __align__(128) unsigned char s_blend[];  // [S x 512 B features][NWB x TW * k records of 8 B]
  __shared__ __align__(8) uint64_t s_bar;
  const int k = K > 0 ? K : krt;
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int chunk = blockIdx.x, b = blockIdx.z;
  const int n0 = blockIdx.y * span;
  const int n1 = min(N, n0 + span);
  // this warp's run of targets [w0, w1): whole pairs, so that a trip's two targets are neighbours
  const int per_warp = (((n1 - n0) + NWB - 1) / NWB + 1) & ~1;
  const int w0 = min(n1, n0 + warp * per_warp);
  const int w1 = min(n1, w0 + per_warp);
  uint2* s_rec = reinterpret_cast<uint2*>(s_blend + ((static_cast<size_t>(S) * rowb + 15) & ~static_cast<size_t>(15))) + warp * (TW * k);

  if (t == 0) {
    tma_prefetch_desc(&fmap);
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, static_cast<unsigned>(S) * rowb);
    tma_load_2d(s_blend, &fmap, chunk * kBlendCh, b * S, &s_bar);
  }
  // records of one batch: entry e <-> flat (target, j) position, contiguous in idx / weight
  constexpr int PE = K > 0 ? (TW * K + kWarp - 1) / kWarp : 1;
  uint2 pre[PE];
  auto fetch = [&](int s0) {  // K > 0 only: the batch's records into registers
    const int cnt = (min(w1, s0 + TW) - s0) * k;
    const size_t p0 = (static_cast<size_t>(b) * N + s0) * k;
#pragma unroll
    for (int u = 0; u < PE; ++u) {
      const int e = lane + u * kWarp;
      pre[u] = e < cnt ? make_uint2(static_cast<unsigned>(__ldg(idx + p0 + e)) * rowb,
                                    __float_as_uint(__ldg(weight + p0 + e)))
                       : make_uint2(0u, 0u);
    }
  };
  auto commit = [&]() {
#pragma unroll
    for (int u = 0; u < PE; ++u) {
      const int e = lane + u * kWarp;
      if (e < TW * k) s_rec[e] = pre[u];
    }
  };
  if (K > 0 && w0 < w1) {
    fetch(w0);
    commit();
  }
  __syncthreads();  // barrier init visible to every thread (and this warp's first record batch in place)
  mbar_wait(&s_bar, 0);

  const uint32_t feat = smem_u32(s_blend) + lane * 16u;
  const f32x2 A2 = pack2(alpha, alpha);
  const int col = chunk * kBlendCh + lane * 4;
  const bool lane_on = lane * 16u < rowb;  // narrow rows: the upper lanes compute on whatever follows and store nothing
  for (int s0 = w0; s0 < w1; s0 += TW) {
    const int cnt = min(w1, s0 + TW) - s0;
    if (K > 0) {
      if (s0 + TW < w1) fetch(s0 + TW);  // next batch: loads in flight under this batch's blend
    } else {
      __syncwarp();
      for (int e = lane; e < cnt * k; e += kWarp) {
        const size_t p = (static_cast<size_t>(b) * N + s0) * k + e;
        s_rec[e] = make_uint2(static_cast<unsigned>(__ldg(idx + p)) * rowb, __float_as_uint(__ldg(weight + p)));
      }
      __syncwarp();
    }
    // two targets per trip: 2k independent LDS.128 in flight
    for (int tl = 0; tl < cnt; tl += 2) {
      const bool two = tl + 1 < cnt;
      const int tb = two ? tl + 1 : tl;
      const uint2* ra = s_rec + tl * k;
      const uint2* rb = s_rec + tb * k;
      f32x2 a0 = pack2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
      auto neighbour = [&](int j) {
        const uint2 ea = ra[j], eb = rb[j];
        const ulonglong2 fa = lds_v2u64(feat + ea.x), fb = lds_v2u64(feat + eb.x);
        const f32x2 WA = pack2(__uint_as_float(ea.y), __uint_as_float(ea.y));
        const f32x2 WB = pack2(__uint_as_float(eb.y), __uint_as_float(eb.y));
        a0 = add2(a0, mul2(fa.x, WA));
        a1 = add2(a1, mul2(fa.y, WA));
        b0 = add2(b0, mul2(fb.x, WB));
        b1 = add2(b1, mul2(fb.y, WB));
      };
      if constexpr (K > 0) {
#pragma unroll
        for (int j = 0; j < K; ++j) neighbour(j);
      } else {
        for (int j = 0; j < k; ++j) neighbour(j);
      }
      ulonglong2 oa, ob;
      oa.x = mul2(A2, a0);
      oa.y = mul2(A2, a1);
      ob.x = mul2(A2, b0);
      ob.y = mul2(A2, b1);
      const size_t rowa = static_cast<size_t>(b) * N + s0 + tl;
      const size_t rowb2 = static_cast<size_t>(b) * N + s0 + tb;
      if (base && lane_on) {
        const ulonglong2 pa = __ldg(reinterpret_cast<const ulonglong2*>(base + rowa * C + col));
        const ulonglong2 pb = __ldg(reinterpret_cast<const ulonglong2*>(base + rowb2 * C + col));
        oa.x = add2(pa.x, oa.x);
        oa.y = add2(pa.y, oa.y);
        ob.x = add2(pb.x, ob.x);
        ob.y = add2(pb.y, ob.y);
      }
      if (lane_on) {
        __stcs(reinterpret_cast<ulonglong2*>(out + rowa * C + col), oa);
        if (two) __stcs(reinterpret_cast<ulonglong2*>(out + rowb2 * C + col), ob);
      }
    }
    if (K > 0 && s0 + TW < w1) {
      __syncwarp();  // every lane is done with this batch's records
      commit();
      __syncwarp();
    }
  }
}

// Backward, target side (only when coordinate gradients are wanted).  With G = alpha * grad_out[b,n,:],
// dot_j = <G, f_j>, m = sum_j w_j dot_j, r_j = 1/(d_j + eps):
//   d loss / d d_j = -(r_j w_j)(dot_j - m)                      (w_j = r_j / sum r  =>  r_j^2 / sum r = r_j w_j)
//   grad_xyz1[b,n] = sum_j (d loss / d d_j) * 2 (x1 - x2_j)      (d/dx1 of -2 x1.x2 + |x1|^2 + |x2|^2)
// gd_out (B,N,k) carries d loss / d d_j to the source-side kernel.  One warp per target; the k dot products
// run four neighbours at a time over float4 loads (the grad_out row is read once per group of four).
template <bool VEC4>
__global__ void __launch_bounds__(256)
    interp_bwd_target_kernel(const float* __restrict__ gout, const float* __restrict__ feat2,
                             const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                             const int32_t* __restrict__ idx, const float* __restrict__ weight,
                             const float* __restrict__ distk, float alpha, float eps, int N, int S, int C, int k,
                             size_t rows, float* __restrict__ gd_out, float* __restrict__ gxyz1) {
  const int lane = threadIdx.x & 31;
  for (size_t row = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += static_cast<size_t>(gridDim.x) * (blockDim.x >> 5)) {
    const size_t b = row / N;
    const float* g = gout + row * C;
    const float* fb = feat2 + b * S * C;
    int ij = 0;
    float w = 0.f, d = 0.f;
    if (lane < k) {
      ij = idx[row * k + lane];
      w = weight[row * k + lane];
      d = distk[row * k + lane];
    }
    float mydot = 0.f;
    for (int j0 = 0; j0 < k; j0 += 4) {
      const float* f[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) f[v] = fb + static_cast<size_t>(__shfl_sync(0xffffffffu, ij, min(j0 + v, k - 1))) * C;
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      if (VEC4) {
        for (int c = lane * 4; c < C; c += 128) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(g + c));
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 fv = __ldg(reinterpret_cast<const float4*>(f[v] + c));
            part[v] = __fmaf_rn(gv.w, fv.w, __fmaf_rn(gv.z, fv.z, __fmaf_rn(gv.y, fv.y, __fmaf_rn(gv.x, fv.x, part[v]))));
          }
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const float gv = __ldg(g + c);
#pragma unroll
          for (int v = 0; v < 4; ++v) part[v] = __fmaf_rn(gv, __ldg(f[v] + c), part[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float tot = warp_sum(part[v]);
        if (lane == j0 + v) mydot = __fmul_rn(alpha, tot);
      }
    }
    const float m = warp_sum(lane < k ? __fmul_rn(w, mydot) : 0.f);
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (lane < k) {
      const float r = __fdiv_rn(1.0f, __fadd_rn(d, eps));
      const float gd = -__fmul_rn(__fmul_rn(r, w), mydot - m);
      gd_out[row * k + lane] = gd;
      const float* p1 = xyz1 + row * 3;
      const float* p2 = xyz2 + (b * S + ij) * 3;
      const float t = __fmul_rn(2.0f, gd);
      vx = __fmul_rn(t, __ldg(p1) - __ldg(p2));
      vy = __fmul_rn(t, __ldg(p1 + 1) - __ldg(p2 + 1));
      vz = __fmul_rn(t, __ldg(p1 + 2) - __ldg(p2 + 2));
    }
    vx = warp_sum(vx);
    vy = warp_sum(vy);
    vz = warp_sum(vz);
    if (lane == 0 && gxyz1) {
      gxyz1[row * 3] = vx;
      gxyz1[row * 3 + 1] = vy;
      gxyz1[row * 3 + 2] = vz;
    }
  }
}

// Backward, source side: one CTA per source point (b, s).  The cloud's (N*k) neighbour list is scanned in
// order (each warp a contiguous segment, matches compacted by ballot); the matches (n, j) are then dealt
// round-robin to G thread groups of CG = 256 / G threads (G = 1 for C > 128 ... 8 for C <= 32), each group adding
// w * grad_out[b,n,:] for its matches to register accumulators, four matches in flight at a time; the G partial
// rows are combined through shared memory in group order.  Fixed assignment + fixed order: deterministic, no
// atomics (the scatter-add formulation would issue B*N*k*C float atomics).  Consecutive CTAs belong to the same
// cloud, whose grad_out rows (read k times in total) stay in L2.
constexpr int kSrcThreads = 256;
constexpr int kSrcChunk = 2048;  // list entries per round: 8 warps x 256
constexpr int kSrcAcc = 8;       // channels per thread and pass

template <int G>
__global__ void __launch_bounds__(kSrcThreads)
    interp_bwd_source_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                             const float* __restrict__ weight, const float* __restrict__ gd,
                             const float* __restrict__ xyz1, const float* __restrict__ xyz2, float alpha, int N,
                             int S, int C, int k, float* __restrict__ gfeat2, float* __restrict__ gxyz2) {
  constexpr int CG = kSrcThreads / G;          // threads per group
  constexpr int kPass = kSrcAcc * CG;          // channels per pass
  constexpr int kSeg = kSrcChunk / (kSrcThreads / 32);
  __shared__ int s_list[kSrcChunk];            // this round's matches, in list order
  __shared__ int s_cnt[kSrcThreads / 32 + 1];
  __shared__ float s_part[G > 1 ? kSrcThreads * kSrcAcc : 1];
  __shared__ float s_xyz[kSrcThreads / 32][3];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int grp = t / CG, tc = t % CG;
  const int s = blockIdx.x, b = blockIdx.y;
  const int L = N * k;
  const int32_t* il = idx + static_cast<size_t>(b) * L;
  const float* wl = weight + static_cast<size_t>(b) * L;
  const float* gl = gd ? gd + static_cast<size_t>(b) * L : nullptr;
  const float* gb = gout + static_cast<size_t>(b) * N * C;
  const unsigned below = (1u << lane) - 1u;
  // coordinate gradient: lanes 0..2 of every warp take the matches e = warp (mod 8) of each round
  const float x2c = (gl && lane < 3) ? __ldg(xyz2 + (static_cast<size_t>(b) * S + s) * 3 + lane) : 0.f;
  float gx2 = 0.f;

  for (int cpass = 0; cpass < C || cpass == 0; cpass += kPass) {
    float acc[kSrcAcc];
#pragma unroll
    for (int u = 0; u < kSrcAcc; ++u) acc[u] = 0.f;
    for (int chunk = 0; chunk < L; chunk += kSrcChunk) {
      // ---- ordered compaction of this round's matches ----
      unsigned masks[kSeg / 32];
      int cnt = 0;
#pragma unroll
      for (int it = 0; it < kSeg / 32; ++it) {
        const int p = chunk + warp * kSeg + it * 32 + lane;
        masks[it] = __ballot_sync(0xffffffffu, p < L && __ldg(il + p) == s);
        cnt += __popc(masks[it]);
      }
      if (lane == 0) s_cnt[warp] = cnt;
      __syncthreads();
      int off = 0, total = 0;
#pragma unroll
      for (int wi = 0; wi < kSrcThreads / 32; ++wi) {
        const int cw = s_cnt[wi];
        off += wi < warp ? cw : 0;
        total += cw;
      }
#pragma unroll
      for (int it = 0; it < kSeg / 32; ++it) {
        if (masks[it] >> lane & 1u) s_list[off + __popc(masks[it] & below)] = chunk + warp * kSeg + it * 32 + lane;
        off += __popc(masks[it]);
      }
      __syncthreads();
      // ---- accumulate: group `grp` takes matches grp, grp + G, ...; four in flight ----
      for (int e0 = grp; e0 < total; e0 += 4 * G) {
        int pn[4];
        float wt[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int e = e0 + v * G;
          const int p = e < total ? s_list[e] : -1;
          pn[v] = p < 0 ? -1 : p / k;
          wt[v] = p < 0 ? 0.f : __ldg(wl + p);
        }
        float gv[4][kSrcAcc];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int u = 0; u < kSrcAcc; ++u) {
            const int c = cpass + tc + u * CG;
            gv[v][u] = (pn[v] >= 0 && c < C) ? __ldg(gb + static_cast<size_t>(pn[v]) * C + c) : 0.f;
          }
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int u = 0; u < kSrcAcc; ++u) acc[u] = __fmaf_rn(wt[v], gv[v][u], acc[u]);
      }
      if (gl && cpass == 0 && lane < 3) {
        for (int e = warp; e < total; e += kSrcThreads / 32) {
          const int p = s_list[e];
          const float x1c = __ldg(xyz1 + (static_cast<size_t>(b) * N + p / k) * 3 + lane);
          gx2 = __fmaf_rn(-__fmul_rn(2.0f, __ldg(gl + p)), x1c - x2c, gx2);
        }
      }
      __syncthreads();
    }
    // ---- combine the G partial rows in group order, scale, store ----
    float* orow = gfeat2 + (static_cast<size_t>(b) * S + s) * C;
    if (G > 1) {
#pragma unroll
      for (int u = 0; u < kSrcAcc; ++u) s_part[(grp * kSrcAcc + u) * CG + tc] = acc[u];
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int u = 0; u < kSrcAcc; ++u) {
          float tot = acc[u];
          for (int gi = 1; gi < G; ++gi) tot += s_part[(gi * kSrcAcc + u) * CG + tc];
          const int c = cpass + tc + u * CG;
          if (c < C) orow[c] = __fmul_rn(alpha, tot);
        }
      }
      __syncthreads();
    } else {
#pragma unroll
      for (int u = 0; u < kSrcAcc; ++u) {
        const int c = cpass + tc + u * CG;
        if (c < C) orow[c] = __fmul_rn(alpha, acc[u]);
      }
    }
  }
  if (gl && gxyz2) {  // eight per-warp partial sums of the coordinate gradient, added in warp order
    if (lane < 3) s_xyz[warp][lane] = gx2;
    __syncthreads();
    if (t < 3) {
      float tot = 0.f;
      for (int wi = 0; wi < kSrcThreads / 32; ++wi) tot += s_xyz[wi][t];
      gxyz2[(static_cast<size_t>(b) * S + s) * 3 + t] = tot;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward for wide features, feature gradient only (no coordinate terms): STREAM grad_out once.
// interp_bwd_source_kernel reads every grad_out row k times through L2 (one CTA per source point
// gathering its matches: 207 us at the seg shape, 1.5 TB/s).  Here a CTA owns (cloud, 128-channel
// chunk) and pulls that chunk of ALL N grad_out rows through a 3-stage TMA pipeline exactly once,
// 64 targets per tile; warp w keeps the accumulators of sources 16w .. 16w+15 in registers (one float4
// per lane and source) and adds weight * row for the tile's matches of each of its sources, in list
// order.  Which matches those are comes from a per-tile CSR block built once per call (and shared by
// all channel chunks) by interp_csr_kernel: a stable counting sort of the tile's (target, j) entries
// by source.  Per source the additions happen in ascending (n, j) order, tile after tile: the same
// sequence of fmaf as the source-side kernel's -- deterministic, no atomics, bit-identical to it.
constexpr int kBsTile = 64;                 // targets per streamed tile
constexpr int kBsSrc = 128;                 // sources covered: 8 warps x 16 register accumulators
constexpr int kBsWarps = 8;                 // consumer warps (+ 1 producer warp)
constexpr int kBsStages = 3;
constexpr int kBsMaxK = 8;                  // streamed kernel: entries per tile kBsTile * k <= 512
constexpr int kBsMaxKNarrow = 16;           // CSR gather kernel (C <= 128): entries per tile <= 1024
constexpr unsigned kBsOffBytes = 272;       // 129 uint16 offsets, padded to a multiple of 16 B
constexpr unsigned kBsRowBytes = kBlendCh * 4;
constexpr unsigned kBsStagePad = 16;         // the entry prefetch reads one slot past the last entry

// CSR block of one (cloud, tile): [uint16 off[129] (pad)][E entries of 8 B][E bytes: neighbour slot j of each entry],
// E = kBsTile * k.  The streamed kernel copies the first two parts only; the j bytes serve the coordinate gradient.
__host__ __device__ inline unsigned bs_block_copy_bytes(int k) { return kBsOffBytes + static_cast<unsigned>(kBsTile) * k * 8u; }
__host__ __device__ inline unsigned bs_block_bytes(int k) { return bs_block_copy_bytes(k) + static_cast<unsigned>(kBsTile) * k; }

// stage = [tile][CSR block][pad for the entry prefetch], rounded to the 128-byte alignment TMA tile loads need
__host__ __device__ inline unsigned bs_stage_bytes(int k) {
  return (kBsTile * kBsRowBytes + bs_block_copy_bytes(k) + kBsStagePad + 127u) & ~127u;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One warp per (cloud, tile): block = [uint16 off[129] (pad)][entries sorted by (source, list position)],
// entry = {byte offset of the target's row inside the staged tile, weight bits}.
template <int kRounds>  // 32-entry rounds held in registers: 16 for k <= 8, 32 for k <= 16
__global__ void __launch_bounds__(256)
    interp_csr_kernel(const int32_t* __restrict__ idx, const float* __restrict__ weight, int B, int N, int k,
                      int tiles, unsigned char* __restrict__ csr) {
  __shared__ int s_cnt[8][kBsSrc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long wg = static_cast<long>(blockIdx.x) * 8 + warp;
  if (wg >= static_cast<long>(B) * tiles) return;  // whole warps leave; no CTA barrier below
  const int b = static_cast<int>(wg / tiles), tile = static_cast<int>(wg % tiles);
  int* cnt = s_cnt[warp];
#pragma unroll
  for (int i = 0; i < kBsSrc / 32; ++i) cnt[lane + 32 * i] = 0;
  __syncwarp();
  const int nt = min(kBsTile, N - tile * kBsTile);
  const int E = nt * k;
  const size_t p0 = (static_cast<size_t>(b) * N + static_cast<size_t>(tile) * kBsTile) * k;
  // the tile's whole list in registers first (one exposed memory latency instead of one per round)
  int si[kRounds];
  float sw[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int e = r * 32 + lane;
    si[r] = (r * 32 < E && e < E) ? __ldg(idx + p0 + e) : -1 - lane;  // distinct negatives: invalid lanes match nobody
    sw[r] = (r * 32 < E && e < E) ? __ldg(weight + p0 + e) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < kRounds; ++r)
    if (r * 32 < E && si[r] >= 0) atomicAdd(&cnt[si[r]], 1);
  __syncwarp();
  // exclusive prefix over the 128 sources: four consecutive counts per lane + a warp scan
  int c[4], tot = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i] = cnt[4 * lane + i]; tot += c[i]; }
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  unsigned char* blk = csr + static_cast<size_t>(wg) * bs_block_bytes(k);
  uint16_t* off = reinterpret_cast<uint16_t*>(blk);
  uint2* ent = reinterpret_cast<uint2*>(blk + kBsOffBytes);
  unsigned char* jb = blk + bs_block_copy_bytes(k);
  int run = incl - tot;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    off[4 * lane + i] = static_cast<uint16_t>(run);
    cnt[4 * lane + i] = run;  // becomes the placement cursor
    run += c[i];
  }
  if (lane == 31) off[kBsSrc] = static_cast<uint16_t>(run);
  __syncwarp();
  const unsigned below = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {  // rounds in list order, lanes in list order: a stable sort
    if (r * 32 < E) {                  // warp-uniform
      const int e = r * 32 + lane;
      const int s = si[r];
      const bool valid = s >= 0;
      const unsigned m = __match_any_sync(0xffffffffu, s);
      const int rank = __popc(m & below);
      const int pos = valid ? cnt[s] + rank : 0;
      __syncwarp();
      if (valid && rank == 0) cnt[s] += __popc(m);
      __syncwarp();
      if (valid) {
        ent[pos] = make_uint2(static_cast<unsigned>(e / k) * kBsRowBytes, __float_as_uint(sw[r]));
        jb[pos] = static_cast<unsigned char>(e % k);
      }
    }
  }
}

__global__ void __launch_bounds__((kBsWarps + 1) * kWarp)
    interp_bwd_stream_kernel(const __grid_constant__ CUtensorMap gmap, const unsigned char* __restrict__ csr, float alpha,
                             int N, int S, int C, int k, int tiles, float* __restrict__ gfeat2) {
  extern __shared__ __align__(128) unsigned char s_stage[];  // kBsStages x [tile 64 x 512 B][CSR block]
  __shared__ __align__(8) uint64_t s_full[kBsStages], s_empty[kBsStages];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int chunk = blockIdx.x, b = blockIdx.y;
  const unsigned blk_bytes = bs_block_copy_bytes(k);  // what a stage holds of a block
  const unsigned blk_stride = bs_block_bytes(k);      // distance between blocks in the workspace
  const unsigned stage_bytes = bs_stage_bytes(k);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kBsStages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kBsWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kBsWarps) {  // ---- producer: one 2-D TMA tile (64 rows x 512 B) + the tile's CSR block per stage ----
    if (lane == 0) {
      tma_prefetch_desc(&gmap);
      const unsigned char* csrc = csr + static_cast<size_t>(b) * tiles * blk_stride;
      for (int t = 0; t < tiles; ++t) {
        const int st = t % kBsStages;
        unsigned char* dst = s_stage + static_cast<size_t>(st) * stage_bytes;
        if (t >= kBsStages) mbar_wait(&s_empty[st], static_cast<unsigned>(t / kBsStages - 1) & 1u);
        // rows past this cloud's N (next cloud's, or zero fill past the tensor) are loaded but never referenced
        mbar_expect_tx(&s_full[st], kBsTile * kBsRowBytes + blk_bytes);
        tma_load_2d(dst, &gmap, chunk * kBlendCh, b * N + t * kBsTile, &s_full[st]);
        bulk_g2s(dst + kBsTile * kBsRowBytes, csrc + static_cast<size_t>(t) * blk_stride, blk_bytes, &s_full[st]);
      }
    }
    return;
  }

  // ---- consumer warps ----
  // Shared memory is addressed with explicit 32-bit shared-window addresses (ld.shared): the generic-pointer form made
  // the compiler rebuild the stage address in front of every one of the 16 per-source loops.
  f32x2 acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = pack2(0.f, 0.f);
  const uint32_t stage0 = smem_u32(s_stage);
  for (int t = 0; t < tiles; ++t) {
    const int st = t % kBsStages;
    const uint32_t tile = stage0 + static_cast<uint32_t>(st) * stage_bytes;
    const uint32_t blk = tile + kBsTile * kBsRowBytes;
    mbar_wait(&s_full[st], static_cast<unsigned>(t / kBsStages) & 1u);
    // this warp's 17 offsets (warp-uniform): two LDS.128 + one LDS.U16
    const uint4 oa = lds_u128(blk + 32u * warp);
    const uint4 ob = lds_u128(blk + 32u * warp + 16u);
    const unsigned olast = lds_u16(blk + 32u * warp + 32u);
    const unsigned ow[9] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w, olast};
    uint32_t ent = blk + kBsOffBytes;
    uint32_t rowp = tile + lane * 16u;
    asm volatile("" : "+r"(ent), "+r"(rowp));  // opaque: kept in registers, not rebuilt from %tid / the shared window per source
    // one cursor over the warp's contiguous entry range; the next entry is always already in flight.  (Also keeping
    // the next grad_out row in flight -- two entries + one row ahead -- was measured slower: 66 vs 56 us, the loop is
    // issue-bound, not latency-bound, and the deeper pipeline costs three more instructions per match.)
    unsigned m = ow[0] & 0xffffu;
    uint2 en = lds_u64(ent + m * 8u);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const unsigned end = (i & 1) ? (ow[(i >> 1) + 1] & 0xffffu) : (ow[i >> 1] >> 16);
#pragma unroll 1
      for (; m < end; ++m) {
        const uint2 cur = en;
        en = lds_u64(ent + (m + 1u) * 8u);  // may be the slot after the last entry: the stage is padded for it
        const ulonglong2 g = lds_v2u64(rowp + cur.x);
        const f32x2 W2 = pack2(__uint_as_float(cur.y), __uint_as_float(cur.y));
        acc[i][0] = fma2(W2, g.x, acc[i][0]);  // per half == fmaf(w, g, acc): the source-side kernel's arithmetic
        acc[i][1] = fma2(W2, g.y, acc[i][1]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[st]);
  }
  const f32x2 A2 = pack2(alpha, alpha);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int s = 16 * warp + i;
    if (s < S) {
      ulonglong2 o;
      o.x = mul2(A2, acc[i][0]);
      o.y = mul2(A2, acc[i][1]);
      *reinterpret_cast<ulonglong2*>(gfeat2 + (static_cast<size_t>(b) * S + s) * C + chunk * kBlendCh + lane * 4) = o;
    }
  }
}

// Feature gradient for NARROW features (C <= 128, k <= 16; the rectify-prompter propagation 1096 <- 32 sources, 96
// channels, k = 16) from the same CSR: a CTA per (cloud, source), warp w walks tiles w, w + NWG, ... and adds
// weight * grad_out row (one float4 per lane, up to four rows in flight) for the source's entries of each tile in
// list order; the NWG partial rows are combined through shared memory in warp order.  No list scan (the source-side
// kernel compacts the cloud's N * k entries once per source: 116 us at that shape), deterministic, no atomics.
constexpr int kGatherWarps = 8;

__global__ void __launch_bounds__(kGatherWarps * kWarp)
    interp_bwd_gather_csr_kernel(const float* __restrict__ gout, const unsigned char* __restrict__ csr, float alpha, int N,
                                 int S, int C, int k, int tiles, float* __restrict__ gfeat2) {
  __shared__ float4 s_part[kGatherWarps][kWarp];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x, b = blockIdx.y;
  const bool lane_on = lane * 4 < C;
  const unsigned stride = bs_block_bytes(k);
  const float* gb = gout + static_cast<size_t>(b) * N * C + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = warp; t < tiles; t += kGatherWarps) {
    const unsigned char* blk = csr + (static_cast<size_t>(b) * tiles + t) * stride;
    const uint16_t* off = reinterpret_cast<const uint16_t*>(blk);
    const uint2* ent = reinterpret_cast<const uint2*>(blk + kBsOffBytes);
    const int beg = off[s], end = off[s + 1];
    for (int m0 = beg; m0 < end; m0 += 4) {
      float w[4];
      float4 g[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int m = m0 + v;
        const uint2 e = m < end ? ent[m] : make_uint2(0u, 0u);  // weight 0 past the end: adds +0 to the row of target 0
        w[v] = __uint_as_float(e.y);
        const int n = t * kBsTile + static_cast<int>(e.x / kBsRowBytes);
        g[v] = (lane_on && m < end) ? __ldg(reinterpret_cast<const float4*>(gb + static_cast<size_t>(n) * C))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (m0 + v < end) {  // warp-uniform
          acc.x = __fmaf_rn(w[v], g[v].x, acc.x);
          acc.y = __fmaf_rn(w[v], g[v].y, acc.y);
          acc.z = __fmaf_rn(w[v], g[v].z, acc.z);
          acc.w = __fmaf_rn(w[v], g[v].w, acc.w);
        }
      }
    }
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && lane_on) {
    float4 tot = s_part[0][lane];
    for (int wi = 1; wi < kGatherWarps; ++wi) {
      const float4 p = s_part[wi][lane];
      tot.x += p.x; tot.y += p.y; tot.z += p.z; tot.w += p.w;
    }
    float4 o;
    o.x = __fmul_rn(alpha, tot.x);
    o.y = __fmul_rn(alpha, tot.y);
    o.z = __fmul_rn(alpha, tot.z);
    o.w = __fmul_rn(alpha, tot.w);
    *reinterpret_cast<float4*>(gfeat2 + (static_cast<size_t>(b) * S + s) * C + lane * 4) = o;
  }
}

// Feature gradient for SMALL source blocks (S * C <= 3072 floats, C <= 128: the rectify-prompter propagation, 32
// sources x 96 channels): grad_out is read ONCE and no sort is needed.  A CTA takes a span of one cloud's targets;
// every warp keeps a private copy of all S accumulator rows in shared memory (12 KB each) and walks its targets in
// order: one coalesced row load (a float4 per lane), then k shared-memory read-modify-writes acc[idx_j] += w_j * row
// (the k sources of a target are distinct, and a warp owns its copy: no conflicts, no atomics).  The 8 copies are
// combined in warp order into a per-CTA partial; interp_bwd_acc_combine_kernel adds the partials of a cloud in span
// order and applies alpha.  Fixed assignment + fixed order: deterministic.
constexpr int kAccWarps = 8;
constexpr int kAccMaxFloats = 3072;  // S * C per accumulator copy: 8 copies = 96 KB, two CTAs per SM

// spans per cloud: about two CTAs per SM, at least 64 targets each
__host__ inline int acc_spans(int B, int N) {
  int sp = 2 * 148 / B;  // every CTA resident at once (a second, nearly empty wave would double the time)
  const int smax = (N + 63) / 64;
  sp = sp > smax ? smax : sp;
  return sp < 1 ? 1 : sp;
}

__global__ void __launch_bounds__(kAccWarps * kWarp)
    interp_bwd_acc_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx, const float* __restrict__ weight,
                          int N, int S, int C, int k, int span, float* __restrict__ partial) {
  extern __shared__ __align__(16) float s_acc[];  // kAccWarps x S * C
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * span, n1 = min(N, n0 + span);
  const int SC = S * C;
  for (int f = t * 4; f < kAccWarps * SC; f += kAccWarps * kWarp * 4)
    *reinterpret_cast<float4*>(s_acc + f) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const bool lane_on = lane * 4 < C;
  float* acc = s_acc + warp * SC + lane * 4;
  const float* gb = gout + static_cast<size_t>(b) * N * C + lane * 4;
  // the next target's selection and row are in flight while this one is accumulated
  int ni = 0;
  float nw = 0.f;
  float4 ng = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int n) {
    if (n < n1) {
      const size_t row = static_cast<size_t>(b) * N + n;
      if (lane < k) {
        ni = __ldg(idx + row * k + lane);
        nw = __ldg(weight + row * k + lane);
      }
      if (lane_on) ng = __ldg(reinterpret_cast<const float4*>(gb + static_cast<size_t>(n) * C));
    }
  };
  fetch(n0 + warp);
  for (int n = n0 + warp; n < n1; n += kAccWarps) {
    const int ci = ni;
    const float cw = nw;
    const float4 g = ng;
    fetch(n + kAccWarps);
    // four neighbours at a time: their rows are DISTINCT (a target's k sources are), so the four read-modify-writes
    // are independent -- loads first, then the stores (one at a time they serialise on possible aliasing)
    for (int j0 = 0; j0 < k; j0 += 4) {
      int sj[4];
      float wj[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        sj[v] = __shfl_sync(0xffffffffu, ci, (j0 + v) & 31);
        wj[v] = __shfl_sync(0xffffffffu, cw, (j0 + v) & 31);
      }
      if (lane_on) {
        float4 r[4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (j0 + v < k) r[v] = *reinterpret_cast<const float4*>(acc + sj[v] * C);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (j0 + v < k) {
            r[v].x = __fmaf_rn(wj[v], g.x, r[v].x);
            r[v].y = __fmaf_rn(wj[v], g.y, r[v].y);
            r[v].z = __fmaf_rn(wj[v], g.z, r[v].z);
            r[v].w = __fmaf_rn(wj[v], g.w, r[v].w);
          }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (j0 + v < k) *reinterpret_cast<float4*>(acc + sj[v] * C) = r[v];
      }
    }
  }
  __syncthreads();
  float* out = partial + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * SC;
  for (int f = t * 4; f < SC; f += kAccWarps * kWarp * 4) {
    float4 tot = *reinterpret_cast<const float4*>(s_acc + f);
#pragma unroll
    for (int w = 1; w < kAccWarps; ++w) {
      const float4 p = *reinterpret_cast<const float4*>(s_acc + w * SC + f);
      tot.x += p.x; tot.y += p.y; tot.z += p.z; tot.w += p.w;
    }
    *reinterpret_cast<float4*>(out + f) = tot;
  }
}

__global__ void __launch_bounds__(256)
    interp_bwd_acc_combine_kernel(const float* __restrict__ partial, float alpha, int SC, int spans,
                                  float* __restrict__ gfeat2) {
  const int b = blockIdx.y;
  const int f = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (f >= SC) return;
  const float* p = partial + static_cast<size_t>(b) * spans * SC + f;
  float4 tot = *reinterpret_cast<const float4*>(p);
  for (int sp = 1; sp < spans; ++sp) {
    const float4 q = *reinterpret_cast<const float4*>(p + static_cast<size_t>(sp) * SC);
    tot.x += q.x; tot.y += q.y; tot.z += q.z; tot.w += q.w;
  }
  float4 o;
  o.x = __fmul_rn(alpha, tot.x);
  o.y = __fmul_rn(alpha, tot.y);
  o.z = __fmul_rn(alpha, tot.z);
  o.w = __fmul_rn(alpha, tot.w);
  *reinterpret_cast<float4*>(gfeat2 + static_cast<size_t>(b) * SC + f) = o;
}

// Coordinate gradient of the sources from the CSR (replaces the source-side kernel's list scan when the streamed
// path runs): one warp per (cloud, source), lane l takes tiles l, l + 32, ...; per match
//   grad_xyz2[b,s] += -2 * gd[p] * (x1[n] - x2[s])        (d/dx2 of -2 x1.x2 + |x1|^2 + |x2|^2 is -2 (x1 - x2))
// accumulated per lane in tile order, then combined by a fixed shuffle tree: deterministic, no atomics.
__global__ void __launch_bounds__(256)
    interp_bwd_xyz2_kernel(const unsigned char* __restrict__ csr, const float* __restrict__ gd,
                           const float* __restrict__ xyz1, const float* __restrict__ xyz2, int B, int N, int S, int k,
                           int tiles, float* __restrict__ gxyz2) {
  const int lane = threadIdx.x & 31;
  const long wg = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (wg >= static_cast<long>(B) * S) return;
  const int b = static_cast<int>(wg / S), s = static_cast<int>(wg % S);
  const float* p2 = xyz2 + (static_cast<size_t>(b) * S + s) * 3;
  const float sx = __ldg(p2), sy = __ldg(p2 + 1), sz = __ldg(p2 + 2);
  const unsigned stride = bs_block_bytes(k);
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int t = lane; t < tiles; t += 32) {
    const unsigned char* blk = csr + (static_cast<size_t>(b) * tiles + t) * stride;
    const uint16_t* off = reinterpret_cast<const uint16_t*>(blk);
    const uint2* ent = reinterpret_cast<const uint2*>(blk + kBsOffBytes);
    const unsigned char* jb = blk + bs_block_copy_bytes(k);
    const int beg = off[s], end = off[s + 1];
    for (int m = beg; m < end; ++m) {
      const int n = t * kBsTile + static_cast<int>(ent[m].x / kBsRowBytes);
      const size_t row = static_cast<size_t>(b) * N + n;
      const float g = __fmul_rn(-2.0f, __ldg(gd + row * k + jb[m]));
      ax = __fmaf_rn(g, __ldg(xyz1 + row * 3) - sx, ax);
      ay = __fmaf_rn(g, __ldg(xyz1 + row * 3 + 1) - sy, ay);
      az = __fmaf_rn(g, __ldg(xyz1 + row * 3 + 2) - sz, az);
    }
  }
  ax = warp_sum(ax);
  ay = warp_sum(ay);
  az = warp_sum(az);
  if (lane == 0) {
    float* o = gxyz2 + (static_cast<size_t>(b) * S + s) * 3;
    o[0] = ax;
    o[1] = ay;
    o[2] = az;
  }
}

// UPP_INTERP_PATH (tuning / test aid): 0 = never take the wide-feature paths, 1 = take them whenever the
// shape allows, unset = when the shape allows AND the problem is large enough to pay for the extra launch.
static int interp_path_env() {
  const char* v = tuning_env("UPP_INTERP_PATH");
  return v ? atoi(v) : -1;
}

static bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) % 16) == 0;
}

// Target spans per (cloud, chunk).  Every CTA pays its feature-tile load (64 KB, ~3 us of exposed latency: CTAs of a
// wave start together, so nothing overlaps it) before it streams, so FEWER, LONGER CTAs win as long as every SM has
// work: B200 sweep at the seg shape (288 (cloud, chunk) items, forward us): 1 span 63.0, 2: 67.1, 3: 68.6, 6: 67.6,
// 12: 69.1, 24: 75.3.  Hence: the fewest spans that give about 1.5 CTAs per SM.
static int blend_pick_spans(long items0, int N) {
  const long want = (3 * 148 / 2 + items0 - 1) / items0;
  const long smax = N / 128 < 1 ? 1 : N / 128;
  return static_cast<int>(want < 1 ? 1 : (want > smax ? smax : want));
}

// The shared-memory blend over a 2-D TMA tile of the source features (second launch of the two-phase forward, and the
// whole of a forward that REUSES a cached selection).
static int launch_blend(const CUtensorMap& fmap, const float* base, float alpha, const int32_t* idx, const float* weight,
                        int B, int N, int S, int C, int k, float* out, cudaStream_t st) {
  const int chunks = C >= kBlendCh ? C / kBlendCh : 1;
  const int chw = min(C, kBlendCh);  // channels of a staged row
  const char* wv = tuning_env("UPP_BLEND_WARPS");  // tuning aid: 8 or 16 warps per CTA
  const int nwb = (wv && atoi(wv) == 8) ? 8 : 16;
  const int kk = (k == 3 || k == 4 || k == 8 || k == 16) ? k : 0;  // the instantiated neighbour counts; 0 = run-time loop
  const size_t bsmem = ((static_cast<size_t>(S) * chw * sizeof(float) + 15) & ~static_cast<size_t>(15)) +
                       static_cast<size_t>(nwb) * blend_tw(kk) * k * 8;
  const char* sv2 = tuning_env("UPP_BLEND_SPANS");  // tuning aid: force the number of target spans per (cloud, chunk)
  const int spans = sv2 ? max(1, atoi(sv2)) : blend_pick_spans(static_cast<long>(chunks) * B, N);
  const int span = ((N + spans - 1) / spans + 1) & ~1;  // whole pairs of targets (a warp blends two per trip)
  dim3 bgrid(chunks, (N + span - 1) / span, B);
#define UPP_BLEND(K_, W_)                                                                                               \
  do {                                                                                                                   \
    if (bsmem > 40 * 1024) {                                                                                             \
      cudaError_t e = cudaFuncSetAttribute(interp_blend_kernel<K_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                           static_cast<int>(bsmem));                                                     \
      if (e != cudaSuccess) return static_cast<int>(e);                                                                  \
    }                                                                                                                    \
    interp_blend_kernel<K_, W_><<<bgrid, W_ * kWarp, bsmem, st>>>(fmap, base, alpha, idx, weight, N, S, C, k, span, out); \
  } while (0)
#define UPP_BLEND_K(W_)              \
  do {                               \
    if (k == 3) UPP_BLEND(3, W_);    \
    else if (k == 4) UPP_BLEND(4, W_); \
    else if (k == 8) UPP_BLEND(8, W_); \
    else if (k == 16) UPP_BLEND(16, W_); \
    else UPP_BLEND(0, W_);           \
  } while (0)
  if (nwb == 8) UPP_BLEND_K(8);
  else UPP_BLEND_K(16);
#undef UPP_BLEND_K
#undef UPP_BLEND
  count_launch();
  return launch_status();
}

int interp_fwd_launch(const float* xyz1, const float* xyz2, const float* feat2, const float* base, float alpha,
                      float eps, int B, int N, int S, int C, int k, float* out, int32_t* idx, float* weight,
                      float* distk, cudaStream_t st) {
  const bool vec4 = (C % 4 == 0) && aligned16(feat2, out, base);
  // wide features: selection alone, then the shared-memory blend (interp_blend_kernel)
  const int env = interp_path_env();
  // 128-channel chunks, or one narrower chunk holding the whole row (C < 128)
  const bool blend_ok = vec4 && C > 0 && (C % kBlendCh == 0 || C < kBlendCh) && S <= kBlendMaxS && B <= 65535;
  // thread-per-target selection: instantiated neighbour counts, cost ~ S * k compare/selects per target
  const bool sel_k = k == 1 || k == 2 || k == 3 || k == 4 || k == 6 || k == 8 || k == 16;
  const bool sel_ok = sel_k && S <= 1024 && static_cast<long>(S) * k <= 1024;
  const bool blend_big = static_cast<size_t>(B) * N * C >= (static_cast<size_t>(8) << 20) && N >= 256;
  // narrow / mid-size problems whose one-launch cost is the per-target warp selection (k = 16 of 32 sources: a 15-stage
  // sort and a 16-step weight sum per target) also split, as long as the selection can go thread-per-target
  const bool blend_mid = sel_ok && static_cast<long>(B) * N >= 8192 && N >= 256;
  bool two_phase = blend_ok && (env == 1 || (env < 0 && (blend_big || blend_mid)));
  CUtensorMap fmap;  // feat2 as (B*S rows) x C, box = S rows x 128 channels
  if (two_phase && make_tmap_2d_f32(&fmap, feat2, static_cast<uint64_t>(C), static_cast<uint64_t>(B) * S,
                                    static_cast<uint64_t>(C) * sizeof(float), static_cast<uint32_t>(min(C, kBlendCh)),
                                    static_cast<uint32_t>(S)) != UPP_OK)
    two_phase = false;  // no tensor-map encoder in this driver: the one-launch kernel serves every shape
  const int csel = two_phase ? 0 : C;
  // UPP_INTERP_SELECT (test aid): 0 = never the thread-per-target selection, 1 = whenever k <= 4 and S <= 1024
  const char* sv = tuning_env("UPP_INTERP_SELECT");
  const int senv = sv ? atoi(sv) : -1;
  // (C == 0: a selection-only call, upp_interp_select_f32 -- the fast selection whenever it applies)
  const bool thread_select = (two_phase || C == 0) && sel_ok && senv != 0 && (senv == 1 || static_cast<long>(B) * N >= 4096);
  if (thread_select) {
    const char* tv = tuning_env("UPP_INTERP_TPT");  // tuning aid: threads per target (1, 2, 4)
    const int tpt = tv ? atoi(tv) : 1;
#define UPP_SELECT(K_, T_)                                                                                          \
  do {                                                                                                             \
    dim3 sgrid((N + kSelThreads / T_ - 1) / (kSelThreads / T_), B);                                                \
    const size_t ssm = (static_cast<size_t>((S + T_ - 1) / T_ + 1) * 4 * T_ + static_cast<size_t>(S) * 3) * sizeof(float); \
    interp_select_kernel<K_, T_><<<sgrid, kSelThreads, ssm, st>>>(xyz1, xyz2, eps, N, S, idx, weight, distk);      \
  } while (0)
#define UPP_SELECT_K(T_)            \
  do {                              \
    if (k == 1) UPP_SELECT(1, T_);  \
    else if (k == 2) UPP_SELECT(2, T_); \
    else if (k == 3) UPP_SELECT(3, T_); \
    else UPP_SELECT(4, T_);         \
  } while (0)
    if (k == 6) UPP_SELECT(6, 1);
    else if (k == 8) UPP_SELECT(8, 1);
    else if (k == 16) UPP_SELECT(16, 1);
    else if (tpt == 4) UPP_SELECT_K(4);
    else if (tpt == 2) UPP_SELECT_K(2);
    else UPP_SELECT_K(1);
#undef UPP_SELECT_K
#undef UPP_SELECT
  } else {
  // targets per warp: as many as keep >= 4 residency waves (3 CTAs x 148 SMs) of CTAs in the grid
  int tpw = 1;
  if (S <= kKnnTile)
    for (int cand = 8; cand > 1; cand >>= 1)
      if (static_cast<long>(B) * ((N + kKnnWarps * cand - 1) / (kKnnWarps * cand)) >= 4L * 3 * 148) { tpw = cand; break; }
  dim3 grid((N + kKnnWarps * tpw - 1) / (kKnnWarps * tpw), B);
  const size_t smem = static_cast<size_t>(min(S, kKnnTile)) * 3 * sizeof(float);
  const int threads = kKnnWarps * kWarp;
#define UPP_INTERP(V_, SL_) \
  interp_fwd_kernel<V_, SL_><<<grid, threads, smem, st>>>(xyz1, xyz2, feat2, base, alpha, eps, N, S, csel, k, tpw, out, idx, weight, distk)
  if (vec4) {
    if (S <= 128) UPP_INTERP(true, 4);
    else if (S <= 256) UPP_INTERP(true, 8);
    else UPP_INTERP(true, 32);
  } else {
    if (S <= 128) UPP_INTERP(false, 4);
    else if (S <= 256) UPP_INTERP(false, 8);
    else UPP_INTERP(false, 32);
  }
#undef UPP_INTERP
  }
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK || !two_phase) return rc;

  return launch_blend(fmap, base, alpha, idx, weight, B, N, S, C, k, out, st);
}

// Forward from a CACHED selection (the SA-units call propagate six times on identical geometry,
// models/Point_MAE_pretask_dev.py:298 -- the selection is paid once, every later call is this blend alone).
// Generic shapes: one warp per target reads its k saved {index, weight} pairs and runs the gather of the one-launch
// kernel (same helper, same arithmetic order: bit-identical output); wide / narrow-chunk shapes take the shared-memory
// blend kernel exactly as the two-phase forward does.
template <bool VEC4>
__global__ void __launch_bounds__(kKnnWarps * kWarp)
    interp_apply_kernel(const float* __restrict__ feat2, const float* __restrict__ base, float alpha,
                        const int32_t* __restrict__ idx, const float* __restrict__ weight, int N, int S, int C, int k,
                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int b = blockIdx.y;
  const int n = blockIdx.x * kKnnWarps + warp;
  if (n >= N) return;  // whole warp
  const size_t row = static_cast<size_t>(b) * N + n;
  int li = 0;
  float w = 0.f;
  if (lane < k) {
    li = __ldg(idx + row * k + lane);
    w = __ldg(weight + row * k + lane);
  }
  const float* fb = feat2 + static_cast<size_t>(b) * S * C;
  float* orow = out + row * C;
  const float* brow = base ? base + row * C : nullptr;
  if (VEC4) {
    int cs = 0;
    for (; cs + kInterpCB * 128 <= C; cs += kInterpCB * 128)
      interp_gather_blocks<kInterpCB, false>(fb, brow, orow, C, cs, k, w, li, alpha);
    for (; cs < C; cs += 128) interp_gather_blocks<1, true>(fb, brow, orow, C, cs, k, w, li, alpha);
  } else {
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        const float wj = __shfl_sync(0xffffffffu, w, j);
        const int ij = __shfl_sync(0xffffffffu, li, j);
        if (c < C) acc = __fadd_rn(acc, __fmul_rn(__ldg(fb + static_cast<size_t>(ij) * C + c), wj));
      }
      if (c < C) {
        float o = __fmul_rn(alpha, acc);
        if (brow) o = __fadd_rn(__ldg(brow + c), o);
        orow[c] = o;
      }
    }
  }
}

int interp_blend_launch(const float* feat2, const float* base, float alpha, const int32_t* idx, const float* weight,
                        int B, int N, int S, int C, int k, float* out, cudaStream_t st) {
  const bool vec4 = (C % 4 == 0) && aligned16(feat2, out, base);
  const int env = interp_path_env();
  const bool blend_ok = vec4 && C > 0 && (C % kBlendCh == 0 || C < kBlendCh) && S <= kBlendMaxS && B <= 65535;
  // the staged-tile blend pays from a few thousand targets per launch on; below that the per-target gather is cheaper
  bool tiled = blend_ok && (env == 1 || (env < 0 && static_cast<long>(B) * N >= 8192 && N >= 256));
  CUtensorMap fmap;
  if (tiled && make_tmap_2d_f32(&fmap, feat2, static_cast<uint64_t>(C), static_cast<uint64_t>(B) * S,
                                static_cast<uint64_t>(C) * sizeof(float), static_cast<uint32_t>(min(C, kBlendCh)),
                                static_cast<uint32_t>(S)) != UPP_OK)
    tiled = false;
  if (tiled) return launch_blend(fmap, base, alpha, idx, weight, B, N, S, C, k, out, st);
  dim3 grid((N + kKnnWarps - 1) / kKnnWarps, B);
  if (vec4) interp_apply_kernel<true><<<grid, kKnnWarps * kWarp, 0, st>>>(feat2, base, alpha, idx, weight, N, S, C, k, out);
  else interp_apply_kernel<false><<<grid, kKnnWarps * kWarp, 0, st>>>(feat2, base, alpha, idx, weight, N, S, C, k, out);
  count_launch();
  return launch_status();
}

// Workspace of the streamed backward (one CSR block per cloud and 64-target tile); 0 when the shape has no such path.
// [CSR blocks][per-span partial sums of the small-source-block kernel]; either part may be absent
static size_t interp_bwd_csr_bytes(int B, int N, int S, int C, int k) {
  if (B <= 0 || N <= 0 || C <= 0 || S > kBsSrc || B > 65535) return 0;
  const bool wide = C % kBlendCh == 0 && k <= kBsMaxK;                     // streamed kernel
  const bool narrow = C <= kBlendCh && C % 4 == 0 && k <= kBsMaxKNarrow;   // CSR gather kernel / coordinate terms
  if (!wide && !narrow) return 0;
  const size_t tiles = (static_cast<size_t>(N) + kBsTile - 1) / kBsTile;
  return (static_cast<size_t>(B) * tiles * bs_block_bytes(k) + 255) & ~static_cast<size_t>(255);
}
static bool interp_bwd_acc_shape(int B, int N, int S, int C, int k) {
  return B > 0 && B <= 65535 && N > 0 && S > 0 && C > 0 && C <= kBlendCh && C % 4 == 0 && k <= 32 &&
         static_cast<long>(S) * C <= kAccMaxFloats;
}
size_t interp_bwd_workspace_bytes(int B, int N, int S, int C, int k) {
  size_t bytes = interp_bwd_csr_bytes(B, N, S, C, k);
  if (interp_bwd_acc_shape(B, N, S, C, k))
    bytes += static_cast<size_t>(B) * acc_spans(B, N) * S * C * sizeof(float);
  return bytes;
}

int interp_bwd_launch(const float* gout, const int32_t* idx, const float* weight, const float* distk,
                      const float* feat2, const float* xyz1, const float* xyz2, float alpha, float eps, int B,
                      int N, int S, int C, int k, float* gfeat2, float* gxyz1, float* gxyz2, float* gd_ws,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool want_xyz = gd_ws != nullptr;
  if (want_xyz) {
    const size_t rows = static_cast<size_t>(B) * N;
    const size_t want = (rows + 7) / 8;
    const int blocks = static_cast<int>(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
    const bool vec4 = (C % 4 == 0) && aligned16(gout, feat2);
    if (vec4)
      interp_bwd_target_kernel<true><<<blocks, 256, 0, st>>>(gout, feat2, xyz1, xyz2, idx, weight, distk, alpha, eps, N,
                                                             S, C, k, rows, gd_ws, gxyz1);
    else
      interp_bwd_target_kernel<false><<<blocks, 256, 0, st>>>(gout, feat2, xyz1, xyz2, idx, weight, distk, alpha, eps, N,
                                                              S, C, k, rows, gd_ws, gxyz1);
    count_launch();
    int rc = launch_status();
    if (rc != UPP_OK) return rc;
  }
  // wide features: the feature gradient streams grad_out once (interp_bwd_stream_kernel) and the sources' coordinate
  // gradient is read off the same CSR (interp_bwd_xyz2_kernel); the source-side kernel is not launched at all
  const int env = interp_path_env();
  const size_t need = interp_bwd_workspace_bytes(B, N, S, C, k);
  const size_t csr_bytes = interp_bwd_csr_bytes(B, N, S, C, k);
  const bool ws_ok = need > 0 && ws != nullptr && ws_bytes >= need && aligned16(gout, gfeat2, ws);
  const bool csr_ok = ws_ok && csr_bytes > 0;
  const bool wide = C % kBlendCh == 0 && k <= kBsMaxK;
  const bool stream_big = static_cast<long>(C / kBlendCh) * B >= 120 && N >= 512;
  bool streamed = csr_ok && wide && (env == 1 || (env < 0 && stream_big));
  // narrow features (one chunk of <= 128 channels): gather from the CSR, no streaming pipeline
  const bool gather_big = static_cast<long>(N) * k >= 4096;
  // small source blocks: per-warp accumulator copies in shared memory, grad_out read once (coordinate terms still
  // come off the CSR, so with them the CSR must exist too)
  const bool accumulated = ws_ok && !streamed && interp_bwd_acc_shape(B, N, S, C, k) && (!want_xyz || csr_ok) &&
                           (env == 1 || (env < 0 && gather_big));
  const bool gathered = csr_ok && !streamed && !accumulated && C <= kBlendCh && C % 4 == 0 && k <= kBsMaxKNarrow &&
                        (env == 1 || (env < 0 && gather_big));
  CUtensorMap gmap;  // grad_out as (B*N rows) x C, box = 64 rows x 128 channels
  if (streamed && make_tmap_2d_f32(&gmap, gout, static_cast<uint64_t>(C), static_cast<uint64_t>(B) * N,
                                   static_cast<uint64_t>(C) * sizeof(float), kBlendCh, kBsTile) != UPP_OK)
    streamed = false;  // no tensor-map encoder in this driver: the source-side kernel serves every shape
  if (accumulated) {
    const int spans = acc_spans(B, N);
    const int span = (N + spans - 1) / spans;
    const size_t asmem = static_cast<size_t>(kAccWarps) * S * C * sizeof(float);
    float* partial = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + csr_bytes);
    if (asmem > 40 * 1024) {
      cudaError_t ae = cudaFuncSetAttribute(interp_bwd_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(asmem));
      if (ae != cudaSuccess) return static_cast<int>(ae);
    }
    interp_bwd_acc_kernel<<<dim3((N + span - 1) / span, B), kAccWarps * kWarp, asmem, st>>>(gout, idx, weight, N, S, C, k,
                                                                                           span, partial);
    count_launch();
    int arc = launch_status();
    if (arc != UPP_OK) return arc;
    interp_bwd_acc_combine_kernel<<<dim3((S * C / 4 + 255) / 256, B), 256, 0, st>>>(partial, alpha, S * C,
                                                                                     (N + span - 1) / span, gfeat2);
    count_launch();
    arc = launch_status();
    if (arc != UPP_OK || !want_xyz) return arc;
  }
  if (!streamed && !gathered && !accumulated) {
    dim3 grid(S, B);
    const float* gdp = want_xyz ? gd_ws : nullptr;
    float* g2p = want_xyz ? gxyz2 : nullptr;
#define UPP_SRC(G_) \
  interp_bwd_source_kernel<G_><<<grid, kSrcThreads, 0, st>>>(gout, idx, weight, gdp, xyz1, xyz2, alpha, N, S, C, k, gfeat2, g2p)
    if (C > 1024) UPP_SRC(1);       // 256 threads x 8 channels per pass
    else if (C > 512) UPP_SRC(2);
    else if (C > 256) UPP_SRC(4);
    else UPP_SRC(8);                // 32 threads x 8 channels = 256 channels per group
#undef UPP_SRC
    count_launch();
    return launch_status();
  }
  const int tiles = (N + kBsTile - 1) / kBsTile;
  const long warps = static_cast<long>(B) * tiles;
  if (k <= kBsMaxK)
    interp_csr_kernel<kBsTile * kBsMaxK / 32><<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(
        idx, weight, B, N, k, tiles, static_cast<unsigned char*>(ws));
  else
    interp_csr_kernel<kBsTile * kBsMaxKNarrow / 32><<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(
        idx, weight, B, N, k, tiles, static_cast<unsigned char*>(ws));
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  if (want_xyz && gxyz2 != nullptr) {
    const long w2 = static_cast<long>(B) * S;
    interp_bwd_xyz2_kernel<<<static_cast<unsigned>((w2 + 7) / 8), 256, 0, st>>>(static_cast<const unsigned char*>(ws), gd_ws,
                                                                              xyz1, xyz2, B, N, S, k, tiles, gxyz2);
    count_launch();
    rc = launch_status();
    if (rc != UPP_OK) return rc;
  }
  if (accumulated) return UPP_OK;  // (coordinate terms only: the features were done above)
  if (gathered) {
    interp_bwd_gather_csr_kernel<<<dim3(S, B), kGatherWarps * kWarp, 0, st>>>(gout, static_cast<const unsigned char*>(ws),
                                                                              alpha, N, S, C, k, tiles, gfeat2);
    count_launch();
    return launch_status();
  }
  const size_t ssmem = static_cast<size_t>(kBsStages) * bs_stage_bytes(k);
  cudaError_t e = cudaFuncSetAttribute(interp_bwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(ssmem));
  if (e != cudaSuccess) return static_cast<int>(e);
  interp_bwd_stream_kernel<<<dim3(C / kBlendCh, B), (kBsWarps + 1) * kWarp, ssmem, st>>>(
      gmap, static_cast<const unsigned char*>(ws), alpha, N, S, C, k, tiles, gfeat2);
  count_launch();
  return launch_status();
}

}  // namespace upp
