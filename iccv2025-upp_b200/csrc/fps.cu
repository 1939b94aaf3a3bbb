// fps.cu -- farthest point sampling for sm_100a.
//
// Replaces pointnet2_ops furthest_point_sampling_kernel (reference call site utils/misc.py:18).
// Upstream: one CTA per cloud, running min-distance array in GLOBAL memory (L2 round trip every
// iteration), ~9 __syncthreads per iteration for a shared-memory tree arg-max.
//
// Here (fps_reg_kernel): one cloud per CTA,
//   * the cloud is staged once into shared memory with a 1-D TMA bulk copy (UBLKCP);
//   * every thread keeps its P points AND their running min-distance in REGISTERS for the
//     whole kernel -- HBM traffic is 12N in + 4M(+12M) out per cloud, nothing else;
//   * the block arg-max is two REDUX stages around ONE __syncthreads per iteration (a single
//     warp needs neither the barrier nor the second stage):
//     warp max of the (order-preserving) float bit pattern, lowest matching point index by
//     REDUX.MIN, per-warp winners double-buffered in shared memory;
//   * ties resolve to the lowest point index, deterministically.
// Semantics kept from upstream: start at index 0, temp = 1e10, d2 = min(d, temp), points with
// x^2+y^2+z^2 <= 1e-3 (double compare) never selected, d = fma(dz,dz,fma(dx,dx,dy*dy)).
#include <limits.h>

#include <stdio.h>

#include "fps_round.cuh"

namespace upp {

constexpr int kFpsMaxRegPoints = 8192; // largest N the register-resident kernel covers

// THREADS is any multiple of 32; one warp means no barrier and no second stage at all.
// TT == 0 is the tuning variant: thread count taken from blockDim.x at run time.
template <int TT, int P>
__global__ void __launch_bounds__(TT ? TT : (P > 8 ? 512 : 1024), 1)
    fps_reg_kernel(const float* __restrict__ xyz, int N, int M, int32_t* __restrict__ idx_out,
                   float* __restrict__ centers_out) {
  const int THREADS = TT ? TT : static_cast<int>(blockDim.x);
  const int NWARPS = THREADS / kWarp;
  extern __shared__ __align__(16) float s_xyz[];  // 3*N floats (AoS, as in global memory)
  __shared__ int2 s_slot[2][kWarp];
  __shared__ __align__(8) uint64_t s_bar;

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);

  float x[P], y[P], z[P], md[P];
#pragma unroll
  for (int r = 0; r < P; ++r) {
    const int i = t + r * THREADS;  // ascending in r: lowest-index tie-break inside a thread
    if (i < N) {
      x[r] = s_xyz[3 * i];
      y[r] = s_xyz[3 * i + 1];
      z[r] = s_xyz[3 * i + 2];
      md[r] = fps_initial_md(x[r], y[r], z[r]);
    } else {
      x[r] = y[r] = z[r] = 0.f;
      md[r] = kOutOfRange;
    }
  }

  float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
  if (t == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }

  for (int j = 1; j < M; ++j) {
    int best = INT_MIN;
#pragma unroll
    for (int r = 0; r < P; ++r) {
      const float d = dist_yxz(x[r] - cx, y[r] - cy, z[r] - cz);
      md[r] = fminf(md[r], d);
      best = max(best, __float_as_int(md[r]));
    }
    // stage 1: warp winner (value, lowest index)
    const int wbest = redux_max_s32(best);
    unsigned cand = 0xffffffffu;
    if (best == wbest) {
#pragma unroll
      for (int r = P - 1; r >= 0; --r)
        if (__float_as_int(md[r]) == wbest) cand = static_cast<unsigned>(t + r * THREADS);
    }
    int sel = static_cast<int>(redux_min_u32(cand));
    if (NWARPS > 1) {
      // stage 2: block winner over the per-warp winners (one barrier, double-buffered slots)
      int2* slot = s_slot[j & 1];
      if (lane == 0) slot[warp] = make_int2(wbest, sel);
      __syncthreads();
      const int2 s = (lane < NWARPS) ? slot[lane] : make_int2(INT_MIN, INT_MAX);
      const int bbest = redux_max_s32(s.x);
      sel = static_cast<int>(redux_min_u32(s.x == bbest ? static_cast<unsigned>(s.y) : 0xffffffffu));
    }
    cx = s_xyz[3 * sel];
    cy = s_xyz[3 * sel + 1];
    cz = s_xyz[3 * sel + 2];
    if (t == 0) {
      out[j] = sel;
      if (cen) { cen[3 * j] = cx; cen[3 * j + 1] = cy; cen[3 * j + 2] = cz; }
    }
  }
}

// Variant for small / medium clouds: FOUR warps per cloud, one per SM sub-partition, so the P
// points of a thread issue back to back with no contention, BAR.SYNC is at its cheapest (21 cycles
// at 4 warps vs 45 at 16) and the second stage is a register scan of the four per-warp winners
// (two LDS.128 + three compare/selects) instead of two more REDUX.  In-thread max and lowest-index
// resolution are balanced trees (VIMNMX3), not serial chains.  128-thread CTAs also let up to 16
// clouds share an SM when the batch is large.
template <int P>
__global__ void __launch_bounds__(128, 1)
    fps_w4_kernel(const float* __restrict__ xyz, int N, int M, int32_t* __restrict__ idx_out,
                  float* __restrict__ centers_out) {
  constexpr int THREADS = 128;
  extern __shared__ __align__(16) float s_xyz[];
  __shared__ __align__(16) int2 s_slot[2][4];
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);

  float x[P], y[P], z[P], md[P];
#pragma unroll
  for (int r = 0; r < P; ++r) {
    const int i = t + r * THREADS;
    if (i < N) {
      x[r] = s_xyz[3 * i];
      y[r] = s_xyz[3 * i + 1];
      z[r] = s_xyz[3 * i + 2];
      md[r] = fps_initial_md(x[r], y[r], z[r]);
    } else {
      x[r] = y[r] = z[r] = 0.f;
      md[r] = kOutOfRange;
    }
  }
  float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
  if (t == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }

  for (int j = 1; j < M; ++j) {
    int key[P];
#pragma unroll
    for (int r = 0; r < P; ++r) {
      const float d = dist_yxz(x[r] - cx, y[r] - cy, z[r] - cz);
      md[r] = fminf(md[r], d);
      key[r] = __float_as_int(md[r]);
    }
    // balanced max tree over the thread's keys
    int red[P];
#pragma unroll
    for (int r = 0; r < P; ++r) red[r] = key[r];
#pragma unroll
    for (int n = P; n > 1; n = (n + 2) / 3) {
#pragma unroll
      for (int q = 0; q < (n + 2) / 3; ++q) {
        int v = red[3 * q];
        if (3 * q + 1 < n) v = max(v, red[3 * q + 1]);
        if (3 * q + 2 < n) v = max(v, red[3 * q + 2]);
        red[q] = v;
      }
    }
    const int wbest = redux_max_s32(red[0]);
    // lowest point index among this thread's points equal to the warp maximum (balanced min tree)
    unsigned cnd[P];
#pragma unroll
    for (int r = 0; r < P; ++r) cnd[r] = key[r] == wbest ? static_cast<unsigned>(t + r * THREADS) : 0xffffffffu;
#pragma unroll
    for (int n = P; n > 1; n = (n + 2) / 3) {
#pragma unroll
      for (int q = 0; q < (n + 2) / 3; ++q) {
        unsigned v = cnd[3 * q];
        if (3 * q + 1 < n) v = min(v, cnd[3 * q + 1]);
        if (3 * q + 2 < n) v = min(v, cnd[3 * q + 2]);
        cnd[q] = v;
      }
    }
    const int widx = static_cast<int>(redux_min_u32(cnd[0]));
    int2* slot = s_slot[j & 1];
    if (lane == 0) slot[warp] = make_int2(wbest, widx);
    __syncthreads();
    // stage 2 in registers: scan the four per-warp winners (value desc, index asc)
    const int4 s01 = *reinterpret_cast<const int4*>(&slot[0]);
    const int4 s23 = *reinterpret_cast<const int4*>(&slot[2]);
    int bv = s01.x, bi = s01.y;
    if (s01.z > bv || (s01.z == bv && s01.w < bi)) { bv = s01.z; bi = s01.w; }
    int cv = s23.x, ci = s23.y;
    if (s23.z > cv || (s23.z == cv && s23.w < ci)) { cv = s23.z; ci = s23.w; }
    if (cv > bv || (cv == bv && ci < bi)) { bv = cv; bi = ci; }
    const int sel = bi;
    cx = s_xyz[3 * sel];
    cy = s_xyz[3 * sel + 1];
    cz = s_xyz[3 * sel + 2];
    if (t == 0) {
      out[j] = sel;
      if (cen) { cen[3 * j] = cx; cen[3 * j + 1] = cy; cen[3 * j + 2] = cz; }
    }
  }
}

// Ternary max tree over a thread's NS keys with its intermediate levels kept (for SEARCH == 2): the
// slot of the lowest-index maximum is then found by DESCENDING the tree (first child equal to the
// maximum, ~2 compares per level) instead of comparing every slot.  Only the one lane per warp that
// posts the warp's winner walks it, so for 16-32 points per thread the search costs ~16 issued
// instructions per warp and round instead of 2 per slot (the integer/ALU pipe, 16 lanes per scheduler,
// is what binds large clouds: profiles/r01d_fps8k_nw32.txt).
template <int NS>
struct KeyTree {
  static constexpr int N1 = (NS + 2) / 3, N2 = (N1 + 2) / 3, N3 = (N2 + 2) / 3, N4 = (N3 + 2) / 3;
  static_assert(N4 == 1, "KeyTree covers up to 81 keys");
  int l0[NS], l1[N1], l2[N2], l3[N3], l4[1];

  template <int NI, int NO>
  static __device__ __forceinline__ void fold(const int (&in)[NI], int (&out)[NO]) {
#pragma unroll
    for (int q = 0; q < NO; ++q) {
      int v = in[3 * q];
      if (3 * q + 1 < NI) v = max(v, in[3 * q + 1]);
      if (3 * q + 2 < NI) v = max(v, in[3 * q + 2]);
      out[q] = v;
    }
  }
  __device__ __forceinline__ int build() {
    fold(l0, l1); fold(l1, l2); fold(l2, l3); fold(l3, l4);
    return l4[0];
  }
  template <int LEVEL>
  __device__ __forceinline__ int at(int q) const {
    if constexpr (LEVEL == 0) return l0[q];
    else if constexpr (LEVEL == 1) return l1[q];
    else if constexpr (LEVEL == 2) return l2[q];
    else if constexpr (LEVEL == 3) return l3[q];
    else return l4[q];
  }
  template <int LEVEL>
  static __host__ __device__ constexpr int size() {
    return LEVEL == 0 ? NS : LEVEL == 1 ? N1 : LEVEL == 2 ? N2 : LEVEL == 3 ? N3 : 1;
  }
  // lowest slot whose key equals `best` (== the root), node Q of level LEVEL known to contain it
  template <int LEVEL, int Q>
  __device__ __forceinline__ int descend(int best) const {
    if constexpr (LEVEL == 0) {
      return Q;
    } else {
      constexpr int n = size<LEVEL - 1>();
      constexpr int c0 = 3 * Q;
      if constexpr (c0 + 1 >= n) {
        return descend<LEVEL - 1, c0>(best);
      } else {
        if (at<LEVEL - 1>(c0) == best) return descend<LEVEL - 1, c0>(best);
        if constexpr (c0 + 2 < n) {
          if (at<LEVEL - 1>(c0 + 1) == best) return descend<LEVEL - 1, c0 + 1>(best);
          return descend<LEVEL - 1, c0 + 2>(best);
        } else {
          return descend<LEVEL - 1, c0 + 1>(best);
        }
      }
    }
  }
  __device__ __forceinline__ int find(int best) const { return descend<4, 0>(best); }
};

// ---------------------------------------------------------------------------------------------
// v2 (fps_blk_kernel): the same register-resident scheme re-cut around what the B200 measurements
// say binds one round (scripts/microbench2.cu): issue slots and the dependent REDUX/BAR/LDS chain.
//   * BLOCKED ownership: thread t holds points [t*P, t*P+P), P = 2*P2.  A lower lane / lower warp
//     then always means lower point indices, so "lowest index among equal maxima" is decided by
//     position: in-thread lowest slot, ballot + find-first-set across lanes, lowest warp across
//     warps.  No index REDUX anywhere (v1: two REDUX.MIN on the dependent chain).
//   * PACKED fp32x2 math: the thread's points sit in 64-bit register pairs {x_2r, x_2r+1}; one
//     FADD2/FMUL2/FFMA2 updates two points (6 issue slots per 2 points instead of 12, results
//     bit-identical to the scalar spelling).
//   * the in-thread slot search runs in the shadow of the REDUX (it needs only the thread's own
//     maximum), the warp's winning lane posts {key, index} and ONE __syncthreads later every
//     thread resolves the block winner itself: S2 == 0 scans the NW posted pairs with vector
//     LDS + a select tree (NW <= 8), S2 == 1 does REDUX.MAX + ballot over them (NW >= 16).
//   * a single warp (NW == 1) needs no shared-memory round trip at all.
// (Posting {key, x, y, z} per warp -- every lane fetching its own candidate's coordinates while the REDUX / vote are in
//  flight, so that the select tree after the barrier yields the next centre with no second shared-memory round
//  trip -- was built and measured at N = 1228: 0.207 us per round against 0.153; the 3 extra divergent LDS per lane
//  and round and the wider select tree cost more than the dependent load they remove.)
// SEARCH: 0 = in-thread slot search in the shadow of the REDUX (select chain / min tree; small clouds,
// latency bound), 2 = the posting lane descends a kept max tree after the vote (large clouds, ALU bound).
template <int NW, int P2, int S2, int SEARCH = 0>
__global__ void __launch_bounds__(NW * 32)
    fps_blk_kernel(const float* __restrict__ xyz, int N, int M, int32_t* __restrict__ idx_out,
                   float* __restrict__ centers_out) {
  constexpr int P = 2 * P2;
  extern __shared__ __align__(16) float s_xyz[];  // 3*N floats (AoS, as in global memory)
  __shared__ __align__(16) int2 s_slot[2][NW];
  __shared__ __align__(8) uint64_t s_bar;

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  pdl_trigger();  // (programmatic dependent launch, common.cuh: the next kernel of the chain may be scheduled now)
  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();     // the cloud may come from the kernel in front of this one
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);

  const int base = t * P;
  f32x2 X[P2], Y[P2], Z[P2];
  float md[P];
#pragma unroll
  for (int r = 0; r < P2; ++r) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = base + 2 * r + h;
      if (i < N) {
        c[h][0] = s_xyz[3 * i];
        c[h][1] = s_xyz[3 * i + 1];
        c[h][2] = s_xyz[3 * i + 2];
        md[2 * r + h] = fps_initial_md(c[h][0], c[h][1], c[h][2]);
      } else {
        c[h][0] = c[h][1] = c[h][2] = 0.f;
        md[2 * r + h] = kOutOfRange;
      }
    }
    X[r] = pack2(c[0][0], c[1][0]);
    Y[r] = pack2(c[0][1], c[1][1]);
    Z[r] = pack2(c[0][2], c[1][2]);
  }

  float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
  if (t == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }

  const unsigned lanes_below = (1u << lane) - 1u;
  for (int j = 1; j < M; ++j) {
    const f32x2 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
    // The in-thread slot search needs only the thread's own maximum, so it runs in the shadow of the
    // REDUX.  (Splitting the pairs into two spans so that the first span's search issues under the
    // second span's FMA work was measured: no gain -- one warp per scheduler is latency-, not pipe-bound.)
    int best, ls = 0;
    KeyTree<SEARCH == 2 ? P : 1> kt;
    if constexpr (SEARCH == 2) {
      static_assert(NW > 1, "the deferred search posts through shared memory");
      f32x2 D[P2];
#pragma unroll
      for (int r = 0; r < P2; ++r) D[r] = sub2(Y[r], CY);
#pragma unroll
      for (int r = 0; r < P2; ++r) D[r] = mul2(D[r], D[r]);
#pragma unroll
      for (int r = 0; r < P2; ++r) { const f32x2 dx = sub2(X[r], CX); D[r] = fma2(dx, dx, D[r]); }
#pragma unroll
      for (int r = 0; r < P2; ++r) { const f32x2 dz = sub2(Z[r], CZ); D[r] = fma2(dz, dz, D[r]); }
#pragma unroll
      for (int r = 0; r < P2; ++r) {
        float d0, d1;
        unpack2(D[r], d0, d1);
        md[2 * r] = fminf(md[2 * r], d0);
        md[2 * r + 1] = fminf(md[2 * r + 1], d1);
        kt.l0[2 * r] = __float_as_int(md[2 * r]);
        kt.l0[2 * r + 1] = __float_as_int(md[2 * r + 1]);
      }
      best = kt.build();
    } else {
      fps_span<P2, 0, P2, (NW <= 4)>(X, Y, Z, md, CX, CY, CZ, best, ls);
    }
    const int wbest = redux_max_s32(best);
    const unsigned winners = __ballot_sync(0xffffffffu, best == wbest);
    int sel;
    if constexpr (NW == 1) {
      sel = __shfl_sync(0xffffffffu, base + ls, __ffs(winners) - 1);
    } else {
      int2* slot = s_slot[j & 1];
      // the lowest lane holding the warp maximum posts (lower lane == lower point indices)
      if (best == wbest && (winners & lanes_below) == 0u) {
        if constexpr (SEARCH == 2) ls = kt.find(wbest);
        slot[warp] = make_int2(wbest, base + ls);
      }
      __syncthreads();
      if constexpr (S2 == 0) {
        int v[NW], ix[NW];
#pragma unroll
        for (int w = 0; w < NW; w += 2) {
          const int4 a = *reinterpret_cast<const int4*>(&slot[w]);
          v[w] = a.x; ix[w] = a.y; v[w + 1] = a.z; ix[w + 1] = a.w;
        }
#pragma unroll
        for (int stride = 1; stride < NW; stride <<= 1)
#pragma unroll
          for (int w = 0; w + stride < NW; w += 2 * stride)
            if (v[w + stride] > v[w]) { v[w] = v[w + stride]; ix[w] = ix[w + stride]; }  // strict: lower warp on ties
        sel = ix[0];
      } else {
        const int2 s = (lane < NW) ? slot[lane] : make_int2(INT_MIN, 0);
        const int bbest = redux_max_s32(s.x);
        const unsigned wm = __ballot_sync(0xffffffffu, s.x == bbest);
        sel = __shfl_sync(0xffffffffu, s.y, __ffs(wm) - 1);
      }
    }
    cx = s_xyz[3 * sel];
    cy = s_xyz[3 * sel + 1];
    cz = s_xyz[3 * sel + 2];
    if (t == 0) out[j] = sel;
  }
  // centres (the fused gather of utils/misc.py:19) after the serial chain, by the whole CTA
  if (cen) {
    __syncthreads();  // thread 0's out[] stores are visible to the block
    for (int j = 1 + t; j < M; j += NW * 32) {
      const int sel = out[j];
      cen[3 * j] = s_xyz[3 * sel];
      cen[3 * j + 1] = s_xyz[3 * sel + 1];
      cen[3 * j + 2] = s_xyz[3 * sel + 2];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Cluster FPS: ONE cloud per thread-block CLUSTER of CS CTAs (SURVEY.md 8e "cluster-split FPS for small
// per-GPU batches").  With few large clouds -- seprate_point_cloud's B x 6144 -> 1024 (utils/misc.py:241-249),
// or C4's 8192-point clouds once the batch is sharded over 8 GPUs (16 clouds on 148 SMs) -- one CTA per cloud
// leaves most SMs idle while each round costs ~900-1050 cycles of distance updates.  Here CTA r of the cluster
// owns the contiguous slice [r * Nc, (r + 1) * Nc) of the cloud (blocked ownership as in fps_blk_kernel: a lower
// rank / warp / lane / slot is a lower point index), every CTA stages the WHOLE cloud once (coordinate look-up
// of any winner), and a round is
//   distance update + in-thread max -> REDUX.MAX (key) -> REDUX.MIN (lowest index among the lanes holding it)
//   -> lanes 0..CS-1 of every warp push the warp's {key, index} into slot [rank * NW + warp] of EVERY CTA of the
//      cluster with st.async (distributed shared memory, completes bytes on the destination's mbarrier)
//   -> every thread waits on its own CTA's mbarrier, lane i reads entry i (CS * NW <= 32 entries),
//      REDUX.MAX + REDUX.MIN pick the cluster winner -> its coordinates come from the local copy of the cloud.
// No intra-CTA barrier and no cluster barrier inside the loop: two entry buffers / two mbarriers alternate by
// round parity (a CTA can only send round j+2 after it has received every CTA's round j+1, i.e. after every
// CTA has consumed round j).  Same selection rule as every other FPS kernel here, bit-identical indices.
template <int CS, int NW, int P2>
__global__ void __launch_bounds__(NW * 32, 1)
    fps_cluster_kernel(const float* __restrict__ xyz, int N, int M, int Nc, int32_t* __restrict__ idx_out,
                       float* __restrict__ centers_out) {
  static_assert(CS * NW <= 32, "one entry per lane in the cluster stage");
  constexpr int P = 2 * P2;
  extern __shared__ __align__(16) float s_xyz[];  // 3*N floats: the whole cloud (AoS, as in global memory)
  __shared__ __align__(8) int2 s_x[2][32];        // {key, index} of every warp of the cluster, by round parity
  __shared__ __align__(8) uint64_t s_bar, s_xbar[2];

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / CS;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_init(&s_xbar[0], 1);
    mbar_init(&s_xbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anyone sends to them
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);

  const int lo = static_cast<int>(rank) * Nc;
  const int hi = min(N, lo + Nc);
  const int base = lo + t * P;
  f32x2 X[P2], Y[P2], Z[P2];
  float md[P];
#pragma unroll
  for (int r = 0; r < P2; ++r) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = base + 2 * r + h;
      if (i < hi) {
        c[h][0] = s_xyz[3 * i];
        c[h][1] = s_xyz[3 * i + 1];
        c[h][2] = s_xyz[3 * i + 2];
        md[2 * r + h] = fps_initial_md(c[h][0], c[h][1], c[h][2]);
      } else {
        c[h][0] = c[h][1] = c[h][2] = 0.f;
        md[2 * r + h] = kOutOfRange;
      }
    }
    X[r] = pack2(c[0][0], c[1][0]);
    Y[r] = pack2(c[0][1], c[1][1]);
    Z[r] = pack2(c[0][2], c[1][2]);
  }
  // lane r < CS of every warp sends to CTA r: destination entry / barrier addresses for both parities
  const unsigned dst = lane < CS ? lane : 0;
  const int my_entry = static_cast<int>(rank) * NW + warp;
  const uint32_t r_entry0 = map_to_cta(smem_u32(&s_x[0][my_entry]), dst);
  const uint32_t r_entry1 = map_to_cta(smem_u32(&s_x[1][my_entry]), dst);
  const uint32_t r_bar0 = map_to_cta(smem_u32(&s_xbar[0]), dst);
  const uint32_t r_bar1 = map_to_cta(smem_u32(&s_xbar[1]), dst);

  float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
  if (t == 0 && rank == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }
  for (int j = 1; j < M; ++j) {
    const int h = j & 1;
    if (t == 0) mbar_expect_tx(&s_xbar[h], CS * NW * 8u);  // this round's CS * NW entries of 8 bytes
    const f32x2 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
    int best, ls = 0;
    fps_span<P2, 0, P2, true>(X, Y, Z, md, CX, CY, CZ, best, ls);
    const int wbest = redux_max_s32(best);
    const unsigned widx = redux_min_u32(best == wbest ? static_cast<unsigned>(base + ls) : 0xffffffffu);
    if (lane < CS) st_async_b64(h ? r_entry1 : r_entry0, wbest, static_cast<int>(widx), h ? r_bar1 : r_bar0);
    mbar_wait(&s_xbar[h], static_cast<unsigned>((j - 1) >> 1) & 1u);  // buffer h is on its ((j - 1) / 2)-th use
    const int2 e = lane < CS * NW ? s_x[h][lane] : make_int2(INT_MIN, INT_MAX);
    const int cbest = redux_max_s32(e.x);
    const int sel = static_cast<int>(redux_min_u32(e.x == cbest ? static_cast<unsigned>(e.y) : 0xffffffffu));
    cx = s_xyz[3 * sel];
    cy = s_xyz[3 * sel + 1];
    cz = s_xyz[3 * sel + 2];
    if (t == 0 && rank == 0) out[j] = sel;
  }
  cluster_sync_all();  // nobody leaves while a peer may still be sending to it
  if (cen && rank == 0) {
    __syncthreads();  // thread 0's out[] stores are visible to the block
    for (int j = 1 + t; j < M; j += NW * 32) {
      const int sel = out[j];
      cen[3 * j] = s_xyz[3 * sel];
      cen[3 * j + 1] = s_xyz[3 * sel + 1];
      cen[3 * j + 2] = s_xyz[3 * sel + 2];
    }
  }
}

// Cluster FPS with ONE LOOK-AHEAD (fps_cluster2_kernel): the same exchange, but a round may yield TWO selections.
// A cluster round is ~600 cycles of which ~290 are the DSMEM flight and the mbarrier wake-up, whatever the payload; so
// every warp also sends an upper bound of everything else it holds (its runner-up: one more REDUX, issued beside the
// index REDUX), and after the exchange every thread checks whether the round's runner-up is already decided:
//   * p1 = the arg-max as before (largest key, lowest index);
//   * candidates for the next selection: every other entry's key, and for p1's warp its runner-up bound.  If their
//     maximum M2 > 0 is held by exactly ONE entry, that entry is not p1's warp's bound, and the point p2 it names is not
//     reached by p1 -- fma(dz,dz,fma(dx,dx,dy*dy)) of (p2 - p1), the very value the update would compute, is >= M2 -- then
//     after the update with p1 every point is still <= its old min-distance < M2 (or == M2 at a higher index inside p2's
//     warp), p2 keeps M2, p1 drops to 0: the next round's arg-max IS p2.  It is selected now, and the next round updates
//     with both centres (min3).  Any doubt (a tie between entries, the bound of p1's own warp, p2 inside p1's reach,
//     M2 <= 0, the last selection) ends the round with one selection, as before.
// Same selections bit for bit as every other FPS kernel here (the parity tests force it on the cluster shapes it is built
// for).  MEASURED (B200, round 2), NOT ADOPTED: the look-ahead is accepted in 88 % of the rounds -- 1023 selections in 545
// exchange rounds on C4's clouds (B16 x 8192), 557 at B32 x 6144 -- and the launch is no faster: 297-315 us against
// 308-324 us (B16 x 8192), 296-306 against 291 (B32 x 6144), 62-68 against 53 (M = 128: few acceptances early on).  A
// round with the second resolve costs ~1100 cycles against 620: one more REDUX before the exchange, REDUX + two votes +
// shuffle + three LDS + the distance test + the second centre's update after it -- ~50 more DEPENDENT instructions at 5-6
// cycles each plus two more 60-70-cycle REDUX round trips, which is what a second 600-cycle round would have cost.
// UPP_FPS_CLUSTER_AHEAD=1 (under UPP_TUNING=1) selects it.
constexpr int kAheadFrom = 48;

template <int CS, int NW, int P2>
__global__ void __launch_bounds__(NW * 32, 1)
    fps_cluster2_kernel(const float* __restrict__ xyz, int N, int M, int Nc, int32_t* __restrict__ idx_out,
                        float* __restrict__ centers_out) {
  static_assert(CS * NW <= 32, "one entry per lane in the cluster stage");
  constexpr int P = 2 * P2;
  extern __shared__ __align__(16) float s_xyz[];  // 3*N floats: the whole cloud (AoS, as in global memory)
  __shared__ __align__(16) int4 s_x[2][32];       // {key, index, runner-up bound, -} of every warp of the cluster, by round parity
  __shared__ __align__(8) uint64_t s_bar, s_xbar[2];

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / CS;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_init(&s_xbar[0], 1);
    mbar_init(&s_xbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anyone sends to them
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);

  const int lo = static_cast<int>(rank) * Nc;
  const int hi = min(N, lo + Nc);
  const int base = lo + t * P;
  f32x2 X[P2], Y[P2], Z[P2];
  float md[P];
#pragma unroll
  for (int r = 0; r < P2; ++r) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = base + 2 * r + h;
      if (i < hi) {
        c[h][0] = s_xyz[3 * i];
        c[h][1] = s_xyz[3 * i + 1];
        c[h][2] = s_xyz[3 * i + 2];
        md[2 * r + h] = fps_initial_md(c[h][0], c[h][1], c[h][2]);
      } else {
        c[h][0] = c[h][1] = c[h][2] = 0.f;
        md[2 * r + h] = kOutOfRange;
      }
    }
    X[r] = pack2(c[0][0], c[1][0]);
    Y[r] = pack2(c[0][1], c[1][1]);
    Z[r] = pack2(c[0][2], c[1][2]);
  }
  const unsigned dst = lane < CS ? lane : 0;
  const int my_entry = static_cast<int>(rank) * NW + warp;
  const uint32_t r_entry0 = map_to_cta(smem_u32(&s_x[0][my_entry]), dst);
  const uint32_t r_entry1 = map_to_cta(smem_u32(&s_x[1][my_entry]), dst);
  const uint32_t r_bar0 = map_to_cta(smem_u32(&s_xbar[0]), dst);
  const uint32_t r_bar1 = map_to_cta(smem_u32(&s_xbar[1]), dst);
  const unsigned lanes_below = (1u << lane) - 1u;

  float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
  float ex = cx, ey = cy, ez = cz;  // the second centre of a round (when `two`)
  bool two = false;
  if (t == 0 && rank == 0) out[0] = 0;
  int rnd = 0;
  for (int j = 1; j < M;) {
    ++rnd;
    const int h = rnd & 1;
    if (t == 0) mbar_expect_tx(&s_xbar[h], CS * NW * 16u);  // this round's CS * NW entries of 16 bytes
    // distances to the round's centre(s), running-min update, in-thread arg-max: the tuned span of every FPS kernel here
    int best, ls = 0;
    if (two) {  // cluster-uniform: the second centre of the previous round first (its maximum is recomputed below)
      const f32x2 EX = pack2(ex, ex), EY = pack2(ey, ey), EZ = pack2(ez, ez);
      f32x2 D[P2];
#pragma unroll
      for (int r = 0; r < P2; ++r) D[r] = sub2(Y[r], EY);
#pragma unroll
      for (int r = 0; r < P2; ++r) D[r] = mul2(D[r], D[r]);
#pragma unroll
      for (int r = 0; r < P2; ++r) { const f32x2 dx = sub2(X[r], EX); D[r] = fma2(dx, dx, D[r]); }
#pragma unroll
      for (int r = 0; r < P2; ++r) { const f32x2 dz = sub2(Z[r], EZ); D[r] = fma2(dz, dz, D[r]); }
#pragma unroll
      for (int r = 0; r < P2; ++r) {
        float d0, d1;
        unpack2(D[r], d0, d1);
        md[2 * r] = fminf(md[2 * r], d0);
        md[2 * r + 1] = fminf(md[2 * r + 1], d1);
      }
    }
    {
      const f32x2 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
      fps_span<P2, 0, P2, true>(X, Y, Z, md, CX, CY, CZ, best, ls);
    }
    const int wbest = redux_max_s32(best);
    // (in the shadow of the REDUX) the largest of the thread's other slots: independent selects + a max tree
    int second;
    {
      int o[P];
#pragma unroll
      for (int s = 0; s < P; ++s) o[s] = s == ls ? INT_MIN : __float_as_int(md[s]);
#pragma unroll
      for (int n = P; n > 1; n = (n + 2) / 3) {
#pragma unroll
        for (int q = 0; q < (n + 2) / 3; ++q) {
          int v = o[3 * q];
          if (3 * q + 1 < n) v = max(v, o[3 * q + 1]);
          if (3 * q + 2 < n) v = max(v, o[3 * q + 2]);
          o[q] = v;
        }
      }
      second = o[0];
    }
    const unsigned winners = __ballot_sync(0xffffffffu, best == wbest);
    const bool poster = best == wbest && (winners & lanes_below) == 0u;  // lower lane == lower point indices
    const unsigned widx = redux_min_u32(best == wbest ? static_cast<unsigned>(base + ls) : 0xffffffffu);
    const int wsec = redux_max_s32(poster ? second : best);  // everything this warp holds besides the point it names
    if (lane < CS) st_async_b128(h ? r_entry1 : r_entry0, wbest, static_cast<int>(widx), wsec, 0, h ? r_bar1 : r_bar0);
    mbar_wait(&s_xbar[h], static_cast<unsigned>((rnd - 1) >> 1) & 1u);  // buffer h is on its ((rnd - 1) / 2)-th use
    const int4 e = lane < CS * NW ? s_x[h][lane] : make_int4(INT_MIN, INT_MAX, INT_MIN, 0);
    const int k1 = redux_max_s32(e.x);
    const int sel = static_cast<int>(redux_min_u32(e.x == k1 ? static_cast<unsigned>(e.y) : 0xffffffffu));
    cx = s_xyz[3 * sel];
    cy = s_xyz[3 * sel + 1];
    cz = s_xyz[3 * sel + 2];
    // the look-ahead: is the runner-up already decided?  (Not tried during the first selections: the min-distances are
    // still so large that p1 reaches nearly everything.)
    bool ok2 = false;
    int sel2 = 0;
    if (j >= kAheadFrom && j + 1 < M) {  // cluster-uniform
      const bool is_top = e.x == k1 && e.y == sel;  // the entry naming p1
      const int cand = is_top ? e.z : e.x;
      const int m2c = redux_max_s32(cand);
      const unsigned h2 = __ballot_sync(0xffffffffu, cand == m2c && !is_top);
      const unsigned hall = __ballot_sync(0xffffffffu, cand == m2c);
      if (m2c > 0 && h2 == hall && (h2 & (h2 - 1u)) == 0u) {
        sel2 = __shfl_sync(0xffffffffu, e.y, __ffs(h2) - 1);
        ex = s_xyz[3 * sel2];
        ey = s_xyz[3 * sel2 + 1];
        ez = s_xyz[3 * sel2 + 2];
        ok2 = !(dist_yxz(ex - cx, ey - cy, ez - cz) < __int_as_float(m2c));  // p1 does not reach p2
      }
    }
    if (t == 0 && rank == 0) {
      out[j] = sel;
      if (ok2) out[j + 1] = sel2;
    }
    two = ok2;
    j += ok2 ? 2 : 1;
  }
#ifdef UPP_CLUSTER_STATS
  if (blockIdx.x == 0 && t == 0) printf("cluster look-ahead: %d selections in %d rounds\n", M - 1, rnd);
#endif
  cluster_sync_all();  // nobody leaves while a peer may still be sending to it
  if (cen && rank == 0) {
    __syncthreads();  // thread 0's out[] stores are visible to the block
    for (int j = t; j < M; j += NW * 32) {
      const int sel = out[j];
      cen[3 * j] = s_xyz[3 * sel];
      cen[3 * j + 1] = s_xyz[3 * sel + 1];
      cen[3 * j + 2] = s_xyz[3 * sel + 2];
    }
  }
}

// Any-N fallback: min-distance array in a global workspace (L2-resident), xyz re-read from
// global memory every iteration.  Same selection rule, same two-stage arg-max.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    fps_global_kernel(const float* __restrict__ xyz, int N, int M, int32_t* __restrict__ idx_out,
                      float* __restrict__ centers_out, float* __restrict__ temp) {
  constexpr int NWARPS = THREADS / kWarp;
  __shared__ int2 s_slot[2][kWarp];
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;
  float* md = temp + static_cast<size_t>(b) * N;
  int32_t* out = idx_out + static_cast<size_t>(b) * M;
  float* cen = centers_out ? centers_out + static_cast<size_t>(b) * M * 3 : nullptr;

  for (int i = t; i < N; i += THREADS) md[i] = fps_initial_md(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  float cx = p[0], cy = p[1], cz = p[2];
  if (t == 0) {
    out[0] = 0;
    if (cen) { cen[0] = cx; cen[1] = cy; cen[2] = cz; }
  }
  for (int j = 1; j < M; ++j) {
    int best = INT_MIN;
    unsigned besti = 0xffffffffu;
    for (int i = t; i < N; i += THREADS) {
      const float d = dist_yxz(p[3 * i] - cx, p[3 * i + 1] - cy, p[3 * i + 2] - cz);
      const float m = fminf(md[i], d);
      md[i] = m;
      const int key = __float_as_int(m);
      if (key > best) { best = key; besti = static_cast<unsigned>(i); }
    }
    const int wbest = redux_max_s32(best);
    int sel = static_cast<int>(redux_min_u32(best == wbest ? besti : 0xffffffffu));
    int2* slot = s_slot[j & 1];
    if (lane == 0) slot[warp] = make_int2(wbest, sel);
    __syncthreads();
    const int2 s = (lane < NWARPS) ? slot[lane] : make_int2(INT_MIN, INT_MAX);
    const int bbest = redux_max_s32(s.x);
    sel = static_cast<int>(redux_min_u32(s.x == bbest ? static_cast<unsigned>(s.y) : 0xffffffffu));
    cx = p[3 * sel];
    cy = p[3 * sel + 1];
    cz = p[3 * sel + 2];
    if (t == 0) {
      out[j] = sel;
      if (cen) { cen[3 * j] = cx; cen[3 * j + 1] = cy; cen[3 * j + 2] = cz; }
    }
  }
}

// ---- host side -------------------------------------------------------------------------

struct FpsConfig {
  int threads, p;
};

static int env_int(const char* name, int dflt) { return tuning_env_int(name, dflt); }  // UPP_TUNING=1 only

// FPS is a serial latency chain: one warp per scheduler, ~40 % issue utilisation, every instruction on
// the critical path.  Any co-resident CTA of a throughput kernel (measured: the packed Chamfer kernel
// running on another stream, 5 CTAs x 4 warps per SM) takes issue slots from it and stretches every
// round -- the whole-step graph went from 0.27 ms to 0.34-0.43 ms.  So when there are no more clouds
// than SMs, each FPS CTA asks for (nearly) all of the SM's shared memory: nothing with a shared-memory
// footprint can be placed beside it, the cloud has the SM to itself, and the other 148 - B SMs carry
// the concurrent work.  With more clouds than SMs the request is the real footprint (clouds share SMs).
constexpr int kNumSMs = 148;
constexpr size_t kExclusiveSmem = 232448 - 2048;  // 227 KB opt-in limit minus static + headroom
static size_t fps_smem_request(size_t need, int B) {
  if (env_int("UPP_FPS_SHARE_SM", 0) == 1) return need;  // A/B aid: allow co-residency
  return (B <= kNumSMs && need < kExclusiveSmem) ? kExclusiveSmem : need;
}

static const int kPs[] = {1, 2, 3, 4, 6, 8, 12, 16};

static int round_p(int p) {
  for (int v : kPs)
    if (v >= p) return v;
  return 0;
}

// Register-resident configuration for N <= kFpsMaxRegPoints, from B200 measurements
// (scripts/time_ops.py --sweep-fps, scripts/microbench.cu).  One iteration is a latency chain:
//   LDS centroid 34 -> distance/min/max ~30 -> REDUX 23 -> REDUX 23 -> STS/BAR.SYNC 21..50 ->
//   LDS 34 -> REDUX 23 -> REDUX 23 -> LDS ...   ~ 350-400 cycles (0.18-0.20 us) for any N <= 1280,
// plus ~8 issue slots per point per warp.  Measured best everywhere: TWO points per thread
// (more per thread lengthens the in-thread max/select chains; one per thread doubles the warps
// and the barrier cost), i.e. N/2 threads up to 640; beyond that 512 threads x ceil(N/512) points.
FpsConfig fps_pick_config(int N) {
  int threads, p;
  if (N <= 32) {
    threads = 32;
    p = 1;
  } else if (N <= 1280) {
    p = 2;
    threads = ((N + 1) / 2 + 31) / 32 * 32;
  } else {
    threads = 512;
    p = round_p((N + 511) / 512);
  }
  if (p == 0 || static_cast<long>(threads) * p < N) return {0, 0};
  return {threads, p};
}

template <int THREADS, int P>
static int launch_fps_reg(const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
                          cudaStream_t st) {
  const size_t smem = fps_smem_request(static_cast<size_t>(N) * 3 * sizeof(float), B);
  auto kern = fps_reg_kernel<THREADS, P>;
  if (smem > 40 * 1024) {  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  kern<<<B, THREADS, smem, st>>>(xyz, N, M, idx, centers);
  count_launch();
  return launch_status();
}

// tuning variant (UPP_FPS_THREADS + UPP_FPS_P both set): any thread count, run-time stride
template <int P>
static int launch_fps_rt(int threads, const float* xyz, int B, int N, int M, int32_t* idx,
                         float* centers, cudaStream_t st) {
  const size_t smem = fps_smem_request(static_cast<size_t>(N) * 3 * sizeof(float), B);
  auto kern = fps_reg_kernel<0, P>;
  if (smem > 40 * 1024) {  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  kern<<<B, threads, smem, st>>>(xyz, N, M, idx, centers);
  count_launch();
  return launch_status();
}

template <int P>
static int launch_fps_w4(const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
                         cudaStream_t st) {
  const size_t smem = fps_smem_request(static_cast<size_t>(N) * 3 * sizeof(float), B);
  auto kern = fps_w4_kernel<P>;
  if (smem > 40 * 1024) {  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  kern<<<B, 128, smem, st>>>(xyz, N, M, idx, centers);
  count_launch();
  return launch_status();
}

static int dispatch_fps_w4(int p, const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
                           cudaStream_t st) {
  switch (p) {
    case 1: return launch_fps_w4<1>(xyz, B, N, M, idx, centers, st);
    case 2: return launch_fps_w4<2>(xyz, B, N, M, idx, centers, st);
    case 3: return launch_fps_w4<3>(xyz, B, N, M, idx, centers, st);
    case 4: return launch_fps_w4<4>(xyz, B, N, M, idx, centers, st);
    case 6: return launch_fps_w4<6>(xyz, B, N, M, idx, centers, st);
    case 8: return launch_fps_w4<8>(xyz, B, N, M, idx, centers, st);
    case 12: return launch_fps_w4<12>(xyz, B, N, M, idx, centers, st);
    case 16: return launch_fps_w4<16>(xyz, B, N, M, idx, centers, st);
    default: return UPP_ERR_UNSUPPORTED;
  }
}

// ---- v2 dispatch ---------------------------------------------------------------------------
struct FpsBlkConfig {
  int nw, p2, s2, search;
};

template <int NW, int P2, int S2, int SEARCH = 0>
static int launch_fps_blk(const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
                          cudaStream_t st) {
  const size_t smem = fps_smem_request(static_cast<size_t>(N) * 3 * sizeof(float), B);
  auto kern = fps_blk_kernel<NW, P2, S2, SEARCH>;
  if (smem > 40 * 1024) {  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  cudaError_t le = launch_pdl(kern, dim3(B), dim3(NW * 32), smem, st, xyz, N, M, idx, centers);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return launch_status();
}

static int dispatch_fps_blk(FpsBlkConfig c, const float* xyz, int B, int N, int M, int32_t* idx,
                            float* centers, cudaStream_t st) {
#define UPP_BLK(NW_, P2_, S2_) \
  if (c.nw == NW_ && c.p2 == P2_ && c.s2 == S2_) return launch_fps_blk<NW_, P2_, S2_>(xyz, B, N, M, idx, centers, st);
#define UPP_BLK_ROW(NW_, S2_) \
  UPP_BLK(NW_, 1, S2_) UPP_BLK(NW_, 2, S2_) UPP_BLK(NW_, 3, S2_) UPP_BLK(NW_, 4, S2_) \
  UPP_BLK(NW_, 5, S2_) UPP_BLK(NW_, 6, S2_) UPP_BLK(NW_, 7, S2_) UPP_BLK(NW_, 8, S2_)
  UPP_BLK_ROW(1, 0) UPP_BLK_ROW(2, 0) UPP_BLK_ROW(4, 0) UPP_BLK_ROW(4, 1) UPP_BLK_ROW(8, 0) UPP_BLK_ROW(8, 1)
  UPP_BLK_ROW(16, 0) UPP_BLK_ROW(16, 1)
  UPP_BLK(32, 1, 1) UPP_BLK(32, 2, 1) UPP_BLK(32, 3, 1) UPP_BLK(32, 4, 1)
#undef UPP_BLK_ROW
#undef UPP_BLK
  return UPP_ERR_UNSUPPORTED;
}

// large clouds (2048 < N <= 8192): deferred tree search, 8 or 16 warps, up to 32 points per thread
static int dispatch_fps_blk_big(FpsBlkConfig c, const float* xyz, int B, int N, int M, int32_t* idx,
                                float* centers, cudaStream_t st) {
#define UPP_BIG(NW_, P2_, S2_) \
  if (c.nw == NW_ && c.p2 == P2_ && c.s2 == S2_) return launch_fps_blk<NW_, P2_, S2_, 2>(xyz, B, N, M, idx, centers, st);
#define UPP_BIG_ROW(NW_, S2_) \
  UPP_BIG(NW_, 3, S2_) UPP_BIG(NW_, 4, S2_) UPP_BIG(NW_, 5, S2_) UPP_BIG(NW_, 6, S2_) UPP_BIG(NW_, 7, S2_) UPP_BIG(NW_, 8, S2_)
  UPP_BIG_ROW(4, 0) UPP_BIG_ROW(8, 0) UPP_BIG_ROW(8, 1) UPP_BIG_ROW(16, 0) UPP_BIG_ROW(16, 1)
  UPP_BIG(8, 10, 0) UPP_BIG(8, 12, 0) UPP_BIG(8, 14, 0) UPP_BIG(8, 16, 0)
  UPP_BIG(8, 10, 1) UPP_BIG(8, 12, 1) UPP_BIG(8, 14, 1) UPP_BIG(8, 16, 1)
  UPP_BIG(32, 2, 1) UPP_BIG(32, 3, 1) UPP_BIG(32, 4, 1)
#undef UPP_BIG_ROW
#undef UPP_BIG
  return UPP_ERR_UNSUPPORTED;
}

template <int CS, int NW, int P2, bool AHEAD = false>
static int launch_fps_cluster(const float* xyz, int B, int N, int M, int Nc, int32_t* idx, float* centers,
                              cudaStream_t st) {
  // as for one CTA per cloud: while every CTA can have an SM of its own, keep throughput kernels off that SM --
  // up to 4 CTAs per cluster (measured: a cluster of 8 CTAs asking for the whole 227 KB each is not placed at all)
  const size_t need = static_cast<size_t>(N) * 3 * sizeof(float);
  auto kern = AHEAD ? fps_cluster2_kernel<CS, NW, P2> : fps_cluster_kernel<CS, NW, P2>;
  size_t smem = CS <= 4 ? fps_smem_request(need, B * CS) : need;
  // ... but only if ALL B whole-SM clusters fit at once on this very GPU (GPC sizes differ from chip to chip): a cluster
  // left over runs as a second wave and doubles the launch
  if (smem > need && max_active_clusters(kern, CS, NW * 32, smem) < B) smem = need;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(B) * CS);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, N, M, Nc, idx, centers);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}

// Cluster size for a batch of B clouds of N points, from the B200 sweep (scripts/time_ops.py --sweep-fps-cluster,
// profiles/r01e_sweep_fps_cluster.jsonl; us per round, one CTA per cloud -> cluster):
//   B32 N6144 0.453 -> 0.282 (4)   B32 N4096 0.359 -> 0.254 (4)   B16 N8192 0.543 -> 0.299 (8)   B8 N8192 0.542 -> 0.283 (8)
//   B64 N8192 0.543 -> 0.364 (4, two CTAs per SM)   B32 N2048 0.193 -> 0.231 (slower: the ~300-cycle exchange dominates)
//   mid-size clouds (r01e_sweep_fps_cluster_mid.jsonl, B32): N3584 0.345 -> 0.260 (4), N3072 0.270 -> 0.250, N2560 0.255 -> 0.247
// so: clouds of more than 3072 points (where one CTA per cloud moves to its 8-warp kernels); 8 CTAs when they all
// get an SM of their own and the cloud has >= 5120 points, else 4 CTAs while every cluster is still resident at two
// CTAs per SM.  Two-CTA clusters (slices of up to 4096 points, 32 points per thread) lose and are never chosen.
// UPP_FPS_CLUSTER: 0 = never, 2/4/8 = force that size (tuning / tests).  Returns 0 for "one CTA per cloud".
static int fps_pick_cluster(int B, int N) {
  const int forced = env_int("UPP_FPS_CLUSTER", -1);
  if (forced == 0 || N > kFpsMaxRegPoints) return 0;
  int cs = 0;
  if (forced == 2 || forced == 4 || forced == 8) cs = forced;
  else if (N > 3072) cs = (B * 8 <= kNumSMs && N >= 5120) ? 8 : (B * 4 <= 2 * kNumSMs ? 4 : 0);
  while (cs > 1 && N / cs < 256) cs >>= 1;  // a slice is at least 256 points
  return cs >= 2 ? cs : 0;
}

static int dispatch_fps_cluster(int cs, const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
                                cudaStream_t st) {
  const int nc = (((N + cs - 1) / cs) + 1) & ~1;  // slice length: whole point pairs
  // warps per CTA: 4, or 8 for clusters of up to 4 CTAs (CS * NW <= 32 entries; UPP_FPS_CLUSTER_NW forces 4 / 8);
  // point pairs per thread sized to the slice (instantiated: 1, 2, 3, 4, 6, 8, 12, 16)
  const int nw_env = env_int("UPP_FPS_CLUSTER_NW", 0);
  const int nw = (cs <= 4 && (nw_env == 8 || (nw_env == 0 && cs == 2))) ? 8 : 4;  // measured: 8 warps only pay for 2-CTA clusters
  const int want = (nc + nw * 64 - 1) / (nw * 64);
  const int p2 = want <= 4 ? want : (want <= 6 ? 6 : (want <= 8 ? 8 : (want <= 12 ? 12 : 16)));
  if (p2 > 16 || static_cast<long>(p2) * nw * 64 < nc) return UPP_ERR_UNSUPPORTED;
  // one look-ahead per round (fps_cluster2_kernel): MEASURED, NOT FASTER (see its header) -- off unless UPP_FPS_CLUSTER_AHEAD=1,
  // and instantiated only for the shapes the heuristic itself picks (8 x 4 warps: 2-4 pairs, 4 x 4 warps: 4-8 pairs)
  if (env_int("UPP_FPS_CLUSTER_AHEAD", 0) == 1 && M > 2 && nw == 4) {
#define UPP_CLA(CS_, P2_) \
  if (cs == CS_ && p2 == P2_) return launch_fps_cluster<CS_, 4, P2_, true>(xyz, B, N, M, nc, idx, centers, st);
    UPP_CLA(8, 2) UPP_CLA(8, 3) UPP_CLA(8, 4) UPP_CLA(4, 4) UPP_CLA(4, 6) UPP_CLA(4, 8)
#undef UPP_CLA
  }
#define UPP_CL(CS_, NW_, P2_) \
  if (cs == CS_ && nw == NW_ && p2 == P2_) return launch_fps_cluster<CS_, NW_, P2_>(xyz, B, N, M, nc, idx, centers, st);
#define UPP_CL_ROW(CS_, NW_) UPP_CL(CS_, NW_, 1) UPP_CL(CS_, NW_, 2) UPP_CL(CS_, NW_, 3) UPP_CL(CS_, NW_, 4) \
  UPP_CL(CS_, NW_, 6) UPP_CL(CS_, NW_, 8) UPP_CL(CS_, NW_, 12) UPP_CL(CS_, NW_, 16)
  UPP_CL_ROW(2, 4) UPP_CL_ROW(4, 4) UPP_CL_ROW(8, 4) UPP_CL_ROW(2, 8) UPP_CL_ROW(4, 8)
#undef UPP_CL_ROW
#undef UPP_CL
  return UPP_ERR_UNSUPPORTED;
}

static bool fps_blk_valid(FpsBlkConfig c, int N) {
  const bool nw_ok = c.nw == 1 || c.nw == 2 || c.nw == 4 || c.nw == 8 || c.nw == 16 || c.nw == 32;
  if (c.search == 2) {  // the combinations dispatch_fps_blk_big instantiates
    const bool p_ok = (c.nw == 4 && c.p2 >= 3 && c.p2 <= 8 && c.s2 == 0) ||
                      (c.nw == 8 && ((c.p2 >= 3 && c.p2 <= 8) || c.p2 == 10 || c.p2 == 12 || c.p2 == 14 || c.p2 == 16)) ||
                      (c.nw == 16 && c.p2 >= 3 && c.p2 <= 8) || (c.nw == 32 && c.p2 >= 2 && c.p2 <= 4 && c.s2 == 1);
    return p_ok && (c.s2 == 0 || c.s2 == 1) && static_cast<long>(c.nw) * 64 * c.p2 >= N;
  }
  if (c.search != 0) return false;
  if (!nw_ok || c.p2 < 1 || c.p2 > (c.nw == 32 ? 4 : 8)) return false;
  if (c.s2 != 0 && c.s2 != 1) return false;
  if (c.nw <= 2 && c.s2 != 0) return false;
  if (c.nw == 32 && c.s2 != 1) return false;
  return static_cast<long>(c.nw) * 64 * c.p2 >= N;
}

// Warps x point-pairs per thread for a cloud of N points (N <= kFpsMaxRegPoints); B200 sweep
// (scripts/time_ops.py --sweep-fps2): see DESIGN.md "FPS" for the table this encodes.
FpsBlkConfig fps_pick_blk(int N, int B) {
  int nw;
  if (N <= 256) nw = 1;                               // one warp: no barrier, no shared-memory round trip
  else if (B >= 2 * 148 && N <= 1024) nw = 2;         // several clouds per SM: fewer, fatter warps
  else nw = 4;                                        // one warp per SM sub-partition
  const int p2 = (N + nw * 64 - 1) / (nw * 64);
  return {nw, p2, 0, 0};
}

#define UPP_FPS_CASE_P(T, PP) \
  case PP: return launch_fps_reg<T, PP>(xyz, B, N, M, idx, centers, st);
#define UPP_FPS_CASE_T(T) \
  case T: return launch_fps_reg<T, 2>(xyz, B, N, M, idx, centers, st);

int fps_pruned_launch(const float*, int, int, int, int32_t*, float*, cudaStream_t);  // fps_pruned.cu
int fps_rows_launch(const float*, int, int, int, int32_t*, float*, cudaStream_t);    // fps_pruned.cu

int fps_launch(const float* xyz, int B, int N, int M, int32_t* idx, float* centers,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
  // Exact spatial pruning on a Z-order-sorted copy of the cloud (fps_pruned.cu): EXPERIMENTS, off by default -- bit-identical
  // selections with 85 % of the distance updates skipped, and not faster: the pruned round trades distance work for dependent
  // instructions (0.50-0.65 us per round at N = 8192 against 0.54 for the plain kernel and 0.31 on clusters).
  // UPP_FPS_PRUNED=1 (shared-memory buckets) / 2 (register-resident rows), under UPP_TUNING=1, route clouds of more than
  // UPP_FPS_PRUNED_MIN (2048) points to them: the parity tests do.
  const int pruned = env_int("UPP_FPS_PRUNED", 0);
  if (pruned >= 1 && N > env_int("UPP_FPS_PRUNED_MIN", 2048) && N <= kFpsMaxRegPoints && M >= 32 &&
      env_int("UPP_FPS_CLUSTER", -1) < 1 && env_int("UPP_FPS_NW", 0) == 0 && env_int("UPP_FPS_IMPL", 2) != 1) {
    const int rc = pruned == 2 ? fps_rows_launch(xyz, B, N, M, idx, centers, st) : fps_pruned_launch(xyz, B, N, M, idx, centers, st);
    if (rc != UPP_ERR_UNSUPPORTED) return rc;
  }
  if (const int cs = fps_pick_cluster(B, N)) {  // few large clouds: one cloud per cluster of CTAs
    if (dispatch_fps_cluster(cs, xyz, B, N, M, idx, centers, st) == UPP_OK) return UPP_OK;
    (void)cudaGetLastError();  // a cluster that cannot be placed (or an uninstantiated shape): one CTA per cloud below
  }
  if (N <= kFpsMaxRegPoints && env_int("UPP_FPS_IMPL", 2) != 1) {
    // v2 (blocked ownership, packed fp32x2) everywhere the register-resident scheme reaches; the v1 kernels
    // below stay for A/B timing (UPP_FPS_IMPL=1) and as a parity cross-check.
    FpsBlkConfig c = fps_pick_blk(N, B);
    const FpsBlkConfig forced = {env_int("UPP_FPS_NW", 0), env_int("UPP_FPS_P2", 0), env_int("UPP_FPS_S2", 0),
                                 env_int("UPP_FPS_SEARCH", 0)};
    const bool use_forced = forced.nw > 0 && forced.p2 > 0 && fps_blk_valid(forced, N);  // tuning aid
    if (use_forced) c = forced;
    if (!use_forced && N > 2048) {
      // (A Morton-bucketed variant -- cloud sorted along a Z-curve in the CTA, one spatial bucket per warp, buckets whose
      //  box is farther from the new centre than their largest running min-distance skipped, exact results -- was
      //  built and measured: 0.63 us per round at N = 8192 against 0.54 here; with 32 buckets per cloud too few are
      //  skipped and the per-warp reduction overhead of 32 warps dominates.  profiles/r01d_sweep_fps_big_bucket_experiment.jsonl)
      // large clouds: 8 fat warps (B200 sweep, us per round v1 -> here): N 2500 0.335 -> 0.255, 3000 0.335 -> 0.270,
      // 4096 0.388 -> 0.367, 6144 0.475 -> 0.452, 8192 0.570 -> 0.544 (deferred tree search from 18 points per thread)
      const int p2 = (N + 511) / 512;
      if (N <= 3072) c = {8, p2, 1, 0};
      else if (N <= 4096) c = {8, p2, 0, 0};
      else c = {8, p2 + (p2 & 1), 0, 2};
    }
    if (use_forced || fps_blk_valid(c, N)) {
      if (!fps_blk_valid(c, N)) return UPP_ERR_UNSUPPORTED;
      if (c.search == 2) return dispatch_fps_blk_big(c, xyz, B, N, M, idx, centers, st);
      return dispatch_fps_blk(c, xyz, B, N, M, idx, centers, st);
    }
  }
  if (N <= kFpsMaxRegPoints) {  // v1 kernels (UPP_FPS_IMPL=1): kept for A/B timing and parity tests
    // 4-warp variant: measured better for 1280 < N <= 2048 at any batch, and for every N <= 2048 once
    // the batch is large enough that several clouds share an SM (B=512, N=1024: 108 vs 143 us).
    const int w4 = env_int("UPP_FPS_W4", -1);  // tuning aid: 1 force, 0 forbid
    if (N > 32 && N <= 2048 && (w4 == 1 || (w4 != 0 && (N > 1280 || B >= 2 * 148))))
      return dispatch_fps_w4(round_p((N + 127) / 128), xyz, B, N, M, idx, centers, st);
    const int ft = env_int("UPP_FPS_THREADS", 0), fp = env_int("UPP_FPS_P", 0);
    if (ft >= 32 && ft <= 1024 && ft % 32 == 0 && fp > 0 && round_p(fp) == fp &&
        static_cast<long>(ft) * fp >= N && ft <= (fp > 8 ? 512 : 1024)) {
      switch (fp) {
        case 1: return launch_fps_rt<1>(ft, xyz, B, N, M, idx, centers, st);
        case 2: return launch_fps_rt<2>(ft, xyz, B, N, M, idx, centers, st);
        case 3: return launch_fps_rt<3>(ft, xyz, B, N, M, idx, centers, st);
        case 4: return launch_fps_rt<4>(ft, xyz, B, N, M, idx, centers, st);
        case 6: return launch_fps_rt<6>(ft, xyz, B, N, M, idx, centers, st);
        case 8: return launch_fps_rt<8>(ft, xyz, B, N, M, idx, centers, st);
        case 12: return launch_fps_rt<12>(ft, xyz, B, N, M, idx, centers, st);
        default: return launch_fps_rt<16>(ft, xyz, B, N, M, idx, centers, st);
      }
    }
    const FpsConfig c = fps_pick_config(N);
    if (c.p == 1 && c.threads == 32) return launch_fps_reg<32, 1>(xyz, B, N, M, idx, centers, st);
    if (c.p == 2) {
      switch (c.threads) {
        UPP_FPS_CASE_T(32) UPP_FPS_CASE_T(64) UPP_FPS_CASE_T(96) UPP_FPS_CASE_T(128)
        UPP_FPS_CASE_T(160) UPP_FPS_CASE_T(192) UPP_FPS_CASE_T(224) UPP_FPS_CASE_T(256)
        UPP_FPS_CASE_T(288) UPP_FPS_CASE_T(320) UPP_FPS_CASE_T(352) UPP_FPS_CASE_T(384)
        UPP_FPS_CASE_T(416) UPP_FPS_CASE_T(448) UPP_FPS_CASE_T(480) UPP_FPS_CASE_T(512)
        UPP_FPS_CASE_T(544) UPP_FPS_CASE_T(576) UPP_FPS_CASE_T(608) UPP_FPS_CASE_T(640)
        default: return UPP_ERR_UNSUPPORTED;
      }
    }
    if (c.threads == 512) {
      switch (c.p) {
        UPP_FPS_CASE_P(512, 3) UPP_FPS_CASE_P(512, 4) UPP_FPS_CASE_P(512, 6)
        UPP_FPS_CASE_P(512, 8) UPP_FPS_CASE_P(512, 12) UPP_FPS_CASE_P(512, 16)
        default: return UPP_ERR_UNSUPPORTED;
      }
    }
    return UPP_ERR_UNSUPPORTED;
  }
  const size_t need = static_cast<size_t>(B) * N * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) return UPP_ERR_WORKSPACE;
  fps_global_kernel<1024><<<B, 1024, 0, st>>>(xyz, N, M, idx, centers, static_cast<float*>(workspace));
  count_launch();
  return launch_status();
}

size_t fps_workspace_bytes(int B, int N) {
  if (N <= kFpsMaxRegPoints) return 0;
  return static_cast<size_t>(B) * N * sizeof(float);
}

}  // namespace upp
