// interp.cu -- k-nearest inverse-distance feature interpolation for sm_100a.
//
// Replaces the body of the reference's pure-torch
//   propagate()                         models/Point_MAE_unify.py:22-48      (k = de_neighbors, eps 1e-8, x0.3 + points1)
//   PointNetFeaturePropagation.forward  models/Point_MAE_unify_segment.py:289-313, models/Point_MAE_pretask_dev.py:437-461
//                                                                            (k = interpolate_neighbors, eps 1e-4)
// which build the full (B,N,S) square_distance matrix, SORT every row, slice k columns, materialise
// index_points(points2, idx) as a (B,N,k,C) tensor, multiply by the weights and sum.
//
// Here, forward is ONE launch: the S source points of a cloud are staged once per CTA by TMA bulk copy,
// one warp per target point selects its k nearest sources with the register-resident warp top-k of
// topk.cuh on the reference's own distance form (square_distance: -2 a.b + |a|^2 + |b|^2), turns them
// into weights (1/(d+eps), normalised) and immediately gathers the k feature rows (float4, coalesced
// over channels) into the output row -- no distance matrix, no sort, no (B,N,k,C) tensor.
// Backward is deterministic and atomic-free: a target-side kernel (one warp per target) produces
// d loss / d dist and grad_xyz1; a source-side kernel (one CTA per source point) scans the cloud's
// (N,k) index list in order and accumulates grad_points2 / grad_xyz2 rows in registers.
#include "topk.cuh"

namespace upp {

// out[b,n,:] = (base ? base[b,n,:] : 0) + alpha * sum_j w_j feat2[b, idx_j, :]
// Gather: channels in super-blocks of 8 x 128 (one float4 per lane and block, 8 independent LDG.128 in flight per
// neighbour), neighbours four at a time with their weight / row pointer shuffled once per super-block; products and
// sums as packed fp32x2 (FMUL2 / FADD2: bit-identical to the scalar mul-then-add of the torch expression).
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

int interp_fwd_launch(const float* xyz1, const float* xyz2, const float* feat2, const float* base, float alpha,
                      float eps, int B, int N, int S, int C, int k, float* out, int32_t* idx, float* weight,
                      float* distk, cudaStream_t st) {
  // targets per warp: as many as keep >= 4 residency waves (3 CTAs x 148 SMs) of CTAs in the grid
  int tpw = 1;
  if (S <= kKnnTile)
    for (int cand = 8; cand > 1; cand >>= 1)
      if (static_cast<long>(B) * ((N + kKnnWarps * cand - 1) / (kKnnWarps * cand)) >= 4L * 3 * 148) { tpw = cand; break; }
  dim3 grid((N + kKnnWarps * tpw - 1) / (kKnnWarps * tpw), B);
  const size_t smem = static_cast<size_t>(min(S, kKnnTile)) * 3 * sizeof(float);
  const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat2) | reinterpret_cast<uintptr_t>(out) |
                                      reinterpret_cast<uintptr_t>(base)) % 16 == 0);
  const int threads = kKnnWarps * kWarp;
#define UPP_INTERP(V_, SL_) \
  interp_fwd_kernel<V_, SL_><<<grid, threads, smem, st>>>(xyz1, xyz2, feat2, base, alpha, eps, N, S, C, k, tpw, out, idx, weight, distk)
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
  count_launch();
  return launch_status();
}

int interp_bwd_launch(const float* gout, const int32_t* idx, const float* weight, const float* distk,
                      const float* feat2, const float* xyz1, const float* xyz2, float alpha, float eps, int B,
                      int N, int S, int C, int k, float* gfeat2, float* gxyz1, float* gxyz2, float* gd_ws,
                      cudaStream_t st) {
  const bool want_xyz = gd_ws != nullptr;
  if (want_xyz) {
    const size_t rows = static_cast<size_t>(B) * N;
    const size_t want = (rows + 7) / 8;
    const int blocks = static_cast<int>(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
    const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(feat2)) % 16 == 0);
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

}  // namespace upp
