// chamfer.cu -- Chamfer distance forward / backward for sm_100a.
//
// Replaces extensions/chamfer_dist/chamfer.cu: chamfer_dist_kernel (:15-145, launched twice with a
// fixed <<<(32,16),512>>> grid, one query per thread, one LDS per 1.3 pairs) and
// chamfer_dist_grad_kernel (:173-201, <<<(1,16),256>>> = 16 CTAs with the batch loop serial
// inside and six float atomics per point).
//
// Forward here: ONE launch covers both directions (blockIdx.z) and the whole batch (blockIdx.y).
// The reference cloud is staged in 2048-point tiles by TMA bulk copy; each thread owns R queries
// in registers and reads the tile as broadcast LDS.128 (3 loads feed 4 refs x R queries), so the
// inner loop is issue bound on the distance itself: 3 FADD + FMUL + 2 FFMA + 0.5 FMNMX3 per pair.
// d = fma(dz,dz,fma(dx,dx,dy*dy)), d* = ref - query; strict '<' in ascending ref order gives the
// lowest index on ties (== reference).  The grid is sized from the problem, not fixed.
//
// Backward here: pass 1 WRITES each point's own term (no atomics, no memset), pass 2 scatters the
// partner term with float RED.ADD -- half the reference's atomics, all B*(N+M) points in flight.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace upp {

constexpr int kChTile = 2048;  // reference points per shared-memory tile (24 KB)

// Min-tracking costs as much as the distance itself if done per pair (FSETP+FSEL+SEL ~ 4.4 issue
// cycles vs 6 for the distance, scripts/microbench.cu), so the inner loop tracks per GROUP of 8
// refs: two FMNMX3 chains fold 8 distances into one min (0.5 instr/pair), ONE compare/select pair
// per group keeps (best value, group base).  The arg-min inside the winning group is recovered once
// per query in the epilogue by recomputing its 8 distances (first one equal to the best value).
// Strict '<' across groups in ascending order + first match inside => lowest index on ties.
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
    chamfer_fwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int M,
                       float* __restrict__ dist1, float* __restrict__ dist2,
                       int32_t* __restrict__ idx1, int32_t* __restrict__ idx2) {
  __shared__ __align__(16) float s_ref[kChTile * 3];
  __shared__ __align__(8) uint64_t s_bar;
  const int dir = blockIdx.z;
  const int b = blockIdx.y;
  const int nq = dir == 0 ? N : M;  // queries
  const int nr = dir == 0 ? M : N;  // references
  const int q0 = blockIdx.x * (THREADS * R);
  if (q0 >= nq) return;  // whole CTA leaves together (before any barrier)
  const float* qp = (dir == 0 ? xyz1 : xyz2) + static_cast<size_t>(b) * nq * 3;
  const float* rp = (dir == 0 ? xyz2 : xyz1) + static_cast<size_t>(b) * nr * 3;
  float* dout = (dir == 0 ? dist1 : dist2) + static_cast<size_t>(b) * nq;
  int32_t* iout = (dir == 0 ? idx1 : idx2) + static_cast<size_t>(b) * nq;
  const int t = threadIdx.x;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;

  float qx[R], qy[R], qz[R], best[R];
  int bg[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = q0 + r * THREADS + t;
    const int jj = j < nq ? j : nq - 1;  // clamp: computes a duplicate, never stored
    qx[r] = __ldg(qp + 3 * jj);
    qy[r] = __ldg(qp + 3 * jj + 1);
    qz[r] = __ldg(qp + 3 * jj + 2);
    best[r] = __int_as_float(0x7f800000);
    bg[r] = 0;
  }

  for (int base = 0; base < nr; base += kChTile) {
    const int tile = min(kChTile, nr - base);
    const int tile8 = (tile + 7) & ~7;
    if (base > 0) __syncthreads();
    // pad the last group of eight with NaN: (NaN - q)^2 = NaN, and fminf() drops NaN operands
    for (int i = tile * 3 + t; i < tile8 * 3; i += THREADS) s_ref[i] = __int_as_float(0x7fc00000);
    stage_points(s_ref, rp + static_cast<size_t>(base) * 3, tile, &s_bar, parity);

    const float4* s4 = reinterpret_cast<const float4*>(s_ref);
    for (int k = 0; k < tile8; k += 8) {
      float m[R];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = s4[(k >> 2) * 3 + h * 3 + 0];  // x0 y0 z0 x1
        const float4 c = s4[(k >> 2) * 3 + h * 3 + 1];  // y1 z1 x2 y2
        const float4 e = s4[(k >> 2) * 3 + h * 3 + 2];  // z2 x3 y3 z3
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float d0 = dist_yxz(a.x - qx[r], a.y - qy[r], a.z - qz[r]);
          const float d1 = dist_yxz(a.w - qx[r], c.x - qy[r], c.y - qz[r]);
          const float d2 = dist_yxz(c.z - qx[r], c.w - qy[r], e.x - qz[r]);
          const float d3 = dist_yxz(e.y - qx[r], e.z - qy[r], e.w - qz[r]);
          if (h == 0) m[r] = fminf(fminf(fminf(d0, d1), d2), d3);
          else m[r] = fminf(fminf(fminf(fminf(m[r], d0), d1), d2), d3);
        }
      }
      const int kk = base + k;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (m[r] < best[r]) { best[r] = m[r]; bg[r] = kk; }
    }
  }
  // epilogue: arg-min inside the winning group of 8 (refs re-read from global / L2)
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = q0 + r * THREADS + t;
    if (j < nq) {
      int bi = bg[r];
      const int lim = min(8, nr - bg[r]);
      for (int u = lim - 1; u >= 0; --u) {
        const float* p = rp + static_cast<size_t>(bg[r] + u) * 3;
        const float d = dist_yxz(__ldg(p) - qx[r], __ldg(p + 1) - qy[r], __ldg(p + 2) - qz[r]);
        if (d == best[r]) bi = bg[r] + u;
      }
      dout[j] = best[r];
      iout[j] = bi;
    }
  }
}

// Deterministic whole-call sums { sum d1, sum d2, sum sqrt d1, sum sqrt d2 } -- the send buffer of
// the one NCCL all-reduce the batch-sharded loss needs.  One thread-block CLUSTER of 8 CTAs: each
// CTA reduces a fixed interleaved slice (4 independent accumulators per quantity keep loads in
// flight), leaves 4 floats in its shared memory, and after one cluster barrier CTA 0 adds the 8
// partials in rank order through distributed shared memory.  Fixed partition + fixed order =>
// run-to-run identical value, no scratch buffer, no atomics.
constexpr int kSumCluster = 8;
constexpr int kSumThreads = 512;

__global__ void __cluster_dims__(kSumCluster, 1, 1) __launch_bounds__(kSumThreads)
    chamfer_sums_kernel(const float* __restrict__ dist1, size_t n1, const float* __restrict__ dist2,
                        size_t n2, float* __restrict__ sums) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_part[4][kSumThreads / 32];
  __shared__ float s_out[4];
  const unsigned rank = cluster.block_rank();
  const size_t stride = static_cast<size_t>(kSumCluster) * kSumThreads;
  const size_t first = static_cast<size_t>(rank) * kSumThreads + threadIdx.x;
  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[a][q][u] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const float* d = a == 0 ? dist1 : dist2;
    const size_t n = a == 0 ? n1 : n2;
    for (size_t i = first; i < n; i += 4 * stride) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t j = i + u * stride;
        const float v = j < n ? __ldg(d + j) : 0.f;
        acc[a][0][u] += v;
        acc[a][1][u] += __fsqrt_rn(v);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float v = warp_sum((acc[a][q][0] + acc[a][q][1]) + (acc[a][q][2] + acc[a][q][3]));
      if (lane == 0) s_part[q * 2 + a][warp] = v;
    }
  __syncthreads();
  if (warp < 4) {
    const float v = warp_sum(lane < kSumThreads / 32 ? s_part[warp][lane] : 0.f);
    if (lane == 0) s_out[warp] = v;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < 4) {
    float tot = 0.f;
    for (unsigned r = 0; r < kSumCluster; ++r) tot += *cluster.map_shared_rank(&s_out[threadIdx.x], r);
    sums[threadIdx.x] = tot;
  }
  cluster.sync();  // keep every CTA's shared memory alive until CTA 0 has read it
}

// Backward pass 1: own term, plain stores.  grad_a[j] = 2 g[j] (a_j - b_idx[j]) for both clouds.
__global__ void __launch_bounds__(256)
    chamfer_bwd_own_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                           const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                           const float* __restrict__ g1, const float* __restrict__ g2, int B, int N,
                           int M, float* __restrict__ gx1, float* __restrict__ gx2) {
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const bool first = i < n1;
    const size_t p = first ? i : i - n1;
    const int na = first ? N : M, nb = first ? M : N;
    const size_t b = p / na;
    const float* a = (first ? xyz1 : xyz2) + p * 3;
    const int j2 = (first ? idx1 : idx2)[p];
    const float* o = (first ? xyz2 : xyz1) + (b * nb + j2) * 3;
    const float g = __fmul_rn((first ? g1 : g2)[p], 2.0f);
    float* out = (first ? gx1 : gx2) + p * 3;
    out[0] = __fmul_rn(g, a[0] - o[0]);
    out[1] = __fmul_rn(g, a[1] - o[1]);
    out[2] = __fmul_rn(g, a[2] - o[2]);
  }
}

// Backward pass 2: partner term, scattered with RED.ADD.F32.  grad_b[idx[j]] += -(2 g[j] (a_j - b_idx[j])).
__global__ void __launch_bounds__(256)
    chamfer_bwd_scatter_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                               const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                               const float* __restrict__ g1, const float* __restrict__ g2, int B,
                               int N, int M, float* __restrict__ gx1, float* __restrict__ gx2) {
  const size_t n1 = static_cast<size_t>(B) * N, n2 = static_cast<size_t>(B) * M;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const bool first = i < n1;
    const size_t p = first ? i : i - n1;
    const int na = first ? N : M, nb = first ? M : N;
    const size_t b = p / na;
    const float* a = (first ? xyz1 : xyz2) + p * 3;
    const int j2 = (first ? idx1 : idx2)[p];
    const size_t op = (b * nb + j2) * 3;
    const float* o = (first ? xyz2 : xyz1) + op;
    const float g = __fmul_rn((first ? g1 : g2)[p], 2.0f);
    float* out = (first ? gx2 : gx1) + op;
    atomicAdd(out + 0, -__fmul_rn(g, a[0] - o[0]));
    atomicAdd(out + 1, -__fmul_rn(g, a[1] - o[1]));
    atomicAdd(out + 2, -__fmul_rn(g, a[2] - o[2]));
  }
}

template <int R, int THREADS>
static void launch_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                               float* dist2, int32_t* idx1, int32_t* idx2, cudaStream_t st) {
  const int per = R * THREADS;
  dim3 grid((max(N, M) + per - 1) / per, B, 2);
  chamfer_fwd_kernel<R, THREADS><<<grid, THREADS, 0, st>>>(xyz1, xyz2, N, M, dist1, dist2, idx1, idx2);
}

int chamfer_fwd_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* sums, cudaStream_t st) {
  const char* v = getenv("UPP_CH_VARIANT");  // tuning aid: "R,T"
  const int variant = v ? atoi(v) : 0;
  switch (variant) {
    case 1: launch_chamfer_fwd<4, 64>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 2: launch_chamfer_fwd<8, 64>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 3: launch_chamfer_fwd<8, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 4: launch_chamfer_fwd<4, 256>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 5: launch_chamfer_fwd<4, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 6: launch_chamfer_fwd<2, 256>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 7: launch_chamfer_fwd<2, 64>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 8: launch_chamfer_fwd<1, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    case 9: launch_chamfer_fwd<1, 256>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    // measured best on B200 (scripts/time_ops.py --sweep-chamfer): more warps beat more queries per thread
    default: launch_chamfer_fwd<2, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
  }
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK || sums == nullptr) return rc;
  chamfer_sums_kernel<<<kSumCluster, kSumThreads, 0, st>>>(dist1, static_cast<size_t>(B) * N, dist2,
                                                          static_cast<size_t>(B) * M, sums);
  count_launch();
  return launch_status();
}

int chamfer_bwd_launch(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                       const float* g1, const float* g2, int B, int N, int M, float* gx1, float* gx2,
                       cudaStream_t st) {
  const size_t total = static_cast<size_t>(B) * (static_cast<size_t>(N) + M);
  const size_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want > 148 * 16 ? 148 * 16 : want);
  chamfer_bwd_own_kernel<<<blocks, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, g1, g2, B, N, M, gx1, gx2);
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  chamfer_bwd_scatter_kernel<<<blocks, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, g1, g2, B, N, M, gx1, gx2);
  count_launch();
  return launch_status();
}

}  // namespace upp
