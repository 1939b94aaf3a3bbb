// group.cu -- the Group divider in ONE launch: farthest-point sampling with the kNN grouping pipelined behind it.
//
// Replaces the body of Group.forward (models/Point_MAE_unify.py:58-92: misc.fps -> KNN -> index gather -> centre
// subtraction).  The two-launch path (fps.cu then knn.cu) serialises two things that are not
// dependent: the neighbours of centre j need centre j only, not the centres after it.  Here one thread-block CLUSTER
// owns one cloud:
//   * every CTA of the cluster stages the cloud once (TMA bulk copy);
//   * the first NW warps of CTA 0 are the PRODUCER: the register-resident FPS chain of fps_blk_kernel (blocked ownership,
//     packed fp32x2 updates, REDUX arg-max, one named barrier among the NW warps per round).  When round j has picked
//     its point, lanes 0..CS-1 of warp 0 publish the index into slot j of EVERY CTA's centre table (one fire-and-forget
//     red.shared::cluster each) and the chain goes on -- nothing waits;
//   * every other warp of the cluster is a CONSUMER: consumer c serves centres c, c + NC, c + 2 NC, ...; it polls slot j
//     of its own CTA's table (one lane, atomic read, back-off by nanosleep), then runs the warp top-k of topk.cuh against
//     the staged cloud and writes centre j's outputs: neighbourhood (already centre-subtracted), centre, indices.
// So the kNN of the first centres runs while the FPS chain is still producing the later ones; the launch ends one kNN
// query after the last FPS round instead of a whole kNN kernel after it.  With few clouds (B * CS <= 148) the CTAs ask
// for a whole SM each, so the latency-bound FPS warps never share issue slots with the throughput-bound consumers; with
// many clouds the cluster is one CTA and producer and consumers share the SM (the SMs are full either way).
// Arithmetic, selection rule and tie-breaks are those of fps.cu / knn.cu: outputs are bit-identical to the two-launch path.
#include <limits.h>

#include "fps_round.cuh"
#include "topk.cuh"

namespace upp {

constexpr int kGroupMaxThreads = 512;

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int NW, int P2, int SLOTS>
__global__ void __launch_bounds__(kGroupMaxThreads, 1)
    group_fused_kernel(const float* __restrict__ xyz, int N, int G, int k, int cw0, int atomic_sync, float* __restrict__ nb_out,
                       float* __restrict__ center_out, int64_t* __restrict__ idx_out,
                       int32_t* __restrict__ cidx_out) {
  constexpr int P = 2 * P2;
  extern __shared__ __align__(16) float s_xyz[];  // 3*N floats (AoS), then the centre table: G ints
  __shared__ __align__(16) int2 s_slot[2][NW];
  __shared__ __align__(8) uint64_t s_bar;
  int* s_cidx = reinterpret_cast<int*>(s_xyz + ((3 * N + 3) & ~3));

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int W = static_cast<int>(blockDim.x) >> 5;
  const unsigned cs = cluster_nctarank();
  const unsigned rank = cs > 1 ? cluster_ctarank() : 0u;
  const int b = blockIdx.x / cs;
  const float* p = xyz + static_cast<size_t>(b) * N * 3;

  pdl_trigger();  // (programmatic dependent launch, common.cuh)
  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  for (int j = t; j < G; j += blockDim.x) s_cidx[j] = j == 0 ? 0 : -1;  // FPS starts at point 0
  __syncthreads();
  pdl_wait();     // the cloud may come from the kernel in front of this one
  unsigned parity = 0;
  stage_points(s_xyz, p, N, &s_bar, parity);
  if (cs > 1) cluster_sync_all();  // every CTA's table is initialised before the producer writes into it

  const bool producer = rank == 0 && warp < NW;
  if (producer) {
    // ---- FPS chain (fps_blk_kernel's round, S2 = 0 / SEARCH = 0), publishing instead of storing ----
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
    // lane r < cs of warp 0 publishes into CTA r's table
    const uint32_t table = map_to_cta(smem_u32(s_cidx), static_cast<unsigned>(lane) < cs ? lane : 0);
    float cx = s_xyz[0], cy = s_xyz[1], cz = s_xyz[2];
    const unsigned lanes_below = (1u << lane) - 1u;
    for (int j = 1; j < G; ++j) {
      const f32x2 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
      int best, ls = 0;
      fps_span<P2, 0, P2, (NW <= 4)>(X, Y, Z, md, CX, CY, CZ, best, ls);
      const int wbest = redux_max_s32(best);
      const unsigned winners = __ballot_sync(0xffffffffu, best == wbest);
      int sel;
      if constexpr (NW == 1) {
        sel = __shfl_sync(0xffffffffu, base + ls, __ffs(winners) - 1);
      } else {
        int2* slot = s_slot[j & 1];
        if (best == wbest && (winners & lanes_below) == 0u) slot[warp] = make_int2(wbest, base + ls);
        named_barrier(1, NW * 32);
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
      }
      if (warp == 0 && static_cast<unsigned>(lane) < cs) {
        if (atomic_sync & 1) publish_cluster_s32(table + 4u * static_cast<unsigned>(j), sel);
        else st_cluster_s32(table + 4u * static_cast<unsigned>(j), sel);
      }
      cx = s_xyz[3 * sel];
      cy = s_xyz[3 * sel + 1];
      cz = s_xyz[3 * sel + 2];
    }
  } else {
    // ---- consumers: kNN + gather - centre for the centres as they appear ----
    // consumer ids interleave the CTAs (consecutive centres go to different SMs): the last centres -- the ones whose
    // kNN nothing overlaps any more -- then spread over all consumer SMs instead of queueing on one
    const int nc = cw0 + (static_cast<int>(cs) - 1) * W;
    const int c = rank == 0 ? warp - NW : cw0 + warp * (static_cast<int>(cs) - 1) + (static_cast<int>(rank) - 1);
    if (c < nc && (rank != 0 || warp - NW < cw0)) {
      for (int j = c; j < G; j += nc) {
        int ci;
        const int sleep_ns = (atomic_sync >> 4) & 0xff;  // tuning: UPP_GROUP_SYNC bits 4..11
        const bool all_lanes = (atomic_sync & 2) != 0;   //         bit 1: every lane polls (plain flavour only)
        do {
          int v = 0;
          if (all_lanes) v = ld_volatile_shared_s32(s_cidx + j);
          else if (lane == 0) v = (atomic_sync & 1) ? poll_shared_s32(s_cidx + j) : ld_volatile_shared_s32(s_cidx + j);
          ci = __shfl_sync(0xffffffffu, v, 0);
          if (ci < 0 && sleep_ns > 0) __nanosleep(sleep_ns);
        } while (ci < 0);
        const float qx = s_xyz[3 * ci], qy = s_xyz[3 * ci + 1], qz = s_xyz[3 * ci + 2];
        DistDirect dist;
        dist.set(qx, qy, qz);
        float ld;
        int li;
        if (SLOTS <= 8 && k <= 8) {
          warp_topk_small<DistDirect, SLOTS>(s_xyz, N, k, dist, ld, li);
        } else {
          TopkState st;
          st.init();
          warp_topk_tile<DistDirect, SLOTS>(s_xyz, N, 0, k, dist, st);
          li = st.li;
        }
        const size_t g = static_cast<size_t>(b) * G + j;
        if (lane < k) {
          const size_t o = g * k + lane;
          if (idx_out) idx_out[o] = static_cast<int64_t>(li);
          nb_out[3 * o + 0] = s_xyz[3 * li] - qx;
          nb_out[3 * o + 1] = s_xyz[3 * li + 1] - qy;
          nb_out[3 * o + 2] = s_xyz[3 * li + 2] - qz;
        }
        if (lane == 0) {
          center_out[3 * g] = qx;
          center_out[3 * g + 1] = qy;
          center_out[3 * g + 2] = qz;
          if (cidx_out) cidx_out[g] = ci;
        }
      }
    }
  }
  if (cs > 1) cluster_sync_all();  // nobody leaves while the producer may still be writing into its table
}

// ---- host side -------------------------------------------------------------------------

constexpr int kNumSMsG = 148;
constexpr size_t kExclusiveSmemG = 232448 - 2048;  // 227 KB opt-in limit minus static + headroom (as fps.cu)

template <int NW, int P2, int SLOTS>
static int launch_group_fused(const float* xyz, int B, int N, int G, int k, float* nb, float* center, int64_t* idx,
                              int32_t* cidx, int cs_forced, int warps, cudaStream_t st) {
  const size_t need = static_cast<size_t>((3 * N + 3) & ~3) * sizeof(float) + static_cast<size_t>(G) * sizeof(int);
  auto kern = group_fused_kernel<NW, P2, SLOTS>;
  const bool may_reserve = need < kExclusiveSmemG && tuning_env_int("UPP_FPS_SHARE_SM", 0) != 1;
  // Cluster size: the consumers of a cloud want SMs of their own (the FPS warps are latency-bound, see fps.cu), so the
  // largest cluster of 4 / 3 / 2 whole-SM CTAs of which ALL B fit at once; failing that one CTA per cloud with producer
  // and consumers side by side (many clouds: the SMs are full either way).  Clusters of 8 measured slower (B16: 20.6 us
  // with 4, 22.7 with 8) and are only reachable through UPP_GROUP_CLUSTER.
  int cs = 1;
  bool exclusive = false;
  if (cs_forced > 0) {
    cs = cs_forced;
    exclusive = may_reserve && cs <= 4 && cs > 1 && max_active_clusters(kern, cs, warps * 32, kExclusiveSmemG) >= B;
  } else if (may_reserve && N > 256) {  // (tiny clouds: one warp of FPS, a handful of queries -- one CTA does it all)
    for (int tryc = 4; tryc >= 2; --tryc) {
      if ((tryc - 1) * warps >= 2 * G && tryc > 2) continue;  // far more consumer warps than centres: a smaller cluster
      if (max_active_clusters(kern, tryc, warps * 32, kExclusiveSmemG) >= B) { cs = tryc; exclusive = true; break; }
    }
  }
  if (cs == 1 && cs_forced <= 0 && B > kNumSMsG && warps > 8) warps = 8;  // several clouds per SM
  const size_t smem = exclusive ? kExclusiveSmemG : need;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (warps <= NW) return UPP_ERR_UNSUPPORTED;
  const int cw0 = cs > 1 ? 0 : warps - NW;  // consumers beside the producer only when the cluster is one CTA
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(B) * cs);
  cfg.blockDim = dim3(warps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tuning_env_int("UPP_PDL", 0) == 1 ? 2 : 1;
  // hand-off flavour: bit 0 atomic publish / poll (else plain store + volatile load), bit 1 every lane polls,
  // bits 4.. back-off in ns while a centre is not there yet (default: atomics, 40 ns)
  const int atomic_sync = tuning_env_int("UPP_GROUP_SYNC", 1 | (40 << 4));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, N, G, k, cw0, atomic_sync, nb, center, idx, cidx);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}

struct FpsBlkConfig {
  int nw, p2, s2, search;
};
FpsBlkConfig fps_pick_blk(int N, int B);  // fps.cu: warps x point pairs per thread for the producer

// UPP_OK when the single-launch kernel took the call; UPP_ERR_UNSUPPORTED when the shape is outside it (the caller then
// composes fps_launch + knn_launch).  Covered: N <= 2048 (every Group call of the UPP configs: 32 ... 2048 points),
// k <= 32, the table + cloud within shared memory.
int group_fused_launch(const float* xyz, int B, int N, int G, int k, float* nb, float* center, int64_t* idx,
                       int32_t* cidx, cudaStream_t st) {
  if (N > 2048 || k > 32 || G < 1 || G > 4096) return UPP_ERR_UNSUPPORTED;
  if (tuning_env_int("UPP_GROUP_FUSED", 1) == 0) return UPP_ERR_UNSUPPORTED;  // A/B: the two-launch path
  const FpsBlkConfig c = fps_pick_blk(N, B);
  if (c.nw > 4 || c.p2 > 8 || static_cast<long>(c.nw) * 64 * c.p2 < N) return UPP_ERR_UNSUPPORTED;
  // cluster size: chosen per launch from the occupancy of this very GPU (launch_group_fused); UPP_GROUP_CLUSTER /
  // UPP_GROUP_WARPS force it (tests / tuning)
  const int forced = tuning_env_int("UPP_GROUP_CLUSTER", 0);
  const int cs = (forced == 1 || forced == 2 || forced == 3 || forced == 4 || forced == 8) ? forced : 0;
  int warps = 16;
  const int fw = tuning_env_int("UPP_GROUP_WARPS", 0);
  if (fw == 8 || fw == 16) warps = fw;
  if (warps <= c.nw) return UPP_ERR_UNSUPPORTED;
#define UPP_G(NW_, P2_, SL_) \
  if (c.nw == NW_ && c.p2 == P2_) return launch_group_fused<NW_, P2_, SL_>(xyz, B, N, G, k, nb, center, idx, cidx, cs, warps, st);
  UPP_G(1, 1, 4) UPP_G(1, 2, 4) UPP_G(1, 3, 8) UPP_G(1, 4, 8)
  UPP_G(2, 1, 32) UPP_G(2, 2, 32) UPP_G(2, 3, 32) UPP_G(2, 4, 32) UPP_G(2, 5, 32) UPP_G(2, 6, 32) UPP_G(2, 7, 32) UPP_G(2, 8, 32)
  UPP_G(4, 1, 32) UPP_G(4, 2, 32) UPP_G(4, 3, 32) UPP_G(4, 4, 32) UPP_G(4, 5, 32) UPP_G(4, 6, 32) UPP_G(4, 7, 32) UPP_G(4, 8, 32)
#undef UPP_G
  return UPP_ERR_UNSUPPORTED;
}

}  // namespace upp
