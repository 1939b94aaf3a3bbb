// common.cuh -- shared device helpers for libupp_geom (sm_100a only).
#pragma once
#include <cuda.h>  // CUtensorMap: types only -- the encoder is fetched from the driver at run time (no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "upp_geom.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libupp_geom is written for sm_100a (B200) only"
#endif

namespace upp {

constexpr int kWarp = 32;

// ---- launch accounting (bench.py's gpu_launches, tests) -------------------------------
void count_launch(int n = 1);

// ---- tuning / A-B switches --------------------------------------------------------------
// The UPP_FPS_* / UPP_CH_* / UPP_INTERP_* / UPP_BLEND_* / UPP_GROUP_* environment switches select kernel variants
// for timing sweeps and for the parity tests that force every variant.  They are honoured ONLY when
// UPP_TUNING=1 is set as well: a stray variable in a user's environment never changes the product path.
const char* tuning_env(const char* name);
int tuning_env_int(const char* name, int dflt);

// ---- distance forms: spelled with explicit roundings, never left to contraction -------
// chamfer.cu:40-43 / upstream sampling_gpu.cu: nvcc contracts dx*dx + dy*dy + dz*dz to this.
__device__ __forceinline__ float dist_yxz(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// KNN_CUDA knn.cu: ssd = 0; ssd += tmp*tmp over x,y,z.  fma(dx,dx,0) == rn(dx*dx).
__device__ __forceinline__ float dist_xyz_acc(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) ---------------------------
// One issue slot performs the IEEE-rn operation on both 32-bit halves of a 64-bit register pair:
// results are bit-identical to the scalar __fsub_rn/__fmul_rn/__fmaf_rn forms above, at half the
// issue slots -- which is what binds these kernels (scripts/microbench2.cu).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// two distances at once, each half == dist_yxz / dist_xyz_acc of the scalar forms
__device__ __forceinline__ f32x2 dist2_yxz(f32x2 dx, f32x2 dy, f32x2 dz) {
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}
__device__ __forceinline__ f32x2 dist2_xyz_acc(f32x2 dx, f32x2 dy, f32x2 dz) {
  return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}

// ---- warp reductions in one instruction (REDUX) ----------------------------------------
__device__ __forceinline__ int redux_max_s32(int v) {
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ unsigned redux_min_u32(unsigned v) {
  unsigned r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ unsigned redux_max_u32(unsigned v) {
  unsigned r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 2-D TMA tile load (cp.async.bulk.tensor -> SASS UTMALDG): the box of `map` whose first element is
// (c_inner, c_row) lands densely in shared memory (row after row, dst 128-byte aligned) and completes
// box-bytes on `bar`; elements outside the tensor arrive as zeros and still count.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c_inner, int c_row,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_row), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Host: row-major fp32 matrix (rows x inner, row pitch in bytes a multiple of 16, base 16-byte aligned)
// as a tensor map with a (box_rows x box_inner) box, no swizzle, zero fill.  Returns UPP_OK or an error code.
int make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_rows);

// Stage `npts` xyz triples (AoS, 12 B each) from global into shared memory.
// The 16-byte-aligned body goes through one TMA bulk copy issued by thread 0; a misaligned
// source or the (<4 point) tail is copied with plain loads.  All threads of the CTA must
// call this; `parity` is the caller-tracked phase bit of `bar` (flipped here when used).
// On return the data is visible to every thread.
__device__ __forceinline__ void stage_points(float* dst, const float* __restrict__ src, int npts,
                                             uint64_t* bar, unsigned& parity) {
  const int nfl = npts * 3;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  const int body = aligned ? (nfl & ~3) : 0;  // floats moved by TMA (multiple of 16 B)
  if (body > 0 && threadIdx.x == 0) {
    mbar_expect_tx(bar, static_cast<unsigned>(body) * 4u);
    bulk_g2s(dst, src, static_cast<unsigned>(body) * 4u, bar);
  }
  for (int i = body + threadIdx.x; i < nfl; i += blockDim.x) dst[i] = __ldg(src + i);
  if (body > 0) {
    mbar_wait(bar, parity);
    parity ^= 1u;
  }
  __syncthreads();
}

// ---- thread-block clusters / distributed shared memory --------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, int lo, int hi, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
               "r"(lo), "r"(hi), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b128(uint32_t remote_addr, int a, int b, int c, int d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// Flag-style publish / poll of one 32-bit word in (distributed) shared memory, both sides ATOMIC so that the pair is a
// synchronising access in the memory model (and in compute-sanitizer racecheck) rather than a racing store / load:
//   publisher: red.max into the word of CTA `rank`-mapped address (mapa result) -- fire and forget, the slot starts at -1
//              and receives a value >= 0 exactly once;
//   poller:    atom.max with -1, i.e. an atomic read.
__device__ __forceinline__ void publish_cluster_s32(uint32_t remote_addr, int v) {
  asm volatile("red.relaxed.cluster.shared::cluster.max.s32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
// (the plain-store / volatile-load flavour of the same hand-off, kept for A/B timing: UPP_GROUP_SYNC=0)
__device__ __forceinline__ void st_cluster_s32(uint32_t remote_addr, int v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_shared_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ int poll_shared_s32(const int* p) {
  int v;
  asm volatile("atom.relaxed.cluster.shared::cta.max.s32 %0, [%1], -1;" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// How many clusters of `cs` CTAs of this kernel can be resident at once with `smem` bytes of dynamic shared memory per
// CTA (cudaOccupancyMaxActiveClusters; cached per configuration).  The GPC sizes of a B200 differ from chip to chip
// (floor-sweeping), so "32 clusters of 4 whole-SM CTAs fit on 148 SMs" holds on one board and not on the next: a
// cluster that does not fit runs as a second wave and the launch takes twice as long.
template <class Kern>
static int max_active_clusters(Kern kern, int cs, int threads, size_t smem) {
  struct Key { int cs, threads; size_t smem; int n; };
  static std::mutex mu;
  static std::vector<Key> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (const Key& k : cache)
    if (k.cs == cs && k.threads == threads && k.smem == smem) return k.n;
  int n = 0;
  if (smem <= 40 * 1024 ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(cs) * 148);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
  }
  cache.push_back({cs, threads, smem, n});
  return n;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// The chains on this path are runs of short DEPENDENT kernels (FPS -> FPS -> Group -> Group -> four gradient kernels in
// the UPP step), each edge costing ~2.4 us of launch latency in a CUDA graph.  A kernel launched with the programmatic
// stream-serialization attribute may start (be scheduled, run its prologue) while its predecessor is still running; it
// calls pdl_wait() before it touches anything the predecessor wrote (griddepcontrol.wait returns when the prerequisite
// grids have completed and their memory is visible).  pdl_trigger() lets this grid's own dependents start early; it is
// issued first thing, which is always safe because they wait the same way.  Kernels launched WITHOUT the attribute
// execute both instructions as no-ops, and a predecessor that never triggers (a torch kernel) simply releases its
// dependents when it exits.
// MEASURED AND LEFT OFF (round 2, bench.py headline step, two runs each): device-resident 0.2859 -> 0.2887 ms without it
// (-1 %), but the end-to-end arm -- H2D copies and autograd's backward thread in the same graph -- went from 0.324 to
// 0.376 ms with it: dependents that are resident early and parked in griddepcontrol.wait hold whole-SM shared-memory
// reservations the other branches of the step are waiting for.  UPP_PDL=1 (under UPP_TUNING=1) turns it on for A/B.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tuning_env_int("UPP_PDL", 0) == 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- host-side launch check -----------------------------------------------------------
inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UPP_OK : static_cast<int>(e);
}

}  // namespace upp
