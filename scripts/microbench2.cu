// microbench2.cu -- B200 packed-FP32 (FADD2 / FMUL2 / FFMA2, PTX *.f32x2) issue and pipe rates, alone
// and mixed with ALU-pipe min/max, plus the block-reduction round trips the FPS kernels choose from.
// Tuning aid, not product code.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o gpurun_out/microbench2 scripts/microbench2.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }

// MIX: 0 FFMA2 only; 1 FADD2 only; 2 packed pair (3 FADD2 + FMUL2 + 2 FFMA2 = two distances);
//      3 packed pair + 2 FMNMX; 4 packed pair + 1 FMNMX3; 5 FFMA2 + FMNMX 1:1; 6 scalar FFMA + FMNMX 1:1;
//      7 packed pair + 3 FMNMX (single-pass chamfer mix); 8 FFMA2 + 2 FMNMX; 9 scalar FFMA + IMNMX 1:1
//      10 packed pair + 2 FMNMX + 1 VIMNMX3 (FPS mix)
template <int MIX>
__global__ void __launch_bounds__(1024) tput_kernel(long long* out, float* sink, int iters, float a0) {
  u64 acc[8];
  float m[8];
  int key = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = pk(a0 + j + threadIdx.x, a0 - j); m[j] = a0 + j; }
  const u64 q = pk(a0 * 0.5f, a0 * 0.5f), qy = pk(a0 * 0.25f, a0 * 0.25f), qz = pk(a0 * 0.125f, a0 * 0.125f);
  const float fq = a0 * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MIX == 0 || MIX == 5 || MIX == 8) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(q), "l"(qy));
      if (MIX == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[j]) : "l"(q));
      if (MIX == 6 || MIX == 9) { asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(m[j]) : "f"(fq)); }
      if (MIX == 5 || MIX == 6) asm volatile("min.f32 %0, %0, %1;" : "+f"(m[(j + 4) & 7]) : "f"(fq + (float)j));
      if (MIX == 9) asm volatile("max.s32 %0, %0, %1;" : "+r"(key) : "r"(i + j));
      if (MIX == 8) {
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m[(j + 4) & 7]) : "f"(fq));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(m[(j + 5) & 7]) : "f"(fq));
      }
      if (MIX == 2 || MIX == 3 || MIX == 4 || MIX == 7 || MIX == 10) {
        u64 dx, dy, dz, d;
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(acc[j]), "l"(q));
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(acc[(j + 1) & 7]), "l"(qy));
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(acc[(j + 2) & 7]), "l"(qz));
        asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
        asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dx));
        asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
        const float d0 = lo(d), d1 = hi(d);
        if (MIX == 2) m[j] = d0 + 0.f * d1;  // keep both halves live without extra work being counted... (1 FFMA)
        if (MIX == 3) { m[j] = fminf(m[j], d0); m[(j + 1) & 7] = fminf(m[(j + 1) & 7], d1); }
        if (MIX == 4) m[j] = fminf(fminf(m[j], d0), d1);
        if (MIX == 7) { m[j] = fminf(fminf(m[j], d0), d1); m[(j + 3) & 7] = fminf(m[(j + 3) & 7], d0); m[(j + 5) & 7] = fminf(m[(j + 5) & 7], d1); }
        if (MIX == 10) {
          m[j] = fminf(m[j], d0); m[(j + 1) & 7] = fminf(m[(j + 1) & 7], d1);
          key = max(max(key, __float_as_int(m[j])), __float_as_int(m[(j + 1) & 7]));
        }
      }
    }
  }
  long long t1 = clock64();
  float s = (float)key;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += lo(acc[j]) + hi(acc[j]) + m[j];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int MIX>
static int run_tput(const char* name, long long* d_out, float* d_sink, int per_iter_instr) {
  const int iters = 2048;
  for (int nw : {4, 8, 16, 32}) {
    tput_kernel<MIX><<<148, nw * 32>>>(d_out, d_sink, iters, 1.0f);
    tput_kernel<MIX><<<148, nw * 32>>>(d_out, d_sink, iters, 1.0f);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    const double winstr = (double)iters * 8 * per_iter_instr * nw;  // warp-instructions per SM
    printf("tput %-44s nw=%2d  %6.3f warp-instr/cycle/SM  (%.2f cycles per unrolled body per sub-partition-warp)\n", name, nw,
           winstr / (double)h, (double)h / iters / 8 / ((nw + 3) / 4));
  }
  return 0;
}

// ---- block arg-max round trips, W warps: the candidates for the FPS cross-warp stage ----
__device__ __forceinline__ int redux_max(int v) { int r; asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ unsigned redux_min(unsigned v) { unsigned r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }

// VAR 0: REDUX.MAX, REDUX.MIN(idx), STS int2, BAR, LDS (lane<NW), REDUX.MAX, REDUX.MIN, LDS.128 point  (current kernel)
// VAR 1: REDUX.MAX, ballot/ffs, winner lane STS {val},{x,y,z,idx}, BAR, LDS.128 vals (NW<=8: all), tree, LDS.128 point
// VAR 2: as 1 but stage 2 = LDS val (lane<NW), REDUX.MAX, ballot/ffs, LDS.128 point
// VAR 3: NW==4 only: REDUX.MAX, ballot, winner STS.128 {val,idx,x,y}+z, BAR, 4 x LDS.128 + select tree (no second LDS)
template <int VAR>
__global__ void __launch_bounds__(1024) argmax_round(long long* out, int iters, int* sink) {
  __shared__ __align__(16) int s_val[2][32];
  __shared__ __align__(16) float4 s_pt[2][32];
  __shared__ __align__(16) float s_xyz[3 * 1024];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, NW = blockDim.x >> 5;
  for (int i = t; i < 3 * 1024; i += blockDim.x) s_xyz[i] = (float)i;
  __syncthreads();
  float cx = 1.f, cy = 2.f, cz = 3.f;
  int v = t * 7919;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    // stand-in for the distance update: a short dependent FP chain from the centroid to the key
    float d = __fmaf_rn(cz, cz, __fmaf_rn(cx, cx, __fmul_rn(cy, cy)));
    int best = (__float_as_int(d) ^ v) & 0x7fffffff;
    const int b = i & 1;
    int sel;
    if (VAR == 0) {
      const int wb = redux_max(best);
      sel = (int)redux_min(best == wb ? (unsigned)t : 0xffffffffu);
      if (lane == 0) { s_val[b][warp] = wb; reinterpret_cast<int*>(&s_pt[b][warp])[0] = sel; }
      __syncthreads();
      const int sv = lane < NW ? s_val[b][lane] : -1;
      const int si = lane < NW ? reinterpret_cast<int*>(&s_pt[b][lane])[0] : 0x7fffffff;
      const int bb = redux_max(sv);
      sel = (int)redux_min(sv == bb ? (unsigned)si : 0xffffffffu) & 1023;
      cx = s_xyz[3 * sel]; cy = s_xyz[3 * sel + 1]; cz = s_xyz[3 * sel + 2];
    } else {
      const int wb = redux_max(best);
      const unsigned mk = __ballot_sync(0xffffffffu, best == wb);
      if (lane == __ffs(mk) - 1) {
        s_val[b][warp] = wb;
        s_pt[b][warp] = make_float4(cx + lane, cy, cz, __int_as_float(t));
      }
      __syncthreads();
      int w = 0;
      if (VAR == 1) {
        int bv = -1;
        for (int q4 = 0; q4 < NW; q4 += 4) {
          const int4 a = *reinterpret_cast<const int4*>(&s_val[b][q4]);
          if (a.x > bv) { bv = a.x; w = q4; }
          if (q4 + 1 < NW && a.y > bv) { bv = a.y; w = q4 + 1; }
          if (q4 + 2 < NW && a.z > bv) { bv = a.z; w = q4 + 2; }
          if (q4 + 3 < NW && a.w > bv) { bv = a.w; w = q4 + 3; }
        }
      } else if (VAR == 2) {
        const int sv = lane < NW ? s_val[b][lane] : -1;
        const int bb = redux_max(sv);
        w = __ffs(__ballot_sync(0xffffffffu, sv == bb)) - 1;
      }
      const float4 p = s_pt[b][w];
      cx = p.x; cy = p.y; cz = p.z; sel = __float_as_int(p.w);
    }
    v += sel;
  }
  long long t1 = clock64();
  if (t == 0) { out[0] = t1 - t0; sink[0] = v; }
}

template <int VAR>
static int run_round(const char* name, long long* d_out, int* d_sink) {
  const int iters = 4096;
  for (int nw : {1, 2, 4, 8, 16, 20, 32}) {
    argmax_round<VAR><<<1, nw * 32>>>(d_out, iters, d_sink);
    argmax_round<VAR><<<1, nw * 32>>>(d_out, iters, d_sink);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("round %-60s nw=%2d  %6.1f cycles\n", name, nw, (double)h / iters);
  }
  return 0;
}

int main() {
  long long* d_out; float* d_sink; int* d_isink;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_sink, 64)); CK(cudaMalloc(&d_isink, 64));
  run_tput<0>("FFMA2", d_out, d_sink, 1);
  run_tput<1>("FADD2", d_out, d_sink, 1);
  run_tput<2>("2 pairs packed: 3FADD2+FMUL2+2FFMA2 (+1 FFMA)", d_out, d_sink, 7);
  run_tput<3>("2 pairs packed + 2 FMNMX", d_out, d_sink, 8);
  run_tput<4>("2 pairs packed + 1 FMNMX3", d_out, d_sink, 7);
  run_tput<7>("2 pairs packed + 3 FMNMX (1-pass chamfer)", d_out, d_sink, 9);
  run_tput<10>("2 pts packed + 2 FMNMX + VIMNMX3 (FPS)", d_out, d_sink, 9);
  run_tput<5>("FFMA2 + FMNMX 1:1", d_out, d_sink, 2);
  run_tput<8>("FFMA2 + 2 FMNMX", d_out, d_sink, 3);
  run_tput<6>("FFMA + FMNMX 1:1 (scalar)", d_out, d_sink, 2);
  run_tput<9>("FFMA + IMNMX 1:1 (scalar)", d_out, d_sink, 2);
  run_round<0>("REDUX,REDUX,STS,BAR,LDS,REDUX,REDUX,LDSx3 (v1 kernel)", d_out, d_isink);
  run_round<1>("REDUX,ballot,STS,BAR,LDS.128 vals+scan,LDS.128 pt", d_out, d_isink);
  run_round<2>("REDUX,ballot,STS,BAR,LDS,REDUX,ballot,LDS.128 pt", d_out, d_isink);
  return 0;
}
