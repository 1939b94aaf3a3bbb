// microbench.cu -- B200 primitive latencies / throughputs that shape the FPS, kNN and Chamfer
// kernels (tuning aid, not product code).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o gpurun_out/microbench scripts/microbench.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ int redux_max(int v) { int r; asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ unsigned redux_min(unsigned v) { unsigned r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }

// ---- dependent-chain latencies (one warp) ----
template <int OP>
__global__ void lat_kernel(long long* out, int iters, int seed) {
  __shared__ int sm[64];
  __shared__ unsigned long long sm64[4];
  int v = seed + threadIdx.x;
  float f = (float)v;
  sm[threadIdx.x & 63] = v;
  if (threadIdx.x < 4) sm64[threadIdx.x] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) v = redux_max(v) + 1;
    if (OP == 1) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    if (OP == 2) v = __popc(__ballot_sync(0xffffffffu, v & 1)) + v;
    if (OP == 3) v = sm[(v & 31)] + 1;                                  // LDS dependent
    if (OP == 4) { sm[threadIdx.x & 31] = v; __syncwarp(); v = sm[(threadIdx.x + 1) & 31] + 1; __syncwarp(); }  // STS->LDS
    if (OP == 5) { asm volatile("min.f32 %0, %0, %1;" : "+f"(f) : "f"((float)i)); }
    if (OP == 6) v = max(v, i) + 1;
    if (OP == 7) v = atomicMax(&sm[0], v) + 1;                          // ATOMS.MAX.32 with return
    if (OP == 8) { f = __fmaf_rn(f, f, 1.0f); }
    if (OP == 9) v = (int)redux_min((unsigned)redux_max(v)) + 1;        // two dependent REDUX
    if (OP == 10) v = __shfl_sync(0xffffffffu, v, 3) + 1;
    if (OP == 11) v = __shfl_up_sync(0xffffffffu, v, 1) + 1;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v + (int)f; }
}

// ---- block barrier round trip: STS, BAR, LDS, repeated; W warps ----
__global__ void bar_kernel(long long* out, int iters) {
  __shared__ int sm[2][32];
  int v = threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (lane == 0) sm[i & 1][warp] = v;
    __syncthreads();
    v = sm[i & 1][lane & ((blockDim.x >> 5) - 1)] + 1;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
__global__ void bar_only_kernel(long long* out, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; }
}

// ---- throughput: NW warps per CTA, 1 CTA per SM, unrolled independent ops ----
// MIX: 0 = FFMA 3 distinct regs; 1 = FFMA a*a+c; 2 = FADD; 3 = chamfer pair (3 FADD, FMUL, 2 FFMA);
//      4 = chamfer pair + FSETP/FSEL/SEL tracking; 5 = chamfer pair + FMNMX only; 6 = FMNMX only; 7 = pair + 0.5 FMNMX3
//      8 = independent REDUX.MIN (throughput); 9 = independent SHFL.bfly; 10 = independent ballot
template <int MIX>
__global__ void __launch_bounds__(1024) tput_kernel(long long* out, float* sink, int iters, float a0) {
  float acc[8];
  int bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = a0 + j + threadIdx.x; bi[j] = 0; }
  float qx = a0 * 0.5f, qy = a0 * 0.25f, qz = a0 * 0.125f;
  float rx = a0 + threadIdx.x * 1e-3f, ry = rx * 1.5f, rz = rx * 2.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MIX == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j]) : "f"(qx), "f"(qy));
      if (MIX == 1) asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(acc[j]) : "f"(qx));
      if (MIX == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc[j]) : "f"(qx));
      if (MIX >= 3 && MIX != 6) {
        float dx, dy, dz, d;
        const float sx = rx + (float)j, sy = ry, sz = rz;
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dx) : "f"(sx), "f"(qx));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dy) : "f"(sy), "f"(acc[j]));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dz) : "f"(sz), "f"(qz));
        asm volatile("mul.rn.f32 %0, %1, %1;" : "=f"(d) : "f"(dy));
        asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(d) : "f"(dx));
        asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(d) : "f"(dz));
        if (MIX == 3) acc[j] = d;
        if (MIX == 4) { if (d < acc[j]) { acc[j] = d; bi[j] = i; } }
        if (MIX == 5) acc[j] = fminf(acc[j], d);
        if (MIX == 7) { if (j & 1) acc[j] = fminf(fminf(acc[j], acc[j - 1]), d); else acc[j] = d; }
      }
      if (MIX == 6) asm volatile("min.f32 %0, %0, %1;" : "+f"(acc[j]) : "f"(qx));
      if (MIX == 8) { unsigned r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_uint(acc[j]) + i)); bi[j] += r; }
      if (MIX == 9) { bi[j] += __shfl_xor_sync(0xffffffffu, bi[j] + i, 1); }
      if (MIX == 10) { bi[j] += __ballot_sync(0xffffffffu, (bi[j] + i) & 1); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j] + bi[j];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
static int run_lat(const char* name, long long* d_out) {
  const int iters = 4096;
  lat_kernel<OP><<<1, 32>>>(d_out, iters, 1);
  lat_kernel<OP><<<1, 32>>>(d_out, iters, 1);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("latency %-28s %7.1f cycles/op\n", name, (double)h[0] / iters);
  return 0;
}

template <int MIX>
static int run_tput(const char* name, long long* d_out, float* d_sink, int per_iter_instr) {
  const int iters = 2048;
  for (int nw : {4, 8, 16, 32}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput_kernel<MIX><<<148, nw * 32>>>(d_out, d_sink, iters, 1.0f);
    cudaEventRecord(e0);
    tput_kernel<MIX><<<148, nw * 32>>>(d_out, d_sink, iters, 1.0f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h;
    CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    const double winstr = (double)iters * 8 * per_iter_instr * nw;  // warp-instructions per SM
    printf("tput %-34s nw=%2d  %6.3f warp-instr/cycle/SM  (%.0f cycles, %.1f us, clock %.0f MHz)\n", name, nw,
           winstr / (double)h, (double)h, ms * 1e3, (double)h / (ms * 1e3));
  }
  return 0;
}

int main() {
  long long* d_out; float* d_sink;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_sink, 64));
  run_lat<0>("REDUX.max.s32", d_out);
  run_lat<9>("REDUX+REDUX dependent", d_out);
  run_lat<1>("SHFL.bfly", d_out);
  run_lat<10>("SHFL.idx", d_out);
  run_lat<11>("SHFL.up", d_out);
  run_lat<2>("ballot+popc", d_out);
  run_lat<3>("LDS dependent", d_out);
  run_lat<4>("STS->syncwarp->LDS", d_out);
  run_lat<5>("FMNMX", d_out);
  run_lat<6>("IMNMX+IADD", d_out);
  run_lat<7>("ATOMS.MAX.32 (ret)", d_out);
  run_lat<8>("FFMA", d_out);
  for (int nw : {1, 2, 4, 8, 16, 32}) {
    const int iters = 2048;
    long long h[2];
    bar_kernel<<<1, nw * 32>>>(d_out, iters);
    bar_kernel<<<1, nw * 32>>>(d_out, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    double a = (double)h[0] / iters;
    bar_only_kernel<<<1, nw * 32>>>(d_out, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("barrier nw=%2d: STS+BAR+LDS round %6.1f cycles, BAR only %6.1f cycles\n", nw, a, (double)h[0] / iters);
  }
  run_tput<0>("FFMA 3 distinct src", d_out, d_sink, 1);
  run_tput<1>("FFMA a*a+c", d_out, d_sink, 1);
  run_tput<2>("FADD", d_out, d_sink, 1);
  run_tput<6>("FMNMX", d_out, d_sink, 1);
  run_tput<3>("pair: 3FADD+FMUL+2FFMA", d_out, d_sink, 6);
  run_tput<5>("pair + FMNMX", d_out, d_sink, 7);
  run_tput<4>("pair + FSETP/FSEL/SEL", d_out, d_sink, 9);
  run_tput<7>("pair + 0.5 FMNMX3", d_out, d_sink, 6);
  run_tput<8>("REDUX.MIN independent (+IADD x2)", d_out, d_sink, 3);
  run_tput<9>("SHFL.bfly independent (+IADD x2)", d_out, d_sink, 3);
  run_tput<10>("ballot independent (+LOP,IADD x2)", d_out, d_sink, 4);
  return 0;
}
