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

#include "scatter.cuh"

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
#pragma unroll
      for (int u = 7; u >= 0; --u) {  // independent (clamped) loads, lowest match wins
        const float* p = rp + static_cast<size_t>(min(bg[r] + u, nr - 1)) * 3;
        const float d = dist_yxz(__ldg(p) - qx[r], __ldg(p + 1) - qy[r], __ldg(p + 2) - qz[r]);
        if (d == best[r] && bg[r] + u < nr) bi = bg[r] + u;
      }
      dout[j] = best[r];
      iout[j] = bi;
    }
  }
}

// Predicated RED.MIN.U64 of the key (bits << 32 | code), issued only by lanes whose value equals the
// warp minimum.  The key is assembled outside the predicate so that ptxas keeps a single predicated
// REDG instead of a divergent branch (BSSY / BRA / BSYNC) around it.
__device__ __forceinline__ void red_min_key_if_equal(unsigned long long* addr, unsigned mine,
                                                     unsigned warp_min, unsigned code) {
  const unsigned long long key = (static_cast<unsigned long long>(warp_min) << 32) | code;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, %1, %2;\n\t"
      "@p red.global.min.u64 [%0], %3;\n\t"
      "}" ::"l"(addr),
      "r"(mine), "r"(warp_min), "l"(key)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Single-pass forward: every pairwise distance is computed ONCE and serves both directions
// (the matrix is bitwise symmetric: (-a)^2 == a^2), halving the FP32 work of the reference's two
// directed launches.  Rows (cloud A) live in registers, R per thread, lane-major so that a lower
// lane owns lower row indices; columns (cloud B) stream through shared memory in TMA-staged tiles,
// the CTA's WC warps taking interleaved groups of 8 columns.
//   row side : as chamfer_fwd_kernel -- FMNMX3 folds pairs of columns into a per-group min, one
//              compare/select per 8 columns; the WC partial results per row are combined through
//              shared memory and the arg-min inside the winning group is recomputed once per row.
//   col side : per column, an FMNMX3 tree over the thread's R rows, REDUX.MIN over the warp on the
//              (non-negative) float bits, and the lane(s) holding the minimum issue one
//              RED.MIN.U64 of (distance bits << 32 | row-block id) into a (B, M8) workspace.
//              64-bit min == smallest distance, then lowest row block.  chamfer_cols_finalize_kernel
//              then turns the keys into dist/idx by recomputing the R distances of the winning
//              row block -- first row equal to the minimum => lowest index on ties.
// The workspace is preset to 0xFF bytes (keys = +max) by a memset node in front of the kernel.
template <int R, int WC>
__global__ void __launch_bounds__(WC * 32)
    chamfer_fwd_both_kernel(const float* __restrict__ xyzA, const float* __restrict__ xyzB, int NA,
                            int NB, int NB8, float* __restrict__ distA, int32_t* __restrict__ idxA,
                            float* __restrict__ distB, int32_t* __restrict__ idxB,
                            unsigned long long* __restrict__ colbest) {
  constexpr int ROWS = 32 * R;
  constexpr int THREADS = WC * 32;
  static_assert(WC * ROWS * 8 <= kChTile * 12, "row-combine scratch must fit in the tile buffer");
  __shared__ __align__(16) float s_ref[kChTile * 3];
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x, lane = t & 31;
  // warp index through a shuffle: lets the compiler treat it (and the column loop) as warp-uniform,
  // so the REDUX in the loop needs no divergence guard (BSSY / BRA.DIV / BSYNC)
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int b = blockIdx.y;
  const float* ap = xyzA + static_cast<size_t>(b) * NA * 3;
  const float* bp = xyzB + static_cast<size_t>(b) * NB * 3;
  unsigned long long* cb = colbest + static_cast<size_t>(b) * NB8;
  const int row0 = blockIdx.x * ROWS + lane * R;
  const unsigned code = static_cast<unsigned>(blockIdx.x * 32 + lane);  // row block id = row0 / R

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
    const int j = row0 + r;
    const bool ok = j < NA;  // rows past the end are NaN: they never win a min
    qx[r] = ok ? __ldg(ap + 3 * j) : __int_as_float(0x7fc00000);
    qy[r] = ok ? __ldg(ap + 3 * j + 1) : __int_as_float(0x7fc00000);
    qz[r] = ok ? __ldg(ap + 3 * j + 2) : __int_as_float(0x7fc00000);
    best[r] = __int_as_float(0x7f800000);
    bg[r] = 0;
  }

  for (int base = 0; base < NB; base += kChTile) {
    const int tile = min(kChTile, NB - base);
    const int tile8 = (tile + 7) & ~7;
    if (base > 0) __syncthreads();
    for (int i = tile * 3 + t; i < tile8 * 3; i += THREADS) s_ref[i] = __int_as_float(0x7fc00000);
    stage_points(s_ref, bp + static_cast<size_t>(base) * 3, tile, &s_bar, parity);

    const int ngroups = tile8 >> 3;
    for (int g = warp; g < ngroups; g += WC) {
      const float4* s4 = reinterpret_cast<const float4*>(s_ref) + g * 6;
      unsigned long long* cbk = cb + base + g * 8;
      float m[R];
#pragma unroll
      for (int h = 0; h < 4; ++h) {  // four column pairs per group of 8
        // refs 2h and 2h+1 of the group: 6 consecutive floats starting at float 6h
        const float4 v0 = s4[(6 * h) >> 2];
        const float4 v1 = s4[((6 * h) >> 2) + 1];
        float ax, ay, az, bx, by, bz;
        if ((h & 1) == 0) { ax = v0.x; ay = v0.y; az = v0.z; bx = v0.w; by = v1.x; bz = v1.y; }
        else              { ax = v0.z; ay = v0.w; az = v1.x; bx = v1.y; by = v1.z; bz = v1.w; }
        float ca = 0.f, cbm = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float da = dist_yxz(ax - qx[r], ay - qy[r], az - qz[r]);
          const float db = dist_yxz(bx - qx[r], by - qy[r], bz - qz[r]);
          m[r] = h == 0 ? fminf(da, db) : fminf(fminf(m[r], da), db);
          ca = r == 0 ? da : fminf(ca, da);
          cbm = r == 0 ? db : fminf(cbm, db);
        }
        const unsigned ua = __float_as_uint(ca), ub = __float_as_uint(cbm);
        const unsigned wa = redux_min_u32(ua), wb = redux_min_u32(ub);
        red_min_key_if_equal(cbk + 2 * h, ua, wa, code);
        red_min_key_if_equal(cbk + 2 * h + 1, ub, wb, code);
      }
      const int kk = base + g * 8;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (m[r] < best[r]) { best[r] = m[r]; bg[r] = kk; }
    }
  }

  // ---- row side: combine the WC column-chunk partials, resolve the arg-min inside the group ----
  __syncthreads();  // every warp is done reading the tile
  float* s_best = s_ref;
  int* s_bg = reinterpret_cast<int*>(s_ref + WC * ROWS);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s_best[warp * ROWS + lane * R + r] = best[r];
    s_bg[warp * ROWS + lane * R + r] = bg[r];
  }
  __syncthreads();
  for (int rho = t; rho < ROWS; rho += THREADS) {
    const int j = blockIdx.x * ROWS + rho;
    if (j >= NA) continue;
    float bb = __int_as_float(0x7f800000);
    int gg = 0;
#pragma unroll
    for (int w = 0; w < WC; ++w) {
      const float v = s_best[w * ROWS + rho];
      const int g = s_bg[w * ROWS + rho];
      if (v < bb || (v == bb && g < gg)) { bb = v; gg = g; }
    }
    const float px = __ldg(ap + 3 * j), py = __ldg(ap + 3 * j + 1), pz = __ldg(ap + 3 * j + 2);
    int bi = gg;
#pragma unroll
    for (int u = 7; u >= 0; --u) {  // 8 independent (clamped) loads in flight, lowest match wins
      const int kcol = min(gg + u, NB - 1);
      const float* p = bp + static_cast<size_t>(kcol) * 3;
      const float d = dist_yxz(__ldg(p) - px, __ldg(p + 1) - py, __ldg(p + 2) - pz);
      if (d == bb && gg + u < NB) bi = gg + u;
    }
    distA[static_cast<size_t>(b) * NA + j] = bb;
    idxA[static_cast<size_t>(b) * NA + j] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// Single-pass forward, packed: the same algorithm as chamfer_fwd_both_kernel with the two changes the
// B200 measurements asked for (scripts/microbench2.cu, profiles/r01b_chamfer_both_R8W8.txt):
//   * PACKED fp32x2 distances: a thread's R rows sit in R/2 register pairs; one FADD2/FMUL2/FFMA2
//     evaluates a column against two rows -- 6 issue slots per 2 pairs instead of 12, bit-identical
//     results -- which frees the issue port for the FMNMX bookkeeping (the FMA pipe needs 2 cycles
//     per packed instruction): the loop is FMA-pipe bound at 6 cycles per pair instead of
//     issue bound at ~8.25.
//   * COLUMN CHUNKS (blockIdx.z): a work item is (cloud, 32*R rows, one chunk of columns), so the
//     grid can be cut to a whole number of SM-waves (the R8W8 profile ran 1.15 waves).  With more
//     than one chunk the row side also merges through RED.MIN.U64 keys (distance bits << 32 | column
//     group) and chamfer_finalize_kernel resolves both sides; with one chunk rows are written directly.
template <int R, int WC, int MINB = 0>
__global__ void __launch_bounds__(WC * 32, MINB)
    chamfer_fwd_packed_kernel(const float* __restrict__ xyzA, const float* __restrict__ xyzB, int NA,
                              int NB, int NA8, int NB8, int chunk, float* __restrict__ distA,
                              int32_t* __restrict__ idxA, unsigned long long* __restrict__ rowbest,
                              unsigned long long* __restrict__ colbest) {
  static_assert(R % 2 == 0, "rows are processed in packed pairs");
  constexpr int ROWS = 32 * R;
  constexpr int THREADS = WC * 32;
  constexpr int RP = R / 2;
  static_assert(WC * ROWS * 8 <= kChTile * 12, "row-combine scratch must fit in the tile buffer");
  __shared__ __align__(16) float s_ref[kChTile * 3];
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);  // warp-uniform for the compiler
  const int b = blockIdx.y;
  const int c0 = blockIdx.z * chunk;
  const int c1 = min(NB, c0 + chunk);
  const float* ap = xyzA + static_cast<size_t>(b) * NA * 3;
  const float* bp = xyzB + static_cast<size_t>(b) * NB * 3;
  unsigned long long* cb = colbest + static_cast<size_t>(b) * NB8;
  const int row0 = blockIdx.x * ROWS + lane * R;
  const unsigned code = static_cast<unsigned>(blockIdx.x * 32 + lane);  // row block id = row0 / R

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;

  f32x2 QX[RP], QY[RP], QZ[RP];
  float best[R];
  int bg[R];
#pragma unroll
  for (int rp = 0; rp < RP; ++rp) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = row0 + 2 * rp + h;
      const bool ok = j < NA;  // rows past the end are NaN: they never win a min
#pragma unroll
      for (int a = 0; a < 3; ++a) c[h][a] = ok ? __ldg(ap + 3 * j + a) : __int_as_float(0x7fc00000);
    }
    QX[rp] = pack2(c[0][0], c[1][0]);
    QY[rp] = pack2(c[0][1], c[1][1]);
    QZ[rp] = pack2(c[0][2], c[1][2]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    best[r] = __int_as_float(0x7f800000);
    bg[r] = 0;
  }

  for (int base = c0; base < c1; base += kChTile) {
    const int tile = min(kChTile, c1 - base);
    const int tile8 = (tile + 7) & ~7;
    if (base > c0) __syncthreads();
    // pad the last group of eight with NaN: (q - NaN)^2 = NaN, and fminf() drops NaN operands
    for (int i = tile * 3 + t; i < tile8 * 3; i += THREADS) s_ref[i] = __int_as_float(0x7fc00000);
    stage_points(s_ref, bp + static_cast<size_t>(base) * 3, tile, &s_bar, parity);

    const int ngroups = tile8 >> 3;
    for (int g = warp; g < ngroups; g += WC) {
      const float4* s4 = reinterpret_cast<const float4*>(s_ref) + g * 6;
      unsigned long long* cbk = cb + base + g * 8;
      float m[R];
#pragma unroll
      for (int h = 0; h < 4; ++h) {  // four column pairs per group of 8
        const float4 v0 = s4[(6 * h) >> 2];
        const float4 v1 = s4[((6 * h) >> 2) + 1];
        float ax, ay, az, bx, by, bz;
        if ((h & 1) == 0) { ax = v0.x; ay = v0.y; az = v0.z; bx = v0.w; by = v1.x; bz = v1.y; }
        else              { ax = v0.z; ay = v0.w; az = v1.x; bx = v1.y; by = v1.z; bz = v1.w; }
        // (q - ref): the sign is squared away, and packed-minus-broadcast is one FADD2
        const f32x2 AX = pack2(ax, ax), AY = pack2(ay, ay), AZ = pack2(az, az);
        const f32x2 BX = pack2(bx, bx), BY = pack2(by, by), BZ = pack2(bz, bz);
        float ca = 0.f, cbm = 0.f;
#pragma unroll
        for (int rp = 0; rp < RP; ++rp) {
          float a0, a1, b0, b1;
          unpack2(dist2_yxz(sub2(QX[rp], AX), sub2(QY[rp], AY), sub2(QZ[rp], AZ)), a0, a1);
          unpack2(dist2_yxz(sub2(QX[rp], BX), sub2(QY[rp], BY), sub2(QZ[rp], BZ)), b0, b1);
          m[2 * rp] = h == 0 ? fminf(a0, b0) : fminf(fminf(m[2 * rp], a0), b0);
          m[2 * rp + 1] = h == 0 ? fminf(a1, b1) : fminf(fminf(m[2 * rp + 1], a1), b1);
          ca = rp == 0 ? fminf(a0, a1) : fminf(fminf(ca, a0), a1);
          cbm = rp == 0 ? fminf(b0, b1) : fminf(fminf(cbm, b0), b1);
        }
        const unsigned ua = __float_as_uint(ca), ub = __float_as_uint(cbm);
        const unsigned wa = redux_min_u32(ua), wb = redux_min_u32(ub);
        red_min_key_if_equal(cbk + 2 * h, ua, wa, code);
        red_min_key_if_equal(cbk + 2 * h + 1, ub, wb, code);
      }
      const int kk = base + g * 8;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (m[r] < best[r]) { best[r] = m[r]; bg[r] = kk; }
    }
  }

  // ---- row side: combine the WC partials per row, then either resolve + write, or merge by key ----
  __syncthreads();  // every warp is done reading the tile
  float* s_best = s_ref;
  int* s_bg = reinterpret_cast<int*>(s_ref + WC * ROWS);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s_best[warp * ROWS + lane * R + r] = best[r];
    s_bg[warp * ROWS + lane * R + r] = bg[r];
  }
  __syncthreads();
  for (int rho = t; rho < ROWS; rho += THREADS) {
    const int j = blockIdx.x * ROWS + rho;
    if (j >= NA) continue;
    float bb = __int_as_float(0x7f800000);
    int gg = 0;
#pragma unroll
    for (int w = 0; w < WC; ++w) {
      const float v = s_best[w * ROWS + rho];
      const int g = s_bg[w * ROWS + rho];
      if (v < bb || (v == bb && g < gg)) { bb = v; gg = g; }
    }
    if (rowbest != nullptr) {  // several column chunks: smallest distance, then lowest column group
      const unsigned long long key =
          (static_cast<unsigned long long>(__float_as_uint(bb)) << 32) | static_cast<unsigned>(gg);
      atomicMin(rowbest + static_cast<size_t>(b) * NA8 + j, key);
      continue;
    }
    const float px = __ldg(ap + 3 * j), py = __ldg(ap + 3 * j + 1), pz = __ldg(ap + 3 * j + 2);
    int bi = gg;
#pragma unroll
    for (int u = 7; u >= 0; --u) {  // 8 independent (clamped) loads in flight, lowest match wins
      const int kcol = min(gg + u, NB - 1);
      const float* p = bp + static_cast<size_t>(kcol) * 3;
      const float d = dist_yxz(__ldg(p) - px, __ldg(p + 1) - py, __ldg(p + 2) - pz);
      if (d == bb && gg + u < NB) bi = gg + u;
    }
    distA[static_cast<size_t>(b) * NA + j] = bb;
    idxA[static_cast<size_t>(b) * NA + j] = bi;
  }
}

// Second half of the key merge: turn each packed key (distance bits << 32 | code) of cloud `self`
// into (distance, lowest matching index in `other`) by recomputing the `cnt` (<= 8) distances of the
// winning span other[code * mult .. + cnt).  One thread per point, all loads independent.
__global__ void __launch_bounds__(256)
    chamfer_finalize_kernel(const float* __restrict__ self, const float* __restrict__ other, int n_self,
                            int n_other, int n_self8, const unsigned long long* __restrict__ keys,
                            int mult, int cnt, float* __restrict__ dist_out, int32_t* __restrict__ idx_out) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_self) return;
  const float* op = other + static_cast<size_t>(b) * n_other * 3;
  const float* sp = self + (static_cast<size_t>(b) * n_self + k) * 3;
  const unsigned long long key = keys[static_cast<size_t>(b) * n_self8 + k];
  const unsigned bits = static_cast<unsigned>(key >> 32);
  const int j0 = static_cast<int>(static_cast<unsigned>(key)) * mult;
  const float rx = __ldg(sp), ry = __ldg(sp + 1), rz = __ldg(sp + 2);
  int bi = j0;
#pragma unroll
  for (int u = 7; u >= 0; --u) {
    const int j = min(j0 + u, n_other - 1);
    const float* p = op + static_cast<size_t>(j) * 3;
    const float d = dist_yxz(rx - __ldg(p), ry - __ldg(p + 1), rz - __ldg(p + 2));
    if (u < cnt && __float_as_uint(d) == bits && j0 + u < n_other) bi = j0 + u;
  }
  dist_out[static_cast<size_t>(b) * n_self + k] = __uint_as_float(bits);
  idx_out[static_cast<size_t>(b) * n_self + k] = bi;
}

// ---------------------------------------------------------------------------------------------
// Packed path, second (and last) launch: ONE kernel resolves the merged keys of both sides and
// produces the call's partial sums, replacing {cols finalize, rows finalize, sums} = 3 launches.
//   blockIdx.z == 0 : column side (cloud B): key -> (distance, lowest matching row of A)
//   blockIdx.z == 1 : row side (cloud A): from keys when the columns were cut into chunks, else the
//                     main kernel already wrote distA/idxA and these CTAs only read distA for the sums
//                     (the z == 1 slice is not launched at all when neither is needed).
// Sums { sum d, sum sqrt d } per side: fixed-shape block reduction -> partials[cta]; the CTA that
// draws the last ticket adds the partials in index order.  Fixed partition + fixed order: the value
// is identical run to run whichever CTA happens to be last.  The ticket word sits behind the keys and
// is preset to 0xFFFFFFFF by the same memset node, so the last of T CTAs draws ticket T - 2 (mod 2^32).
// ---- fused all-reduce of the 4 sums over NVLink peer memory (batch sharded across GPUs) ----------------
// Every rank owns an exchange buffer slots[2][world][8 floats] that all ranks have mapped (CUDA IPC).  The CTA
// that finishes a rank's local sums stores them, tagged with the call's sequence number, into slot
// [seq & 1][rank] of EVERY rank's buffer (plain NVLink stores, release at system scope), then waits until its own
// buffer holds the tag in all `world` slots and adds the world's contributions in RANK ORDER -- every rank
// computes bit-identical global sums, deterministically, with no NCCL launch and no host involvement.
// Two parities make slot reuse safe: rank r can only write call s+2 after it has finished call s+1, which needed
// every peer's call-s+1 contribution, which a peer sends only after it has read call s.
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerXchg {
  float* slots[UPP_MAX_PEERS];  // slots[r] = rank r's exchange buffer (device pointer valid in THIS process)
  int rank, world;
  unsigned* seq;                // this rank's call counter (device memory, zero before the first call)
  int defer;                    // 1: only SEND here; peer_finish_kernel (a later launch) waits and adds
  unsigned* status;             // nullable: receives the sequence number of a wait that timed out (host-visible memory)
  long long timeout;            // SM clocks a wait may last (0: kPeerTimeoutDefault)
};

// A peer that is late is NORMAL under data-parallel training (rank-0 checkpointing or evaluation, a data-loader stall):
// the wait is long (2^38 clocks, ~2.3 minutes at 1.965 GHz -- the order of NCCL's own watchdog, not of a kernel), and a
// wait that does run out is REPORTED: the sums are poisoned with NaN and, when the caller gave a status word, the
// sequence number of the failed call is stored there with system scope so that the host can raise (parallel.PeerExchange
// .check()); nothing is skipped silently.
constexpr long long kPeerTimeoutDefault = 1LL << 38;

__device__ __forceinline__ bool peer_wait_tag(const float* src, unsigned seq, const PeerXchg& px) {
  const long long limit = px.timeout > 0 ? px.timeout : kPeerTimeoutDefault;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (ld_acquire_sys_u32(reinterpret_cast<const unsigned*>(src + 4)) != seq) {
    if ((++spins & 1023u) == 0u && clock64() - t0 > limit) {
      if (px.status != nullptr) {
        *reinterpret_cast<volatile unsigned*>(px.status) = seq;
        __threadfence_system();
      }
      return false;
    }
  }
  return true;
}

// Called by ONE CTA (256 threads) after `local` (4 floats, thread 0's registers) is final.  Returns the global
// sums in thread 0.
__device__ __forceinline__ void peer_allreduce4(const PeerXchg& px, float (&v)[4], float (*s_x)[4]) {
  __shared__ unsigned s_seq;
  if (threadIdx.x == 0) {
    s_seq = *px.seq + 1u;
    *px.seq = s_seq;
#pragma unroll
    for (int q = 0; q < 4; ++q) s_x[0][q] = v[q];
  }
  __syncthreads();
  const unsigned seq = s_seq;
  const int t = threadIdx.x;
  if (t < px.world) {
    // send: my contribution into slot [parity][rank] of peer t
    float* dst = px.slots[t] + ((seq & 1u) * px.world + px.rank) * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(s_x[0][0], s_x[0][1], s_x[0][2], s_x[0][3]);
    st_release_sys_u32(reinterpret_cast<unsigned*>(dst + 4), seq);
  }
  __syncthreads();
  if (px.defer) return;  // the wait happens in peer_finish_kernel, late in the step: a kernel that spins early would
                         // hold up whatever the hardware queues behind it (measured: +8..24 us per step at 2 GPUs)
  if (t < px.world) {
    // receive: rank t's contribution from my own buffer
    const float* src = px.slots[px.rank] + ((seq & 1u) * px.world + t) * 8;
    const bool ok = peer_wait_tag(src, seq, px);  // long, reported wait (see kPeerTimeoutDefault)
    const volatile float* vs = src;  // after the acquire: the payload written before the tag
    const float nan = __int_as_float(0x7fc00000);
    s_x[1 + t][0] = ok ? vs[0] : nan; s_x[1 + t][1] = ok ? vs[1] : nan;
    s_x[1 + t][2] = ok ? vs[2] : nan; s_x[1 + t][3] = ok ? vs[3] : nan;
  }
  __syncthreads();
  if (t == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float tot = 0.f;
      for (int r = 0; r < px.world; ++r) tot += s_x[1 + r][q];
      v[q] = tot;
    }
  }
}

// Second half of a deferred exchange: wait for the world's contributions of the CURRENT sequence number in this
// rank's own buffer, add them in rank order.  One warp; peers had the rest of the step to deliver.
__global__ void __launch_bounds__(32) peer_finish_kernel(const PeerXchg px, float* __restrict__ global_sums) {
  __shared__ float s_v[UPP_MAX_PEERS][4];
  const int t = threadIdx.x;
  const unsigned seq = *px.seq;
  for (int r = t; r < px.world; r += 32) {
    const float* src = px.slots[px.rank] + ((seq & 1u) * px.world + r) * 8;
    const bool ok = peer_wait_tag(src, seq, px);
    const volatile float* vs = src;
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int q = 0; q < 4; ++q) s_v[r][q] = ok ? vs[q] : nan;
  }
  __syncwarp();
  if (t < 4) {
    float tot = 0.f;
    for (int r = 0; r < px.world; ++r) tot += s_v[r][t];
    global_sums[t] = tot;
  }
}

// The exchange on its own: all-reduce (SUM, rank order) of 4 floats that are already in device memory -- an EMPTY shard's
// contribution to a sharded Chamfer call (local == nullptr: zeros; every rank must take part in every exchange), or the
// gradient statistics of chamfer_bwd_stats.  One CTA.
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerXchg px, const float* __restrict__ local,
                                                             float* __restrict__ global_sums) {
  __shared__ float s_x[1 + UPP_MAX_PEERS][4];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (threadIdx.x == 0 && local != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = local[q];
  }
  peer_allreduce4(px, v, s_x);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) global_sums[q] = v[q];
  }
}

__device__ __forceinline__ float2 block_sum2_256(float a, float b, float2* s_w) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_w[warp] = make_float2(a, b);
  __syncthreads();
  float2 r = make_float2(0.f, 0.f);
  if (warp == 0) {
    const float2 v = lane < 8 ? s_w[lane] : make_float2(0.f, 0.f);
    r.x = warp_sum(v.x);
    r.y = warp_sum(v.y);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__global__ void __launch_bounds__(256)
    chamfer_finalize2_kernel(const float* __restrict__ xyzA, const float* __restrict__ xyzB, int NA, int NB,
                             int NA8, int NB8, const unsigned long long* __restrict__ colkeys,
                             const unsigned long long* __restrict__ rowkeys, int R,
                             float* __restrict__ distA, int32_t* __restrict__ idxA,
                             float* __restrict__ distB, int32_t* __restrict__ idxB,
                             float2* __restrict__ partials, unsigned* __restrict__ ticket,
                             float* __restrict__ sums, int swapped, const PeerXchg px) {
  __shared__ float2 s_w[8];
  __shared__ bool s_last;
  __shared__ float s_x[1 + UPP_MAX_PEERS][4];
  const int side = blockIdx.z;  // 0: B (columns), 1: A (rows)
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_self = side == 0 ? NB : NA, n_other = side == 0 ? NA : NB;
  float d = 0.f;
  if (k < n_self) {
    if (side == 1 && rowkeys == nullptr) {
      d = distA[static_cast<size_t>(b) * NA + k];  // written by the main kernel (previous launch)
    } else {
      const float* self = side == 0 ? xyzB : xyzA;
      const float* op = (side == 0 ? xyzA : xyzB) + static_cast<size_t>(b) * n_other * 3;
      const float* sp = self + (static_cast<size_t>(b) * n_self + k) * 3;
      const unsigned long long key = side == 0 ? colkeys[static_cast<size_t>(b) * NB8 + k]
                                               : rowkeys[static_cast<size_t>(b) * NA8 + k];
      const int mult = side == 0 ? R : 1, cnt = side == 0 ? R : 8;
      const unsigned bits = static_cast<unsigned>(key >> 32);
      const int j0 = static_cast<int>(static_cast<unsigned>(key)) * mult;
      const float rx = __ldg(sp), ry = __ldg(sp + 1), rz = __ldg(sp + 2);
      int bi = j0;
#pragma unroll
      for (int u = 15; u >= 0; --u) {  // independent (clamped) loads, lowest match wins
        if (u < cnt) {
          const int j = min(j0 + u, n_other - 1);
          const float* p = op + static_cast<size_t>(j) * 3;
          const float dd = dist_yxz(rx - __ldg(p), ry - __ldg(p + 1), rz - __ldg(p + 2));
          if (__float_as_uint(dd) == bits && j0 + u < n_other) bi = j0 + u;
        }
      }
      d = __uint_as_float(bits);
      (side == 0 ? distB : distA)[static_cast<size_t>(b) * n_self + k] = d;
      (side == 0 ? idxB : idxA)[static_cast<size_t>(b) * n_self + k] = bi;
    }
  }
  if (partials == nullptr) return;
  const unsigned per_side = gridDim.x * gridDim.y;
  const unsigned cta = side * per_side + blockIdx.y * gridDim.x + blockIdx.x;
  const float2 part = block_sum2_256(d, k < n_self ? __fsqrt_rn(d) : 0.f, s_w);
  if (threadIdx.x == 0) {
    partials[cta] = part;
    __threadfence();
    const unsigned total = per_side * gridDim.z;
    s_last = (atomicAdd(ticket, 1u) + 2u == total);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: side 0 = cloud B, side 1 = cloud A; output order { d1, d2, sqrt d1, sqrt d2 }
  float loc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int sd = 0; sd < 2; ++sd) {
    float a0 = 0.f, a1 = 0.f;
    if (sd < static_cast<int>(gridDim.z)) {
      const volatile float2* pp = partials + sd * per_side;
      for (unsigned i = threadIdx.x; i < per_side; i += 256) {
        a0 += pp[i].x;
        a1 += pp[i].y;
      }
    }
    const float2 tot = block_sum2_256(a0, a1, s_w);
    const int first = (sd == 1) != (swapped != 0);  // is this side the caller's xyz1?
    loc[first ? 0 : 1] = tot.x;                     // (valid in thread 0)
    loc[first ? 2 : 3] = tot.y;
  }
  if (px.world > 1) peer_allreduce4(px, loc, s_x);  // sums of the WHOLE sharded batch, same bits on every rank
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sums[q] = loc[q];
  }
}

// Column side, second half: turn each packed key into (distance, lowest matching row index) by
// recomputing the R distances of the winning row block.  One thread per column, all loads independent.
template <int R>
__global__ void __launch_bounds__(256)
    chamfer_cols_finalize_kernel(const float* __restrict__ xyzA, const float* __restrict__ xyzB, int NA,
                                 int NB, int NB8, const unsigned long long* __restrict__ colbest,
                                 float* __restrict__ distB, int32_t* __restrict__ idxB) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= NB) return;
  const float* ap = xyzA + static_cast<size_t>(b) * NA * 3;
  const float* bp = xyzB + (static_cast<size_t>(b) * NB + k) * 3;
  const unsigned long long key = colbest[static_cast<size_t>(b) * NB8 + k];
  const unsigned bits = static_cast<unsigned>(key >> 32);
  const int j0 = static_cast<int>(static_cast<unsigned>(key)) * R;
  const float rx = __ldg(bp), ry = __ldg(bp + 1), rz = __ldg(bp + 2);
  int bi = j0;
#pragma unroll
  for (int u = R - 1; u >= 0; --u) {
    const int j = min(j0 + u, NA - 1);
    const float* p = ap + static_cast<size_t>(j) * 3;
    const float d = dist_yxz(rx - __ldg(p), ry - __ldg(p + 1), rz - __ldg(p + 2));
    if (__float_as_uint(d) == bits && j0 + u < NA) bi = j0 + u;
  }
  distB[static_cast<size_t>(b) * NB + k] = __uint_as_float(bits);
  idxB[static_cast<size_t>(b) * NB + k] = bi;
}

// ---------------------------------------------------------------------------------------------
// Single-pass forward, ONE kernel (round 2): chamfer_fwd_packed_kernel's distance loop, with the three things around
// it folded in -- the key-workspace memset, the finalize launch and the atomics:
//   * NO ATOMICS, NO PRESET.  A column's minimum over a CTA's 32*R rows is produced by exactly one warp (the CTA's
//     warps take disjoint column groups), so the CTA owns the slot colpart[cloud][row block][column] and writes
//     {ballot of the lanes holding the minimum, distance bits} with one plain 8-byte store.  (The RED.MIN.U64 it
//     replaces was compiled to BSSY / BRA / MOV / REDG / BSYNC: 8 instructions per column and warp against 5, an L2
//     atomic per column and row block, and a 0xFF memset of every key in front of the kernel.)  With several column
//     chunks the row side does the same into rowpart[cloud][chunk][row].
//   * THE FINALIZE RUNS INSIDE.  Tickets count the CTAs that have contributed to a (cloud, column chunk) and to a
//     (cloud, row block); the CTA that draws the last ticket resolves those columns / rows itself: minimum over the
//     partials (strict '<' in ascending row-block / chunk order = lowest index on ties), arg-min inside the winning span
//     by recomputing its <= 16 distances, dist/idx written once.  This work overlaps the distance loops of the CTAs
//     still running instead of waiting for the grid to drain behind a second launch.
//   * THE SUMS, and the all-reduce over NVLink peer memory when the batch is sharded, close the same kernel: every
//     finalizing CTA leaves {sum d, sum sqrt d} of its points in a fixed slot, and the CTA that draws the last global
//     ticket adds the slots in index order (deterministic whichever CTA that is) and runs peer_allreduce4.
//   * CG = 16: the row side keeps (best value, group) per 16 columns instead of 8 (one FSETP/FSEL/SEL per row and 16
//     columns), the arg-min scan in the finalize covers 16 candidates.
// Only the tickets (a few KB) are zeroed in front of the launch.
struct ChFused {
  const float* xyzA;
  const float* xyzB;
  int NA, NB, NA8, NB8, chunk, nchunks, RB;
  float* distA;
  int32_t* idxA;
  float* distB;
  int32_t* idxB;
  uint2* colpart;      // [B][RB][NB8]   {ballot, distance bits}
  uint2* rowpart;      // [B][nchunks][NA8] {column group base, distance bits}; unused when nchunks == 1
  unsigned* tickets;   // [B * (RB + nchunks)] + 1 global, zero on entry
  float2* partials;    // [B * (RB + nchunks)]  {sum d, sum sqrt d} of the points each finalizing CTA resolved
  float* sums;         // nullable: 4 floats { d1, d2, sqrt d1, sqrt d2 }
  int swapped;
};

template <int THREADS>
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* s_w) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_w[warp] = make_float2(a, b);
  __syncthreads();
  float2 r = make_float2(0.f, 0.f);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) { r.x += s_w[w].x; r.y += s_w[w].y; }
  }
  __syncthreads();
  return r;  // valid in thread 0
}

template <int R, int WC, int CG, int MINB = 0, bool FOLD = true>
__global__ void __launch_bounds__(WC * 32, MINB)
    chamfer_fwd_fused_kernel(const ChFused a, const PeerXchg px) {
  static_assert(R % 2 == 0 && R <= 16, "rows are processed in packed pairs; the finalize scans at most 16 rows");
  static_assert(CG == 8 || CG == 16, "column groups of 8 or 16");
  constexpr int ROWS = 32 * R;
  constexpr int THREADS = WC * 32;
  constexpr int RP = R / 2;
  static_assert(WC * ROWS * 8 <= kChTile * 12, "row-combine scratch must fit in the tile buffer");
  __shared__ __align__(16) float s_ref[kChTile * 3];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float2 s_w[WC];
  __shared__ int s_flags[3];
  __shared__ float s_x[1 + UPP_MAX_PEERS][4];
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);  // warp-uniform for the compiler
  const int b = blockIdx.y, rb = blockIdx.x, z = blockIdx.z;
  const int NA = a.NA, NB = a.NB;
  const int c0 = z * a.chunk;
  const int c1 = min(NB, c0 + a.chunk);
  const float* ap = a.xyzA + static_cast<size_t>(b) * NA * 3;
  const float* bp = a.xyzB + static_cast<size_t>(b) * NB * 3;
  uint2* cpart = a.colpart + (static_cast<size_t>(b) * a.RB + rb) * a.NB8;
  const int row0 = rb * ROWS + lane * R;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;

  f32x2 QX[RP], QY[RP], QZ[RP];
  float best[R];
  int bg[R];
#pragma unroll
  for (int rp = 0; rp < RP; ++rp) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = row0 + 2 * rp + h;
      const bool ok = j < NA;  // rows past the end are NaN: they never win a min
#pragma unroll
      for (int q = 0; q < 3; ++q) c[h][q] = ok ? __ldg(ap + 3 * j + q) : __int_as_float(0x7fc00000);
    }
    QX[rp] = pack2(c[0][0], c[1][0]);
    QY[rp] = pack2(c[0][1], c[1][1]);
    QZ[rp] = pack2(c[0][2], c[1][2]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    best[r] = __int_as_float(0x7f800000);
    bg[r] = 0;
  }

  for (int base = c0; base < c1; base += kChTile) {
    const int tile = min(kChTile, c1 - base);
    const int tileg = (tile + CG - 1) / CG * CG;
    if (base > c0) __syncthreads();
    // pad the last group with NaN: (q - NaN)^2 = NaN, and fminf() drops NaN operands
    for (int i = tile * 3 + t; i < tileg * 3; i += THREADS) s_ref[i] = __int_as_float(0x7fc00000);
    stage_points(s_ref, bp + static_cast<size_t>(base) * 3, tile, &s_bar, parity);

    const int ngroups = tileg / CG;
    for (int g = warp; g < ngroups; g += WC) {
      const float4* s4 = reinterpret_cast<const float4*>(s_ref) + g * (CG * 3 / 4);
      uint2* cbk = cpart + base + g * CG;
      float m[R];
#pragma unroll
      for (int h = 0; h < CG / 2; ++h) {  // column pairs of the group
        const float4 v0 = s4[(6 * h) >> 2];
        const float4 v1 = s4[((6 * h) >> 2) + 1];
        float ax, ay, az, bx, by, bz;
        if ((h & 1) == 0) { ax = v0.x; ay = v0.y; az = v0.z; bx = v0.w; by = v1.x; bz = v1.y; }
        else              { ax = v0.z; ay = v0.w; az = v1.x; bx = v1.y; by = v1.z; bz = v1.w; }
        const f32x2 AX = pack2(ax, ax), AY = pack2(ay, ay), AZ = pack2(az, az);
        const f32x2 BX = pack2(bx, bx), BY = pack2(by, by), BZ = pack2(bz, bz);
        float ca = 0.f, cbm = 0.f;
#pragma unroll
        for (int rp = 0; rp < RP; ++rp) {
          float a0, a1, b0, b1;
          unpack2(dist2_yxz(sub2(QX[rp], AX), sub2(QY[rp], AY), sub2(QZ[rp], AZ)), a0, a1);
          unpack2(dist2_yxz(sub2(QX[rp], BX), sub2(QY[rp], BY), sub2(QZ[rp], BZ)), b0, b1);
          m[2 * rp] = h == 0 ? fminf(a0, b0) : fminf(fminf(m[2 * rp], a0), b0);
          m[2 * rp + 1] = h == 0 ? fminf(a1, b1) : fminf(fminf(m[2 * rp + 1], a1), b1);
          ca = rp == 0 ? fminf(a0, a1) : fminf(fminf(ca, a0), a1);
          cbm = rp == 0 ? fminf(b0, b1) : fminf(fminf(cbm, b0), b1);
        }
        // column side: warp minimum of the (non-negative) float bits + which lanes hold it -- one plain store each
        const unsigned ua = __float_as_uint(ca), ub = __float_as_uint(cbm);
        const unsigned wa = redux_min_u32(ua), wb = redux_min_u32(ub);
        const unsigned ma = __ballot_sync(0xffffffffu, ua == wa), mb = __ballot_sync(0xffffffffu, ub == wb);
        cbk[2 * h] = make_uint2(ma, wa);
        cbk[2 * h + 1] = make_uint2(mb, wb);
      }
      const int kk = base + g * CG;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (m[r] < best[r]) { best[r] = m[r]; bg[r] = kk; }
    }
  }

  // ---- row side: combine the WC partials per row, then resolve (one chunk) or leave a partial (several) ----
  __syncthreads();  // every warp is done reading the tile
  float* s_best = s_ref;
  int* s_bg = reinterpret_cast<int*>(s_ref + WC * ROWS);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s_best[warp * ROWS + lane * R + r] = best[r];
    s_bg[warp * ROWS + lane * R + r] = bg[r];
  }
  __syncthreads();
  const bool direct = a.nchunks == 1;
  float sd = 0.f, ss = 0.f;  // this thread's share of { sum d, sum sqrt d } of the rows resolved here
  for (int rho = t; rho < ROWS; rho += THREADS) {
    const int j = rb * ROWS + rho;
    if (j >= NA) continue;
    float bb = __int_as_float(0x7f800000);
    int gg = 0;
#pragma unroll
    for (int w = 0; w < WC; ++w) {
      const float v = s_best[w * ROWS + rho];
      const int g = s_bg[w * ROWS + rho];
      if (v < bb || (v == bb && g < gg)) { bb = v; gg = g; }
    }
    if (!direct) {
      a.rowpart[(static_cast<size_t>(b) * a.nchunks + z) * a.NA8 + j] = make_uint2(static_cast<unsigned>(gg), __float_as_uint(bb));
      continue;
    }
    const float px0 = __ldg(ap + 3 * j), py0 = __ldg(ap + 3 * j + 1), pz0 = __ldg(ap + 3 * j + 2);
    int bi = gg;
#pragma unroll
    for (int u = CG - 1; u >= 0; --u) {  // independent (clamped) loads in flight, lowest match wins
      const int kcol = min(gg + u, NB - 1);
      const float* q = bp + static_cast<size_t>(kcol) * 3;
      const float d = dist_yxz(__ldg(q) - px0, __ldg(q + 1) - py0, __ldg(q + 2) - pz0);
      if (d == bb && gg + u < NB) bi = gg + u;
    }
    a.distA[static_cast<size_t>(b) * NA + j] = bb;
    a.idxA[static_cast<size_t>(b) * NA + j] = bi;
    sd += bb;
    ss += __fsqrt_rn(bb);
  }

  if constexpr (!FOLD) {
    // two-kernel form: chamfer_finalize3_kernel (next launch) resolves the partials and produces the sums; it counts its
    // own CTAs with tickets[0], zeroed here (this kernel is complete before that one starts)
    if (t == 0 && rb == 0 && b == 0 && z == 0) a.tickets[0] = 0u;
    return;
  }
  // ---- tickets: am I the last contributor to this row block (over chunks) / to this column chunk (over row blocks)? ----
  const unsigned slots = static_cast<unsigned>(a.RB + a.nchunks);
  unsigned* tk = a.tickets + static_cast<size_t>(b) * slots;
  __syncthreads();  // every thread's partial stores are ordered before thread 0's fence (cumulative) ...
  if (t == 0) {
    __threadfence();  // ... which makes them visible device-wide before the tickets are drawn: one fence per CTA, not 128
    const unsigned r0 = direct ? 0u : atomicAdd(tk + rb, 1u);   // two independent round trips, issued back to back
    const unsigned r1 = atomicAdd(tk + a.RB + z, 1u);
    s_flags[0] = direct ? 1 : (r0 + 1u == static_cast<unsigned>(a.nchunks));
    s_flags[1] = (r1 + 1u == static_cast<unsigned>(a.RB));
    __threadfence();
  }
  __syncthreads();
  const bool fin_rows = s_flags[0] != 0, fin_cols = s_flags[1] != 0;
  int finished = 0;

  if (fin_rows) {
    if (!direct) {  // rows of this block: minimum over the chunks' partials, arg-min inside the winning group
      for (int rho = t; rho < ROWS; rho += THREADS) {
        const int j = rb * ROWS + rho;
        if (j >= NA) continue;
        unsigned bits = 0xffffffffu, gg = 0;
        for (int zz = 0; zz < a.nchunks; ++zz) {
          const uint2 e = __ldcg(a.rowpart + (static_cast<size_t>(b) * a.nchunks + zz) * a.NA8 + j);
          if (e.y < bits) { bits = e.y; gg = e.x; }  // strict: the lower chunk (lower columns) keeps ties
        }
        const float px0 = __ldg(ap + 3 * j), py0 = __ldg(ap + 3 * j + 1), pz0 = __ldg(ap + 3 * j + 2);
        int bi = static_cast<int>(gg);
#pragma unroll
        for (int u = CG - 1; u >= 0; --u) {
          const int kcol = min(static_cast<int>(gg) + u, NB - 1);
          const float* q = bp + static_cast<size_t>(kcol) * 3;
          const float d = dist_yxz(__ldg(q) - px0, __ldg(q + 1) - py0, __ldg(q + 2) - pz0);
          if (__float_as_uint(d) == bits && static_cast<int>(gg) + u < NB) bi = static_cast<int>(gg) + u;
        }
        const float dd = __uint_as_float(bits);
        a.distA[static_cast<size_t>(b) * NA + j] = dd;
        a.idxA[static_cast<size_t>(b) * NA + j] = bi;
        sd += dd;
        ss += __fsqrt_rn(dd);
      }
    }
    if (a.sums != nullptr) {
      const float2 tot = block_sum2<THREADS>(sd, ss, s_w);
      if (t == 0) a.partials[static_cast<size_t>(b) * slots + rb] = tot;
    }
    ++finished;
  }
  if (fin_cols) {  // columns of this chunk: minimum over the row blocks' partials, arg-min inside the winning lane's R rows
    float cd = 0.f, csq = 0.f;
    for (int col = c0 + t; col < c1; col += THREADS) {
      unsigned bits = 0xffffffffu, mask = 0u;
      int wrb = 0;
      for (int rr = 0; rr < a.RB; ++rr) {
        const uint2 e = __ldcg(a.colpart + (static_cast<size_t>(b) * a.RB + rr) * a.NB8 + col);
        if (e.y < bits) { bits = e.y; mask = e.x; wrb = rr; }  // strict: the lower row block keeps ties
      }
      const int j0 = wrb * ROWS + (__ffs(mask) - 1) * R;  // lowest lane holding the minimum = lowest rows
      const float* q = bp + static_cast<size_t>(col) * 3;
      const float rx = __ldg(q), ry = __ldg(q + 1), rz = __ldg(q + 2);
      int bi = j0;
#pragma unroll
      for (int u = R - 1; u >= 0; --u) {
        const int j = min(j0 + u, NA - 1);
        const float* pr = ap + static_cast<size_t>(j) * 3;
        const float d = dist_yxz(rx - __ldg(pr), ry - __ldg(pr + 1), rz - __ldg(pr + 2));
        if (__float_as_uint(d) == bits && j0 + u < NA) bi = j0 + u;
      }
      const float dd = __uint_as_float(bits);
      a.distB[static_cast<size_t>(b) * NB + col] = dd;
      a.idxB[static_cast<size_t>(b) * NB + col] = bi;
      cd += dd;
      csq += __fsqrt_rn(dd);
    }
    if (a.sums != nullptr) {
      const float2 tot = block_sum2<THREADS>(cd, csq, s_w);
      if (t == 0) a.partials[static_cast<size_t>(b) * slots + a.RB + z] = tot;
    }
    ++finished;
  }
  if (a.sums == nullptr || finished == 0) return;

  // ---- global ticket: the CTA that completes the last of the B * (RB + nchunks) finalizations adds the partials ----
  const unsigned total = gridDim.y * slots;
  if (t == 0) {
    __threadfence();
    s_flags[2] = (atomicAdd(a.tickets + static_cast<size_t>(gridDim.y) * slots, static_cast<unsigned>(finished)) +
                  static_cast<unsigned>(finished) == total);
  }
  __syncthreads();
  if (s_flags[2] == 0) return;
  __threadfence();
  // fixed order: thread t adds slots t, t + THREADS, ... per side; then a fixed-shape block sum
  float2 side_sum[2];
#pragma unroll
  for (int sd2 = 0; sd2 < 2; ++sd2) {  // 0: cloud A (row slots), 1: cloud B (column slots)
    float a0 = 0.f, a1 = 0.f;
    const unsigned per = sd2 == 0 ? static_cast<unsigned>(a.RB) : static_cast<unsigned>(a.nchunks);
    const unsigned off = sd2 == 0 ? 0u : static_cast<unsigned>(a.RB);
    for (unsigned i = t; i < gridDim.y * per; i += THREADS) {
      const float2 v = __ldcg(a.partials + static_cast<size_t>(i / per) * slots + off + i % per);
      a0 += v.x;
      a1 += v.y;
    }
    side_sum[sd2] = block_sum2<THREADS>(a0, a1, s_w);  // (valid in thread 0)
  }
  // output order { d1, d2, sqrt d1, sqrt d2 } of the CALLER's xyz1 / xyz2: cloud A is xyz1 unless the clouds were swapped
  const float2 s1 = a.swapped ? side_sum[1] : side_sum[0], s2 = a.swapped ? side_sum[0] : side_sum[1];
  float loc[4] = {s1.x, s2.x, s1.y, s2.y};
  if (px.world > 1) peer_allreduce4(px, loc, s_x);  // sums of the WHOLE sharded batch, same bits on every rank
  if (t == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) a.sums[q] = loc[q];
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

// Backward: ONE launch, deterministic (ordered gather, scatter.cuh) -- no atomics, no memset.  blockIdx.z picks the
// side; a lane owns destinations j of that side's cloud `self`:
//   grad_self[j] = 2 g_self[j] (self_j - other[idx_self[j]])                      own term (chamfer.cu:186-191)
//                - sum_{k : idx_other[k] == j} 2 g_other[k] (other_k - self_j)    partner terms (chamfer.cu:192-198),
// the partner terms added in ascending k.  inf * 0 = NaN lands exactly where the reference's atomics put it (a row is
// poisoned by its own term or by a partner that points at it, never by anything else).
struct ChamferBwdOp {
  const float* self;        // (B,Ns,3) this side's cloud
  const float* other;       // (B,No,3)
  const int32_t* idx_self;  // (B,Ns) nearest `other` point of every `self` point
  const int32_t* idx_other; // (B,No) nearest `self` point of every `other` point: the scatter list
  const float* g_self;      // (B,Ns)
  const float* g_other;     // (B,No)
  float* grad;              // (B,Ns,3)
  int b, Ns, No;
  __device__ __forceinline__ int entries() const { return No; }
  __device__ __forceinline__ int dst(int e) const { return __ldg(idx_other + static_cast<size_t>(b) * No + e); }
  __device__ __forceinline__ void fetch(int e, int j, float (&v)[3]) const {
    const size_t k = static_cast<size_t>(b) * No + e;
    const float* o = other + k * 3;
    const float* a = self + (static_cast<size_t>(b) * Ns + j) * 3;  // j == idx_other[k]: the destination itself
    const float g = __fmul_rn(__ldg(g_other + k), 2.0f);
    v[0] = -__fmul_rn(g, __ldg(o) - __ldg(a));
    v[1] = -__fmul_rn(g, __ldg(o + 1) - __ldg(a + 1));
    v[2] = -__fmul_rn(g, __ldg(o + 2) - __ldg(a + 2));
  }
  __device__ __forceinline__ void init(int j, float (&acc)[3]) const {
    const size_t p = static_cast<size_t>(b) * Ns + j;
    const float* a = self + p * 3;
    const float* o = other + (static_cast<size_t>(b) * No + __ldg(idx_self + p)) * 3;
    const float g = __fmul_rn(__ldg(g_self + p), 2.0f);
    acc[0] = __fmul_rn(g, __ldg(a) - __ldg(o));
    acc[1] = __fmul_rn(g, __ldg(a + 1) - __ldg(o + 1));
    acc[2] = __fmul_rn(g, __ldg(a + 2) - __ldg(o + 2));
  }
  __device__ __forceinline__ void store(int j, const float (&acc)[3]) const {
    float* out = grad + (static_cast<size_t>(b) * Ns + j) * 3;
    out[0] = acc[0]; out[1] = acc[1]; out[2] = acc[2];
  }
};

template <int Q>
__global__ void __launch_bounds__(256)
    chamfer_bwd_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                       const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                       const float* __restrict__ g1, const float* __restrict__ g2, int N, int M, int blocks1,
                       float* __restrict__ gx1, float* __restrict__ gx2, float* __restrict__ sq_partials,
                       unsigned* __restrict__ ticket, float* __restrict__ sq_out, const PeerXchg px) {
  extern __shared__ int s_dst[];
  const int b = blockIdx.y;
  // blockIdx.x < blocks1: destinations are cloud 1's points; else cloud 2's (one grid, both sides)
  const bool first = static_cast<int>(blockIdx.x) < blocks1;
  const ChamferBwdOp op{first ? xyz1 : xyz2, first ? xyz2 : xyz1, first ? idx1 : idx2, first ? idx2 : idx1,
                        first ? g1 : g2,     first ? g2 : g1,     first ? gx1 : gx2,   b,
                        first ? N : M,       first ? M : N};
  float sq = ordered_scatter_cta<Q>(op, op.Ns, s_dst, first ? 0 : blocks1);
  if (sq_partials == nullptr) return;
  // ---- gradient statistics: sum ||grad||^2 per side.  Fixed-shape block reduction -> one partial per CTA; the CTA that
  //      draws the last ticket adds the partials in CTA order (deterministic whichever CTA that is), then -- batch sharded
  //      over GPUs -- all-reduces the two sums over NVLink peer memory like the forward's loss sums. ----
  __shared__ float s_w[8], s_w2[8];
  __shared__ bool s_last;
  __shared__ float s_x[1 + UPP_MAX_PEERS][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  sq = warp_sum(sq);
  if (lane == 0) s_w[warp] = sq;
  __syncthreads();
  const unsigned ncta = gridDim.x * gridDim.y;
  const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += s_w[w];
    sq_partials[cta] = tot;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) + 1u == ncta);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed partition (thread t takes partials t, t + blockDim, ...) and a fixed-shape block sum: deterministic
  float s1 = 0.f, s2 = 0.f;
  for (unsigned i = threadIdx.x; i < ncta; i += blockDim.x) {
    const float v = __ldcg(sq_partials + i);
    if ((i % gridDim.x) < static_cast<unsigned>(blocks1)) s1 += v; else s2 += v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  __syncthreads();
  if (lane == 0) { s_w[warp] = s1; s_w2[warp] = s2; }
  __syncthreads();
  float loc[4] = {0.f, 0.f, 0.f, 0.f};
  if (threadIdx.x == 0) {
    for (int w = 0; w < nwarps; ++w) { loc[0] += s_w[w]; loc[1] += s_w2[w]; }
    *ticket = 0u;  // leave the ticket as it was found
  }
  if (px.world > 1) peer_allreduce4(px, loc, s_x);
  if (threadIdx.x == 0) {
    sq_out[0] = loc[0];
    sq_out[1] = loc[1];
  }
}

// Second (and last) launch of the two-kernel form of the partial-slot path: one thread per point.
//   blockIdx.z == 0 : columns (cloud B): minimum over the RB row-block partials {ballot, bits} (strict '<' in ascending
//                     row-block order), arg-min among the R rows of the lowest lane holding it;
//   blockIdx.z == 1 : rows (cloud A): minimum over the chunk partials {group base, bits}, arg-min among the CG columns of
//                     the winning group -- or, with one chunk, distA as the main kernel wrote it (sums only).
// Sums and the fused all-reduce exactly as chamfer_finalize2_kernel (ticketed last CTA, fixed order, peer_allreduce4).
template <int R, int CG>
__global__ void __launch_bounds__(256)
    chamfer_finalize3_kernel(const ChFused a, const PeerXchg px) {
  __shared__ float2 s_w[8];
  __shared__ bool s_last;
  __shared__ float s_x[1 + UPP_MAX_PEERS][4];
  constexpr int ROWS = 32 * R;
  const int side = blockIdx.z;  // 0: B (columns), 1: A (rows)
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int NA = a.NA, NB = a.NB;
  const float* ap = a.xyzA + static_cast<size_t>(b) * NA * 3;
  const float* bp = a.xyzB + static_cast<size_t>(b) * NB * 3;
  float d = 0.f;
  bool live = false;
  if (side == 0 && k < NB) {
    live = true;
    unsigned bits = 0xffffffffu, mask = 0u;
    int wrb = 0;
    for (int rr = 0; rr < a.RB; ++rr) {
      const uint2 e = __ldg(a.colpart + (static_cast<size_t>(b) * a.RB + rr) * a.NB8 + k);
      if (e.y < bits) { bits = e.y; mask = e.x; wrb = rr; }
    }
    const int j0 = wrb * ROWS + (__ffs(mask) - 1) * R;
    const float* q = bp + static_cast<size_t>(k) * 3;
    const float rx = __ldg(q), ry = __ldg(q + 1), rz = __ldg(q + 2);
    int bi = j0;
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
      const int j = min(j0 + u, NA - 1);
      const float* pr = ap + static_cast<size_t>(j) * 3;
      const float dd = dist_yxz(rx - __ldg(pr), ry - __ldg(pr + 1), rz - __ldg(pr + 2));
      if (__float_as_uint(dd) == bits && j0 + u < NA) bi = j0 + u;
    }
    d = __uint_as_float(bits);
    a.distB[static_cast<size_t>(b) * NB + k] = d;
    a.idxB[static_cast<size_t>(b) * NB + k] = bi;
  } else if (side == 1 && k < NA) {
    live = true;
    if (a.nchunks == 1) {
      d = a.distA[static_cast<size_t>(b) * NA + k];
    } else {
      unsigned bits = 0xffffffffu, gg = 0;
      for (int zz = 0; zz < a.nchunks; ++zz) {
        const uint2 e = __ldg(a.rowpart + (static_cast<size_t>(b) * a.nchunks + zz) * a.NA8 + k);
        if (e.y < bits) { bits = e.y; gg = e.x; }
      }
      const float px0 = __ldg(ap + 3 * k), py0 = __ldg(ap + 3 * k + 1), pz0 = __ldg(ap + 3 * k + 2);
      int bi = static_cast<int>(gg);
#pragma unroll
      for (int u = CG - 1; u >= 0; --u) {
        const int kcol = min(static_cast<int>(gg) + u, NB - 1);
        const float* q = bp + static_cast<size_t>(kcol) * 3;
        const float dd = dist_yxz(__ldg(q) - px0, __ldg(q + 1) - py0, __ldg(q + 2) - pz0);
        if (__float_as_uint(dd) == bits && static_cast<int>(gg) + u < NB) bi = static_cast<int>(gg) + u;
      }
      d = __uint_as_float(bits);
      a.distA[static_cast<size_t>(b) * NA + k] = d;
      a.idxA[static_cast<size_t>(b) * NA + k] = bi;
    }
  }
  if (a.sums == nullptr) return;
  const unsigned per_side = gridDim.x * gridDim.y;
  const unsigned cta = side * per_side + blockIdx.y * gridDim.x + blockIdx.x;
  const float2 part = block_sum2_256(d, live ? __fsqrt_rn(d) : 0.f, s_w);
  if (threadIdx.x == 0) {
    a.partials[cta] = part;
    __threadfence();
    s_last = (atomicAdd(a.tickets, 1u) + 1u == per_side * gridDim.z);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float2 side_sum[2];
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
    float a0 = 0.f, a1 = 0.f;
    const volatile float2* pp = a.partials + sd * per_side;
    for (unsigned i = threadIdx.x; i < per_side; i += 256) {
      a0 += pp[i].x;
      a1 += pp[i].y;
    }
    side_sum[sd] = block_sum2_256(a0, a1, s_w);  // (valid in thread 0)
  }
  // side 1 = cloud A = the caller's xyz1 unless the clouds were swapped; output { d1, d2, sqrt d1, sqrt d2 }
  const float2 s1 = a.swapped ? side_sum[0] : side_sum[1], s2 = a.swapped ? side_sum[1] : side_sum[0];
  float loc[4] = {s1.x, s2.x, s1.y, s2.y};
  if (px.world > 1) peer_allreduce4(px, loc, s_x);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) a.sums[q] = loc[q];
  }
}

static PeerXchg make_px(const upp_peer_exchange* peers) {
  PeerXchg px;
  px.world = 1;
  px.rank = 0;
  px.seq = nullptr;
  px.defer = 0;
  px.status = nullptr;
  px.timeout = 0;
  for (int r = 0; r < UPP_MAX_PEERS; ++r) px.slots[r] = nullptr;
  if (peers != nullptr && peers->world > 1) {
    px.world = peers->world;
    px.rank = peers->rank;
    px.seq = peers->seq;
    px.defer = peers->defer ? 1 : 0;
    px.status = peers->status;
    px.timeout = peers->timeout_cycles;
    for (int r = 0; r < peers->world; ++r) px.slots[r] = peers->slots[r];
  }
  return px;
}

template <int R, int THREADS>
static void launch_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                               float* dist2, int32_t* idx1, int32_t* idx2, cudaStream_t st) {
  const int per = R * THREADS;
  dim3 grid((max(N, M) + per - 1) / per, B, 2);
  chamfer_fwd_kernel<R, THREADS><<<grid, THREADS, 0, st>>>(xyz1, xyz2, N, M, dist1, dist2, idx1, idx2);
}

// Column chunks for the packed kernel.  An SM keeps up to 5 of these CTAs resident and needs several of
// them (one warp per scheduler each) to keep its FMA pipe fed, so the grid should hold at least one full
// residency wave (5 x 148 items) and then fill whole waves.  B200 sweep (scripts/time_ops.py
// --sweep-chamfer): B32 1024x1024 1 chunk 33.8 us -> 4 chunks 27.6; B64 2048x2048 1 chunk 107.5 -> 4 chunks
// 93.2; B64 2048x8192 (2048 items already) best with 1.  Chunks stay >= 256 columns and a multiple of 64.
static int chamfer_pick_chunks(long items0, int NB, int per_sm = 5) {
  const long wave = static_cast<long>(per_sm) * 148;
  const int cap = NB / 256 < 1 ? 1 : (NB / 256 > 32 ? 32 : NB / 256);
  long cmin = (wave + items0 - 1) / items0;
  if (cmin < 1) cmin = 1;
  if (cmin >= cap) return cap;
  int best = static_cast<int>(cmin);
  double best_eff = 0.0;
  for (int c = static_cast<int>(cmin); c <= cap && c <= 2 * cmin; ++c) {
    const long items = items0 * c;
    const double eff = static_cast<double>(items) / (static_cast<double>(wave) * ((items + wave - 1) / wave));
    if (eff > best_eff + 0.03) { best_eff = eff; best = c; }  // fewest chunks within 3 % of the best fill
  }
  return best;
}

static int env_chunks() {
  const char* v = tuning_env("UPP_CH_CHUNKS");  // tuning aid: force the number of column chunks
  return v ? atoi(v) : 0;
}

// Plan of the fused single-kernel path (chamfer_fwd_fused_kernel<8, 4, 16>): rows = the larger cloud, columns = the
// smaller one cut into chunks; one function serves the workspace query and the launch so that they always agree.
struct ChFusedPlan {
  int na, nb, na8, nb8, rowblocks, chunks, chunk;
  size_t colpart, rowpart, partials, tickets, bytes;  // byte offsets / total
};
constexpr int kFusedR = 8, kFusedWC = 4, kFusedCG = 16;
constexpr size_t kFusedWorkspaceCap = static_cast<size_t>(512) << 20;  // beyond this the keyed two-launch path runs

static ChFusedPlan chamfer_fused_plan(int B, int N, int M) {
  ChFusedPlan p;
  p.na = N >= M ? N : M;
  p.nb = N >= M ? M : N;
  p.na8 = (p.na + 15) & ~15;
  p.nb8 = (p.nb + 15) & ~15;  // column slots are written in whole groups of kFusedCG
  p.rowblocks = (p.na + 32 * kFusedR - 1) / (32 * kFusedR);
  const int fc = env_chunks();
  int chunks = fc > 0 ? fc : chamfer_pick_chunks(static_cast<long>(B) * p.rowblocks, p.nb, 5);
  p.chunk = ((p.nb + chunks - 1) / chunks + 63) & ~63;
  p.chunks = (p.nb + p.chunk - 1) / p.chunk;
  const size_t slots = static_cast<size_t>(B) * (p.rowblocks + p.chunks);
  p.colpart = 0;
  p.rowpart = p.colpart + static_cast<size_t>(B) * p.rowblocks * p.nb8 * 8;
  p.partials = p.rowpart + (p.chunks > 1 ? static_cast<size_t>(B) * p.chunks * p.na8 * 8 : 0);
  const size_t fin_ctas = 2 * static_cast<size_t>(B) * ((static_cast<size_t>(p.na) + 255) / 256);  // two-kernel form
  p.tickets = p.partials + (slots > fin_ctas ? slots : fin_ctas) * 8;
  p.bytes = p.tickets + (slots + 4) * 4;
  p.bytes = (p.bytes + 15) & ~static_cast<size_t>(15);
  return p;
}

static size_t chamfer_keyed_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  // one 64-bit key per point of either cloud (the column side always merges by key, the row side
  // when the columns are cut into chunks)
  const size_t n8 = (static_cast<size_t>(N) + 7) & ~static_cast<size_t>(7);
  const size_t m8 = (static_cast<size_t>(M) + 7) & ~static_cast<size_t>(7);
  // + the ticket word (16 B) and one float2 partial per finalize CTA (2 sides x B x ceil(max/256))
  const size_t ctas = 2 * static_cast<size_t>(B) * ((static_cast<size_t>(N > M ? N : M) + 255) / 256);
  return static_cast<size_t>(B) * (n8 + m8) * 8 + 16 + ctas * 8;
}

size_t chamfer_fwd_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  const size_t keyed = chamfer_keyed_workspace_bytes(B, N, M);
  const size_t fused = chamfer_fused_plan(B, N, M).bytes;
  return fused <= kFusedWorkspaceCap && fused > keyed ? fused : keyed;
}

template <int R, int WC, int MINB = 0>
static int launch_chamfer_packed(const float* a, const float* bpts, int B, int NA, int NB, float* dA,
                                 int32_t* iA, float* dB, int32_t* iB, void* ws, int force_chunks,
                                 float* sums, bool swapped, const PeerXchg& px, cudaStream_t st) {
  static_assert(R <= 16, "chamfer_finalize2_kernel scans at most 16 rows per key");
  const int na8 = (NA + 7) & ~7, nb8 = (NB + 7) & ~7;
  const int rowblocks = (NA + 32 * R - 1) / (32 * R);
  int chunks = force_chunks > 0 ? force_chunks : chamfer_pick_chunks(static_cast<long>(B) * rowblocks, NB, MINB > 1 ? MINB : 5);
  int chunk = ((NB + chunks - 1) / chunks + 63) & ~63;
  chunks = (NB + chunk - 1) / chunk;
  // workspace: [col keys B*nb8][row keys B*na8][ticket, 16 B][partials]; keys and ticket preset to 0xFF
  unsigned long long* colkeys = static_cast<unsigned long long*>(ws);
  unsigned long long* rowkeys_all = colkeys + static_cast<size_t>(B) * nb8;
  unsigned long long* rowkeys = chunks > 1 ? rowkeys_all : nullptr;
  unsigned* ticket = reinterpret_cast<unsigned*>(rowkeys_all + static_cast<size_t>(B) * na8);
  float2* partials = reinterpret_cast<float2*>(ticket + 4);
  const size_t preset = static_cast<size_t>(B) * (nb8 + na8) * 8 + 16;
  cudaError_t e = cudaMemsetAsync(ws, 0xFF, preset, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(rowblocks, B, chunks);
  chamfer_fwd_packed_kernel<R, WC, MINB><<<grid, WC * 32, 0, st>>>(a, bpts, NA, NB, na8, nb8, chunk, dA, iA, rowkeys, colkeys);
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  const bool rows_too = chunks > 1 || sums != nullptr;
  dim3 fgrid(((rows_too ? max(NA, NB) : NB) + 255) / 256, B, rows_too ? 2 : 1);
  chamfer_finalize2_kernel<<<fgrid, 256, 0, st>>>(a, bpts, NA, NB, na8, nb8, colkeys, rowkeys, R, dA, iA, dB, iB,
                                                  sums ? partials : nullptr, ticket, sums, swapped ? 1 : 0, px);
  count_launch();
  return launch_status();
}

template <int R, int WC>
static int launch_chamfer_both(const float* a, const float* bpts, int B, int NA, int NB, float* dA,
                               int32_t* iA, float* dB, int32_t* iB, void* ws, cudaStream_t st) {
  const int nb8 = (NB + 7) & ~7;
  const size_t keys = static_cast<size_t>(B) * nb8 * 8;
  cudaError_t e = cudaMemsetAsync(ws, 0xFF, keys, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid((NA + 32 * R - 1) / (32 * R), B);
  chamfer_fwd_both_kernel<R, WC><<<grid, WC * 32, 0, st>>>(a, bpts, NA, NB, nb8, dA, iA, dB, iB,
                                                          static_cast<unsigned long long*>(ws));
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  chamfer_cols_finalize_kernel<R><<<dim3((NB + 255) / 256, B), 256, 0, st>>>(
      a, bpts, NA, NB, nb8, static_cast<const unsigned long long*>(ws), dB, iB);
  return UPP_OK;
}


// The fused single-kernel forward: tickets zeroed (a few KB), then ONE launch.
static int launch_chamfer_fused(const float* a, const float* bpts, int B, const ChFusedPlan& p, float* dA, int32_t* iA,
                                float* dB, int32_t* iB, void* ws, float* sums, bool swapped, const PeerXchg& px,
                                cudaStream_t st) {
  char* w = static_cast<char*>(ws);
  ChFused f;
  f.xyzA = a; f.xyzB = bpts;
  f.NA = p.na; f.NB = p.nb; f.NA8 = p.na8; f.NB8 = p.nb8; f.chunk = p.chunk; f.nchunks = p.chunks; f.RB = p.rowblocks;
  f.distA = dA; f.idxA = iA; f.distB = dB; f.idxB = iB;
  f.colpart = reinterpret_cast<uint2*>(w + p.colpart);
  f.rowpart = reinterpret_cast<uint2*>(w + p.rowpart);
  f.partials = reinterpret_cast<float2*>(w + p.partials);
  f.tickets = reinterpret_cast<unsigned*>(w + p.tickets);
  f.sums = sums;
  f.swapped = swapped ? 1 : 0;
  const size_t slots = static_cast<size_t>(B) * (p.rowblocks + p.chunks);
  if (tuning_env_int("UPP_CH_FUSED", 0) >= 1 && tuning_env_int("UPP_CH_FUSED", 0) <= 4) {  // folded forms count contributors
    cudaError_t e = cudaMemsetAsync(f.tickets, 0, (slots + 4) * 4, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  dim3 grid(p.rowblocks, B, p.chunks);
  // default: the two-kernel form (distance kernel without an epilogue, then one finalize launch at full occupancy).
  // Folding the finalize into the distance kernel (tickets, last contributor resolves) was built and measured slower:
  // its CTAs hold their SM slots through latency-bound tails (B64 2048^2 + sums: keyed 92.3 us, folded 97-112 us,
  // two-kernel slots 88-92 us; profiles/r02_quick_ops.jsonl).  UPP_CH_FUSED: 1 = folded cg8 6 CTAs/SM, 2 = folded cg16
  // 5/SM, 3 = folded cg8 5/SM, 4 = folded cg16 6/SM, 5 = two-kernel cg16, 6 = two-kernel cg8 capped for 6/SM (tuning / tests).
  const int v = tuning_env_int("UPP_CH_FUSED", 0);
  const bool fold = v >= 1 && v <= 4;
  if (fold) {
    if (v == 1) chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 8, 6><<<grid, kFusedWC * 32, 0, st>>>(f, px);
    else if (v == 2) chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 16, 0><<<grid, kFusedWC * 32, 0, st>>>(f, px);
    else if (v == 3) chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 8, 0><<<grid, kFusedWC * 32, 0, st>>>(f, px);
    else chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 16, 6><<<grid, kFusedWC * 32, 0, st>>>(f, px);
    count_launch();
    return launch_status();
  }
  const int cg = v == 5 ? 16 : 8;
  if (v == 5) chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 16, 6, false><<<grid, kFusedWC * 32, 0, st>>>(f, px);
  else if (v == 6) chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 8, 6, false><<<grid, kFusedWC * 32, 0, st>>>(f, px);
  else chamfer_fwd_fused_kernel<kFusedR, kFusedWC, 8, 0, false><<<grid, kFusedWC * 32, 0, st>>>(f, px);
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK) return rc;
  const bool rows_too = p.chunks > 1 || sums != nullptr;
  dim3 fgrid(((rows_too ? p.na : p.nb) + 255) / 256, B, rows_too ? 2 : 1);
  if (cg == 16) chamfer_finalize3_kernel<kFusedR, 16><<<fgrid, 256, 0, st>>>(f, px);
  else chamfer_finalize3_kernel<kFusedR, 8><<<fgrid, 256, 0, st>>>(f, px);
  count_launch();
  return launch_status();
}

int chamfer_fwd_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* sums, void* workspace,
                       size_t workspace_bytes, const upp_peer_exchange* peers, cudaStream_t st) {
  const PeerXchg px = make_px(peers);
  const char* v = tuning_env("UPP_CH_VARIANT");  // tuning aid
  const int variant = v ? atoi(v) : -1;
  const bool have_ws = workspace != nullptr && workspace_bytes >= chamfer_fwd_workspace_bytes(B, N, M);
  // single pass: rows = the larger cloud (fills the 32*R-row tiles), columns = the smaller one (any size);
  // two tiny clouds stay on the directed kernel
  const bool single = have_ws && max(N, M) >= 128 && (variant < 0 || variant >= 20);
  // the fused peer all-reduce lives in the packed path's finalize kernel
  if (px.world > 1 && !(single && (variant < 0 || variant >= 30) && sums != nullptr)) return UPP_ERR_UNSUPPORTED;
  if (single) {
    // rows = the larger cloud (more CTAs), columns = the smaller one (fewer keys)
    const bool swap = M > N;
    const float* a = swap ? xyz2 : xyz1;
    const float* c = swap ? xyz1 : xyz2;
    const int na = swap ? M : N, nb = swap ? N : M;
    float* dA = swap ? dist2 : dist1; float* dB = swap ? dist1 : dist2;
    int32_t* iA = swap ? idx2 : idx1; int32_t* iB = swap ? idx1 : idx2;
    const long ctas8 = static_cast<long>(B) * ((na + 255) / 256);
    int rc;
    const ChFusedPlan plan = chamfer_fused_plan(B, N, M);
    const bool fused_ok = plan.bytes <= kFusedWorkspaceCap && workspace_bytes >= plan.bytes && B <= 65535;
    // default: the fused single-kernel path; 30 = the keyed two-launch packed path (round 1; also the fallback when the
    // partial slots would not fit the workspace cap); 20..22 = the scalar single-pass kernel (A/B timing)
    int pick = variant >= 20 ? variant : (fused_ok ? 50 : 30);
    if (pick == 50 && !fused_ok) pick = 30;
    const int fc = env_chunks();
    (void)ctas8;
    switch (pick) {
      case 20: rc = launch_chamfer_both<8, 8>(a, c, B, na, nb, dA, iA, dB, iB, workspace, st); break;
      case 21: rc = launch_chamfer_both<4, 8>(a, c, B, na, nb, dA, iA, dB, iB, workspace, st); break;
      case 22: rc = launch_chamfer_both<8, 4>(a, c, B, na, nb, dA, iA, dB, iB, workspace, st); break;
      case 31: rc = launch_chamfer_packed<4, 8>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 32: rc = launch_chamfer_packed<8, 8>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 33: rc = launch_chamfer_packed<6, 8>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      // measured best on B200: 8 rows per thread, 4 warps per CTA (5 CTAs / SM)
      case 34: rc = launch_chamfer_packed<12, 4>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 35: rc = launch_chamfer_packed<16, 4>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 36: rc = launch_chamfer_packed<16, 2>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 37: rc = launch_chamfer_packed<8, 4, 6>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 38: rc = launch_chamfer_packed<8, 4, 4>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 39: rc = launch_chamfer_packed<8, 4, 3>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 40: rc = launch_chamfer_packed<8, 8, 3>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
      case 50: rc = launch_chamfer_fused(a, c, B, plan, dA, iA, dB, iB, workspace, sums, swap, px, st); break;
      default: rc = launch_chamfer_packed<8, 4>(a, c, B, na, nb, dA, iA, dB, iB, workspace, fc, sums, swap, px, st); break;
    }
    if (rc != UPP_OK) return rc;
    if (pick >= 30) return UPP_OK;  // the packed path's finalize kernel has produced the sums too
  } else {
    switch (variant) {
      case 1: launch_chamfer_fwd<4, 64>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      case 5: launch_chamfer_fwd<4, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      case 6: launch_chamfer_fwd<2, 256>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      case 7: launch_chamfer_fwd<2, 64>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      case 8: launch_chamfer_fwd<1, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      case 9: launch_chamfer_fwd<1, 256>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
      // measured best on B200 (scripts/time_ops.py --sweep-chamfer): more warps beat more queries per thread
      default: launch_chamfer_fwd<2, 128>(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, st); break;
    }
  }
  count_launch();
  int rc = launch_status();
  if (rc != UPP_OK || sums == nullptr) return rc;
  chamfer_sums_kernel<<<kSumCluster, kSumThreads, 0, st>>>(dist1, static_cast<size_t>(B) * N, dist2,
                                                          static_cast<size_t>(B) * M, sums);
  count_launch();
  return launch_status();
}

int peer_finish_launch(const upp_peer_exchange* peers, float* global_sums, cudaStream_t st) {
  PeerXchg px = make_px(peers);
  px.defer = 0;
  peer_finish_kernel<<<1, 32, 0, st>>>(px, global_sums);
  count_launch();
  return launch_status();
}

size_t chamfer_bwd_stats_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  // ticket (16 B) + one float per CTA; CTAs per cloud <= ceil(N / 32) + ceil(M / 32) (one-warp CTAs, one destination per lane, at worst)
  return 16 + static_cast<size_t>(B) * ((N + 31) / 32 + (M + 31) / 32 + 2) * sizeof(float);
}

// sq_out == nullptr: plain backward.  Otherwise sq_out[0..1] = sum ||grad_xyz1||^2, sum ||grad_xyz2||^2 (over all ranks
// when peers is given), workspace as sized by chamfer_bwd_stats_workspace_bytes.
int chamfer_bwd_launch(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                       const float* g1, const float* g2, int B, int N, int M, float* gx1, float* gx2,
                       float* sq_out, void* workspace, size_t workspace_bytes, const upp_peer_exchange* peers,
                       cudaStream_t st) {
  // one grid for both sides: the same CTA shape serves both (the smaller side just needs fewer CTAs)
  const int q = scatter_pick_q(B, N + M);
  const ScatterGrid a = scatter_grid(N, M, q), c = scatter_grid(M, N, q);
  const int warps = a.warps > c.warps ? a.warps : c.warps;
  const int per = warps * 32 * q;
  const int blocks1 = (N + per - 1) / per, blocks2 = (M + per - 1) / per;
  const int lmax = N > M ? N : M;  // the longer list sizes the staging area for both sides
  const size_t smem = scatter_smem(lmax, warps, q);
  unsigned* ticket = nullptr;
  float* partials = nullptr;
  if (sq_out != nullptr) {
    if (workspace == nullptr || workspace_bytes < chamfer_bwd_stats_workspace_bytes(B, N, M)) return UPP_ERR_WORKSPACE;
    ticket = static_cast<unsigned*>(workspace);
    partials = reinterpret_cast<float*>(ticket + 4);
    cudaError_t e = cudaMemsetAsync(ticket, 0, 16, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const PeerXchg px = make_px(peers);
  cudaError_t le = q == 1 ? launch_pdl(chamfer_bwd_kernel<1>, dim3(blocks1 + blocks2, B), dim3(warps * 32), smem, st, xyz1, xyz2, idx1, idx2,
                                       g1, g2, N, M, blocks1, gx1, gx2, partials, ticket, sq_out, px)
                          : launch_pdl(chamfer_bwd_kernel<2>, dim3(blocks1 + blocks2, B), dim3(warps * 32), smem, st, xyz1, xyz2, idx1, idx2,
                                       g1, g2, N, M, blocks1, gx1, gx2, partials, ticket, sq_out, px);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return launch_status();
}

int peer_allreduce_launch(const upp_peer_exchange* peers, const float* local, float* global_sums, cudaStream_t st) {
  peer_allreduce_kernel<<<1, 256, 0, st>>>(make_px(peers), local, global_sums);
  count_launch();
  return launch_status();
}

}  // namespace upp
