// fps_round.cuh -- the pieces of one farthest-point-sampling round shared by fps.cu (the FPS kernels) and group.cu
// (the single-launch Group divider): upstream's initial running min-distance with its |p|^2 <= 1e-3 skip rule, and the
// packed fp32x2 distance update + in-thread arg-max of a span of a thread's points.
#pragma once
#include "common.cuh"

namespace upp {

constexpr float kSkipped = -1.0f;      // |p|^2 <= 1e-3 : bits 0xBF800000, s32 -1082130432
constexpr float kOutOfRange = -0.5f;   // slot >= N     : bits 0xBF000000, s32 -1090519040 (lower)

__device__ __forceinline__ float fps_initial_md(float x, float y, float z) {
  const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
  return (static_cast<double>(mag) <= 1e-3) ? kSkipped : 1e10f;
}

// One span [R0, R1) of a thread's point pairs: distances to the new centre, running-min update,
// span maximum (order-preserving float bits) and the lowest slot holding it.  TREE: independent
// compare/selects + a VIMNMX3 tree (short dependent chain) instead of one select chain.
template <int P2, int R0, int R1, bool TREE>
__device__ __forceinline__ void fps_span(const f32x2 (&X)[P2], const f32x2 (&Y)[P2], const f32x2 (&Z)[P2],
                                         float (&md)[2 * P2], f32x2 CX, f32x2 CY, f32x2 CZ, int& best,
                                         int& ls) {
  constexpr int NP = R1 - R0, NS = 2 * NP;
  // all packed distance chains are independent: spelled stage by stage so that they are scheduled
  // interleaved (one warp per scheduler has nobody else to hide a dependent chain)
  f32x2 D[NP];
#pragma unroll
  for (int r = 0; r < NP; ++r) D[r] = sub2(Y[R0 + r], CY);
#pragma unroll
  for (int r = 0; r < NP; ++r) D[r] = mul2(D[r], D[r]);
#pragma unroll
  for (int r = 0; r < NP; ++r) { const f32x2 dx = sub2(X[R0 + r], CX); D[r] = fma2(dx, dx, D[r]); }
#pragma unroll
  for (int r = 0; r < NP; ++r) { const f32x2 dz = sub2(Z[R0 + r], CZ); D[r] = fma2(dz, dz, D[r]); }
  int key[NS];
#pragma unroll
  for (int r = 0; r < NP; ++r) {
    float d0, d1;
    unpack2(D[r], d0, d1);
    md[2 * (R0 + r)] = fminf(md[2 * (R0 + r)], d0);
    md[2 * (R0 + r) + 1] = fminf(md[2 * (R0 + r) + 1], d1);
    key[2 * r] = __float_as_int(md[2 * (R0 + r)]);
    key[2 * r + 1] = __float_as_int(md[2 * (R0 + r) + 1]);
  }
  int red[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) red[s] = key[s];
#pragma unroll
  for (int n = NS; n > 1; n = (n + 2) / 3) {  // balanced VIMNMX3 tree
#pragma unroll
    for (int q = 0; q < (n + 2) / 3; ++q) {
      int v = red[3 * q];
      if (3 * q + 1 < n) v = max(v, red[3 * q + 1]);
      if (3 * q + 2 < n) v = max(v, red[3 * q + 2]);
      red[q] = v;
    }
  }
  best = red[0];
  if constexpr (TREE) {
    int cnd[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) cnd[s] = key[s] == best ? 2 * R0 + s : 2 * P2;
#pragma unroll
    for (int n = NS; n > 1; n = (n + 2) / 3) {
#pragma unroll
      for (int q = 0; q < (n + 2) / 3; ++q) {
        int v = cnd[3 * q];
        if (3 * q + 1 < n) v = min(v, cnd[3 * q + 1]);
        if (3 * q + 2 < n) v = min(v, cnd[3 * q + 2]);
        cnd[q] = v;
      }
    }
    ls = cnd[0];
  } else {
    ls = 2 * R0 + NS - 1;
#pragma unroll
    for (int s = NS - 2; s >= 0; --s)
      if (key[s] == best) ls = 2 * R0 + s;
  }
}

}  // namespace upp
