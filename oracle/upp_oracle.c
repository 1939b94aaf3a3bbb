/*
 * upp_oracle.c -- CPU restatement of the UPP point-geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing outside tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link, load or call
 * this file.  The product path (iccv2025-upp_b200/) never routes through it.
 *
 * Parity status per function (see DESIGN.md "Oracle"):
 *   chamfer_fwd / chamfer_bwd : PINNED   -- follows /root/reference/extensions/
 *       chamfer_dist/chamfer.cu, and is checked on the GPU box against that
 *       very file compiled unmodified into oracle/_ref/chamfer_ref*.so, plus
 *       golden vectors generated from the reference's Python modules
 *       (tests/golden/make_golden.py).
 *   fps / gather / knn        : PARITY UNPINNED -- the arithmetic lives in two
 *       third-party packages whose sources are NOT under /root/reference:
 *         pointnet2_ops 3.0.0 (erikwijmans/Pointnet2_PyTorch, un-pinned commit,
 *                              reference README.md:73)
 *         KNN_CUDA 0.2        (unlimblue/KNN_CUDA wheel, reference README.md:76)
 *       Their published algorithms are restated below; parity is anchored on
 *       the reference's call sites (utils/misc.py:13-20,
 *       models/Point_MAE_unify.py:51-92) and on its own pure-torch /
 *       numpy formulations (models/modules.py:13-51, models/dgcnn_group.py:8-19,
 *       datasets/ModelNetDataset.py:29-50).
 *
 * Floating point: every distance is spelled with explicit fmaf() in the order
 * nvcc 12.9 contracts the upstream expressions for sm_100a (verified in SASS):
 *   chamfer / fps : d   = fma(dz,dz, fma(dx,dx, dy*dy))
 *   fps skip test : mag = fma(z,z,  fma(x,x,  y*y)),  (double)mag <= 1e-3
 *   knn           : d   = fma(dz,dz, fma(dy,dy, fma(dx,dx, 0)))
 * Build with -ffp-contract=off so gcc adds no contraction of its own.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ---- tiny pthread parallel-for over clouds (libgomp is not in the image) */

typedef void (*cloud_fn)(int b, void* ctx);
typedef struct { cloud_fn fn; void* ctx; int B; int next; pthread_mutex_t mu; } pf_t;

static int g_threads = 0; /* 0 = all online cores */

ORACLE_API void upp_oracle_set_threads(int n) { g_threads = n; }
ORACLE_API int upp_oracle_get_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static void* pf_worker(void* arg) {
  pf_t* pf = (pf_t*)arg;
  for (;;) {
    pthread_mutex_lock(&pf->mu);
    int b = pf->next++;
    pthread_mutex_unlock(&pf->mu);
    if (b >= pf->B) return NULL;
    pf->fn(b, pf->ctx);
  }
}

static void parallel_for(int B, cloud_fn fn, void* ctx) {
  int nt = upp_oracle_get_threads();
  if (nt > B) nt = B;
  if (nt <= 1) { for (int b = 0; b < B; ++b) fn(b, ctx); return; }
  pf_t pf = {fn, ctx, B, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, pf_worker, &pf);
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
  free(th);
}

/* ---- distance forms ---------------------------------------------------- */

/* chamfer.cu:40-43 `x2*x2 + y2*y2 + z2*z2` with x2 = ref - query, as nvcc
 * contracts it; upstream sampling_gpu.cu uses the same expression shape. */
static inline float dist_yxz(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* KNN_CUDA knn.cu cuComputeDistanceGlobal: ssd = 0; ssd += tmp*tmp over
 * dim = x,y,z (padding rows are 0-0). */
static inline float dist_xyz_acc(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, 0.0f)));
}

/* ---- farthest point sampling ------------------------------------------ */

/*
 * Restates upstream furthest_point_sampling_kernel (pointnet2_ops
 * _ext-src/src/sampling_gpu.cu), called from reference utils/misc.py:18.
 *   - first sample is index 0
 *   - temp[k] starts at 1e10, d2 = min(d, temp[k]), strict `d2 > best`
 *   - points with x^2+y^2+z^2 <= 1e-3 (compared in double) are skipped:
 *     never selected, temp never updated
 *   - no valid point at all -> index 0
 * block_size == 0 : global lowest-index tie-break (the contract BASELINE.json
 *                   north_star states; what the CUDA kernel implements).
 * block_size  > 0 : emulate upstream's thread-strided scan plus shared-memory
 *                   tree reduction with that many threads (power of two);
 *                   differs from 0 only on exact float ties.
 */
typedef struct { const float* xyz; int N, M; int32_t* idx_out; int block_size; } fps_ctx;

static void fps_cloud(int b, void* vctx) {
  const fps_ctx* c = (const fps_ctx*)vctx;
  const float* xyz = c->xyz; const int N = c->N, M = c->M, block_size = c->block_size;
  int32_t* idx_out = c->idx_out;
  {
    const float* p = xyz + (size_t)b * N * 3;
    int32_t* out = idx_out + (size_t)b * M;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    unsigned char* skip = (unsigned char*)malloc((size_t)N);
    int bs = block_size > 0 ? block_size : 1;
    float* tv = (float*)malloc(sizeof(float) * (size_t)bs);
    int* ti = (int*)malloc(sizeof(int) * (size_t)bs);
    for (int k = 0; k < N; ++k) {
      temp[k] = 1e10f;
      float x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
      float mag = fmaf(z, z, fmaf(x, x, y * y));
      skip[k] = ((double)mag <= 1e-3) ? 1 : 0;
    }
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < M; ++j) {
      float x1 = p[3 * old], y1 = p[3 * old + 1], z1 = p[3 * old + 2];
      if (block_size <= 0) {
        float best = -1.0f;
        int besti = 0;
        for (int k = 0; k < N; ++k) {
          if (skip[k]) continue;
          float d = dist_yxz(p[3 * k] - x1, p[3 * k + 1] - y1, p[3 * k + 2] - z1);
          float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
        old = besti;
      } else {
        for (int t = 0; t < bs; ++t) {
          float best = -1.0f;
          int besti = 0;
          for (int k = t; k < N; k += bs) {
            if (skip[k]) continue;
            float d = dist_yxz(p[3 * k] - x1, p[3 * k + 1] - y1, p[3 * k + 2] - z1);
            float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            besti = d2 > best ? k : besti;
            best = d2 > best ? d2 : best;
          }
          tv[t] = best;
          ti[t] = besti;
        }
        for (int half = bs / 2; half >= 1; half /= 2) {
          for (int t = 0; t < half; ++t) {
            float v1 = tv[t], v2 = tv[t + half];
            int i1 = ti[t], i2 = ti[t + half];
            tv[t] = v1 > v2 ? v1 : v2;
            ti[t] = v2 > v1 ? i2 : i1;
          }
        }
        old = ti[0];
      }
      out[j] = old;
    }
    free(temp); free(skip); free(tv); free(ti);
  }
}

ORACLE_API void upp_oracle_fps(const float* xyz, int B, int N, int M,
                               int32_t* idx_out, int block_size) {
  if (M <= 0 || N <= 0) return;
  fps_ctx c = {xyz, N, M, idx_out, block_size};
  parallel_for(B, fps_cloud, &c);
}

/* upstream opt_n_threads(): largest power of two <= N, clamped to [1, 512]. */
ORACLE_API int upp_oracle_fps_upstream_block(int N) {
  int p = 1;
  while (p * 2 <= N && p * 2 <= 512) p *= 2;
  return p;
}

/* ---- gather ------------------------------------------------------------ */

/* upstream gather_points_kernel: out[b,c,j] = feat[b,c,idx[b,j]]
 * (channel-first; reference utils/misc.py:19). */
ORACLE_API void upp_oracle_gather(const float* feat, const int32_t* idx, int B,
                                  int C, int N, int M, float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < M; ++j)
        out[((size_t)b * C + c) * M + j] =
            feat[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

/* upstream gather_points_grad_kernel: gfeat[b,c,idx[b,j]] += gout[b,c,j],
 * summed here in ascending j (upstream uses atomicAdd: order unspecified). */
ORACLE_API void upp_oracle_gather_grad(const float* gout, const int32_t* idx,
                                       int B, int C, int N, int M, float* gfeat) {
  memset(gfeat, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < M; ++j)
        gfeat[((size_t)b * C + c) * N + idx[(size_t)b * M + j]] +=
            gout[((size_t)b * C + c) * M + j];
}

/* ---- kNN --------------------------------------------------------------- */

/*
 * Restates KNN_CUDA 0.2 (knn.cu: cuComputeDistanceGlobal -> cuInsertionSort ->
 * cuParallelSqrt) as driven by knn_cuda.KNN(k, transpose_mode=True).forward
 * (reference models/Point_MAE_unify.py:56,69): per cloud, for every query the
 * full column of squared distances to the N refs, then upstream's own
 * two-phase insertion sort (strict '<' against the running k-th, insert before
 * the first strictly greater entry => ascending, ties keep the lower index
 * first), then sqrtf, indices converted to 0-based int64.
 * Layout: ref (B,N,3), query (B,Q,3), dist/idx (B,Q,k).  Returns -1 if k > N
 * (upstream reads out of bounds there; the replacement rejects it).
 */
typedef struct { const float* ref; const float* query; int N, Q, k; float* dist_out; int64_t* idx_out; int squared; } knn_ctx;

static void knn_cloud(int b, void* vctx) {
  const knn_ctx* c = (const knn_ctx*)vctx;
  const float* ref = c->ref; const float* query = c->query;
  const int N = c->N, Q = c->Q, k = c->k;
  float* dist_out = c->dist_out; int64_t* idx_out = c->idx_out;
  {
    float* col = (float*)malloc(sizeof(float) * (size_t)N);
    int64_t* ind = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    for (int q = 0; q < Q; ++q) {
      const float* qp = query + ((size_t)b * Q + q) * 3;
      for (int n = 0; n < N; ++n) {
        const float* rp = ref + ((size_t)b * N + n) * 3;
        col[n] = dist_xyz_acc(rp[0] - qp[0], rp[1] - qp[1], rp[2] - qp[2]);
      }
      /* cuInsertionSort, part 1: sort the first k entries */
      float max_dist = col[0];
      ind[0] = 1;
      for (int l = 1; l < k; ++l) {
        float curr = col[l];
        if (curr < max_dist) {
          int i = l - 1;
          for (int a = 0; a < l - 1; ++a)
            if (col[a] > curr) { i = a; break; }
          for (int j = l; j > i; --j) { col[j] = col[j - 1]; ind[j] = ind[j - 1]; }
          col[i] = curr;
          ind[i] = l + 1;
        } else {
          ind[l] = l + 1;
        }
        max_dist = col[l];
      }
      /* part 2: insert the remaining N-k entries */
      for (int l = k; l < N; ++l) {
        float curr = col[l];
        if (curr < max_dist) {
          int i = k - 1;
          for (int a = 0; a < k - 1; ++a)
            if (col[a] > curr) { i = a; break; }
          for (int j = k - 1; j > i; --j) { col[j] = col[j - 1]; ind[j] = ind[j - 1]; }
          col[i] = curr;
          ind[i] = l + 1;
          max_dist = col[k - 1];
        }
      }
      for (int j = 0; j < k; ++j) {
        dist_out[((size_t)b * Q + q) * k + j] = c->squared ? col[j] : sqrtf(col[j]);
        idx_out[((size_t)b * Q + q) * k + j] = ind[j] - 1;
      }
    }
    free(col); free(ind);
  }
}

ORACLE_API int upp_oracle_knn(const float* ref, const float* query, int B, int N,
                              int Q, int k, float* dist_out, int64_t* idx_out) {
  if (k > N || k <= 0) return -1;
  knn_ctx c = {ref, query, N, Q, k, dist_out, idx_out, 0};
  parallel_for(B, knn_cloud, &c);
  return 0;
}

/* pytorch3d.ops.knn_points convention (reference call site models/Point_MAE_pretask_dev.py:680; the
 * package itself is third-party and absent -> parity unpinned, like KNN_CUDA): same selection, SQUARED
 * distances, p1 = queries, p2 = references. */
ORACLE_API int upp_oracle_knn_points(const float* p1, const float* p2, int B, int N1, int N2, int K,
                                     float* dist2_out, int64_t* idx_out) {
  if (K > N2 || K <= 0) return -1;
  knn_ctx c = {p2, p1, N2, N1, K, dist2_out, idx_out, 1};
  parallel_for(B, knn_cloud, &c);
  return 0;
}

/* ---- Chamfer ----------------------------------------------------------- */

/* chamfer.cu:15-145 chamfer_dist_kernel, one direction: for each query j of
 * cloud A, min over refs of B and the lowest index attaining it (strict '<'
 * inside a tile, strict '>' across tiles). */
static void chamfer_dir(const float* a, int n, const float* bpts, int m,
                        float* dist, int32_t* idx) {
  for (int j = 0; j < n; ++j) {
    float x1 = a[3 * j], y1 = a[3 * j + 1], z1 = a[3 * j + 2];
    float best = 0.0f;
    int besti = 0;
    for (int k = 0; k < m; ++k) {
      float d = dist_yxz(bpts[3 * k] - x1, bpts[3 * k + 1] - y1, bpts[3 * k + 2] - z1);
      if (k == 0 || d < best) { best = d; besti = k; }
    }
    dist[j] = best;
    idx[j] = besti;
  }
}

typedef struct { const float* xyz1; const float* xyz2; int N, M; float* dist1; float* dist2; int32_t* idx1; int32_t* idx2; } cf_ctx;

static void cf_cloud(int b, void* vctx) {
  const cf_ctx* c = (const cf_ctx*)vctx;
  const float* a = c->xyz1 + (size_t)b * c->N * 3;
  const float* d = c->xyz2 + (size_t)b * c->M * 3;
  chamfer_dir(a, c->N, d, c->M, c->dist1 + (size_t)b * c->N, c->idx1 + (size_t)b * c->N);
  chamfer_dir(d, c->M, a, c->N, c->dist2 + (size_t)b * c->M, c->idx2 + (size_t)b * c->M);
}

/* chamfer.cu:147-171 chamfer_cuda_forward: outputs start as zeros (kept for
 * the degenerate m == 0 case), two directed passes. */
ORACLE_API void upp_oracle_chamfer_fwd(const float* xyz1, const float* xyz2,
                                       int B, int N, int M, float* dist1,
                                       float* dist2, int32_t* idx1, int32_t* idx2) {
  memset(dist1, 0, sizeof(float) * (size_t)B * N);
  memset(dist2, 0, sizeof(float) * (size_t)B * M);
  memset(idx1, 0, sizeof(int32_t) * (size_t)B * N);
  memset(idx2, 0, sizeof(int32_t) * (size_t)B * M);
  if (N <= 0 || M <= 0) return;
  cf_ctx c = {xyz1, xyz2, N, M, dist1, dist2, idx1, idx2};
  parallel_for(B, cf_cloud, &c);
}

/* chamfer.cu:173-201 chamfer_dist_grad_kernel, one pass:
 *   g = grad_dist[j]*2;  gA[j] += g*(A_j - B_idx);  gB[idx] += -(g*(A_j - B_idx))
 * Accumulated in ascending j (the CUDA kernels use float atomics: order
 * unspecified, so comparisons against this are tolerance-based).  inf*0 = NaN
 * is produced exactly where the reference produces it. */
static void chamfer_grad_dir(const float* a, int n, const float* bpts,
                             const float* g, const int32_t* idx, float* ga,
                             float* gb) {
  for (int j = 0; j < n; ++j) {
    int j2 = idx[j];
    float gg = g[j] * 2.0f;
    for (int c = 0; c < 3; ++c) {
      float v = gg * (a[3 * j + c] - bpts[3 * j2 + c]);
      ga[3 * j + c] += v;
      gb[3 * j2 + c] += -v;
    }
  }
}

/* chamfer.cu:203-229 chamfer_cuda_backward. */
ORACLE_API void upp_oracle_chamfer_bwd(const float* xyz1, const float* xyz2,
                                       const int32_t* idx1, const int32_t* idx2,
                                       const float* g1, const float* g2, int B,
                                       int N, int M, float* gx1, float* gx2) {
  memset(gx1, 0, sizeof(float) * (size_t)B * N * 3);
  memset(gx2, 0, sizeof(float) * (size_t)B * M * 3);
  for (int b = 0; b < B; ++b) {
    const float* a = xyz1 + (size_t)b * N * 3;
    const float* c = xyz2 + (size_t)b * M * 3;
    float* ga = gx1 + (size_t)b * N * 3;
    float* gc = gx2 + (size_t)b * M * 3;
    chamfer_grad_dir(a, N, c, g1 + (size_t)b * N, idx1 + (size_t)b * N, ga, gc);
    chamfer_grad_dir(c, M, a, g2 + (size_t)b * M, idx2 + (size_t)b * M, gc, ga);
  }
}

/* ---- Group divider (reference models/Point_MAE_unify.py:58-92) --------- */

/* FPS -> kNN -> gather -> subtract centre, composed from the pieces above;
 * neighborhood (B,G,k,3), center (B,G,3), idx (B,G,k) int64 local indices,
 * center_idx (B,G) int32. */
ORACLE_API int upp_oracle_group(const float* xyz, int B, int N, int G, int k,
                                float* neighborhood, float* center,
                                int64_t* idx, int32_t* center_idx) {
  upp_oracle_fps(xyz, B, N, G, center_idx, 0);
  for (int b = 0; b < B; ++b)
    for (int g = 0; g < G; ++g)
      for (int c = 0; c < 3; ++c)
        center[((size_t)b * G + g) * 3 + c] =
            xyz[((size_t)b * N + center_idx[(size_t)b * G + g]) * 3 + c];
  float* d = (float*)malloc(sizeof(float) * (size_t)B * G * k);
  int rc = upp_oracle_knn(xyz, center, B, N, G, k, d, idx);
  free(d);
  if (rc) return rc;
  for (int b = 0; b < B; ++b)
    for (int g = 0; g < G; ++g)
      for (int j = 0; j < k; ++j)
        for (int c = 0; c < 3; ++c) {
          size_t o = (((size_t)b * G + g) * k + j) * 3 + c;
          int64_t src = idx[((size_t)b * G + g) * k + j];
          neighborhood[o] = xyz[((size_t)b * N + src) * 3 + c] -
                            center[((size_t)b * G + g) * 3 + c];
        }
  return 0;
}

/* ---- k-nearest inverse-distance interpolation ------------------------------
 * Restates the reference's pure-torch
 *   propagate()                         models/Point_MAE_unify.py:22-48
 *   PointNetFeaturePropagation.forward  models/Point_MAE_unify_segment.py:289-313
 *                                       (same code: models/Point_MAE_pretask_dev.py:437-461)
 * on top of square_distance (models/modules.py:13-32, the EXPANDED form
 * -2 a.b + |a|^2 + |b|^2) and index_points (models/modules.py:35-51):
 *   dists = square_distance(xyz1, xyz2); sort ascending; keep k
 *   dist_recip = 1/(dists + eps); weight = dist_recip / sum(dist_recip)
 *   interpolated = sum_j index_points(points2, idx)[...,j,:] * weight[...,j]
 *   out = (base ? base : 0) + alpha * interpolated     (propagate: base=points1, alpha=0.3)
 * The sort is restated as a stable selection on (distance, index) -- equal
 * distances keep the lower source index (torch.sort makes no promise there).
 * Parity of THIS restatement against the reference's own Python: goldens made by
 * tests/golden/make_golden.py from the lifted functions (outputs to tolerance:
 * the matmul inside square_distance rounds differently from the fma chain here).
 */
static inline float sqdist_expanded(const float* a, const float* b) {
  const float s1 = (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2];
  const float s2 = (b[0] * b[0] + b[1] * b[1]) + b[2] * b[2];
  const float dot = fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
  return (-2.0f * dot + s1) + s2;
}

/* xyz1 (B,N,3), xyz2 (B,S,3), feat2 (B,S,C), base (B,N,C) or NULL ->
 * out (B,N,C), idx (B,N,k) int32, weight (B,N,k), dist (B,N,k).  Returns -1 if k > S. */
ORACLE_API int upp_oracle_interp_fwd(const float* xyz1, const float* xyz2, const float* feat2,
                                     const float* base, float alpha, float eps, int B, int N,
                                     int S, int C, int k, float* out, int32_t* idx, float* weight,
                                     float* dist) {
  if (k <= 0 || k > S) return -1;
  float* col = (float*)malloc(sizeof(float) * (size_t)S);
  unsigned char* used = (unsigned char*)malloc((size_t)S);
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n) {
      const size_t row = (size_t)b * N + n;
      const float* p = xyz1 + row * 3;
      for (int s = 0; s < S; ++s) { col[s] = sqdist_expanded(p, xyz2 + ((size_t)b * S + s) * 3); used[s] = 0; }
      float norm = 0.f;
      for (int j = 0; j < k; ++j) { /* stable selection: strict '<' keeps the lower index */
        int best = -1;
        for (int s = 0; s < S; ++s)
          if (!used[s] && (best < 0 || col[s] < col[best])) best = s;
        used[best] = 1;
        idx[row * k + j] = best;
        dist[row * k + j] = col[best];
        const float r = 1.0f / (col[best] + eps);
        weight[row * k + j] = r;
        norm = norm + r;
      }
      for (int j = 0; j < k; ++j) weight[row * k + j] = weight[row * k + j] / norm;
      for (int c = 0; c < C; ++c) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j)
          acc = acc + feat2[((size_t)b * S + idx[row * k + j]) * C + c] * weight[row * k + j];
        float o = alpha * acc;
        if (base) o = base[row * C + c] + o;
        out[row * C + c] = o;
      }
    }
  free(col); free(used);
  return 0;
}

/* Analytic gradients of the forward above (what autograd derives through the reference code):
 *   grad_feat2[b,s,:] = alpha * sum_{(n,j): idx=s} w * grad_out[b,n,:]
 *   with G = alpha*grad_out[b,n,:], dot_j = <G, f_j>, m = sum_j w_j dot_j, r_j = 1/(d_j+eps):
 *   g_dj = -(r_j w_j)(dot_j - m);  grad_xyz1[b,n] = sum_j g_dj 2(x1 - x2_j);  grad_xyz2[b,s] -= g_dj 2(x1 - x2_s)
 * Accumulated in double (this is the checker, not a bit-level restatement). */
ORACLE_API void upp_oracle_interp_bwd(const float* gout, const float* feat2, const float* xyz1,
                                      const float* xyz2, const int32_t* idx, const float* weight,
                                      const float* dist, float alpha, float eps, int B, int N, int S,
                                      int C, int k, float* gfeat2, float* gxyz1, float* gxyz2) {
  double* gf = (double*)calloc((size_t)B * S * C, sizeof(double));
  double* g2 = (double*)calloc((size_t)B * S * 3, sizeof(double));
  double* dot = (double*)malloc(sizeof(double) * (size_t)k);
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n) {
      const size_t row = (size_t)b * N + n;
      double m = 0.0;
      for (int j = 0; j < k; ++j) {
        const int s = idx[row * k + j];
        double d = 0.0;
        for (int c = 0; c < C; ++c) {
          const double g = (double)alpha * gout[row * C + c];
          d += g * feat2[((size_t)b * S + s) * C + c];
          gf[((size_t)b * S + s) * C + c] += g * weight[row * k + j];
        }
        dot[j] = d;
        m += (double)weight[row * k + j] * d;
      }
      double a[3] = {0, 0, 0};
      for (int j = 0; j < k; ++j) {
        const int s = idx[row * k + j];
        const double r = 1.0 / ((double)dist[row * k + j] + (double)eps);
        const double gd = -(r * weight[row * k + j]) * (dot[j] - m);
        for (int c = 0; c < 3; ++c) {
          const double v = gd * 2.0 * ((double)xyz1[row * 3 + c] - (double)xyz2[((size_t)b * S + s) * 3 + c]);
          a[c] += v;
          g2[((size_t)b * S + s) * 3 + c] -= v;
        }
      }
      for (int c = 0; c < 3; ++c) gxyz1[row * 3 + c] = (float)a[c];
    }
  for (size_t i = 0; i < (size_t)B * S * C; ++i) gfeat2[i] = (float)gf[i];
  for (size_t i = 0; i < (size_t)B * S * 3; ++i) gxyz2[i] = (float)g2[i];
  free(gf); free(g2); free(dot);
}

/* ------------------------------------------------------------------------------------------------
 * Viewpoint crop order of misc.seprate_point_cloud (reference utils/misc.py:232-233):
 *   distance_matrix = torch.norm(center - points, p=2, dim=-1);  idx = torch.argsort(distance_matrix)
 * restated as: d = sqrtf(fma(dz,dz,fma(dy,dy,dx*dx))) with d* = center - point (the sum of squares in x, y, z order,
 * then the square root), ascending, equal distances keep the lower point index first (stable).  order (B,n) int32.
 * Pinned by tests/golden/golden_seprate.npz (the reference's own function run unmodified). */
static int crop_key_cmp(const void* a, const void* b) {
  const unsigned long long x = *(const unsigned long long*)a, y = *(const unsigned long long*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

ORACLE_API void upp_oracle_crop_order(const float* xyz, const float* centers, int B, int n, int32_t* order) {
  unsigned long long* keys = (unsigned long long*)malloc((size_t)(n > 0 ? n : 1) * sizeof(unsigned long long));
  for (int b = 0; b < B; ++b) {
    const float* p = xyz + (size_t)b * n * 3;
    const float cx = centers[3 * b], cy = centers[3 * b + 1], cz = centers[3 * b + 2];
    for (int i = 0; i < n; ++i) {
      const float dx = cx - p[3 * i], dy = cy - p[3 * i + 1], dz = cz - p[3 * i + 2];
      const float d = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
      uint32_t bits;
      memcpy(&bits, &d, 4);
      keys[i] = ((unsigned long long)bits << 32) | (unsigned)i;
    }
    qsort(keys, (size_t)n, sizeof(unsigned long long), crop_key_cmp);
    for (int i = 0; i < n; ++i) order[(size_t)b * n + i] = (int32_t)(keys[i] & 0xffffffffull);
  }
  free(keys);
}
