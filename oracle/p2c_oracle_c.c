/* C restatement of the integer-valued point operators of the reference (TEST INFRASTRUCTURE ONLY).
 *
 * Why a second oracle: the torch oracle (p2c_oracle.py) follows the reference literally, including its (B,S,N)
 * distance tensor and full sort, which at the stress configuration (B=128, N=32768) is 17 GB and minutes of CPU
 * time.  These loops compute the same indices with O(N) memory so the kernels can be compared bit-exactly at
 * BASELINE.json's full sizes.  Parity status: PINNED - tests/test_oracle_c.py checks every function against the
 * reference goldens (tests/golden/pointops_*.npz) and against the torch oracle on random and tie-heavy inputs.
 *
 * Arithmetic follows what torch computes on float32:
 *   FPS  (models/pointnet_util.py:63-84)   d = (dx*dx + dy*dy) + dz*dz, no contraction; running = min; FIRST arg-max
 *   ball (models/pointnet_util.py:87-107)  d = ((-2 * dot) + |c|^2) + |p|^2 with dot = fma(c2,p2, fma(c1,p1, c0*p0))
 *                                          (what the K=3 batched matmul does), out-of-ball iff d > (float)(r*r)
 *   3-NN (models/pointnet_util.py:301-308) same expanded distance, three smallest (ties: lowest index first),
 *                                          w = 1/(d+1e-8) normalised
 * Compile with -ffp-contract=off so the compiler adds no FMAs of its own (oracle/c_oracle.py does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline float sq_norm(const float* p) { return (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]; }

static inline float expanded_dist(const float* c, float c2, const float* p, float p2) {
  float dot = c[0] * p[0];
  dot = fmaf(c[1], p[1], dot);
  dot = fmaf(c[2], p[2], dot);
  return ((dot * -2.0f) + c2) + p2;
}

/* xyz (B,N,3), start (B) -> idx (B,npoint) */
void p2c_oracle_fps(const float* xyz, const int64_t* start, int B, int N, int npoint, int64_t* idx) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float* P = xyz + (size_t)b * N * 3;
    float* running = (float*)malloc(sizeof(float) * (size_t)N);
    for (int n = 0; n < N; ++n) running[n] = 1e10f;
    int64_t far = start[b];
    for (int i = 0; i < npoint; ++i) {
      idx[(size_t)b * npoint + i] = far;
      const float cx = P[far * 3], cy = P[far * 3 + 1], cz = P[far * 3 + 2];
      float best = -INFINITY;
      int64_t arg = 0;
      for (int n = 0; n < N; ++n) {
        const float dx = P[n * 3] - cx, dy = P[n * 3 + 1] - cy, dz = P[n * 3 + 2] - cz;
        const float d = (dx * dx + dy * dy) + dz * dz;
        if (d < running[n]) running[n] = d;
        if (running[n] > best) { best = running[n]; arg = n; }      /* strict: first maximum */
      }
      far = arg;
    }
    free(running);
  }
}

/* xyz (B,N,3), new_xyz (B,S,3) -> idx (B,S,nsample); a ball without any hit yields N everywhere, like the reference */
void p2c_oracle_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq, int nsample,
                           int64_t* idx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int s = 0; s < S; ++s) {
      const float* P = xyz + (size_t)b * N * 3;
      const float* c = new_xyz + ((size_t)b * S + s) * 3;
      const float c2 = sq_norm(c);
      int64_t* out = idx + ((size_t)b * S + s) * nsample;
      int cnt = 0;
      for (int n = 0; n < N && cnt < nsample; ++n) {
        const float d = expanded_dist(c, c2, P + (size_t)n * 3, sq_norm(P + (size_t)n * 3));
        if (!(d > radius_sq)) out[cnt++] = n;
      }
      const int64_t pad = cnt > 0 ? out[0] : (int64_t)N;
      for (int j = cnt; j < nsample; ++j) out[j] = pad;
    }
  }
}

/* xyz1 (B,N,3) queries, xyz2 (B,S,3) sources, S >= 3 -> idx (B,N,3), weight (B,N,3), dist (B,N,3) */
void p2c_oracle_three_nn(const float* xyz1, const float* xyz2, int B, int N, int S, int64_t* idx, float* weight,
                         float* dist) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int n = 0; n < N; ++n) {
      const float* q = xyz1 + ((size_t)b * N + n) * 3;
      const float* Q = xyz2 + (size_t)b * S * 3;
      const float q2 = sq_norm(q);
      float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
      int64_t i0 = 0, i1 = 0, i2 = 0;
      for (int s = 0; s < S; ++s) {
        const float d = expanded_dist(q, q2, Q + (size_t)s * 3, sq_norm(Q + (size_t)s * 3));
        if (d < d0) { d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = s; }
        else if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = s; }
        else if (d < d2) { d2 = d; i2 = s; }
      }
      const size_t o = ((size_t)b * N + n) * 3;
      idx[o] = i0; idx[o + 1] = i1; idx[o + 2] = i2;
      dist[o] = d0; dist[o + 1] = d1; dist[o + 2] = d2;
      const float r0 = 1.0f / (d0 + 1e-8f), r1 = 1.0f / (d1 + 1e-8f), r2 = 1.0f / (d2 + 1e-8f);
      const float sum = (r0 + r1) + r2;
      weight[o] = r0 / sum; weight[o + 1] = r1 / sum; weight[o + 2] = r2 / sum;
    }
  }
}
