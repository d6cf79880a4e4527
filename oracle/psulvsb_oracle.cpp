// psulvsb_oracle.cpp -- CPU parity oracle for the PSULVSB solve path.
//
// TEST INFRASTRUCTURE ONLY (see psulvsb_oracle.h).  A dependency-free restatement, in plain
// C++17, of what the reference computes on this path; it shares no code with the product.
// Citations: REG = /root/reference/teaser/src/registration.cc,
//            REGH = /root/reference/teaser/include/teaser/registration.h,
//            UTILS = /root/reference/teaser/include/teaser/utils.h.
// Third-party arithmetic the reference pulls in and this file restates:
//   Eigen3 (unpinned system package): JacobiSVD<Matrix3d> (UTILS:127, REG:550) -> svd3() below,
//     a two-sided Jacobi SVD; R = V U^T is unique for non-degenerate H so any accurate SVD agrees.
//   Boost.Math (unpinned): gamma_p(1.5, z) (REG:616) -> closed form erf(sqrt z) - 2 sqrt(z/pi) e^-z.
//   PMC (git, no tag; teaser/src/graph.cc:12-125) -> exact branch-and-bound max clique below;
//     parity for that branch is unpinned (maximum cliques are not unique).
//   libc rand()/srand(time) and std::random_device -> replaced by a keyed Philox4x32-10 stream so
//     that every draw can be replayed (reference behaviour is unreproducible by construction).
//
// Reference defects and the position taken here:
//   1. inlier masks are zeroed before being set (REG:1676-1691, REG:197-202 leave stale bits).
//   2. REG:1438 assignment-in-condition is reproduced literally: a non-inlier point draws u and
//      clears final_inliers[j] iff u > Q(residual_history[j]); inlier_history[j] := 0.
//   3. the out-of-bounds read at REG:169 has no effect and is not reproduced.
//   6. the first local iteration's sub-solvers use the caller's Params, later ones the in-loop
//      overrides (REG:937-945); reproduced.
//   7. tau = 2*score_noise_bound*(1 + (float)C/M) (REG:36,669,1424); reproduced incl. float.
//   8. refinement starts from *_best_sampled (REG:1508-1509); reproduced.
//  10. the early return at REG:1032-1036 resets first_time/longholi here (reference leaks them).
// abs() at REG:79,93,199,1262 is taken as the double overload (the evident intent).

#include "psulvsb_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// small 3x3 helpers (row-major storage: m[r][c])
// ---------------------------------------------------------------------------------------------
struct M3 {
  double m[3][3];
};
struct V3 {
  double v[3];
};

M3 m3_identity() {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
  return r;
}
M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
  return r;
}
M3 m3_transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
double m3_det(const M3& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
M3 m3_from_colmajor(const double* p) {
  M3 r;
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr) r.m[rr][c] = p[c * 3 + rr];
  return r;
}
void m3_to_colmajor(const M3& a, double* p) {
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr) p[c * 3 + rr] = a.m[rr][c];
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD: two-sided Jacobi (stands in for Eigen::JacobiSVD<Matrix3d>, UTILS:127 / REG:550).
// A = U diag(S) V^T, S sorted descending and non-negative.
// ---------------------------------------------------------------------------------------------
struct Rot2 {
  double c, s;
};  // [ c s; -s c ]

// Jacobi rotation J such that J^T [x y; y z] J is diagonal.
Rot2 sym_jacobi(double x, double y, double z) {
  Rot2 j;
  double deno = 2.0 * std::fabs(y);
  if (deno < std::numeric_limits<double>::min()) {
    j.c = 1.0;
    j.s = 0.0;
    return j;
  }
  double tau = (x - z) / deno;
  double w = std::sqrt(tau * tau + 1.0);
  double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sign_t = t > 0 ? 1.0 : -1.0;
  double n = 1.0 / std::sqrt(t * t + 1.0);
  j.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  j.c = n;
  return j;
}

// rows p,q of A: A <- J^T-applied-on-the-left, with J = [c s; -s c]:
//   row_p' = c*row_p - s*row_q ; row_q' = s*row_p + c*row_q      (this is "apply J^T... adjoint")
void rot_left(M3& a, int p, int q, double c, double s) {
  for (int k = 0; k < 3; ++k) {
    double xp = a.m[p][k], xq = a.m[q][k];
    a.m[p][k] = c * xp + s * xq;
    a.m[q][k] = -s * xp + c * xq;
  }
}
// columns p,q of A: A <- A * [c s; -s c]
void rot_right(M3& a, int p, int q, double c, double s) {
  for (int k = 0; k < 3; ++k) {
    double xp = a.m[k][p], xq = a.m[k][q];
    a.m[k][p] = c * xp - s * xq;
    a.m[k][q] = s * xp + c * xq;
  }
}

void svd3(const M3& Ain, M3& U, double S[3], M3& V) {
  const double eps = std::numeric_limits<double>::epsilon();
  const double precision = 2.0 * eps;
  const double tiny = std::numeric_limits<double>::min();
  double scale = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(Ain.m[i][j]));
  if (!(scale > 0.0) || !std::isfinite(scale)) scale = 1.0;
  M3 W;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) W.m[i][j] = Ain.m[i][j] / scale;
  U = m3_identity();
  V = m3_identity();
  double max_diag = std::max(std::fabs(W.m[0][0]), std::max(std::fabs(W.m[1][1]), std::fabs(W.m[2][2])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 200) {
    finished = true;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        double threshold = std::max(tiny, precision * max_diag);
        if (std::fabs(W.m[p][q]) > threshold || std::fabs(W.m[q][p]) > threshold) {
          finished = false;
          // 2x2 block [[a b],[c d]] on indices (p,q)
          double a = W.m[p][p], b = W.m[p][q], c = W.m[q][p], d = W.m[q][q];
          // step 1: rotation making the block symmetric
          double t = a + d, dd = c - b;
          double r1c, r1s;
          if (std::fabs(dd) < tiny) {
            r1c = 1.0;
            r1s = 0.0;
          } else {
            double u = t / dd;
            double tmp = std::sqrt(1.0 + u * u);
            r1s = 1.0 / tmp;
            r1c = u / tmp;
          }
          // apply rot1 on the left of the block: rows (p,q)
          double a1 = r1c * a + r1s * c, b1 = r1c * b + r1s * d;
          double d1 = -r1s * b + r1c * d;
          // step 2: diagonalise the symmetric block
          Rot2 jr = sym_jacobi(a1, b1, d1);
          // j_left = rot1 * jr^T ; rotations compose as complex-like pairs
          // rot(c1,s1)*rot(c2,s2) = rot(c1c2 - s1s2, c1s2 + s1c2); transpose flips sign of s
          double jlc = r1c * jr.c + r1s * jr.s;
          double jls = -r1c * jr.s + r1s * jr.c;
          // W <- j_left applied on the left (rows p,q), then jr on the right (cols p,q)
          rot_left(W, p, q, jlc, jls);
          rot_right(U, p, q, jlc, -jls);  // U <- U * j_left^T
          rot_right(W, p, q, jr.c, jr.s);
          rot_right(V, p, q, jr.c, jr.s);
          max_diag = std::max(max_diag, std::max(std::fabs(W.m[p][p]), std::fabs(W.m[q][q])));
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    double a = W.m[i][i];
    S[i] = std::fabs(a) * scale;
    if (a < 0)
      for (int k = 0; k < 3; ++k) U.m[k][i] = -U.m[k][i];
  }
  // sort descending (selection, swaps columns of U and V)
  for (int i = 0; i < 3; ++i) {
    int best = i;
    for (int k = i + 1; k < 3; ++k)
      if (S[k] > S[best]) best = k;
    if (best != i) {
      std::swap(S[i], S[best]);
      for (int k = 0; k < 3; ++k) {
        std::swap(U.m[k][i], U.m[k][best]);
        std::swap(V.m[k][i], V.m[k][best]);
      }
    }
  }
}

// UTILS:121-136.  X,Y: 3xK column-major, W: K weights.
M3 svd_rot(const double* X, const double* Y, const double* W, long long K) {
  M3 H;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) H.m[r][c] = 0.0;
  for (long long k = 0; k < K; ++k) {
    const double w = W[k];
    for (int r = 0; r < 3; ++r) {
      const double xw = X[3 * k + r] * w;
      for (int c = 0; c < 3; ++c) H.m[r][c] += xw * Y[3 * k + c];
    }
  }
  M3 U, V;
  double S[3];
  svd3(H, U, S, V);
  if (m3_det(U) * m3_det(V) < 0)
    for (int k = 0; k < 3; ++k) V.m[k][2] = -V.m[k][2];
  return m3_mul(V, m3_transpose(U));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the replayable sample stream
//   counter = (block_lo, block_hi, event, domain), key = (seed_lo, seed_hi)
// ---------------------------------------------------------------------------------------------
void philox4x32_10(uint64_t seed, uint32_t domain, uint32_t event, uint64_t block, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)block, c1 = (uint32_t)(block >> 32), c2 = event, c3 = domain;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

enum : uint32_t {
  DOMAIN_L_SAMPLED = 1,  // REG:852-861, event = host round
  DOMAIN_BASIC = 2,      // REG:916-932, event = global local-iteration index
  DOMAIN_UNIFORM = 3,    // REG:1428/1438 generateRandom01, event = host scoring index, k = point j
  DOMAIN_SCALE = 4,      // REG:90, event = scale-estimate call index
};

uint32_t rand31(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k) {
  uint32_t o[4];
  philox4x32_10(seed, domain, event, k >> 2, o);
  return o[k & 3] >> 1;  // 31 bits, like glibc rand()
}
double uniform01(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k) {
  uint32_t o[4];
  philox4x32_10(seed, domain, event, k, o);
  return ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) / 9007199254740992.0;
}

// REG:852-861 / REG:916-932: do { r = rand() % n; } while (used[r]);
long long sample_without_replacement(uint64_t seed, uint32_t domain, uint32_t event, long long n,
                                     long long count, int64_t* out) {
  std::vector<uint8_t> used((size_t)n, 0);
  uint64_t k = 0;
  for (long long i = 0; i < count; ++i) {
    long long r;
    do {
      r = (long long)(rand31(seed, domain, event, k++) % (uint64_t)n);
    } while (used[(size_t)r]);
    used[(size_t)r] = 1;
    out[i] = r;
  }
  return (long long)k;  // draws consumed
}

// ---------------------------------------------------------------------------------------------
// File-scope state of the reference (REG:36-50), made explicit
// ---------------------------------------------------------------------------------------------
struct State {
  int unknown_scale = 1;
  int first_time = 1;
  double scale_noise = 0;
  double translation_noise = 0;
  double scale_last_best = 1.0;
  M3 rotation_last_best = m3_identity();
  double translation_last_best[3] = {0, 0, 0};
  bool longholi = false;
  uint32_t scale_calls = 0;
};

inline double norm3(const double* v) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

// REG:418-434
inline bool length_consistent(const double* sv, const double* tv, double beta) {
  return std::fabs(norm3(sv) - norm3(tv)) <= beta;
}

// REG:66-120 (scale branch of ScalarTLSEstimator::estimate)
int tls_scale_estimate(const std::vector<double>& X, const std::vector<double>& ranges, State& st,
                       uint64_t seed, double* estimate, uint8_t* inliers) {
  const long long N = (long long)X.size();
  int best_inliers_count = 0;
  double confidence = 0;
  int iteration = 0;
  const uint32_t event = st.scale_calls++;
  uint64_t k = 0;
  if (!st.first_time) {
    iteration++;
    for (long long j = 0; j < N; ++j)
      if (std::fabs(X[j] - st.scale_last_best) <= ranges[j]) best_inliers_count++;
    *estimate = st.scale_last_best;
    confidence = 1.0 - std::pow(1.0 - ((double)best_inliers_count / (double)N), iteration);
  }
  while (confidence < 0.99) {
    iteration++;
    long long ran = (long long)(rand31(seed, DOMAIN_SCALE, event, k++) % (uint64_t)N);
    int curr_count = 0;
    for (long long j = 0; j < N; ++j)
      if (std::fabs(X[j] - X[ran]) <= ranges[j]) curr_count++;
    if (curr_count > best_inliers_count) {
      best_inliers_count = curr_count;
      *estimate = X[ran];
    }
    confidence = 1.0 - std::pow(1.0 - ((double)best_inliers_count / (double)N), iteration);
    if (iteration > 100000) break;  // guard: degenerate inputs (N == 0 / NaN) never converge
  }
  double sum_left = 0, sum_right = 0;
  for (long long i = 0; i < N; ++i) {
    bool in = std::fabs(X[i] - *estimate) <= ranges[i];
    if (inliers) inliers[i] = in ? 1 : 0;
    if (in) {
      sum_left += 1.0 / (ranges[i] * ranges[i]);
      sum_right += X[i] / (ranges[i] * ranges[i]);
    }
  }
  if (!std::isnan(sum_right) && !std::isnan(sum_left)) *estimate = sum_right / sum_left;
  return iteration;
}

// REG:121-203 (translation branch): max-stabbing of intervals x_k +- translation_noise.
// last_best_axis: pointer to the pseudo-measurement (REG:136-161) or nullptr when first_time.
void tls_translation_axis(const std::vector<double>& X, double noise, const double* last_best_axis,
                          double* estimate, uint8_t* inliers) {
  long long N = (long long)X.size();
  std::vector<std::pair<double, int>> h;
  h.reserve((size_t)(2 * (N + 1)));
  for (long long i = 0; i < N; ++i) {
    h.emplace_back(X[i] - noise, (int)i);
    h.emplace_back(X[i] + noise, (int)i);
  }
  double transAxis = 0;
  const bool pseudo = last_best_axis != nullptr;
  if (pseudo) {
    transAxis = *last_best_axis;
    h.emplace_back(transAxis - noise, (int)N);
    h.emplace_back(transAxis + noise, (int)N);
    N++;
  }
  // REG:162 uses std::sort (unstable); ties are resolved here by insertion order.
  std::stable_sort(h.begin(), h.end(),
                   [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });
  std::vector<int> record((size_t)N, 0);
  long long currLine = 0, bestLine = 0, remainingLine = N;
  double sum_left = 0, sum_right = 0;
  const double inv = 1.0 / (noise * noise);
  for (size_t i = 0; i < h.size(); ++i) {
    const int id = h[i].second;
    double x = (pseudo && id == N - 1) ? transAxis : X[(size_t)id];
    if (record[(size_t)id] == 0) {
      sum_left += inv;
      sum_right += x / (noise * noise);
      currLine++;
      remainingLine--;
      record[(size_t)id] = 1;
    } else {
      if (currLine > bestLine) {
        bestLine = currLine;
        if (!std::isnan(sum_right) && !std::isnan(sum_left))
          *estimate = sum_right / sum_left;
        else
          *estimate = x;
      }
      sum_left -= inv;
      sum_right -= x / (noise * noise);
      currLine--;
      record[(size_t)id] = 0;
      if (currLine + remainingLine <= bestLine) break;
    }
  }
  if (inliers) {
    for (size_t i = 0; i < X.size(); ++i) inliers[i] = (std::fabs(X[i] - *estimate) <= noise) ? 1 : 0;
  }
}

// REG:436-463.  src,dst 3xN column-major.
void tls_translation(const double* src, const double* dst, int N, double noise_bound, double cbar2,
                     State& st, double t_out[3], uint8_t* inliers) {
  st.translation_noise = noise_bound * std::sqrt(cbar2);
  std::vector<uint8_t> tmp((size_t)N, 1), acc((size_t)N, 1);
  std::vector<double> row((size_t)N);
  for (int axis = 0; axis < 3; ++axis) {
    for (int k = 0; k < N; ++k) row[(size_t)k] = dst[3 * k + axis] - src[3 * k + axis];
    const double* lb = st.first_time ? nullptr : &st.translation_last_best[axis];
    double est = t_out[axis];
    tls_translation_axis(row, st.translation_noise, lb, &est, tmp.data());
    t_out[axis] = est;
    for (int k = 0; k < N; ++k) acc[(size_t)k] = acc[(size_t)k] & tmp[(size_t)k];
  }
  if (inliers) std::memcpy(inliers, acc.data(), (size_t)N);
}

// REG:1563-1692.  sv,tv: 3xK column-major.  Returns iterations executed.
int gnc_tls(const double* sv, const double* tv, long long K, double noise_bound, int max_iterations,
            double gnc_factor, double cost_threshold, const M3* R_init, M3* R_out, uint8_t* inliers,
            double* cost_out) {
  double mu = 1;
  double prev_cost = std::numeric_limits<double>::infinity();
  double cost = std::numeric_limits<double>::infinity();
  double noise_bound_sq = noise_bound * noise_bound;
  if (noise_bound_sq < 1e-16) noise_bound_sq = 1e-2;
  std::vector<double> weights((size_t)K, 1.0), res((size_t)K, 0.0);
  bool use_init = R_init != nullptr;
  M3 R = m3_identity();
  int it_done = 0;
  for (int i = 0; i < max_iterations; ++i) {
    it_done = i + 1;
    if (use_init) {
      R = *R_init;
      use_init = false;
    } else {
      R = svd_rot(sv, tv, weights.data(), K);
    }
    for (long long k = 0; k < K; ++k) {
      const double* s = sv + 3 * k;
      const double* t = tv + 3 * k;
      double d0 = t[0] - ((R.m[0][0] * s[0] + R.m[0][1] * s[1]) + R.m[0][2] * s[2]);
      double d1 = t[1] - ((R.m[1][0] * s[0] + R.m[1][1] * s[1]) + R.m[1][2] * s[2]);
      double d2 = t[2] - ((R.m[2][0] * s[0] + R.m[2][1] * s[1]) + R.m[2][2] * s[2]);
      res[(size_t)k] = (d0 * d0 + d1 * d1) + d2 * d2;
    }
    if (i == 0) {
      double max_residual = K > 0 ? res[0] : 0.0;
      for (long long k = 1; k < K; ++k) max_residual = std::max(max_residual, res[(size_t)k]);
      mu = 1 / (2 * max_residual / noise_bound_sq - 1);
      if (mu <= 0) break;
    }
    double th1 = (mu + 1) / mu * noise_bound_sq;
    double th2 = mu / (mu + 1) * noise_bound_sq;
    cost = 0;
    for (long long k = 0; k < K; ++k) {
      cost += weights[(size_t)k] * res[(size_t)k];
      if (res[(size_t)k] >= th1)
        weights[(size_t)k] = 0;
      else if (res[(size_t)k] <= th2)
        weights[(size_t)k] = 1;
      else
        weights[(size_t)k] = std::sqrt(noise_bound_sq * mu * (mu + 1) / res[(size_t)k]) - mu;
    }
    double cost_diff = std::fabs(cost - prev_cost);
    mu = mu * gnc_factor;
    prev_cost = cost;
    if (cost_diff < cost_threshold) break;
  }
  if (inliers) {
    long long gf = 0;
    for (long long k = 0; k < K; ++k) {
      inliers[k] = weights[(size_t)k] >= 0.5 ? 1 : 0;
      gf += inliers[k];
    }
    if (gf <= 10)
      for (long long k = 0; k < K; ++k) inliers[k] = 1;
  }
  *R_out = R;
  if (cost_out) *cost_out = cost;
  return it_done;
}

// REG:611-619
double inlier_probability(double r, double sigma) {
  double z = (r * r) / (2.0 * sigma * sigma);
  if (!(z > 0)) return 1.0;
  double sq = std::sqrt(z);
  // 1 - P(3/2, z) = erfc(sqrt z) + 2 sqrt(z/pi) exp(-z)
  return std::erfc(sq) + 2.0 * std::sqrt(z / M_PI) * std::exp(-z);
}

// score of one transform over N points (REG:1303-1336, REG:1403-1424): res = | q - s (R p + t) |
inline double residual(const double* p, const double* q, double s, const M3& R, const double t[3]) {
  // (s * TRANSFORM) * [p;1]: entries of s*TRANSFORM are formed first (REG:1303,1329,1417)
  double d[3];
  for (int r = 0; r < 3; ++r) {
    double x = (((s * R.m[r][0]) * p[0] + (s * R.m[r][1]) * p[1]) + (s * R.m[r][2]) * p[2]) + (s * t[r]) * 1.0;
    d[r] = q[r] - x;
  }
  return std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
}

// REG:526-569.  T matrices are row-major 4x4 here.
struct M4 {
  double m[4][4];
};
M4 m4_identity() {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0 : 0.0;
  return r;
}
M4 m4_from_rt(const M3& R, const double t[3]) {
  M4 T = m4_identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.m[i][j] = R.m[i][j];
    T.m[i][3] = t[i];
  }
  return T;
}
M4 m4_mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}
inline void m4_apply(const M4& T, const double* p, double out[3]) {
  for (int r = 0; r < 3; ++r) out[r] = ((T.m[r][0] * p[0] + T.m[r][1] * p[1]) + T.m[r][2] * p[2]) + T.m[r][3];
}

M4 weighted_svd(const double* src, const double* tgt, const int* w, int M, const M4& init) {
  std::vector<double> ts((size_t)M * 3);
  double total = 0;
  double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
  for (int k = 0; k < M; ++k) {
    m4_apply(init, src + 3 * k, &ts[(size_t)3 * k]);
    double wk = (double)w[k];
    total += wk;
    for (int r = 0; r < 3; ++r) {
      cs[r] += ts[(size_t)3 * k + r] * wk;
      ct[r] += tgt[3 * k + r] * wk;
    }
  }
  for (int r = 0; r < 3; ++r) {
    cs[r] /= total;
    ct[r] /= total;
  }
  M3 cov;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov.m[r][c] = 0;
  for (int k = 0; k < M; ++k) {
    double wk = (double)w[k];
    for (int r = 0; r < 3; ++r) {
      double a = (ts[(size_t)3 * k + r] - cs[r]) * wk;
      for (int c = 0; c < 3; ++c) cov.m[r][c] += a * (tgt[3 * k + c] - ct[c]);
    }
  }
  M3 U, V;
  double S[3];
  svd3(cov, U, S, V);
  M3 R = m3_mul(V, m3_transpose(U));
  if (m3_det(R) < 0) {
    for (int k = 0; k < 3; ++k) V.m[k][2] = -V.m[k][2];
    R = m3_mul(V, m3_transpose(U));
  }
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = ct[r] - ((R.m[r][0] * cs[0] + R.m[r][1] * cs[1]) + R.m[r][2] * cs[2]);
  return m4_mul(m4_from_rt(R, t), init);
}

// REG:571-602.  Returns NaN when no inlier (the reference throws, REG:597-599).
double rmse(const double* src, const double* tgt, const int* mask, int M, const M4& T) {
  double sse = 0;
  int cnt = 0;
  for (int i = 0; i < M; ++i) {
    if (mask[i] == 1) {
      double x[3];
      m4_apply(T, src + 3 * i, x);
      double e0 = x[0] - tgt[3 * i], e1 = x[1] - tgt[3 * i + 1], e2 = x[2] - tgt[3 * i + 2];
      sse += (e0 * e0 + e1 * e1) + e2 * e2;
      cnt++;
    }
  }
  if (cnt == 0) return std::numeric_limits<double>::quiet_NaN();
  return std::sqrt(sse / cnt);
}

// ---------------------------------------------------------------------------------------------
// exact maximum clique (bitset branch and bound with greedy colouring bound)
// ---------------------------------------------------------------------------------------------
struct CliqueSolver {
  int n, words;
  std::vector<uint64_t> adj;  // n x words, vertices renumbered by degree order
  std::vector<int> best, cur;
  std::chrono::steady_clock::time_point t0;
  double time_limit = 3600;
  bool timed_out = false;
  const uint64_t* row(int v) const { return &adj[(size_t)v * words]; }
  void expand(std::vector<uint64_t>& P) {
    // greedy colouring of P -> order + bounds
    std::vector<int> order, bound;
    std::vector<uint64_t> U = P, Q((size_t)words);
    int colour = 0;
    auto any = [&](const std::vector<uint64_t>& s) {
      for (int w = 0; w < words; ++w)
        if (s[(size_t)w]) return true;
      return false;
    };
    while (any(U)) {
      colour++;
      Q = U;
      while (any(Q)) {
        int v = -1;
        for (int w = 0; w < words; ++w)
          if (Q[(size_t)w]) {
            v = w * 64 + __builtin_ctzll(Q[(size_t)w]);
            break;
          }
        Q[(size_t)(v >> 6)] &= ~(1ull << (v & 63));
        U[(size_t)(v >> 6)] &= ~(1ull << (v & 63));
        const uint64_t* a = row(v);
        for (int w = 0; w < words; ++w) Q[(size_t)w] &= ~a[w];
        order.push_back(v);
        bound.push_back(colour);
      }
    }
    for (int idx = (int)order.size() - 1; idx >= 0; --idx) {
      if ((int)cur.size() + bound[(size_t)idx] <= (int)best.size()) return;
      int v = order[(size_t)idx];
      cur.push_back(v);
      std::vector<uint64_t> NP((size_t)words);
      bool nonempty = false;
      const uint64_t* a = row(v);
      for (int w = 0; w < words; ++w) {
        NP[(size_t)w] = P[(size_t)w] & a[w];
        nonempty |= NP[(size_t)w] != 0;
      }
      if (nonempty) {
        expand(NP);
      } else if (cur.size() > best.size()) {
        best = cur;
      }
      cur.pop_back();
      P[(size_t)(v >> 6)] &= ~(1ull << (v & 63));
    }
  }
};

std::vector<int> max_clique(int n, const std::vector<std::pair<int, int>>& edges) {
  std::vector<int> deg((size_t)n, 0);
  for (auto& e : edges) {
    if (e.first == e.second) continue;
    deg[(size_t)e.first]++;
    deg[(size_t)e.second]++;
  }
  // renumber: ascending degree (stable), so high-degree vertices are branched on first (expand
  // walks the colour order from the back)
  std::vector<int> perm((size_t)n);
  for (int i = 0; i < n; ++i) perm[(size_t)i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return deg[(size_t)a] < deg[(size_t)b]; });
  std::vector<int> inv((size_t)n);
  for (int i = 0; i < n; ++i) inv[(size_t)perm[(size_t)i]] = i;
  CliqueSolver cs;
  cs.n = n;
  cs.words = (n + 63) / 64;
  cs.adj.assign((size_t)n * cs.words, 0);
  for (auto& e : edges) {
    if (e.first == e.second) continue;
    int a = inv[(size_t)e.first], b = inv[(size_t)e.second];
    cs.adj[(size_t)a * cs.words + (b >> 6)] |= 1ull << (b & 63);
    cs.adj[(size_t)b * cs.words + (a >> 6)] |= 1ull << (a & 63);
  }
  std::vector<uint64_t> P((size_t)cs.words, 0);
  for (int i = 0; i < n; ++i)
    if (deg[(size_t)perm[(size_t)i]] > 0) P[(size_t)(i >> 6)] |= 1ull << (i & 63);
  cs.expand(P);
  std::vector<int> out;
  for (int v : cs.best) out.push_back(perm[(size_t)v]);
  if (out.empty() && n > 0) out.push_back(0);
  std::sort(out.begin(), out.end());
  // WHICH maximum clique PMC returns is not pinned by anything in the reference (un-vendored, unpinned, threaded).
  // Position taken, the same in the product (csrc/k5_clique.cu): of all cliques of the maximum size omega, the one
  // whose ascending vertex list is lexicographically smallest.  Found by a depth-first search over the ORIGINAL
  // numbering, candidates ascending, cut only where omega cannot be reached.
  const int omega = (int)out.size();
  if (omega >= 2) {
    const int words = (n + 63) / 64;
    std::vector<uint64_t> A((size_t)n * words, 0);
    for (auto& e : edges) {
      if (e.first == e.second) continue;
      A[(size_t)e.first * words + (e.second >> 6)] |= 1ull << (e.second & 63);
      A[(size_t)e.second * words + (e.first >> 6)] |= 1ull << (e.first & 63);
    }
    std::vector<int> cur;
    std::function<bool(const std::vector<uint64_t>&)> dfs = [&](const std::vector<uint64_t>& cand) -> bool {
      if ((int)cur.size() == omega) return true;
      int cnt = 0;
      for (int w = 0; w < words; ++w) cnt += __builtin_popcountll(cand[(size_t)w]);
      if ((int)cur.size() + cnt < omega) return false;
      std::vector<uint64_t> rest = cand, next((size_t)words);
      for (int w = 0; w < words; ++w)
        while (rest[(size_t)w]) {
          const int v = w * 64 + __builtin_ctzll(rest[(size_t)w]);
          rest[(size_t)w] &= rest[(size_t)w] - 1;
          // later candidates adjacent to v (everything at or below v is gone from `rest`)
          for (int x = 0; x < words; ++x) next[(size_t)x] = rest[(size_t)x] & A[(size_t)v * words + x];
          cur.push_back(v);
          if (dfs(next)) return true;
          cur.pop_back();
        }
      return false;
    };
    std::vector<uint64_t> all((size_t)words, 0);
    for (int v = 0; v < n; ++v)
      if (deg[(size_t)v] >= omega - 1) all[(size_t)(v >> 6)] |= 1ull << (v & 63);
    if (dfs(all)) out = cur;
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// the solver (REG:622-1535)
// ---------------------------------------------------------------------------------------------
struct SubParams {  // what reset(params_) hands to the sub-solvers (REGH:747-783)
  double noise_bound, cbar2;
  int estimate_scaling;
  int rot_max_it;
  double rot_gnc, rot_cost;
};

int solve_impl(const oracle_params_t& P, const double* src_in, const double* dst_in, int C0,
               const double* ori_src, const double* ori_dst, int M, const int* keep_mask_in,
               const int* reduce_map_in, oracle_solution_t* out, oracle_trace_t* trace) {
  State st;
  st.unknown_scale = P.estimate_scaling;
  const double PrNoise = 2 * P.score_noise_bound;  // REG:36
  double scale_best_sampled = 1.0, scale_best_host = 1.0;
  M3 rotation_best_sampled = m3_identity(), rotation_best_host = m3_identity();
  double translation_best_sampled[3] = {0, 0, 0}, translation_best_host[3] = {0, 0, 0};
  // working correspondence set (grows under self-update, REG:800-806)
  std::vector<double> src(src_in, src_in + (size_t)3 * C0), dst(dst_in, dst_in + (size_t)3 * C0);
  int C = C0;
  const float adoptive_thr_multiplier = 1 + (((float)C0) / (long)M);  // REG:669
  const double tau = PrNoise * adoptive_thr_multiplier;

  std::vector<int> inlier_counter((size_t)M, 0);
  std::vector<int> keep_mask(keep_mask_in, keep_mask_in + M);
  std::vector<int> reduce_map(reduce_map_in, reduce_map_in + M);
  std::vector<int> new_corr((size_t)M, 0);
  std::vector<double> residual_history((size_t)M, 0);
  std::vector<int> inlier_history((size_t)M, -1);
  std::vector<int> final_inliers((size_t)M, 0);
  int new_corr_count = 0;

  // sub-solver parameters: caller's until the first in-loop reset+override (defect 6)
  SubParams sp{P.noise_bound, P.cbar2, P.estimate_scaling, P.rotation_max_iterations,
               P.rotation_gnc_factor, P.rotation_cost_threshold};
  const SubParams sp_inloop{P.inloop_noise_bound, P.inloop_cbar2, P.estimate_scaling,
                            P.inloop_max_iterations, P.inloop_gnc_factor, P.inloop_cost_threshold};

  // ---- L set and L reduced set (REG:682-767); line vectors are never materialised: the set
  // is kept as endpoint pairs (map(0,l), map(1,l)) in reference order, sv = s[b]-s[a].
  const long long L0 = (long long)C0 * (C0 - 1) / 2;
  std::vector<int> red_a, red_b;  // L_reduced_set as endpoint pairs
  if (P.estimate_scaling) {
    // ratio histogram REG:687-731
    long long MaxScale = 10000;
    const int binsize = 20;
    std::vector<int> Hcount((size_t)(MaxScale * binsize), 0);
    long long max_H_index = 0;
    int max_H_height = 0;
    auto bin_of = [&](double X) -> long long {
      if (X > (double)MaxScale) {
        MaxScale = (long long)std::ceil((double)MaxScale + X);
        Hcount.resize((size_t)(MaxScale * binsize), 0);
      }
      double Hsize = (double)Hcount.size();
      double f = std::floor((X - 0) / (double)MaxScale * Hsize);
      long long H_index;
      if (!std::isfinite(f))
        H_index = 0;
      else
        H_index = (long long)f;
      if (H_index == (long long)Hcount.size())
        H_index--;
      else if (H_index > (long long)Hcount.size() || H_index < 0)
        H_index = 0;
      return H_index;
    };
    std::vector<int> bins;  // bin of each line vector, in order (needed for the second pass)
    bins.reserve((size_t)L0);
    for (int i = 0; i < C0 - 1; ++i)
      for (int j = i + 1; j < C0; ++j) {
        double sv[3], tv[3];
        for (int r = 0; r < 3; ++r) {
          sv[r] = src[3 * j + r] - src[3 * i + r];
          tv[r] = dst[3 * j + r] - dst[3 * i + r];
        }
        double X = norm3(tv) / norm3(sv);
        long long b = bin_of(X);
        bins.push_back((int)b);
        int h = ++Hcount[(size_t)b];
        if (h > max_H_height) {
          max_H_height = h;
          max_H_index = b;
        }
      }
    // NOTE: a MaxScale growth mid-way changes H.size() and thus later bin indices exactly as in
    // the reference, because bin_of() evaluates the same expression in the same order.
    long long want[3] = {max_H_index, max_H_index != 0 ? max_H_index - 1 : -1,
                         max_H_index != (long long)Hcount.size() - 1 ? max_H_index + 1 : -1};
    for (int w = 0; w < 3; ++w) {
      if (want[w] < 0) continue;
      size_t l = 0;
      for (int i = 0; i < C0 - 1; ++i)
        for (int j = i + 1; j < C0; ++j, ++l)
          if (bins[l] == (int)want[w]) {
            red_a.push_back(i);
            red_b.push_back(j);
          }
    }
  } else {
    const double beta = 2 * P.noise_bound * std::sqrt(P.cbar2);  // REG:429 with caller's params
    for (int i = 0; i < C0 - 1; ++i)
      for (int j = i + 1; j < C0; ++j) {
        double sv[3], tv[3];
        for (int r = 0; r < 3; ++r) {
          sv[r] = src[3 * j + r] - src[3 * i + r];
          tv[r] = dst[3 * j + r] - dst[3 * i + r];
        }
        if (length_consistent(sv, tv, beta)) {
          red_a.push_back(i);
          red_b.push_back(j);
        }
      }
  }
  const long long n_reduced0 = (long long)red_a.size();
  if (n_reduced0 == 0) {
    // The reference spins forever here (p_local = NaN never exceeds 0.99, REG:1352/1399);
    // the oracle reports an invalid solution instead.
    std::memset(out, 0, sizeof(*out));
    out->n_line_vectors = L0;
    out->scale = 1.0;
    out->rotation[0] = out->rotation[4] = out->rotation[8] = 1.0;
    return 0;
  }

  int best_inliers_count_host = 0, host_r = 0;
  double pro_host = 0.0;
  bool pro_host_not_over = true;
  double L_sampled_rate = 0.1, b_sampled_rate = 0.3;  // REG:776-777
  auto begin = std::chrono::steady_clock::now();
  std::vector<int> inlier_map;
  int qr_round_bound_limit = P.host_round_limit;
  int host_round = 0, local_iter_global = 0, host_scorings = 0, escalations = 0;
  double solution_scale = 1.0;
  M3 solution_rotation = m3_identity();
  double solution_translation[3] = {0, 0, 0};
  bool aborted = false;

  while (pro_host_not_over && qr_round_bound_limit > 0 && !aborted) {
    qr_round_bound_limit--;
    // ---- self-update append (REG:786-832)
    if (new_corr_count != 0 && P.self_update) {
      const int ori_corr_count = C;
      src.resize((size_t)3 * (C + new_corr_count));
      dst.resize((size_t)3 * (C + new_corr_count));
      for (int i = 0; i < new_corr_count; ++i) {
        const int o = new_corr[(size_t)i];
        for (int r = 0; r < 3; ++r) {
          src[(size_t)3 * (ori_corr_count + i) + r] = ori_src[3 * o + r];
          dst[(size_t)3 * (ori_corr_count + i) + r] = ori_dst[3 * o + r];
        }
        for (size_t j = 0; j < inlier_map.size(); ++j) {
          red_a.push_back(ori_corr_count + i);  // map(0,L) = new, map(1,L) = inlier: sv = s[inl]-s[new]
          red_b.push_back(inlier_map[j]);
        }
        keep_mask[(size_t)o] = 1;
        reduce_map[(size_t)o] = ori_corr_count + i;
        inlier_map.push_back(ori_corr_count + i);
      }
      C += new_corr_count;
    }
    new_corr_count = 0;
    inlier_map.clear();
    int sampled_first_time = 1;

    // ---- L sampled set (REG:837-894)
    const long long n_red = (long long)red_a.size();
    long long L_sampled_set_size = (long long)std::floor((double)n_red * L_sampled_rate);
    std::vector<int64_t> L_sampled;
    if (L_sampled_set_size == 0) {
      L_sampled_set_size = n_red;
      L_sampled.resize((size_t)n_red);
      for (long long i = 0; i < n_red; ++i) L_sampled[(size_t)i] = i;
    } else {
      L_sampled.resize((size_t)L_sampled_set_size);
      sample_without_replacement(P.seed, DOMAIN_L_SAMPLED, (uint32_t)host_round, n_red, L_sampled_set_size,
                                 L_sampled.data());
    }
    // unique endpoints in first-appearance order (REG:870-894)
    std::vector<double> src_sampled, dst_sampled;
    {
      std::vector<uint8_t> dub((size_t)C, 0);
      for (long long i = 0; i < L_sampled_set_size; ++i) {
        const int e[2] = {red_a[(size_t)L_sampled[(size_t)i]], red_b[(size_t)L_sampled[(size_t)i]]};
        for (int q = 0; q < 2; ++q)
          if (!dub[(size_t)e[q]]) {
            dub[(size_t)e[q]] = 1;
            for (int r = 0; r < 3; ++r) {
              src_sampled.push_back(src[(size_t)3 * e[q] + r]);
              dst_sampled.push_back(dst[(size_t)3 * e[q] + r]);
            }
          }
      }
    }
    const int n_sampled_pts = (int)(src_sampled.size() / 3);

    int best_inliers_count_sampled = 0, local_r = 0;
    double pro_local = 0;
    bool pro_local_not_over = true;

    while (pro_local_not_over) {
      // ---- basic subset (REG:908-933)
      const int basic_choose = (int)((double)L_sampled_set_size * b_sampled_rate);
      std::vector<int64_t> basic((size_t)basic_choose);
      if (basic_choose > 0)
        sample_without_replacement(P.seed, DOMAIN_BASIC, (uint32_t)local_iter_global, L_sampled_set_size,
                                   basic_choose, basic.data());
      std::vector<double> bsv((size_t)3 * basic_choose), btv((size_t)3 * basic_choose);
      std::vector<int> bma((size_t)basic_choose), bmb((size_t)basic_choose);
      for (int i = 0; i < basic_choose; ++i) {
        const long long l = L_sampled[(size_t)basic[(size_t)i]];
        const int a = red_a[(size_t)l], b = red_b[(size_t)l];
        bma[(size_t)i] = a;
        bmb[(size_t)i] = b;
        for (int r = 0; r < 3; ++r) {
          bsv[(size_t)3 * i + r] = src[(size_t)3 * b + r] - src[(size_t)3 * a + r];
          btv[(size_t)3 * i + r] = dst[(size_t)3 * b + r] - dst[(size_t)3 * a + r];
        }
      }
      // reset(params_) then overrides (REG:937-945): THIS iteration's solvers see `sp`
      const SubParams cur = sp;
      sp = sp_inloop;

      // ---- scale (REG:958-991)
      std::vector<uint8_t> scale_mask((size_t)basic_choose, 1);
      std::vector<double> psv, ptv;  // pruned TIMs handed to the rotation solver
      std::vector<int> pma, pmb;
      if (cur.estimate_scaling) {
        std::vector<double> X((size_t)basic_choose), alphas((size_t)basic_choose);
        const double beta = 2 * cur.noise_bound * std::sqrt(cur.cbar2);
        st.scale_noise = beta;  // REG:411
        for (int i = 0; i < basic_choose; ++i) {
          double v1 = norm3(&bsv[(size_t)3 * i]), v2 = norm3(&btv[(size_t)3 * i]);
          X[(size_t)i] = v2 / v1;
          alphas[(size_t)i] = beta * (1.0 / v1);
        }
        if (basic_choose > 0)
          tls_scale_estimate(X, alphas, st, P.seed, &solution_scale, scale_mask.data());
        for (int i = 0; i < basic_choose; ++i)
          if (scale_mask[(size_t)i]) {
            for (int r = 0; r < 3; ++r) {
              psv.push_back(bsv[(size_t)3 * i + r]);
              ptv.push_back(btv[(size_t)3 * i + r]);
            }
            pma.push_back(bma[(size_t)i]);
            pmb.push_back(bmb[(size_t)i]);
          }
      } else {
        solution_scale = 1;
        const double beta = 2 * cur.noise_bound * std::sqrt(cur.cbar2);
        for (int i = 0; i < basic_choose; ++i)
          scale_mask[(size_t)i] = length_consistent(&bsv[(size_t)3 * i], &btv[(size_t)3 * i], beta) ? 1 : 0;
        psv = bsv;
        ptv = btv;
        pma = bma;
        pmb = bmb;
      }
      const long long Kp = (long long)pma.size();

      // ---- max clique escalation (REG:1000-1085)
      std::vector<int> clique_pts;
      bool use_clique_pts = false;
      if (b_sampled_rate == 1.0) {
        use_clique_pts = true;
        if (P.inlier_selection_mode != 3) {
          std::vector<std::pair<int, int>> edges;
          for (int i = 0; i < basic_choose; ++i)
            if (scale_mask[(size_t)i]) edges.emplace_back(bma[(size_t)i], bmb[(size_t)i]);
          clique_pts = max_clique(C, edges);
          if (clique_pts.size() <= 1) {
            aborted = true;  // REG:1032-1036
            break;
          }
        } else {
          for (int i = 0; i < C; ++i) clique_pts.push_back(i);
        }
      }

      // ---- rotation (REG:1102-1111)
      const double inv_scale = 1 / solution_scale;
      for (auto& v : ptv) v *= inv_scale;
      const double rot_noise = cur.noise_bound * (2 / solution_scale);
      std::vector<uint8_t> rot_inl((size_t)Kp, 0);
      M3 Rinit = st.rotation_last_best;
      M3 Rsol = m3_identity();
      double gnc_cost = 0;
      if (st.first_time == 1) st.rotation_last_best = m3_identity();  // REG:1606-1610
      int gnc_its = gnc_tls(psv.data(), ptv.data(), Kp, rot_noise, cur.rot_max_it, cur.rot_gnc, cur.rot_cost,
                            st.first_time ? nullptr : &Rinit, &Rsol, rot_inl.data(), &gnc_cost);
      solution_rotation = Rsol;

      // ---- unique endpoints of rotation inliers (REG:1114-1155)
      std::vector<double> rps, rpd;
      int n_rot_inl = 0;
      {
        std::vector<uint8_t> dub((size_t)C, 0);
        for (long long i = 0; i < Kp; ++i)
          if (rot_inl[(size_t)i]) {
            n_rot_inl++;
            const int e[2] = {pma[(size_t)i], pmb[(size_t)i]};
            for (int q = 0; q < 2; ++q)
              if (!dub[(size_t)e[q]]) {
                dub[(size_t)e[q]] = 1;
                for (int r = 0; r < 3; ++r) {
                  rps.push_back(src[(size_t)3 * e[q] + r]);
                  rpd.push_back(dst[(size_t)3 * e[q] + r]);
                }
              }
          }
      }
      if (use_clique_pts) {  // REG:1238-1244
        rps.clear();
        rpd.clear();
        for (int v : clique_pts)
          for (int r = 0; r < 3; ++r) {
            rps.push_back(src[(size_t)3 * v + r]);
            rpd.push_back(dst[(size_t)3 * v + r]);
          }
      }
      const int n_rot_pts = (int)(rps.size() / 3);

      // ---- translation (REG:1248-1250): v1 = (s*R) * P
      {
        std::vector<double> v1((size_t)3 * n_rot_pts);
        for (int k = 0; k < n_rot_pts; ++k)
          for (int r = 0; r < 3; ++r)
            v1[(size_t)3 * k + r] = ((solution_scale * Rsol.m[r][0]) * rps[(size_t)3 * k] +
                                     (solution_scale * Rsol.m[r][1]) * rps[(size_t)3 * k + 1]) +
                                    (solution_scale * Rsol.m[r][2]) * rps[(size_t)3 * k + 2];
        tls_translation(v1.data(), rpd.data(), n_rot_pts, cur.noise_bound, cur.cbar2, st, solution_translation,
                        nullptr);
        for (int r = 0; r < 3; ++r) solution_translation[r] /= solution_scale;
      }

      // ---- similarity test / local scoring (REG:1261-1397)
      bool similar = false;
      int curr_count = -1;
      if (!st.first_time) {
        M3 RtR = m3_mul(m3_transpose(st.rotation_last_best), solution_rotation);
        double tr = RtR.m[0][0] + RtR.m[1][1] + RtR.m[2][2];
        double ang = std::fabs(std::acos(std::fmin(std::fmax((tr - 1) / 2, -1.0), 1.0)));
        double dt[3] = {st.translation_last_best[0] - solution_translation[0],
                        st.translation_last_best[1] - solution_translation[1],
                        st.translation_last_best[2] - solution_translation[2]};
        similar = std::fabs(st.scale_last_best - solution_scale) <= st.scale_noise && ang <= P.rotation_similar &&
                  norm3(dt) <= st.translation_noise;
      }
      if (similar) {
        if (sampled_first_time)
          local_r += host_r + 1;
        else
          local_r++;
        pro_local = 1.0;
        scale_best_sampled = solution_scale;
        rotation_best_sampled = solution_rotation;
        for (int r = 0; r < 3; ++r) translation_best_sampled[r] = solution_translation[r];
      } else {
        local_r++;
        if (!st.first_time && b_sampled_rate < 1.0) {  // REG:1289-1315
          int cnt = 0;
          for (int j = 0; j < n_sampled_pts; ++j)
            if (residual(&src_sampled[(size_t)3 * j], &dst_sampled[(size_t)3 * j], st.scale_last_best,
                         st.rotation_last_best, st.translation_last_best) <= tau)
              cnt++;
          best_inliers_count_sampled = cnt;
          scale_best_sampled = st.scale_last_best;
          rotation_best_sampled = st.rotation_last_best;
          for (int r = 0; r < 3; ++r) translation_best_sampled[r] = st.translation_last_best[r];
        }
        curr_count = 0;
        for (int j = 0; j < n_sampled_pts; ++j)
          if (residual(&src_sampled[(size_t)3 * j], &dst_sampled[(size_t)3 * j], solution_scale, solution_rotation,
                       solution_translation) <= tau)
            curr_count++;
        if (curr_count > best_inliers_count_sampled || st.first_time) {
          scale_best_sampled = solution_scale;
          rotation_best_sampled = solution_rotation;
          for (int r = 0; r < 3; ++r) translation_best_sampled[r] = solution_translation[r];
          best_inliers_count_sampled = curr_count;
        }
        st.scale_last_best = scale_best_sampled;
        st.rotation_last_best = rotation_best_sampled;
        for (int r = 0; r < 3; ++r) st.translation_last_best[r] = translation_best_sampled[r];
        pro_local = 1.0 - std::pow(1.0 - (double)((double)best_inliers_count_sampled / (double)n_sampled_pts), local_r);
        st.first_time = 0;
        if ((local_r >= P.local_max_iter && pro_local <= 0.2) || b_sampled_rate == 1.0) {
          pro_local = 1.0;
          if (L_sampled_rate == 0.1 && b_sampled_rate == 0.3) {
            L_sampled_rate = 0.2;
            b_sampled_rate = 0.3;
            escalations++;
          } else if (L_sampled_rate == 0.2 && b_sampled_rate == 0.3) {
            L_sampled_rate = 0.5;
            b_sampled_rate = 0.3;
            escalations++;
          } else if (L_sampled_rate == 0.5 && b_sampled_rate == 0.3) {
            L_sampled_rate = 1.0;
            b_sampled_rate = 1.0;
            escalations++;
          }
        }
      }
      if (trace && trace->local && trace->local_n < trace->local_cap) {
        oracle_local_trace_t& T = trace->local[trace->local_n++];
        T.host_round = host_round;
        T.local_iter = local_iter_global;
        T.n_sampled_lines = (int)L_sampled_set_size;
        T.n_sampled_points = n_sampled_pts;
        T.basic_choose = basic_choose;
        T.gnc_iterations = gnc_its;
        T.rot_inliers = n_rot_inl;
        T.n_rot_points = n_rot_pts;
        T.similar = similar ? 1 : 0;
        T.curr_count = curr_count;
        T.best_count = best_inliers_count_sampled;
        T.local_r = local_r;
        T.p_local = pro_local;
        T.l_rate = L_sampled_rate;
        T.b_rate = b_sampled_rate;
        T.scale = solution_scale;
        m3_to_colmajor(solution_rotation, T.R);
        for (int r = 0; r < 3; ++r) T.t[r] = solution_translation[r];
      }
      local_iter_global++;

      // ---- host scoring + self-update decision (REG:1399-1488)
      if (pro_local > P.tpro_local) {
        host_r += local_r;
        int curr = 0;
        const uint32_t ev = (uint32_t)host_scorings++;
        for (int j = 0; j < M; ++j) {
          double res = residual(ori_src + 3 * j, ori_dst + 3 * j, scale_best_sampled, rotation_best_sampled,
                                translation_best_sampled);
          if (res <= tau) {
            curr++;
            inlier_counter[(size_t)j]++;
            bool add = false;
            if (keep_mask[(size_t)j] == 0) {
              const int hst = inlier_history[(size_t)j];
              if (hst == -1 || hst == 1)
                add = true;
              else if (hst == 0)
                add = uniform01(P.seed, DOMAIN_UNIFORM, ev, (uint64_t)j) <= inlier_probability(res, P.score_noise_bound);
            }
            if (add) {
              new_corr[(size_t)new_corr_count++] = j;
              final_inliers[(size_t)j] = 1;
            } else if (keep_mask[(size_t)j] == 1) {
              inlier_map.push_back(reduce_map[(size_t)j]);
              final_inliers[(size_t)j] = 1;
            }
            inlier_history[(size_t)j] = 1;
          } else {
            // REG:1438 (defect 2): the draw is always consumed; clears final_inliers when u > Q
            double u = uniform01(P.seed, DOMAIN_UNIFORM, ev, (uint64_t)j);
            if (u > inlier_probability(residual_history[(size_t)j], P.score_noise_bound)) final_inliers[(size_t)j] = 0;
            inlier_history[(size_t)j] = 0;
          }
          residual_history[(size_t)j] = res;
        }
        if (curr > best_inliers_count_host || pro_host == 0.0 ||
            (b_sampled_rate == 1.0 && curr >= best_inliers_count_host)) {
          scale_best_host = scale_best_sampled;
          rotation_best_host = rotation_best_sampled;
          for (int r = 0; r < 3; ++r) translation_best_host[r] = translation_best_sampled[r];
          best_inliers_count_host = curr;
        }
        st.scale_last_best = scale_best_host;
        st.rotation_last_best = rotation_best_host;
        for (int r = 0; r < 3; ++r) st.translation_last_best[r] = translation_best_host[r];
        pro_host = 1.0 - std::pow(1.0 - (double)((double)best_inliers_count_host / (double)M), host_r);
        double curr_time =
            std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - begin).count() /
            1000000.0;
        bool timeup = P.wallclock_cap_s > 0 && curr_time > P.wallclock_cap_s;
        if (pro_host > P.tpro_host || st.longholi || timeup) pro_host_not_over = false;
        pro_local_not_over = false;
        if (L_sampled_rate == 1.0 && b_sampled_rate == 1.0) st.longholi = true;
        if (trace && trace->host && trace->host_n < trace->host_cap) {
          oracle_host_trace_t& T = trace->host[trace->host_n++];
          T.host_round = host_round;
          T.curr_count = curr;
          T.best_host = best_inliers_count_host;
          T.new_corr_count = P.self_update ? new_corr_count : 0;
          T.inlier_map_size = (int)inlier_map.size();
          T.host_r = host_r;
          T.p_host = pro_host;
        }
      }
      sampled_first_time = 0;
    }
    host_round++;
  }

  std::memset(out, 0, sizeof(*out));
  out->host_rounds = host_round;
  out->local_iters = local_iter_global;
  out->n_line_vectors = L0;
  out->n_reduced = n_reduced0;
  out->final_C = C;
  out->escalations = escalations;
  if (aborted) {
    out->valid = 0;
    out->scale = solution_scale;
    m3_to_colmajor(solution_rotation, out->rotation);
    for (int r = 0; r < 3; ++r) out->translation[r] = solution_translation[r];
    return 0;
  }
  // ---- refinement (REG:1499-1525)
  solution_rotation = rotation_best_host;
  for (int r = 0; r < 3; ++r) solution_translation[r] = translation_best_host[r];
  if (best_inliers_count_host != 0) {
    M4 init = m4_from_rt(rotation_best_sampled, translation_best_sampled);
    M4 adj = weighted_svd(ori_src, ori_dst, inlier_counter.data(), M, init);
    double adj_rmse = rmse(ori_src, ori_dst, final_inliers.data(), M, adj);
    double ori_rmse = rmse(ori_src, ori_dst, final_inliers.data(), M, init);
    if (!std::isnan(adj_rmse) && !std::isnan(ori_rmse) && adj_rmse < ori_rmse) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) solution_rotation.m[r][c] = adj.m[r][c];
        solution_translation[r] = adj.m[r][3];
      }
      out->refined = 1;
    }
  }
  out->valid = 1;
  out->scale = scale_best_host;
  out->final_inlier_count = best_inliers_count_host;
  m3_to_colmajor(solution_rotation, out->rotation);
  for (int r = 0; r < 3; ++r) out->translation[r] = solution_translation[r];
  if (trace) {
    if (trace->final_inliers) std::memcpy(trace->final_inliers, final_inliers.data(), sizeof(int) * (size_t)M);
    if (trace->inlier_counter) std::memcpy(trace->inlier_counter, inlier_counter.data(), sizeof(int) * (size_t)M);
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C interface
// ---------------------------------------------------------------------------------------------
extern "C" {

void oracle_default_params(oracle_params_t* p) {
  p->noise_bound = 0.01;  // REGH:383
  p->cbar2 = 1;
  p->estimate_scaling = 1;
  p->rotation_max_iterations = 100;
  p->rotation_gnc_factor = 1.4;
  p->rotation_cost_threshold = 1e-6;
  p->inlier_selection_mode = 0;
  p->kcore_heuristic_threshold = 0.5;
  p->score_noise_bound = 0.01;
  p->inloop_noise_bound = 0.05;
  p->inloop_cbar2 = 1;
  p->inloop_max_iterations = 100;
  p->inloop_gnc_factor = 1.4;
  p->inloop_cost_threshold = 0.005;
  p->rotation_similar = 0.01;
  p->local_max_iter = 10;
  p->tpro_host = 0.99;
  p->tpro_local = 0.99;
  p->host_round_limit = 5;
  p->wallclock_cap_s = 60.0;
  p->self_update = 1;
  p->seed = 0;
}

int oracle_solve(const oracle_params_t* p, const double* src, const double* dst, int C, const double* ori_src,
                 const double* ori_dst, int M, const int* keep_mask, const int* reduce_map, oracle_solution_t* out,
                 oracle_trace_t* trace) {
  if (!p || !src || !dst || !ori_src || !ori_dst || !keep_mask || !reduce_map || !out || C < 2 || M < 1) return 1;
  return solve_impl(*p, src, dst, C, ori_src, ori_dst, M, keep_mask, reduce_map, out, trace);
}

void oracle_consistency_mask(const double* src, const double* dst, int n, double beta, uint8_t* mask,
                             double* margin) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (i == j) {
        mask[(size_t)i * n + j] = 0;
        if (margin) margin[(size_t)i * n + j] = 0;
        continue;
      }
      double sv[3], tv[3];
      // evaluated on the unordered pair (min,max) exactly as REG:697-698 does for i<j
      const int a = i < j ? i : j, b = i < j ? j : i;
      for (int r = 0; r < 3; ++r) {
        sv[r] = src[3 * b + r] - src[3 * a + r];
        tv[r] = dst[3 * b + r] - dst[3 * a + r];
      }
      double d = std::fabs(norm3(sv) - norm3(tv));
      mask[(size_t)i * n + j] = d <= beta ? 1 : 0;
      if (margin) margin[(size_t)i * n + j] = d - beta;
    }
}

void oracle_scale_inliers(const double* sv, const double* tv, long long K, double beta, uint8_t* mask) {
  for (long long k = 0; k < K; ++k) mask[k] = length_consistent(sv + 3 * k, tv + 3 * k, beta) ? 1 : 0;
}

long long oracle_reduced_set(const double* src, const double* dst, int n, double beta, int* pair_i, int* pair_j,
                             long long cap) {
  long long cnt = 0;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      double sv[3], tv[3];
      for (int r = 0; r < 3; ++r) {
        sv[r] = src[3 * j + r] - src[3 * i + r];
        tv[r] = dst[3 * j + r] - dst[3 * i + r];
      }
      if (length_consistent(sv, tv, beta)) {
        if (cnt < cap && pair_i && pair_j) {
          pair_i[cnt] = i;
          pair_j[cnt] = j;
        }
        cnt++;
      }
    }
  return cnt;
}

void oracle_svd_rot(const double* X, const double* Y, const double* W, long long K, double* R_colmajor) {
  M3 R = svd_rot(X, Y, W, K);
  m3_to_colmajor(R, R_colmajor);
}

void oracle_svd3(const double* A, double* U, double* S, double* V) {
  M3 a = m3_from_colmajor(A), u, v;
  svd3(a, u, S, v);
  m3_to_colmajor(u, U);
  m3_to_colmajor(v, V);
}

int oracle_gnc_tls(const double* sv, const double* tv, long long K, double noise_bound, int max_iterations,
                   double gnc_factor, double cost_threshold, const double* R_init, double* R_colmajor,
                   uint8_t* inliers, double* cost) {
  M3 init, R;
  if (R_init) init = m3_from_colmajor(R_init);
  int its = gnc_tls(sv, tv, K, noise_bound, max_iterations, gnc_factor, cost_threshold, R_init ? &init : nullptr, &R,
                    inliers, cost);
  m3_to_colmajor(R, R_colmajor);
  return its;
}

void oracle_tls_translation(const double* src, const double* dst, int N, double noise_bound, double cbar2,
                            const double* last_best, double* t_out, uint8_t* inliers) {
  State st;
  st.first_time = last_best ? 0 : 1;
  if (last_best)
    for (int r = 0; r < 3; ++r) st.translation_last_best[r] = last_best[r];
  double t[3] = {0, 0, 0};
  tls_translation(src, dst, N, noise_bound, cbar2, st, t, inliers);
  for (int r = 0; r < 3; ++r) t_out[r] = t[r];
}

int oracle_tls_scale(const double* sv, const double* tv, long long K, double noise_bound, double cbar2,
                     const double* last_best, uint64_t seed, uint32_t event, double* scale, uint8_t* inliers) {
  State st;
  st.first_time = last_best ? 0 : 1;
  if (last_best) st.scale_last_best = *last_best;
  st.scale_calls = event;
  std::vector<double> X((size_t)K), alphas((size_t)K);
  const double beta = 2 * noise_bound * std::sqrt(cbar2);
  for (long long i = 0; i < K; ++i) {
    double v1 = norm3(sv + 3 * i), v2 = norm3(tv + 3 * i);
    X[(size_t)i] = v2 / v1;
    alphas[(size_t)i] = beta * (1.0 / v1);
  }
  return tls_scale_estimate(X, alphas, st, seed, scale, inliers);
}

int oracle_score(const double* Pp, const double* Q, int N, double scale, const double* R_colmajor, const double* t,
                 double tau, uint8_t* inliers, double* residuals) {
  M3 R = m3_from_colmajor(R_colmajor);
  int cnt = 0;
  for (int j = 0; j < N; ++j) {
    double res = residual(Pp + 3 * j, Q + 3 * j, scale, R, t);
    bool in = res <= tau;
    if (inliers) inliers[j] = in ? 1 : 0;
    if (residuals) residuals[j] = res;
    cnt += in ? 1 : 0;
  }
  return cnt;
}

static M4 m4_from_colmajor(const double* p) {
  M4 T;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) T.m[r][c] = p[c * 4 + r];
  return T;
}
static void m4_to_colmajor(const M4& T, double* p) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) p[c * 4 + r] = T.m[r][c];
}

void oracle_weighted_svd(const double* src, const double* tgt, const int* w, int M, const double* T_init,
                         double* T_out) {
  M4 out = weighted_svd(src, tgt, w, M, m4_from_colmajor(T_init));
  m4_to_colmajor(out, T_out);
}

double oracle_rmse(const double* src, const double* tgt, const int* mask, int M, const double* T) {
  return rmse(src, tgt, mask, M, m4_from_colmajor(T));
}

double oracle_inlier_probability(double r, double sigma) { return inlier_probability(r, sigma); }

void oracle_philox4x32(uint64_t seed, uint32_t domain, uint32_t event, uint64_t block, uint32_t out[4]) {
  philox4x32_10(seed, domain, event, block, out);
}
uint32_t oracle_rand31(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k) {
  return rand31(seed, domain, event, k);
}
double oracle_uniform01(uint64_t seed, uint32_t domain, uint32_t event, uint64_t k) {
  return uniform01(seed, domain, event, k);
}
long long oracle_sample_without_replacement(uint64_t seed, uint32_t domain, uint32_t event, long long n,
                                            long long count, int64_t* out) {
  return sample_without_replacement(seed, domain, event, n, count, out);
}

int oracle_max_clique(int n_vertices, const int* edge_u, const int* edge_v, long long n_edges, int* clique_out) {
  std::vector<std::pair<int, int>> e;
  e.reserve((size_t)n_edges);
  for (long long i = 0; i < n_edges; ++i) e.emplace_back(edge_u[i], edge_v[i]);
  std::vector<int> c = max_clique(n_vertices, e);
  for (size_t i = 0; i < c.size(); ++i) clique_out[i] = c[i];
  return (int)c.size();
}

}  // extern "C"
