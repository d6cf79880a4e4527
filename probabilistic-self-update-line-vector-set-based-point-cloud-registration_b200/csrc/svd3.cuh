// svd3.cuh -- FP64 3x3 Kabsch rotation on the device (stands in for Eigen::JacobiSVD<Matrix3d> in
// teaser::utils::svdRot, utils.h:121-136, and in weightedSVD, registration.cc:550-558).
//
// One-sided (Hestenes) Jacobi: columns of A = H are rotated until mutually orthogonal while the
// same rotations accumulate in V, so that H = U diag(sigma) V^T with U = A / sigma column-wise.
// R = V U^T = sum_i v_i u_i^T does not depend on the ordering of the singular triplets; the
// reflection fix (utils.h:131-133: negate V.col(2), the smallest singular value after Eigen's
// descending sort) is applied to the triplet with the smallest sigma.
#pragma once

#include "common.cuh"

namespace psulvsb {

__device__ __forceinline__ double det3(const double a[3][3]) {
  return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
         a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

// ------------------------------------------------------------------------------------------------------------
// Two-sided Jacobi SVD of a 3x3 matrix in the operation order of Eigen::JacobiSVD<Matrix3d> as the CPU oracle
// restates it (oracle/psulvsb_oracle.cpp svd3; utils.h:127, registration.cc:550): sweeps over (p, q) =
// (1,0), (2,0), (2,1); each 2x2 block is first made symmetric by a left rotation, then diagonalised by a
// symmetric Jacobi rotation; threshold max(DBL_MIN, 2 eps max|diag|); singular values made non-negative by
// flipping U's columns, then sorted descending.  This file is compiled with -fmad=false, so every line below
// rounds exactly like the x86-64 build of the reference: for a rank-deficient H (a basic subset of ONE line
// vector gives H = w x y^T) the null-space completion -- which decides R = V U^T there and is determined by
// 1-ulp entries -- comes out identical to the oracle's, not merely "also valid".
// ------------------------------------------------------------------------------------------------------------
struct Svd3 {
  double U[3][3], S[3], V[3][3];
};

__device__ inline void svd3_two_sided(const double Ain[3][3], Svd3& o) {
  const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
  const double precision = 2.0 * eps;
  double scale = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double a = fabs(Ain[i][j]);
      scale = (scale < a) ? a : scale;
    }
  if (!(scale > 0.0) || !isfinite(scale)) scale = 1.0;
  double W[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      W[i][j] = Ain[i][j] / scale;
      o.U[i][j] = (i == j) ? 1.0 : 0.0;
      o.V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  auto mx = [](double a, double b) { return (a < b) ? b : a; };
  double max_diag = mx(fabs(W[0][0]), mx(fabs(W[1][1]), fabs(W[2][2])));
  // W <- rows (p, q) combined:  row_p' = c row_p + s row_q,  row_q' = -s row_p + c row_q
  auto rows = [](double M[3][3], int p, int q, double c, double s) {
    for (int k = 0; k < 3; ++k) {
      const double xp = M[p][k], xq = M[q][k];
      M[p][k] = c * xp + s * xq;
      M[q][k] = -s * xp + c * xq;
    }
  };
  // M <- columns (p, q) combined:  col_p' = c col_p - s col_q,  col_q' = s col_p + c col_q
  auto cols = [](double M[3][3], int p, int q, double c, double s) {
    for (int k = 0; k < 3; ++k) {
      const double xp = M[k][p], xq = M[k][q];
      M[k][p] = c * xp - s * xq;
      M[k][q] = s * xp + c * xq;
    }
  };
  bool finished = false;
  for (int guard = 0; !finished && guard < 200; ++guard) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        const double threshold = mx(tiny, precision * max_diag);
        if (fabs(W[p][q]) > threshold || fabs(W[q][p]) > threshold) {
          finished = false;
          const double a = W[p][p], b = W[p][q], c = W[q][p], d = W[q][q];
          // left rotation that makes the block symmetric
          const double t = a + d, dd = c - b;
          double r1c, r1s;
          if (fabs(dd) < tiny) {
            r1c = 1.0;
            r1s = 0.0;
          } else {
            const double u = t / dd;
            const double tmp = sqrt(1.0 + u * u);
            r1s = 1.0 / tmp;
            r1c = u / tmp;
          }
          const double a1 = r1c * a + r1s * c, b1 = r1c * b + r1s * d, d1 = -r1s * b + r1c * d;
          // symmetric 2x2 Jacobi rotation of [a1 b1; b1 d1]
          double jc, js;
          const double deno = 2.0 * fabs(b1);
          if (deno < tiny) {
            jc = 1.0;
            js = 0.0;
          } else {
            const double tau = (a1 - d1) / deno;
            const double w = sqrt(tau * tau + 1.0);
            const double tt = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
            const double sign_t = tt > 0 ? 1.0 : -1.0;
            const double n = 1.0 / sqrt(tt * tt + 1.0);
            js = -sign_t * (b1 / fabs(b1)) * fabs(tt) * n;
            jc = n;
          }
          const double lc = r1c * jc + r1s * js, ls = -r1c * js + r1s * jc;
          rows(W, p, q, lc, ls);
          cols(o.U, p, q, lc, -ls);
          cols(W, p, q, jc, js);
          cols(o.V, p, q, jc, js);
          max_diag = mx(max_diag, mx(fabs(W[p][p]), fabs(W[q][q])));
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    const double a = W[i][i];
    o.S[i] = fabs(a) * scale;
    if (a < 0)
      for (int k = 0; k < 3; ++k) o.U[k][i] = -o.U[k][i];
  }
  for (int i = 0; i < 3; ++i) {  // selection sort, descending; columns of U and V follow
    int best = i;
    for (int k = i + 1; k < 3; ++k)
      if (o.S[k] > o.S[best]) best = k;
    if (best != i) {
      double t = o.S[i];
      o.S[i] = o.S[best];
      o.S[best] = t;
      for (int k = 0; k < 3; ++k) {
        t = o.U[k][i];
        o.U[k][i] = o.U[k][best];
        o.U[k][best] = t;
        t = o.V[k][i];
        o.V[k][i] = o.V[k][best];
        o.V[k][best] = t;
      }
    }
  }
}

// svdRot's tail (utils.h:129-135): flip V.col(2) when det(U) det(V) < 0, R = V U^T with the oracle's association
__device__ inline void rotation_from_svd(Svd3& d, double R[3][3]) {
  if (det3(d.U) * det3(d.V) < 0)
    for (int k = 0; k < 3; ++k) d.V[k][2] = -d.V[k][2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = (d.V[i][0] * d.U[j][0] + d.V[i][1] * d.U[j][1]) + d.V[i][2] * d.U[j][2];
}

// (out of line: rare, and its arrays must not weigh on the register allocation of the callers' loops)
static __device__ __noinline__ void kabsch_rank_deficient(const double* Hin, double* R) {
  double H[3][3], Rm[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = Hin[i * 3 + j];
  Svd3 d;
  svd3_two_sided(H, d);
  rotation_from_svd(d, Rm);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rm[i][j];
}

// H: row-major 3x3 (H = sum w x y^T, rows index x).  Returns R = V U^T (row-major) where
// H = U S V^T.  det_mode 0: flip when det(U) det(V) < 0 (svdRot); 1: flip when det(V U^T) < 0
// (weightedSVD) -- the same condition, kept separate to mirror the two reference sites.
// Vw (optional, in/out): Jacobi warm start -- an orthogonal matrix close to the right singular
// vectors (the V of a nearby H); the rotations accumulate on top of it, so the converged result is
// the same SVD, reached in 1-2 sweeps instead of 4-5.
__device__ inline void kabsch_rotation(const double Hin[3][3], double R[3][3], double Vw[3][3] = nullptr) {
  double A[3][3], V[3][3];
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) scale = fmax(scale, fabs(Hin[i][j]));
  if (!(scale > 0.0) || !isfinite(scale)) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  const double inv = 1.0 / scale;
  if (Vw) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        V[i][j] = Vw[i][j];
        A[i][j] = ((Hin[i][0] * inv) * Vw[0][j] + (Hin[i][1] * inv) * Vw[1][j]) + (Hin[i][2] * inv) * Vw[2][j];
      }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        A[i][j] = Hin[i][j] * inv;
        V[i][j] = (i == j) ? 1.0 : 0.0;
      }
  }
  // FP64 latency, not throughput, is what the single thread running this pays, so the rotation is
  // derived with the shortest dependent chain: one rsqrt gives 1/h, from which t and c follow side
  // by side (no sqrt(alpha beta) in the convergence test, no rsqrt(1 + t^2) after t).
  const double eps2 = 4.930380657631324e-32;  // (2^-52)^2
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      const double alpha = fma(A[2][p], A[2][p], fma(A[1][p], A[1][p], A[0][p] * A[0][p]));
      const double beta = fma(A[2][q], A[2][q], fma(A[1][q], A[1][q], A[0][q] * A[0][q]));
      const double gamma = fma(A[2][p], A[2][q], fma(A[1][p], A[1][q], A[0][p] * A[0][q]));
      if (gamma * gamma > eps2 * (alpha * beta) && fabs(gamma) > 1e-300) {
        rotated = true;
        const double tau = beta - alpha, g = 2.0 * gamma;
        const double h2 = fma(tau, tau, g * g);
        const double ih = rsqrt(h2);
        const double h = h2 * ih;
        const double t = g / (tau + copysign(h, tau));       // smaller root of t^2 + 2 (tau / g) t - 1 = 0
        const double c = sqrt(fma(0.5 * fabs(tau), ih, 0.5));  // cos: c^2 = (1 + |tau| / h) / 2
        const double sn = c * t;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double ap = A[k][p], aq = A[k][q];
          A[k][p] = fma(c, ap, -sn * aq);
          A[k][q] = fma(sn, ap, c * aq);
          const double vp = V[k][p], vq = V[k][q];
          V[k][p] = fma(c, vp, -sn * vq);
          V[k][q] = fma(sn, vp, c * vq);
        }
      }
    }
    if (!rotated) break;
  }
  double a2[3], v2[3], ia[3], iv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    a2[j] = fma(A[2][j], A[2][j], fma(A[1][j], A[1][j], A[0][j] * A[0][j]));
    v2[j] = fma(V[2][j], V[2][j], fma(V[1][j], V[1][j], V[0][j] * V[0][j]));
    ia[j] = a2[j] > 0.0 ? rsqrt(a2[j]) : 0.0;
    iv[j] = rsqrt(v2[j]);
  }
  if (Vw) {
    // warm start for the next solve: the normalised V (orthonormal up to rounding; re-orthogonalised
    // by one Gram-Schmidt pass so that the drift cannot accumulate over a hundred warm-started solves)
#pragma unroll
    for (int k = 0; k < 3; ++k) Vw[k][0] = V[k][0] * iv[0];
    const double d01 = (Vw[0][0] * V[0][1] + Vw[1][0] * V[1][1] + Vw[2][0] * V[2][1]) * iv[1];
    double c1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c1[k] = fma(V[k][1], iv[1], -d01 * Vw[k][0]);
    const double n1 = rsqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) Vw[k][1] = c1[k] * n1;
    const double cx = Vw[1][0] * Vw[2][1] - Vw[2][0] * Vw[1][1];
    const double cy = Vw[2][0] * Vw[0][1] - Vw[0][0] * Vw[2][1];
    const double cz = Vw[0][0] * Vw[1][1] - Vw[1][0] * Vw[0][1];
    const double sg = (cx * V[0][2] + cy * V[1][2] + cz * V[2][2]) < 0.0 ? -1.0 : 1.0;
    Vw[0][2] = sg * cx;
    Vw[1][2] = sg * cy;
    Vw[2][2] = sg * cz;
  }
  // order the triplets by descending sigma_j = |A_j| / |V_j| (indices only, compared cross-multiplied)
  int i0 = 0, i1 = 1, i2 = 2;
  auto less = [&](int x, int y) { return a2[x] * v2[y] < a2[y] * v2[x]; };
  if (less(i0, i1)) { int t = i0; i0 = i1; i1 = t; }
  if (less(i1, i2)) { int t = i1; i1 = i2; i2 = t; }
  if (less(i0, i1)) { int t = i0; i0 = i1; i1 = t; }
  double U[3][3];  // columns: left singular vectors, in the order (i0, i1, i2)
  double Vs[3][3];
  // sigma_x > 1e-14 sigma_0  <=>  a2[x] v2[i0] > 1e-28 a2[i0] v2[x]
  const bool ok1 = a2[i1] * v2[i0] > 1e-28 * (a2[i0] * v2[i1]);
  const bool ok2 = a2[i2] * v2[i0] > 1e-28 * (a2[i0] * v2[i2]);
  if (!ok1 || !ok2) {
    // rank-deficient H: R = V U^T is not unique; take the completion of the reference's own algorithm
    kabsch_rank_deficient(&Hin[0][0], &R[0][0]);
    return;
  }
  {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      U[k][0] = A[k][i0] * ia[i0];
      Vs[k][0] = V[k][i0] * iv[i0];
      Vs[k][1] = V[k][i1] * iv[i1];
      Vs[k][2] = V[k][i2] * iv[i2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      U[k][1] = A[k][i1] * ia[i1];
      U[k][2] = A[k][i2] * ia[i2];
    }
  }
  if (det3(U) * det3(Vs) < 0.0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) Vs[k][2] = -Vs[k][2];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r][c] = Vs[r][0] * U[c][0] + Vs[r][1] * U[c][1] + Vs[r][2] * U[c][2];
}

// Warm-started alternative to the SVD: R = V U^T (with the determinant rule) is the maximiser of tr(R H) over SO(3)
// (Wahba / Kabsch), and inside the GNC loop the previous iteration's R is already close to it.  Newton's method on
// SO(3) from that start: with A = R H, the gradient of w -> tr(exp([w]x) A) at 0 is g = vee(A^T - A) and the negated
// Hessian is G = tr(A) I - sym(A); solve G w = g, apply the Cayley rotation of w, repeat.  The generic objective
// has a single local maximum (the other critical points are saddles or the minimum), so a converged iterate with G
// positive definite IS the SVD answer; it agrees with it to ~1e-15.  Returns false -- leaving R untouched -- when
// G is not safely positive definite (rank-deficient or nearly degenerate H, a start in the wrong basin) or the
// step does not shrink below 1e-7 rad within 5 iterations: the caller then runs the Jacobi SVD above.
// One division per step, ~1/5 of the SVD's dependent FP64 chain for the usual 2-3 steps.
__device__ inline bool rotation_newton(const double H[3][3], double R[3][3]) {
  double Rn[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rn[i][j] = R[i][j];
  for (int step = 0; step < 5; ++step) {
    double A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) A[i][j] = fma(Rn[i][2], H[2][j], fma(Rn[i][1], H[1][j], Rn[i][0] * H[0][j]));
    const double g0 = A[1][2] - A[2][1], g1 = A[2][0] - A[0][2], g2 = A[0][1] - A[1][0];
    const double G00 = A[1][1] + A[2][2], G11 = A[0][0] + A[2][2], G22 = A[0][0] + A[1][1];
    const double G01 = -0.5 * (A[0][1] + A[1][0]), G02 = -0.5 * (A[0][2] + A[2][0]), G12 = -0.5 * (A[1][2] + A[2][1]);
    const double c00 = fma(G11, G22, -G12 * G12), c01 = fma(G02, G12, -G01 * G22), c02 = fma(G01, G12, -G02 * G11);
    const double c11 = fma(G00, G22, -G02 * G02), c12 = fma(G01, G02, -G00 * G12), c22 = fma(G00, G11, -G01 * G01);
    const double det = fma(G00, c00, fma(G01, c01, G02 * c02));
    const double trg = G00 + G11 + G22;  // = 2 tr(A) = 2 (s1 + s2 +- s3) at the optimum
    // positive definite with margin: eigenvalues of G at the optimum are (s2 +- s3, s1 +- s3, s1 + s2)
    if (!(G00 > 0.0 && c22 > 0.0 && det > 1e-7 * (trg * trg * trg))) return false;
    const double u0 = 0.5 * fma(c00, g0, fma(c01, g1, c02 * g2));
    const double u1 = 0.5 * fma(c01, g0, fma(c11, g1, c12 * g2));
    const double u2 = 0.5 * fma(c02, g0, fma(c12, g1, c22 * g2));
    const double n = det * det, uu = fma(u0, u0, fma(u1, u1, u2 * u2));
    const double inv = 1.0 / (n + uu);
    const double d = (n - uu) * inv, k2 = 2.0 * inv, kd = k2 * det;
    double C[3][3];  // Cayley rotation of h = u / det: ((1 - |h|^2) I + 2 h h^T + 2 [h]x) / (1 + |h|^2)
    C[0][0] = fma(k2 * u0, u0, d);
    C[1][1] = fma(k2 * u1, u1, d);
    C[2][2] = fma(k2 * u2, u2, d);
    C[0][1] = fma(k2 * u0, u1, -kd * u2);
    C[1][0] = fma(k2 * u0, u1, kd * u2);
    C[0][2] = fma(k2 * u0, u2, kd * u1);
    C[2][0] = fma(k2 * u0, u2, -kd * u1);
    C[1][2] = fma(k2 * u1, u2, -kd * u0);
    C[2][1] = fma(k2 * u1, u2, kd * u0);
    double T[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) T[i][j] = fma(C[i][2], Rn[2][j], fma(C[i][1], Rn[1][j], C[i][0] * Rn[0][j]));
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rn[i][j] = T[i][j];
    if (uu < 2.5e-15 * n) {  // |w| = 2 |h| < 1e-7: the next step would move R by ~1e-14
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i][j] = Rn[i][j];
      return true;
    }
  }
  return false;
}

}  // namespace psulvsb
