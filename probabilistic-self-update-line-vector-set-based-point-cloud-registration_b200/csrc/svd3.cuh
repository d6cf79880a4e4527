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
  const double eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      const double alpha = A[0][p] * A[0][p] + A[1][p] * A[1][p] + A[2][p] * A[2][p];
      const double beta = A[0][q] * A[0][q] + A[1][q] * A[1][q] + A[2][q] * A[2][q];
      const double gamma = A[0][p] * A[0][q] + A[1][p] * A[1][q] + A[2][p] * A[2][q];
      if (fabs(gamma) > eps * sqrt(alpha * beta) && fabs(gamma) > 1e-300) {
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = rsqrt(1.0 + t * t);
        const double s = c * t;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double ap = A[k][p], aq = A[k][q];
          A[k][p] = c * ap - s * aq;
          A[k][q] = s * ap + c * aq;
          const double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq;
          V[k][q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  if (Vw) {
    // keep V orthonormal over many warm-started solves (one Gram-Schmidt pass; the drift per solve is O(eps))
    double n0 = rsqrt(V[0][0] * V[0][0] + V[1][0] * V[1][0] + V[2][0] * V[2][0]);
#pragma unroll
    for (int k = 0; k < 3; ++k) Vw[k][0] = V[k][0] * n0;
    double d01 = Vw[0][0] * V[0][1] + Vw[1][0] * V[1][1] + Vw[2][0] * V[2][1];
    double c1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c1[k] = V[k][1] - d01 * Vw[k][0];
    double n1 = rsqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) Vw[k][1] = c1[k] * n1;
    // third column: +- cross(v0, v1), sign of the current third column
    double cx = Vw[1][0] * Vw[2][1] - Vw[2][0] * Vw[1][1];
    double cy = Vw[2][0] * Vw[0][1] - Vw[0][0] * Vw[2][1];
    double cz = Vw[0][0] * Vw[1][1] - Vw[1][0] * Vw[0][1];
    const double sg = (cx * V[0][2] + cy * V[1][2] + cz * V[2][2]) < 0.0 ? -1.0 : 1.0;
    Vw[0][2] = sg * cx;
    Vw[1][2] = sg * cy;
    Vw[2][2] = sg * cz;
  }
  double sig[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  // order the triplets by descending sigma (indices only)
  int i0 = 0, i1 = 1, i2 = 2;
  if (sig[i0] < sig[i1]) { int t = i0; i0 = i1; i1 = t; }
  if (sig[i1] < sig[i2]) { int t = i1; i1 = i2; i2 = t; }
  if (sig[i0] < sig[i1]) { int t = i0; i0 = i1; i1 = t; }
  double U[3][3];  // columns: left singular vectors, in the order (i0, i1, i2)
  double Vs[3][3];
  const double tiny = 1e-14 * sig[i0];
  // first two columns
  {
    const double n0 = sig[i0];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      U[k][0] = A[k][i0] / n0;
      Vs[k][0] = V[k][i0];
      Vs[k][1] = V[k][i1];
      Vs[k][2] = V[k][i2];
    }
    if (sig[i1] > tiny) {
#pragma unroll
      for (int k = 0; k < 3; ++k) U[k][1] = A[k][i1] / sig[i1];
    } else {
      // rank 1: any unit vector orthogonal to u0 (the reference's completion is arbitrary too)
      double ax = fabs(U[0][0]), ay = fabs(U[1][0]), az = fabs(U[2][0]);
      double e[3] = {0, 0, 0};
      if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
      const double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
      double w[3] = {e[0] - d * U[0][0], e[1] - d * U[1][0], e[2] - d * U[2][0]};
      const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) U[k][1] = w[k] / nw;
    }
    if (sig[i2] > tiny) {
#pragma unroll
      for (int k = 0; k < 3; ++k) U[k][2] = A[k][i2] / sig[i2];
    } else {
      U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
      U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
      U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
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

}  // namespace psulvsb
