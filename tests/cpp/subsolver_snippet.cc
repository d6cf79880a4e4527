// The reference's sub-solver call sequences (test/teaser/scale-solver-test.cc:71-130,
// rotation-solver-test.cc:137-251, translation-solver-test.cc:21-113, registration-test.cc:286-291 computeTIMs)
// against include/teaser/registration.h.  Prints one "name ok|FAIL" line per check; exit code = failures
// (3 = no CUDA device: the library has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <random>

#include "teaser/registration.h"

typedef Eigen::Matrix<double, 3, Eigen::Dynamic> Mat3X;
typedef Eigen::Matrix<bool, 1, Eigen::Dynamic> Mask;

static int failures = 0;
static void check(const char* name, bool ok) {
  std::printf("%s %s\n", name, ok ? "ok" : "FAIL");
  if (!ok) ++failures;
}

int main() {
  std::mt19937_64 gen(7);
  std::uniform_real_distribution<double> uni(-1.0, 1.0);
  const int N = 40;
  Mat3X src(3, N), dst(3, N);
  const double th = 1.2345;
  const double R[3][3] = {{std::cos(th), -std::sin(th), 0}, {std::sin(th), std::cos(th), 0}, {0, 0, 1}};
  const double t[3] = {0.5, -1.0, 2.0};
  for (int i = 0; i < N; ++i) {
    for (int r = 0; r < 3; ++r) src(r, i) = uni(gen);
    for (int r = 0; r < 3; ++r) dst(r, i) = R[r][0] * src(0, i) + R[r][1] * src(1, i) + R[r][2] * src(2, i) + t[r];
  }
  teaser::RobustRegistrationSolver::Params params;
  params.noise_bound = 0.01;
  params.cbar2 = 1;
  params.estimate_scaling = false;
  params.rotation_max_iterations = 100;
  params.rotation_gnc_factor = 1.4;
  params.rotation_cost_threshold = 1e-6;
  teaser::RobustRegistrationSolver solver(params);

  // computeTIMs (registration.cc:471-505)
  Eigen::Matrix<int, 2, Eigen::Dynamic> smap, dmap;
  Mat3X sv = solver.computeTIMs(src, &smap);
  Mat3X tv = solver.computeTIMs(dst, &dmap);
  if (solver.lastStatus() == PSULVSB_ERR_NO_DEVICE) {
    std::printf("no CUDA device: %s\n", psulvsb_last_error());
    return 3;
  }
  bool tims_ok = sv.cols() == N * (N - 1) / 2 && smap.cols() == sv.cols();
  long l = 0;
  for (int i = 0; i < N - 1 && tims_ok; ++i)
    for (int j = i + 1; j < N; ++j, ++l) {
      tims_ok = tims_ok && smap(0, l) == i && smap(1, l) == j;
      for (int r = 0; r < 3; ++r) tims_ok = tims_ok && sv(r, l) == src(r, j) - src(r, i);
    }
  check("computeTIMs", tims_ok);

  // ScaleInliersSelector (scale-solver-test.cc FixedScale): rigid motion keeps every length
  {
    teaser::ScaleInliersSelector sel(params.noise_bound, params.cbar2);
    double scale = 0;
    Mask inl;
    sel.solveForScale(sv, tv, &scale, &inl);
    bool all = inl.cols() == sv.cols() && scale == 1;
    for (long k = 0; k < inl.cols(); ++k) all = all && inl(0, k);
    check("ScaleInliersSelector.all_inliers", all);
    Mat3X big = tv;
    for (long k = 0; k < big.cols(); ++k)
      for (int r = 0; r < 3; ++r) big(r, k) = 3 * tv(r, k) + 10;
    sel.solveForScale(sv, big, &scale, &inl);
    bool none = true;
    for (long k = 0; k < inl.cols(); ++k) none = none && !inl(0, k);
    check("ScaleInliersSelector.no_inliers", none);
  }
  // solveForScale / Rotation / Translation of the solver object (registration.cc:1537-1560)
  check("solveForScale", solver.solveForScale(sv, tv) == 1.0 && solver.getScaleInliersMask().cols() == sv.cols());
  Eigen::Matrix3d Rg = solver.solveForRotation(sv, tv);
  double tr = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) tr += Rg(r, c) * R[r][c];
  check("solveForRotation", std::acos(std::fmin(1.0, std::fmax(-1.0, (tr - 1) / 2))) < 1e-5);
  Mat3X rotated(3, N);
  for (int i = 0; i < N; ++i)
    for (int r = 0; r < 3; ++r) rotated(r, i) = Rg(r, 0) * src(0, i) + Rg(r, 1) * src(1, i) + Rg(r, 2) * src(2, i);
  Eigen::Vector3d tg = solver.solveForTranslation(rotated, dst);
  check("solveForTranslation", std::fabs(tg(0, 0) - t[0]) < 1e-5 && std::fabs(tg(1, 0) - t[1]) < 1e-5 &&
                                   std::fabs(tg(2, 0) - t[2]) < 1e-5);
  bool masks = solver.getRotationInliersMask().cols() == sv.cols() && solver.getTranslationInliersMask().cols() == N;
  for (int i = 0; i < N && masks; ++i) masks = masks && solver.getTranslationInliersMask()(0, i);
  check("inlier_masks", masks);

  // GNCTLSRotationSolver with outliers + warm start (rotation-solver-test.cc)
  {
    teaser::GNCRotationSolver::Params gp = {100, 0.005, 1.4, 0.05};
    teaser::GNCTLSRotationSolver rot(gp);
    Mat3X tvo = tv;
    for (long k = 0; k < tvo.cols(); k += 3)
      for (int r = 0; r < 3; ++r) tvo(r, k) = 3 * uni(gen);
    Eigen::Matrix3d Ro;
    Mask inl;
    rot.solveForRotation(sv, tvo, &Ro, &inl);
    double tr2 = 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) tr2 += Ro(r, c) * R[r][c];
    long n_in = 0;
    for (long k = 0; k < inl.cols(); ++k) n_in += inl(0, k) ? 1 : 0;
    check("GNCTLSRotationSolver.outliers", std::acos(std::fmin(1.0, (tr2 - 1) / 2)) < 1e-3 && n_in >= sv.cols() / 2 &&
                                               n_in < sv.cols() && rot.getCostAtTermination() >= 0);
    rot.setLastBest(Ro);
    Eigen::Matrix3d Rw;
    rot.solveForRotation(sv, tvo, &Rw, &inl);
    double d = 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) d = std::fmax(d, std::fabs(Rw(r, c) - Ro(r, c)));
    check("GNCTLSRotationSolver.warm_start", d < 1e-3);
  }
  // TLSScaleSolver: dst line vectors 2.5 x longer
  {
    Mat3X tvs = tv;
    for (long k = 0; k < tvs.cols(); ++k)
      for (int r = 0; r < 3; ++r) tvs(r, k) = 2.5 * tv(r, k);
    teaser::TLSScaleSolver sc(params.noise_bound, params.cbar2, 3);
    double s = 0;
    Mask inl;
    sc.solveForScale(sv, tvs, &s, &inl);
    check("TLSScaleSolver", std::fabs(s - 2.5) < 1e-6 && inl.cols() == sv.cols());
  }
  return failures;
}
