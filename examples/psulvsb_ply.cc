// psulvsb_ply.cc -- the reference's bunny experiment (examples/teaser_cpp_ply/PSULVSB.cc:224-514) against
// the B200 library: load a PLY, apply a random rigid transform, +-0.05 uniform noise and 90 % gross
// outliers (+-U[5,10] per axis), estimate normals (k = 20), then -- the timed part -- the normal-angle
// histogram pre-filter, the reduced set and RobustRegistrationSolver::solve; report the errors.
//
//   psulvsb_ply <file.ply> [trials = 5] [seed = 1] [outlier_rate = 0.9] [vertex_scale = 1]
//
// vertex_scale multiplies the vertices on load: the Stanford bunny is 0.15 m across, so the reference's
// +-0.05 noise is a third of the object and the rotation is only weakly determined; with vertex_scale 10
// the same noise is 3 % of the object.
//
// Differences from the reference driver, on purpose: a seeded generator instead of srand(time(NULL)) /
// random_device; normals from the library's GPU k-NN PCA instead of PCL; no sleep(3) between trials; the
// scalar N_OUTLIERS_RATE macro is used as a scalar (PSULVSB.cc:204 indexes it and does not compile).
#include <teaser/ply_io.h>
#include <teaser/registration.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

#include "psulvsb_io.h"

#define NOISE_BOUND 0.05

typedef Eigen::Matrix<double, 3, Eigen::Dynamic> Mat3X;

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s file.ply [trials] [seed] [outlier_rate]\n", argv[0]);
    return 64;
  }
  const int trials = argc > 2 ? std::atoi(argv[2]) : 5;
  const unsigned long long seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 1ull;
  const double outlier_rate = argc > 4 ? std::atof(argv[4]) : 0.9;
  const double vertex_scale = argc > 5 ? std::atof(argv[5]) : 1.0;

  teaser::PLYReader reader;
  teaser::PointCloud src_cloud;
  if (reader.read(argv[1], src_cloud) != 0) {
    std::fprintf(stderr, "cannot read %s: %s\n", argv[1], psulvsb_last_error());
    return 65;
  }
  const int N = static_cast<int>(src_cloud.size());
  Mat3X src(3, N);
  for (int i = 0; i < N; ++i) {
    src(0, i) = vertex_scale * src_cloud[static_cast<size_t>(i)].x;
    src(1, i) = vertex_scale * src_cloud[static_cast<size_t>(i)].y;
    src(2, i) = vertex_scale * src_cloud[static_cast<size_t>(i)].z;
  }
  std::printf("loaded %d vertices from %s\n", N, argv[1]);

  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> G(0.0, 1.0);
  double sum_re = 0, sum_te = 0, sum_ms = 0;
  int ok = 0;
  for (int trial = 0; trial < trials; ++trial) {
    // random axis-angle rotation (angle <= pi) and translation with |t| <= 3   (PSULVSB.cc:256-278)
    double ax[3] = {G(gen), G(gen), G(gen)};
    const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (double& v : ax) v /= an;
    const double ang = M_PI * U(gen), c = std::cos(ang), s = std::sin(ang), C1 = 1 - c;
    const double R[3][3] = {{c + ax[0] * ax[0] * C1, ax[0] * ax[1] * C1 - ax[2] * s, ax[0] * ax[2] * C1 + ax[1] * s},
                            {ax[1] * ax[0] * C1 + ax[2] * s, c + ax[1] * ax[1] * C1, ax[1] * ax[2] * C1 - ax[0] * s},
                            {ax[2] * ax[0] * C1 - ax[1] * s, ax[2] * ax[1] * C1 + ax[0] * s, c + ax[2] * ax[2] * C1}};
    double t[3] = {G(gen), G(gen), G(gen)};
    const double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), tl = 3.0 * U(gen);
    for (double& v : t) v *= tl / tn;

    // tgt = T src, noise, outliers   (PSULVSB.cc:281-286 -> :190-222)
    Mat3X tgt(3, N);
    for (int i = 0; i < N; ++i)
      for (int r = 0; r < 3; ++r)
        tgt(r, i) = R[r][0] * src(0, i) + R[r][1] * src(1, i) + R[r][2] * src(2, i) + t[r] + NOISE_BOUND * (2 * U(gen) - 1);
    std::vector<char> is_outlier(static_cast<size_t>(N), 0);
    const int outliers = static_cast<int>(N * outlier_rate);
    for (int k = 0; k < outliers;) {
      const int idx = static_cast<int>(U(gen) * N) % N;
      if (is_outlier[static_cast<size_t>(idx)]) continue;
      is_outlier[static_cast<size_t>(idx)] = 1;
      for (int r = 0; r < 3; ++r) tgt(r, idx) += (U(gen) <= 0.5 ? -1.0 : 1.0) * (5.0 + 5.0 * U(gen));
      ++k;
    }

    // normals (not timed in the reference either: PSULVSB.cc:307 precedes the timer)
    Mat3X src_normals(3, N), tgt_normals(3, N);
    if (psulvsb_estimate_normals_host(src.data(), N, 20, nullptr, src_normals.data()) != PSULVSB_OK ||
        psulvsb_estimate_normals_host(tgt.data(), N, 20, nullptr, tgt_normals.data()) != PSULVSB_OK) {
      std::fprintf(stderr, "normal estimation failed: %s\n", psulvsb_last_error());
      return 66;
    }

    const auto t0 = std::chrono::steady_clock::now();  // PSULVSB.cc:309
    std::vector<int> keep_mask(static_cast<size_t>(N), 0);
    int remain = 0;
    psulvsb_histogram_outlier_removal(src_normals.data(), tgt_normals.data(), N, keep_mask.data(), &remain);
    Mat3X src_reduce(3, N), tgt_reduce(3, N);
    std::vector<int> dense(static_cast<size_t>(N), -1);
    int Cn = 0;
    psulvsb_mask_filter(src.data(), tgt.data(), keep_mask.data(), N, src_reduce.data(), tgt_reduce.data(), dense.data(), &Cn);
    src_reduce.conservativeResize(3, Cn);
    tgt_reduce.conservativeResize(3, Cn);
    std::map<int, int> reduce_map;
    for (int i = 0; i < N; ++i)
      if (dense[static_cast<size_t>(i)] >= 0) reduce_map[i] = dense[static_cast<size_t>(i)];

    teaser::RobustRegistrationSolver::Params params;  // PSULVSB.cc:291-299, :319-324
    params.noise_bound = NOISE_BOUND;
    params.cbar2 = 1;
    params.estimate_scaling = false;
    params.rotation_max_iterations = 100;
    params.rotation_gnc_factor = 1.4;
    params.rotation_estimation_algorithm = teaser::RobustRegistrationSolver::ROTATION_ESTIMATION_ALGORITHM::GNC_TLS;
    params.rotation_cost_threshold = 0.005;
    params.ori_src = src;
    params.ori_dst = tgt;
    params.keep_mask = keep_mask;
    params.reduce_map = reduce_map;
    params.seed = seed * 1000 + static_cast<unsigned long long>(trial);
    teaser::RobustRegistrationSolver solver(params);
    solver.solve(src_reduce, tgt_reduce);
    const auto t1 = std::chrono::steady_clock::now();  // PSULVSB.cc:329
    auto solution = solver.getSolution();
    const double ms = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() / 1000.0;

    if (!solution.valid) {
      std::printf("trial %d: valid=0 status=%d (%s)\n", trial, solver.lastStatus(), psulvsb_last_error());
      if (solver.lastStatus() == PSULVSB_ERR_NO_DEVICE) return 3;
      continue;
    }
    double tr = 0, te = 0;
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) tr += R[k][r] * solution.rotation(k, r);
    const double re = std::fabs(std::acos(std::fmin(std::fmax((tr - 1) / 2, -1.0), 1.0))) * 180.0 / M_PI;
    for (int r = 0; r < 3; ++r) te += (solution.translation(r, 0) - t[r]) * (solution.translation(r, 0) - t[r]);
    te = std::sqrt(te);
    std::printf("trial %d: kept=%d/%d inliers=%d rot_err_deg=%.4f trans_err=%.4f time_ms=%.2f final_C=%ld\n", trial, Cn, N,
                solution.final_inlier_count, re, te, ms, static_cast<long>(src_reduce.cols()));
    sum_re += re;
    sum_te += te;
    sum_ms += ms;
    ++ok;
  }
  if (ok) std::printf("average over %d valid trials: rot_err_deg=%.4f trans_err=%.4f time_ms=%.2f\n", ok, sum_re / ok, sum_te / ok, sum_ms / ok);
  return ok == trials ? 0 : 1;
}
