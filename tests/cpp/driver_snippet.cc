// The call sequence of the reference's experiment driver (examples/teaser_cpp_ply/PSULVSB.cc:291-331)
// against include/teaser/registration.h: Params -> solver(params) -> solve(src_reduce, tgt_reduce) ->
// getSolution().  Synthetic input: N points, rotation about a fixed axis, translation, uniform noise,
// a fraction of gross outliers, an emulated keep_mask pre-filter (so self-update has work to do).
#include <teaser/registration.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

static unsigned long long lcg_state = 88172645463325252ull;
static double urand() {  // xorshift64*, [0, 1)
  lcg_state ^= lcg_state >> 12;
  lcg_state ^= lcg_state << 25;
  lcg_state ^= lcg_state >> 27;
  return (double)((lcg_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 1500;
  const double outlier_ratio = 0.8;
  Eigen::Matrix<double, 3, Eigen::Dynamic> src(3, N), tgt(3, N);
  const double ang = 0.7, c = std::cos(ang), s = std::sin(ang);
  const double R[3][3] = {{c, -s, 0}, {s, c, 0}, {0, 0, 1}};
  const double t[3] = {0.5, -1.0, 0.25};
  for (int i = 0; i < N; ++i) {
    double p[3] = {3 * urand() - 1.5, 3 * urand() - 1.5, 3 * urand() - 1.5};
    for (int r = 0; r < 3; ++r) {
      src(r, i) = p[r];
      tgt(r, i) = R[r][0] * p[0] + R[r][1] * p[1] + R[r][2] * p[2] + t[r] + 0.02 * urand() - 0.01;
    }
    if (urand() < outlier_ratio)
      for (int r = 0; r < 3; ++r) tgt(r, i) += (urand() < 0.5 ? -1 : 1) * (5 + 5 * urand());
  }
  // emulated histogram pre-filter (PSULVSB.cc:87-188): keep ~half of the correspondences
  std::vector<int> keep_mask(N, 0);
  std::map<int, int> reduce_map;
  int C = 0;
  for (int i = 0; i < N; ++i)
    if (urand() < 0.5) {
      keep_mask[i] = 1;
      reduce_map[i] = C++;
    }
  Eigen::Matrix<double, 3, Eigen::Dynamic> src_reduce(3, C), tgt_reduce(3, C);
  for (auto& kv : reduce_map)
    for (int r = 0; r < 3; ++r) {
      src_reduce(r, kv.second) = src(r, kv.first);
      tgt_reduce(r, kv.second) = tgt(r, kv.first);
    }

  teaser::RobustRegistrationSolver::Params params;
  params.noise_bound = 0.05;
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
  params.replay = true;

  teaser::RobustRegistrationSolver solver(params);
  solver.solve(src_reduce, tgt_reduce);
  auto solution = solver.getSolution();

  if (!solution.valid) {
    std::printf("valid=0 status=%d message=%s\n", solver.lastStatus(), psulvsb_last_error());
    return solver.lastStatus() == PSULVSB_ERR_NO_DEVICE ? 3 : 1;
  }
  double tr = 0;
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) tr += R[k][r] * solution.rotation(k, r);
  double cosv = (tr - 1) / 2;
  cosv = cosv > 1 ? 1 : (cosv < -1 ? -1 : cosv);
  const double rot_err = std::fabs(std::acos(cosv));
  double te = 0;
  for (int r = 0; r < 3; ++r) te += (solution.translation(r, 0) - t[r]) * (solution.translation(r, 0) - t[r]);
  std::printf("valid=1 inliers=%d rot_err=%.6f trans_err=%.6f C=%d final_C=%ld\n", solution.final_inlier_count, rot_err,
              std::sqrt(te), C, (long)src_reduce.cols());
  return (rot_err < 0.02 && std::sqrt(te) < 0.05) ? 0 : 2;
}
