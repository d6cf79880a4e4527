// teaser/registration.h -- source-compatible facade of teaser::RobustRegistrationSolver for the PSULVSB
// path, implemented inline on the C ABI of libpsulvsb_b200.so (include/psulvsb.h).
//
// Same names, fields, enums and call sequence as the reference header
// (/root/reference/teaser/include/teaser/registration.h): RegistrationSolution (:34-41), Params incl.
// the PSULVSB additions ori_src / ori_dst / keep_mask / reduce_map (:378-473), the constructor (:484),
// solve(PointCloud, PointCloud, correspondences) (:503-505), solve(Eigen 3xN&, Eigen 3xN&) (:512-513),
// getSolution() (:553), getParams() (:548), reset() (:747).  The reference's drivers
// (examples/teaser_cpp_ply/PSULVSB.cc:291-331) compile against it unchanged; see INTEGRATION.md.
//
// Not carried over (outside the PSULVSB path): the abstract sub-solver classes,
// set*Estimator, the TIM / mask getters, computeTIMs, the certifier.
#pragma once

#include <Eigen/Core>

#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "../psulvsb.h"
#include "geometry.h"

namespace teaser {

struct RegistrationSolution {
  bool valid = true;
  double scale;
  int final_inlier_count;
  Eigen::Vector3d translation;
  Eigen::Matrix3d rotation;
};

class RobustRegistrationSolver {
public:
  enum class ROTATION_ESTIMATION_ALGORITHM { GNC_TLS = 0, FGR = 1 };
  enum class INLIER_GRAPH_FORMULATION { CHAIN = 0, COMPLETE = 1 };
  enum class INLIER_SELECTION_MODE { PMC_EXACT = 0, PMC_HEU = 1, KCORE_HEU = 2, NONE = 3 };

  struct Params {
    double noise_bound = 0.01;
    double cbar2 = 1;
    bool estimate_scaling = true;
    ROTATION_ESTIMATION_ALGORITHM rotation_estimation_algorithm = ROTATION_ESTIMATION_ALGORITHM::GNC_TLS;
    double rotation_gnc_factor = 1.4;
    size_t rotation_max_iterations = 100;
    double rotation_cost_threshold = 1e-6;
    INLIER_GRAPH_FORMULATION rotation_tim_graph = INLIER_GRAPH_FORMULATION::CHAIN;
    INLIER_SELECTION_MODE inlier_selection_mode = INLIER_SELECTION_MODE::PMC_EXACT;
    double kcore_heuristic_threshold = 0.5;
    bool use_max_clique = true;
    bool max_clique_exact_solution = true;
    double max_clique_time_limit = 3600;
    // PSULVSB additions (registration.h:469-472)
    Eigen::Matrix<double, 3, Eigen::Dynamic> ori_src;
    Eigen::Matrix<double, 3, Eigen::Dynamic> ori_dst;
    std::vector<int> keep_mask;
    std::map<int, int> reduce_map;
    // not in the reference (which seeds rand() with time(NULL) and stops after 60 s of wall clock):
    uint64_t seed = 0;    // key of the replayable Philox sample stream
    bool replay = false;  // true: disable the wall-clock stop rule (registration.cc:1475)
    int device = 0;       // CUDA device of this solver
  };

  RobustRegistrationSolver() = default;
  explicit RobustRegistrationSolver(const Params& params) { reset(params); }
  RobustRegistrationSolver(const RobustRegistrationSolver&) = delete;
  RobustRegistrationSolver& operator=(const RobustRegistrationSolver&) = delete;
  ~RobustRegistrationSolver() {
    if (handle_) psulvsb_destroy(handle_);
  }

  void reset(const Params& params) {
    params_ = params;
    solution_.valid = true;
    solution_.scale = 1;
    solution_.final_inlier_count = 0;
    solution_.translation.setZero();
    solution_.rotation.setIdentity();
  }
  Params getParams() { return params_; }
  RegistrationSolution getSolution() { return solution_; }
  // status / message of the last solve (PSULVSB_OK = 0); diagnostics of the run
  int lastStatus() const { return last_status_; }
  const psulvsb_solution_t& diagnostics() const { return raw_; }

  /// registration.cc:511-524: gathers the corresponded points into 3xN matrices and calls solve().
  RegistrationSolution solve(const teaser::PointCloud& src_cloud, const teaser::PointCloud& dst_cloud,
                             const std::vector<std::pair<int, int>> correspondences) {
    Eigen::Matrix<double, 3, Eigen::Dynamic> src, dst;
    src.resize(3, static_cast<long>(correspondences.size()));
    dst.resize(3, static_cast<long>(correspondences.size()));
    for (size_t i = 0; i < correspondences.size(); ++i) {
      const auto& s = src_cloud[static_cast<size_t>(correspondences[i].first)];
      const auto& d = dst_cloud[static_cast<size_t>(correspondences[i].second)];
      src(0, static_cast<long>(i)) = s.x;
      src(1, static_cast<long>(i)) = s.y;
      src(2, static_cast<long>(i)) = s.z;
      dst(0, static_cast<long>(i)) = d.x;
      dst(1, static_cast<long>(i)) = d.y;
      dst(2, static_cast<long>(i)) = d.z;
    }
    return solve(src, dst);
  }

  /// registration.cc:622-1535.  src / dst: the (reduced) correspondences as 3xC matrices.  Like the
  /// reference, self-update appends the correspondences it adopts to src / dst (registration.cc:800-806).
  RegistrationSolution solve(Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                             Eigen::Matrix<double, 3, Eigen::Dynamic>& dst) {
    solution_.valid = false;
    if (!handle_) {
      last_status_ = psulvsb_create(&handle_, params_.device);
      if (last_status_ != PSULVSB_OK) return solution_;
    }
    psulvsb_params_t p;
    psulvsb_default_params(&p);
    p.noise_bound = params_.noise_bound;
    p.cbar2 = params_.cbar2;
    p.estimate_scaling = params_.estimate_scaling ? 1 : 0;
    p.rotation_max_iterations = static_cast<int>(params_.rotation_max_iterations);
    p.rotation_gnc_factor = params_.rotation_gnc_factor;
    p.rotation_cost_threshold = params_.rotation_cost_threshold;
    p.inlier_selection_mode = static_cast<int>(params_.inlier_selection_mode);
    p.kcore_heuristic_threshold = params_.kcore_heuristic_threshold;
    p.seed = params_.seed;
    if (params_.replay) p.wallclock_cap_s = 0.0;
    const int C = static_cast<int>(src.cols());
    // ori_src / ori_dst absent (upstream-style callers): the reduced set is the whole set
    const bool have_ori = params_.ori_src.cols() > 0;
    const int M = have_ori ? static_cast<int>(params_.ori_src.cols()) : C;
    std::vector<int> keep(static_cast<size_t>(M), 1), dense(static_cast<size_t>(M), -1);
    if (have_ori && static_cast<int>(params_.keep_mask.size()) == M) {
      keep = params_.keep_mask;
      for (const auto& kv : params_.reduce_map)
        if (kv.first >= 0 && kv.first < M) dense[static_cast<size_t>(kv.first)] = kv.second;
    } else {
      for (int j = 0; j < M; ++j) dense[static_cast<size_t>(j)] = j;
    }
    psulvsb_problem_t prob;
    prob.src = src.data();
    prob.dst = dst.data();
    prob.C = C;
    prob.ori_src = have_ori ? params_.ori_src.data() : src.data();
    prob.ori_dst = have_ori ? params_.ori_dst.data() : dst.data();
    prob.M = M;
    prob.keep_mask = keep.data();
    prob.reduce_map = dense.data();
    std::vector<int> final_inliers(static_cast<size_t>(M), 0);
    psulvsb_trace_t trace = {};
    trace.final_inliers = final_inliers.data();
    last_status_ = psulvsb_solve(handle_, &p, &prob, &raw_, &trace);
    if (last_status_ != PSULVSB_OK) return solution_;
    if (raw_.status != PSULVSB_OK) {
      last_status_ = raw_.status;
      return solution_;
    }
    solution_.valid = raw_.valid != 0;
    solution_.scale = raw_.scale;
    solution_.final_inlier_count = raw_.final_inlier_count;
    for (int r = 0; r < 3; ++r) {
      solution_.translation(r, 0) = raw_.translation[r];
      for (int c = 0; c < 3; ++c) solution_.rotation(r, c) = raw_.rotation[c * 3 + r];
    }
    final_inliers_ = final_inliers;
    if (raw_.final_C > C) {  // the working set grew: keep the caller's matrices in step with the reference
      src.conservativeResize(3, raw_.final_C);
      dst.conservativeResize(3, raw_.final_C);
    }
    return solution_;
  }

  /// 1 for the original correspondences the solver ended with as inliers (final_inliers, registration.cc:1427)
  const std::vector<int>& getFinalInliers() const { return final_inliers_; }

private:
  Params params_;
  RegistrationSolution solution_;
  psulvsb_handle_t handle_ = nullptr;
  psulvsb_solution_t raw_ = {};
  int last_status_ = PSULVSB_OK;
  std::vector<int> final_inliers_;
};

} // namespace teaser
