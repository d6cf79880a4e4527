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
// Also declared, with the reference's names and signatures, because the reference's own unit tests and
// solveForScale / solveForRotation / solveForTranslation call them: the abstract sub-solver interfaces (:47-101),
// ScaleInliersSelector (:203-215), TLSScaleSolver (:180-198), TLSTranslationSolver (:221-239), GNCRotationSolver /
// GNCTLSRotationSolver (:244-290) and computeTIMs (:523) -- each one call into the C ABI (CUDA).
//
// Not carried over (outside the PSULVSB path): FastGlobalRegistrationSolver,
// ScalarTLSEstimator::estimate_tiled, set*Estimator, the TIM / mask getters of the upstream pipeline (the fork's
// solve() no longer maintains them), the certifier.
#pragma once

#include <Eigen/Core>

#include <cstdint>
#include <map>
#include <memory>
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

// ---- sub-solvers (registration.h:47-317) ------------------------------------------------------
namespace detail {
template <typename Mask>
inline void fill_mask(Mask* inliers, const std::vector<unsigned char>& m) {
  if (!inliers) return;
  inliers->resize(1, static_cast<long>(m.size()));
  for (size_t i = 0; i < m.size(); ++i) (*inliers)(0, static_cast<long>(i)) = m[i] != 0;
}
} // namespace detail

class AbstractScaleSolver { // registration.h:47-61
public:
  virtual ~AbstractScaleSolver() {}
  virtual void solveForScale(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                             const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, double* scale,
                             Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) = 0;
};

class AbstractRotationSolver { // registration.h:66-81
public:
  virtual ~AbstractRotationSolver() {}
  virtual void solveForRotation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                                const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, Eigen::Matrix3d* rotation,
                                Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) = 0;
};

class AbstractTranslationSolver { // registration.h:86-101
public:
  virtual ~AbstractTranslationSolver() {}
  virtual void solveForTranslation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                                   const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, Eigen::Vector3d* translation,
                                   Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) = 0;
};

/// registration.cc:418-434: known scale; inliers = line vectors whose lengths agree within 2 noise_bound sqrt(cbar2)
class ScaleInliersSelector : public AbstractScaleSolver {
public:
  ScaleInliersSelector() = delete;
  explicit ScaleInliersSelector(double noise_bound, double cbar2) : noise_bound_(noise_bound), cbar2_(cbar2) {}
  void solveForScale(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                     const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, double* scale,
                     Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) override {
    std::vector<unsigned char> m(static_cast<size_t>(src.cols()), 0);
    status_ = psulvsb_scale_inliers_host(src.data(), dst.data(), static_cast<unsigned long long>(src.cols()),
                                         noise_bound_, cbar2_, m.data());
    if (scale) *scale = 1;
    detail::fill_mask(inliers, m);
  }
  int lastStatus() const { return status_; }
  double noiseBound() const { return noise_bound_; }
  double cbar2() const { return cbar2_; }

private:
  double noise_bound_;
  double cbar2_;
  int status_ = PSULVSB_OK;
};

/// registration.cc:397-415 -> :66-120.  The reference draws its RANSAC candidates with rand() seeded by time();
/// here draw k of call e comes from philox(seed; PSULVSB_DOMAIN_SCALE, e, k).  setLastBest() supplies the
/// reference's file-scope scale_last_best (first candidate of later calls, :75-86).
class TLSScaleSolver : public AbstractScaleSolver {
public:
  TLSScaleSolver() = delete;
  explicit TLSScaleSolver(double noise_bound, double cbar2, uint64_t seed = 0)
      : noise_bound_(noise_bound), cbar2_(cbar2), seed_(seed) {}
  void setLastBest(double scale) {
    last_best_ = scale;
    have_last_ = true;
  }
  void solveForScale(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                     const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, double* scale,
                     Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) override {
    std::vector<unsigned char> m(static_cast<size_t>(src.cols()), 0);
    double s = 1;
    status_ = psulvsb_tls_scale_host(src.data(), dst.data(), static_cast<int>(src.cols()), noise_bound_, cbar2_, seed_,
                                     calls_++, have_last_ ? &last_best_ : nullptr, &s, m.data());
    if (scale) *scale = s;
    detail::fill_mask(inliers, m);
  }
  int lastStatus() const { return status_; }
  double noiseBound() const { return noise_bound_; }
  double cbar2() const { return cbar2_; }

private:
  double noise_bound_;
  double cbar2_;
  uint64_t seed_;
  uint32_t calls_ = 0;
  double last_best_ = 1;
  bool have_last_ = false;
  int status_ = PSULVSB_OK;
};

/// registration.cc:436-463 -> :121-203: per-axis max-stabbing.  setLastBest(): translation_last_best (:136-161).
class TLSTranslationSolver : public AbstractTranslationSolver {
public:
  TLSTranslationSolver() = delete;
  explicit TLSTranslationSolver(double noise_bound, double cbar2) : noise_bound_(noise_bound), cbar2_(cbar2) {}
  void setLastBest(const Eigen::Vector3d& t) {
    for (int r = 0; r < 3; ++r) last_best_[r] = t(r, 0);
    have_last_ = true;
  }
  void solveForTranslation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                           const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, Eigen::Vector3d* translation,
                           Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) override {
    std::vector<unsigned char> m(static_cast<size_t>(src.cols()), 0);
    double t[3] = {0, 0, 0};
    status_ = psulvsb_tls_translation_host(src.data(), dst.data(), static_cast<int>(src.cols()), noise_bound_, cbar2_,
                                           have_last_ ? last_best_ : nullptr, t, m.data());
    if (translation)
      for (int r = 0; r < 3; ++r) (*translation)(r, 0) = t[r];
    detail::fill_mask(inliers, m);
  }
  int lastStatus() const { return status_; }
  double noiseBound() const { return noise_bound_; }
  double cbar2() const { return cbar2_; }

private:
  double noise_bound_;
  double cbar2_;
  double last_best_[3] = {0, 0, 0};
  bool have_last_ = false;
  int status_ = PSULVSB_OK;
};

class GNCRotationSolver : public AbstractRotationSolver { // registration.h:244-262
public:
  struct Params {
    size_t max_iterations;
    double cost_threshold;
    double gnc_factor;
    double noise_bound;
  };
  GNCRotationSolver(Params params) : params_(params) {}
  Params getParams() { return params_; }
  void setParams(Params params) { params_ = params; }
  double getCostAtTermination() { return cost_; }

protected:
  Params params_;
  double cost_ = 0;
};

/// registration.cc:1563-1692 + utils.h:121-136.  setLastBest(): rotation_last_best, the warm start of later calls.
class GNCTLSRotationSolver : public GNCRotationSolver {
public:
  GNCTLSRotationSolver() = delete;
  explicit GNCTLSRotationSolver(Params params) : GNCRotationSolver(params) {}
  void setLastBest(const Eigen::Matrix3d& R) {
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) last_best_[c * 3 + r] = R(r, c);
    have_last_ = true;
  }
  void solveForRotation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& src,
                        const Eigen::Matrix<double, 3, Eigen::Dynamic>& dst, Eigen::Matrix3d* rotation,
                        Eigen::Matrix<bool, 1, Eigen::Dynamic>* inliers) override {
    std::vector<unsigned char> m(static_cast<size_t>(src.cols()), 0);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    status_ = psulvsb_gnc_tls_rotation_host(src.data(), dst.data(), static_cast<unsigned long long>(src.cols()),
                                            params_.noise_bound, static_cast<int>(params_.max_iterations),
                                            params_.gnc_factor, params_.cost_threshold,
                                            have_last_ ? last_best_ : nullptr, R, m.data(), &cost_, &iterations_);
    if (rotation)
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) (*rotation)(r, c) = R[c * 3 + r];
    detail::fill_mask(inliers, m);
  }
  int getIterations() const { return iterations_; }
  int lastStatus() const { return status_; }

private:
  double last_best_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool have_last_ = false;
  int iterations_ = 0;
  int status_ = PSULVSB_OK;
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

  /// registration.h:559-580.  The reference's solve() calls whatever estimator objects it holds; here the path runs
  /// as CUDA kernels, so an estimator is accepted when it is one of this header's own classes (the kernels behind
  /// them ARE the path) -- its parameters then replace the corresponding Params fields, as installing it in the
  /// reference would -- and any other AbstractScaleSolver / GNCRotationSolver / AbstractTranslationSolver subclass
  /// makes the next solve() fail with PSULVSB_ERR_UNSUPPORTED (valid = false) instead of being silently ignored.
  void setScaleEstimator(std::unique_ptr<AbstractScaleSolver> estimator) { scale_solver_ = std::move(estimator); }
  void setRotationEstimator(std::unique_ptr<GNCRotationSolver> estimator) { rotation_solver_ = std::move(estimator); }
  void setTranslationEstimator(std::unique_ptr<AbstractTranslationSolver> estimator) {
    translation_solver_ = std::move(estimator);
  }
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
    // deprecated fields first, exactly as registration.cc:628-637 (the second one wins when both are cleared)
    if (!params_.use_max_clique) params_.inlier_selection_mode = INLIER_SELECTION_MODE::NONE;
    if (!params_.max_clique_exact_solution) params_.inlier_selection_mode = INLIER_SELECTION_MODE::PMC_HEU;
    p.inlier_selection_mode = static_cast<int>(params_.inlier_selection_mode);
    p.kcore_heuristic_threshold = params_.kcore_heuristic_threshold;
    p.seed = params_.seed;
    if (params_.replay) p.wallclock_cap_s = 0.0;
    if (scale_solver_) {
      if (auto* sel = dynamic_cast<ScaleInliersSelector*>(scale_solver_.get())) {
        p.estimate_scaling = 0;
        p.noise_bound = sel->noiseBound();
        p.cbar2 = sel->cbar2();
      } else if (auto* tls = dynamic_cast<TLSScaleSolver*>(scale_solver_.get())) {
        p.estimate_scaling = 1;
        p.noise_bound = tls->noiseBound();
        p.cbar2 = tls->cbar2();
      } else {
        last_status_ = PSULVSB_ERR_UNSUPPORTED;
        return solution_;
      }
    }
    if (rotation_solver_) {
      if (auto* gnc = dynamic_cast<GNCTLSRotationSolver*>(rotation_solver_.get())) {
        const GNCRotationSolver::Params rp = gnc->getParams();
        p.rotation_max_iterations = static_cast<int>(rp.max_iterations);
        p.rotation_cost_threshold = rp.cost_threshold;
        p.rotation_gnc_factor = rp.gnc_factor;
      } else {
        last_status_ = PSULVSB_ERR_UNSUPPORTED;
        return solution_;
      }
    }
    if (translation_solver_ && !dynamic_cast<TLSTranslationSolver*>(translation_solver_.get())) {
      last_status_ = PSULVSB_ERR_UNSUPPORTED;
      return solution_;
    }
    const int C = static_cast<int>(src.cols());
    // ori_src / ori_dst absent (upstream-style callers): the reduced set is the whole set
    const bool have_ori = params_.ori_src.cols() > 0;
    const int M = have_ori ? static_cast<int>(params_.ori_src.cols()) : C;
    std::vector<int> keep(static_cast<size_t>(M), 1), dense(static_cast<size_t>(M), -1);
    if (have_ori && static_cast<int>(params_.keep_mask.size()) == M) {
      keep = params_.keep_mask;
      for (const auto& kv : params_.reduce_map)
        if (kv.first >= 0 && kv.first < M) dense[static_cast<size_t>(kv.first)] = kv.second;
      // the reference reads params_.reduce_map[j] through std::map::operator[] (registration.cc:1433): a kept
      // correspondence without an entry yields column 0
      for (int j = 0; j < M; ++j)
        if (keep[static_cast<size_t>(j)] == 1 && dense[static_cast<size_t>(j)] < 0) dense[static_cast<size_t>(j)] = 0;
    } else if (have_ori && M != C) {
      // ori_src / ori_dst without a keep_mask of M entries: the reference would index keep_mask out of bounds
      last_status_ = PSULVSB_ERR_INVALID;
      return solution_;
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
    std::vector<int> final_inliers(static_cast<size_t>(M), 0), map_out(static_cast<size_t>(M), -1);
    psulvsb_trace_t trace = {};
    trace.final_inliers = final_inliers.data();
    trace.reduce_map_out = map_out.data();
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
    if (raw_.final_C > C && have_ori) {  // self-update grew the working set: append the adopted correspondences to the
      src.conservativeResize(3, raw_.final_C);  // caller's matrices, as the reference does (registration.cc:800-806)
      dst.conservativeResize(3, raw_.final_C);
      for (int j = 0; j < M; ++j) {
        const int col = map_out[static_cast<size_t>(j)];
        if (col < C || col >= raw_.final_C) continue;
        for (int r = 0; r < 3; ++r) {
          src(r, col) = params_.ori_src(r, j);
          dst(r, col) = params_.ori_dst(r, j);
        }
      }
    }
    return solution_;
  }

  /// registration.cc:471-505: all pairwise differences v_j - v_i (i < j), column i*N - i(i+1)/2 + (j-i-1); map = (i, j)
  Eigen::Matrix<double, 3, Eigen::Dynamic> computeTIMs(const Eigen::Matrix<double, 3, Eigen::Dynamic>& v,
                                                      Eigen::Matrix<int, 2, Eigen::Dynamic>* map) {
    const long N = static_cast<long>(v.cols());
    const long L = N * (N - 1) / 2;
    Eigen::Matrix<double, 3, Eigen::Dynamic> vtilde;
    vtilde.resize(3, L);
    if (map) map->resize(2, L);
    last_status_ = psulvsb_compute_tims_host(v.data(), static_cast<int>(N), vtilde.data(), map ? map->data() : nullptr);
    return vtilde;
  }

  /// registration.cc:1537-1560: the configured sub-solvers on explicit line vectors / points
  double solveForScale(const Eigen::Matrix<double, 3, Eigen::Dynamic>& v1,
                       const Eigen::Matrix<double, 3, Eigen::Dynamic>& v2) {
    Eigen::Matrix<bool, 1, Eigen::Dynamic>& mask = scale_inliers_mask_;
    double s = 1;
    if (params_.estimate_scaling) {
      TLSScaleSolver solver(params_.noise_bound, params_.cbar2, params_.seed);
      solver.solveForScale(v1, v2, &s, &mask);
      last_status_ = solver.lastStatus();
    } else {
      ScaleInliersSelector solver(params_.noise_bound, params_.cbar2);
      solver.solveForScale(v1, v2, &s, &mask);
      last_status_ = solver.lastStatus();
    }
    solution_.scale = s;
    return s;
  }
  Eigen::Matrix3d solveForRotation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& v1,
                                   const Eigen::Matrix<double, 3, Eigen::Dynamic>& v2) {
    GNCRotationSolver::Params gp = {params_.rotation_max_iterations, params_.rotation_cost_threshold,
                                    params_.rotation_gnc_factor, params_.noise_bound};
    GNCTLSRotationSolver solver(gp);
    solver.solveForRotation(v1, v2, &solution_.rotation, &rotation_inliers_mask_);
    last_status_ = solver.lastStatus();
    return solution_.rotation;
  }
  Eigen::Vector3d solveForTranslation(const Eigen::Matrix<double, 3, Eigen::Dynamic>& v1,
                                      const Eigen::Matrix<double, 3, Eigen::Dynamic>& v2) {
    TLSTranslationSolver solver(params_.noise_bound, params_.cbar2);
    solver.solveForTranslation(v1, v2, &solution_.translation, &translation_inliers_mask_);
    last_status_ = solver.lastStatus();
    return solution_.translation;
  }

  /// masks of the last solveForScale / solveForRotation / solveForTranslation call (registration.h:588-613)
  Eigen::Matrix<bool, 1, Eigen::Dynamic> getScaleInliersMask() { return scale_inliers_mask_; }
  Eigen::Matrix<bool, 1, Eigen::Dynamic> getRotationInliersMask() { return rotation_inliers_mask_; }
  Eigen::Matrix<bool, 1, Eigen::Dynamic> getTranslationInliersMask() { return translation_inliers_mask_; }

  /// 1 for the original correspondences the solver ended with as inliers (final_inliers, registration.cc:1427)
  const std::vector<int>& getFinalInliers() const { return final_inliers_; }

private:
  Params params_;
  RegistrationSolution solution_;
  psulvsb_handle_t handle_ = nullptr;
  std::unique_ptr<AbstractScaleSolver> scale_solver_;
  std::unique_ptr<GNCRotationSolver> rotation_solver_;
  std::unique_ptr<AbstractTranslationSolver> translation_solver_;
  psulvsb_solution_t raw_ = {};
  int last_status_ = PSULVSB_OK;
  std::vector<int> final_inliers_;
  Eigen::Matrix<bool, 1, Eigen::Dynamic> scale_inliers_mask_, rotation_inliers_mask_, translation_inliers_mask_;
};

} // namespace teaser
