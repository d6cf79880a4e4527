// The rest of the reference's public surface used by its tests and drivers: setScale/Rotation/TranslationEstimator
// (registration.h:559-580) and PLYWriter (ply_io.h:35-50), against include/teaser/*.h.
#include <teaser/ply_io.h>
#include <teaser/registration.h>

#include <cmath>
#include <cstdio>
#include <memory>

struct ForeignTranslationSolver : teaser::AbstractTranslationSolver {
  void solveForTranslation(const Eigen::Matrix<double, 3, Eigen::Dynamic>&, const Eigen::Matrix<double, 3, Eigen::Dynamic>&,
                           Eigen::Vector3d*, Eigen::Matrix<bool, 1, Eigen::Dynamic>*) override {}
};

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/tmp/facade_extras.ply";
  int fails = 0;
  // ---- PLYWriter -> PLYReader round trip, ascii and binary
  teaser::PointCloud cloud;
  for (int i = 0; i < 257; ++i) cloud.push_back({0.125f * i, -1.5f + 0.01f * i, 1.0f / (1 + i)});
  for (int binary = 0; binary < 2; ++binary) {
    teaser::PLYWriter w;
    if (w.write(path, cloud, binary != 0) != 0) { std::printf("FAIL write %d\n", binary); ++fails; continue; }
    teaser::PLYReader r;
    teaser::PointCloud back;
    if (r.read(path, back) != 0 || back.size() != cloud.size()) { std::printf("FAIL read %d\n", binary); ++fails; continue; }
    double worst = 0;
    for (size_t i = 0; i < cloud.size(); ++i)
      worst = std::fmax(worst, std::fmax(std::fabs(back[i].x - cloud[i].x),
                                         std::fmax(std::fabs(back[i].y - cloud[i].y), std::fabs(back[i].z - cloud[i].z))));
    if (worst > (binary ? 0.0 : 1e-6)) { std::printf("FAIL roundtrip %d worst %g\n", binary, worst); ++fails; }
    else std::printf("ply roundtrip %s ok\n", binary ? "binary" : "ascii");
  }
  // ---- estimator setters: this header's own classes are accepted, a foreign subclass is refused loudly
  teaser::RobustRegistrationSolver::Params params;
  params.noise_bound = 0.05;
  params.estimate_scaling = false;
  teaser::RobustRegistrationSolver solver(params);
  solver.setScaleEstimator(std::make_unique<teaser::ScaleInliersSelector>(0.05, 1.0));
  solver.setRotationEstimator(std::make_unique<teaser::GNCTLSRotationSolver>(
      teaser::GNCRotationSolver::Params{100, 0.005, 1.4, 0.05}));
  solver.setTranslationEstimator(std::make_unique<ForeignTranslationSolver>());
  Eigen::Matrix<double, 3, Eigen::Dynamic> src(3, 8), dst(3, 8);
  for (int i = 0; i < 8; ++i)
    for (int r = 0; r < 3; ++r) {
      src(r, i) = 0.1 * (i + 1) * (r + 1) + 0.01 * i * i;
      dst(r, i) = src(r, i) + 1.0;
    }
  auto sol = solver.solve(src, dst);
  // without a device psulvsb_create fails first (PSULVSB_ERR_NO_DEVICE); with one, the foreign estimator is refused
  if (sol.valid || (solver.lastStatus() != PSULVSB_ERR_UNSUPPORTED && solver.lastStatus() != PSULVSB_ERR_NO_DEVICE)) {
    std::printf("FAIL foreign estimator accepted (status %d)\n", solver.lastStatus());
    ++fails;
  } else {
    std::printf("foreign estimator refused ok (status %d)\n", solver.lastStatus());
  }
  return fails ? 1 : 0;
}
