// teaser/geometry.h -- PointXYZ / PointCloud with the reference's names and members
// (reference: teaser/include/teaser/geometry.h:15-70).  Boundary types of
// RobustRegistrationSolver::solve(const PointCloud&, const PointCloud&, correspondences).
#pragma once

#include <cstddef>
#include <vector>

namespace teaser {

struct PointXYZ {
  float x;
  float y;
  float z;
  friend inline bool operator==(const PointXYZ& lhs, const PointXYZ& rhs) {
    return lhs.x == rhs.x && lhs.y == rhs.y && lhs.z == rhs.z;
  }
  friend inline bool operator!=(const PointXYZ& lhs, const PointXYZ& rhs) { return !(lhs == rhs); }
};

class PointCloud {
public:
  PointCloud() = default;
  std::vector<PointXYZ>::iterator begin() { return points_.begin(); }
  std::vector<PointXYZ>::iterator end() { return points_.end(); }
  std::vector<PointXYZ>::const_iterator begin() const { return points_.begin(); }
  std::vector<PointXYZ>::const_iterator end() const { return points_.end(); }
  std::size_t size() const { return points_.size(); }
  void reserve(std::size_t n) { points_.reserve(n); }
  bool empty() const { return points_.empty(); }
  PointXYZ& operator[](std::size_t i) { return points_[i]; }
  const PointXYZ& operator[](std::size_t i) const { return points_[i]; }
  PointXYZ& at(std::size_t n) { return points_.at(n); }
  const PointXYZ& at(std::size_t n) const { return points_.at(n); }
  void push_back(const PointXYZ& pt) { points_.push_back(pt); }
  void push_back(PointXYZ& pt) { points_.push_back(pt); }
  void clear() { points_.clear(); }

private:
  std::vector<PointXYZ> points_;
};

} // namespace teaser
