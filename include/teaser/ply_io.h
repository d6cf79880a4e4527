// teaser/ply_io.h -- PLYReader with the reference's name and read() signature
// (reference: teaser/include/teaser/ply_io.h + teaser/src/ply_io.cc:26-79, which wrap tinyply), implemented
// on the dependency-free reader of libpsulvsb_b200.so (include/psulvsb_io.h).  Vertices only (x, y, z,
// float32 or float64, ascii or binary) -- all the PSULVSB drivers read.
#pragma once

#include <string>
#include <vector>

#include "../psulvsb_io.h"
#include "geometry.h"

namespace teaser {

class PLYReader {
public:
  /// returns 0 on success, -1 on failure (as the reference does)
  int read(const std::string& file_name, PointCloud& cloud) {
    long long n = 0;
    if (psulvsb_ply_vertex_count(file_name.c_str(), &n) != PSULVSB_OK) return -1;
    std::vector<float> xyz(static_cast<size_t>(3 * (n > 0 ? n : 1)));
    if (psulvsb_ply_read_xyz(file_name.c_str(), xyz.data(), n, &n) != PSULVSB_OK) return -1;
    cloud.reserve(cloud.size() + static_cast<size_t>(n));
    for (long long i = 0; i < n; ++i) cloud.push_back({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
    return 0;
  }
};

} // namespace teaser
