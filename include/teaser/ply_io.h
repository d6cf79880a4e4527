// teaser/ply_io.h -- PLYReader with the reference's name and read() signature
// (reference: teaser/include/teaser/ply_io.h + teaser/src/ply_io.cc:26-79, which wrap tinyply), implemented
// on the dependency-free reader of libpsulvsb_b200.so (include/psulvsb_io.h).  Vertices only (x, y, z,
// float32 or float64, ascii or binary) -- all the PSULVSB drivers read.
#pragma once

#include <cstdint>
#include <cstring>
#include <fstream>
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

/// teaser/include/teaser/ply_io.h:35-50 (+ teaser/src/ply_io.cc:81-130): vertices as float32 x, y, z, ascii or
/// binary little endian.  Returns 0 on success, -1 on failure.
class PLYWriter {
public:
  int write(const std::string& file_name, const PointCloud& cloud, bool binary_mode = false) {
    std::ofstream f(file_name, binary_mode ? (std::ios::out | std::ios::binary) : std::ios::out);
    if (!f) return -1;
    f << "ply\nformat " << (binary_mode ? "binary_little_endian" : "ascii") << " 1.0\n"
      << "element vertex " << cloud.size() << "\nproperty float x\nproperty float y\nproperty float z\nend_header\n";
    if (binary_mode) {
      std::vector<unsigned char> buf(12 * cloud.size());
      for (size_t i = 0; i < cloud.size(); ++i) {
        const float v[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
        for (int k = 0; k < 3; ++k) {
          uint32_t w;
          std::memcpy(&w, &v[k], 4);
          for (int b = 0; b < 4; ++b) buf[12 * i + 4 * k + b] = static_cast<unsigned char>((w >> (8 * b)) & 0xFFu);
        }
      }
      f.write(reinterpret_cast<const char*>(buf.data()), static_cast<std::streamsize>(buf.size()));
    } else {
      f.precision(9);
      for (size_t i = 0; i < cloud.size(); ++i) f << cloud[i].x << " " << cloud[i].y << " " << cloud[i].z << "\n";
    }
    return f.good() ? 0 : -1;
  }
};

} // namespace teaser
