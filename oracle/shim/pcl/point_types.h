// Minimal stand-in for <pcl/point_types.h>: the three point records the reference uses, with
// PCL 1.8's memory layout (16-byte aligned).  See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_POINT_TYPES_H
#define IFTWT_SHIM_PCL_POINT_TYPES_H
#include <cmath>
#include <cstdint>
#include <memory>
#define pcl_isfinite(x) std::isfinite(x)
namespace Eigen {
template <class T>
using aligned_allocator = std::allocator<T>;  // PCL's point vectors use Eigen's; alignment is moot here
}
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  PointXYZ() {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) PointXYZI {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  float intensity = 0.f, pad2_[3] = {0.f, 0.f, 0.f};
  PointXYZI() {}
  PointXYZI(float intensity_) : intensity(intensity_) {}
};
struct alignas(16) PointXYZRGB {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  union {
    struct {
      std::uint8_t b, g, r, a;
    };
    float rgb;
    std::uint32_t rgba;
  };
  float pad2_[3] = {0.f, 0.f, 0.f};
  PointXYZRGB() : rgba(0xff000000u) {}
};
struct alignas(16) Normal {
  float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, pad_ = 0.f;
  float curvature = 0.f, pad2_[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl
#endif
