// Minimal stand-in for <pcl/point_cloud.h>.  See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_POINT_CLOUD_H
#define IFTWT_SHIM_PCL_POINT_CLOUD_H
#include <boost/shared_ptr.hpp>
#include <cstddef>
#include <cstdint>
#include <vector>
namespace pcl {
template <class PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = (std::uint32_t)points.size();
    height = 1;
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  void swap(PointCloud& o) {
    points.swap(o.points);
    std::swap(width, o.width);
    std::swap(height, o.height);
  }
  void clear() {
    points.clear();
    width = height = 0;
  }
};
template <class PointT>
void copyPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out) {
  out = in;
}
}  // namespace pcl
#endif
