// Minimal stand-in for <pcl/search/kdtree.h>: exact k-NN by brute force, results in the canonical
// (d2, index) order, d2 accumulated as FLANN's L2_Simple<float> does.
// See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_SEARCH_KDTREE_H
#define IFTWT_SHIM_PCL_SEARCH_KDTREE_H
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <algorithm>
#include <utility>
#include <vector>
namespace pcl {
namespace search {
template <class PointT>
class KdTree {
 public:
  typedef boost::shared_ptr<KdTree<PointT>> Ptr;
  typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
  void setInputCloud(const CloudConstPtr& c) { cloud_ = c; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    const std::size_t n = cloud_->points.size();
    std::vector<std::pair<float, int>> best;
    best.reserve((std::size_t)k + 1);
    for (std::size_t i = 0; i < n; ++i) {
      const PointT& p = cloud_->points[i];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      const float d = ((dx * dx) + (dy * dy)) + (dz * dz);
      const std::pair<float, int> key(d, (int)i);
      if ((int)best.size() < k) {
        best.insert(std::upper_bound(best.begin(), best.end(), key), key);
      } else if (key < best.back()) {
        best.pop_back();
        best.insert(std::upper_bound(best.begin(), best.end(), key), key);
      }
    }
    idx.resize(best.size());
    d2.resize(best.size());
    for (std::size_t i = 0; i < best.size(); ++i) {
      d2[i] = best[i].first;
      idx[i] = best[i].second;
    }
    return (int)best.size();
  }
  int nearestKSearch(int index, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    return nearestKSearch(cloud_->points[(std::size_t)index], k, idx, d2);
  }
 private:
  CloudConstPtr cloud_;
};
}  // namespace search
}  // namespace pcl
#endif
