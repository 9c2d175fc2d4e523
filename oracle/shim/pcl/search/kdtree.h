// Minimal stand-in for <pcl/search/kdtree.h>: EXACT k-NN, results in the canonical (d2, index)
// order, d2 accumulated as FLANN's L2_Simple<float> does.  The search walks
// outwards from the query in a copy of the cloud sorted by x and stops a side once dx*dx
// exceeds the k-th best — a pruned brute force, no approximation.  See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_SEARCH_KDTREE_H
#define IFTWT_SHIM_PCL_SEARCH_KDTREE_H
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>
namespace pcl {
namespace search {
template <class PointT>
class KdTree {
 public:
  typedef boost::shared_ptr<KdTree<PointT>> Ptr;
  typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
  void setInputCloud(const CloudConstPtr& c) {
    cloud_ = c;
    const std::size_t n = c->points.size();
    order_.clear();
    for (std::size_t i = 0; i < n; ++i)
      if (std::isfinite(c->points[i].x) && std::isfinite(c->points[i].y) && std::isfinite(c->points[i].z))
        order_.push_back((int)i);
    std::sort(order_.begin(), order_.end(), [&](int a, int b) {
      return c->points[a].x < c->points[b].x || (c->points[a].x == c->points[b].x && a < b);
    });
    xs_.resize(order_.size());
    for (std::size_t i = 0; i < order_.size(); ++i) xs_[i] = c->points[order_[i]].x;
  }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    std::vector<std::pair<float, int>> best;  // ascending (d2, index), at most k
    best.reserve((std::size_t)k + 1);
    auto offer = [&](int i) {
      const PointT& p = cloud_->points[(std::size_t)i];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      const std::pair<float, int> key(((dx * dx) + (dy * dy)) + (dz * dz), i);
      if ((int)best.size() < k) {
        best.insert(std::upper_bound(best.begin(), best.end(), key), key);
      } else if (key < best.back()) {
        best.pop_back();
        best.insert(std::upper_bound(best.begin(), best.end(), key), key);
      }
    };
    const long n = (long)order_.size();
    long r = (long)(std::lower_bound(xs_.begin(), xs_.end(), q.x) - xs_.begin()), l = r - 1;
    bool go_l = l >= 0, go_r = r < n;
    while (go_l || go_r) {
      if (go_r) {
        const float dx = xs_[(std::size_t)r] - q.x;
        if ((int)best.size() == k && dx * dx > best.back().first) go_r = false;
        else {
          offer(order_[(std::size_t)r]);
          go_r = ++r < n;
        }
      }
      if (go_l) {
        const float dx = q.x - xs_[(std::size_t)l];
        if ((int)best.size() == k && dx * dx > best.back().first) go_l = false;
        else {
          offer(order_[(std::size_t)l]);
          go_l = --l >= 0;
        }
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
  std::vector<int> order_;
  std::vector<float> xs_;
};
}  // namespace search
}  // namespace pcl
#endif
