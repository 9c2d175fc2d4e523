// Minimal stand-in for <pcl/octree/octree_pointcloud_adjacency.h>: the occupied-voxel set and its
// 26-adjacency as PCL 1.8's OctreePointCloudAdjacency produces them, restated from the published
// algorithm (SURVEY App. A.2 [PCL-recall]): bounding box over the finite points, getKeyBitSize
// (depth, symmetric padding to side 2^depth * res, binary64), key = (unsigned)((p - min') / res),
// vertex = voxel CENTRE, neighbours = occupied voxels at key offsets {-1,0,1}^3 clipped to the key
// range INCLUDING the voxel itself (a self loop).  Leaves are created in ascending Morton order of
// the key (x = most significant axis bit).  See oracle/shim/README.md.
#ifndef IFTWT_SHIM_PCL_OCTREE_ADJACENCY_H
#define IFTWT_SHIM_PCL_OCTREE_ADJACENCY_H
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>
namespace pcl {
namespace octree {
template <class PointT>
class OctreePointCloudAdjacency {
 public:
  typedef boost::shared_ptr<OctreePointCloudAdjacency<PointT>> Ptr;
  typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
  explicit OctreePointCloudAdjacency(double resolution) : res_(resolution) { last_resolution() = resolution; }
  static double& last_resolution() {  // shim extension: Graph keeps voxel_size private
    static double r = 0.0;
    return r;
  }
  void setInputCloud(const CloudConstPtr& c) { cloud_ = c; }
  std::size_t size() const { return keys_.size(); }
  double getVoxelSquaredSideLen() const { return res_ * res_; }
  unsigned getTreeDepth() const { return depth_; }
  void addPointsFromInputCloud() {
    double mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const PointT& p : cloud_->points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const float c[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; ++a) {
        if (c[a] < mn[a]) mn[a] = c[a];
        if (c[a] > mx[a]) mx[a] = c[a];
      }
    }
    const float minValue = FLT_EPSILON;  // getKeyBitSize
    unsigned mk[3];
    for (int a = 0; a < 3; ++a) mk[a] = (unsigned)std::ceil((mx[a] - mn[a] - minValue) / res_);
    const unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
    depth_ = std::max(std::min(32u, (unsigned)std::ceil(std::log((double)max_voxels) / std::log(2.0) - minValue)), 0u);
    const double side = (double)(1u << depth_) * res_;
    for (int a = 0; a < 3; ++a) {
      const double over = (side - (mx[a] - mn[a])) / 2.0;
      if (over > minValue) {
        mn[a] -= over;
        mx[a] += over;
      }
      min_[a] = mn[a];
    }
    std::map<std::uint64_t, std::array<std::uint32_t, 3>> vox;  // ascending Morton code
    for (const PointT& p : cloud_->points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const float c[3] = {p.x, p.y, p.z};
      std::array<std::uint32_t, 3> k;
      for (int a = 0; a < 3; ++a) k[a] = (std::uint32_t)(((double)c[a] - mn[a]) / res_);
      vox.emplace(morton(k[0], k[1], k[2]), k);
    }
    keys_.clear();
    codes_.clear();
    for (const auto& kv : vox) {
      codes_.push_back(kv.first);
      keys_.push_back(kv.second);
    }
  }
  template <class Alloc>
  int getOccupiedVoxelCenters(std::vector<PointT, Alloc>& out) const {
    out.clear();
    for (std::size_t v = 0; v < keys_.size(); ++v) out.push_back(centre(v));
    return (int)out.size();
  }
  template <class GraphT>
  void computeVoxelAdjacencyGraph(GraphT& g) {
    g.clear();
    std::vector<typename GraphT::vertex_descriptor> vd(keys_.size());
    for (std::size_t v = 0; v < keys_.size(); ++v) {
      vd[v] = add_vertex(g);
      g[vd[v]] = centre(v);
    }
    const std::uint32_t max_key = (1u << depth_) - 1u;
    for (std::size_t v = 0; v < keys_.size(); ++v) {
      const auto& k = keys_[v];
      for (int dx = (k[0] > 0 ? -1 : 0); dx <= (k[0] == max_key ? 0 : 1); ++dx)
        for (int dy = (k[1] > 0 ? -1 : 0); dy <= (k[1] == max_key ? 0 : 1); ++dy)
          for (int dz = (k[2] > 0 ? -1 : 0); dz <= (k[2] == max_key ? 0 : 1); ++dz) {
            const std::uint64_t code = morton(k[0] + dx, k[1] + dy, k[2] + dz);
            auto it = std::lower_bound(codes_.begin(), codes_.end(), code);
            if (it == codes_.end() || *it != code) continue;
            const std::size_t u = (std::size_t)(it - codes_.begin());
            auto e = add_edge(vd[v], vd[u], g);
            if (e.second) {
              const PointT a = g[vd[v]], b = g[vd[u]];
              const float ex = a.x - b.x, ey = a.y - b.y, ez = a.z - b.z;
              g[e.first] = std::sqrt(ex * ex + ey * ey + ez * ez);
            }
          }
    }
  }
 private:
  static std::uint64_t spread(std::uint64_t v) {
    std::uint64_t x = v & 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
  }
  static std::uint64_t morton(std::uint32_t x, std::uint32_t y, std::uint32_t z) {
    return (spread(x) << 2) | (spread(y) << 1) | spread(z);
  }
  PointT centre(std::size_t v) const {  // genLeafNodeCenterFromOctreeKey
    PointT p;
    p.x = (float)(((double)keys_[v][0] + 0.5f) * res_ + min_[0]);
    p.y = (float)(((double)keys_[v][1] + 0.5f) * res_ + min_[1]);
    p.z = (float)(((double)keys_[v][2] + 0.5f) * res_ + min_[2]);
    return p;
  }
  double res_;
  double min_[3] = {0, 0, 0};
  unsigned depth_ = 0;
  CloudConstPtr cloud_;
  std::vector<std::array<std::uint32_t, 3>> keys_;
  std::vector<std::uint64_t> codes_;
};
}  // namespace octree
}  // namespace pcl
#endif
