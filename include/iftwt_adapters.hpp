// iftwt_adapters.hpp — header-only C++14 adapters that keep the reference's class and method
// names (Graph, FlatPoint, IFT_PCD, Cluster; the block main.cpp:91-129) on top of the C ABI
// of include/iftwt.h.
//
// Point types: real PCL types when <pcl/point_cloud.h> is available, otherwise a POD shim
// with PCL's memory layout [PCL-recall]: PointXYZ 16 B {x,y,z,pad}; PointXYZI 32 B
// {x,y,z,pad,intensity,pad x3}; Normal 32 B {nx,ny,nz,pad,curvature,pad x3}.  Either way the
// clouds cross the ABI zero-copy through the stride/offset form of iftwt_set_cloud.
//
// What changes for a caller, and why (INTEGRATION.md has the side-by-side):
//   * BGL-typed getters (graph_i, std::map<vd,vd>) become index-based (CSR, vectors): the
//     reference's vertex descriptors are heap addresses (setS), there is nothing to keep.
//   * Graph drops the original cloud and carries the voxel-centre cloud, as graph.hpp:307 does.
#ifndef IFTWT_ADAPTERS_HPP
#define IFTWT_ADAPTERS_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "iftwt.h"

#if defined(__has_include)
#if __has_include(<pcl/point_cloud.h>) && !defined(IFTWT_NO_PCL)
#define IFTWT_HAVE_PCL 1
#endif
#endif

#ifdef IFTWT_HAVE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace iftwt {
namespace types {
using PointXYZ = pcl::PointXYZ;
using PointXYZI = pcl::PointXYZI;
using PointXYZRGB = pcl::PointXYZRGB;
using Normal = pcl::Normal;
template <class P>
using PointCloud = pcl::PointCloud<P>;
}  // namespace types
}  // namespace iftwt
#else
namespace iftwt {
namespace types {
struct alignas(16) PointXYZ {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
};
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  float intensity = 0, pad2_[3] = {0, 0, 0};
};
struct alignas(16) PointXYZRGB {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  std::uint8_t b = 0, g = 0, r = 0, a = 255;
  float pad2_[3] = {0, 0, 0};
};
struct alignas(16) Normal {
  float normal_x = 0, normal_y = 0, normal_z = 0, pad_ = 0;
  float curvature = 0, pad2_[3] = {0, 0, 0};
};
template <class P>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  std::vector<P> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void push_back(const P& p) {
    points.push_back(p);
    width = (std::uint32_t)points.size();
  }
  P& operator[](std::size_t i) { return points[i]; }
  const P& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace types
}  // namespace iftwt
#endif

static_assert(sizeof(iftwt::types::PointXYZ) == 16, "PointXYZ layout");
static_assert(sizeof(iftwt::types::PointXYZI) == 32, "PointXYZI layout");
static_assert(sizeof(iftwt::types::Normal) == 32, "Normal layout");

// globals.hpp:29-36
namespace global {
typedef iftwt::types::PointXYZ Point;
typedef iftwt::types::PointCloud<Point> Cloud;
typedef iftwt::types::PointXYZRGB PointT;
typedef iftwt::types::PointCloud<PointT> CloudT;
typedef iftwt::types::PointXYZI PointI;
typedef iftwt::types::PointCloud<PointI> CloudI;
}  // namespace global

namespace iftwt {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// One ctx shared by the adapters of one pipeline (the reference hands its Graph around).
class Session {
 public:
  explicit Session(int device = 0) {
    int rc = iftwt_create(&ctx_, device);
    if (rc != IFTWT_OK) throw Error(rc, iftwt_last_error(nullptr));
  }
  ~Session() { iftwt_destroy(ctx_); }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;
  iftwt_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != IFTWT_OK) throw Error(rc, iftwt_last_error(ctx_));
  }

 private:
  iftwt_ctx* ctx_ = nullptr;
};

// index-based stand-in for global::graph_i (globals.hpp:37): CSR, rows ascending
struct CsrGraph {
  std::vector<std::uint32_t> off, adj;
};

}  // namespace iftwt

enum class MORPH : char { bin_erode = 1, bin_dilate = 2, dilate = 3, erode = 4 };  // graph.hpp:13

// graph.hpp:32-321
class Graph {
 public:
  // Graph(cloud) / Graph(cloud, overlap): the reference's voxel-26 graph (graph.hpp:34-50)
  explicit Graph(global::CloudI::Ptr cloud, double ovrlp = 1.01, int device = 0)
      : s_(std::make_shared<iftwt::Session>(device)) {
    upload(*cloud);
    double voxel = 0;
    s_->check(iftwt_cloud_resolution(s_->get(), ovrlp, &voxel));
    voxel_size_ = voxel;
    s_->check(iftwt_voxel_graph(s_->get(), voxel, &n_));
  }
  // graph builder (B): symmetric k-NN graph on the points themselves (BASELINE configs 2-5)
  static Graph knn(global::CloudI::Ptr cloud, int k, int device = 0) {
    Graph g(device);
    g.upload(*cloud);
    std::size_t arcs = 0;
    g.s_->check(iftwt_knn_graph(g.s_->get(), k, nullptr, nullptr, 0, &arcs));
    return g;
  }
  iftwt::CsrGraph getAdjacencyList() {  // graph.hpp:63-66
    iftwt::CsrGraph g;
    std::size_t arcs = 0;
    s_->check(iftwt_graph_csr(s_->get(), nullptr, nullptr, 0, &arcs));
    g.off.resize(n_ + 1);
    g.adj.resize(arcs);
    s_->check(iftwt_graph_csr(s_->get(), g.off.data(), g.adj.data(), arcs, &arcs));
    return g;
  }
  global::CloudI::Ptr getCloud() {  // graph.hpp:177-180
    std::vector<float> buf(4 * n_);
    std::size_t n = 0;
    s_->check(iftwt_get_cloud(s_->get(), buf.data(), n_, &n));
    global::CloudI::Ptr out(new global::CloudI());
    out->points.resize(n);
    for (std::size_t i = 0; i < n; ++i) {
      out->points[i].x = buf[4 * i];
      out->points[i].y = buf[4 * i + 1];
      out->points[i].z = buf[4 * i + 2];
      out->points[i].intensity = buf[4 * i + 3];
    }
    out->width = (std::uint32_t)n;
    out->height = 1;
    return out;
  }
  global::CloudI::Ptr morph_bin_erode(global::CloudI::Ptr out) { return morph(IFTWT_MORPH_BIN_ERODE, out); }
  global::CloudI::Ptr morph_bin_dilate(global::CloudI::Ptr out) { return morph(IFTWT_MORPH_BIN_DILATE, out); }
  global::CloudI::Ptr morph_erode(global::CloudI::Ptr out) { return morph(IFTWT_MORPH_ERODE, out); }
  global::CloudI::Ptr morph_dilate(global::CloudI::Ptr out) { return morph(IFTWT_MORPH_DILATE, out); }
  global::CloudI::Ptr morph_gradient() {  // graph.hpp:147-163
    return morph(IFTWT_MORPH_GRADIENT, global::CloudI::Ptr(new global::CloudI()));
  }
  // graph.hpp:164-176 — the reference returns a graph copy carrying the cloud's intensities;
  // here the graph itself takes them (what morph_erode_gradient does with the result, :203-208)
  void pc_to_g(global::CloudI::Ptr cloud) {
    std::vector<float> v(cloud->points.size());
    for (std::size_t i = 0; i < v.size(); ++i) v[i] = cloud->points[i].intensity;
    if (v.size() != n_) throw iftwt::Error(IFTWT_ERR_INVALID, "pc_to_g: cloud size differs from the graph");
    s_->check(iftwt_set_intensity(s_->get(), v.data()));
  }
  std::vector<std::uint32_t> getRegmin() {  // graph.hpp:209-226
    std::size_t ns = 0;
    s_->check(iftwt_regmin_seeds(s_->get(), nullptr, 0, &ns));
    std::vector<std::int32_t> s(ns);
    if (ns) s_->check(iftwt_regmin_seeds(s_->get(), s.data(), ns, &ns));
    return std::vector<std::uint32_t>(s.begin(), s.end());
  }
  std::size_t size() const { return n_; }
  double voxelSize() const { return voxel_size_; }
  const std::shared_ptr<iftwt::Session>& session() const { return s_; }

 private:
  explicit Graph(int device) : s_(std::make_shared<iftwt::Session>(device)) {}
  void upload(const global::CloudI& cloud) {
    n_ = cloud.points.size();
    s_->check(iftwt_set_cloud(s_->get(), cloud.points.data(), sizeof(global::PointI), n_,
                              (int)offsetof(global::PointI, x), (int)offsetof(global::PointI, intensity)));
  }
  global::CloudI::Ptr morph(int op, global::CloudI::Ptr out) {
    global::CloudI::Ptr c = getCloud();
    std::vector<float> v(n_);
    s_->check(iftwt_morph(s_->get(), op, v.data()));
    for (std::size_t i = 0; i < n_; ++i) c->points[i].intensity = v[i];
    if (out) {
      out->points.swap(c->points);
      out->width = (std::uint32_t)out->points.size();
      out->height = 1;
      return out;
    }
    return c;
  }
  std::shared_ptr<iftwt::Session> s_;
  std::size_t n_ = 0;
  double voxel_size_ = 0;
};

enum class MORPH_COLOR : char { dilate = 3, erode = 4 };  // graco.hpp:36

// graco.hpp:38-352 — the RGB twin of Graph; what the reference's live main() runs
// (GraCo gc(cloud_t, 5.0); gc.morph_erode(cloud_g), main.cpp:87-89).  The packed colour word
// of pcl::PointXYZRGB travels in the ctx's 4th channel untouched.
class GraCo {
 public:
  explicit GraCo(global::CloudT::Ptr cloud, double ovrlp = 1.01, int device = 0)
      : s_(std::make_shared<iftwt::Session>(device)) {
    static_assert(sizeof(global::PointT) == 32, "PointXYZRGB layout");
    n_ = cloud->points.size();
    s_->check(iftwt_set_cloud(s_->get(), cloud->points.data(), sizeof(global::PointT), n_, 0, 16));
    double voxel = 0;
    s_->check(iftwt_cloud_resolution(s_->get(), ovrlp, &voxel));
    voxel_size_ = voxel;
    s_->check(iftwt_voxel_graph(s_->get(), voxel, &n_));
  }
  global::CloudT::Ptr getCloud() { return cloud_with(nullptr); }  // graco.hpp:137-140
  global::CloudT::Ptr morph_erode(global::CloudT::Ptr out) { return morph(IFTWT_MORPH_ERODE, out); }
  global::CloudT::Ptr morph_dilate(global::CloudT::Ptr out) { return morph(IFTWT_MORPH_DILATE, out); }
  iftwt::CsrGraph getAdjacencyList() {
    iftwt::CsrGraph g;
    std::size_t arcs = 0;
    s_->check(iftwt_graph_csr(s_->get(), nullptr, nullptr, 0, &arcs));
    g.off.resize(n_ + 1);
    g.adj.resize(arcs);
    s_->check(iftwt_graph_csr(s_->get(), g.off.data(), g.adj.data(), arcs, &arcs));
    return g;
  }
  std::size_t size() const { return n_; }
  double voxelSize() const { return voxel_size_; }
  const std::shared_ptr<iftwt::Session>& session() const { return s_; }

 private:
  global::CloudT::Ptr cloud_with(const std::vector<std::uint32_t>* rgb) {
    std::vector<float> buf(4 * n_);
    std::size_t n = 0;
    s_->check(iftwt_get_cloud(s_->get(), buf.data(), n_, &n));
    global::CloudT::Ptr out(new global::CloudT());
    out->points.resize(n);
    for (std::size_t i = 0; i < n; ++i) {
      global::PointT& p = out->points[i];
      p.x = buf[4 * i];
      p.y = buf[4 * i + 1];
      p.z = buf[4 * i + 2];
      std::uint32_t w;
      if (rgb) w = (*rgb)[i];
      else std::memcpy(&w, &buf[4 * i + 3], 4);
      p.r = (std::uint8_t)(w >> 16);
      p.g = (std::uint8_t)(w >> 8);
      p.b = (std::uint8_t)w;
    }
    out->width = (std::uint32_t)n;
    out->height = 1;
    return out;
  }
  global::CloudT::Ptr morph(int op, global::CloudT::Ptr out) {
    std::vector<std::uint32_t> rgb(n_);
    s_->check(iftwt_morph_rgb(s_->get(), op, rgb.data()));
    global::CloudT::Ptr c = cloud_with(&rgb);
    if (out) {
      out->points.swap(c->points);
      out->width = (std::uint32_t)out->points.size();
      out->height = 1;
      return out;
    }
    return c;
  }
  std::shared_ptr<iftwt::Session> s_;
  std::size_t n_ = 0;
  double voxel_size_ = 0;
};

// main.cpp:16-76: cloudGray (otsu = false) and cloudRGB2GRAY (otsu = true).  The conversion runs
// on the device of a scratch ctx; the XYZI cloud it returns feeds Graph(cloud_i) (main.cpp:97).
inline global::CloudI::Ptr cloudGray(global::CloudT::Ptr cloud, bool otsu = false, double* threshold = nullptr,
                                     int device = 0) {
  iftwt::Session s(device);
  const std::size_t n = cloud->points.size();
  s.check(iftwt_set_cloud(s.get(), cloud->points.data(), sizeof(global::PointT), n, 0, 16));
  double t = -1;
  s.check(iftwt_rgb_to_gray(s.get(), otsu ? 1 : 0, &t));
  if (threshold) *threshold = t;
  std::vector<float> buf(4 * n);
  std::size_t m = 0;
  s.check(iftwt_get_cloud(s.get(), buf.data(), n, &m));
  global::CloudI::Ptr out(new global::CloudI());
  out->points.resize(m);
  for (std::size_t i = 0; i < m; ++i) {
    out->points[i].x = buf[4 * i];
    out->points[i].y = buf[4 * i + 1];
    out->points[i].z = buf[4 * i + 2];
    out->points[i].intensity = buf[4 * i + 3];
  }
  out->width = cloud->width;
  out->height = cloud->height;
  return out;
}
inline global::CloudI::Ptr cloudRGB2GRAY(global::CloudT::Ptr cloud, double* threshold = nullptr, int device = 0) {
  return cloudGray(cloud, true, threshold, device);
}

namespace iftwt {
// pcl::NormalEstimationOMP<PointI, Normal> as used at main.cpp:104-109, on the graph's cloud
class NormalEstimation {
 public:
  explicit NormalEstimation(Graph& g) : s_(g.session()), n_(g.size()) {}
  void setKSearch(int k) { k_ = k; }
  void compute(types::PointCloud<types::Normal>& out) {
    out.points.resize(n_);
    s_->check(iftwt_normals(s_->get(), k_, out.points.data(), sizeof(types::Normal),
                            (int)offsetof(types::Normal, normal_x), (int)offsetof(types::Normal, curvature)));
    out.width = (std::uint32_t)n_;
    out.height = 1;
  }

 private:
  std::shared_ptr<Session> s_;
  std::size_t n_;
  int k_ = 30;
};
}  // namespace iftwt

// flat_point.hpp:9-92 (pcl::RegionGrowing setters of main.cpp:111-119)
template <typename PointT = global::PointI, typename NormalT = iftwt::types::Normal>
class FlatPoint {
 public:
  explicit FlatPoint(Graph& g) : s_(g.session()), n_(g.size()) {
    p_.min_cluster_size = 1;
    p_.max_cluster_size = 0x7FFFFFFF;
    p_.number_of_neighbours = 30;
    p_.smoothness_threshold = 30.f / 180.f * 3.14159265f;
    p_.curvature_threshold = 0.05f;
  }
  void setMinClusterSize(int v) { p_.min_cluster_size = v; }
  void setMaxClusterSize(int v) { p_.max_cluster_size = v; }
  void setNumberOfNeighbours(int v) { p_.number_of_neighbours = v; }
  void setSmoothnessThreshold(float v) { p_.smoothness_threshold = v; }
  void setCurvatureThreshold(float v) { p_.curvature_threshold = v; }
  // seed = the input point nearest to each kept cluster's centroid (flat_point.hpp:50-56, 65-91)
  global::Cloud::Ptr getCentroidsCloud() {
    std::size_t ns = 0;
    s_->check(iftwt_region_seeds(s_->get(), &p_, nullptr, nullptr, 0, &ns));
    seeds_.resize(ns);
    if (ns) s_->check(iftwt_region_seeds(s_->get(), &p_, nullptr, seeds_.data(), ns, &ns));
    std::vector<float> buf(4 * n_);
    std::size_t n = 0;
    s_->check(iftwt_get_cloud(s_->get(), buf.data(), n_, &n));
    global::Cloud::Ptr c(new global::Cloud());
    for (std::int32_t i : seeds_) {
      global::Point p;
      p.x = buf[4 * (std::size_t)i];
      p.y = buf[4 * (std::size_t)i + 1];
      p.z = buf[4 * (std::size_t)i + 2];
      c->push_back(p);
    }
    return c;
  }
  const std::vector<std::int32_t>& seedIndices() const { return seeds_; }

 private:
  std::shared_ptr<iftwt::Session> s_;
  std::size_t n_;
  iftwt_rg_params p_;
  std::vector<std::int32_t> seeds_;
};

// ift.hpp:29-269
class IFT_PCD {
 public:
  // seeds given as points: each is snapped to its nearest graph vertex (add_seeds, ift.hpp:117-142)
  IFT_PCD(Graph& g, global::Cloud::Ptr seeds) : s_(g.session()), n_(g.size()) {
    std::vector<std::int32_t> idx(seeds->points.size());
    if (!idx.empty())
      s_->check(iftwt_knn_query(s_->get(), seeds->points.data(), sizeof(global::Point), idx.size(), 1, idx.data(),
                                nullptr));
    // NULL means "the seeds already on the device" to iftwt_ift: an empty seed cloud must not alias it
    // (the reference with no seeds leaves every vertex at MAX_COST, ift.hpp:101)
    static const std::int32_t none = 0;
    run(idx.empty() ? &none : idx.data(), idx.size());
  }
  // regional-minima seeds (ift.hpp:38-44)
  explicit IFT_PCD(Graph& g) : s_(g.session()), n_(g.size()) {
    std::size_t ns = 0;
    s_->check(iftwt_regmin_seeds(s_->get(), nullptr, 0, &ns));
    run(nullptr, 0);
  }
  const std::vector<std::uint32_t>& getRoots() const { return label_; }  // root_m, ift.hpp:48-50
  const std::vector<std::uint32_t>& getCosts() const { return cost_; }
  const std::shared_ptr<iftwt::Session>& session() const { return s_; }
  std::size_t size() const { return n_; }

 private:
  void run(const std::int32_t* seeds, std::size_t ns) {
    cost_.resize(n_);
    label_.resize(n_);
    s_->check(iftwt_ift(s_->get(), nullptr, seeds, ns, cost_.data(), label_.data()));
  }
  std::shared_ptr<iftwt::Session> s_;
  std::size_t n_;
  std::vector<std::uint32_t> cost_, label_;
};

// cluster.hpp:23-229
class Cluster {
 public:
  Cluster(IFT_PCD& ift, int num_regions) : cluster_(ift.size()) {
    ift.session()->check(iftwt_cluster(ift.session()->get(), num_regions, cluster_.data(), &n_clusters_));
  }
  const std::vector<std::uint32_t>& getLabels() const { return cluster_; }  // region id per vertex
  std::size_t regions() const { return n_clusters_; }

 private:
  std::vector<std::uint32_t> cluster_;
  std::size_t n_clusters_ = 0;
};

#endif  // IFTWT_ADAPTERS_HPP
