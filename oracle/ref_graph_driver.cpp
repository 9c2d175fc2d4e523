// ref_graph_driver.cpp — runs the REFERENCE's own Graph (graph.hpp) and IFT_PCD (ift.hpp) on
// it, both #included unmodified from /root/reference: `Graph gg(cloud, overlap)` followed by
// `IFT_PCD ift(gg, seeds)` is main.cpp:97 + 123.
//
// Test infrastructure (like everything under oracle/).  Built by `make -C oracle ref` into
// oracle/_ref/libref_graph.so against the stand-in headers of oracle/shim.
//
// What is the reference's here: computeCloudResolution, initialize, optimizeGraph, find_p,
// getCloud / g_to_pc, getAdjacencyList, copy_g, pc_to_g, morph_base and its five wrappers,
// morph_gradient, morph_erode_gradient, getRegmin (graph.hpp:15-31, 34-66, 130-226, 228-371), all of
// IFT_PCD, and GraCo's constructor, optimizeGraph, getCloud and colour erode / dilate
// (graco.hpp:19-64, 137-144, 176-179, 246-296, 301-351).  What is the shim's: k-NN search, the occupied-voxel set and its adjacency
// (PCL's part of the work, restated from the published algorithm — SURVEY App. A.1 / A.2) and
// the Boost.Graph containers.
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "globals.hpp"  // /root/reference
#include <pcl/octree/octree_pointcloud_adjacency.h>
#include <pcl/search/kdtree.h>
#include <array>
#include "graph.hpp"    // /root/reference
#include "ift.hpp"      // /root/reference
#include "graco.hpp"    // /root/reference

namespace {
struct Quiet {  // the reference prints counts and timings on stdout: park fd 1 on /dev/null
  int saved, devnull;
  Quiet() {
    fflush(stdout);
    std::cout.flush();
    saved = dup(1);
    devnull = open("/dev/null", O_WRONLY);
    if (devnull >= 0) dup2(devnull, 1);
  }
  ~Quiet() {
    fflush(stdout);
    std::cout.flush();
    if (saved >= 0) {
      dup2(saved, 1);
      close(saved);
    }
    if (devnull >= 0) close(devnull);
  }
};
struct Handle {
  Graph* g;
  double resolution;
};
typedef global::graph_i GI;
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// Graph(CloudI::Ptr, double) — graph.hpp:43-50
REF_API void* ref_graph_create(const float* pts4, int64_t n, double overlap) {
  global::CloudI::Ptr cloud(new global::CloudI());
  for (int64_t i = 0; i < n; ++i) {
    global::PointI p;
    p.x = pts4[4 * i];
    p.y = pts4[4 * i + 1];
    p.z = pts4[4 * i + 2];
    p.intensity = pts4[4 * i + 3];
    cloud->push_back(p);
  }
  Quiet q;
  Handle* h = new Handle;
  h->g = new Graph(cloud, overlap);
  h->resolution = pcl::octree::OctreePointCloudAdjacency<global::PointI>::last_resolution();  // = voxel_size
  return h;
}
REF_API void ref_graph_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  delete h->g;
  delete h;
}
REF_API double ref_graph_resolution(void* hv) { return ((Handle*)hv)->resolution; }
// getCloud() — graph.hpp:177-180; returns V, writes V x {x, y, z, intensity} when out != NULL
REF_API int64_t ref_graph_cloud(void* hv, float* out4) {
  global::CloudI::Ptr c = ((Handle*)hv)->g->getCloud();
  if (out4)
    for (size_t i = 0; i < c->size(); ++i) {
      out4[4 * i] = c->points[i].x;
      out4[4 * i + 1] = c->points[i].y;
      out4[4 * i + 2] = c->points[i].z;
      out4[4 * i + 3] = c->points[i].intensity;
    }
  return (int64_t)c->size();
}
// getAdjacencyList() — graph.hpp:63-66 — as CSR in vertex order, the self loop left out;
// returns the number of arcs, self_loops = vertices that carry one
REF_API int64_t ref_graph_csr(void* hv, uint32_t* off, uint32_t* adj, int64_t* self_loops) {
  GI g = ((Handle*)hv)->g->getAdjacencyList();
  int64_t m = 0, v = 0, loops = 0;
  for (auto u : boost::make_iterator_range(boost::vertices(g))) {
    off[v++] = (uint32_t)m;
    for (auto t : boost::make_iterator_range(boost::adjacent_vertices(u, g))) {
      if (t == u) {
        ++loops;
        continue;
      }
      if (adj) adj[m] = (uint32_t)GI::index_of(t);
      ++m;
    }
  }
  off[v] = (uint32_t)m;
  if (self_loops) *self_loops = loops;
  return m;
}
// op: 1 bin_erode, 2 bin_dilate, 3 dilate, 4 erode (enum MORPH, graph.hpp:13), 5 morph_gradient
// (graph.hpp:147-163), 6 morph_erode_gradient (graph.hpp:203-208; REPLACES the graph's values by
// the gradient through pc_to_g, graph.hpp:164-176)
REF_API int ref_graph_morph(void* hv, int op, float* out) {
  Graph* g = ((Handle*)hv)->g;
  global::CloudI::Ptr c(new global::CloudI());
  Quiet q;
  switch (op) {
    case 1: c = g->morph_bin_erode(c); break;
    case 2: c = g->morph_bin_dilate(c); break;
    case 3: c = g->morph_dilate(c); break;
    case 4: c = g->morph_erode(c); break;
    case 5: c = g->morph_gradient(); break;
    case 6: c = g->morph_erode_gradient(); break;
    default: return -1;
  }
  for (size_t i = 0; i < c->size(); ++i) out[i] = c->points[i].intensity;
  return 0;
}
// getRegmin() — graph.hpp:209-226
REF_API int64_t ref_graph_regmin(void* hv, int32_t* idx) {
  auto v = ((Handle*)hv)->g->getRegmin();
  for (size_t i = 0; i < v.size(); ++i) idx[i] = (int32_t)GI::index_of(v[i]);
  return (int64_t)v.size();
}
// IFT_PCD(Graph&, Cloud::Ptr) — ift.hpp:31-37 — on the reference's own graph
REF_API int64_t ref_graph_ift(void* hv, const float* seeds_xyz, int64_t ns, uint32_t* cost, uint32_t* root,
                              uint32_t* mst_a, uint32_t* mst_b, double* mst_w, int64_t cap) {
  global::Cloud::Ptr seeds(new global::Cloud());
  for (int64_t i = 0; i < ns; ++i)
    seeds->push_back(global::Point(seeds_xyz[3 * i], seeds_xyz[3 * i + 1], seeds_xyz[3 * i + 2]));
  Quiet q;
  IFT_PCD ift(*((Handle*)hv)->g, seeds);
  typedef global::graph_v GV;
  GV gv = ift.getGraph();
  int64_t i = 0;
  for (auto v : boost::make_iterator_range(boost::vertices(gv))) cost[i++] = gv[v].cost;
  for (const auto& kv : ift.getRoots()) root[GV::index_of(kv.first)] = (uint32_t)GV::index_of(kv.second);
  int64_t m = 0;
  for (const auto& kv : ift.getMST()) {
    if (m < cap) {
      mst_a[m] = (uint32_t)GV::index_of(kv.first.first);
      mst_b[m] = (uint32_t)GV::index_of(kv.first.second);
      mst_w[m] = kv.second;
    }
    ++m;
  }
  return m;
}

// ---- GraCo (graco.hpp): `GraCo gc(cloud_t, 5.0); gc.morph_erode(cloud_g)` is what the live main()
// runs (main.cpp:87-89).  pts4: n x {x, y, z, bits of 0xAARRGGBB} ----
namespace {
inline uint32_t pack_rgb(const global::PointT& p) { return ((uint32_t)p.r << 16) | ((uint32_t)p.g << 8) | p.b; }  // 0x00RRGGBB
}
REF_API void* ref_graco_create(const float* pts4, int64_t n, double overlap) {
  global::CloudT::Ptr cloud(new global::CloudT());
  for (int64_t i = 0; i < n; ++i) {
    global::PointT p;
    p.x = pts4[4 * i];
    p.y = pts4[4 * i + 1];
    p.z = pts4[4 * i + 2];
    uint32_t w;
    memcpy(&w, pts4 + 4 * i + 3, 4);
    p.r = (w >> 16) & 0xFF;
    p.g = (w >> 8) & 0xFF;
    p.b = w & 0xFF;
    cloud->push_back(p);
  }
  Quiet q;
  return new GraCo(cloud, overlap);
}
REF_API void ref_graco_destroy(void* h) { delete (GraCo*)h; }
REF_API double ref_graco_resolution(void*) {
  return pcl::octree::OctreePointCloudAdjacency<global::PointT>::last_resolution();
}
static int64_t graco_out(const global::CloudT::Ptr& c, float* xyz, uint32_t* rgb) {
  for (size_t i = 0; i < c->size() && rgb; ++i) {
    if (xyz) {
      xyz[3 * i] = c->points[i].x;
      xyz[3 * i + 1] = c->points[i].y;
      xyz[3 * i + 2] = c->points[i].z;
    }
    rgb[i] = pack_rgb(c->points[i]);
  }
  return (int64_t)c->size();
}
// getCloud() — graco.hpp:176-179
REF_API int64_t ref_graco_cloud(void* h, float* xyz, uint32_t* rgb) { return graco_out(((GraCo*)h)->getCloud(), xyz, rgb); }
// op: 3 dilate, 4 erode (enum MORPH_COLOR, graco.hpp:18)
REF_API int64_t ref_graco_morph(void* h, int op, uint32_t* rgb) {
  global::CloudT::Ptr c(new global::CloudT());
  Quiet q;
  if (op == 3) c = ((GraCo*)h)->morph_dilate(c);
  else if (op == 4) c = ((GraCo*)h)->morph_erode(c);
  else return -1;
  return graco_out(c, nullptr, rgb);
}
