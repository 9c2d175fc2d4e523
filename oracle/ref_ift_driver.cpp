// ref_ift_driver.cpp — runs the REFERENCE's own IFT (IFT_PCD in ift.hpp, custom_priority_queue
// in globals.hpp, both #included unmodified from /root/reference) on a caller-given graph.
//
// Test infrastructure (like everything under oracle/): nothing under iftwt_rg_b200/ links or
// calls it.  Built by `make -C oracle ref` into oracle/_ref/libref_ift.so against the
// stand-in headers of oracle/shim (README there); used by tests/ to pin the oracle's literal
// restatement (orc_ift_literal_mst) and by tests/golden/make_ref_ift_golden.py to write the fixtures.
//
// What is the reference's here: copy_graph, add_seeds, search_p, compute_watershed,
// find_and_swap, r_info (ift.hpp:5-23, 97-250) and the priority queue with its O(|Q|) remove()
// (globals.hpp:57-81).  What is not: the `Graph` class below is a stand-in with the three
// methods IFT_PCD calls (getAdjacencyList / getCloud / getRegmin, graph.hpp:63, 177, 209) —
// graph.hpp itself needs PCL's octree and VTK — and the containers come from oracle/shim.
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>

#include "globals.hpp"          // /root/reference
#include <pcl/search/kdtree.h>  // oracle/shim

// stand-in for graph.hpp's Graph: holds a graph_i built from caller arrays.  Vertex i of the
// caller is the i-th vertex created, hence (shim arena) the i-th in address order.
class Graph {
 public:
  Graph(const float* pts4, int64_t n, const uint32_t* off, const uint32_t* adj) {
    std::vector<global::graph_i::vertex_descriptor> vd((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      global::PointI p;
      p.x = pts4[4 * i];
      p.y = pts4[4 * i + 1];
      p.z = pts4[4 * i + 2];
      p.intensity = pts4[4 * i + 3];
      vd[(size_t)i] = boost::add_vertex(p, g_);
    }
    for (int64_t i = 0; i < n; ++i)
      for (uint32_t e = off[i]; e < off[i + 1]; ++e)
        if ((int64_t)adj[e] >= i) boost::add_edge(vd[(size_t)i], vd[adj[e]], 0.f, g_);
  }
  global::graph_i getAdjacencyList() { return g_; }  // graph.hpp:63-66 (by value)
  global::CloudI::Ptr getCloud() {                   // graph.hpp:177-180 -> g_to_pc, graph.hpp:15-31
    global::CloudI::Ptr out(new global::CloudI());
    const auto vd_v = boost::vertices(g_);
    for (auto v = vd_v.first; v != vd_v.second; ++v) out->push_back(g_[*v]);
    return out;
  }
  std::vector<global::pcd_vx_descriptor> getRegmin() { return {}; }  // the IFT_PCD(Graph&) path is not driven

 private:
  global::graph_i g_;
};

#include "ift.hpp"  // /root/reference

extern "C" {
// pts4: n x {x, y, z, intensity}; (off, adj): symmetric CSR; seeds_xyz: ns x {x, y, z}.
// cost[n], root[n] (vertex index of the seed whose region conquered the vertex; self when
// unreached).  mst rows {s_root, r_root, weight} in the map's order, at most cap rows are written.
// Returns the number of mst rows.
__attribute__((visibility("default")))
int64_t ref_ift_run(const float* pts4, int64_t n, const uint32_t* off, const uint32_t* adj, const float* seeds_xyz,
                    int64_t ns, uint32_t* cost, uint32_t* root, uint32_t* mst_a, uint32_t* mst_b, double* mst_w,
                    int64_t cap) {
  Graph g(pts4, n, off, adj);
  global::Cloud::Ptr seeds(new global::Cloud());
  for (int64_t i = 0; i < ns; ++i)
    seeds->push_back(global::Point(seeds_xyz[3 * i], seeds_xyz[3 * i + 1], seeds_xyz[3 * i + 2]));
  // the reference prints counts and timings on stdout: park fd 1 on /dev/null while it runs
  fflush(stdout);
  std::cout.flush();
  const int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
  if (devnull >= 0) dup2(devnull, 1);
  IFT_PCD ift(g, seeds);
  fflush(stdout);
  std::cout.flush();
  if (saved >= 0) {
    dup2(saved, 1);
    close(saved);
  }
  if (devnull >= 0) close(devnull);
  typedef global::graph_v GV;
  GV gv = ift.getGraph();  // a copy: same creation order
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
}
