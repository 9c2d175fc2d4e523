// Checks of the stand-in headers under oracle/shim that the reference's graph.hpp / ift.hpp are compiled
// against (oracle/shim/README.md): the container semantics the reference's results depend on — setS
// iteration in address (= creation) order, no parallel edges, one entry for a self loop, copies that keep
// the order — the exact (d2, index) k-NN order, and the voxel set / centres / closed 26-adjacency of the
// octree stand-in on a hand-checkable cloud.  Prints "shim ok".
#include <cstdio>
#include <vector>

#include <boost/graph/adjacency_list.hpp>
#include <pcl/octree/octree_pointcloud_adjacency.h>
#include <pcl/search/kdtree.h>

#define CHECK(x)                                               \
  do {                                                         \
    if (!(x)) {                                                \
      std::printf("FAILED line %d: %s\n", __LINE__, #x);       \
      return 1;                                                \
    }                                                          \
  } while (0)

typedef boost::adjacency_list<boost::setS, boost::setS, boost::undirectedS, int, float> G;

int main() {
  {  // containers
    G g;
    std::vector<G::vertex_descriptor> v;
    for (int i = 0; i < 100; ++i) v.push_back(boost::add_vertex(i * 10, g));
    int i = 0;
    for (auto u : boost::make_iterator_range(boost::vertices(g))) CHECK(g[u] == 10 * i++);  // creation order
    CHECK(boost::add_edge(v[5], v[3], 1.5f, g).second);
    CHECK(boost::add_edge(v[5], v[9], 2.5f, g).second);
    CHECK(!boost::add_edge(v[9], v[5], 9.f, g).second);  // parallel edge rejected (setS), property kept
    CHECK(boost::add_edge(v[5], v[5], 0.f, g).second);   // self loop
    CHECK(boost::add_edge(v[5], v[1], 0.f, g).second);
    std::vector<int> nb;
    for (auto t : boost::make_iterator_range(boost::adjacent_vertices(v[5], g))) nb.push_back(g[t]);
    CHECK((nb == std::vector<int>{10, 30, 50, 90}));  // ascending target address, the self loop once
    nb.clear();
    for (auto t : boost::make_iterator_range(boost::adjacent_vertices(v[9], g))) nb.push_back(g[t]);
    CHECK((nb == std::vector<int>{50}));
    CHECK(boost::num_vertices(g) == 100 && boost::num_edges(g) == 4);
    int e = 0;
    const int src[4] = {50, 50, 50, 50}, dst[4] = {30, 90, 50, 10};
    BGL_FORALL_EDGES(ed, g, G) {  // insertion order
      CHECK(g[boost::source(ed, g)] == src[e] && g[boost::target(ed, g)] == dst[e]);
      ++e;
    }
    CHECK(g[boost::add_edge(v[5], v[9], g).first] == 2.5f);
    G h = g;  // copy: same order, same properties, its own nodes
    i = 0;
    for (auto u : boost::make_iterator_range(boost::vertices(h))) {
      CHECK(h[u] == 10 * i && G::index_of(u) == (std::size_t)i);
      ++i;
    }
    CHECK(*boost::vertices(h).first != *boost::vertices(g).first);
    h.clear();
    CHECK(boost::num_vertices(h) == 0 && boost::num_vertices(g) == 100);
  }
  {  // k-NN: canonical (d2, index) order, ties by index, the query point itself first
    pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
    const float xs[7] = {0.f, 1.f, -1.f, 0.f, 0.f, 2.f, 1.f};
    const float ys[7] = {0.f, 0.f, 0.f, 1.f, -1.f, 0.f, 0.f};
    for (int i = 0; i < 7; ++i) {
      pcl::PointXYZI p;
      p.x = xs[i];
      p.y = ys[i];
      c->push_back(p);
    }
    pcl::search::KdTree<pcl::PointXYZI> t;
    t.setInputCloud(c);
    std::vector<int> idx;
    std::vector<float> d2;
    CHECK(t.nearestKSearch(0, 6, idx, d2) == 6);
    CHECK((idx == std::vector<int>{0, 1, 2, 3, 4, 6}));  // four ties at d2 = 1, point 6 duplicates point 1
    CHECK(d2[0] == 0.f && d2[1] == 1.f && d2[5] == 1.f);
    CHECK(t.nearestKSearch(5, 9, idx, d2) == 7);          // fewer points than k
    CHECK(idx[0] == 5 && idx[1] == 1 && idx[2] == 6);
  }
  {  // octree stand-in: resolution 1, points in voxels (0,0,0) (1,0,0) (1,1,1) (3,3,3) of a 4^3 key range
    pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
    const float pts[5][3] = {{0.1f, 0.2f, 0.3f}, {1.5f, 0.5f, 0.5f}, {1.2f, 1.7f, 1.1f}, {3.9f, 3.9f, 3.9f}, {0.4f, 0.1f, 0.2f}};
    for (auto& q : pts) {
      pcl::PointXYZI p;
      p.x = q[0];
      p.y = q[1];
      p.z = q[2];
      c->push_back(p);
    }
    pcl::octree::OctreePointCloudAdjacency<pcl::PointXYZI> o(1.0);
    o.setInputCloud(c);
    o.addPointsFromInputCloud();
    CHECK(o.size() == 4 && o.getTreeDepth() == 2);
    boost::adjacency_list<boost::setS, boost::setS, boost::undirectedS, pcl::PointXYZI, float> g;
    o.computeVoxelAdjacencyGraph(g);
    CHECK(boost::num_vertices(g) == 4);
    std::vector<std::vector<int>> nb;
    std::vector<pcl::PointXYZI> cen;
    for (auto u : boost::make_iterator_range(boost::vertices(g))) {
      cen.push_back(g[u]);
      nb.emplace_back();
      for (auto t : boost::make_iterator_range(boost::adjacent_vertices(u, g)))
        nb.back().push_back((int)decltype(g)::index_of(t));
    }
    // extent 3.8 -> 4 voxels per axis, box padded by 0.1 on both sides: min' = 0.0 (0.1 - 0.1) in x
    CHECK(std::fabs(cen[0].x - 0.5f) < 1e-6f && std::fabs(cen[3].x - 3.5f) < 1e-6f);
    // Morton order (x most significant): (0,0,0), (1,0,0), (1,1,1), (3,3,3); closed neighbourhoods
    CHECK((nb[0] == std::vector<int>{0, 1, 2}));
    CHECK((nb[1] == std::vector<int>{0, 1, 2}));
    CHECK((nb[2] == std::vector<int>{0, 1, 2}));
    CHECK((nb[3] == std::vector<int>{3}));
  }
  std::puts("shim ok");
  return 0;
}
