// The reference's live main() (main.cpp:80-89) through the adapters: a coloured cloud, GraCo with
// overlap 5.0, colour erosion; then the grey / Otsu prep feeding Graph.  Exits 0 and prints
// "graco ok" when the erosion never raises r+g+b above the input's closed-neighbourhood minimum.
#include <cstdio>
#include <cstdlib>

#include "iftwt_adapters.hpp"

int main() {
  global::CloudT::Ptr cloud_t(new global::CloudT());
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (float)((s >> 40) & 0xFFFFFF) / 16777216.f;
  };
  for (int i = 0; i < 40000; ++i) {
    global::PointT p;
    p.x = rnd() * 4.f; p.y = rnd() * 4.f; p.z = 0.02f * rnd();
    p.r = (std::uint8_t)(255.f * rnd()); p.g = (std::uint8_t)(p.x * 60.f); p.b = (std::uint8_t)(p.y * 60.f);
    cloud_t->push_back(p);
  }
  try {
    GraCo gc(cloud_t, 5.0);
    global::CloudT::Ptr cloud_g(new global::CloudT());
    gc.morph_erode(cloud_g);
    global::CloudT::Ptr cloud_v = gc.getCloud();
    iftwt::CsrGraph g = gc.getAdjacencyList();
    if (cloud_g->size() != gc.size() || cloud_v->size() != gc.size()) return 2;
    for (std::size_t v = 0; v < gc.size(); ++v) {
      auto sum = [](const global::PointT& p) { return (int)p.r + p.g + p.b; };
      int m = sum(cloud_v->points[v]);
      for (std::uint32_t e = g.off[v]; e < g.off[v + 1]; ++e) m = std::min(m, sum(cloud_v->points[g.adj[e]]));
      if (sum(cloud_g->points[v]) != m) {
        std::printf("vertex %zu: eroded sum %d, neighbourhood minimum %d\n", v, sum(cloud_g->points[v]), m);
        return 3;
      }
    }
    double thr = 0;
    global::CloudI::Ptr cloud_i = cloudRGB2GRAY(cloud_t, &thr);
    if (cloud_i->size() != cloud_t->size() || thr <= 0 || thr >= 255) return 4;
    for (const auto& p : cloud_i->points)
      if (p.intensity != 0.f && p.intensity != 255.f) return 5;
    Graph gg(cloud_i);
    std::printf("graco ok: vertices %zu voxel %.6g otsu %.0f grey-graph vertices %zu\n", gc.size(), gc.voxelSize(), thr,
                gg.size());
  } catch (const iftwt::Error& e) {
    std::printf("iftwt error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
