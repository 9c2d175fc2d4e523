// The calls of the reference's pipeline block (main.cpp:97-123) made through the adapters —
// same class names, same method names — on a cloud read from a file; every result is dumped so
// that tests/test_cpp_adapters.py can compare it with what the REFERENCE's own classes
// (graph.hpp / ift.hpp compiled unmodified, tests/golden/ref_graph.json) produced on that cloud.
//
//   adapter_vs_ref <cloud.f32 (n x {x,y,z,intensity})> <overlap> <seed vertex ids .i64> <out dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "iftwt_adapters.hpp"

template <class T>
static std::vector<T> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  std::vector<T> v((std::size_t)f.tellg() / sizeof(T));
  f.seekg(0);
  f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T>
static void dump(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}
static std::vector<float> intensities(const global::CloudI::Ptr& c) {
  std::vector<float> v(c->size());
  for (std::size_t i = 0; i < v.size(); ++i) v[i] = c->points[i].intensity;
  return v;
}

int main(int argc, char** argv) {
  if (argc != 5) return 2;
  const std::vector<float> raw = slurp<float>(argv[1]);
  const double overlap = std::atof(argv[2]);
  const std::vector<long long> seed_ids = slurp<long long>(argv[3]);
  const std::string out = argv[4];
  global::CloudI::Ptr cloud_i(new global::CloudI());
  for (std::size_t i = 0; i + 3 < raw.size(); i += 4) {
    global::PointI p;
    p.x = raw[i]; p.y = raw[i + 1]; p.z = raw[i + 2]; p.intensity = raw[i + 3];
    cloud_i->push_back(p);
  }
  try {
    Graph gg(cloud_i, overlap);                                  // main.cpp:97
    global::CloudI::Ptr cloud_v = gg.getCloud();                 // graph.hpp:177
    std::vector<float> cen(4 * cloud_v->size());
    for (std::size_t i = 0; i < cloud_v->size(); ++i) {
      cen[4 * i] = cloud_v->points[i].x; cen[4 * i + 1] = cloud_v->points[i].y;
      cen[4 * i + 2] = cloud_v->points[i].z; cen[4 * i + 3] = cloud_v->points[i].intensity;
    }
    dump(out + "/centres.f32", cen);
    iftwt::CsrGraph g = gg.getAdjacencyList();                   // graph.hpp:63
    dump(out + "/off.u32", g.off);
    dump(out + "/adj.u32", g.adj);
    std::vector<double> res(1, gg.voxelSize());
    dump(out + "/resolution.f64", res);
    global::CloudI::Ptr c(new global::CloudI());
    dump(out + "/morph1.f32", intensities(gg.morph_bin_erode(c)));
    dump(out + "/morph3.f32", intensities(gg.morph_dilate(c)));
    dump(out + "/morph4.f32", intensities(gg.morph_erode(c)));
    global::CloudI::Ptr cloud_g = gg.morph_gradient();           // main.cpp:99
    dump(out + "/morph5.f32", intensities(cloud_g));
    dump(out + "/regmin.u32", gg.getRegmin());                   // graph.hpp:209
    global::Cloud::Ptr seeds(new global::Cloud());               // stands for reg.getCentroidsCloud(), main.cpp:121
    for (long long s : seed_ids) {
      global::Point p;
      p.x = cloud_v->points[(std::size_t)s].x; p.y = cloud_v->points[(std::size_t)s].y; p.z = cloud_v->points[(std::size_t)s].z;
      seeds->push_back(p);
    }
    IFT_PCD ift(gg, seeds);                                      // main.cpp:123
    dump(out + "/cost.u32", ift.getCosts());
    dump(out + "/roots.u32", ift.getRoots());                    // ift.hpp:48
    gg.pc_to_g(cloud_g);                                         // main.cpp:100
    dump(out + "/after_pc_to_g.f32", intensities(gg.getCloud()));
    dump(out + "/erode_gradient.f32", intensities(gg.morph_erode(c)));   // = morph_erode_gradient, graph.hpp:203
  } catch (const std::exception& e) {
    std::fprintf(stderr, "adapter_vs_ref: %s\n", e.what());
    return 1;
  }
  std::puts("adapter ok");
  return 0;
}
