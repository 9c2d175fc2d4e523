// The reference's pipeline block (main.cpp:91-129), line for line, on the adapters of
// include/iftwt_adapters.hpp.  Build:
//   g++ -std=c++14 -Iinclude examples/main_pipeline.cpp -Liftwt_rg_b200/lib -liftwt
//       -Wl,-rpath,$PWD/iftwt_rg_b200/lib -o main_pipeline
// Input: a .ply file (include/iftwt_io.hpp), a raw little-endian float32 file of {x,y,z,intensity} records, or a small built-in
// scene when no file is given (the reference's ../cloud/lucy_gray.ply is absent from its tree).
#include <cstdio>
#include <cstdlib>
#include <map>

#include "iftwt_adapters.hpp"
#include "iftwt_io.hpp"

static global::CloudI::Ptr load_or_make(const char* path) {
  global::CloudI::Ptr cloud(new global::CloudI());
  if (path && std::string(path).size() > 4 && std::string(path).substr(std::string(path).size() - 4) == ".ply") {
    iftwt::io::loadPLYFile(path, *cloud);                                        // main.cpp:92
    return cloud;
  }
  if (path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return cloud;
    float r[4];
    while (std::fread(r, sizeof(float), 4, f) == 4) {
      global::PointI p;
      p.x = r[0]; p.y = r[1]; p.z = r[2]; p.intensity = r[3];
      cloud->push_back(p);
    }
    std::fclose(f);
    return cloud;
  }
  unsigned s = 12345u;
  auto u = [&s]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
  for (int i = 0; i < 60000; ++i) {
    global::PointI p;
    if (i % 2) {  // plane, dark
      p.x = 4.f * u() - 2.f; p.y = 4.f * u() - 2.f; p.z = 0.002f * (u() - 0.5f);
      p.intensity = (float)(40 + (int)(8.f * u()));
    } else {      // sphere, bright
      float z = 2.f * u() - 1.f, a = 6.2831853f * u(), r = std::sqrt(1.f - z * z);
      p.x = r * std::cos(a); p.y = r * std::sin(a); p.z = 1.2f + z;
      p.intensity = (float)(200 + (int)(8.f * u()));
    }
    cloud->push_back(p);
  }
  return cloud;
}

int main(int argc, char** argv) {
  try {
    global::CloudI::Ptr cloud_i = load_or_make(argc > 1 ? argv[1] : nullptr);   // main.cpp:91-95
    if (cloud_i->size() == 0) {
      std::printf("Cloud reading failed.\n");
      return -1;
    }
    Graph gg(cloud_i, 2.5);                                                      // main.cpp:97 (+ overlap)
    global::CloudI::Ptr cloud_g = gg.morph_gradient();                           // main.cpp:99
    gg.pc_to_g(cloud_g);                                                         // main.cpp:100

    iftwt::types::PointCloud<iftwt::types::Normal> normals;                      // main.cpp:102-109
    iftwt::NormalEstimation normal_estimator(gg);
    normal_estimator.setKSearch(30);
    normal_estimator.compute(normals);

    FlatPoint<global::PointI, iftwt::types::Normal> reg(gg);                     // main.cpp:111-119
    reg.setMinClusterSize(20);
    reg.setMaxClusterSize(100000);
    reg.setNumberOfNeighbours(50);
    reg.setSmoothnessThreshold(20.0f / 180.0f * 3.14159265f);
    reg.setCurvatureThreshold(1.0f);
    global::Cloud::Ptr centroid_cloud = reg.getCentroidsCloud();                 // main.cpp:121

    if (centroid_cloud->size() == 0) std::printf("no region-growing seeds: try IFT_PCD ift(gg) (regional minima)\n");
    IFT_PCD ift(gg, centroid_cloud);                                             // main.cpp:123
    Cluster c(ift, 10);                                                          // main.cpp:124

    std::map<std::uint32_t, std::size_t> sizes;
    for (std::uint32_t l : c.getLabels()) sizes[l]++;
    std::printf("points %zu  vertices %zu  voxel %.6g  seeds %zu  regions %zu\n", cloud_i->size(), gg.size(),
                gg.voxelSize(), centroid_cloud->size(), c.regions());
    for (auto& kv : sizes) std::printf("  region %u: %zu vertices\n", kv.first, kv.second);
    return 0;
  } catch (const iftwt::Error& e) {
    std::fprintf(stderr, "iftwt error %d: %s\n", e.code, e.what());
    return 1;
  }
}
