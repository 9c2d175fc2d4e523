// The reference's pipeline block (main.cpp:91-129), line for line, on the adapters of
// include/iftwt_adapters.hpp.  Build:
//   g++ -std=c++14 -Iinclude examples/main_pipeline.cpp -Liftwt_rg_b200/lib -liftwt
//       -Wl,-rpath,$PWD/iftwt_rg_b200/lib -o main_pipeline
// Input: a .ply or .pcd file (include/iftwt_io.hpp), a raw little-endian float32 file of {x,y,z,intensity} records, or a small built-in
// scene when no file is given (the reference's ../cloud/lucy_gray.ply is absent from its tree).
#include <cstdio>
#include <cstdlib>
#include <map>

#include "iftwt_adapters.hpp"
#include "iftwt_io.hpp"

static global::CloudI::Ptr load_or_make(const char* path) {
  global::CloudI::Ptr cloud(new global::CloudI());
  const std::string ext = path && std::string(path).size() > 4 ? std::string(path).substr(std::string(path).size() - 4) : "";
  if (ext == ".ply" || ext == ".pcd") {
    iftwt::io::load(path, *cloud);                                               // main.cpp:92 (loadPLYFile / loadPCDFile)
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
  // a staircase of eight steps: sixteen planar patches of 800 - 2,000 voxels each, so that the reference's
  // own region-growing parameters (main.cpp:112-119: 50 / 10000 / 50 / 3 degrees / 1.0) yield seeds
  unsigned s = 12345u;
  auto u = [&s]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
  for (int i = 0; i < 60000; ++i) {
    global::PointI p;
    const int step = i % 8;
    if ((i / 8) % 3 != 2) {  // tread, dark
      p.x = 0.5f * ((float)step + u()); p.y = 2.f * u(); p.z = 0.25f * (float)step + 0.001f * (u() - 0.5f);
      p.intensity = (float)(40 + 10 * step + (int)(4.f * u()));
    } else {                 // riser, bright
      p.x = 0.5f * (float)step + 0.001f * (u() - 0.5f); p.y = 2.f * u(); p.z = 0.25f * ((float)step - u());
      p.intensity = (float)(200 - 10 * step + (int)(4.f * u()));
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
    reg.setMinClusterSize(50);
    reg.setMaxClusterSize(10000);
    reg.setNumberOfNeighbours(50);
    reg.setSmoothnessThreshold(3.0f / 180.0f * 3.14159265f);
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
