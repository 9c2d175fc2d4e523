// PLY reader/writer of include/iftwt_io.hpp: the PCL header layout of the reference's
// cmake-build-debug/test.ply (vertex element + camera element), ASCII and binary round trips.
#include <cmath>
#include <cstdio>
#include <fstream>

#include "iftwt_io.hpp"

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : ".";
  {  // what PCL writes: x y z intensity + a camera element the reader must skip
    std::ofstream f(dir + "/pcl.ply");
    f << "ply\nformat ascii 1.0\ncomment PCL generated\nelement vertex 3\nproperty float x\nproperty float y\n"
         "property float z\nproperty float intensity\nelement camera 1\nproperty float view_px\n"
         "property float view_py\nproperty int viewportx\nend_header\n"
         "0.2877613 0.24348953 -1.0507743 0\n1 2 3 17\n-4.5 5.25 6.125 255\n0 0 0 640\n";
  }
  global::CloudI c;
  CHECK(iftwt::io::loadPLYFile(dir + "/pcl.ply", c) == 0);
  CHECK(c.size() == 3 && c.width == 3 && c.height == 1);
  CHECK(c.points[0].x == 0.2877613f && c.points[0].z == -1.0507743f && c.points[0].intensity == 0.f);
  CHECK(c.points[1].intensity == 17.f && c.points[2].y == 5.25f && c.points[2].intensity == 255.f);
  for (int binary = 0; binary < 2; ++binary) {
    CHECK(iftwt::io::savePLYFile(dir + "/rt.ply", c, binary != 0) == 0);
    global::CloudI d;
    CHECK(iftwt::io::loadPLYFile(dir + "/rt.ply", d) == 0);
    CHECK(d.size() == c.size());
    for (std::size_t i = 0; i < c.size(); ++i)
      CHECK(d.points[i].x == c.points[i].x && d.points[i].y == c.points[i].y && d.points[i].z == c.points[i].z &&
            d.points[i].intensity == c.points[i].intensity);
  }
  // colour clouds (main.cpp:81): separate channels, packed rgb, and grey conversion on an XYZI load
  global::CloudT t;
  global::PointT p;
  p.x = 1; p.y = 2; p.z = 3; p.r = 200; p.g = 100; p.b = 50;
  t.push_back(p);
  for (int binary = 0; binary < 2; ++binary) {
    CHECK(iftwt::io::savePLYFile(dir + "/rgb.ply", t, binary != 0) == 0);
    global::CloudT u;
    CHECK(iftwt::io::loadPLYFile(dir + "/rgb.ply", u) == 0);
    CHECK(u.size() == 1 && u.points[0].r == 200 && u.points[0].g == 100 && u.points[0].b == 50);
    global::CloudI g;
    CHECK(iftwt::io::loadPLYFile(dir + "/rgb.ply", g) == 0);
    CHECK(g.points[0].intensity == (float)(int)(0.299 * 200 + 0.587 * 100 + 0.114 * 50 + 0.5));
  }
  CHECK(iftwt::io::loadPLYFile(dir + "/does_not_exist.ply", c) == -1);   // main.cpp:83-85
  std::printf("io ok\n");
  return 0;
}
