// PLY and PCD readers/writers of include/iftwt_io.hpp: the PCL header layout of the reference's
// cmake-build-debug/test.ply (vertex element + camera element), ASCII and binary round trips.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

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
  {  // an unknown property type is an error, not a zero-byte read
    std::ofstream f(dir + "/bad.ply");
    f << "ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty quux y\nend_header\n";
  }
  CHECK(iftwt::io::loadPLYFile(dir + "/bad.ply", c) == -1);

  // ---- PCD (BASELINE config 1: cloud/lucy.pcd) ----
  {  // the header PCL writes, ascii, with a nan and a multi-count field the reader must skip
    std::ofstream f(dir + "/pcl.pcd");
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity hist\nSIZE 4 4 4 4 4\n"
         "TYPE F F F F F\nCOUNT 1 1 1 1 2\nWIDTH 3\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 3\nDATA ascii\n"
         "0.2877613 0.24348953 -1.0507743 0 9 9\n1 2 3 17 9 9\nnan 5.25 6.125 255 9 9\n";
  }
  global::CloudI pc;
  CHECK(iftwt::io::loadPCDFile(dir + "/pcl.pcd", pc) == 0);
  CHECK(pc.size() == 3 && pc.width == 3 && pc.height == 1);
  CHECK(pc.points[0].x == 0.2877613f && pc.points[0].z == -1.0507743f && pc.points[1].intensity == 17.f);
  CHECK(std::isnan(pc.points[2].x) && pc.points[2].y == 5.25f && pc.points[2].intensity == 255.f);
  pc.points[2].x = -4.5f;
  for (int binary = 0; binary < 2; ++binary) {
    CHECK(iftwt::io::savePCDFile(dir + "/rt.pcd", pc, binary != 0) == 0);
    global::CloudI d;
    CHECK(iftwt::io::load(dir + "/rt.pcd", d) == 0);
    CHECK(d.size() == pc.size());
    for (std::size_t i = 0; i < pc.size(); ++i)
      CHECK(d.points[i].x == pc.points[i].x && d.points[i].y == pc.points[i].y && d.points[i].z == pc.points[i].z &&
            d.points[i].intensity == pc.points[i].intensity);
    CHECK(iftwt::io::savePCDFile(dir + "/rgb.pcd", t, binary != 0) == 0);
    global::CloudT u;
    CHECK(iftwt::io::loadPCDFile(dir + "/rgb.pcd", u) == 0);
    CHECK(u.size() == 1 && u.points[0].x == 1.f && u.points[0].r == 200 && u.points[0].g == 100 && u.points[0].b == 50);
  }
  {  // PCL 1.8 declares the packed colour as F: ascii prints the float whose bits are 0x00RRGGBB
    const std::uint32_t bits = (200u << 16) | (100u << 8) | 50u;
    float as_float;
    std::memcpy(&as_float, &bits, 4);
    std::ofstream f(dir + "/rgbf.pcd");
    f.precision(9);
    f << "VERSION .7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n"
      << "1 2 3 " << as_float << "\n";
  }
  {
    global::CloudT u;
    CHECK(iftwt::io::loadPCDFile(dir + "/rgbf.pcd", u) == 0);
    CHECK(u.size() == 1 && u.points[0].r == 200 && u.points[0].g == 100 && u.points[0].b == 50);
  }
  {  // binary_compressed: LZF over the structure-of-arrays image.  Two points {1,2,3,7} {1,2,3,9}: the x, y, z
     // columns repeat, so the stream below uses a literal run and back references (written by hand from the
     // codec's definition).
    const float col[8] = {1.f, 1.f, 2.f, 2.f, 3.f, 3.f, 7.f, 9.f};
    unsigned char raw[32];
    std::memcpy(raw, col, 32);
    std::vector<unsigned char> z;
    for (int fld = 0; fld < 3; ++fld) {   // 4 literal bytes, then a back reference of length 4 at distance 4
      z.push_back(3);
      for (int b = 0; b < 4; ++b) z.push_back(raw[8 * fld + b]);
      z.push_back((unsigned char)((2u << 5) | 0u));  // len 2 + 2 = 4, distance high bits 0
      z.push_back(3);                                // distance 3 + 1 = 4
    }
    z.push_back(7);                                  // 8 literal bytes: the intensity column
    for (int b = 0; b < 8; ++b) z.push_back(raw[24 + b]);
    std::ofstream f(dir + "/lzf.pcd", std::ios::binary);
    f << "VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\n"
         "DATA binary_compressed\n";
    const std::uint32_t cs = (std::uint32_t)z.size(), us = 32;
    f.write((const char*)&cs, 4);
    f.write((const char*)&us, 4);
    f.write((const char*)z.data(), (std::streamsize)z.size());
  }
  {
    global::CloudI d;
    CHECK(iftwt::io::loadPCDFile(dir + "/lzf.pcd", d) == 0);
    CHECK(d.size() == 2 && d.points[0].x == 1.f && d.points[1].x == 1.f && d.points[0].y == 2.f && d.points[1].z == 3.f);
    CHECK(d.points[0].intensity == 7.f && d.points[1].intensity == 9.f);
  }
  CHECK(iftwt::io::loadPCDFile(dir + "/does_not_exist.pcd", pc) == -1);
  std::printf("io ok\n");
  return 0;
}
