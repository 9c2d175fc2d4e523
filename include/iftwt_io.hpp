// iftwt_io.hpp — minimal PLY reader / writer for the clouds the reference moves around
// (pcl::io::loadPLYFile<PointXYZI> at main.cpp:92, loadPLYFile<PointXYZRGB> at main.cpp:81,
// savePLYFile at main.cpp:128-129), so the pipeline runs without PCL.  Header-only C++14.
//
// Reads `format ascii 1.0` and `format binary_little_endian 1.0`, the `vertex` element with
// any scalar properties (x y z required; `intensity` / `scalar_intensity`; `red green blue` or
// packed `rgb`), skipping other elements such as the `camera` element PCL appends (see the
// header of the reference's cmake-build-debug/test.ply).  Writes what PCL writes: ASCII or
// binary vertex list `x y z intensity` (XYZI) or `x y z red green blue` (XYZRGB).
#ifndef IFTWT_IO_HPP
#define IFTWT_IO_HPP

#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "iftwt_adapters.hpp"

namespace iftwt {
namespace io {

namespace detail {
struct Prop {
  std::string name, type;
  bool is_list = false;
  std::string count_type;
};
struct Element {
  std::string name;
  std::size_t count = 0;
  std::vector<Prop> props;
};
inline int type_size(const std::string& t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 0;
}
inline double read_bin(const unsigned char* p, const std::string& t) {
  if (t == "char" || t == "int8") return (double)*(const std::int8_t*)p;
  if (t == "uchar" || t == "uint8") return (double)*p;
  if (t == "short" || t == "int16") { std::int16_t v; std::memcpy(&v, p, 2); return v; }
  if (t == "ushort" || t == "uint16") { std::uint16_t v; std::memcpy(&v, p, 2); return v; }
  if (t == "int" || t == "int32") { std::int32_t v; std::memcpy(&v, p, 4); return v; }
  if (t == "uint" || t == "uint32") { std::uint32_t v; std::memcpy(&v, p, 4); return v; }
  if (t == "float" || t == "float32") { float v; std::memcpy(&v, p, 4); return v; }
  double v; std::memcpy(&v, p, 8); return v;
}
struct Row {
  double x = 0, y = 0, z = 0, intensity = 0, r = 0, g = 0, b = 0;
  bool has_i = false, has_rgb = false;
};
// the packed `rgb` property of PCL is a float whose bits are 0x00RRGGBB
inline void unpack_rgb(double v, const std::string& type, Row& row) {
  std::uint32_t bits;
  if (type == "float" || type == "float32") { float f = (float)v; std::memcpy(&bits, &f, 4); }
  else bits = (std::uint32_t)v;
  row.r = (bits >> 16) & 255; row.g = (bits >> 8) & 255; row.b = bits & 255;
  row.has_rgb = true;
}
inline void assign(const Prop& p, double v, Row& row) {
  const std::string& n = p.name;
  if (n == "x") row.x = v;
  else if (n == "y") row.y = v;
  else if (n == "z") row.z = v;
  else if (n == "intensity" || n == "scalar_intensity" || n == "scalar_Intensity") { row.intensity = v; row.has_i = true; }
  else if (n == "red" || n == "r") { row.r = v; row.has_rgb = true; }
  else if (n == "green" || n == "g") { row.g = v; row.has_rgb = true; }
  else if (n == "blue" || n == "b") { row.b = v; row.has_rgb = true; }
  else if (n == "rgb" || n == "rgba") unpack_rgb(v, p.type, row);
}
template <class F>
int read_ply(const std::string& path, F&& emit) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::string line;
  if (!std::getline(f, line) || line.substr(0, 3) != "ply") return -1;
  bool ascii = true;
  std::vector<Element> elems;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ss(line);
    std::string tok;
    ss >> tok;
    if (tok == "format") {
      std::string fmt;
      ss >> fmt;
      if (fmt == "ascii") ascii = true;
      else if (fmt == "binary_little_endian") ascii = false;
      else return -1;  // big endian files are not produced by PCL on x86
    } else if (tok == "element") {
      Element e;
      ss >> e.name >> e.count;
      elems.push_back(e);
    } else if (tok == "property" && !elems.empty()) {
      Prop p;
      ss >> p.type;
      if (p.type == "list") {
        p.is_list = true;
        ss >> p.count_type >> p.type;
      }
      ss >> p.name;
      elems.back().props.push_back(p);
    } else if (tok == "end_header") {
      break;
    }
  }
  for (const Element& e : elems) {
    const bool want = e.name == "vertex";
    for (std::size_t i = 0; i < e.count; ++i) {
      Row row;
      if (ascii) {
        if (!std::getline(f, line)) return -1;
        std::istringstream ss(line);
        for (const Prop& p : e.props) {
          if (p.is_list) {
            long cnt = 0;
            ss >> cnt;
            double d;
            for (long c = 0; c < cnt; ++c) ss >> d;
          } else {
            double v = 0;
            ss >> v;
            if (want) assign(p, v, row);
          }
        }
      } else {
        for (const Prop& p : e.props) {
          unsigned char buf[8];
          if (p.is_list) {
            f.read((char*)buf, type_size(p.count_type));
            long cnt = (long)read_bin(buf, p.count_type);
            f.ignore((std::streamsize)cnt * type_size(p.type));
          } else {
            f.read((char*)buf, type_size(p.type));
            if (want) assign(p, read_bin(buf, p.type), row);
          }
        }
        if (!f) return -1;
      }
      if (want) emit(row);
    }
  }
  return 0;
}
}  // namespace detail

// pcl::io::loadPLYFile<pcl::PointXYZI>: 0 on success, -1 on failure (main.cpp:92-95).  A colour
// file without an intensity property gets the BT.601 grey of its colour (cloudGray, main.cpp:55-76).
inline int loadPLYFile(const std::string& path, global::CloudI& cloud) {
  cloud.points.clear();
  int rc = detail::read_ply(path, [&](const detail::Row& r) {
    global::PointI p;
    p.x = (float)r.x; p.y = (float)r.y; p.z = (float)r.z;
    p.intensity = r.has_i ? (float)r.intensity
                          : (r.has_rgb ? (float)(int)(0.299 * r.r + 0.587 * r.g + 0.114 * r.b + 0.5) : 0.f);
    cloud.points.push_back(p);
  });
  cloud.width = (std::uint32_t)cloud.points.size();
  cloud.height = 1;
  return rc;
}
// pcl::io::loadPLYFile<pcl::PointXYZRGB> (main.cpp:81)
inline int loadPLYFile(const std::string& path, global::CloudT& cloud) {
  cloud.points.clear();
  int rc = detail::read_ply(path, [&](const detail::Row& r) {
    global::PointT p;
    p.x = (float)r.x; p.y = (float)r.y; p.z = (float)r.z;
    p.r = (std::uint8_t)r.r; p.g = (std::uint8_t)r.g; p.b = (std::uint8_t)r.b;
    cloud.points.push_back(p);
  });
  cloud.width = (std::uint32_t)cloud.points.size();
  cloud.height = 1;
  return rc;
}
// pcl::io::savePLYFile (main.cpp:128-129)
inline int savePLYFile(const std::string& path, const global::CloudI& cloud, bool binary = false) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  f << "ply\nformat " << (binary ? "binary_little_endian" : "ascii") << " 1.0\ncomment iftwt generated\n"
    << "element vertex " << cloud.points.size()
    << "\nproperty float x\nproperty float y\nproperty float z\nproperty float intensity\nend_header\n";
  f.precision(9);
  for (const auto& p : cloud.points) {
    if (binary) {
      float v[4] = {p.x, p.y, p.z, p.intensity};
      f.write((const char*)v, 16);
    } else {
      f << p.x << ' ' << p.y << ' ' << p.z << ' ' << p.intensity << '\n';
    }
  }
  return f ? 0 : -1;
}
inline int savePLYFile(const std::string& path, const global::CloudT& cloud, bool binary = false) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  f << "ply\nformat " << (binary ? "binary_little_endian" : "ascii") << " 1.0\ncomment iftwt generated\n"
    << "element vertex " << cloud.points.size()
    << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
       "property uchar blue\nend_header\n";
  f.precision(9);
  for (const auto& p : cloud.points) {
    if (binary) {
      float v[3] = {p.x, p.y, p.z};
      unsigned char c[3] = {p.r, p.g, p.b};
      f.write((const char*)v, 12);
      f.write((const char*)c, 3);
    } else {
      f << p.x << ' ' << p.y << ' ' << p.z << ' ' << (int)p.r << ' ' << (int)p.g << ' ' << (int)p.b << '\n';
    }
  }
  return f ? 0 : -1;
}

}  // namespace io
}  // namespace iftwt
#endif  // IFTWT_IO_HPP
