// iftwt_io.hpp — minimal PLY and PCD readers / writers for the clouds the reference moves around
// (pcl::io::loadPLYFile<PointXYZI> at main.cpp:92, loadPLYFile<PointXYZRGB> at main.cpp:81,
// savePLYFile at main.cpp:128-129; BASELINE config 1 names cloud/lucy.pcd), so the pipeline runs
// without PCL.  Header-only C++14.
//
// PCD (the format of pcl::io::loadPCDFile / savePCDFile, "# .PCD v0.7"): header keys VERSION, FIELDS,
// SIZE, TYPE, COUNT, WIDTH, HEIGHT, VIEWPOINT, POINTS, DATA; DATA ascii, binary (array of records)
// and binary_compressed (LZF over a structure-of-arrays image of the records) are read; ascii and
// binary are written, with the field lists PCL writes for PointXYZI (`x y z intensity`) and
// PointXYZRGB (`x y z rgb`, the colour packed in a 4-byte field).
//
// Reads `format ascii 1.0` and `format binary_little_endian 1.0`, the `vertex` element with
// any scalar properties (x y z required; `intensity` / `scalar_intensity`; `red green blue` or
// packed `rgb`), skipping other elements such as the `camera` element PCL appends (see the
// header of the reference's cmake-build-debug/test.ply).  Writes what PCL writes: ASCII or
// binary vertex list `x y z intensity` (XYZI) or `x y z red green blue` (XYZRGB).
#ifndef IFTWT_IO_HPP
#define IFTWT_IO_HPP

#include <cstdint>
#include <cstdlib>
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
      if (type_size(p.type) == 0 || (p.is_list && type_size(p.count_type) == 0)) return -1;  // unknown property type
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

// ---------------------------------------------------------------------------------------------
// PCD
// ---------------------------------------------------------------------------------------------
namespace detail {
struct PcdField {
  std::string name;
  int size = 4;
  char type = 'F';
  int count = 1;
  std::size_t offset = 0;  // byte offset inside a record
};
// LZF (the codec of PCL's binary_compressed): a control byte < 32 starts a literal run of ctrl + 1
// bytes; otherwise it is a back reference of length (ctrl >> 5) + 2 (one more length byte follows
// when ctrl >> 5 == 7) at distance ((ctrl & 31) << 8 | next byte) + 1.
inline bool lzf_decompress(const unsigned char* in, std::size_t in_len, unsigned char* out, std::size_t out_len) {
  std::size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const std::size_t run = ctrl + 1;
      if (ip + run > in_len || op + run > out_len) return false;
      std::memcpy(out + op, in + ip, run);
      ip += run;
      op += run;
    } else {
      std::size_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= in_len) return false;
        len += in[ip++];
      }
      if (ip >= in_len) return false;
      const std::size_t dist = ((std::size_t)(ctrl & 31) << 8 | in[ip++]) + 1;
      len += 2;
      if (dist > op || op + len > out_len) return false;
      for (std::size_t i = 0; i < len; ++i, ++op) out[op] = out[op - dist];  // may overlap: byte by byte
    }
  }
  return op == out_len;
}
inline double pcd_value(const unsigned char* p, const PcdField& f) {
  switch (f.type) {
    case 'F':
      if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
      if (f.size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      return 0;
    case 'U':
      if (f.size == 1) return *p;
      if (f.size == 2) { std::uint16_t v; std::memcpy(&v, p, 2); return v; }
      if (f.size == 4) { std::uint32_t v; std::memcpy(&v, p, 4); return v; }
      if (f.size == 8) { std::uint64_t v; std::memcpy(&v, p, 8); return (double)v; }
      return 0;
    default:
      if (f.size == 1) return *(const std::int8_t*)p;
      if (f.size == 2) { std::int16_t v; std::memcpy(&v, p, 2); return v; }
      if (f.size == 4) { std::int32_t v; std::memcpy(&v, p, 4); return v; }
      if (f.size == 8) { std::int64_t v; std::memcpy(&v, p, 8); return (double)v; }
      return 0;
  }
}
inline void pcd_assign(const PcdField& f, const unsigned char* p, Row& row) {
  const std::string& n = f.name;
  if (n == "rgb" || n == "rgba") {  // 4 bytes 0xAARRGGBB whatever the declared type
    if (f.size != 4) return;
    std::uint32_t bits;
    std::memcpy(&bits, p, 4);
    row.r = (bits >> 16) & 255; row.g = (bits >> 8) & 255; row.b = bits & 255;
    row.has_rgb = true;
    return;
  }
  const double v = pcd_value(p, f);
  if (n == "x") row.x = v;
  else if (n == "y") row.y = v;
  else if (n == "z") row.z = v;
  else if (n == "intensity") { row.intensity = v; row.has_i = true; }
}
// one ascii token -> the bytes of the field (so that ascii and binary share pcd_assign)
inline void pcd_parse_token(const char* tok, const PcdField& f, unsigned char* out) {
  if (f.type == 'F') {
    const double v = std::strtod(tok, nullptr);  // accepts nan / inf, unlike operator>>
    if (f.size == 4) { float x = (float)v; std::memcpy(out, &x, 4); }
    else if (f.size == 8) std::memcpy(out, &v, 8);
  } else if (f.type == 'U') {
    const unsigned long long v = std::strtoull(tok, nullptr, 10);
    std::memcpy(out, &v, (std::size_t)f.size);  // little endian
  } else {
    const long long v = std::strtoll(tok, nullptr, 10);
    std::memcpy(out, &v, (std::size_t)f.size);
  }
}
template <class F>
int read_pcd(const std::string& path, F&& emit) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::vector<PcdField> fields;
  std::size_t width = 0, height = 1, points = 0;
  bool have_points = false;
  std::string data, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS" || key == "COLUMNS") {
      std::string n;
      while (ss >> n) {
        PcdField fd;
        fd.name = n;
        fields.push_back(fd);
      }
    } else if (key == "SIZE") {
      for (auto& fd : fields) if (!(ss >> fd.size)) return -1;
    } else if (key == "TYPE") {
      for (auto& fd : fields) if (!(ss >> fd.type)) return -1;
    } else if (key == "COUNT") {
      for (auto& fd : fields) if (!(ss >> fd.count)) return -1;
    } else if (key == "WIDTH") {
      ss >> width;
    } else if (key == "HEIGHT") {
      ss >> height;
    } else if (key == "POINTS") {
      ss >> points;
      have_points = true;
    } else if (key == "DATA") {
      ss >> data;
      break;
    }  // VERSION, VIEWPOINT: nothing to keep
  }
  if (fields.empty() || data.empty()) return -1;
  if (!have_points) points = width * height;
  std::size_t rec = 0;
  for (auto& fd : fields) {
    if (fd.size != 1 && fd.size != 2 && fd.size != 4 && fd.size != 8) return -1;
    if (fd.type != 'F' && fd.type != 'U' && fd.type != 'I') return -1;
    if (fd.count < 0) return -1;
    fd.offset = rec;
    rec += (std::size_t)fd.size * (std::size_t)fd.count;
  }
  if (rec == 0) return -1;
  auto emit_record = [&](const unsigned char* r) {
    Row row;
    for (const auto& fd : fields)
      if (fd.count == 1) pcd_assign(fd, r + fd.offset, row);
    emit(row);
  };
  if (data == "ascii") {
    std::vector<unsigned char> r(rec);
    for (std::size_t i = 0; i < points; ++i) {
      if (!std::getline(f, line)) return -1;
      std::istringstream ss(line);
      std::string tok;
      for (const auto& fd : fields)
        for (int c = 0; c < fd.count; ++c) {
          if (!(ss >> tok)) return -1;
          pcd_parse_token(tok.c_str(), fd, r.data() + fd.offset + (std::size_t)c * fd.size);
        }
      emit_record(r.data());
    }
    return 0;
  }
  if (data == "binary") {
    std::vector<unsigned char> buf(rec * 4096);
    for (std::size_t done = 0; done < points;) {
      const std::size_t m = points - done < 4096 ? points - done : 4096;
      f.read((char*)buf.data(), (std::streamsize)(m * rec));
      if (!f) return -1;
      for (std::size_t i = 0; i < m; ++i) emit_record(buf.data() + i * rec);
      done += m;
    }
    return 0;
  }
  if (data == "binary_compressed") {
    std::uint32_t csize = 0, usize = 0;
    f.read((char*)&csize, 4);
    f.read((char*)&usize, 4);
    if (!f || (std::size_t)usize != rec * points) return -1;
    std::vector<unsigned char> cbuf(csize), soa(usize), r(rec);
    f.read((char*)cbuf.data(), csize);
    if (!f || !lzf_decompress(cbuf.data(), csize, soa.data(), usize)) return -1;
    // the image is field after field: all x, then all y, ...
    for (std::size_t i = 0; i < points; ++i) {
      for (const auto& fd : fields) {
        const std::size_t fsz = (std::size_t)fd.size * (std::size_t)fd.count;
        std::memcpy(r.data() + fd.offset, soa.data() + fd.offset * points + i * fsz, fsz);
      }
      emit_record(r.data());
    }
    return 0;
  }
  return -1;
}
inline void pcd_header(std::ostream& f, const char* fields, const char* sizes, const char* types, const char* counts,
                       std::size_t n, bool binary) {
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS " << fields << "\nSIZE " << sizes
    << "\nTYPE " << types << "\nCOUNT " << counts << "\nWIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS "
    << n << "\nDATA " << (binary ? "binary" : "ascii") << "\n";
}
}  // namespace detail

// pcl::io::loadPCDFile<pcl::PointXYZI>: 0 on success, -1 on failure
inline int loadPCDFile(const std::string& path, global::CloudI& cloud) {
  cloud.points.clear();
  int rc = detail::read_pcd(path, [&](const detail::Row& r) {
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
// pcl::io::loadPCDFile<pcl::PointXYZRGB>
inline int loadPCDFile(const std::string& path, global::CloudT& cloud) {
  cloud.points.clear();
  int rc = detail::read_pcd(path, [&](const detail::Row& r) {
    global::PointT p;
    p.x = (float)r.x; p.y = (float)r.y; p.z = (float)r.z;
    p.r = (std::uint8_t)r.r; p.g = (std::uint8_t)r.g; p.b = (std::uint8_t)r.b;
    cloud.points.push_back(p);
  });
  cloud.width = (std::uint32_t)cloud.points.size();
  cloud.height = 1;
  return rc;
}
// pcl::io::savePCDFile(file, cloud, binary_mode = false)
inline int savePCDFile(const std::string& path, const global::CloudI& cloud, bool binary = false) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  detail::pcd_header(f, "x y z intensity", "4 4 4 4", "F F F F", "1 1 1 1", cloud.points.size(), binary);
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
inline int savePCDFile(const std::string& path, const global::CloudT& cloud, bool binary = false) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  // the colour as one unsigned 4-byte field (what PCL >= 1.9 writes; PCL 1.8 declares the same four
  // bytes as F — the reader takes either)
  detail::pcd_header(f, "x y z rgb", "4 4 4 4", "F F F U", "1 1 1 1", cloud.points.size(), binary);
  f.precision(9);
  for (const auto& p : cloud.points) {
    const std::uint32_t rgb = ((std::uint32_t)p.r << 16) | ((std::uint32_t)p.g << 8) | (std::uint32_t)p.b;
    if (binary) {
      float v[3] = {p.x, p.y, p.z};
      f.write((const char*)v, 12);
      f.write((const char*)&rgb, 4);
    } else {
      f << p.x << ' ' << p.y << ' ' << p.z << ' ' << rgb << '\n';
    }
  }
  return f ? 0 : -1;
}
inline int savePCDFileASCII(const std::string& path, const global::CloudI& c) { return savePCDFile(path, c, false); }
inline int savePCDFileBinary(const std::string& path, const global::CloudI& c) { return savePCDFile(path, c, true); }
inline int savePCDFileASCII(const std::string& path, const global::CloudT& c) { return savePCDFile(path, c, false); }
inline int savePCDFileBinary(const std::string& path, const global::CloudT& c) { return savePCDFile(path, c, true); }

// by extension: .pcd or .ply (pcl::io::load)
template <class CloudType>
inline int load(const std::string& path, CloudType& cloud) {
  const std::size_t n = path.size();
  if (n > 4 && (path.substr(n - 4) == ".pcd" || path.substr(n - 4) == ".PCD")) return loadPCDFile(path, cloud);
  return loadPLYFile(path, cloud);
}

}  // namespace io
}  // namespace iftwt
#endif  // IFTWT_IO_HPP
