// =============================================================================
// iftwt_oracle.cpp — CPU ORACLE for the IFTWT_RG segmentation hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
// in __graft_entry__.py and the cpu_baseline / --impl reference legs of
// bench.py may load it.  The product (iftwt_rg_b200/csrc, libiftwt.so) never
// links, includes or calls anything in this directory.
//
// PARITY STATUS: pinned by the reference's OWN code for the rows the reference wrote itself,
// unpinned for the PCL routines it calls.  The reference (enemy537/IFTWT_RG) ships no tests
// and no golden vectors, and as a whole it cannot be compiled here (PCL 1.8, Boost.Graph, FLANN,
// Eigen, VTK are not installed; no network).  This file is a dependency-free C++17 restatement of
//   (a) the reference's own loops, cited as  file:line  of /root/reference.  graph.hpp, ift.hpp
//       and globals.hpp DO compile, unmodified, against the small API stand-ins of oracle/shim
//       (oracle/ref_ift_driver.cpp, oracle/ref_graph_driver.cpp -> oracle/_ref/*.so); their
//       outputs — cloud resolution, voxel graph, morphology, pc_to_g, regional minima, IFT costs,
//       root labels and the `mst` map — are committed under tests/golden/ref_*.{npz,json} and this
//       file reproduces every one of them bit for bit (tests/test_ref_ift.py, test_ref_graph.py);
//   (b) the PCL 1.8.1 / FLANN 1.9.1 / Eigen 3.3 routines those loops call (k-NN search, octree
//       voxelisation, normal estimation, region growing), restated from the published algorithms
//       ("[PCL-recall]": not checkable offline; each is one clearly-named function).  These stay
//       UNPINNED by the reference and are cross-checked in tests/ against scipy.spatial.cKDTree,
//       numpy.linalg, scipy.sparse.csgraph and brute-force Python.
//
// Conventions
//   * points are rows of 4 floats {x, y, z, intensity}
//   * all indices are ORIGINAL point / vertex indices
//   * canonical neighbour order is (d2, index) ascending  
//   * FP32 arithmetic is written so that every rounding is explicit; build
//     with  -ffp-contract=off  (no FMA contraction), see oracle/Makefile.
// =============================================================================
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <numeric>
#include <queue>
#include <set>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

typedef int64_t i64;
typedef uint32_t u32;
typedef uint64_t u64;

// -----------------------------------------------------------------------------
// Squared distance exactly as FLANN's L2_Simple<float> accumulates it
// [PCL-recall App. A.1]: result = ((0 + dx*dx) + dy*dy) + dz*dz, all FP32.
// Call sites in the reference: graph.hpp:245, graph.hpp:264, flat_point.hpp:74,
// ift.hpp:135, main.cpp:102-108 (through pcl::search::KdTree).
// -----------------------------------------------------------------------------
inline float dist2(const float* a, const float* b) {
  float dx = a[0] - b[0];
  float dy = a[1] - b[1];
  float dz = a[2] - b[2];
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

struct Cand {
  float d2;
  int idx;
};
inline bool cand_less(const Cand& a, const Cand& b) {
  return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

// -----------------------------------------------------------------------------
// Exact k-NN on a bounding-box k-d tree.  PCL uses a FLANN KDTreeSingleIndex
// (leaf size 15); the *result* of an exact search does not depend on the tree,
// only on the metric and on the tie order, which we make canonical (d2, idx).
// Pruning uses the FP32 box distance, which is a lower bound of the FP32 point
// distance term by term (rounding is monotone), so no true neighbour is lost.
// -----------------------------------------------------------------------------
struct KdTree {
  const float* pts = nullptr;  // n x 4
  i64 n = 0;
  std::vector<int> perm;  // leaf order -> original index
  struct Node {
    float lo[3], hi[3];
    int left, right;  // children or -1
    int begin, end;   // range in perm
  };
  std::vector<Node> nodes;
  static constexpr int LEAF = 12;

  void build(const float* p, i64 n_) {
    pts = p;
    n = n_;
    perm.resize(n);
    std::iota(perm.begin(), perm.end(), 0);
    nodes.clear();
    nodes.reserve((size_t)(2 * n / LEAF + 16));
    if (n > 0) build_rec(0, (int)n);
  }
  int build_rec(int b, int e) {
    Node nd;
    for (int a = 0; a < 3; ++a) {
      nd.lo[a] = FLT_MAX;
      nd.hi[a] = -FLT_MAX;
    }
    for (int i = b; i < e; ++i) {
      const float* q = pts + 4 * (i64)perm[i];
      for (int a = 0; a < 3; ++a) {
        nd.lo[a] = std::min(nd.lo[a], q[a]);
        nd.hi[a] = std::max(nd.hi[a], q[a]);
      }
    }
    nd.begin = b;
    nd.end = e;
    nd.left = nd.right = -1;
    int id = (int)nodes.size();
    nodes.push_back(nd);
    if (e - b > LEAF) {
      int ax = 0;
      float ext = nd.hi[0] - nd.lo[0];
      for (int a = 1; a < 3; ++a)
        if (nd.hi[a] - nd.lo[a] > ext) {
          ext = nd.hi[a] - nd.lo[a];
          ax = a;
        }
      if (ext > 0.f) {
        int m = (b + e) / 2;
        const float* P = pts;
        std::nth_element(perm.begin() + b, perm.begin() + m, perm.begin() + e, [P, ax](int i, int j) {
          float vi = P[4 * (i64)i + ax], vj = P[4 * (i64)j + ax];
          return vi < vj || (vi == vj && i < j);
        });
        int l = build_rec(b, m);
        int r = build_rec(m, e);
        nodes[id].left = l;
        nodes[id].right = r;
      }
    }
    return id;
  }
  static inline float box_d2(const Node& nd, const float* q) {
    float e[3];
    for (int a = 0; a < 3; ++a) {
      float v = 0.f;
      if (q[a] < nd.lo[a]) v = nd.lo[a] - q[a];
      else if (q[a] > nd.hi[a]) v = q[a] - nd.hi[a];
      e[a] = v;
    }
    float r = e[0] * e[0];
    r = r + e[1] * e[1];
    r = r + e[2] * e[2];
    return r;
  }
  // heap = max-heap of the k best (worst on top)
  void search(const float* q, int k, std::vector<Cand>& heap) const {
    heap.clear();
    if (n == 0) return;
    search_rec(0, q, k, heap);
    std::sort(heap.begin(), heap.end(), cand_less);
  }
  void search_rec(int id, const float* q, int k, std::vector<Cand>& heap) const {
    const Node& nd = nodes[id];
    if ((int)heap.size() == k && box_d2(nd, q) > heap.front().d2) return;
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        Cand c;
        c.idx = perm[i];
        c.d2 = dist2(q, pts + 4 * (i64)c.idx);
        if ((int)heap.size() < k) {
          heap.push_back(c);
          std::push_heap(heap.begin(), heap.end(), cand_less);
        } else if (cand_less(c, heap.front())) {
          std::pop_heap(heap.begin(), heap.end(), cand_less);
          heap.back() = c;
          std::push_heap(heap.begin(), heap.end(), cand_less);
        }
      }
      return;
    }
    float dl = box_d2(nodes[nd.left], q), dr = box_d2(nodes[nd.right], q);
    if (dl <= dr) {
      search_rec(nd.left, q, k, heap);
      search_rec(nd.right, q, k, heap);
    } else {
      search_rec(nd.right, q, k, heap);
      search_rec(nd.left, q, k, heap);
    }
  }
};

void knn_run(const float* pts, i64 n, const float* q, i64 m, int k, int32_t* idx, float* d2) {
  KdTree t;
  t.build(pts, n);
#pragma omp parallel
  {
    std::vector<Cand> heap;
    heap.reserve(k + 1);
#pragma omp for schedule(dynamic, 1024)
    for (i64 i = 0; i < m; ++i) {
      t.search(q + 4 * i, k, heap);
      for (int j = 0; j < k; ++j) {
        if (j < (int)heap.size()) {
          idx[i * k + j] = heap[j].idx;
          if (d2) d2[i * k + j] = heap[j].d2;
        } else {
          idx[i * k + j] = -1;
          if (d2) d2[i * k + j] = INFINITY;
        }
      }
    }
  }
}

// -----------------------------------------------------------------------------
// Elementary functions used by the eigen solver, in two flavours:
//   trig_mode 1  ("libm")  std::atan2 / std::cos / std::sin on float, which is
//                what PCL's computeRoots<float> calls                [PCL-recall]
//   trig_mode 0  ("spec")  the same functions evaluated in binary64 by fixed
//                nested series built only from IEEE + - * / and then rounded
//                once to binary32.  Every step is correctly rounded on any
//                IEEE machine, so a GPU restating this spec produces the same
//                bits.  It differs from libm by at most 1 float ulp.
// The spec (also stated in DESIGN.md):
//   atan on [0, tan(pi/8)]:  s = 1/(2N+1); for n = N-1..0: s = 1/(2n+1) - u2*s;
//                            atan(u) = u*s,  N = 22
//   t > tan(pi/8): atan(t) = pi/4 + atan((t-1)/(t+1))
//   cos(x) = c_0, c_n = 1 - (x2/((2n-1)(2n))) * c_{n+1}, c_{13} = 1
//   sin(x) = x*s_1, s_n = 1 - (x2/((2n)(2n+1))) * s_{n+1}, s_{13} = 1
// -----------------------------------------------------------------------------
const double kPi = 3.141592653589793;
const double kTanPi8 = 0.41421356237309503;

double spec_atan_small(double u) {
  const int N = 22;
  double u2 = u * u;
  double s = 1.0 / (double)(2 * N + 1);
  for (int n = N - 1; n >= 0; --n) s = 1.0 / (double)(2 * n + 1) - u2 * s;
  return u * s;
}
double spec_atan01(double t) {  // t in [0,1]
  if (t > kTanPi8) return kPi / 4.0 + spec_atan_small((t - 1.0) / (t + 1.0));
  return spec_atan_small(t);
}
double spec_atan2_ypos(double y, double x) {  // y >= 0
  double ax = std::fabs(x);
  if (y == 0.0) return x < 0.0 ? kPi : 0.0;
  double r;
  if (ax >= y) r = spec_atan01(y / ax);
  else r = kPi / 2.0 - spec_atan01(ax / y);
  if (x < 0.0) r = kPi - r;
  return r;
}
double spec_cos(double x) {  // |x| <= ~1.1
  double x2 = x * x, c = 1.0;
  for (int n = 13; n >= 1; --n) c = 1.0 - (x2 * (1.0 / (double)((2 * n - 1) * (2 * n)))) * c;
  return c;
}
double spec_sin(double x) {
  double x2 = x * x, s = 1.0;
  for (int n = 13; n >= 1; --n) s = 1.0 - (x2 * (1.0 / (double)((2 * n) * (2 * n + 1)))) * s;
  return x * s;
}

// pcl::computeRoots2 (common/impl/eigen.hpp) [PCL-recall App. A.3]
void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float bb = b * b;
  float d = (float)((double)bb - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

// pcl::computeRoots<Matrix3f> [PCL-recall App. A.3]
void compute_roots(const float m[3][3], float roots[3], int trig_mode) {
  float c0 = m[0][0] * m[1][1] * m[2][2];
  c0 = c0 + 2.0f * m[0][1] * m[0][2] * m[1][2];
  c0 = c0 - m[0][0] * m[1][2] * m[1][2];
  c0 = c0 - m[1][1] * m[0][2] * m[0][2];
  c0 = c0 - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1];
  c1 = c1 - m[0][1] * m[0][1];
  c1 = c1 + m[0][0] * m[2][2];
  c1 = c1 - m[0][2] * m[0][2];
  c1 = c1 + m[1][1] * m[2][2];
  c1 = c1 - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1];
  c2 = c2 + m[2][2];

  if (std::fabs(c0) < FLT_EPSILON) {
    compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  float rho = std::sqrt(-a_over_3);
  float sq = std::sqrt(-q);
  float theta, cos_theta, sin_theta;
  if (trig_mode == 1) {
    theta = std::atan2(sq, half_b) * s_inv3;
    cos_theta = std::cos(theta);
    sin_theta = std::sin(theta);
  } else {
    float at = (float)spec_atan2_ypos((double)sq, (double)half_b);
    theta = at * s_inv3;
    cos_theta = (float)spec_cos((double)theta);
    sin_theta = (float)spec_sin((double)theta);
  }
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.f) compute_roots2(c2, c1, roots);
}

inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// Eigen's fully unrolled 3-element reduction is e0 + (e1 + e2) [PCL-recall]
inline float sqnorm3(const float v[3]) { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); }
inline float dot3_eigen(const float a[3], const float b[3]) {
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
}

// pcl::NormalEstimation::computePointNormal + solvePlaneParameters + eigen33 +
// flipNormalTowardsViewpoint(vp = 0,0,0), the call made at main.cpp:104-109
// [PCL-recall App. A.3].  nbr lists `k` indices in search order (self first).
void point_normal(const float* pts, const int32_t* nbr, int k, const float* self, int trig_mode,
                  float out[4]) {
  int cnt = 0;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
  for (int j = 0; j < k; ++j) {
    if (nbr[j] < 0) continue;
    const float* p = pts + 4 * (i64)nbr[j];
    a0 += p[0] * p[0];
    a1 += p[0] * p[1];
    a2 += p[0] * p[2];
    a3 += p[1] * p[1];
    a4 += p[1] * p[2];
    a5 += p[2] * p[2];
    a6 += p[0];
    a7 += p[1];
    a8 += p[2];
    ++cnt;
  }
  if (cnt < 3) {
    out[0] = out[1] = out[2] = out[3] = NAN;
    return;
  }
  float fc = (float)cnt;
  a0 /= fc; a1 /= fc; a2 /= fc; a3 /= fc; a4 /= fc; a5 /= fc; a6 /= fc; a7 /= fc; a8 /= fc;
  float cov[3][3];
  cov[0][0] = a0 - a6 * a6;
  cov[0][1] = a1 - a6 * a7;
  cov[0][2] = a2 - a6 * a8;
  cov[1][1] = a3 - a7 * a7;
  cov[1][2] = a4 - a7 * a8;
  cov[2][2] = a5 - a8 * a8;
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  // eigen33(mat, eigenvalue, eigenvector)
  float scale = 0.f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) scale = std::max(scale, std::fabs(cov[r][c]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float sm[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) sm[r][c] = cov[r][c] / scale;
  float roots[3];
  compute_roots(sm, roots, trig_mode);
  float eigenvalue = roots[0] * scale;
  sm[0][0] -= roots[0];
  sm[1][1] -= roots[0];
  sm[2][2] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3(sm[0], sm[1], v1);
  cross3(sm[0], sm[2], v2);
  cross3(sm[1], sm[2], v3);
  float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
  float nx, ny, nz;
  if (l1 >= l2 && l1 >= l3) {
    float s = std::sqrt(l1);
    nx = v1[0] / s; ny = v1[1] / s; nz = v1[2] / s;
  } else if (l2 >= l1 && l2 >= l3) {
    float s = std::sqrt(l2);
    nx = v2[0] / s; ny = v2[1] / s; nz = v2[2] / s;
  } else {
    float s = std::sqrt(l3);
    nx = v3[0] / s; ny = v3[1] / s; nz = v3[2] / s;
  }
  float eig_sum = cov[0][0] + cov[1][1];
  eig_sum = eig_sum + cov[2][2];
  float curvature = (eig_sum != 0.f) ? std::fabs(eigenvalue / eig_sum) : 0.f;
  // flipNormalTowardsViewpoint, viewpoint (0,0,0)
  float vx = 0.f - self[0], vy = 0.f - self[1], vz = 0.f - self[2];
  float ct = vx * nx + vy * ny;
  ct = ct + vz * nz;
  if (ct < 0.f) {
    nx *= -1.f; ny *= -1.f; nz *= -1.f;
  }
  out[0] = nx; out[1] = ny; out[2] = nz; out[3] = curvature;
}

}  // namespace

// =============================================================================
// exported API (ctypes)
// =============================================================================
ORC_API int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#endif
}

// kNN of every cloud point against the cloud (self included, d2 = 0 first).
ORC_API void orc_knn_self(const float* pts4, i64 n, int k, int32_t* idx, float* d2) {
  knn_run(pts4, n, pts4, n, k, idx, d2);
}
// kNN of external queries (flat_point.hpp:74, ift.hpp:135, graph.hpp:260-266)
ORC_API void orc_knn_query(const float* pts4, i64 n, const float* q4, i64 m, int k, int32_t* idx,
                           float* d2) {
  knn_run(pts4, n, q4, m, k, idx, d2);
}
// O(n*m) brute force, to validate the tree
ORC_API void orc_knn_brute(const float* pts4, i64 n, const float* q4, i64 m, int k, int32_t* idx,
                           float* d2) {
#pragma omp parallel for schedule(dynamic, 64)
  for (i64 i = 0; i < m; ++i) {
    std::vector<Cand> all((size_t)n);
    for (i64 j = 0; j < n; ++j) {
      all[j].idx = (int)j;
      all[j].d2 = dist2(q4 + 4 * i, pts4 + 4 * j);
    }
    int kk = (int)std::min<i64>(k, n);
    std::partial_sort(all.begin(), all.begin() + kk, all.end(), cand_less);
    for (int j = 0; j < k; ++j) {
      idx[i * k + j] = j < kk ? all[j].idx : -1;
      if (d2) d2[i * k + j] = j < kk ? all[j].d2 : INFINITY;
    }
  }
}

// Normal + curvature per point from its neighbour list (main.cpp:104-109).
ORC_API void orc_normals(const float* pts4, i64 n, const int32_t* nbr, int k, int trig_mode,
                         float* out4) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
    point_normal(pts4, nbr + i * k, k, pts4 + 4 * i, trig_mode, out4 + 4 * i);
}

// Symmetric k-NN graph as CSR with ascending, de-duplicated, self-free rows
// (graph builder (B), SURVEY finding 3).  Returns the arc count; adj may be
// NULL for a counting call.
ORC_API i64 orc_knn_graph(const int32_t* nbr, i64 n, int k, u32* off, u32* adj) {
  // counting form (the vector-of-vectors form took 2.6 s per million points): every valid arc u -> v
  // is filed under both ends, rows are sorted and de-duplicated in place, then compacted.
  // adj == NULL: counting call, returns the arc count only (off is still filled).
  std::vector<u64> start((size_t)n + 1, 0);
  for (i64 u = 0; u < n; ++u)
    for (int j = 0; j < k; ++j) {
      int v = nbr[u * k + j];
      if (v < 0 || v == u) continue;
      ++start[(size_t)u + 1];
      ++start[(size_t)v + 1];
    }
  for (i64 u = 0; u < n; ++u) start[(size_t)u + 1] += start[(size_t)u];
  std::vector<u32> raw((size_t)start[(size_t)n]);
  {
    std::vector<u64> cur(start.begin(), start.end() - 1);
    for (i64 u = 0; u < n; ++u)
      for (int j = 0; j < k; ++j) {
        int v = nbr[u * k + j];
        if (v < 0 || v == u) continue;
        raw[cur[(size_t)u]++] = (u32)v;
        raw[cur[(size_t)v]++] = (u32)u;
      }
  }
  std::vector<u32> deg((size_t)n);
#pragma omp parallel for schedule(dynamic, 4096)
  for (i64 u = 0; u < n; ++u) {
    u32* b = raw.data() + start[(size_t)u];
    u32* e = raw.data() + start[(size_t)u + 1];
    std::sort(b, e);
    deg[(size_t)u] = (u32)(std::unique(b, e) - b);
  }
  i64 tot = 0;
  for (i64 u = 0; u < n; ++u) {
    off[u] = (u32)tot;
    if (adj) std::copy(raw.data() + start[(size_t)u], raw.data() + start[(size_t)u] + deg[(size_t)u], adj + tot);
    tot += (i64)deg[(size_t)u];
  }
  off[n] = (u32)tot;
  return tot;
}

// -----------------------------------------------------------------------------
// pcl::RegionGrowing::extract as driven by FlatPoint::getCentroids
// (flat_point.hpp:23-42; parameters main.cpp:112-119) [PCL-recall App. A.4].
// Sequential, literal: sort by (curvature, index) — the canonical form of the
// unstable std::sort on curvature alone —, seeded BFS with the smooth-mode test
// |n_cur . n_nbr| >= cos(theta), curvature gate on enqueue, then the size
// filter.  point_label[i] = index of the kept cluster (in PCL's emission
// order) or -1.  raw_label[i] (optional) = segment number before the filter.
// Returns the number of kept clusters.
// -----------------------------------------------------------------------------
ORC_API i64 orc_region_growing(const float* normals4, const int32_t* nbr, i64 n, int k, float theta,
                               float curv_thr, int min_sz, int max_sz, int32_t* point_label,
                               int32_t* raw_label, i64* n_segments_out) {
  std::vector<std::pair<float, int>> order((size_t)n);
  for (i64 i = 0; i < n; ++i) order[i] = {normals4[4 * i + 3], (int)i};
  std::sort(order.begin(), order.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
    return a.first < b.first || (a.first == b.first && a.second < b.second);
  });
  const float cosine_threshold = cosf(theta);
  std::vector<int> lab((size_t)n, -1);
  std::vector<int> seg_size;
  i64 segmented = 0, seed_counter = 0;
  int nseg = 0;
  std::queue<int> q;
  while (segmented < n) {
    while (seed_counter < n && lab[order[seed_counter].second] != -1) ++seed_counter;
    if (seed_counter >= n) break;
    int seed = order[seed_counter].second;
    // growRegion
    q.push(seed);
    lab[seed] = nseg;
    int cnt = 1;
    while (!q.empty()) {
      int cur = q.front();
      q.pop();
      const float* nc = normals4 + 4 * (i64)cur;
      for (int j = 0; j < k; ++j) {
        int v = nbr[(i64)cur * k + j];
        if (v < 0 || lab[v] != -1) continue;
        const float* nv = normals4 + 4 * (i64)v;
        float dp = std::fabs(dot3_eigen(nv, nc));
        if (dp < cosine_threshold) continue;  // NaN passes, as in PCL
        lab[v] = nseg;
        ++cnt;
        if (!(nv[3] > curv_thr)) q.push(v);
      }
    }
    seg_size.push_back(cnt);
    segmented += cnt;
    ++nseg;
  }
  std::vector<int> kept_id((size_t)nseg, -1);
  i64 kept = 0;
  for (int s = 0; s < nseg; ++s)
    if (seg_size[s] >= min_sz && seg_size[s] <= max_sz) kept_id[s] = (int)kept++;
  for (i64 i = 0; i < n; ++i) {
    if (raw_label) raw_label[i] = lab[i];
    point_label[i] = lab[i] >= 0 ? kept_id[lab[i]] : -1;
  }
  if (n_segments_out) *n_segments_out = nseg;
  return kept;
}

// The exact parallel form of the same procedure when the curvature gate never
// fires: label(v) = min rank over all u that reach v along
// arcs u->w, w in kNN(u), |n_u.n_w| >= cos(theta).  rank = position in the
// (curvature, index) order.  Output: minrank[v].  Used only by tests, to prove
// the restatement the GPU kernel relies on.
ORC_API void orc_rg_minrank(const float* normals4, const int32_t* nbr, i64 n, int k, float theta,
                            int32_t* minrank) {
  std::vector<std::pair<float, int>> order((size_t)n);
  for (i64 i = 0; i < n; ++i) order[i] = {normals4[4 * i + 3], (int)i};
  std::sort(order.begin(), order.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
    return a.first < b.first || (a.first == b.first && a.second < b.second);
  });
  for (i64 r = 0; r < n; ++r) minrank[order[r].second] = (int)r;
  const float cosine_threshold = cosf(theta);
  bool changed = true;
  while (changed) {
    changed = false;
    for (i64 u = 0; u < n; ++u) {
      const float* nu = normals4 + 4 * u;
      for (int j = 0; j < k; ++j) {
        int v = nbr[u * k + j];
        if (v < 0) continue;
        float dp = std::fabs(dot3_eigen(normals4 + 4 * (i64)v, nu));
        if (dp < cosine_threshold) continue;
        if (minrank[u] < minrank[v]) {
          minrank[v] = minrank[u];
          changed = true;
        }
      }
    }
  }
}

// FlatPoint::getCentroids centroid part (flat_point.hpp:34-38, via
// pcl::compute3DCentroid: FP32 Vector4f accumulation over the cluster's indices
// in ascending point order, as assembleRegions emits them, divided by the
// count) + compute_centroid_info's snap to the nearest input point
// (flat_point.hpp:71-76).  seed_idx[c] = index of the snapped point.
ORC_API void orc_rg_seeds(const float* pts4, i64 n, const int32_t* point_label, i64 n_kept,
                          int32_t* seed_idx, float* centroid4) {
  std::vector<float> sx((size_t)n_kept, 0.f), sy((size_t)n_kept, 0.f), sz((size_t)n_kept, 0.f);
  std::vector<i64> cnt((size_t)n_kept, 0);
  for (i64 i = 0; i < n; ++i) {
    int c = point_label[i];
    if (c < 0) continue;
    sx[c] += pts4[4 * i];
    sy[c] += pts4[4 * i + 1];
    sz[c] += pts4[4 * i + 2];
    cnt[c]++;
  }
  std::vector<float> q((size_t)n_kept * 4);
  for (i64 c = 0; c < n_kept; ++c) {
    float d = (float)cnt[c];
    q[4 * c] = sx[c] / d;
    q[4 * c + 1] = sy[c] / d;
    q[4 * c + 2] = sz[c] / d;
    q[4 * c + 3] = 0.f;
  }
  if (centroid4) std::copy(q.begin(), q.end(), centroid4);
  knn_run(pts4, n, q.data(), n_kept, 1, seed_idx, nullptr);
}

// =============================================================================
// IFT watershed with the f_max path cost (ift.hpp:170-250)
// =============================================================================
static const unsigned MAX_COST = 500;  // ift.hpp:1

// Policy R "first offer": vertices leave in non-decreasing cost, FIFO inside a
// cost bucket, seeds enqueued in ascending vertex id, adjacency ascending;
// a vertex takes the label of the first strictly improving offer
// (ift.hpp:196-205) and is never relabelled on equal cost.
static void ift_policy_R(const u32* off, const u32* adj, i64 n, const float* w, const int32_t* seeds,
                         i64 ns, u32* cost, u32* label) {
  for (i64 v = 0; v < n; ++v) {
    cost[v] = MAX_COST;  // ift.hpp:101
    label[v] = (u32)v;   // ift.hpp:179
  }
  std::vector<int32_t> s(seeds, seeds + ns);
  std::sort(s.begin(), s.end());
  s.erase(std::unique(s.begin(), s.end()), s.end());
  std::vector<std::deque<u32>> bucket(MAX_COST);
  for (int32_t v : s) {
    cost[v] = 0;  // ift.hpp:138 — seed cost forced to 0
    bucket[0].push_back((u32)v);
  }
  for (unsigned c = 0; c < MAX_COST; ++c) {
    auto& b = bucket[c];
    while (!b.empty()) {
      u32 sv = b.front();
      b.pop_front();
      for (u32 e = off[sv]; e < off[sv + 1]; ++e) {
        u32 t = adj[e];
        if (cost[t] > cost[sv]) {                          // ift.hpp:196
          float tmp = std::max(w[t], (float)cost[sv]);     // ift.hpp:198
          if (tmp < (float)cost[t]) {                      // ift.hpp:199
            cost[t] = (unsigned)tmp;                       // ift.hpp:203 (truncation)
            label[t] = label[sv];                          // ift.hpp:204
            bucket[cost[t]].push_back(t);                  // ift.hpp:205
          }
        }
      }
    }
  }
}

// Policy C "min seed": the bit-exact GPU<->oracle contract.
// Lexicographic Dijkstra on (cost, label): a vertex is final when popped, offers
// are made only by final vertices, a vertex keeps the smallest (cost, label)
// offered; seeds are final from the start and are never relabelled.  Weights
// must be integer valued (App. B #6), so truncation is a no-op.
static void ift_policy_C(const u32* off, const u32* adj, i64 n, const float* w, const int32_t* seeds,
                         i64 ns, u32* cost, u32* label) {
  std::vector<u64> key((size_t)n);
  std::vector<uint8_t> is_seed((size_t)n, 0), done((size_t)n, 0);
  for (i64 v = 0; v < n; ++v) key[v] = ((u64)MAX_COST << 32) | (u32)v;
  typedef std::pair<u64, u32> QE;
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
  for (i64 i = 0; i < ns; ++i) {
    u32 v = (u32)seeds[i];
    if (is_seed[v]) continue;
    is_seed[v] = 1;
    key[v] = (u64)v;  // cost 0, label = self
    pq.push({key[v], v});
  }
  while (!pq.empty()) {
    QE top = pq.top();
    pq.pop();
    u32 sv = top.second;
    if (done[sv] || top.first != key[sv]) continue;
    done[sv] = 1;
    u32 cs = (u32)(key[sv] >> 32), ls = (u32)key[sv];
    for (u32 e = off[sv]; e < off[sv + 1]; ++e) {
      u32 t = adj[e];
      if (is_seed[t] || done[t]) continue;
      float tmp = std::max(w[t], (float)cs);
      if (!(tmp < (float)MAX_COST)) continue;
      u64 cand = ((u64)(unsigned)tmp << 32) | ls;
      if (cand < key[t]) {
        key[t] = cand;
        pq.push({cand, t});
      }
    }
  }
  for (i64 v = 0; v < n; ++v) {
    cost[v] = (u32)(key[v] >> 32);
    label[v] = (u32)key[v];
  }
}

// Literal flavour: the reference's own containers and control flow
// (ift.hpp:170-250 with globals.hpp:57-81: std::priority_queue on cost only,
// remove() = find + erase + make_heap), with vertex ids standing in for the
// heap addresses that order every setS iteration in the reference.  Also runs
// the interleaved region-attribute / "MST" block (ift.hpp:5-23, 155-168,
// 207-243).  O(V*|Q|): small inputs only.
struct RInfo {
  double height = 0, area = 0, volume = 0;
  bool status = false;
  void update(double i) {  // ift.hpp:11-16
    area += 1;
    if (i > height) height = i;
    volume += area * height;
  }
};
struct LiteralOut {
  std::vector<u32> root_mst;
  std::map<std::pair<u32, u32>, double> mst;
};
static void ift_literal(const u32* off, const u32* adj, i64 n, const float* w, const int32_t* seeds,
                        i64 ns, u32* cost, u32* label, LiteralOut* lo) {
  typedef std::pair<u32, unsigned> P;
  struct Cmp {
    bool operator()(const P& a, const P& b) const { return a.second > b.second; }
  };
  std::vector<P> heap;
  Cmp cmp;
  std::map<u32, RInfo> r_info;
  for (i64 v = 0; v < n; ++v) cost[v] = MAX_COST;
  for (i64 i = 0; i < ns; ++i) {  // ift.hpp:132-142
    u32 v = (u32)seeds[i];
    cost[v] = 0;
    RInfo ri;
    ri.height = w[v];
    ri.area = 1;
    r_info[v] = ri;
  }
  std::vector<u32> root_mst((size_t)n);
  for (i64 v = 0; v < n; ++v) {  // ift.hpp:177-183
    label[v] = (u32)v;
    root_mst[v] = (u32)v;
    if (cost[v] != MAX_COST) {
      heap.push_back(P((u32)v, cost[v]));
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
  }
  std::map<std::pair<u32, u32>, double> mst;
  auto find_and_swap = [&](u32 old_, u32 new_) {  // ift.hpp:155-168
    for (i64 v = 0; v < n; ++v)
      if (root_mst[v] == old_) root_mst[v] = new_;
  };
  while (!heap.empty()) {  // ift.hpp:188
    std::pop_heap(heap.begin(), heap.end(), cmp);
    u32 s = heap.back().first;
    heap.pop_back();
    u32 s_root = root_mst[s];
    for (u32 e = off[s]; e < off[s + 1]; ++e) {
      u32 t = adj[e];
      if (cost[t] > cost[s]) {
        float tmp = std::max(w[t], (float)cost[s]);
        if (tmp < (float)cost[t]) {
          if (cost[t] != MAX_COST) {  // Q.remove, globals.hpp:63-73
            auto it = std::find(heap.begin(), heap.end(), P(t, cost[t]));
            if (it != heap.end()) {
              heap.erase(it);
              std::make_heap(heap.begin(), heap.end(), cmp);
            }
          }
          cost[t] = (unsigned)tmp;
          label[t] = label[s];
          heap.push_back(P(t, cost[t]));
          std::push_heap(heap.begin(), heap.end(), cmp);
          root_mst[t] = root_mst[s];
          r_info[s_root].update(w[t]);  // ift.hpp:209 (creates the entry if absent, as the map does)
          for (u32 e2 = off[t]; e2 < off[t + 1]; ++e2) {  // ift.hpp:211-243
            u32 r = adj[e2];
            u32 r_root = root_mst[r];
            bool exists = r_info.find(r_root) != r_info.end();
            if (exists && r_root != s_root && !r_info[s_root].status) {
              r_info[s_root].status = true;
              double s_volume = r_info[s_root].volume, r_volume = r_info[r_root].volume, wgt = 0;
              if (s_volume < r_volume) {
                wgt = s_volume;
                find_and_swap(s_root, r_root);
                r_info[r_root].area += r_info[s_root].area;
                r_info[r_root].volume += r_info[s_root].volume;
              } else {
                wgt = r_volume;
                find_and_swap(r_root, s_root);
                r_info[s_root].area += r_info[r_root].area;
                r_info[s_root].volume += r_info[r_root].volume;
              }
              mst[std::make_pair(s_root, r_root)] = wgt;
            }
          }
        }
      }
    }
  }
  if (lo) {
    lo->root_mst = root_mst;
    lo->mst = mst;
  }
}

// policy: 0 = R (first offer, FIFO), 1 = C (min seed, contract), 2 = literal
ORC_API void orc_ift(const u32* off, const u32* adj, i64 n, const float* w, const int32_t* seeds,
                     i64 ns, int policy, u32* cost, u32* label) {
  if (policy == 0) ift_policy_R(off, adj, n, w, seeds, ns, cost, label);
  else if (policy == 1) ift_policy_C(off, adj, n, w, seeds, ns, cost, label);
  else ift_literal(off, adj, n, w, seeds, ns, cost, label, nullptr);
}

// Literal IFT that also returns the "MST" edge map of ift.hpp:207-243:
// mst_out rows {s_root, r_root, weight-as-double-bits} — call with cap rows.
ORC_API i64 orc_ift_literal_mst(const u32* off, const u32* adj, i64 n, const float* w,
                                const int32_t* seeds, i64 ns, u32* cost, u32* label, u32* mst_a,
                                u32* mst_b, double* mst_w, i64 cap) {
  LiteralOut lo;
  ift_literal(off, adj, n, w, seeds, ns, cost, label, &lo);
  i64 m = 0;
  for (auto& kv : lo.mst) {
    if (m < cap) {
      mst_a[m] = kv.first.first;
      mst_b[m] = kv.first.second;
      mst_w[m] = kv.second;
    }
    ++m;
  }
  return m;
}

// Tie-admissibility audit of a (cost, label) map: every reached non-seed vertex
// must be reachable from its own seed through arcs s->t with label[s]==label[t],
// cost[s] <= cost[t] and max(w[t], cost[s]) == cost[t].  Returns the number of
// reached vertices that are NOT so reachable (0 = the labelling is a valid
// tie-break of the reference rule).
ORC_API i64 orc_ift_audit(const u32* off, const u32* adj, i64 n, const float* w, const int32_t* seeds,
                          i64 ns, const u32* cost, const u32* label) {
  std::vector<uint8_t> ok((size_t)n, 0);
  std::vector<u32> stack;
  for (i64 i = 0; i < ns; ++i) {
    u32 s = (u32)seeds[i];
    if (ok[s]) continue;
    if (cost[s] != 0 || label[s] != s) return -1;
    ok[s] = 1;
    stack.push_back(s);
  }
  while (!stack.empty()) {
    u32 s = stack.back();
    stack.pop_back();
    for (u32 e = off[s]; e < off[s + 1]; ++e) {
      u32 t = adj[e];
      if (ok[t] || label[t] != label[s] || cost[t] >= MAX_COST || cost[s] > cost[t]) continue;
      float tmp = std::max(w[t], (float)cost[s]);
      if ((unsigned)tmp != cost[t]) continue;
      ok[t] = 1;
      stack.push_back(t);
    }
  }
  i64 bad = 0;
  for (i64 v = 0; v < n; ++v) {
    if (cost[v] < MAX_COST && !ok[v]) ++bad;
    if (cost[v] >= MAX_COST && label[v] != (u32)v) ++bad;
  }
  return bad;
}

// =============================================================================
// Voxel-26 adjacency graph (graph builder (A), graph.hpp:228-308)
// =============================================================================

// Graph::computeCloudResolution (graph.hpp:228-259): mean over points of the
// distance to the 2nd nearest neighbour, times overlap.
//   mode 1: the literal sequential double accumulation of graph.hpp:248.
//   mode 0: the order-independent definition the GPU implements.  Every term
//           t = sqrtf(d2) is split exactly (in binary64) into
//             A = floor(t * 2^24)            (units of 2^-24)
//             B = floor((t * 2^24 - A) * 2^40)   (units of 2^-64)
//           and the two limbs are summed as 64-bit integers; the mean is
//           (sumA * 2^-24 + sumB * 2^-64) / count in long double, rounded to
//           double once.  It differs from mode 1 only by mode 1's own
//           accumulated rounding (~1e-16 relative).
ORC_API double orc_cloud_resolution(const float* pts4, i64 n, double overlap, int mode) {
  std::vector<int32_t> idx((size_t)n * 2);
  std::vector<float> d2((size_t)n * 2);
  knn_run(pts4, n, pts4, n, 2, idx.data(), d2.data());
  double res = 0.0;
  i64 cnt = 0;
  if (mode == 1) {
    for (i64 i = 0; i < n; ++i) {
      if (idx[2 * i + 1] < 0) continue;
      res += (double)std::sqrt(d2[2 * i + 1]);  // graph.hpp:248 (sqrt on float)
      ++cnt;
    }
    if (cnt != 0) res /= (double)cnt;  // graph.hpp:254
  } else {
    u64 sumA = 0, sumB = 0;
    for (i64 i = 0; i < n; ++i) {
      if (idx[2 * i + 1] < 0) continue;
      double t = (double)std::sqrt(d2[2 * i + 1]) * 16777216.0;
      double a = std::floor(t);
      double b = std::floor((t - a) * 1099511627776.0);
      sumA += (u64)a;
      sumB += (u64)b;
      ++cnt;
    }
    if (cnt != 0) {
      long double tot = (long double)sumA * 0x1p-24L + (long double)sumB * 0x1p-64L;
      res = (double)(tot / (long double)cnt);
    }
  }
  if (cnt != 0) res *= overlap;  // graph.hpp:256
  return res;
}

// -----------------------------------------------------------------------------
// pcl::octree::OctreePointCloudAdjacency as driven by Graph::initialize
// (graph.hpp:293-308) [PCL-recall App. A.2]: bounding box over the points,
// getKeyBitSize (depth, symmetric padding to side 2^depth * res, binary64), key =
// (unsigned)((p - min') / res) per axis, vertex = voxel CENTRE, neighbours = occupied
// voxels at key offsets {-1,0,1}^3 clipped to [0, max_key].  Vertices are numbered in
// ascending Morton order of the key (x = most significant axis bit): the canonical
// stand-in for the reference's address order.  The self loop PCL adds is NOT emitted
// in the CSR (morphology / regmin add the vertex itself explicitly).
// Graph::optimizeGraph (graph.hpp:269-284): vertex intensity = intensity of the
// nearest original point.
// Outputs: centres4[V] {x,y,z,intensity}, keys3[V*3], off[V+1], adj (cap 26 V),
// box[4] = {min_x', min_y', min_z', depth}.  Returns V.
// -----------------------------------------------------------------------------
static inline u64 morton_spread21(u64 v) {
  u64 x = v & 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
static inline u64 morton_xyz(u32 x, u32 y, u32 z) {
  return (morton_spread21(x) << 2) | (morton_spread21(y) << 1) | morton_spread21(z);
}

ORC_API i64 orc_voxel_graph(const float* pts4, i64 n, double resolution, float* centres4, u32* keys3,
                            u32* off, u32* adj, double* box) {
  double mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (i64 i = 0; i < n; ++i) {
    const float* p = pts4 + 4 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; ++a) {
      if (p[a] < mn[a]) mn[a] = p[a];
      if (p[a] > mx[a]) mx[a] = p[a];
    }
  }
  // OctreePointCloud::getKeyBitSize
  const float minValue = FLT_EPSILON;
  unsigned mk[3];
  for (int a = 0; a < 3; ++a) mk[a] = (unsigned)std::ceil((mx[a] - mn[a] - minValue) / resolution);
  unsigned max_voxels = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
  unsigned depth = std::max(std::min(32u, (unsigned)std::ceil(std::log((double)max_voxels) / std::log(2.0) - minValue)), 0u);
  double side = (double)(1u << depth) * resolution;
  for (int a = 0; a < 3; ++a) {
    double over = (side - (mx[a] - mn[a])) / 2.0;
    if (over > minValue) {
      mn[a] -= over;
      mx[a] += over;
    }
  }
  const u32 max_key = (1u << depth) - 1u;
  if (box) {
    box[0] = mn[0]; box[1] = mn[1]; box[2] = mn[2]; box[3] = (double)depth;
  }
  // keys
  std::vector<std::pair<u64, std::array<u32, 3>>> vox;
  vox.reserve((size_t)n);
  for (i64 i = 0; i < n; ++i) {
    const float* p = pts4 + 4 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    std::array<u32, 3> k;
    for (int a = 0; a < 3; ++a) k[a] = (u32)(((double)p[a] - mn[a]) / resolution);
    vox.push_back({morton_xyz(k[0], k[1], k[2]), k});
  }
  std::sort(vox.begin(), vox.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  vox.erase(std::unique(vox.begin(), vox.end(), [](const auto& a, const auto& b) { return a.first == b.first; }),
            vox.end());
  const i64 V = (i64)vox.size();
  std::vector<float> q((size_t)V * 4);
  for (i64 v = 0; v < V; ++v) {
    for (int a = 0; a < 3; ++a) {
      keys3[3 * v + a] = vox[v].second[a];
      q[4 * v + a] = (float)(((double)vox[v].second[a] + 0.5f) * resolution + mn[a]);
    }
    q[4 * v + 3] = 0.f;
  }
  // optimizeGraph: intensity of the nearest original point
  std::vector<int32_t> nn((size_t)V);
  knn_run(pts4, n, q.data(), V, 1, nn.data(), nullptr);
  for (i64 v = 0; v < V; ++v) {
    for (int a = 0; a < 3; ++a) centres4[4 * v + a] = q[4 * v + a];
    centres4[4 * v + 3] = nn[v] >= 0 ? pts4[4 * (i64)nn[v] + 3] : 0.f;
  }
  // computeNeighbors
  i64 tot = 0;
  for (i64 v = 0; v < V; ++v) {
    off[v] = (u32)tot;
    const auto& k = vox[v].second;
    std::vector<u32> row;
    for (int dx = (k[0] > 0 ? -1 : 0); dx <= (k[0] == max_key ? 0 : 1); ++dx)
      for (int dy = (k[1] > 0 ? -1 : 0); dy <= (k[1] == max_key ? 0 : 1); ++dy)
        for (int dz = (k[2] > 0 ? -1 : 0); dz <= (k[2] == max_key ? 0 : 1); ++dz) {
          if (dx == 0 && dy == 0 && dz == 0) continue;
          u64 code = morton_xyz(k[0] + dx, k[1] + dy, k[2] + dz);
          auto it = std::lower_bound(vox.begin(), vox.end(), code,
                                     [](const auto& a, u64 c) { return a.first < c; });
          if (it != vox.end() && it->first == code) row.push_back((u32)(it - vox.begin()));
        }
    std::sort(row.begin(), row.end());
    for (u32 r : row) adj[tot++] = r;
  }
  off[V] = (u32)tot;
  return V;
}

// Graph::morph_base (graph.hpp:323-371) over the CLOSED neighbourhood (PCL's voxel
// adjacency carries a self loop, App. A.2).  op: 1 bin_erode, 2 bin_dilate, 3 dilate,
// 4 erode (enum MORPH, graph.hpp:13), 5 gradient = dilate - erode (graph.hpp:147-163).
// bin_dilate: the reference discards its result (shadowed local, graph.hpp:345-349) and
// always writes 0; this restatement implements the evident intent (App. B #3).
ORC_API void orc_morph(const u32* off, const u32* adj, i64 n, const float* inten, int op, float* out) {
  auto one = [&](i64 v, int o) -> float {
    float r;
    if (o == 4) {
      r = 255.f;
      if (inten[v] < r) r = inten[v];
      for (u32 e = off[v]; e < off[v + 1]; ++e)
        if (inten[adj[e]] < r) r = inten[adj[e]];
    } else if (o == 3) {
      r = 0.f;
      if (inten[v] > r) r = inten[v];
      for (u32 e = off[v]; e < off[v + 1]; ++e)
        if (inten[adj[e]] > r) r = inten[adj[e]];
    } else if (o == 1) {
      bool hit = inten[v] == 0.f;
      for (u32 e = off[v]; e < off[v + 1]; ++e) hit = hit || inten[adj[e]] == 0.f;
      r = hit ? 0.f : 255.f;
    } else {
      bool hit = inten[v] == 255.f;
      for (u32 e = off[v]; e < off[v + 1]; ++e) hit = hit || inten[adj[e]] == 255.f;
      r = hit ? 255.f : 0.f;
    }
    return r;
  };
#pragma omp parallel for schedule(static)
  for (i64 v = 0; v < n; ++v) out[v] = op == 5 ? one(v, 3) - one(v, 4) : one(v, op);
}

// Graph::getRegmin (graph.hpp:209-226): a vertex is a seed if its value is 0 or equals
// the minimum over its closed neighbourhood.  Returns the count; seeds ascending.
ORC_API i64 orc_regmin(const u32* off, const u32* adj, i64 n, const float* inten, int32_t* seeds) {
  i64 m = 0;
  for (i64 v = 0; v < n; ++v) {
    float val = inten[v];
    bool is_min = val == 0.f;
    if (!is_min) {
      float mn = val;
      for (u32 e = off[v]; e < off[v + 1]; ++e) mn = std::min(mn, inten[adj[e]]);
      is_min = val == mn;
    }
    if (is_min) seeds[m++] = (int32_t)v;
  }
  return m;
}

// =============================================================================
// Region hierarchy (I4 / L1)
// =============================================================================
// The reference interleaves region attributes and a first-contact "MST" with the
// heap loop (ift.hpp:207-243) and merges regions in Cluster (cluster.hpp:83-221).
// Both depend on the sequential conquest order and on container address order, so they
// have no order-free equivalent (hard part 7).  Two restatements follow:
//   * the PRODUCT DEFINITION (what the GPU path implements, bit-exact contract):
//     region adjacency with watershed saddles, area/height attributes, single-linkage
//     merge of the regions down to num_regions along the saddle minimum spanning forest;
//   * the LITERAL Cluster of cluster.hpp on the literal MST of ift.hpp (informational).

// Region adjacency: for every pair of adjacent conquered vertices with different labels,
// saddle(a,b) = min over such arcs of max(cost[u], cost[v]).  Output rows sorted by
// (a, b), a < b.  Returns the number of pairs (call with cap = 0 to count).
ORC_API i64 orc_region_adjacency(const u32* off, const u32* adj, i64 n, const u32* cost, const u32* label,
                                 u32* ra, u32* rb, u32* saddle, i64 cap) {
  std::map<std::pair<u32, u32>, u32> tab;
  for (i64 u = 0; u < n; ++u) {
    if (cost[u] >= MAX_COST) continue;
    for (u32 e = off[u]; e < off[u + 1]; ++e) {
      u32 v = adj[e];
      if (cost[v] >= MAX_COST || label[v] == label[u]) continue;
      u32 a = std::min(label[u], label[v]), b = std::max(label[u], label[v]);
      u32 s = std::max(cost[u], cost[v]);
      auto it = tab.find({a, b});
      if (it == tab.end()) tab[{a, b}] = s;
      else if (s < it->second) it->second = s;
    }
  }
  i64 m = 0;
  for (auto& kv : tab) {
    if (m < cap) {
      ra[m] = kv.first.first;
      rb[m] = kv.first.second;
      saddle[m] = kv.second;
    }
    ++m;
  }
  return m;
}

// Region attributes in the spirit of r_info (ift.hpp:5-23), order-free part only:
// area = number of conquered vertices (seed included), height = max vertex value over
// them (the seed's own value included, ift.hpp:140-141).  Regions sorted by seed id.
ORC_API i64 orc_region_info(i64 n, const float* w, const u32* cost, const u32* label, u32* seed, u32* area,
                            float* height, i64 cap) {
  std::map<u32, std::pair<u32, float>> tab;
  for (i64 v = 0; v < n; ++v) {
    if (cost[v] >= MAX_COST) continue;
    auto& e = tab[label[v]];
    e.first += 1;
    e.second = std::max(e.second, w[v]);
  }
  i64 m = 0;
  for (auto& kv : tab) {
    if (m < cap) {
      seed[m] = kv.first;
      area[m] = kv.second.first;
      height[m] = kv.second.second;
    }
    ++m;
  }
  return m;
}

// Single-linkage merge along the saddle minimum spanning forest: Kruskal over the pairs
// in ascending (saddle, a, b), stopping when num_regions clusters remain (or no pair is
// left).  cluster_of[v] = smallest seed id of v's merged cluster, or 0xFFFFFFFF for
// unconquered vertices.  Returns the number of clusters.
ORC_API i64 orc_cluster_saddle(i64 n, const u32* cost, const u32* label, const u32* ra, const u32* rb,
                               const u32* saddle, i64 m, int num_regions, u32* cluster_of) {
  std::vector<u32> seeds;
  for (i64 v = 0; v < n; ++v)
    if (cost[v] < MAX_COST) seeds.push_back(label[v]);
  std::sort(seeds.begin(), seeds.end());
  seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
  const i64 S = (i64)seeds.size();
  auto ord = [&](u32 s) { return (i64)(std::lower_bound(seeds.begin(), seeds.end(), s) - seeds.begin()); };
  std::vector<i64> parent((size_t)S);
  std::iota(parent.begin(), parent.end(), 0);
  std::function<i64(i64)> find = [&](i64 x) { return parent[x] == x ? x : parent[x] = find(parent[x]); };
  std::vector<i64> order((size_t)m);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](i64 i, i64 j) {
    if (saddle[i] != saddle[j]) return saddle[i] < saddle[j];
    if (ra[i] != ra[j]) return ra[i] < ra[j];
    return rb[i] < rb[j];
  });
  i64 clusters = S;
  for (i64 idx : order) {
    if (clusters <= (i64)num_regions) break;
    i64 a = find(ord(ra[idx])), b = find(ord(rb[idx]));
    if (a == b) continue;
    if (a < b) parent[b] = a;  // root = smallest ordinal = smallest seed id
    else parent[a] = b;
    --clusters;
  }
  for (i64 v = 0; v < n; ++v)
    cluster_of[v] = cost[v] < MAX_COST ? seeds[(size_t)find(ord(label[v]))] : 0xFFFFFFFFu;
  return clusters;
}

// The literal Cluster of cluster.hpp:23-229 fed by the literal MST of ift.hpp:207-243,
// vertex ids standing in for addresses.  Informational: quirks kept — the region-id
// rewrite of create_TLC acts on a COPY of the heap's container and has no effect
// (cluster.hpp:100-108); single_linking only ever merges while the first key has a partner
// (its inner iterator is never rewound, cluster.hpp:174-195).  Guards (App. B #10): returns
// -1 when TLC.size() <= num_regions or the distance map empties.
// region_of[v] = index into the final `regions` list, or -1.
ORC_API i64 orc_cluster_literal(i64 n, const u32* label, const u32* mst_a, const u32* mst_b, const double* mst_w,
                                i64 m, int num_regions, int32_t* region_of) {
  // root_translator: first-appearance order over the vertex-ordered root map (cluster.hpp:85-89)
  std::map<u32, int> translator;
  int counter = 0;
  for (i64 v = 0; v < n; ++v)
    if (translator.find(label[v]) == translator.end()) translator[label[v]] = counter++;
  struct Edge {
    int r1, r2;
    double w;
  };
  auto cmp = [](const Edge& a, const Edge& b) { return a.w > b.w; };
  std::priority_queue<Edge, std::vector<Edge>, decltype(cmp)> pq(cmp);
  // the reference's std::map is ordered by (s_root, r_root)
  std::vector<i64> eo((size_t)m);
  std::iota(eo.begin(), eo.end(), 0);
  std::sort(eo.begin(), eo.end(), [&](i64 i, i64 j) {
    return mst_a[i] != mst_a[j] ? mst_a[i] < mst_a[j] : mst_b[i] < mst_b[j];
  });
  for (i64 i : eo) pq.push(Edge{translator[mst_a[i]], translator[mst_b[i]], mst_w[i]});
  std::vector<Edge> tlc;
  while (!pq.empty()) {
    tlc.push_back(pq.top());
    pq.pop();
  }
  for (i64 v = 0; v < n; ++v) region_of[v] = -1;
  if ((i64)tlc.size() <= (i64)num_regions || tlc.size() < 2) return -1;
  typedef std::set<int> Key;
  std::map<Key, double> dist;
  for (int i = 0; i < (int)tlc.size(); ++i)
    for (int j = i + 1; j < (int)tlc.size(); ++j) dist[Key{i, j}] = std::sqrt(std::pow(tlc[i].w - tlc[j].w, 2));
  auto shares = [](const Key& a, const Key& b) {
    for (int x : a)
      if (b.count(x)) return true;
    return false;
  };
  auto minus = [](const Key& a, const Key& k) {
    Key r;
    std::set_difference(a.begin(), a.end(), k.begin(), k.end(), std::inserter(r, r.begin()));
    return r;
  };
  for (size_t it = 0; it < tlc.size() - (size_t)num_regions; ++it) {
    if (dist.empty()) return -1;
    Key key;
    double best = INFINITY;
    for (auto& kv : dist)
      if (kv.second < best) {
        key = kv.first;
        best = kv.second;
      }
    std::vector<Key> keys;
    for (auto& kv : dist)
      if (shares(key, kv.first) && kv.first != key) keys.push_back(kv.first);
    // only the first key ever finds partners; both iterators restart after a merge
    size_t i2 = 0;
    while (!keys.empty() && i2 < keys.size()) {
      if (minus(keys[0], key) == minus(keys[i2], key) && keys[0] != keys[i2]) {
        Key a = keys[0], b = keys[i2];
        if (!(dist[a] < dist[b])) std::swap(a, b);
        double wgt = dist[a];
        Key merged = a;
        merged.insert(key.begin(), key.end());
        dist[merged] = wgt;
        dist.erase(b);
        dist.erase(a);
        keys.erase(keys.begin() + (long)i2);
        keys.erase(keys.begin());
        i2 = 0;
      } else {
        ++i2;
      }
    }
    dist.erase(key);
  }
  if (dist.empty()) return -1;
  std::vector<Key> regions;
  regions.push_back(dist.begin()->first);
  for (auto& kv : dist) {
    const Key& k = kv.first;
    if (k != regions[0] && shares(k, regions[0]) && regions.size() == 1) {
      Key r1, r2, r3;
      std::set_intersection(k.begin(), k.end(), regions[0].begin(), regions[0].end(), std::inserter(r1, r1.begin()));
      r2 = minus(regions[0], r1);
      r3 = minus(k, r1);
      regions.push_back(r1);
      regions.push_back(r2);
      regions.push_back(r3);
      regions.erase(regions.begin());
    } else if (shares(k, regions[0]) && regions.size() > 1) {
      regions.push_back(minus(k, regions[0]));
    }
  }
  // getLabelCloud (cluster.hpp:54-79): root -> first TLC entry naming it -> first region holding that entry
  for (i64 v = 0; v < n; ++v) {
    int r = translator[label[v]];
    int tlc_idx = -1;
    for (int i = 0; i < (int)tlc.size(); ++i)
      if (r == tlc[i].r1 || r == tlc[i].r2) {
        tlc_idx = i;
        break;
      }
    if (tlc_idx < 0) continue;
    for (int c = 0; c < (int)regions.size(); ++c)
      if (regions[c].count(tlc_idx)) {
        region_of[v] = c;
        break;
      }
  }
  return (i64)regions.size();
}

// =============================================================================
// Colour path: GraCo colour morphology and the grey / Otsu prep
// =============================================================================
// GraCo::morph_c_base (graco.hpp:301-351): over the closed neighbourhood, erode takes the
// colour of the neighbour with the smallest r+g+b (starting from white), dilate the largest
// (starting from black); the update is strict, so among equal sums the first neighbour in
// iteration order wins — ascending vertex id here, the canonical stand-in for address order.
// rgb words are 0x00RRGGBB (the alpha byte is ignored and cleared).  op: 3 dilate, 4 erode.
ORC_API void orc_morph_rgb(const u32* off, const u32* adj, i64 n, const u32* rgb, int op, u32* out) {
  auto sum = [](u32 c) { return (int)((c >> 16) & 255) + (int)((c >> 8) & 255) + (int)(c & 255); };
#pragma omp parallel for schedule(static)
  for (i64 v = 0; v < n; ++v) {
    u32 best = op == 4 ? 0x00FFFFFFu : 0u;
    // closed neighbourhood in ascending id: neighbours below v, v itself, neighbours above v
    bool self_done = false;
    auto visit = [&](u32 u) {
      u32 c = rgb[u] & 0x00FFFFFFu;
      if (op == 4 ? sum(c) < sum(best) : sum(c) > sum(best)) best = c;
    };
    for (u32 e = off[v]; e < off[v + 1]; ++e) {
      if (!self_done && adj[e] > (u32)v) {
        visit((u32)v);
        self_done = true;
      }
      visit(adj[e]);
    }
    if (!self_done) visit((u32)v);
    out[v] = best;
  }
}

// cloudGray / cloudRGB2GRAY (main.cpp:16-76) [OpenCV-recall, 3.2]: per point
// cv::cvtColor(CV_RGB2GRAY) on 8-bit data = (r*4899 + g*9617 + b*1868 + 8192) >> 14; with
// otsu != 0 the grey values are binarised at cv::threshold(..., CV_THRESH_OTSU)'s threshold
// (getThreshVal_Otsu_8u): v < t -> 0 else 255 (main.cpp:48-51).  Returns the threshold (or -1).
ORC_API double orc_rgb_to_gray(const u32* rgb, i64 n, int otsu, float* gray) {
  std::vector<unsigned char> g((size_t)n);
  for (i64 i = 0; i < n; ++i) {
    unsigned r = (rgb[i] >> 16) & 255, gg = (rgb[i] >> 8) & 255, b = rgb[i] & 255;
    g[i] = (unsigned char)((r * 4899u + gg * 9617u + b * 1868u + 8192u) >> 14);
  }
  double thr = -1.0;
  if (otsu) {
    const int N = 256;
    i64 h[N] = {0};
    for (i64 i = 0; i < n; ++i) h[g[i]]++;
    double mu = 0, scale = 1. / (double)n;
    for (int i = 0; i < N; ++i) mu += i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < N; ++i) {
      double p_i = h[i] * scale;
      mu1 *= q1;
      q1 += p_i;
      double q2 = 1. - q1;
      if (std::min(q1, q2) < FLT_EPSILON || std::max(q1, q2) > 1. - FLT_EPSILON) continue;
      mu1 = (mu1 + i * p_i) / q1;
      double mu2 = (mu - q1 * mu1) / q2;
      double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
      if (sigma > max_sigma) {
        max_sigma = sigma;
        max_val = i;
      }
    }
    thr = max_val;
  }
  for (i64 i = 0; i < n; ++i) {
    float v = (float)g[i];
    if (otsu) v = v < thr ? 0.f : 255.f;
    gray[i] = v;
  }
  return thr;
}
