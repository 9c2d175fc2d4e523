// iftwt_batch_* from a plain C++ host (no torch, no CUDA headers): a batch of small clouds through 3 lanes
// equals the same clouds through one ctx, one at a time (BASELINE config 5's entry point).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "iftwt.h"

#define CHECK(c)                                                          \
  do {                                                                    \
    if (!(c)) {                                                           \
      std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); \
      return 1;                                                           \
    }                                                                     \
  } while (0)

static std::vector<float> staircase(int n, unsigned seed) {
  std::vector<float> p((size_t)n * 4);
  unsigned s = seed;
  auto u = [&s]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
  for (int i = 0; i < n; ++i) {
    const int step = i % 8;
    float x, y, z, w;
    if ((i / 8) % 3 != 2) {
      x = 0.5f * ((float)step + u()); y = 2.f * u(); z = 0.25f * (float)step + 0.001f * (u() - 0.5f);
      w = (float)(40 + 10 * step + (int)(4.f * u()));
    } else {
      x = 0.5f * (float)step + 0.001f * (u() - 0.5f); y = 2.f * u(); z = 0.25f * ((float)step - u());
      w = (float)(200 - 10 * step + (int)(4.f * u()));
    }
    p[4 * (size_t)i] = x; p[4 * (size_t)i + 1] = y; p[4 * (size_t)i + 2] = z; p[4 * (size_t)i + 3] = w;
  }
  return p;
}

int main() {
  const int B = 5, k = 16;
  std::vector<std::vector<float>> clouds;
  std::vector<const void*> ptrs;
  std::vector<size_t> sizes;
  for (int i = 0; i < B; ++i) {
    clouds.push_back(staircase(40000 + 3000 * i, 777u + (unsigned)i));
    sizes.push_back(clouds.back().size() / 4);
  }
  for (int i = 0; i < B; ++i) ptrs.push_back(clouds[(size_t)i].data());
  iftwt_segment_params p;
  std::memset(&p, 0, sizeof p);
  p.k_graph = p.k_normals = k;
  p.rg.min_cluster_size = 50;
  p.rg.max_cluster_size = 10000;
  p.rg.number_of_neighbours = k;
  p.rg.smoothness_threshold = 3.0f / 180.0f * 3.14159265f;
  p.rg.curvature_threshold = 1.0f;
  // one ctx, one cloud at a time
  std::vector<std::vector<uint32_t>> rc(B), rl(B);
  std::vector<size_t> rs(B);
  iftwt_ctx* ctx = nullptr;
  CHECK(iftwt_create(&ctx, 0) == IFTWT_OK);
  for (int i = 0; i < B; ++i) {
    rc[(size_t)i].resize(sizes[(size_t)i]);
    rl[(size_t)i].resize(sizes[(size_t)i]);
    CHECK(iftwt_set_cloud(ctx, ptrs[(size_t)i], 16, sizes[(size_t)i], 0, 12) == IFTWT_OK);
    CHECK(iftwt_segment(ctx, &p, rc[(size_t)i].data(), rl[(size_t)i].data(), &rs[(size_t)i]) == IFTWT_OK);
  }
  iftwt_destroy(ctx);
  // the batch
  iftwt_batch* bt = nullptr;
  CHECK(iftwt_batch_create(&bt, 0, 3) == IFTWT_OK);
  std::vector<std::vector<uint32_t>> bc(B), bl(B);
  std::vector<uint32_t*> cp, lp;
  for (int i = 0; i < B; ++i) {
    bc[(size_t)i].resize(sizes[(size_t)i]);
    bl[(size_t)i].resize(sizes[(size_t)i]);
    cp.push_back(bc[(size_t)i].data());
    lp.push_back(bl[(size_t)i].data());
  }
  std::vector<size_t> bs(B);
  CHECK(iftwt_batch_segment(bt, ptrs.data(), sizes.data(), B, 0, &p, cp.data(), lp.data(), bs.data()) == IFTWT_OK);
  size_t seeds = 0;
  for (int i = 0; i < B; ++i) {
    CHECK(bs[(size_t)i] == rs[(size_t)i]);
    CHECK(bc[(size_t)i] == rc[(size_t)i] && bl[(size_t)i] == rl[(size_t)i]);
    seeds += bs[(size_t)i];
  }
  CHECK(seeds >= 10);
  CHECK(iftwt_batch_ctx(bt, 2) != nullptr && iftwt_batch_ctx(bt, 3) == nullptr);
  iftwt_batch_destroy(bt);
  CHECK(iftwt_batch_create(&bt, 0, 0) == IFTWT_ERR_INVALID);
  std::printf("batch ok: %d clouds, %zu seeds\n", B, seeds);
  return 0;
}
