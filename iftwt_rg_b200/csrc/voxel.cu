// voxel.cu — graph builder (A): the reference's own graph, plus the graph morphology
// and regional-minima seeds that run on it.
//
// Replaces Graph::computeCloudResolution (graph.hpp:228-259), Graph::initialize /
// OctreePointCloudAdjacency (graph.hpp:293-308), Graph::optimizeGraph (graph.hpp:269-284),
// Graph::getCloud (graph.hpp:177-180), Graph::morph_base / morph_gradient
// (graph.hpp:323-371, 147-163) and Graph::getRegmin (graph.hpp:209-226).
//
// As in the reference, building the voxel graph REPLACES the ctx's cloud by the cloud of
// voxel centres ("cloud_.reset()", graph.hpp:307): every later stage (normals, region
// growing, IFT) runs on the graph's vertices, exactly as main.cpp:99-123 runs on cloud_g.
// Vertices are numbered in ascending Morton order of the octree key (x = most significant
// axis bit) — the canonical stand-in for the reference's address order.  The 26-adjacency
// is stored as fixed-width rows (26 entries, missing = none) in g_fwd with an empty
// reverse-arc list: it is symmetric by construction, so the IFT kernels run unchanged.
#include <cub/cub.cuh>
#include <float.h>
#include <math.h>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

namespace {

// term t = sqrtf(d2) split exactly in binary64: A = floor(t 2^24), B = floor((t 2^24 - A) 2^40)
__global__ void resolution_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2, u32 n,
                                  unsigned long long* __restrict__ acc) {
  unsigned long long a = 0, b = 0, cnt = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (nbr[(size_t)i * 2 + 1] == SID_NONE) continue;
    double t = (double)sqrtf(nd2[(size_t)i * 2 + 1]) * 16777216.0;
    double fa = floor(t);
    double fb = floor((t - fa) * 1099511627776.0);
    a += (unsigned long long)fa;
    b += (unsigned long long)fb;
    cnt += 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&acc[0], a);
    atomicAdd(&acc[1], b);
    atomicAdd(&acc[2], cnt);
  }
}

// OctreePointCloud::genOctreeKeyforPoint: key = (unsigned)((p - min') / res), binary64
__global__ void voxel_key_kernel(const float4* __restrict__ pts, u32 n, double mx, double my, double mz,
                                 double res, u64* __restrict__ keys) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  u32 kx = (u32)(((double)p.x - mx) / res);
  u32 ky = (u32)(((double)p.y - my) / res);
  u32 kz = (u32)(((double)p.z - mz) / res);
  keys[i] = morton3(kx, ky, kz);
}
// genLeafNodeCenterFromOctreeKey: centre = float((key + 0.5) * res + min')
__global__ void voxel_centre_kernel(const u64* __restrict__ vkeys, u32 nv, double mx, double my, double mz,
                                    double res, float4* __restrict__ cent) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  u32 kx, ky, kz;
  demorton3(vkeys[v], kx, ky, kz);
  float x = (float)(((double)kx + 0.5) * res + mx);
  float y = (float)(((double)ky + 0.5) * res + my);
  float z = (float)(((double)kz + 0.5) * res + mz);
  cent[v] = make_float4(x, y, z, 0.f);
}
__global__ void voxel_snap_intensity_kernel(const u32* __restrict__ sid, const u32* __restrict__ perm,
                                            const float4* __restrict__ pts_in, u32 nv, float4* __restrict__ cent) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  u32 s = sid[v];
  cent[v].w = s == SID_NONE ? 0.f : pts_in[perm[s]].w;
}
__global__ void voxel_hash_kernel(const u64* __restrict__ vkeys, u32 nv, u64* __restrict__ hk, u32* __restrict__ hv,
                                  u32 hmask) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  u64 key = vkeys[i];
  u32 slot = hash64(key) & hmask;
  while (true) {
    u64 prev = atomicCAS((unsigned long long*)&hk[slot], (unsigned long long)IFTWT_EMPTY_KEY,
                         (unsigned long long)key);
    if (prev == IFTWT_EMPTY_KEY || prev == key) {
      hv[slot] = i;
      return;
    }
    slot = (slot + 1) & hmask;
  }
}
// OctreePointCloudAdjacency::computeNeighbors: offsets {-1,0,1}^3 clipped to [0, max_key]
__global__ void voxel_rows_kernel(const u64* __restrict__ vkeys, u32 nv, u32 max_key, const u64* __restrict__ hk,
                                  const u32* __restrict__ hv, u32 hmask, u32* __restrict__ rows) {
  const u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= nv) return;
  u32 kx, ky, kz;
  demorton3(vkeys[v], kx, ky, kz);
  u32 out = SID_NONE;
  if (lane < 27 && lane != 13) {
    int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
    long long x = (long long)kx + dx, y = (long long)ky + dy, z = (long long)kz + dz;
    if (x >= 0 && y >= 0 && z >= 0 && x <= max_key && y <= max_key && z <= max_key) {
      u64 key = morton3((u32)x, (u32)y, (u32)z);
      u32 slot = hash64(key) & hmask;
      while (true) {
        u64 kk = hk[slot];
        if (kk == key) {
          out = hv[slot];
          break;
        }
        if (kk == IFTWT_EMPTY_KEY) break;
        slot = (slot + 1) & hmask;
      }
    }
  }
  // 26 slots: lanes 0..12 -> 0..12, lanes 14..26 -> 13..25
  if (lane < 27 && lane != 13) rows[(size_t)v * 26 + (lane < 13 ? lane : lane - 1)] = out;
}
// canonical voxel ids -> sorted space of the new cloud
__global__ void voxel_rows_to_sorted_kernel(const u32* __restrict__ rows, const u32* __restrict__ inv, u32 nv,
                                            u32* __restrict__ fwd, unsigned long long* __restrict__ n_arcs) {
  const u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= nv) return;
  u32 r = lane < 26 ? rows[(size_t)v * 26 + lane] : SID_NONE;
  unsigned m = __ballot_sync(0xffffffffu, r != SID_NONE);
  if (lane < 26) fwd[(size_t)inv[v] * 26 + lane] = r == SID_NONE ? SID_NONE : inv[r];
  if (lane == 0 && m) atomicAdd(n_arcs, (unsigned long long)__popc(m));
}

// ---- morphology / regmin on the generic graph (rows + reverse arcs), closed neighbourhood ----
struct GraphView {
  const u32* fwd;
  int k;
  const u32* roff;
  const u32* radj;
};
template <typename F>
__device__ __forceinline__ void for_each_neighbour(const GraphView& g, u32 s, F f) {
  for (int j = 0; j < g.k; ++j) {
    u32 u = g.fwd[(size_t)s * g.k + j];
    if (u != SID_NONE && u != s) f(u);
  }
  for (u32 a = g.roff[s]; a < g.roff[s + 1]; ++a) f(g.radj[a]);
}
__device__ __forceinline__ float morph_one(const GraphView& g, const float4* __restrict__ pts_in,
                                           const u32* __restrict__ perm, u32 s, int op) {
  const float self = pts_in[perm[s]].w;
  float r;
  if (op == 4) {  // erode (graph.hpp:350-355)
    r = 255.f;
    if (self < r) r = self;
    for_each_neighbour(g, s, [&](u32 u) { float t = pts_in[perm[u]].w; if (t < r) r = t; });
  } else if (op == 3) {  // dilate (graph.hpp:356-361)
    r = 0.f;
    if (self > r) r = self;
    for_each_neighbour(g, s, [&](u32 u) { float t = pts_in[perm[u]].w; if (t > r) r = t; });
  } else if (op == 1) {  // bin_erode (graph.hpp:340-344)
    bool hit = self == 0.f;
    for_each_neighbour(g, s, [&](u32 u) { hit = hit || pts_in[perm[u]].w == 0.f; });
    r = hit ? 0.f : 255.f;
  } else {  // bin_dilate, with the shadowed-local bug of graph.hpp:345-349 fixed (App. B #3)
    bool hit = self == 255.f;
    for_each_neighbour(g, s, [&](u32 u) { hit = hit || pts_in[perm[u]].w == 255.f; });
    r = hit ? 255.f : 0.f;
  }
  return r;
}
__global__ void morph_kernel(GraphView g, const float4* __restrict__ pts_in, const u32* __restrict__ perm, u32 n,
                             int op, float* __restrict__ out_orig) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  float r = op == 5 ? morph_one(g, pts_in, perm, s, 3) - morph_one(g, pts_in, perm, s, 4)
                    : morph_one(g, pts_in, perm, s, op);
  out_orig[perm[s]] = r;
}
__global__ void regmin_kernel(GraphView g, const float4* __restrict__ pts_in, const u32* __restrict__ perm, u32 n,
                              unsigned char* __restrict__ flag_orig) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float val = pts_in[perm[s]].w;
  bool is_min = val == 0.f;
  if (!is_min) {
    float mn = val;
    for_each_neighbour(g, s, [&](u32 u) { mn = fminf(mn, pts_in[perm[u]].w); });
    is_min = val == mn;
  }
  flag_orig[perm[s]] = is_min ? 1 : 0;
}


// ---- colour path: GraCo::morph_c_base (graco.hpp:301-351) and the grey / Otsu prep (main.cpp:16-76) ----
// The cloud's 4th channel carries PCL's packed rgb float (bits 0xAARRGGBB); it is only ever
// moved, never computed on.  Among equal r+g+b the reference keeps the first neighbour in
// (address) iteration order; the canonical choice is the smallest ORIGINAL vertex id, which
// makes the result the order-free min / max of the key (sum, id).
__device__ __forceinline__ int rgb_sum(u32 c) { return (int)((c >> 16) & 255) + (int)((c >> 8) & 255) + (int)(c & 255); }
__global__ void morph_rgb_kernel(GraphView g, const float4* __restrict__ pts_in, const u32* __restrict__ perm, u32 n,
                                 int op, u32* __restrict__ out_orig) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const bool erode = op == 4;
  u32 best = erode ? 0x00FFFFFFu : 0u;
  int best_sum = rgb_sum(best);
  u32 best_id = 0xFFFFFFFFu;
  auto visit = [&](u32 u) {
    const u32 o = perm[u];
    const u32 c = __float_as_uint(pts_in[o].w) & 0x00FFFFFFu;
    const int sm = rgb_sum(c);
    // strict improvement over the initial white / black; among real candidates (sum, id) order
    const bool better = erode ? (sm < best_sum || (sm == best_sum && best_id != 0xFFFFFFFFu && o < best_id))
                              : (sm > best_sum || (sm == best_sum && best_id != 0xFFFFFFFFu && o < best_id));
    if (better) {
      best = c;
      best_sum = sm;
      best_id = o;
    }
  };
  visit(s);
  for_each_neighbour(g, s, visit);
  out_orig[perm[s]] = best;
}
// cv::cvtColor(CV_RGB2GRAY), 8-bit fixed point [OpenCV-recall]; also the 256-bin histogram
__global__ void rgb_gray_kernel(const float4* __restrict__ pts_in, u32 n, unsigned char* __restrict__ gray,
                                unsigned long long* __restrict__ hist) {
  __shared__ u32 sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 c = __float_as_uint(pts_in[i].w);
    const u32 r = (c >> 16) & 255, gg = (c >> 8) & 255, b = c & 255;
    const u32 v = (r * 4899u + gg * 9617u + b * 1868u + 8192u) >> 14;
    gray[i] = (unsigned char)v;
    atomicAdd(&sh[v], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
__global__ void gray_store_kernel(const unsigned char* __restrict__ gray, u32 n, int otsu, float thr,
                                  float4* __restrict__ pts_in) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = (float)gray[i];
  if (otsu) v = v < thr ? 0.f : 255.f;  // main.cpp:48-51
  pts_in[i].w = v;
}

GraphView graph_view(iftwt_ctx* c) {
  GraphView g;
  g.fwd = c->graph_shares_nbr ? c->nbr.as<u32>() : c->g_fwd.as<u32>();
  g.k = c->graph_k;
  g.roff = c->g_off.as<u32>();
  g.radj = c->g_adj.as<u32>();
  return g;
}

}  // namespace

int cloud_resolution_run(iftwt_ctx* c, double overlap, double* res) {
  TRY(knn_run(c, 2));
  const u32 n = (u32)c->n;
  unsigned long long* acc = (unsigned long long*)(c->scalars.as<int>() + 116);  // 3 x u64
  CU_TRY(c, cudaMemsetAsync(acc, 0, 24, c->stream));
  LAUNCH(c, resolution_kernel, c->sm_count * 8, 256, 0, c->nbr.as<u32>(), c->nd2.as<float>(), n, acc);
  CHECK_LAUNCH(c);
  unsigned long long* h = (unsigned long long*)((int*)c->h_scalars + 116);
  CU_TRY(c, cudaMemcpyAsync(h, acc, 24, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  double r = 0.0;
  if (h[2] != 0) {
    long double tot = (long double)h[0] * 0x1p-24L + (long double)h[1] * 0x1p-64L;
    r = (double)(tot / (long double)h[2]);
    r *= overlap;  // graph.hpp:256
  }
  *res = r;
  return IFTWT_OK;
}

int voxel_graph_run(iftwt_ctx* c, double resolution, int k_hint) {
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  if (!(resolution > 0.0)) return ctx_fail(c, IFTWT_ERR_INVALID, "voxel size must be positive");
  if (!c->grid_valid) TRY(grid_build(c, 8));
  StageTimer st(c, IFTWT_STAGE_GRAPH);
  const u32 n = (u32)c->n;
  // ---- OctreePointCloud::getKeyBitSize, binary64 on the host ----
  const float minValue = FLT_EPSILON;
  double mn[3], mx[3];
  unsigned mk[3];
  for (int a = 0; a < 3; ++a) {
    mn[a] = c->bb_lo[a];
    mx[a] = c->bb_hi[a];
    mk[a] = (unsigned)ceil((mx[a] - mn[a] - minValue) / resolution);
  }
  unsigned max_voxels = mk[0] > mk[1] ? mk[0] : mk[1];
  if (mk[2] > max_voxels) max_voxels = mk[2];
  if (max_voxels < 2) max_voxels = 2;
  unsigned depth = (unsigned)ceil(log((double)max_voxels) / log(2.0) - minValue);
  if (depth > 21) return ctx_fail(c, IFTWT_ERR_UNSUPPORTED, "octree depth %u exceeds 21 (63-bit Morton keys)", depth);
  const double side = (double)(1u << depth) * resolution;
  for (int a = 0; a < 3; ++a) {
    double over = (side - (mx[a] - mn[a])) / 2.0;
    if (over > minValue) {
      mn[a] -= over;
      mx[a] += over;
    }
  }
  const u32 max_key = (1u << depth) - 1u;
  // ---- occupied voxels, ascending Morton ----
  ENSURE(c, c->keys, (size_t)n * 8);
  ENSURE(c, c->keys_alt, (size_t)n * 8);
  ENSURE(c, c->vx_keys, (size_t)n * 8);
  u64* k0 = c->keys.as<u64>();
  u64* k1 = c->keys_alt.as<u64>();
  LAUNCH(c, voxel_key_kernel, div_up(n, 256), 256, 0, c->pts_in.as<float4>(), n, mn[0], mn[1], mn[2], resolution, k0);
  CHECK_LAUNCH(c);
  u32* d_nv = (u32*)(c->scalars.as<int>() + 124);
  {
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, (int)n, 0, 3 * (int)depth, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(c->tmp.p, tb, k0, k1, (int)n, 0, 3 * (int)depth, c->stream));
    size_t tb2 = 0;
    CU_TRY(c, cub::DeviceSelect::Unique(nullptr, tb2, k1, c->vx_keys.as<u64>(), d_nv, (int)n, c->stream));
    ENSURE(c, c->tmp, tb2);
    CU_TRY(c, cub::DeviceSelect::Unique(c->tmp.p, tb2, k1, c->vx_keys.as<u64>(), d_nv, (int)n, c->stream));
    c->stats.kernel_launches += 4 + (3 * depth + 7) / 8;
  }
  u32* hs = (u32*)c->h_scalars;
  CU_TRY(c, cudaMemcpyAsync(hs + 124, d_nv, 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const u32 nv = hs[124];
  if (nv == 0 || nv > n) return ctx_fail(c, IFTWT_ERR_CUDA, "voxelisation failed");
  const u64* vkeys = c->vx_keys.as<u64>();
  // ---- centres + snapped intensity (optimizeGraph) ----
  ENSURE(c, c->vx_cent, (size_t)nv * 16);
  ENSURE(c, c->out_a, (size_t)nv * 4);
  float4* cent = c->vx_cent.as<float4>();
  LAUNCH(c, voxel_centre_kernel, div_up(nv, 256), 256, 0, vkeys, nv, mn[0], mn[1], mn[2], resolution, cent);
  TRY(knn_query_run(c, cent, nv, 1, c->out_a.as<u32>(), nullptr));
  LAUNCH(c, voxel_snap_intensity_kernel, div_up(nv, 256), 256, 0, c->out_a.as<u32>(), c->perm.as<u32>(),
         c->pts_in.as<float4>(), nv, cent);
  // ---- 26-adjacency in canonical ids ----
  u32 hsize = 1024;
  while (hsize < 2 * nv) hsize <<= 1;
  ENSURE(c, c->vx_hk, (size_t)hsize * 8);
  ENSURE(c, c->vx_hv, (size_t)hsize * 4);
  ENSURE(c, c->vx_rows, (size_t)nv * 26 * 4);
  CU_TRY(c, cudaMemsetAsync(c->vx_hk.p, 0xFF, (size_t)hsize * 8, c->stream));
  LAUNCH(c, voxel_hash_kernel, div_up(nv, 256), 256, 0, vkeys, nv, c->vx_hk.as<u64>(), c->vx_hv.as<u32>(), hsize - 1);
  LAUNCH(c, voxel_rows_kernel, div_up((size_t)nv * 32, 256), 256, 0, vkeys, nv, max_key, c->vx_hk.as<u64>(),
         c->vx_hv.as<u32>(), hsize - 1, c->vx_rows.as<u32>());
  CHECK_LAUNCH(c);
  // ---- the graph's vertices become the cloud (graph.hpp:307) ----
  ENSURE(c, c->pts_in, (size_t)nv * 16);
  CU_TRY(c, cudaMemcpyAsync(c->pts_in.p, cent, (size_t)nv * 16, cudaMemcpyDeviceToDevice, c->stream));
  ctx_invalidate(c);
  c->n = nv;
  c->stats.n_points = nv;
  c->voxel_size = resolution;
  TRY(grid_build(c, k_hint));
  ENSURE(c, c->g_fwd, (size_t)nv * 26 * 4);
  ENSURE(c, c->g_off, ((size_t)nv + 1) * 4);
  ENSURE(c, c->g_adj, 16);
  unsigned long long* d_arcs = (unsigned long long*)(c->scalars.as<int>() + 52);
  CU_TRY(c, cudaMemsetAsync(d_arcs, 0, 8, c->stream));
  CU_TRY(c, cudaMemsetAsync(c->g_off.p, 0, ((size_t)nv + 1) * 4, c->stream));
  LAUNCH(c, voxel_rows_to_sorted_kernel, div_up((size_t)nv * 32, 256), 256, 0, c->vx_rows.as<u32>(),
         c->inv.as<u32>(), nv, c->g_fwd.as<u32>(), d_arcs);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaMemcpyAsync(hs + 52, d_arcs, 8, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  c->n_arcs = *(unsigned long long*)(hs + 52);
  c->n_rev = 0;
  c->graph_k = 26;
  c->graph_kind = 2;
  c->graph_shares_nbr = false;
  c->stats.n_arcs = c->n_arcs;
  return IFTWT_OK;
}


int morph_rgb_run(iftwt_ctx* c, int op, u32* d_out_orig) {
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  if (op != 3 && op != 4) return ctx_fail(c, IFTWT_ERR_INVALID, "colour morphology has dilate (3) and erode (4) only");
  const u32 n = (u32)c->n;
  LAUNCH(c, morph_rgb_kernel, div_up(n, 128), 128, 0, graph_view(c), c->pts_in.as<float4>(), c->perm.as<u32>(), n, op,
         d_out_orig);
  CHECK_LAUNCH(c);
  return IFTWT_OK;
}

// packed rgb -> grey (optionally binarised at the Otsu threshold) in the cloud's 4th channel
int rgb_to_gray_run(iftwt_ctx* c, int otsu, double* threshold) {
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  const u32 n = (u32)c->n;
  ENSURE(c, c->scalars, 4096);
  ENSURE(c, c->out_b, (size_t)n + 256 * 8 + 16);
  unsigned char* gray = c->out_b.as<unsigned char>() + 256 * 8;
  unsigned long long* hist = c->out_b.as<unsigned long long>();
  CU_TRY(c, cudaMemsetAsync(hist, 0, 256 * 8, c->stream));
  LAUNCH(c, rgb_gray_kernel, c->sm_count * 4, 256, 0, c->pts_in.as<float4>(), n, gray, hist);
  CHECK_LAUNCH(c);
  double thr = -1.0;
  if (otsu) {
    unsigned long long h[256];
    CU_TRY(c, cudaMemcpyAsync(h, hist, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    // getThreshVal_Otsu_8u [OpenCV-recall, 3.2]
    double mu = 0, scale = 1. / (double)n;
    for (int i = 0; i < 256; ++i) mu += i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; ++i) {
      double p_i = h[i] * scale;
      mu1 *= q1;
      q1 += p_i;
      double q2 = 1. - q1;
      if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
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
  LAUNCH(c, gray_store_kernel, div_up(n, 256), 256, 0, gray, n, otsu, (float)thr, c->pts_in.as<float4>());
  CHECK_LAUNCH(c);
  if (threshold) *threshold = thr;
  // the values every later stage reads have changed; the geometry (grid, k-NN, graph) has not
  c->ift_valid = false;
  c->rag_valid = false;
  return IFTWT_OK;
}

int morph_run(iftwt_ctx* c, int op, float* d_out_orig) {
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  if (op < 1 || op > 5) return ctx_fail(c, IFTWT_ERR_INVALID, "unknown morphology operator %d", op);
  const u32 n = (u32)c->n;
  LAUNCH(c, morph_kernel, div_up(n, 128), 128, 0, graph_view(c), c->pts_in.as<float4>(), c->perm.as<u32>(), n, op,
         d_out_orig);
  CHECK_LAUNCH(c);
  return IFTWT_OK;
}

// regional-minima seeds, ascending original id, left in c->seeds
int regmin_run(iftwt_ctx* c) {
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  StageTimer st(c, IFTWT_STAGE_SEEDS);
  const u32 n = (u32)c->n;
  ENSURE(c, c->out_b, (size_t)n);
  ENSURE(c, c->seeds, ((size_t)n + 1) * 4);
  unsigned char* flag = c->out_b.as<unsigned char>();
  LAUNCH(c, regmin_kernel, div_up(n, 128), 128, 0, graph_view(c), c->pts_in.as<float4>(), c->perm.as<u32>(), n, flag);
  CHECK_LAUNCH(c);
  u32* d_cnt = (u32*)(c->scalars.as<int>() + 125);
  cub::CountingInputIterator<u32> ids(0);
  size_t tb = 0;
  CU_TRY(c, cub::DeviceSelect::Flagged(nullptr, tb, ids, flag, c->seeds.as<u32>(), d_cnt, (int)n, c->stream));
  ENSURE(c, c->tmp, tb);
  CU_TRY(c, cub::DeviceSelect::Flagged(c->tmp.p, tb, ids, flag, c->seeds.as<u32>(), d_cnt, (int)n, c->stream));
  c->stats.kernel_launches += 2;
  u32* hs = (u32*)c->h_scalars;
  CU_TRY(c, cudaMemcpyAsync(hs + 125, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  c->n_seeds = hs[125];
  c->seeds_valid = true;
  c->ift_valid = false;
  c->stats.n_seeds = c->n_seeds;
  return IFTWT_OK;
}
