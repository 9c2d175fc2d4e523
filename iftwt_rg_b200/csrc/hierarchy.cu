// hierarchy.cu — K8: region adjacency / saddle table, region attributes, and the merge of
// the IFT regions down to num_regions.
//
// Stands in for the region bookkeeping interleaved with the heap loop (r_info and the
// first-contact "MST", ift.hpp:5-23, 155-168, 207-243) and for Cluster (cluster.hpp:23-229).
// Both depend on the sequential conquest order and on container address order in the
// reference, so no order-free restatement can reproduce them literally (hard part
// 7; App. B #9).  The product definition, implemented here and restated by the oracle
// (oracle: orc_region_adjacency / orc_region_info / orc_cluster_saddle) is:
//   * saddle(a, b) = min over arcs (u, v) with label[u] = a != label[v] = b, both conquered,
//     of max(cost[u], cost[v])                                         (watershed pass value)
//   * area(a) = conquered vertices labelled a, height(a) = max vertex value over them
//     (the order-free part of r_info, ift.hpp:11-16, 140-141)
//   * clusters = single-linkage merge along the saddle minimum spanning forest: Kruskal in
//     ascending (saddle, a, b) until num_regions clusters remain; a cluster is named by its
//     smallest seed id.
// The arc scan is a hash-table atomicMin on the GPU; the Kruskal over the (small) region
// graph runs on the host side of the library, as the reference's Cluster does.
#include <cub/cub.cuh>

#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

namespace {

struct GraphView {
  const u32* fwd;
  int k;
  const u32* roff;
  const u32* radj;
};
GraphView graph_view(iftwt_ctx* c) {
  GraphView g;
  g.fwd = c->graph_shares_nbr ? c->nbr.as<u32>() : c->g_fwd.as<u32>();
  g.k = c->graph_k;
  g.roff = c->g_off.as<u32>();
  g.radj = c->g_adj.as<u32>();
  return g;
}

__device__ __forceinline__ bool boundary_arc(u64 su, u64 sv, u64& key, u32& saddle) {
  const u32 cu = (u32)(su >> 32), cv = (u32)(sv >> 32);
  const u32 lu = (u32)su, lv = (u32)sv;
  if (cv >= IFTWT_MAX_COST || lu >= lv) return false;  // each undirected boundary edge once: lu < lv
  key = ((u64)lu << 32) | (u64)lv;
  saddle = max(cu, cv);
  return true;
}

// pass 0: count boundary arcs (upper bound on distinct pairs).  pass 1: hash atomicMin.
__global__ void rag_kernel(GraphView g, const u64* __restrict__ res, u32 n, int pass, unsigned long long* count,
                           u64* __restrict__ hk, u32* __restrict__ hv, u32 hmask) {
  u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  u32 local = 0;
  if (u < n) {
    const u64 su = res[u];
    if ((u32)(su >> 32) < IFTWT_MAX_COST) {
      auto visit = [&](u32 v) {
        u64 key;
        u32 sad;
        if (!boundary_arc(su, res[v], key, sad)) return;
        if (pass == 0) {
          ++local;
          return;
        }
        u32 slot = hash64(key) & hmask;
        while (true) {
          u64 prev = atomicCAS((unsigned long long*)&hk[slot], (unsigned long long)IFTWT_EMPTY_KEY,
                               (unsigned long long)key);
          if (prev == IFTWT_EMPTY_KEY || prev == key) {
            atomicMin(&hv[slot], sad);
            return;
          }
          slot = (slot + 1) & hmask;
        }
      };
      for (int j = 0; j < g.k; ++j) {
        u32 v = g.fwd[(size_t)u * g.k + j];
        if (v != SID_NONE && v != u) visit(v);
      }
      for (u32 a = g.roff[u]; a < g.roff[u + 1]; ++a) visit(g.radj[a]);
    }
  }
  if (pass == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
  }
}
__global__ void rag_compact_kernel(const u64* __restrict__ hk, const u32* __restrict__ hv, u32 hsize,
                                   u32* __restrict__ counter, u64* __restrict__ keys, u32* __restrict__ vals) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hsize) return;
  u64 k = hk[i];
  if (k == IFTWT_EMPTY_KEY) return;
  u32 p = atomicAdd(counter, 1u);
  keys[p] = k;
  vals[p] = hv[i];
}
__device__ __forceinline__ u32 seed_ordinal(const u32* __restrict__ seeds, u32 ns, u32 s) {
  u32 lo = 0, hi = ns;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (seeds[mid] < s) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__global__ void region_info_kernel(const u64* __restrict__ res, const float* __restrict__ wf, u32 n,
                                   const u32* __restrict__ seeds, u32 ns, u32* __restrict__ area,
                                   u32* __restrict__ height_bits) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const u64 s = res[v];
  if ((u32)(s >> 32) >= IFTWT_MAX_COST) return;
  const u32 o = seed_ordinal(seeds, ns, (u32)s);
  atomicAdd(&area[o], 1u);
  atomicMax(&height_bits[o], __float_as_uint(wf[v]));  // values are >= 0: bit order = value order
}
__global__ void cluster_map_kernel(const u64* __restrict__ res, const u32* __restrict__ perm, u32 n,
                                   const u32* __restrict__ seeds, u32 ns, const u32* __restrict__ cluster_seed,
                                   u32* __restrict__ out_orig) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const u64 s = res[v];
  out_orig[perm[v]] = (u32)(s >> 32) < IFTWT_MAX_COST ? cluster_seed[seed_ordinal(seeds, ns, (u32)s)] : SID_NONE;
}

// sorted, de-duplicated seed list (regions are named by their seed)
int build_region_list(iftwt_ctx* c) {
  const size_t ns = c->n_seeds;
  ENSURE(c, c->h_seeds, (ns + 1) * 4 * 2);
  u32* sorted = c->h_seeds.as<u32>();
  u32* uniq = sorted + (ns + 1);
  u32* d_cnt = (u32*)(c->scalars.as<int>() + 126);
  size_t tb = 0;
  CU_TRY(c, cub::DeviceRadixSort::SortKeys(nullptr, tb, c->seeds.as<u32>(), sorted, (int)ns, 0, 32, c->stream));
  ENSURE(c, c->tmp, tb);
  CU_TRY(c, cub::DeviceRadixSort::SortKeys(c->tmp.p, tb, c->seeds.as<u32>(), sorted, (int)ns, 0, 32, c->stream));
  size_t tb2 = 0;
  CU_TRY(c, cub::DeviceSelect::Unique(nullptr, tb2, sorted, uniq, d_cnt, (int)ns, c->stream));
  ENSURE(c, c->tmp, tb2);
  CU_TRY(c, cub::DeviceSelect::Unique(c->tmp.p, tb2, sorted, uniq, d_cnt, (int)ns, c->stream));
  u32* hs = (u32*)c->h_scalars;
  CU_TRY(c, cudaMemcpyAsync(hs + 126, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  c->n_regions = hs[126];
  return IFTWT_OK;
}

}  // namespace

// pairs sorted by (a, b) in c->h_pairs: keys u64 [n_pairs] then saddles u32 [n_pairs]
int region_adjacency_run(iftwt_ctx* c) {
  if (!c->ift_valid) return ctx_fail(c, IFTWT_ERR_STATE, "no IFT result on the device");
  if (c->rag_valid) return IFTWT_OK;
  TRY(ift_res_ensure(c));
  const u32 n = (u32)c->n;
  GraphView g = graph_view(c);
  const u64* res = c->ift_res.as<u64>();
  unsigned long long* d_count = (unsigned long long*)(c->scalars.as<int>() + 52);
  CU_TRY(c, cudaMemsetAsync(d_count, 0, 8, c->stream));
  LAUNCH(c, rag_kernel, div_up(n, 128), 128, 0, g, res, n, 0, d_count, (u64*)nullptr, (u32*)nullptr, 0u);
  CHECK_LAUNCH(c);
  u32* hs = (u32*)c->h_scalars;
  CU_TRY(c, cudaMemcpyAsync(hs + 52, d_count, 8, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const size_t n_boundary = *(unsigned long long*)(hs + 52);
  size_t hsize = 1024;
  while (hsize < 2 * n_boundary) hsize <<= 1;
  if (hsize > 0x80000000ull) return ctx_fail(c, IFTWT_ERR_UNSUPPORTED, "region graph too large");
  ENSURE(c, c->h_keys, hsize * 8);
  ENSURE(c, c->h_vals, hsize * 4);
  CU_TRY(c, cudaMemsetAsync(c->h_keys.p, 0xFF, hsize * 8, c->stream));
  CU_TRY(c, cudaMemsetAsync(c->h_vals.p, 0xFF, hsize * 4, c->stream));
  size_t n_pairs = 0;
  if (n_boundary) {
    LAUNCH(c, rag_kernel, div_up(n, 128), 128, 0, g, res, n, 1, d_count, c->h_keys.as<u64>(), c->h_vals.as<u32>(),
           (u32)(hsize - 1));
    u32* d_np = (u32*)(c->scalars.as<int>() + 127);
    CU_TRY(c, cudaMemsetAsync(d_np, 0, 4, c->stream));
    ENSURE(c, c->h_pairs, (n_boundary + 1) * 12 * 2);
    u64* pk = c->h_pairs.as<u64>();
    u64* pk_sorted = pk + (n_boundary + 1);
    u32* pv = (u32*)(pk_sorted + (n_boundary + 1));
    u32* pv_sorted = pv + (n_boundary + 1);
    LAUNCH(c, rag_compact_kernel, div_up(hsize, 256), 256, 0, c->h_keys.as<u64>(), c->h_vals.as<u32>(), (u32)hsize,
           d_np, pk, pv);
    CHECK_LAUNCH(c);
    CU_TRY(c, cudaMemcpyAsync(hs + 127, d_np, 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    n_pairs = hs[127];
    if (n_pairs) {
      size_t tb = 0;
      CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, pk, pk_sorted, pv, pv_sorted, (int)n_pairs, 0, 64,
                                                c->stream));
      ENSURE(c, c->tmp, tb);
      CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, pk, pk_sorted, pv, pv_sorted, (int)n_pairs, 0, 64,
                                                c->stream));
      // keep the sorted copies at the front
      CU_TRY(c, cudaMemcpyAsync(pk, pk_sorted, n_pairs * 8, cudaMemcpyDeviceToDevice, c->stream));
      CU_TRY(c, cudaMemcpyAsync(pv, pv_sorted, n_pairs * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->pairs_vals_offset = (size_t)((char*)pv - (char*)c->h_pairs.p);
  }
  c->n_pairs = n_pairs;
  c->rag_valid = true;
  return IFTWT_OK;
}

// c->h_info: area u32 [n_regions], height f32 [n_regions]; region list in c->h_seeds (second half)
int region_info_run(iftwt_ctx* c) {
  if (!c->ift_valid) return ctx_fail(c, IFTWT_ERR_STATE, "no IFT result on the device");
  TRY(ift_res_ensure(c));
  TRY(build_region_list(c));
  const u32 nr = (u32)c->n_regions;
  ENSURE(c, c->h_info, ((size_t)nr + 1) * 8);
  CU_TRY(c, cudaMemsetAsync(c->h_info.p, 0, ((size_t)nr + 1) * 8, c->stream));
  const u32* uniq = c->h_seeds.as<u32>() + (c->n_seeds + 1);
  if (nr)
    LAUNCH(c, region_info_kernel, div_up(c->n, 256), 256, 0, c->ift_res.as<u64>(), c->ift_wf.as<float>(), (u32)c->n,
           uniq, nr, c->h_info.as<u32>(), c->h_info.as<u32>() + (nr + 1));
  CHECK_LAUNCH(c);
  return IFTWT_OK;
}

int cluster_run(iftwt_ctx* c, int num_regions, u32* d_cluster_orig, size_t* n_clusters) {
  TRY(ift_res_ensure(c));
  TRY(region_adjacency_run(c));
  TRY(build_region_list(c));
  const size_t S = c->n_regions, m = c->n_pairs;
  const u32* d_uniq = c->h_seeds.as<u32>() + (c->n_seeds + 1);
  std::vector<u32> seeds(S), sad(m);
  std::vector<u64> keys(m);
  if (S) CU_TRY(c, cudaMemcpyAsync(seeds.data(), d_uniq, S * 4, cudaMemcpyDeviceToHost, c->stream));
  if (m) {
    CU_TRY(c, cudaMemcpyAsync(keys.data(), c->h_pairs.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(sad.data(), (char*)c->h_pairs.p + c->pairs_vals_offset, m * 4, cudaMemcpyDeviceToHost,
                              c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  // Kruskal on the region graph: ascending (saddle, a, b); keys are already (a, b) ascending
  std::vector<u32> order(m);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](u32 i, u32 j) { return sad[i] < sad[j]; });
  std::vector<u32> parent(S);
  std::iota(parent.begin(), parent.end(), 0u);
  auto find = [&](u32 x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  };
  auto ord = [&](u32 s) { return (u32)(std::lower_bound(seeds.begin(), seeds.end(), s) - seeds.begin()); };
  size_t clusters = S;
  for (u32 idx : order) {
    if (clusters <= (size_t)(num_regions < 0 ? 0 : num_regions)) break;
    u32 a = find(ord((u32)(keys[idx] >> 32))), b = find(ord((u32)keys[idx]));
    if (a == b) continue;
    if (a < b) parent[b] = a;
    else parent[a] = b;
    --clusters;
  }
  std::vector<u32> cluster_seed(S);
  for (size_t i = 0; i < S; ++i) cluster_seed[i] = seeds[find((u32)i)];
  if (n_clusters) *n_clusters = clusters;
  ENSURE(c, c->h_info, (S + 1) * 8);
  if (S) CU_TRY(c, cudaMemcpyAsync(c->h_info.p, cluster_seed.data(), S * 4, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, cluster_map_kernel, div_up(c->n, 256), 256, 0, c->ift_res.as<u64>(), c->perm.as<u32>(), (u32)c->n, d_uniq,
         (u32)S, c->h_info.as<u32>(), d_cluster_orig);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaStreamSynchronize(c->stream));  // cluster_seed must outlive the copy
  return IFTWT_OK;
}
