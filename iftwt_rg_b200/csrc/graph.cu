// graph.cu — symmetric k-NN graph as CSR (graph builder (B), SURVEY finding 3).
//
// Stands in for Graph::initialize / getAdjacencyList (graph.hpp:293-308, 63-66)
// on the BASELINE k-NN configurations: vertices are the cloud points, u ~ v iff
// v is among the k nearest neighbours of u or u among those of v.  The reverse
// arcs are found without reading the neighbour's list: u is in kNN(v) exactly
// when (d2(u,v), index(u)) <= the key of v's k-th neighbour, and d2 is bitwise
// symmetric, so one 8-byte gather per arc decides mutuality.
#include <cub/cub.cuh>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

__device__ __forceinline__ bool arc_is_mutual(const float4* __restrict__ spts, const float* __restrict__ nd2,
                                              const u64* __restrict__ kth, size_t arc, u32 u, u32 v) {
  u64 key = ((u64)__float_as_uint(nd2[arc]) << 32) | (u64)__float_as_uint(spts[u].w);
  return key <= kth[v];
}

// forward degree: valid, non-self entries of the row
__global__ void graph_fwd_degree_kernel(const u32* __restrict__ nbr, u32 n, int k, u32* __restrict__ fdeg,
                                        u32* __restrict__ deg) {
  u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  u32 c = 0;
  for (int j = 0; j < k; ++j) {
    u32 v = nbr[(size_t)u * k + j];
    c += (v != SID_NONE && v != u) ? 1u : 0u;
  }
  fdeg[u] = c;
  deg[u] = c;
}

// one thread per (u, j): a one-way arc u->v adds a reverse entry to v's row
__global__ void graph_rev_count_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2,
                                       const u64* __restrict__ kth, const float4* __restrict__ spts,
                                       size_t n_entries, int k, u32* __restrict__ deg) {
  size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_entries) return;
  u32 u = (u32)(a / k);
  u32 v = nbr[a];
  if (v == SID_NONE || v == u) return;
  if (!arc_is_mutual(spts, nd2, kth, a, u, v)) atomicAdd(&deg[v], 1u);
}

__global__ void graph_fill_fwd_kernel(const u32* __restrict__ nbr, u32 n, int k, const u32* __restrict__ off,
                                      u32* __restrict__ adj) {
  u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  u32 o = off[u];
  for (int j = 0; j < k; ++j) {
    u32 v = nbr[(size_t)u * k + j];
    if (v != SID_NONE && v != u) adj[o++] = v;
  }
}

__global__ void graph_fill_rev_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2,
                                      const u64* __restrict__ kth, const float4* __restrict__ spts,
                                      size_t n_entries, int k, const u32* __restrict__ off,
                                      const u32* __restrict__ fdeg, u32* __restrict__ cursor,
                                      u32* __restrict__ adj) {
  size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_entries) return;
  u32 u = (u32)(a / k);
  u32 v = nbr[a];
  if (v == SID_NONE || v == u) return;
  if (!arc_is_mutual(spts, nd2, kth, a, u, v)) {
    u32 p = atomicAdd(&cursor[v], 1u);
    adj[off[v] + fdeg[v] + p] = u;
  }
}

int graph_build(iftwt_ctx* c) {
  if (c->knn_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "k-NN lists are not computed");
  if (c->graph_k == c->knn_k) return IFTWT_OK;
  StageTimer st(c, IFTWT_STAGE_GRAPH);
  const u32 n = (u32)c->n;
  const int k = c->knn_k;
  const size_t ne = (size_t)n * k;
  ENSURE(c, c->g_deg, ((size_t)n + 1) * 4 * 2);  // [0..n] total degree / cursor, [n+1..] forward degree
  ENSURE(c, c->g_off, ((size_t)n + 1) * 4);
  u32* deg = c->g_deg.as<u32>();
  u32* fdeg = deg + (n + 1);
  u32* off = c->g_off.as<u32>();
  const u32* nbr = c->nbr.as<u32>();
  const float* nd2 = c->nd2.as<float>();
  const u64* kth = c->kth.as<u64>();
  const float4* spts = c->spts.as<float4>();
  CU_TRY(c, cudaMemsetAsync(deg + n, 0, 4, c->stream));
  LAUNCH(c, graph_fwd_degree_kernel, div_up(n, 256), 256, 0, nbr, n, k, fdeg, deg);
  LAUNCH(c, graph_rev_count_kernel, div_up(ne, 256), 256, 0, nbr, nd2, kth, spts, ne, k, deg);
  CHECK_LAUNCH(c);
  {
    size_t tb = 0;
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, off, (int)(n + 1), c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, deg, off, (int)(n + 1), c->stream));
    c->stats.kernel_launches += 2;
  }
  u32* hs = (u32*)c->h_scalars;
  CU_TRY(c, cudaMemcpyAsync(hs + 52, off + n, 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const size_t arcs = hs[52];
  ENSURE(c, c->g_adj, (arcs + 1) * 4);
  CU_TRY(c, cudaMemsetAsync(deg, 0, (size_t)n * 4, c->stream));  // now the reverse-fill cursor
  LAUNCH(c, graph_fill_fwd_kernel, div_up(n, 256), 256, 0, nbr, n, k, off, c->g_adj.as<u32>());
  LAUNCH(c, graph_fill_rev_kernel, div_up(ne, 256), 256, 0, nbr, nd2, kth, spts, ne, k, off, fdeg, deg,
         c->g_adj.as<u32>());
  CHECK_LAUNCH(c);
  c->n_arcs = arcs;
  c->graph_k = k;
  c->stats.n_arcs = arcs;
  c->ift_valid = false;
  return IFTWT_OK;
}

// ---- export in original index space, rows ascending ----
__global__ void export_degree_kernel(const u32* __restrict__ off, const u32* __restrict__ inv, u32 n,
                                     u32* __restrict__ deg_o) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    deg_o[i] = 0;
    return;
  }
  u32 s = inv[i];
  deg_o[i] = off[s + 1] - off[s];
}
__global__ void export_rows_kernel(const u32* __restrict__ off, const u32* __restrict__ adj,
                                   const u32* __restrict__ inv, const u32* __restrict__ perm, u32 n,
                                   const u32* __restrict__ off_o, u32* __restrict__ adj_o) {
  u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n) return;
  u32 s = inv[gw];
  u32 b = off[s], e = off[s + 1], o = off_o[gw];
  for (u32 t = b + lane; t < e; t += 32) adj_o[o + (t - b)] = perm[adj[t]];
}

int graph_export(iftwt_ctx* c, u32* h_off, u32* h_adj, size_t cap, size_t* n_arcs) {
  const u32 n = (u32)c->n;
  const size_t arcs = c->n_arcs;
  if (n_arcs) *n_arcs = arcs;
  ENSURE(c, c->out_a, ((size_t)n + 1) * 4 * 2);
  u32* deg_o = c->out_a.as<u32>();
  u32* off_o = deg_o + (n + 1);
  LAUNCH(c, export_degree_kernel, div_up((size_t)n + 1, 256), 256, 0, c->g_off.as<u32>(), c->inv.as<u32>(), n,
         deg_o);
  size_t tb = 0;
  CU_TRY(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, deg_o, off_o, (int)(n + 1), c->stream));
  ENSURE(c, c->tmp, tb);
  CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, deg_o, off_o, (int)(n + 1), c->stream));
  if (h_off) CU_TRY(c, cudaMemcpyAsync(h_off, off_o, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (h_adj) {
    if (cap < arcs) {
      cudaStreamSynchronize(c->stream);
      return ctx_fail(c, IFTWT_ERR_CAPACITY, "adjacency needs %zu entries, caller gave %zu", arcs, cap);
    }
    ENSURE(c, c->out_b, (arcs + 1) * 4 * 2);
    u32* tmp_adj = c->out_b.as<u32>();
    u32* srt_adj = tmp_adj + (arcs + 1);
    LAUNCH(c, export_rows_kernel, div_up((size_t)n * 32, 256), 256, 0, c->g_off.as<u32>(), c->g_adj.as<u32>(),
           c->inv.as<u32>(), c->perm.as<u32>(), n, off_o, tmp_adj);
    CHECK_LAUNCH(c);
    size_t tb2 = 0;
    CU_TRY(c, cub::DeviceSegmentedSort::SortKeys(nullptr, tb2, tmp_adj, srt_adj, (int)arcs, (int)n, off_o,
                                                 off_o + 1, c->stream));
    ENSURE(c, c->tmp, tb2);
    CU_TRY(c, cub::DeviceSegmentedSort::SortKeys(c->tmp.p, tb2, tmp_adj, srt_adj, (int)arcs, (int)n, off_o,
                                                 off_o + 1, c->stream));
    CU_TRY(c, cudaMemcpyAsync(h_adj, srt_adj, arcs * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}
