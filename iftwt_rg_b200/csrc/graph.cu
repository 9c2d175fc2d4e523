// graph.cu — symmetric k-NN graph (graph builder (B), SURVEY finding 3).
//
// Stands in for Graph::initialize / getAdjacencyList (graph.hpp:293-308, 63-66)
// on the BASELINE k-NN configurations: vertices are the cloud points, u ~ v iff
// v is among the k nearest neighbours of u or u among those of v.
//
// Device representation (sorted space), chosen so that the forward arcs are never
// copied: adjacency(u) = the k-NN row of u (k entries, self / missing skipped at
// read time) + a CSR list of the REVERSE arcs only (v with u in kNN(v) but v not
// in kNN(u)).  The reverse arcs are found without reading the neighbour's list:
// u is in kNN(v) exactly when (d2(u,v), index(u)) <= the key of v's k-th neighbour,
// and d2 is bitwise symmetric, so one 8-byte gather per arc decides mutuality.  The
// per-vertex bit mask of one-way arcs is kept: region growing reuses it.
// The ascending, de-duplicated CSR in original indices is only materialised on
// export (iftwt_knn_graph with host buffers).
#include <cub/cub.cuh>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

// One warp per vertex, lane j <-> arc j (two rounds when k > 32).
//   ow_mask[u]  bit j set: arc u -> row[j] is one-way (u is not in kNN(row[j]))
//   rdeg[v]    += 1 for every one-way arc into v          (when rdeg != NULL)
//   n_fwd      += number of valid forward arcs            (block-aggregated)
__global__ void __launch_bounds__(256)
knn_mutual_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2, const u64* __restrict__ kth,
                  const float4* __restrict__ spts, u32 n, int k, u64* __restrict__ ow_mask,
                  u32* __restrict__ rdeg, unsigned long long* __restrict__ n_fwd) {
  __shared__ u32 s_cnt[8];
  const u32 u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 valid_cnt = 0;
  if (u < n) {
    const u32 ou = __float_as_uint(spts[u].w);
    u64 mask = 0;
    for (int base = 0; base < k; base += 32) {
      const int j = base + lane;
      bool valid = false, oneway = false;
      u32 v = SID_NONE;
      if (j < k) {
        v = nbr[(size_t)u * k + j];
        valid = v != SID_NONE && v != u;
        if (valid) {
          u64 key = ((u64)__float_as_uint(nd2[(size_t)u * k + j]) << 32) | (u64)ou;
          oneway = key > kth[v];
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, oneway);
      valid_cnt += (u32)__popc(__ballot_sync(0xffffffffu, valid));
      mask |= (u64)m << base;
      if (oneway && rdeg) atomicAdd(&rdeg[v], 1u);
    }
    if (lane == 0) ow_mask[u] = mask;
  }
  if (n_fwd) {
    if (lane == 0) s_cnt[warp] = valid_cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 t = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_cnt[w];
      if (t) atomicAdd(n_fwd, (unsigned long long)t);
    }
  }
}

// scatter the one-way arcs u -> v into v's reverse list.  One THREAD per vertex walks the set bits
// of its one-way mask (1.5 on average, 5 % of the arcs), two arcs at a time: the chain row entry ->
// atomicAdd (with return) -> store is pure latency, and only whole threads' worth of it is worth
// keeping in flight.  (One warp per vertex, lane <-> arc: 1.69 ms at 10 M points, k = 32; one warp per
// four vertices, four chains per lane: 0.88 ms — 95 % of the lanes had no arc.)
__global__ void __launch_bounds__(256)
graph_fill_rev_kernel(const u32* __restrict__ nbr, const u64* __restrict__ ow_mask, u32 n, int k,
                      const u32* __restrict__ roff, u32* __restrict__ cursor, u32* __restrict__ radj) {
  const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  u64 mask = ow_mask[u];
  const u32* row = nbr + (size_t)u * k;
  while (mask) {
    const int j0 = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    int j1 = -1;
    if (mask) {
      j1 = __ffsll((long long)mask) - 1;
      mask &= mask - 1;
    }
    const u32 v0 = row[j0];
    const u32 v1 = j1 >= 0 ? row[j1] : SID_NONE;
    const u32 o0 = roff[v0];
    const u32 p0 = atomicAdd(&cursor[v0], 1u);
    u32 o1 = 0, p1 = 0;
    if (v1 != SID_NONE) {
      o1 = roff[v1];
      p1 = atomicAdd(&cursor[v1], 1u);
    }
    radj[o0 + p0] = u;
    if (v1 != SID_NONE) radj[o1 + p1] = u;
  }
}

int knn_mutual_run(iftwt_ctx* c, bool for_graph) {
  const u32 n = (u32)c->n;
  const int k = c->knn_k;
  ENSURE(c, c->ow_mask, (size_t)n * 8);
  u32* rdeg = nullptr;
  unsigned long long* d_nfwd = nullptr;
  if (for_graph) {
    ENSURE(c, c->g_deg, ((size_t)n + 1) * 4);
    rdeg = c->g_deg.as<u32>();
    CU_TRY(c, cudaMemsetAsync(rdeg, 0, ((size_t)n + 1) * 4, c->stream));
    d_nfwd = (unsigned long long*)(c->scalars.as<int>() + 52);
    CU_TRY(c, cudaMemsetAsync(d_nfwd, 0, 8, c->stream));
  }
  LAUNCH(c, knn_mutual_kernel, div_up((size_t)n * 32, 256), 256, 0, c->nbr.as<u32>(), c->nd2.as<float>(),
         c->kth.as<u64>(), c->spts.as<float4>(), n, k, c->ow_mask.as<u64>(), rdeg, d_nfwd);
  CHECK_LAUNCH(c);
  c->mask_k = k;
  return IFTWT_OK;
}

int graph_build(iftwt_ctx* c) {
  if (c->knn_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "k-NN lists are not computed");
  if (c->graph_kind == 1 && c->graph_k == c->knn_k) return IFTWT_OK;
  StageTimer st(c, IFTWT_STAGE_GRAPH);
  const u32 n = (u32)c->n;
  const int k = c->knn_k;
  if (!(c->mask_k == k && c->rdeg_k == k)) TRY(knn_mutual_run(c, true));  // else: left by the fused pass
  c->rdeg_k = 0;  // g_deg becomes the fill cursor below
  ENSURE(c, c->g_off, ((size_t)n + 1) * 4);
  u32* rdeg = c->g_deg.as<u32>();
  u32* roff = c->g_off.as<u32>();
  {
    size_t tb = 0;
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, rdeg, roff, (int)(n + 1), c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, rdeg, roff, (int)(n + 1), c->stream));
    c->stats.kernel_launches += 2;
  }
  u32* hs = (u32*)c->h_scalars;
  TRY(host_fetch(c, hs + 54, roff + n, 4));
  TRY(host_fetch(c, hs + 52, c->scalars.as<int>() + 52, 8));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const size_t n_rev = hs[54];
  const size_t n_fwd = *(unsigned long long*)(hs + 52);
  ENSURE(c, c->g_adj, (n_rev + 1) * 4);
  CU_TRY(c, cudaMemsetAsync(rdeg, 0, (size_t)n * 4, c->stream));  // now the fill cursor
  if (n_rev) {
    LAUNCH(c, graph_fill_rev_kernel, div_up(n, 256), 256, 0, c->nbr.as<u32>(), c->ow_mask.as<u64>(),
           n, k, roff, rdeg, c->g_adj.as<u32>());
    CHECK_LAUNCH(c);
  }
  c->n_rev = n_rev;
  c->n_arcs = n_fwd + n_rev;
  c->graph_k = k;
  c->graph_kind = 1;
  c->graph_shares_nbr = true;
  c->stats.n_arcs = c->n_arcs;
  c->ift_valid = false;
  return IFTWT_OK;
}

// the k-NN lists are about to be overwritten with another k: keep the graph's rows
int graph_detach_rows(iftwt_ctx* c) {
  if (c->graph_k == 0 || !c->graph_shares_nbr) return IFTWT_OK;
  const size_t bytes = c->n * (size_t)c->graph_k * 4;
  ENSURE(c, c->g_fwd, bytes);
  CU_TRY(c, cudaMemcpyAsync(c->g_fwd.p, c->nbr.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
  c->graph_shares_nbr = false;
  return IFTWT_OK;
}

// ---- export in original index space, rows ascending ----
__global__ void export_degree_kernel(const u32* __restrict__ fwd, int k, const u32* __restrict__ roff,
                                     const u32* __restrict__ inv, u32 n, u32* __restrict__ deg_o) {
  const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i > n) return;
  if (i == n) {
    if (lane == 0) deg_o[i] = 0;
    return;
  }
  const u32 s = inv[i];
  u32 cnt = 0;
  for (int base = 0; base < k; base += 32) {
    int j = base + lane;
    u32 v = j < k ? fwd[(size_t)s * k + j] : SID_NONE;
    cnt += (u32)__popc(__ballot_sync(0xffffffffu, v != SID_NONE && v != s));
  }
  if (lane == 0) deg_o[i] = cnt + (roff[s + 1] - roff[s]);
}
__global__ void export_rows_kernel(const u32* __restrict__ fwd, int k, const u32* __restrict__ roff,
                                   const u32* __restrict__ radj, const u32* __restrict__ inv,
                                   const u32* __restrict__ perm, u32 n, const u32* __restrict__ off_o,
                                   u32* __restrict__ adj_o) {
  const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const u32 s = inv[i];
  u32 o = off_o[i];
  for (int base = 0; base < k; base += 32) {
    int j = base + lane;
    u32 v = j < k ? fwd[(size_t)s * k + j] : SID_NONE;
    bool valid = v != SID_NONE && v != s;
    unsigned m = __ballot_sync(0xffffffffu, valid);
    if (valid) adj_o[o + __popc(m & ((1u << lane) - 1u))] = perm[v];
    o += (u32)__popc(m);
  }
  const u32 b = roff[s], e = roff[s + 1];
  for (u32 t = b + lane; t < e; t += 32) adj_o[o + (t - b)] = perm[radj[t]];
}

int graph_export(iftwt_ctx* c, u32* h_off, u32* h_adj, size_t cap, size_t* n_arcs) {
  const u32 n = (u32)c->n;
  const size_t arcs = c->n_arcs;
  const int k = c->graph_k;
  const u32* fwd = c->graph_shares_nbr ? c->nbr.as<u32>() : c->g_fwd.as<u32>();
  if (n_arcs) *n_arcs = arcs;
  // the segmented sort below counts its items in an int
  if (arcs > 0x7FFFFFFFull)
    return ctx_fail(c, IFTWT_ERR_UNSUPPORTED, "graph export is limited to 2^31 - 1 arcs (this graph has %zu)", arcs);
  ENSURE(c, c->out_a, ((size_t)n + 1) * 4 * 2);
  u32* deg_o = c->out_a.as<u32>();
  u32* off_o = deg_o + (n + 1);
  LAUNCH(c, export_degree_kernel, div_up(((size_t)n + 1) * 32, 256), 256, 0, fwd, k, c->g_off.as<u32>(),
         c->inv.as<u32>(), n, deg_o);
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
    LAUNCH(c, export_rows_kernel, div_up((size_t)n * 32, 256), 256, 0, fwd, k, c->g_off.as<u32>(),
           c->g_adj.as<u32>(), c->inv.as<u32>(), c->perm.as<u32>(), n, off_o, tmp_adj);
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
