// ift.cu — K7: Image Foresting Transform watershed with the f_max path cost.
//
// Replaces IFT_PCD::copy_graph / add_seeds / compute_watershed (ift.hpp:97-153,
// 170-250) and getRoots (ift.hpp:48-50).  Semantics kept from the reference:
//   cost(t) = min over seeds and paths of max(w[v]) over the vertices after the
//   seed (the arc weight is the DESTINATION vertex value, ift.hpp:198); costs are
//   unsigned ints below MAX_COST = 500 (ift.hpp:1, 203); a seed's cost is 0 whatever
//   its own value (ift.hpp:138); vertices that are never conquered keep cost 500
//   and label = self (ift.hpp:101, 179).
//
// Formulation.  The reference drains a priority queue; a literal GPU restatement
// (bucketed frontier sweeps) needs one grid-wide step per hop of every plateau —
// thousands of steps on a 10 M point surface.  With f_max the Dial buckets have a
// much stronger structure: processing cost levels c = 0,1,2,... in order, the
// vertices conquered at level c are exactly the connected components of the
// "open" set {unconquered, w <= c} that touch an already conquered vertex, every
// vertex of such a component gets cost c, and under the min-seed tie policy
// (DESIGN.md policy C: a vertex keeps the smallest (cost, seed id) offered by a
// final neighbour) the whole component takes the smallest seed label on its rim.
// So each level is one incremental union-find step over the vertices whose weight
// IS c (each vertex and each arc is touched once over the whole run):
//   A. every newly opened vertex t (w[t] == c) unions with its open, unconquered
//      neighbours and remembers the smallest label among its conquered neighbours;
//   B. that label is pushed to the component root with one atomicMin on the packed
//      (cost << 32 | label) word — the root's word is the whole component's result.
// Lakes (open components no conquered vertex touches yet) simply stay in the
// union-find until a later level reaches them.  Levels stay strictly ordered, which
// is what makes the labels schedule independent (SURVEY finding 6); inside a level
// nothing depends on thread order.
#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#define W_CLOSED 511u  // weight >= MAX_COST: the vertex can never be conquered
#define LBL_NONE 0xFFFFFFFFu
// info[v] = FINAL | weight << 32 | label: everything a neighbour needs from v in ONE 8-byte
// gather.  FINAL is set (with the component's label) once v's component is conquered — by the
// finalize blocks that ride along with the next level's launch, or lazily by the first
// visitor that resolves v through the union-find.  A conquered component never changes, so
// every writer stores the same word.
#define INFO_FINAL 0x8000000000000000ull
#define INFO_W(e) ((u32)((e) >> 32) & 0x1FFu)

namespace {

__device__ __forceinline__ u32 uf_find(u32* parent, u32 v) {
  u32 p = parent[v];
  while (p != v) {
    u32 gp = parent[p];
    if (gp != p) parent[v] = gp;
    v = p;
    p = gp;
  }
  return v;
}
// weights (original order) -> u16 per sorted vertex; flags non-integer / negative input
__global__ void ift_init_kernel(const float* __restrict__ w_orig, const float4* __restrict__ pts_in,
                                const u32* __restrict__ perm, u32 n, unsigned short* __restrict__ w16,
                                u64* __restrict__ state, u32* __restrict__ parent, u32* __restrict__ ids,
                                float* __restrict__ wf, u64* __restrict__ info, u32* __restrict__ bad) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  u32 o = perm[v];
  float w = w_orig ? w_orig[o] : pts_in[o].w;
  wf[v] = w;
  unsigned short q;
  if (!(w >= 0.f) || w != floorf(w)) {
    atomicAdd(bad, 1u);
    q = W_CLOSED;
  } else {
    q = w >= (float)IFTWT_MAX_COST ? (unsigned short)W_CLOSED : (unsigned short)w;
  }
  w16[v] = q;
  info[v] = ((u64)q << 32) | (u64)LBL_NONE;
  state[v] = ((u64)IFTWT_MAX_COST << 32) | (u64)o;
  parent[v] = v;
  ids[v] = v;
}
__global__ void ift_seed_kernel(const u32* __restrict__ seeds_orig, u32 ns, const u32* __restrict__ inv, u32 n,
                                unsigned short* __restrict__ w16, u64* __restrict__ state,
                                u64* __restrict__ info, u32* __restrict__ bad) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  u32 o = seeds_orig[i];
  if (o >= n) {
    atomicAdd(bad + 1, 1u);
    return;
  }
  u32 v = inv[o];
  w16[v] = 0;            // open from level 0 on; its own value never matters (ift.hpp:138)
  state[v] = (u64)o;     // cost 0, label = itself
  info[v] = INFO_FINAL | (u64)o;
}
// seg_off[c] = first position in the weight-sorted vertex list with weight >= c, c = 0..512
__global__ void ift_segments_kernel(const unsigned short* __restrict__ w_sorted, u32 n, u32* __restrict__ seg_off) {
  u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > 512) return;
  u32 lo = 0, hi = n;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if ((u32)w_sorted[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  seg_off[c] = lo;
}

// A: one warp per newly opened vertex.  adjacency(t) = its k-NN row + its reverse arcs.
struct OpenCtx {
  u64* info;
  const u64* state;
  u32* parent;
  u32 level;
};
// rt = this lane's belief of t's representative (t itself until a hook says otherwise); a stale
// belief only costs a failed CAS, which returns the way up.
__device__ __forceinline__ void ift_visit(const OpenCtx& x, u32& rt, u32 u, u32& lmin) {
  const u64 e = x.info[u];
  if (e & INFO_FINAL) {  // conquered neighbour: an entry offer at cost `level`
    lmin = min(lmin, (u32)e);
    return;
  }
  if (INFO_W(e) > x.level) return;  // closed at this level
  u32 ru = uf_find(x.parent, u);
  const u64 su = x.state[ru];
  if ((u32)(su >> 32) < IFTWT_MAX_COST) {
    lmin = min(lmin, (u32)su);
    x.info[u] = INFO_FINAL | (e & 0x000001FF00000000ull) | (u64)(u32)su;  // lazy finalize
    return;
  }
  while (rt != ru) {  // hook the larger representative under the smaller
    if (rt < ru) {
      const u32 ret = atomicCAS(&x.parent[ru], ru, rt);
      if (ret == ru) break;
      ru = ret;
    } else {
      const u32 ret = atomicCAS(&x.parent[rt], rt, ru);
      if (ret == rt) {
        rt = ru;
        break;
      }
      rt = ret;
    }
  }
}
// Blocks [0, open_blocks) open the vertices of this level's segment, G lanes per vertex; the
// remaining blocks finalize the previous level's segment (one thread per vertex) — independent
// work that would otherwise need a launch of its own.
// Why G varies: a vertex visit is a chain of ~7 dependent memory accesses, so a level's time is
// (vertices in flight)^-1 x chain latency, not bandwidth.  With a whole warp per vertex only
// 64 x 148 vertices are in flight; big levels use fewer lanes per vertex (down to one thread
// per vertex) to put up to 32x more chains in flight, small levels keep a warp per vertex so
// that each chain is as short as possible.
template <int G>
__global__ void __launch_bounds__(256)
ift_level_open_kernel(const u32* __restrict__ byw, u32 seg_begin, u32 seg_len, u32 level, u32 open_blocks,
                      u32 prev_begin, u32 prev_len, const u32* __restrict__ fwd, int k,
                      const u32* __restrict__ roff, const u32* __restrict__ radj, u64* __restrict__ info,
                      const u64* __restrict__ state, u32* __restrict__ parent, u32* __restrict__ local_min) {
  if (blockIdx.x >= open_blocks) {
    const u32 i = (blockIdx.x - open_blocks) * blockDim.x + threadIdx.x;
    if (i >= prev_len) return;
    const u32 t = byw[prev_begin + i];
    const u64 e = info[t];
    if (e & INFO_FINAL) return;
    const u64 s = state[uf_find(parent, t)];
    if ((u32)(s >> 32) < IFTWT_MAX_COST) info[t] = INFO_FINAL | (e & 0x000001FF00000000ull) | (u64)(u32)s;
    return;
  }
  const u32 gv = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int gl = threadIdx.x % G;
  const bool active = gv < seg_len;
  u32 t = 0;
  bool work = false;
  if (active) {
    t = byw[seg_begin + gv];
    work = !(info[t] & INFO_FINAL);  // a seed is conquered from the start
  }
  u32 lmin = LBL_NONE;
  if (work) {
    OpenCtx x{info, state, parent, level};
    u32 rt = t;
    for (int j = gl; j < k; j += G) {
      u32 u = fwd[(size_t)t * k + j];
      if (u != 0xFFFFFFFFu && u != t) ift_visit(x, rt, u, lmin);
    }
    const u32 b = roff[t], e = roff[t + 1];
    for (u32 a = b + gl; a < e; a += G) ift_visit(x, rt, radj[a], lmin);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
  if (active && gl == 0) local_min[t] = lmin;
}
// B: push the entry labels to the component roots
__global__ void ift_level_settle_kernel(const u32* __restrict__ byw, u32 seg_begin, u32 seg_len, u32 level,
                                        u32* __restrict__ parent, const u32* __restrict__ local_min,
                                        u64* __restrict__ state) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= seg_len) return;
  u32 t = byw[seg_begin + i];
  u32 lm = local_min[t];
  if (lm == LBL_NONE) return;
  u32 r = uf_find(parent, t);
  atomicMin((unsigned long long*)&state[r], (unsigned long long)(((u64)level << 32) | (u64)lm));
}
// every vertex reads its component's word; unconquered vertices keep (500, self)
__global__ void ift_resolve_kernel(u32* __restrict__ parent, u64* __restrict__ state, const u32* __restrict__ perm,
                                   u32 n, u32* __restrict__ cost_o, u32* __restrict__ label_o,
                                   u64* __restrict__ res) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  u32 r = uf_find(parent, v);
  u64 s = state[r];
  u32 o = perm[v];
  u32 cst = (u32)(s >> 32);
  if (cst >= IFTWT_MAX_COST) s = ((u64)IFTWT_MAX_COST << 32) | (u64)o;
  cost_o[o] = (u32)(s >> 32);
  label_o[o] = (u32)s;
  res[v] = s;
}

}  // namespace

int ift_run(iftwt_ctx* c, const float* d_w_orig, const u32* d_seeds_orig, size_t ns) {
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  StageTimer st(c, IFTWT_STAGE_IFT);
  const u32 n = (u32)c->n;
  ENSURE(c, c->ift_state, (size_t)n * 8);
  ENSURE(c, c->ift_parent, (size_t)n * 4);
  ENSURE(c, c->ift_w, (size_t)n * 2 * 2 + 16);
  ENSURE(c, c->ift_byw, (size_t)n * 4 * 2);
  ENSURE(c, c->ift_local, (size_t)n * 4);
  ENSURE(c, c->ift_res, (size_t)n * 8);
  ENSURE(c, c->ift_wf, (size_t)n * 4);
  ENSURE(c, c->ift_info, (size_t)n * 8);
  u64* info = c->ift_info.as<u64>();
  ENSURE(c, c->out_a, (size_t)n * 4 * 2);
  u64* state = c->ift_state.as<u64>();
  u32* parent = c->ift_parent.as<u32>();
  unsigned short* w16 = c->ift_w.as<unsigned short>();
  unsigned short* w16_sorted = w16 + ((n + 7) & ~7u);
  u32* ids = c->ift_byw.as<u32>();
  u32* byw = ids + n;
  u32* local_min = c->ift_local.as<u32>();
  u32* sc = (u32*)c->scalars.as<int>();
  u32* hs = (u32*)c->h_scalars;
  u32* d_bad = sc + 64;
  u32* d_seg = sc + 128;  // 513 entries
  CU_TRY(c, cudaMemsetAsync(d_bad, 0, 8, c->stream));
  LAUNCH(c, ift_init_kernel, div_up(n, 256), 256, 0, d_w_orig, c->pts_in.as<float4>(), c->perm.as<u32>(), n, w16,
         state, parent, ids, c->ift_wf.as<float>(), info, d_bad);
  if (ns) LAUNCH(c, ift_seed_kernel, div_up(ns, 256), 256, 0, d_seeds_orig, (u32)ns, c->inv.as<u32>(), n, w16,
                 state, info, d_bad);
  CHECK_LAUNCH(c);
  {  // vertices grouped by weight (stable: ascending vertex id inside a level)
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, w16, w16_sorted, ids, byw, (int)n, 0, 9, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, w16, w16_sorted, ids, byw, (int)n, 0, 9, c->stream));
    c->stats.kernel_launches += 4;
  }
  LAUNCH(c, ift_segments_kernel, 3, 256, 0, w16_sorted, n, d_seg);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaMemcpyAsync(hs + 128, d_seg, 513 * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaMemcpyAsync(hs + 64, d_bad, 8, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (hs[64])
    return ctx_fail(c, IFTWT_ERR_WEIGHTS, "%u IFT weights are negative, NaN or not integer valued", hs[64]);
  if (hs[65]) return ctx_fail(c, IFTWT_ERR_INVALID, "%u seed indices are out of range", hs[65]);
  const u32* seg = hs + 128;
  const u32* fwd = c->graph_shares_nbr ? c->nbr.as<u32>() : c->g_fwd.as<u32>();
  u64 levels = 0;
  u32 prev_b = 0, prev_len = 0;
  const char* wenv = getenv("IFTWT_IFT_WAVES");
  const double waves = wenv ? atof(wenv) : 4.0;  // swept on B200: 0.5 -> 10.4 ms, 2 -> 7.9, 4 -> 7.6, 8 -> 8.0, warp per vertex -> 9.5
  for (u32 lv = 0; lv < IFTWT_MAX_COST; ++lv) {
    u32 b = seg[lv], len = seg[lv + 1] - seg[lv];
    if (len == 0) continue;
    ++levels;
    // lanes per vertex: keep about two waves of threads in flight
    int G = 32;
    while (G > 1 && (double)len * G > waves * 2048.0 * (double)c->sm_count) G >>= 1;
    const u32 open_blocks = div_up((size_t)len * G, 256);
    auto kern = G == 32 ? ift_level_open_kernel<32> : G == 16 ? ift_level_open_kernel<16>
              : G == 8 ? ift_level_open_kernel<8> : G == 4 ? ift_level_open_kernel<4>
              : G == 2 ? ift_level_open_kernel<2> : ift_level_open_kernel<1>;
    LAUNCH(c, kern, open_blocks + div_up(prev_len, 256), 256, 0, byw, b, len, lv, open_blocks, prev_b,
           prev_len, fwd, c->graph_k, c->g_off.as<u32>(), c->g_adj.as<u32>(), info, state, parent, local_min);
    LAUNCH(c, ift_level_settle_kernel, div_up(len, 256), 256, 0, byw, b, len, lv, parent, local_min, state);
    prev_b = b;
    prev_len = len;
  }
  CHECK_LAUNCH(c);
  u32* cost_o = c->out_a.as<u32>();
  u32* label_o = cost_o + n;
  LAUNCH(c, ift_resolve_kernel, div_up(n, 256), 256, 0, parent, state, c->perm.as<u32>(), n, cost_o, label_o,
         c->ift_res.as<u64>());
  CHECK_LAUNCH(c);
  c->stats.ift_levels = levels;
  c->ift_valid = true;
  c->rag_valid = false;
  return IFTWT_OK;
}
