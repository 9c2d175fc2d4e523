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
// So each level is ONE launch of an incremental union-find step over the vertices whose
// weight IS c (each vertex and each arc is touched once over the whole run): every newly
// opened vertex t (w[t] == c) unions with its open, not yet conquered neighbours, takes the
// smallest label among its neighbours conquered at earlier levels, and offers (c, label) to
// its component's root with one atomicMin.  Unions and offers of the same level interleave
// freely: the link, the result and the weight of a vertex share one 64-bit word (see node[]
// below), so a hook (atomicCAS) and an offer (atomicMin) on the same root are ordered by the
// hardware and a result is never stranded on a root that has just been hooked.
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
// node[v] — ONE 8-byte word per vertex carries the union-find link, the component's result and
// the vertex weight, so that an arc visit is one gather and every update is one atomic on the
// word it concerns (no fences, no second array to keep in step):
//   hi == 0                  PARENT     lo = parent vertex; v is hooked for good
//   hi == 1                  SEED       cost 0, lo = label (= the seed's original index)
//   hi == c + 2, c < 500     CONQUERED  at level c with label lo.  While level c is running the
//                                       word is "in flight": later offers may still lower lo
//   hi == 502 + w            OPEN/CLOSED root that no offer has reached; w = its own weight
// PARENT < every result word < every unconquered word, so atomicMin(word, offer) is a no-op
// on a hooked vertex (and returns the link to follow), lowers an in-flight result, and
// conquers an unconquered root.  A hook is atomicCAS(root word -> PARENT): it succeeds only on
// the exact word it read, so whatever result had landed on the hooked root is known to the
// hooking thread, which forwards it to the new root.  Words of components conquered at an
// earlier level never change again; any thread may copy such a root word over a member's
// PARENT word ("lazy finalize") to cut later chases short.
#define NODE_HI(x) ((u32)((x) >> 32))
#define NODE_LO(x) ((u32)(x))
#define NODE_UNCONQ 502u

namespace {

// follows PARENT links from v; returns the first non-PARENT word, v = its vertex (path halving)
__device__ __forceinline__ u64 node_find(u64* __restrict__ node, u32& v, u64 x) {
  while (NODE_HI(x) == 0) {
    const u32 p = NODE_LO(x);
    const u64 y = node[p];
    if (NODE_HI(y) == 0) node[v] = y;
    v = p;
    x = y;
  }
  return x;
}
// an offer travels up the links until it lands on a root word
__device__ __forceinline__ void node_offer(u64* __restrict__ node, u32 v, u64 val) {
  while (true) {
    const u64 old = atomicMin((unsigned long long*)&node[v], (unsigned long long)val);
    if (NODE_HI(old) != 0) return;
    v = NODE_LO(old);
  }
}
// weights (original order) -> u16 per sorted vertex; flags non-integer / negative input
__global__ void ift_init_kernel(const float* __restrict__ w_orig, const float4* __restrict__ pts_in,
                                const u32* __restrict__ perm, u32 n, unsigned short* __restrict__ w16,
                                u64* __restrict__ node, u32* __restrict__ ids, float* __restrict__ wf,
                                u32* __restrict__ bad) {
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
  node[v] = (u64)(NODE_UNCONQ + q) << 32;
  ids[v] = v;
}
__global__ void ift_seed_kernel(const u32* __restrict__ seeds_orig, u32 ns, const u32* __restrict__ inv, u32 n,
                                unsigned short* __restrict__ w16, u64* __restrict__ node, u32* __restrict__ bad) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  u32 o = seeds_orig[i];
  if (o >= n) {
    atomicAdd(bad + 1, 1u);
    return;
  }
  u32 v = inv[o];
  w16[v] = 0;                      // listed with level 0; its own value never matters (ift.hpp:138)
  node[v] = (1ull << 32) | (u64)o; // cost 0, label = itself, conquered before level 0 starts
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

// One visit of the arc t -> u at level `level`.  (rt, xt) is this lane's belief of t's root and
// of the word it last saw there; xu is node[u] as loaded by the caller.  Every belief may be
// stale: decisions are taken either on words that can no longer change (results of earlier
// levels, weights) or through an atomic that fails — and says where to look — when the belief
// was wrong.
__device__ __forceinline__ void ift_visit(u64* __restrict__ node, u32 level, u32& rt, u64& xt, u32 u, u64 xu,
                                          u32& lmin) {
  u32 ru = u;
  if (NODE_HI(xu) == 0) xu = node_find(node, ru, xu);
  const u32 h = NODE_HI(xu);
  if (h < level + 2u) {  // conquered at an earlier level: an entry offer at cost `level`
    lmin = min(lmin, NODE_LO(xu));
    if (ru != u) node[u] = xu;  // lazy finalize
    return;
  }
  if (h >= NODE_UNCONQ && h - NODE_UNCONQ > level) return;  // closed at this level
  // open and not conquered before this level: same component as t.  Hook the larger root
  // under the smaller; a result that already landed on the hooked root moves to the new root.
  while (rt != ru) {
    if (rt < ru) {
      const u64 old = atomicCAS((unsigned long long*)&node[ru], (unsigned long long)xu, (unsigned long long)rt);
      if (old == xu) {
        if (NODE_HI(xu) < NODE_UNCONQ) node_offer(node, rt, xu);
        break;
      }
      xu = node_find(node, ru, old);
    } else {
      const u64 old = atomicCAS((unsigned long long*)&node[rt], (unsigned long long)xt, (unsigned long long)ru);
      if (old == xt) {
        if (NODE_HI(xt) < NODE_UNCONQ) node_offer(node, ru, xt);
        rt = ru;
        xt = xu;
        break;
      }
      xt = node_find(node, rt, old);
    }
  }
}
// Blocks [0, open_blocks) open the vertices of this level's segment, G lanes per vertex; the
// remaining blocks finalize the previous level's segment (one thread per vertex) — independent
// work that would otherwise need a launch of its own.
// Why G varies: a vertex visit is a chain of dependent memory accesses, so a level's time is
// (vertices in flight)^-1 x chain latency, not bandwidth.  With a whole warp per vertex only
// 64 x 148 vertices are in flight; big levels use fewer lanes per vertex (down to one thread
// per vertex) to put up to 32x more chains in flight, small levels keep a warp per vertex so
// that each chain is as short as possible.  Occupancy is what buys memory
// parallelism here: at 46 registers (5 blocks / SM) the run took 8.2 ms, at 32 registers
// 6.3 ms; fetching 2 or 4 neighbour words ahead of the visits did not help (6.4 / 7.8 ms).
// Measured and rejected as well: streaming (evict-first) row loads (6.28 vs 6.23 ms) and a
// persisting-L2 window over node[] — the carve-out it needs slowed every other stage (grid
// 1.4 -> 2.5 ms) and the IFT itself (6.2 -> 7.9 ms); a w16 pre-gather that skips the node word of
// closed neighbours (one more link in the chain: 6.2 -> 6.5 ms); runs of consecutive small levels
// in one cooperative launch with a grid barrier between levels and ld.cg node reads (6.34 vs
// 6.23 ms: a barrier over 1184 resident blocks is no cheaper than a launch boundary); a pre-pass
// that marks, per vertex, the forward neighbours that are not heavier, so that the level kernels
// skip the arcs to still-closed neighbours (half of the forward visits) — the pre-pass and the
// extra mask load cost more than the skipped visits saved (6.2 -> 7.1 ms).  Round 2 repeated it with the mask made
// for free — by the fused graph / region-growing pass, which has both ends of every arc in hand (one more 2-byte
// gather per arc there), seeds patched in afterwards — and the level kernels skipping the row entry and the node
// word of every masked arc: IFT 5.16 -> 5.10 ms, the graph pass 2.29 -> 2.63 ms.  Half the gathers gone and the
// levels barely faster: they are bound by the length of the dependent chain per vertex, not by L2 sector rate.
#define IFT_PEND 6
#define IFT_AHEAD 2  // 3 / 4 rounds ahead spill at 32 registers: 5.84 / 5.95 ms against 5.70; 4 ahead at 40 registers: 6.2 ms
template <int G>
__global__ void __launch_bounds__(256, 8)  // 32 registers: 2048 chains in flight per SM
ift_level_kernel(const u32* __restrict__ byw, u32 seg_begin, u32 seg_len, u32 level, u32 open_blocks,
                 u32 prev_begin, u32 prev_len, const u32* __restrict__ fwd, int k,
                 const u32* __restrict__ roff, const u32* __restrict__ radj, u64* __restrict__ node) {
  __shared__ u32 s_pend[G == 32 ? 1 : 256 * IFT_PEND];
  // Programmatic dependent launch: the next level may become resident as soon as every block of
  // this one has started.  What a block reads before pdl_wait() — its vertex, the bounds of its
  // row, the first neighbour ids — was written before the level loop began; node[] is touched
  // only after the wait, when the previous level is complete.  This takes the launch latency and
  // three of the dependent loads out of every level's chain (148 of the 227 levels of cfg3 hold
  // fewer than 38 k vertices and are nothing but that chain).
  pdl_launch_dependents();
  if (blockIdx.x >= open_blocks) {
    const u32 i = (blockIdx.x - open_blocks) * blockDim.x + threadIdx.x;
    if (i >= prev_len) return;
    const u32 t = byw[prev_begin + i];
    pdl_wait();
    u64 x = node[t];
    if (NODE_HI(x) != 0) return;
    u32 r = t;
    x = node_find(node, r, x);
    if (NODE_HI(x) < level + 2u) node[t] = x;
    return;
  }
  const u32 gv = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int gl = threadIdx.x % G;
  const bool active = gv < seg_len;
  u32 t = 0, rt = 0, b = 0, na = 0;
  u64 xt = 0;
  bool work = false;
  const u32* row = fwd;
  u32 p0 = 0xFFFFFFFFu, p1 = 0xFFFFFFFFu;  // G == 32: the first two rounds of neighbour ids
  if (active) {
    t = byw[seg_begin + gv];
    b = roff[t];
    na = (u32)k + (roff[t + 1] - b);
    row = fwd + (size_t)t * k;
    // the row's sectors start moving towards L2 while the previous level drains (5.13 -> 5.06 ms; doing the same
    // for node[t] and the first neighbours' words of a warp-per-vertex level changed nothing)
    if (G < 32) {
      for (u32 i = (u32)gl * 8u; i < (u32)k; i += 8u * G)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + i));
    }
    if (G == 32) {
      const u32 i = gl, i1 = gl + G;
      if (i < na) p0 = i < (u32)k ? row[i] : radj[b + i - k];
      if (i1 < na) p1 = i1 < (u32)k ? row[i1] : radj[b + i1 - k];
    }
  }
  pdl_wait();
  u64 q0 = 0, q1 = 0;  // G == 32: the words of the first two neighbours, in flight with node[t]
  if (active) {
    if (G == 32) {
      if (p0 == t) p0 = 0xFFFFFFFFu;
      if (p1 == t) p1 = 0xFFFFFFFFu;
      if (p0 != 0xFFFFFFFFu) q0 = node[p0];
      if (p1 != 0xFFFFFFFFu) q1 = node[p1];
    }
    rt = t;
    xt = node_find(node, rt, node[t]);       // t may already hang under a same-level neighbour
    work = NODE_HI(xt) >= level + 2u;        // a seed is conquered from the start
  }
  u32 lmin = LBL_NONE;
  if (work) {
    if (G == 32) {
      // a warp per vertex runs at most a few rounds: fetch two rounds of neighbour words before
      // working on either (the visits hold atomics the compiler cannot move loads across)
      for (u32 i = gl; i < na; i += 2 * G) {
        const u32 i1 = i + G;
        u32 u0 = p0, u1 = p1;
        u64 x0 = q0, x1 = q1;
        if (i != (u32)gl) {
          u0 = i < (u32)k ? row[i] : radj[b + i - k];
          u1 = i1 < na ? (i1 < (u32)k ? row[i1] : radj[b + i1 - k]) : 0xFFFFFFFFu;
          if (u0 == t) u0 = 0xFFFFFFFFu;
          if (u1 == t) u1 = 0xFFFFFFFFu;
          x0 = u0 != 0xFFFFFFFFu ? node[u0] : 0ull;
          x1 = u1 != 0xFFFFFFFFu ? node[u1] : 0ull;
        }
        if (u0 != 0xFFFFFFFFu) ift_visit(node, level, rt, xt, u0, x0, lmin);
        if (u1 != 0xFFFFFFFFu) ift_visit(node, level, rt, xt, u1, x1, lmin);
      }
    } else {
      // Several vertices share a warp here, and a visit that needs the union-find (a chase, a CAS)
      // would stall all of them: with ~3 same-level neighbours per vertex some lane of the warp
      // needs one in nearly every round.  So the pass over the arcs only gathers — conquered
      // neighbours give their label, closed ones are dropped, IFT_AHEAD rounds of loads in flight — and
      // the few neighbours that need the union-find are parked and worked off afterwards, when
      // every busy lane is doing that same kind of work.
      u32* pend = s_pend + threadIdx.x;  // slot q of this thread at pend[256 q]
      int np = 0;
      for (u32 i0 = gl; i0 < na; i0 += IFT_AHEAD * G) {
        u32 uu[IFT_AHEAD];
        u64 xx[IFT_AHEAD];
#pragma unroll
        for (int q = 0; q < IFT_AHEAD; ++q) {
          const u32 i = i0 + q * G;
          uu[q] = i < na ? (i < (u32)k ? row[i] : radj[b + i - k]) : 0xFFFFFFFFu;
          if (uu[q] == t) uu[q] = 0xFFFFFFFFu;
        }
#pragma unroll
        for (int q = 0; q < IFT_AHEAD; ++q) xx[q] = uu[q] != 0xFFFFFFFFu ? node[uu[q]] : ~0ull;
#pragma unroll
        for (int q = 0; q < IFT_AHEAD; ++q) {
          const u32 u = uu[q];
          const u64 x = xx[q];
          const u32 h = NODE_HI(x);
          if (h != 0 && h < level + 2u) {
            lmin = min(lmin, NODE_LO(x));  // conquered at an earlier level
          } else if (h >= NODE_UNCONQ && h - NODE_UNCONQ > level) {
            // closed at this level (or no neighbour: the filler word)
          } else if (np < IFT_PEND) {
            pend[256 * np++] = u;
          } else {
            ift_visit(node, level, rt, xt, u, x, lmin);
          }
        }
      }
      for (int q = 0; q < np; ++q) {
        const u32 u = pend[256 * q];
        ift_visit(node, level, rt, xt, u, node[u], lmin);
      }
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
  if (work && gl == 0 && lmin != LBL_NONE) node_offer(node, rt, ((u64)(level + 2u) << 32) | (u64)lmin);
}
// every vertex reads its component's word; unconquered vertices keep (500, self).  One thread per ORIGINAL
// index: cost and label leave as coalesced stores and the component word is a gather (the per-sorted-vertex form
// scattered two 4-byte stores per point — 640 MB of read-modify-write sectors, 0.25 ms at 10 M points).
__global__ void ift_resolve_kernel(u64* __restrict__ node, const u32* __restrict__ inv, u32 n,
                                   u32* __restrict__ cost_o, u32* __restrict__ label_o) {
  u32 o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const u32 v = inv[o];
  u32 r = v;
  const u64 x = node_find(node, r, node[v]);
  const u32 h = NODE_HI(x);
  u32 cst = IFTWT_MAX_COST, lbl = o;
  if (h < NODE_UNCONQ) {
    cst = h < 2u ? 0u : h - 2u;
    lbl = NODE_LO(x);
  }
  cost_o[o] = cst;
  label_o[o] = lbl;
}
// (cost << 32 | label) per SORTED vertex, for the region hierarchy: built when it is first asked for
__global__ void ift_res_kernel(u64* __restrict__ node, const u32* __restrict__ perm, u32 n, u64* __restrict__ res) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  u32 r = v;
  const u64 x = node_find(node, r, node[v]);
  const u32 h = NODE_HI(x);
  u32 cst = IFTWT_MAX_COST, lbl = perm[v];
  if (h < NODE_UNCONQ) {
    cst = h < 2u ? 0u : h - 2u;
    lbl = NODE_LO(x);
  }
  res[v] = ((u64)cst << 32) | (u64)lbl;
}

}  // namespace

int ift_run(iftwt_ctx* c, const float* d_w_orig, const u32* d_seeds_orig, size_t ns) {
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  // the buffers of the previous run are overwritten from here on: a failed call must not leave them "valid"
  c->ift_valid = false;
  c->rag_valid = false;
  StageTimer st(c, IFTWT_STAGE_IFT);
  const u32 n = (u32)c->n;
  ENSURE(c, c->ift_w, (size_t)n * 2 * 2 + 16);
  ENSURE(c, c->ift_byw, (size_t)n * 4 * 2);
  ENSURE(c, c->ift_wf, (size_t)n * 4);
  ENSURE(c, c->ift_node, (size_t)n * 8);
  u64* node = c->ift_node.as<u64>();
  ENSURE(c, c->out_a, (size_t)n * 4 * 2);
  unsigned short* w16 = c->ift_w.as<unsigned short>();
  unsigned short* w16_sorted = w16 + ((n + 7) & ~7u);
  u32* ids = c->ift_byw.as<u32>();
  u32* byw = ids + n;
  u32* sc = (u32*)c->scalars.as<int>();
  u32* hs = (u32*)c->h_scalars;
  u32* d_bad = sc + 64;
  u32* d_seg = sc + 128;  // 513 entries
  CU_TRY(c, cudaMemsetAsync(d_bad, 0, 8, c->stream));
  LAUNCH(c, ift_init_kernel, div_up(n, 256), 256, 0, d_w_orig, c->pts_in.as<float4>(), c->perm.as<u32>(), n, w16,
         node, ids, c->ift_wf.as<float>(), d_bad);
  if (ns) LAUNCH(c, ift_seed_kernel, div_up(ns, 256), 256, 0, d_seeds_orig, (u32)ns, c->inv.as<u32>(), n, w16,
                 node, d_bad);
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
  TRY(host_fetch(c, hs + 128, d_seg, 513 * 4));
  TRY(host_fetch(c, hs + 64, d_bad, 8));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (hs[64])
    return ctx_fail(c, IFTWT_ERR_WEIGHTS, "%u IFT weights are negative, NaN or not integer valued", hs[64]);
  if (hs[65]) return ctx_fail(c, IFTWT_ERR_INVALID, "%u seed indices are out of range", hs[65]);
  const u32* seg = hs + 128;
  const u32* fwd = c->graph_shares_nbr ? c->nbr.as<u32>() : c->g_fwd.as<u32>();
  u64 levels = 0;
  u32 prev_b = 0, prev_len = 0;
  const char* wenv = getenv("IFTWT_IFT_WAVES");
  const double waves = wenv ? atof(wenv) : 3.0;  // swept on B200: 1 -> 6.1 ms, 2 -> 5.8, 3 -> 5.65, 4 -> 5.8, 6 -> 6.15, 8 -> 6.4
  for (u32 lv = 0; lv < IFTWT_MAX_COST; ++lv) {
    u32 b = seg[lv], len = seg[lv + 1] - seg[lv];
    if (len == 0) continue;
    ++levels;
    // lanes per vertex: keep about two waves of threads in flight
    int G = 32;
    while (G > 1 && (double)len * G > waves * 2048.0 * (double)c->sm_count) G >>= 1;
    const u32 open_blocks = div_up((size_t)len * G, 256);
    auto kern = G == 32 ? ift_level_kernel<32> : G == 16 ? ift_level_kernel<16>
              : G == 8 ? ift_level_kernel<8> : G == 4 ? ift_level_kernel<4>
              : G == 2 ? ift_level_kernel<2> : ift_level_kernel<1>;
    LAUNCH_PDL(c, kern, open_blocks + div_up(prev_len, 256), 256, 0, (const u32*)byw, b, len, lv, open_blocks, prev_b,
               prev_len, fwd, c->graph_k, (const u32*)c->g_off.as<u32>(), (const u32*)c->g_adj.as<u32>(), node);
    prev_b = b;
    prev_len = len;
  }
  CHECK_LAUNCH(c);
  u32* cost_o = c->out_a.as<u32>();
  u32* label_o = cost_o + n;
  LAUNCH(c, ift_resolve_kernel, div_up(n, 256), 256, 0, node, c->inv.as<u32>(), n, cost_o, label_o);
  CHECK_LAUNCH(c);
  c->ift_res_valid = false;
  c->stats.ift_levels = levels;
  c->ift_valid = true;
  c->rag_valid = false;
  return IFTWT_OK;
}

// the per-sorted-vertex result words the region hierarchy reads (hierarchy.cu); node[] keeps the run's state
int ift_res_ensure(iftwt_ctx* c) {
  if (!c->ift_valid) return ctx_fail(c, IFTWT_ERR_STATE, "no IFT result on the device");
  if (c->ift_res_valid) return IFTWT_OK;
  const u32 n = (u32)c->n;
  ENSURE(c, c->ift_res, (size_t)n * 8);
  LAUNCH(c, ift_res_kernel, div_up(n, 256), 256, 0, c->ift_node.as<u64>(), c->perm.as<u32>(), n, c->ift_res.as<u64>());
  CHECK_LAUNCH(c);
  c->ift_res_valid = true;
  return IFTWT_OK;
}
