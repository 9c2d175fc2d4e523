// knn.cu — K1 / K1': exact k-NN on the Morton-sorted uniform grid.
//
// Replaces pcl::search::KdTree::nearestKSearch as called at graph.hpp:245 (k=2),
// main.cpp:108 (k=30, inside NormalEstimationOMP), main.cpp:115 (k=50, inside
// RegionGrowing), graph.hpp:260-266 / flat_point.hpp:74 / ift.hpp:135 (k=1 with an
// arbitrary query point).  Results are the exact k nearest neighbours in the
// canonical (d2, original index) order; d2 is FLANN's L2_Simple<float>
// accumulation (common.cuh: dist2_exact), bit for bit the oracle's.
//
// Fast path (knn_cell_kernel): one warp per occupied cell.  The warp looks the 27
// cells of the 3x3x3 block up in the cell hash, stages their points in shared
// memory with coalesced float4 loads (32 slots per round, whatever cell they come
// from), and then answers the cell's own queries one by one: 32 lanes evaluate the
// staged candidates, a count-based threshold search (per-lane counts + one warp
// reduction; candidates ~ tau in 2-D) narrows them to at most CS survivors, which
// are compacted lane-major into shared memory and ordered — k <= 16 and k > 32: ranked
// against each other, the rank IS the output slot; 17 <= k <= 32: a 64-key bitonic
// network over packed (quantised d2, list position) keys, at most 48 of them real.  A
// query is finished when its k-th distance is inside the block (per-point bound from
// grid.cu); otherwise it is appended to the fallback list.
//
// General path (knn_general_kernel): one warp per query, for fallback queries and
// for external query points (a handful of k = 1 queries: one block per query,
// knn_nearest_block_kernel).  Stage j scans the 3x3x3 block of super-cells of side
// cell*2^j (contiguous Morton-prefix ranges of the sorted cloud, found by binary
// search in the cell table), keeps the k best in a lane-distributed sorted list,
// and stops when the k-th distance is inside the scanned block.
#include <stdlib.h>

#include "common.cuh"

// warps (= cells) per block.  A block holds its shared memory until its slowest warp is done; k <= 16:
// 4 warps 4.26 ms against 4.51 with 8; k = 32: 5.18 against 5.09
#define KNN_WARPS_OF(E) ((E) == 1 ? 4 : 8)
// staged candidates per warp.  (E == 2 with 192 for one more resident block and the cells above it left to
// the second launch: 5.45 ms against 5.22.)
#define KNN_CAP_OF(E) ((E) <= 2 ? 256 : 512)
// Two compile-time round counts per E: the common case and the full buffer.  (Eight variants
// made 9,144 SASS instructions and the kernel stalled on instruction fetch: ncu no_instruction
// 4.2 per issue, profiles/r01_knn_cell_kernel_v3_ncu_full.txt.  A third count of 6 for E == 2 — a third of
// cfg3's cells need exactly six rounds — grew the kernel from 2,700 to 3,800 instructions: 5.22 -> 5.91 ms;
// with the ordering / row-write part factored out of the round-count template (one copy, 2,350 instructions
// for two counts: 5.25 ms; 2,760 for three: 5.42 ms) it still lost.)
#define KNN_RR_SMALL(E) ((E) == 1 ? 3 : (E) == 2 ? 5 : 10)
#define KNN_RR_MAX(E) (KNN_CAP_OF(E) / 32)
#define KEY_MAX 0xFFFFFFFFFFFFFFFFull
#define GEN_CS 64  // general path: candidates below the cap that are ranked in one go
// the 27 cells of a block ordered by distance from the centre; index = (dx+1) + 3 (dy+1) + 9 (dz+1)
__constant__ unsigned char GEN_ORDER27[27] = {13, 4, 10, 12, 14, 16, 22, 1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25, 0, 2, 6, 8, 18, 20, 24, 26};
__device__ __forceinline__ int gen_order(int i) { return GEN_ORDER27[i]; }
#define SID_NONE 0xFFFFFFFFu

__device__ __forceinline__ u64 make_key(float d2, u32 orig) {
  return ((u64)__float_as_uint(d2) << 32) | (u64)orig;
}

struct CellWarp {  // per-warp shared-memory views + per-cell constants
  const float4* s_pts;
  const u32* s_id;
  u32* s_d2;
  u32* s_slot;
  int M;
  u32 qs, qn, qoff;
};

__device__ __forceinline__ void count_lt(int& acc, u32 x, u32 key) {
  // acc += (x < key) as one compare + one predicated add
  asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(acc) : "r"(x), "r"(key));
}

// The queries of one cell.  RR = rounds of 32 staged candidates, compile-time so the
// per-round code carries no branches; E = survivor slots per lane (CS = 32 E).
template <int E, int RR>
__device__ __forceinline__ void cell_queries(const CellWarp& w, const float* __restrict__ qb2, int k, float tau0,
                                             float margin,
                                             u32* __restrict__ nbr, float* __restrict__ nd2,
                                             u64* __restrict__ kth, u32* __restrict__ fb_list,
                                             u32* __restrict__ fb_count, int lane) {
  constexpr int CS = 32 * E;
  // survivors the threshold search accepts.  E == 2 (17 <= k <= 32): at most 48, so that the second key of every
  // lane >= 16 is a filler and the upper run of the sorting network is a sorted 16 + fillers — its five-step
  // merge to a run of 32 is never needed (the same cut as a warp-uniform branch on cnt <= 48 cost registers and
  // was slower; as a property of every query it costs a narrower window for the search instead)
  constexpr int CSA = E == 2 ? 48 : CS;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int M = w.M;
  const float target = fminf(margin * (float)k + 2.f, 0.5f * (float)(k + CSA));
  const float tiny = tau0 * 1e-6f;
  const float grow = margin + __fdividef(2.f, (float)k);
  const bool small = M <= CSA;  // every staged candidate survives: the threshold stays at +inf
  float tau = small ? INFINITY : tau0;
  // completeness bounds of the cell's first 32 queries, fetched once: a load at the end of every query sat on
  // its dependency chain (6 % of the kernel's stall samples on that line)
  const float qb_first = lane < (int)w.qn ? qb2[w.qs + lane] : 0.f;
  constexpr int NB = RR < 4 ? 2 : RR < 8 ? 3 : RR < 16 ? 4 : 5;  // bits of a lane's survivor count

  for (u32 q = 0; q < w.qn; ++q) {
    const float4 Q = w.s_pts[w.qoff + q];
    const u32 qid = w.qs + q;
    float d[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      const int slot = lane + 32 * r;
      const float4 P = w.s_pts[slot];  // slots >= M hold +inf coordinates: their distance is +inf
      d[r] = dist2_exact(Q.x, Q.y, Q.z, P.x, P.y, P.z);
    }
    // ---- threshold search: k <= #{d < tau} <= CS (candidate count ~ linear in tau on a surface) ----
    int cnt = 0, cl = 0;
    bool ok = false;
    {
      float lo = 0.f, hi = INFINITY, t = tau, flo = 0.f, fhi = (float)M;
#pragma unroll 1
      for (int it = 0; it < 24; ++it) {
        cl = 0;  // counted per lane, summed by one warp reduction (2 instead of 4 instructions per round)
#pragma unroll
        for (int r = 0; r < RR; ++r) cl += (d[r] < t) ? 1 : 0;
        cnt = __reduce_add_sync(FULL, cl);
        if ((unsigned)(cnt - k) <= (unsigned)(CSA - k)) {  // k <= cnt <= CSA: the usual first answer
          ok = true;
          break;
        }
        const float fc = (float)cnt;
        if (cnt < k) {
          lo = t;
          flo = fc;
        } else if (cnt > CSA) {
          hi = t;
          fhi = fc;
        } else {
          ok = true;
          break;
        }
        float nt;
        if (hi == INFINITY) {
          nt = t * fminf(4.f, fmaxf(1.3f, __fdividef(target, fmaxf(fc, 1.f))));
        } else {
          nt = lo + (hi - lo) * __fdividef(target - flo, fhi - flo);
          if (!(nt > lo && nt < hi)) nt = 0.5f * (lo + hi);
          if (!(nt > lo && nt < hi)) break;  // adjacent floats: a tie cluster larger than CS
        }
        t = nt;
      }
      tau = t;
    }
    bool done = false;
    if (ok) {
      // ---- compact the survivors (d2 bits, slot) into shared memory.  The search has left every lane its
      // own survivor count: the lane's first list position is the exclusive prefix of those counts (one ballot
      // per bit of the count), and its survivors follow each other — per round one compare, two predicated
      // stores and an increment, instead of a ballot + popc pair per round.  (The list is lane-major now; its
      // order never mattered: positions only name the survivors, ties go through the (d2, index) ranking.) ----
      int pos = 0;
#pragma unroll
      for (int b = 0; b < NB; ++b) pos += __popc(__ballot_sync(FULL, (cl >> b) & 1) & lt_mask) << b;
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        if (d[r] < tau) {
          w.s_d2[pos] = __float_as_uint(d[r]);
          w.s_slot[pos] = (u32)(lane + 32 * r);
          ++pos;
        }
      }
      __syncwarp();
      u32 kb = 0;  // bits of the k-th distance (set by the lane that owns it)
      bool ranked = false;
      if (E == 2 && tau < INFINITY) {
        // ---- E == 2: a bitonic network over 64 packed keys, two per lane.  key = (q << 6) | position
        // in the survivor list, q = (unsigned)(d2 * 6e7 / tau) < 2^26: the map d2 -> q is monotone, so
        // keys with different q are in the order of their d2; two survivors with the same q (exact
        // ties, or d2 closer than tau / 6e7: ~1e-5 of the queries) cannot be told apart here and the
        // query goes through the comparison ranking below.  21 compare-exchange steps of ~5
        // instructions against cnt^2 / 16 comparisons (ncu: 227 of the 560 instructions per query).
        // (E == 1, one key per lane, 15 steps: slower than ranking ~21 survivors — k = 16: 4.51 -> 4.99 ms.) ----
        const float scale = __fdividef(6.0e7f, tau);
        u32 ka = 0xFFFFFFFFu, kc = 0xFFFFFFFFu;  // list positions lane and lane + 32
        if (lane < cnt) ka = ((u32)(__uint_as_float(w.s_d2[lane]) * scale) << 6) | (u32)lane;
        if (lane + 32 < cnt) kc = ((u32)(__uint_as_float(w.s_d2[lane + 32]) * scale) << 6) | (u32)(lane + 32);
        // every compare-exchange is ascending (the lower position keeps the smaller key): a merge of
        // two sorted runs of k2 / 2 starts with the MIRROR step (partner = position ^ (k2 - 1)), then
        // halves the distance.  Which side a lane is on depends on one bit of its number only.
        const bool l1 = (lane & 1) == 0, l2 = (lane & 2) == 0, l4 = (lane & 4) == 0, l8 = (lane & 8) == 0,
                   l16 = (lane & 16) == 0;
#define CX(mask, lower)                                        \
  {                                                            \
    const u32 ta = __shfl_xor_sync(FULL, ka, mask), tc = __shfl_xor_sync(FULL, kc, mask); \
    ka = (lower) ? min(ka, ta) : max(ka, ta);                  \
    kc = (lower) ? min(kc, tc) : max(kc, tc);                  \
  }
        CX(1, l1);                                            // runs of 2
        CX(3, l2);  CX(1, l1);                                // 4
        CX(7, l4);  CX(2, l2);  CX(1, l1);                    // 8
        CX(15, l8); CX(4, l4);  CX(2, l2); CX(1, l1);         // 16
#define CX1(mask, lower)                                       \
  {                                                            \
    const u32 ta = __shfl_xor_sync(FULL, ka, mask);            \
    ka = (lower) ? min(ka, ta) : max(ka, ta);                  \
  }
        // 32: ka becomes a sorted run; kc (<= 16 survivors in lanes 0..15, fillers above) already is one
        CX1(31, l16); CX1(8, l8); CX1(4, l4); CX1(2, l2); CX1(1, l1);
        // (The same cut as a warp-uniform branch on cnt <= 48, with the search still accepting 64: 56 registers, four
        // blocks per SM instead of five: 4.44 -> 4.84 ms; held at 48 by __launch_bounds__ it spilled: 4.62 ms.  Also
        // rejected: the comparison ranking below as a real call (__noinline__) for k <= 32, to take its ~600 cold
        // instructions out of this loop's instruction stream: the call convention cost 64 registers, or 28 bytes
        // of spills at 48: 5.30 ms.)
        {  // 64: position p of the first run against position 63 - p = second run, lane 31 - p
          const u32 tc = __shfl_xor_sync(FULL, kc, 31), ta = __shfl_xor_sync(FULL, ka, 31);
          ka = min(ka, tc);
          kc = max(kc, ta);
        }
        // k <= 32 here (E == 2): only the lower half is ever written, so only ka is sorted on
        // (five steps on one key instead of two); of the upper half only its minimum matters — it is the
        // successor of position 31 in the tie check
        CX1(16, l16); CX1(8, l8); CX1(4, l4); CX1(2, l2); CX1(1, l1);
#undef CX1
#undef CX
        u32 na = __shfl_down_sync(FULL, ka, 1);
        const u32 c0 = __reduce_min_sync(FULL, kc);
        if (lane == 31) na = c0;
        const bool tie = (ka >> 6) == (na >> 6) && na != 0xFFFFFFFFu;
        if (!__any_sync(FULL, tie)) {
          ranked = true;
          if (lane < k) {  // cnt >= k: the first k positions hold survivors
            const u32 i = ka & 63u;
            const u32 slot = w.s_slot[i], db = w.s_d2[i];
            nbr[(size_t)qid * k + lane] = w.s_id[slot];
            nd2[(size_t)qid * k + lane] = __uint_as_float(db);
            if (lane == k - 1) {
              kth[qid] = ((u64)db << 32) | (u64)__float_as_uint(w.s_pts[slot].w);
              kb = db;
            }
          }
        }
      }
      if (!ranked) {
      // pad the survivor list to a multiple of eight with a value no key is below, so that the
      // ranking loop below runs on whole pairs of uint4 words only
      if (lane < 7 && cnt + lane < CS) w.s_d2[cnt + lane] = 0xFFFFFFFFu;
      __syncwarp();
      // ---- rank = output slot.  Ranks are counted on d2 alone (non-negative floats order
      // as unsigned ints); an exact d2 tie makes two ranks collide, which shows as a rank
      // sum below cnt(cnt-1)/2 and sends the query through the full (d2, index) compare. ----
      u32 md[E];
      int rk[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        md[e] = (lane + 32 * e < cnt) ? w.s_d2[lane + 32 * e] : 0xFFFFFFFFu;
        rk[e] = 0;
      }
      {
        const uint4* v4 = (const uint4*)w.s_d2;
        const int n8 = (cnt + 7) >> 3;
#pragma unroll 1
        for (int i = 0; i < n8; ++i) {
          const uint4 x = v4[2 * i], y = v4[2 * i + 1];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            count_lt(rk[e], x.x, md[e]);
            count_lt(rk[e], x.y, md[e]);
            count_lt(rk[e], x.z, md[e]);
            count_lt(rk[e], x.w, md[e]);
            count_lt(rk[e], y.x, md[e]);
            count_lt(rk[e], y.y, md[e]);
            count_lt(rk[e], y.z, md[e]);
            count_lt(rk[e], y.w, md[e]);
          }
        }
      }
      int rsum = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) rsum += (lane + 32 * e < cnt) ? rk[e] : 0;
      rsum = __reduce_add_sync(FULL, rsum);
      if (rsum != (cnt * (cnt - 1)) / 2) {  // exact distance ties: canonical (d2, original index) order
        u32 mo[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          rk[e] = 0;
          mo[e] = (lane + 32 * e < cnt) ? __float_as_uint(w.s_pts[w.s_slot[lane + 32 * e]].w) : 0xFFFFFFFFu;
        }
        for (int i = 0; i < cnt; ++i) {
          const u32 x = w.s_d2[i];
          const u32 xo = __float_as_uint(w.s_pts[w.s_slot[i]].w);
#pragma unroll
          for (int e = 0; e < E; ++e) rk[e] += (x < md[e] || (x == md[e] && xo < mo[e])) ? 1 : 0;
        }
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (lane + 32 * e < cnt && rk[e] < k) {
          const u32 slot = w.s_slot[lane + 32 * e];
          nbr[(size_t)qid * k + rk[e]] = w.s_id[slot];
          nd2[(size_t)qid * k + rk[e]] = __uint_as_float(md[e]);
          if (rk[e] == k - 1) {
            kth[qid] = ((u64)md[e] << 32) | (u64)__float_as_uint(w.s_pts[slot].w);
            kb = md[e];
          }
        }
      }
      }  // !ranked
      const float kd2 = __uint_as_float(__reduce_max_sync(FULL, kb));
      done = kd2 <= (q < 32u ? __shfl_sync(FULL, qb_first, (int)q) : qb2[qid]);
      tau = small ? INFINITY : fmaxf(kd2 * grow, tiny);
      __syncwarp();  // the next query's compaction overwrites the lists read above
    } else {
      tau = small ? INFINITY : tau0;
    }
    // bit 31: the one-ring block is known to be too small, the general path starts at stage 1
    if (!done && lane == 0) fb_list[atomicAdd(fb_count, 1u)] = ok ? (qid | 0x80000000u) : qid;
  }
}

// CAP = staged candidates per warp.  The first pass (SECOND = false) runs one warp per occupied cell
// with CAP = KNN_CAP_OF(E); a cell whose 3x3x3 block holds more (1 % of the cells of cfg3, 97 k of
// its 104 k fallback queries — all of them below 384 candidates) is appended to `ovf_list` and
// answered by a SECOND launch of the same code with a larger staging buffer (fewer warps resident,
// but few cells) instead of by the general path, which costs ~8 times as much per query.
template <int E, int CAP, bool SECOND>  // E = CS / 32 survivor slots per lane
// (k <= 32, first launch: five blocks per SM is what the shared memory allows and the registers must follow —
// 48 today.  Check -Xptxas -v after every change to this kernel.)
__global__ void __launch_bounds__(KNN_WARPS_OF(E) * 32)
knn_cell_kernel(GridView g, const float4* __restrict__ spts, const float* __restrict__ qb2, int k,
                float tau0, float margin, u32* __restrict__ nbr, float* __restrict__ nd2, u64* __restrict__ kth,
                u32* __restrict__ fb_list, u32* __restrict__ fb_count, u32* __restrict__ ovf_list,
                u32* __restrict__ ovf_count, int cap_second) {
  constexpr int CS = 32 * E;
  constexpr int KNN_CAP = CAP;
  constexpr int KNN_WARPS = KNN_WARPS_OF(E);
  constexpr int RR_SMALL = SECOND ? CAP / 32 : KNN_RR_SMALL(E);
  constexpr int RR_MAX = CAP / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_pts_all = (float4*)smem_raw;                        // [W][CAP] staged candidates
  u32* s_d2_all = (u32*)(s_pts_all + KNN_WARPS * KNN_CAP);      // [W][CS] survivor d2 bits (16 B aligned)
  u32* s_id_all = s_d2_all + KNN_WARPS * CS;                    // [W][CAP] sorted ids of the candidates
  u32* s_slot_all = s_id_all + KNN_WARPS * KNN_CAP;             // [W][CS] survivor slots

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* s_pts = s_pts_all + warp * KNN_CAP;
  u32* s_id = s_id_all + warp * KNN_CAP;
  const unsigned FULL = 0xffffffffu;
  // warps are independent: no block-level barrier below
  const u32 n_items = SECOND ? *ovf_count : g.n_cells;
  const u32 stride = gridDim.x * KNN_WARPS;
  for (u32 item = blockIdx.x * KNN_WARPS + warp; item < n_items; item += stride) {
  const u32 ci = SECOND ? ovf_list[item] : item;

  // ---- the 27 cells of the block ----
  u32 cx, cy, cz;
  demorton3(g.cell_key[ci], cx, cy, cz);
  u32 nstart = 0, ncnt = 0;
  if (lane < 27) {
    int x = (int)cx + (lane % 3) - 1, y = (int)cy + ((lane / 3) % 3) - 1, z = (int)cz + (lane / 9) - 1;
    if (x >= 0 && x < g.dimx && y >= 0 && y < g.dimy && z >= 0 && z < g.dimz) {
      u32 cidx = (lane == 13) ? ci : grid_lookup(g, morton3((u32)x, (u32)y, (u32)z));
      if (cidx != 0xFFFFFFFFu) {
        nstart = g.cell_start[cidx];
        ncnt = g.cell_start[cidx + 1] - nstart;
      }
    }
  }
  u32 incl = ncnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  const u32 excl = incl - ncnt;
  const int M = (int)__shfl_sync(FULL, incl, 31);
  const u32 qs = __shfl_sync(FULL, nstart, 13), qn = __shfl_sync(FULL, ncnt, 13);
  const u32 qoff = __shfl_sync(FULL, excl, 13);

  if (!SECOND && M > KNN_CAP && M <= cap_second) {  // crowded: the second launch stages it
    if (lane == 0) ovf_list[atomicAdd(ovf_count, 1u)] = ci;
    return;
  }
  if (M > KNN_CAP || M < k) {  // too crowded to stage, or not enough candidates: general path
    u32 base = 0;
    if (lane == 0) base = atomicAdd(fb_count, qn);
    base = __shfl_sync(FULL, base, 0);
    for (u32 t = lane; t < qn; t += 32) fb_list[base + t] = qs + t;
    if (SECOND) continue;
    return;
  }

  // ---- stage the candidates (coalesced float4 loads), 32 slots per round whatever cell they come from.
  // A loop over the occupied cells (one or two dozen per block, ~15 points each) spent ~35 instructions per
  // cell on one half-empty round: 32 of the 439 warp instructions per query (ncu source page, round 2).  Here
  // the occupied cells are numbered in slot order, cell c keeps (first point - first slot) in shared memory,
  // and a slot finds its cell by counting the cell starts at or before it: one redux.or + popc per round. ----
  {
    u32* s_delta = s_d2_all + warp * CS;  // <= 27 entries; the survivor list is not in use yet
    const bool occ_here = ncnt != 0;
    const unsigned occupied = __ballot_sync(FULL, occ_here);
    if (occ_here) s_delta[__popc(occupied & ((1u << lane) - 1u))] = nstart - excl;
    __syncwarp();
    int cell_of_round = -1;  // occupied cells that start before this round, minus one
    for (int t0 = 0; t0 < M; t0 += 32) {
      const u32 rel = excl - (u32)t0;
      const unsigned starts = __reduce_or_sync(FULL, (occ_here && rel < 32u) ? (1u << rel) : 0u);
      const int t = t0 + lane;
      if (t < M) {
        const u32 src = s_delta[cell_of_round + __popc(starts & (0xFFFFFFFFu >> (31 - lane)))] + (u32)t;
        s_pts[t] = spts[src];
        s_id[t] = src;
      }
      cell_of_round += __popc(starts);
    }
  }  // (the __syncwarp below also hands s_delta back to the survivor list)
  // the compiled round count covers the staged rounds; the unstaged slots it touches hold +inf
  // coordinates, so their distance is +inf and no threshold ever admits them
  const int R = (M + 31) >> 5;
  const int RRu = R <= RR_SMALL ? RR_SMALL : RR_MAX;
  for (int t = M + lane; t < RRu * 32; t += 32) s_pts[t] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
  __syncwarp();

  CellWarp w;
  w.s_pts = s_pts;
  w.s_id = s_id;
  w.s_d2 = s_d2_all + warp * CS;
  w.s_slot = s_slot_all + warp * CS;
  w.M = M;
  w.qs = qs;
  w.qn = qn;
  w.qoff = qoff;
  if (R <= RR_SMALL) cell_queries<E, RR_SMALL>(w, qb2, k, tau0, margin, nbr, nd2, kth, fb_list, fb_count, lane);
  else cell_queries<E, RR_MAX>(w, qb2, k, tau0, margin, nbr, nd2, kth, fb_list, fb_count, lane);
  if (!SECOND) break;
  __syncwarp();  // the next cell of this warp restages the buffers
  }
}

template <int RL>
struct WarpTopK {
  u64 key[RL];
  u32 sid[RL];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      key[r] = KEY_MAX;
      sid[r] = SID_NONE;
    }
  }
  __device__ __forceinline__ u64 at_key(int p) const {
    u64 v = 0;
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      u64 t = __shfl_sync(0xffffffffu, key[r], p & 31);
      if ((p >> 5) == r) v = t;
    }
    return v;
  }
  // x / xs are warp-uniform
  __device__ __forceinline__ void insert(u64 x, u32 xs, int lane) {
    int pos = 0;
#pragma unroll
    for (int r = 0; r < RL; ++r) pos += __popc(__ballot_sync(0xffffffffu, key[r] < x));
    if (pos >= 32 * RL) return;
#pragma unroll
    for (int r = RL - 1; r >= 0; --r) {
      const int base = r * 32;
      if (pos >= base + 32) continue;
      u64 upk = __shfl_up_sync(0xffffffffu, key[r], 1);
      u32 ups = __shfl_up_sync(0xffffffffu, sid[r], 1);
      if (r > 0) {
        u64 ck = __shfl_sync(0xffffffffu, key[r - 1], 31);
        u32 cs = __shfl_sync(0xffffffffu, sid[r - 1], 31);
        if (lane == 0) {
          upk = ck;
          ups = cs;
        }
      }
      const int g = base + lane;
      if (g == pos) {
        key[r] = x;
        sid[r] = xs;
      } else if (g > pos) {
        key[r] = upk;
        sid[r] = ups;
      }
    }
  }
};

__device__ __forceinline__ u32 lower_bound_u64(const u64* a, u32 n, u64 v) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <int RL, bool EXTERNAL>
__global__ void __launch_bounds__(256)
knn_general_kernel(GridView g, const float4* __restrict__ spts, const float4* __restrict__ queries,
                   const u32* __restrict__ list, const u32* __restrict__ count_ptr, u32 m_external,
                   int k, u32* __restrict__ nbr, float* __restrict__ nd2, u64* __restrict__ kth) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  const u32 total = EXTERNAL ? m_external : *count_ptr;
  WarpTopK<RL> top;
  __shared__ u64 s_ckey[EXTERNAL ? 1 : 8][2 * GEN_CS];  // per warp: collected keys | keys by rank
  __shared__ u32 s_csid[EXTERNAL ? 1 : 8][2 * GEN_CS];
  u64* c_key = s_ckey[EXTERNAL ? 0 : (threadIdx.x >> 5)];
  u64* o_key = c_key + GEN_CS;
  u32* c_sid = s_csid[EXTERNAL ? 0 : (threadIdx.x >> 5)];
  u32* o_sid = c_sid + GEN_CS;
  for (u32 it = gw; it < total; it += nw) {
    u32 row;
    int j0 = 0;
    u64 cap = KEY_MAX;
    float4 Q;
    if (EXTERNAL) {
      row = it;
      Q = queries[it];
    } else {
      row = list[it];
      j0 = (int)(row >> 31);
      row &= 0x7FFFFFFFu;
      Q = spts[row];
      // the fast path found k candidates inside the one-ring block (it only could not prove them
      // complete): the key of its k-th bounds the true k-th from above, so nothing beyond it
      // needs to enter the list
      if (j0) cap = kth[row];
    }
    const double vx = ((double)Q.x - g.ox) * g.inv_cell, vy = ((double)Q.y - g.oy) * g.inv_cell,
                 vz = ((double)Q.z - g.oz) * g.inv_cell;
    const int cx = min(g.dimx - 1, max(0, (int)floor(vx)));
    const int cy = min(g.dimy - 1, max(0, (int)floor(vy)));
    const int cz = min(g.dimz - 1, max(0, (int)floor(vz)));
    for (int j = j0;; ++j) {
      top.reset();
      const int sx = cx >> j, sy = cy >> j, sz = cz >> j;
      const int sdx = ((g.dimx - 1) >> j) + 1, sdy = ((g.dimy - 1) >> j) + 1, sdz = ((g.dimz - 1) >> j) + 1;
      // point ranges of the 27 super-cells (Morton prefix ranges)
      u32 pstart = 0, pend = 0;
      if (lane < 27) {
        int x = sx + (lane % 3) - 1, y = sy + ((lane / 3) % 3) - 1, z = sz + (lane / 9) - 1;
        if (x >= 0 && x < sdx && y >= 0 && y < sdy && z >= 0 && z < sdz) {
          u64 pfx = morton3((u32)x, (u32)y, (u32)z);
          u64 klo = pfx << (3 * j), khi = (pfx + 1) << (3 * j);
          u32 a = lower_bound_u64(g.cell_key, g.n_cells, klo);
          u32 b = lower_bound_u64(g.cell_key, g.n_cells, khi);
          pstart = g.cell_start[a];
          pend = g.cell_start[b];
        }
      }
      // A query the fast path handed over with its k-th key as a cap needs exactly the candidates
      // below the cap: the k it already found plus the few of the second ring that beat them.
      // Inserting them one at a time into the sorted list was 60 % of this kernel's instructions
      // (~80 per insertion, ncu source page); they are COLLECTED instead (ballot compaction into
      // shared memory) and ranked against each other once — the rank is the list position.  More
      // than GEN_CS of them (exact ties on the cap) and the pass is repeated with insertions.
      bool collect = !EXTERNAL && j0 == 1 && j == 1;
      for (int pass = 0; pass < 2; ++pass) {
        u32 ccnt = 0;
        for (int o27 = 0; o27 < 27; ++o27) {
          // nearest cells first (centre, faces, edges, corners): the list fills with near points and
          // its k-th key — the admission threshold — is tight before the far cells are read
          const int c27 = gen_order(o27);
          u32 s = __shfl_sync(FULL, pstart, c27), e = __shfl_sync(FULL, pend, c27);
          for (u32 base = s; base < e; base += 32) {
            u32 pi = base + lane;
            u64 ck = KEY_MAX;
            if (pi < e) {
              float4 P = spts[pi];
              ck = make_key(dist2_exact(Q.x, Q.y, Q.z, P.x, P.y, P.z), __float_as_uint(P.w));
            }
            if (collect) {
              const bool p = ck <= cap;  // KEY_MAX (no point) never is: cap carries a finite distance
              const unsigned m = __ballot_sync(FULL, p);
              const u32 pos = ccnt + (u32)__popc(m & ((1u << lane) - 1u));
              if (p && pos < GEN_CS) {
                c_key[pos] = ck;
                c_sid[pos] = pi;
              }
              ccnt += (u32)__popc(m);
              continue;
            }
            u64 thr = top.at_key(k - 1);
            unsigned mask = __ballot_sync(FULL, ck < thr && ck <= cap);
            while (mask) {
              int src = __ffs(mask) - 1;
              mask &= mask - 1;
              u64 x = __shfl_sync(FULL, ck, src);
              top.insert(x, base + (u32)src, lane);
            }
          }
        }
        if (!collect) break;
        collect = false;
        if (ccnt > GEN_CS) continue;  // does not fit: second pass, by insertion
        __syncwarp();
        // rank = position in the sorted list (keys are unique: (d2, original index))
        u64 mk[2];
        int rk[2] = {0, 0};
#pragma unroll
        for (int e = 0; e < 2; ++e) mk[e] = (u32)(lane + 32 * e) < ccnt ? c_key[lane + 32 * e] : KEY_MAX;
        for (u32 i = 0; i < ccnt; ++i) {
          const u64 x = c_key[i];
          rk[0] += x < mk[0] ? 1 : 0;
          rk[1] += x < mk[1] ? 1 : 0;
        }
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if ((u32)(lane + 32 * e) < ccnt) {
            o_key[rk[e]] = mk[e];
            o_sid[rk[e]] = c_sid[lane + 32 * e];
          }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          const u32 p = (u32)(r * 32 + lane);
          top.key[r] = p < ccnt ? o_key[p] : KEY_MAX;
          top.sid[r] = p < ccnt ? o_sid[p] : SID_NONE;
        }
        __syncwarp();
        break;
      }
      // distance from the query to the nearest unscanned cell: faces of the block
      // that still have cells behind them
      const double cs = g.cell * (double)(1 << j);
      double bound = INFINITY;
      {
        int bx0 = max(sx - 1, 0), bx1 = min(sx + 1, sdx - 1);
        int by0 = max(sy - 1, 0), by1 = min(sy + 1, sdy - 1);
        int bz0 = max(sz - 1, 0), bz1 = min(sz + 1, sdz - 1);
        if (bx0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.x - g.ox) - bx0 * cs));
        if (bx1 < sdx - 1) bound = fmin(bound, fmax(0.0, (bx1 + 1) * cs - ((double)Q.x - g.ox)));
        if (by0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.y - g.oy) - by0 * cs));
        if (by1 < sdy - 1) bound = fmin(bound, fmax(0.0, (by1 + 1) * cs - ((double)Q.y - g.oy)));
        if (bz0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.z - g.oz) - bz0 * cs));
        if (bz1 < sdz - 1) bound = fmin(bound, fmax(0.0, (bz1 + 1) * cs - ((double)Q.z - g.oz)));
      }
      if (isinf(bound)) break;  // the block covers the whole grid
      u64 kk = top.at_key(k - 1);
      if (kk != KEY_MAX) {
        double b = bound * (1.0 - 1e-6);
        float kd2 = __uint_as_float((u32)(kk >> 32));
        if (kd2 <= __double2float_rd(b * b)) break;
      }
    }
    // ---- write the row ----
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      int p = r * 32 + lane;
      if (p < k) {
        nbr[(size_t)row * k + p] = top.sid[r];
        if (nd2)
          nd2[(size_t)row * k + p] =
              top.sid[r] == SID_NONE ? INFINITY : __uint_as_float((u32)(top.key[r] >> 32));
        if (p == k - 1 && kth) kth[row] = top.key[r];
      }
    }
  }
}

// A handful of external queries with k = 1 (the centroid snap of the seed stage: ~10^2 queries): one BLOCK per
// query instead of one warp.  A centroid of a curved cluster lies off the surface, its block has to grow over
// several stages, and a single warp scanning them was a 0.12 ms latency chain for 190 queries; eight warps share the
// 27 super-cells of a stage.  Same stages, same stopping rule, same canonical (d2, original index) order as the
// general path — the nearest point is the nearest point.
__global__ void __launch_bounds__(256)
knn_nearest_block_kernel(GridView g, const float4* __restrict__ spts, const float4* __restrict__ queries, u32 m,
                         u32* __restrict__ out_sid, float* __restrict__ out_d2) {
  __shared__ u32 s_ps[27], s_pe[27];
  __shared__ u64 s_key[8];
  __shared__ u32 s_sid[8];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 it = blockIdx.x;
  if (it >= m) return;
  const float4 Q = queries[it];
  const double vx = ((double)Q.x - g.ox) * g.inv_cell, vy = ((double)Q.y - g.oy) * g.inv_cell,
               vz = ((double)Q.z - g.oz) * g.inv_cell;
  const int cx = min(g.dimx - 1, max(0, (int)floor(vx)));
  const int cy = min(g.dimy - 1, max(0, (int)floor(vy)));
  const int cz = min(g.dimz - 1, max(0, (int)floor(vz)));
  u64 best = KEY_MAX;
  u32 bsid = SID_NONE;
  for (int j = 0;; ++j) {
    const int sx = cx >> j, sy = cy >> j, sz = cz >> j;
    const int sdx = ((g.dimx - 1) >> j) + 1, sdy = ((g.dimy - 1) >> j) + 1, sdz = ((g.dimz - 1) >> j) + 1;
    if (warp == 0 && lane < 27) {  // point ranges of the 27 super-cells (Morton prefix ranges)
      u32 pstart = 0, pend = 0;
      int x = sx + (lane % 3) - 1, y = sy + ((lane / 3) % 3) - 1, z = sz + (lane / 9) - 1;
      if (x >= 0 && x < sdx && y >= 0 && y < sdy && z >= 0 && z < sdz) {
        u64 pfx = morton3((u32)x, (u32)y, (u32)z);
        u64 klo = pfx << (3 * j), khi = (pfx + 1) << (3 * j);
        u32 a = lower_bound_u64(g.cell_key, g.n_cells, klo);
        u32 b = lower_bound_u64(g.cell_key, g.n_cells, khi);
        pstart = g.cell_start[a];
        pend = g.cell_start[b];
      }
      s_ps[lane] = pstart;
      s_pe[lane] = pend;
    }
    __syncthreads();
    best = KEY_MAX;
    bsid = SID_NONE;
    for (int c27 = warp; c27 < 27; c27 += 8) {
      const u32 e = s_pe[c27];
      for (u32 pi = s_ps[c27] + (u32)lane; pi < e; pi += 32u) {
        const float4 P = spts[pi];
        const u64 key = make_key(dist2_exact(Q.x, Q.y, Q.z, P.x, P.y, P.z), __float_as_uint(P.w));
        if (key < best) {
          best = key;
          bsid = pi;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 ok = __shfl_xor_sync(FULL, best, o);
      const u32 os = __shfl_xor_sync(FULL, bsid, o);
      if (ok < best) {
        best = ok;
        bsid = os;
      }
    }
    if (lane == 0) {
      s_key[warp] = best;
      s_sid[warp] = bsid;
    }
    __syncthreads();
    best = s_key[0];
    bsid = s_sid[0];
#pragma unroll
    for (int w2 = 1; w2 < 8; ++w2)
      if (s_key[w2] < best) {
        best = s_key[w2];
        bsid = s_sid[w2];
      }
    // distance from the query to the nearest unscanned cell (as in knn_general_kernel)
    const double cs = g.cell * (double)(1 << j);
    double bound = INFINITY;
    {
      int bx0 = max(sx - 1, 0), bx1 = min(sx + 1, sdx - 1);
      int by0 = max(sy - 1, 0), by1 = min(sy + 1, sdy - 1);
      int bz0 = max(sz - 1, 0), bz1 = min(sz + 1, sdz - 1);
      if (bx0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.x - g.ox) - bx0 * cs));
      if (bx1 < sdx - 1) bound = fmin(bound, fmax(0.0, (bx1 + 1) * cs - ((double)Q.x - g.ox)));
      if (by0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.y - g.oy) - by0 * cs));
      if (by1 < sdy - 1) bound = fmin(bound, fmax(0.0, (by1 + 1) * cs - ((double)Q.y - g.oy)));
      if (bz0 > 0) bound = fmin(bound, fmax(0.0, ((double)Q.z - g.oz) - bz0 * cs));
      if (bz1 < sdz - 1) bound = fmin(bound, fmax(0.0, (bz1 + 1) * cs - ((double)Q.z - g.oz)));
    }
    if (isinf(bound)) break;  // the block covers the whole grid
    if (best != KEY_MAX) {
      double b = bound * (1.0 - 1e-6);
      float kd2 = __uint_as_float((u32)(best >> 32));
      if (kd2 <= __double2float_rd(b * b)) break;
    }
    __syncthreads();  // every thread has read this stage's tables before the next stage rewrites them
  }
  if (threadIdx.x == 0) {
    out_sid[it] = bsid;
    if (out_d2) out_d2[it] = bsid == SID_NONE ? INFINITY : __uint_as_float((u32)(best >> 32));
  }
}

// ---------------------------------------------------------------------------
static size_t cell_kernel_smem(int E, int cap) {
  int CS = 32 * E;
  return (size_t)KNN_WARPS_OF(E) * (cap * 16 + cap * 4 + CS * 4 + CS * 4) + 16;
}
#define KNN_CAP2 384  // second launch (k <= 32): every crowded cell of cfg3 holds fewer candidates than this

int knn_run(iftwt_ctx* c, int k) {
  if (k < 1 || k > 64) return ctx_fail(c, IFTWT_ERR_INVALID, "k must be in [1, 64], got %d", k);
  TRY(grid_build(c, k));
  if (c->knn_k == k) return IFTWT_OK;
  TRY(graph_detach_rows(c));  // the graph keeps its rows if they alias the lists we overwrite
  c->mask_k = 0;
  c->rdeg_k = 0;
  c->rg_pre_valid = false;
  StageTimer st(c, IFTWT_STAGE_KNN);
  const size_t n = c->n;
  ENSURE(c, c->nbr, n * k * 4);
  ENSURE(c, c->nd2, n * k * 4);
  ENSURE(c, c->kth, n * 8);
  ENSURE(c, c->fb_list, n * 8);  // [n] fallback queries | [n] crowded cells for the second launch
  u32* fb_count = (u32*)(c->scalars.as<int>() + 48);
  u32* ovf_count = fb_count + 1;
  u32* ovf_list = c->fb_list.as<u32>() + n;
  CU_TRY(c, cudaMemsetAsync(fb_count, 0, 8, c->stream));
  const GridView& g = c->grid;
  float tau0 = (float)(g.cell * g.cell * 0.6);
  const float4* spts = c->spts.as<float4>();
  cudaEventRecord(c->ev0[IFTWT_STAGE_KNN_CELL], c->stream);
  {
    const int E = k <= 16 ? 1 : (k <= 32 ? 2 : 4);
    auto kern = E == 1 ? knn_cell_kernel<1, KNN_CAP_OF(1), false>
              : (E == 2 ? knn_cell_kernel<2, KNN_CAP_OF(2), false> : knn_cell_kernel<4, KNN_CAP_OF(4), false>);
    size_t sm = cell_kernel_smem(E, KNN_CAP_OF(E));
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const char* menv = getenv("IFTWT_KNN_MARGIN");
    // survivors aimed at = margin * k + 2.  Comparison ranking (cost ~ cnt^2) wants few: swept on B200
    // (cfg3): 1.05 -> 5.76 ms, 1.12 -> 5.77, 1.25 -> 5.87, 1.4 -> 6.01.  The 64-key network of E == 2 costs
    // the same for any cnt <= 64, so a wider margin only saves threshold iterations: 1.1 -> 5.13 ms,
    // 1.2 -> 5.10, 1.3 / 1.4 -> 5.09
    // (E == 2 with at most 48 survivors accepted: 1.3 -> 4.36 ms, 1.2 -> 4.33, 1.15 -> 4.33)
    const float margin = menv ? (float)atof(menv) : (E == 2 ? 1.2f : 1.1f);
    const unsigned blocks = div_up(g.n_cells, KNN_WARPS_OF(E));
    const int cap2 = E <= 2 ? KNN_CAP2 : 0;
    LAUNCH(c, kern, blocks, KNN_WARPS_OF(E) * 32, sm, g, spts, (const float*)c->qb2.as<float>(), k, tau0, margin,
           c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>(), c->fb_list.as<u32>(), fb_count, ovf_list,
           ovf_count, cap2);
    if (cap2) {  // crowded cells: the same code with a larger staging buffer, warps stride over the list
      auto kern2 = E == 1 ? knn_cell_kernel<1, KNN_CAP2, true> : knn_cell_kernel<2, KNN_CAP2, true>;
      const size_t sm2 = cell_kernel_smem(E, KNN_CAP2);
      CU_TRY(c, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      LAUNCH(c, kern2, (unsigned)c->sm_count * 3, KNN_WARPS_OF(E) * 32, sm2, g, spts, (const float*)c->qb2.as<float>(),
             k, tau0, margin, c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>(), c->fb_list.as<u32>(), fb_count,
             ovf_list, ovf_count, 0);
    }
  }
  cudaEventRecord(c->ev1[IFTWT_STAGE_KNN_CELL], c->stream);
  c->ev_set[IFTWT_STAGE_KNN_CELL] = true;
  CHECK_LAUNCH(c);
  unsigned gblocks = c->sm_count * 8;
  {
    auto kern = (k <= 32) ? knn_general_kernel<1, false> : knn_general_kernel<2, false>;
    LAUNCH(c, kern, gblocks, 256, 0, g, spts, (const float4*)nullptr, (const u32*)c->fb_list.as<u32>(),
           (const u32*)fb_count, 0u, k, c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>());
  }
  CHECK_LAUNCH(c);
  int* hs = (int*)c->h_scalars;
  TRY(host_fetch(c, hs + 48, fb_count, 4));
  c->knn_k = k;
  return IFTWT_OK;
}

// external queries (device float4 {x,y,z,_}); idx in SORTED space, SID_NONE when missing
int knn_query_run(iftwt_ctx* c, const float4* d_q, size_t m, int k, u32* d_idx_sorted, float* d_d2) {
  if (k < 1 || k > 64) return ctx_fail(c, IFTWT_ERR_INVALID, "k must be in [1, 64], got %d", k);
  if (!c->grid_valid) TRY(grid_build(c, k < 8 ? 8 : k));
  if (m == 0) return IFTWT_OK;
  if (k == 1 && m <= 2048 && !(getenv("IFTWT_NEAREST_BLOCK") && atoi(getenv("IFTWT_NEAREST_BLOCK")) == 0)) {
    LAUNCH(c, knn_nearest_block_kernel, (unsigned)m, 256, 0, c->grid, (const float4*)c->spts.as<float4>(), d_q, (u32)m,
           d_idx_sorted, d_d2);
    CHECK_LAUNCH(c);
    return IFTWT_OK;
  }
  unsigned gblocks = (unsigned)((m + 7) / 8);
  if (gblocks > (unsigned)c->sm_count * 8) gblocks = c->sm_count * 8;
  {
    auto kern = (k <= 32) ? knn_general_kernel<1, true> : knn_general_kernel<2, true>;
    LAUNCH(c, kern, gblocks, 256, 0, c->grid, (const float4*)c->spts.as<float4>(), d_q,
           (const u32*)nullptr, (const u32*)nullptr, (u32)m, k, d_idx_sorted, d_d2, (u64*)nullptr);
  }
  CHECK_LAUNCH(c);
  return IFTWT_OK;
}
