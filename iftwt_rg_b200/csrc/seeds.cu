// seeds.cu — K4 / K4': region-growing seed selection.
//
// Replaces FlatPoint::getCentroidsCloud (flat_point.hpp:50-56 -> compute_centroid_info
// 65-91 -> getCentroids 23-42 -> pcl::RegionGrowing::extract, parameters main.cpp:112-119).
//
// PCL's region growing is a sequential seeded BFS over points sorted by curvature.
// With the curvature gate inactive (it is for the reference's settings: curvature
// <= 1/3 < threshold 1.0) its result has an exact order-free form (SURVEY App. A.4,
// proved against the literal procedure in tests/test_oracle_ift_rg.py):
//     segment(v) = min rank over all u that reach v along arcs u -> w,
//                  w in kNN(u), |n_u . n_w| >= cos(theta),
// rank = position in the (curvature, index) order.  It is computed here as
//   1. the order as a 64-bit key per vertex (curvature bits << 32 | index), never as ranks;
//   2. union-find over the MUTUAL smooth arcs (both u in kNN(w) and w in kNN(u);
//      the test itself is symmetric), hooking + path halving;
//   3. component min-key, then label-min sweeps over the remaining ONE-WAY smooth
//      arcs between components until nothing changes;
//   4. segment sizes, the [min, max] filter, PCL's emission order (ascending seed
//      key: a sort of the kept clusters only), and the literal FP32 centroid: pcl::compute3DCentroid sums a
//      cluster's points in ascending point index (assembleRegions order), so the
//      members are sorted by (cluster, index) and one thread per cluster adds them
//      sequentially;
//   5. the snap of every centroid to its nearest input point (k-NN general path).
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

namespace {

__device__ __forceinline__ u32 uf_find(u32* parent, u32 v) {
  u32 p = parent[v];
  while (p != v) {
    u32 gp = parent[p];
    if (gp != p) parent[v] = gp;  // path halving (benign race)
    v = p;
    p = gp;
  }
  return v;
}
// Eigen's unrolled 3-element reduction: e0 + (e1 + e2)   (oracle: dot3_eigen)
__device__ __forceinline__ bool smooth_pass(const float4 nu, const float4 nv, float cos_thr) {
  float dp = fabsf(nv.x * nu.x + (nv.y * nu.y + nv.z * nu.z));
  return !(dp < cos_thr);  // NaN passes, as in PCL
}

// The (curvature, original index) order PCL's region growing seeds by is carried as a 64-bit key
// per vertex — curvature bits (non-negative floats order as unsigned ints) << 32 | index — and
// never materialised as ranks: only comparisons are needed, so the 10 M-pair radix sort the
// ranks used to cost (0.55 ms) is replaced by a sort of the KEPT clusters' keys alone.
__device__ __forceinline__ u64 rg_key(const float4* __restrict__ nrm, const float4* __restrict__ spts, u32 v) {
  return ((u64)__float_as_uint(nrm[v].w) << 32) | (u64)__float_as_uint(spts[v].w);
}
// Region-growing arcs in two kernels.
//
// rg_classify_pow2_kernel — the form every BASELINE config runs (k = 16 / 32): as below, but a round of 32 entries
// is the whole row of one vertex (or of two), so the masks of a vertex are ballots over its round and the vertex's
// first parent is its hook neighbour with the smallest id (redux.min); see the kernel.
//
// rg_classify_kernel (any k) — one warp per tile of 32 consecutive (Morton-adjacent) vertices, one lane
// per ENTRY (vertex, j) of the tile's 32 x k row block.  The 32 entries of a request are
// consecutive in memory (coalesced index / distance loads) and are the neighbours of one or
// two vertices, so the gathers of the neighbour's normal and k-th key fall on the few lines one
// neighbourhood occupies in Morton order.  RGA_B rounds of loads are in flight per lane.  Per
// entry: valid, one-way (u is not among the k nearest of v), smooth; per vertex three bit masks
// leave the kernel: one-way arcs (graph.cu's ow_mask), smooth MUTUAL arcs to smaller ids (to be
// hooked), smooth ONE-WAY arcs (to be listed).
// FUSED (iftwt_segment, one shared neighbourhood size): the pass also decides mutuality itself
// — one gather of the neighbour's k-th key per arc — and leaves ow_mask / rdeg / n_fwd for
// graph_build, so the rows and the neighbours are visited once instead of twice.
//
// rg_hook_kernel — one thread per vertex u (ECL-CC style): the arcs of the hook mask go into the
// union-find with u's representative cached in a register (no two lanes of a warp fight over
// the same root: a warp-per-vertex variant spent 19 ms in CAS retries); the arcs of the list
// mask are appended to (arc_u, arc_v) with one global atomic per block.
//
// Measured and rejected (round 2): the tile's 32 rows staged in shared memory by coalesced loads before the hooks
// (the scattered row[j] reads fetch 1.29 GB of DRAM sectors at 1 TB/s): 34 KB of shared memory per block, 62 %
// occupancy instead of 97 %, every row read whether it is needed or not — classify + hook 2.48 -> 3.29 ms.
//
// One thread per vertex doing both (rows and distances read at a 4k-byte stride, two gathers
// per arc each on its own line, finds and CAS in between) ran at 88 % L1 throughput with 45
// long-scoreboard stall cycles per issue: 3.9 ms at 10 M points, k = 32
// (profiles/r01_top_kernels_ncu_full.txt).
#define RGA_WARPS 8
__host__ __device__ inline int rga_flag_stride(int k) {  // bytes per vertex, (bytes / 4) odd
  int w = (k + 3) / 4;
  return 4 * (w | 1);
}
static size_t rga_smem_bytes(int k) { return (size_t)RGA_WARPS * (32 * 16 + 32 * 8 + 32 * rga_flag_stride(k)); }
// RGA_B = rounds of loads in flight per lane, MINB = blocks per SM the register allocation must allow
// ((4, 4) was round 1's: 63 registers, 47 % occupancy; variants selectable with IFTWT_RGA_VARIANT)
template <bool FUSED, int RGA_B, int MINB>
__global__ void __launch_bounds__(RGA_WARPS * 32, MINB)
rg_classify_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2, const u64* __restrict__ kth,
                   const float4* __restrict__ spts, u64* __restrict__ ow_mask, const float4* __restrict__ nrm, u32 n,
                   int k, float cos_thr, u64* __restrict__ hook_mask, u64* __restrict__ list_mask,
                   u32* __restrict__ rdeg, unsigned long long* __restrict__ n_fwd, u32* __restrict__ parent) {
  extern __shared__ __align__(16) unsigned char rga_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fs = rga_flag_stride(k);
  unsigned char* base = rga_smem + (size_t)warp * (32 * 16 + 32 * 8 + 32 * fs);
  float4* s_nu = (float4*)base;                            // normal of each tile vertex
  u64* s_aux = (u64*)(base + 32 * 16);                     // FUSED: original index; else: the one-way mask
  unsigned char* s_flags = base + 32 * 16 + 32 * 8;        // [32][fs] 1 valid | 2 one-way | 4 smooth | 8 v < u
  const u32 u0 = (blockIdx.x * RGA_WARPS + warp) * 32u;
  if (u0 >= n) return;  // warps are independent: no block-level barrier below
  const u32 u = u0 + lane;
  if (u < n) {
    s_nu[lane] = nrm[u];
    s_aux[lane] = FUSED ? (u64)__float_as_uint(spts[u].w) : ow_mask[u];
  }
  __syncwarp();
  {
    int p = lane / k, j = lane % k;
    const size_t row0 = (size_t)u0 * k;
    for (int r0 = 0; r0 < k; r0 += RGA_B) {  // 32 k entries, 32 per round
      int pp[RGA_B], jj[RGA_B];
      u32 v[RGA_B];
      float d2[RGA_B];
      u64 kk[RGA_B];
      float4 nv[RGA_B];
      bool ok[RGA_B];
#pragma unroll
      for (int b = 0; b < RGA_B; ++b) {
        pp[b] = p;
        jj[b] = j;
        const u32 e = (u32)lane + 32u * (u32)(r0 + b);
        ok[b] = r0 + b < k && u0 + (u32)p < n;
        v[b] = ok[b] ? nbr[row0 + e] : SID_NONE;
        d2[b] = (FUSED && ok[b]) ? nd2[row0 + e] : 0.f;
        j += 32;
        while (j >= k) {
          j -= k;
          ++p;
        }
      }
#pragma unroll
      for (int b = 0; b < RGA_B; ++b) {
        const bool live = v[b] != SID_NONE && v[b] != u0 + (u32)pp[b];
        kk[b] = (FUSED && live) ? kth[v[b]] : 0ull;
        nv[b] = live ? nrm[v[b]] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < RGA_B; ++b) {
        if (!ok[b]) continue;
        const u32 up = u0 + (u32)pp[b];
        unsigned flag = 0;
        if (v[b] != SID_NONE && v[b] != up) {
          flag = 1;
          bool oneway;
          if (FUSED) {
            const u64 key = ((u64)__float_as_uint(d2[b]) << 32) | s_aux[pp[b]];
            oneway = key > kk[b];
            if (oneway) atomicAdd(&rdeg[v[b]], 1u);
          } else {
            oneway = (s_aux[pp[b]] >> jj[b]) & 1ull;  // bit j: up is NOT among the k nearest of v (graph.cu)
          }
          if (oneway) flag |= 2;
          if (smooth_pass(s_nu[pp[b]], nv[b], cos_thr)) flag |= 4;
          if (v[b] < up) flag |= 8;
        }
        s_flags[pp[b] * fs + jj[b]] = (unsigned char)flag;
      }
    }
  }
  __syncwarp();
  u32 valid_cnt = 0;
  if (u < n) {
    const unsigned char* fl = s_flags + lane * fs;
    u64 ow = 0, hk = 0, ls = 0;
    for (int j = 0; j < k; ++j) {
      const unsigned f = fl[j];
      valid_cnt += f & 1u;
      if (f & 2u) ow |= 1ull << j;
      if ((f & 6u) == 6u) ls |= 1ull << j;   // smooth one-way
      if ((f & 14u) == 12u) hk |= 1ull << j;  // smooth mutual, v < u: the twin arc is mutual and smooth
    }                                         // too; one of the pair hooks
    if (FUSED) ow_mask[u] = ow;
    // ECL-CC's initialisation: u starts under its nearest hook neighbour (a smaller id, so the
    // links stay acyclic) — one plain store instead of two finds and a CAS, and the hook kernel
    // starts from a forest instead of n singletons
    u32 par = u;
    if (hk) {
      par = nbr[(size_t)u * k + (__ffsll((long long)hk) - 1)];
      hk &= hk - 1;
    }
    parent[u] = par;
    hook_mask[u] = hk;
    list_mask[u] = ls;
  }
  if (FUSED) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid_cnt += __shfl_xor_sync(0xffffffffu, valid_cnt, o);
    if (lane == 0 && valid_cnt) atomicAdd(n_fwd, (unsigned long long)valid_cnt);
  }
}

// rg_classify_kernel for a neighbourhood size that divides 32 (KK = 16, 32: every BASELINE config).  A round of
// 32 entries is then the whole row of one vertex (or of two): the three masks of a vertex are three ballots
// over the round, kept by the lane that owns the vertex — no flag bytes through shared memory, no 32-step loop
// per lane to assemble them, no division to find (vertex, j) of an entry.  The generic kernel spent 21 % of its
// instructions on that bookkeeping and 26 % on the flags (ncu source page, round 2): 112 -> ~60 per vertex.
template <bool FUSED, int RGA_B, int MINB, int KK>
__global__ void __launch_bounds__(RGA_WARPS * 32, MINB)
rg_classify_pow2_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2, const u64* __restrict__ kth,
                        const float4* __restrict__ spts, u64* __restrict__ ow_mask, const float4* __restrict__ nrm,
                        u32 n, float cos_thr, u64* __restrict__ hook_mask, u64* __restrict__ list_mask,
                        u32* __restrict__ rdeg, unsigned long long* __restrict__ n_fwd, u32* __restrict__ parent) {
  static_assert(KK == 16 || KK == 32, "one or two vertices per round");
  constexpr int VPR = 32 / KK;        // vertices per round
  constexpr u32 KM = KK == 32 ? 0xFFFFFFFFu : ((1u << (KK & 31)) - 1u);
  __shared__ float4 s_nu_all[RGA_WARPS][32];
  __shared__ u64 s_aux_all[RGA_WARPS][32];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* s_nu = s_nu_all[warp];   // normal of each tile vertex
  u64* s_aux = s_aux_all[warp];    // FUSED: original index; else: the one-way mask
  const u32 u0 = (blockIdx.x * RGA_WARPS + warp) * 32u;
  if (u0 >= n) return;  // warps are independent: no block-level barrier below
  const u32 u = u0 + lane;
  if (u < n) {
    s_nu[lane] = nrm[u];
    s_aux[lane] = FUSED ? (u64)__float_as_uint(spts[u].w) : ow_mask[u];
  }
  __syncwarp();
  const int sub = lane / KK, j = lane % KK;  // this lane's vertex within a round, its row entry
  const size_t row0 = (size_t)u0 * KK;
  u32 ow = 0, hk = 0, ls = 0, valid_cnt = 0, par_first = SID_NONE;
  for (int r0 = 0; r0 < KK; r0 += RGA_B) {  // KK rounds of 32 entries
    u32 v[RGA_B];
    float d2[RGA_B];
    u64 kk[RGA_B];
    float4 nv[RGA_B];
#pragma unroll
    for (int b = 0; b < RGA_B; ++b) {
      const int p = (r0 + b) * VPR + sub;
      const bool ok = r0 + b < KK && u0 + (u32)p < n;
      const u32 e = (u32)lane + 32u * (u32)(r0 + b);
      v[b] = ok ? nbr[row0 + e] : SID_NONE;
      d2[b] = (FUSED && ok) ? nd2[row0 + e] : 0.f;
      if (v[b] == u0 + (u32)p) v[b] = SID_NONE;  // the point itself
    }
#pragma unroll
    for (int b = 0; b < RGA_B; ++b) {
      const bool live = v[b] != SID_NONE;
      kk[b] = (FUSED && live) ? kth[v[b]] : 0ull;
      nv[b] = live ? nrm[v[b]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int b = 0; b < RGA_B; ++b) {
      const int p = (r0 + b) * VPR + sub;
      const bool live = v[b] != SID_NONE;
      bool oneway = false;
      if (live) {
        if (FUSED) {
          const u64 key = ((u64)__float_as_uint(d2[b]) << 32) | s_aux[p];
          oneway = key > kk[b];
          if (oneway) atomicAdd(&rdeg[v[b]], 1u);
        } else {
          oneway = (s_aux[p] >> j) & 1ull;  // bit j: the tile vertex is NOT among the k nearest of v (graph.cu)
        }
      }
      const bool smooth = live && smooth_pass(s_nu[p], nv[b], cos_thr);
      const unsigned b_valid = __ballot_sync(FULL, live);
      const unsigned b_ow = __ballot_sync(FULL, oneway);
      const unsigned b_ls = __ballot_sync(FULL, smooth && oneway);
      // smooth mutual, v < u: the twin arc is mutual and smooth too; one of the pair hooks
      const bool hookable = smooth && !oneway && v[b] < u0 + (u32)p;
      const unsigned b_hk = __ballot_sync(FULL, hookable);
      // the hook neighbour with the SMALLEST id becomes the vertex's first parent (below): it reaches much
      // further back along the Morton curve than the nearest one, so the initial forest is shallow
      const unsigned seg = KK == 32 ? FULL : (KM << (sub * KK));
      const u32 vmin = __reduce_min_sync(seg, hookable ? v[b] : SID_NONE);
      const unsigned b_first = __ballot_sync(FULL, hookable && v[b] == vmin);
      const u32 vmin_hi = VPR == 1 ? vmin : __shfl_sync(FULL, vmin, 16);  // the round's second vertex (KK == 16)
      const u32 vmin_lo = VPR == 1 ? vmin : __shfl_sync(FULL, vmin, 0);
      valid_cnt += (u32)__popc(b_valid);
      if (lane / VPR == r0 + b) {  // the lane that owns one of this round's vertices keeps its masks
        const int sh = (lane % VPR) * KK;
        ow = (b_ow >> sh) & KM;
        ls = (b_ls >> sh) & KM;
        hk = ((b_hk & ~b_first) >> sh) & KM;
        par_first = (lane % VPR) ? vmin_hi : vmin_lo;
      }
    }
  }
  if (u < n) {
    if (FUSED) ow_mask[u] = (u64)ow;
    // ECL-CC's initialisation: u starts under one of its hook neighbours (a smaller id, so the links stay
    // acyclic) — one plain store instead of two finds and a CAS
    parent[u] = par_first != SID_NONE ? par_first : u;
    hook_mask[u] = (u64)hk;
    list_mask[u] = (u64)ls;
  }
  if (FUSED && lane == 0 && valid_cnt) atomicAdd(n_fwd, (unsigned long long)valid_cnt);
}

#define HOOK_S 8
__global__ void __launch_bounds__(256)
rg_hook_kernel(const u32* __restrict__ nbr, const u64* __restrict__ hook_mask, const u64* __restrict__ list_mask,
               u32 n, int k, u32* __restrict__ parent, u32* __restrict__ counter, u32* __restrict__ arc_u,
               u32* __restrict__ arc_v, int hook_jumps) {
  __shared__ u32 s_warp[8];
  __shared__ u32 s_base;
  __shared__ u32 s_v[256 * HOOK_S];
  const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 mask = 0, hk = 0;
  if (u < n) {
    hk = hook_mask[u];
    mask = list_mask[u];
  }
  // ---- block-aggregated append of the one-way arcs.  FIRST: every thread does the same amount of work up to
  // the barriers, and none of the variable-length union-find work below is followed by one (with the append
  // last, 16 % of the kernel's stall samples sat on its __syncthreads: finished threads waiting for the
  // block's slowest chase) ----
  {
    u32 cnt = (u32)__popcll(mask);
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        u32 c = s_warp[w];
        s_warp[w] = tot;
        tot += c;
      }
      s_base = tot ? atomicAdd(counter, tot) : 0u;
    }
    __syncthreads();
    if (mask) {
      u32 p = s_base + s_warp[warp] + (incl - cnt);
      const u32* row = nbr + (size_t)u * k;
      while (mask) {
        int j = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        arc_u[p] = u;
        arc_v[p] = row[j];
        ++p;
      }
    }
  }
  if (hk) {
    const u32* row = nbr + (size_t)u * k;
    u32 ru = 0xFFFFFFFFu;
    // HOOK_S arcs at a time: first all the row entries, then all their parents — independent
    // loads the hardware overlaps — and only then the finds and hooks, which are chains
    u32* sv = s_v + threadIdx.x;  // slot q of this thread at sv[256 q]
    while (hk) {
      int m = 0;
      u64 h2 = hk;
#pragma unroll
      for (int q = 0; q < HOOK_S; ++q) {
        if (h2) {
          const int j = __ffsll((long long)h2) - 1;
          h2 &= h2 - 1;
          sv[256 * q] = row[j];
          m = q + 1;
        }
      }
      hk = h2;
#pragma unroll
      for (int q = 0; q < HOOK_S; ++q)
        if (q < m) sv[256 * q] = parent[sv[256 * q]];  // one step up: v itself is never needed again
      // ... and HOOK_JUMPS more steps for all of them at once (a root is its own parent: harmless there).  The
      // finds below walk one chain at a time; every step taken here is one round trip for the whole batch
      // instead of one per arc.
      u32 pu = ru == 0xFFFFFFFFu ? parent[u] : ru;
#pragma unroll 1
      for (int rep = 0; rep < hook_jumps; ++rep) {
#pragma unroll
        for (int q = 0; q < HOOK_S; ++q)
          if (q < m) sv[256 * q] = parent[sv[256 * q]];
        if (ru == 0xFFFFFFFFu) pu = parent[pu];
      }
      if (ru == 0xFFFFFFFFu) ru = uf_find(parent, pu);
      for (int q = 0; q < m; ++q) {
        u32 rv = sv[256 * q];
        if (rv == ru) continue;  // v hangs directly under u's representative
        rv = uf_find(parent, rv);
        while (ru != rv) {
          if (ru < rv) {
            u32 ret = atomicCAS(&parent[rv], rv, ru);
            if (ret == rv) break;
            rv = ret;
          } else {
            u32 ret = atomicCAS(&parent[ru], ru, rv);
            if (ret == ru) {
              ru = rv;
              break;
            }
            ru = ret;
          }
        }
      }
    }
    // u hangs (well) below its representative: point it there.  Safe as path halving is — a vertex that is not a
    // root never becomes one again and ru was an ancestor of u when it was read, so it still is; a ROOT must keep
    // its own word, another thread may be hooking it this instant.  The flatten kernel then finds short chains.
    if (ru != 0xFFFFFFFFu && ru != u) parent[u] = ru;
  }
}

// root[] is a separate array: writing the root back into parent[] would race with
// the path halving of concurrent finds, which may overwrite it with a non-root ancestor.
// Lanes that share a root (spatial neighbours mostly do) combine before the atomic.
__global__ void rg_flatten_kernel(u32* __restrict__ parent, const float4* __restrict__ nrm,
                                  const float4* __restrict__ spts, u32 n, u32* __restrict__ root,
                                  u64* __restrict__ L) {
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = v < n;
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  u32 r = uf_find(parent, v);
  root[v] = r;
  const unsigned grp = __match_any_sync(act, r);
  const u64 key = rg_key(nrm, spts, v);
  const u32 hi = __reduce_min_sync(grp, (u32)(key >> 32));
  const u32 lo = __reduce_min_sync(grp, (u32)(key >> 32) == hi ? (u32)key : 0xFFFFFFFFu);
  const u64 mn = ((u64)hi << 32) | (u64)lo;
  if ((int)(threadIdx.x & 31) == __ffs(grp) - 1 && mn < *(volatile u64*)&L[r])
    atomicMin((unsigned long long*)&L[r], (unsigned long long)mn);
}
// arcs to root space; only the arcs that join two DIFFERENT components matter to the sweeps
// (inside a wall or a floor nearly every one-way arc stays within its component), so those are
// appended to a compact list (out_count) and the sweeps read that list alone
__global__ void rg_arcs_to_roots_kernel(const u32* __restrict__ root, u32 m, const u32* __restrict__ arc_u,
                                        const u32* __restrict__ arc_v, u32* __restrict__ out_u,
                                        u32* __restrict__ out_v, u32* __restrict__ out_count) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 ru = 0, rv = 0;
  if (i < m) {
    ru = root[arc_u[i]];
    rv = root[arc_v[i]];
  }
  const bool keep = ru != rv;
  const unsigned mk = __ballot_sync(0xffffffffu, keep);
  if (!mk) return;
  const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
  u32 base = 0;
  if (lane == leader) base = atomicAdd(out_count, (u32)__popc(mk));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (keep) {
    const u32 pos = base + (u32)__popc(mk & ((1u << lane) - 1u));
    out_u[pos] = ru;
    out_v[pos] = rv;
  }
}
// m is read on the device: the grid is sized for the uncompacted list.  A launch relaxes every arc
// SWEEP_ITERS times: the labels are read past L1 (ld.cg — the atomics resolve in L2), so a thread sees what
// the others have lowered meanwhile and a chain of improvements advances several arcs per launch.  The
// fixpoint of a monotone min-propagation does not depend on the schedule; a launch in which nothing
// changed has seen every arc satisfied by labels that never moved, i.e. the fixpoint.  (One relaxation per
// launch needed 12 launches of ~15 us on cfg3.)
#define SWEEP_ITERS 4
__global__ void rg_sweep_kernel(const u32* __restrict__ arc_u, const u32* __restrict__ arc_v,
                                const u32* __restrict__ m_ptr, u64* L, u32* __restrict__ changed) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *m_ptr) return;
  const u32 ru = arc_u[i], rv = arc_v[i];
  bool any = false;
#pragma unroll 1
  for (int it = 0; it < SWEEP_ITERS; ++it) {
    const u64 a = __ldcg((const unsigned long long*)&L[ru]);
    if (a < __ldcg((const unsigned long long*)&L[rv])) {
      atomicMin((unsigned long long*)&L[rv], (unsigned long long)a);
      any = true;
    }
  }
  if (any) *changed = 1u;
}
// per point: segment = the vertex whose key is L[root] (the segment's seed point); sizes are
// counted at the seed vertex.  A wall or a floor is ONE segment of 10^5 points, so one counter took an
// atomic from nearly every warp of the cloud (312 k serialised atomics on a handful of addresses: 0.28 ms):
// a block of 1024 consecutive (Morton-adjacent) vertices mostly lies in one segment, so the count of the block's
// first vertex's segment is gathered in shared memory and leaves as ONE atomic; the other segments of the
// block are warp-aggregated as before.
#define SIZES_THREADS 1024
__global__ void __launch_bounds__(SIZES_THREADS)
rg_sizes_kernel(const u32* __restrict__ root, const u64* __restrict__ L, const u32* __restrict__ inv, u32 n,
                u32* __restrict__ seg_of, u32* __restrict__ size_at_seed) {
  __shared__ u32 s_pivot, s_count;
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = v < n;
  u32 s = 0xFFFFFFFFu;
  if (valid) {
    s = inv[(u32)L[root[v]]];  // low half of the key = original index of the seed point
    seg_of[v] = s;
  }
  if (threadIdx.x == 0) {
    s_pivot = s;  // the block's first vertex is always valid
    s_count = 0u;
  }
  __syncthreads();
  const u32 pivot = s_pivot;
  const unsigned on_pivot = __ballot_sync(0xffffffffu, valid && s == pivot);
  const int lane = threadIdx.x & 31;
  if (lane == 0 && on_pivot) atomicAdd(&s_count, (u32)__popc(on_pivot));
  const bool other = valid && s != pivot;
  const unsigned act = __ballot_sync(0xffffffffu, other);
  if (other) {
    const unsigned grp = __match_any_sync(act, s);
    if (lane == __ffs(grp) - 1) atomicAdd(&size_at_seed[s], (u32)__popc(grp));
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_count) atomicAdd(&size_at_seed[pivot], s_count);
}
// keep[s] for every seed vertex s; the kept seeds' (key, vertex) pairs are appended in any
// order — the sort that follows puts them in PCL's emission order (ascending key).  The segment count
// leaves as one atomic per block (one per warp: 0.8 M atomics on one address, 0.15 ms).
__global__ void rg_keep_flags_kernel(const u32* __restrict__ size_at_seed, const float4* __restrict__ nrm,
                                     const float4* __restrict__ spts, u32 n, u32 min_sz, u32 max_sz,
                                     u32* __restrict__ keep, u32* __restrict__ nseg, u32* __restrict__ n_kept,
                                     u64* __restrict__ kept_key, u32* __restrict__ kept_vertex) {
  __shared__ u32 s_nseg;
  if (threadIdx.x == 0) s_nseg = 0u;
  __syncthreads();
  u32 v = blockIdx.x * blockDim.x + threadIdx.x;
  u32 s = v < n ? size_at_seed[v] : 0u;
  const bool any = s > 0;
  const bool kp = s >= min_sz && s <= max_sz && any;
  if (v < n) keep[v] = kp ? 1u : 0u;
  const unsigned m = __ballot_sync(0xffffffffu, any);
  const unsigned mk = __ballot_sync(0xffffffffu, kp);
  const int lane = threadIdx.x & 31;
  if (lane == 0 && m) atomicAdd(&s_nseg, (u32)__popc(m));
  if (mk) {
    u32 base = 0;
    if (lane == __ffs(mk) - 1) base = atomicAdd(n_kept, (u32)__popc(mk));
    base = __shfl_sync(0xffffffffu, base, __ffs(mk) - 1);
    if (kp) {
      const u32 pos = base + (u32)__popc(mk & ((1u << lane) - 1u));
      kept_key[pos] = rg_key(nrm, spts, v);
      kept_vertex[pos] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_nseg) atomicAdd(nseg, s_nseg);
}
// kept cluster c = position of its seed in ascending key order
__global__ void rg_kept_ids_kernel(const u32* __restrict__ kept_vertex_sorted, const u32* __restrict__ size_at_seed,
                                   u32 n_kept, u32* __restrict__ kept_id, u32* __restrict__ ksize) {
  u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_kept) return;
  const u32 v = kept_vertex_sorted[c];
  kept_id[v] = c;
  ksize[c] = size_at_seed[v];
}
// members in ORIGINAL order: key = kept cluster (n_kept for dropped points), value =
// original index; a stable sort on the key groups clusters with ascending indices
__global__ void rg_member_keys_kernel(const u32* __restrict__ seg_of, const u32* __restrict__ keep,
                                      const u32* __restrict__ kept_id, const u32* __restrict__ inv, u32 n,
                                      u32 n_kept, u32* __restrict__ keys, u32* __restrict__ vals,
                                      int* __restrict__ point_label) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 s = seg_of[inv[i]];
  bool kp = keep[s] != 0;
  u32 c = kp ? kept_id[s] : n_kept;
  keys[i] = c;
  vals[i] = i;
  point_label[i] = kp ? (int)c : -1;
}
// One warp per kept cluster: pcl::compute3DCentroid, FP32, ascending point index.
// The 32 gathers of a batch are issued in parallel and parked in shared memory; the additions
// stay strictly sequential (every lane carries the same running sums, read by broadcast), so
// the result is the literal left-to-right FP32 sum.  (Broadcasting through shuffles put a
// shuffle on the dependency chain of every addition: 0.29 ms for 190 clusters.)
#define CEN_WARPS 4
#define CEN_DEPTH 4  // batches of 32 gathers in flight per warp: one ahead left every batch waiting ~0.4 us for
                     // its gather (313 batches in a 10,000-point cluster: 0.13 ms for cfg3's 190 clusters)
__global__ void __launch_bounds__(CEN_WARPS * 32)
rg_centroid_kernel(const u32* __restrict__ members, const u32* __restrict__ kstart,
                   const u32* __restrict__ ksize, u32 n_kept,
                   const float4* __restrict__ pts_in, float4* __restrict__ centroid) {
  __shared__ float4 s_p[CEN_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 c = blockIdx.x * CEN_WARPS + warp;
  if (c >= n_kept) return;
  const u32 b = kstart[c], cnt = ksize[c];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  float4 nxt[CEN_DEPTH];
#pragma unroll
  for (int d = 0; d < CEN_DEPTH; ++d) {
    nxt[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (32u * d + lane < cnt) nxt[d] = pts_in[members[b + 32u * d + lane]];
  }
  for (u32 base = 0; base < cnt; base += 32u * CEN_DEPTH) {
#pragma unroll
    for (int d = 0; d < CEN_DEPTH; ++d) {
      const u32 bd = base + 32u * d;
      if (bd >= cnt) break;
      __syncwarp();  // the previous batch has been read by every lane
      s_p[warp][lane] = nxt[d];
      __syncwarp();
      const u32 ahead = bd + 32u * CEN_DEPTH + lane;
      if (ahead < cnt) nxt[d] = pts_in[members[b + ahead]];  // CEN_DEPTH batches ahead
      const int m = (int)min(32u, cnt - bd);
      const float4* q = s_p[warp];
      if (m == 32) {  // whole batch: eight broadcast reads ahead of the 24 dependent additions they feed
#pragma unroll
        for (int t0 = 0; t0 < 32; t0 += 8) {
          float4 r[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) r[t] = q[t0 + t];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            sx += r[t].x;
            sy += r[t].y;
            sz += r[t].z;
          }
        }
      } else {
        for (int t = 0; t < m; ++t) {
          const float4 p = q[t];
          sx += p.x;
          sy += p.y;
          sz += p.z;
        }
      }
    }
  }
  if (lane == 0) {
    float d = (float)cnt;
    centroid[c] = make_float4(sx / d, sy / d, sz / d, 0.f);
  }
}
__global__ void rg_seed_ids_kernel(const u32* __restrict__ sid_sorted, const u32* __restrict__ perm, u32 m,
                                   u32* __restrict__ seeds) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  u32 s = sid_sorted[i];
  seeds[i] = s == SID_NONE ? SID_NONE : perm[s];
}

}  // namespace

static int hook_jumps() {
  const char* e = getenv("IFTWT_HOOK_JUMPS");
  return e ? atoi(e) : 2;
}

// iftwt_segment fast path: graph mutuality + region-growing arcs in one pass.  Needs the k-NN
// lists and the normals; leaves ow_mask / rdeg / n_fwd for graph_build and the union-find +
// one-way arc list for seeds_run.
int fused_graph_rg_run(iftwt_ctx* c, const iftwt_rg_params* p) {
  if (c->knn_k == 0 || c->nrm_k != c->knn_k || p->number_of_neighbours != c->knn_k)
    return ctx_fail(c, IFTWT_ERR_STATE, "fused graph/region-growing pass needs one shared neighbourhood size");
  StageTimer st(c, IFTWT_STAGE_GRAPH_RG);
  const u32 n = (u32)c->n;
  const int k = c->knn_k;
  const size_t ne = (size_t)n * k;
  const float cos_thr = cosf(p->smoothness_threshold);
  ENSURE(c, c->ow_mask, (size_t)n * 8);
  ENSURE(c, c->g_deg, ((size_t)n + 1) * 4);
  ENSURE(c, c->rg_parent, (size_t)n * 4);
  ENSURE(c, c->rg_aux2, (ne + 1) * 8);
  u32* sc = (u32*)c->scalars.as<int>();
  u32* d_cnt = sc + 56;
  unsigned long long* d_nfwd = (unsigned long long*)(c->scalars.as<int>() + 52);
  CU_TRY(c, cudaMemsetAsync(d_cnt, 0, 16, c->stream));
  CU_TRY(c, cudaMemsetAsync(d_nfwd, 0, 8, c->stream));
  CU_TRY(c, cudaMemsetAsync(c->g_deg.p, 0, ((size_t)n + 1) * 4, c->stream));
  u32* arc_u = c->rg_aux2.as<u32>();
  ENSURE(c, c->rg_masks, (size_t)n * 16);
  u64* hook_mask = c->rg_masks.as<u64>();
  u64* list_mask = hook_mask + n;
  const int variant = getenv("IFTWT_RGA_VARIANT") ? atoi(getenv("IFTWT_RGA_VARIANT")) : 0;
#define RGA_LAUNCH(B, MINB)                                                                                          \
  {                                                                                                                  \
    CU_TRY(c, cudaFuncSetAttribute(rg_classify_kernel<true, B, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                   (int)rga_smem_bytes(k)));                                                         \
    LAUNCH(c, (rg_classify_kernel<true, B, MINB>), div_up(n, RGA_WARPS * 32), RGA_WARPS * 32, rga_smem_bytes(k),     \
           c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>(), c->spts.as<float4>(), c->ow_mask.as<u64>(),       \
           c->nrm.as<float4>(), n, k, cos_thr, hook_mask, list_mask, c->g_deg.as<u32>(), d_nfwd,                     \
           c->rg_parent.as<u32>());                                                                                  \
  }
#define RGA_LAUNCH_P2(B, MINB, KK)                                                                                 \
  LAUNCH(c, (rg_classify_pow2_kernel<true, B, MINB, KK>), div_up(n, RGA_WARPS * 32), RGA_WARPS * 32, 0,            \
         c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>(), c->spts.as<float4>(), c->ow_mask.as<u64>(),       \
         c->nrm.as<float4>(), n, cos_thr, hook_mask, list_mask, c->g_deg.as<u32>(), d_nfwd, c->rg_parent.as<u32>())
  const bool generic = getenv("IFTWT_RGA_GENERIC") && atoi(getenv("IFTWT_RGA_GENERIC")) != 0;
  if (!generic && k == 32) {
    // cfg3 on B200, classify + hook: (4 rounds in flight, 4 blocks per SM: 64 registers) 2.36 ms, (4, 5: 48
    // registers, a few spilled words since the smallest-id parent) 2.45; the generic kernel 2.89
    if (variant == 1) RGA_LAUNCH_P2(4, 5, 32);
    else if (variant == 3) RGA_LAUNCH_P2(2, 6, 32);
    else RGA_LAUNCH_P2(4, 4, 32);
  } else if (!generic && k == 16) {
    // cfg3 on B200, classify + hook: (4 rounds in flight, 4 blocks per SM: 64 registers) 2.36 ms, (4, 5: 48
    // registers, a few spilled words since the smallest-id parent) 2.45; the generic kernel 2.89
    if (variant == 1) RGA_LAUNCH_P2(4, 5, 16);
    else if (variant == 3) RGA_LAUNCH_P2(2, 6, 16);
    else RGA_LAUNCH_P2(4, 4, 16);
  } else
  switch (variant) {
    case 1: RGA_LAUNCH(4, 5); break;
    case 2: RGA_LAUNCH(4, 6); break;
    case 3: RGA_LAUNCH(2, 6); break;
    case 4: RGA_LAUNCH(2, 8); break;
    case 5: RGA_LAUNCH(8, 4); break;
    default: RGA_LAUNCH(4, 4); break;
  }
#undef RGA_LAUNCH
#undef RGA_LAUNCH_P2
  LAUNCH(c, rg_hook_kernel, div_up(n, 256), 256, 0, c->nbr.as<u32>(), hook_mask, list_mask, n, k,
         c->rg_parent.as<u32>(), d_cnt, arc_u, arc_u + (ne + 1), hook_jumps());
  CHECK_LAUNCH(c);
  c->mask_k = k;
  c->rdeg_k = k;
  c->rg_pre_valid = true;
  c->rg_pre_k = k;
  c->rg_pre_cos = cos_thr;
  return IFTWT_OK;
}

int seeds_run(iftwt_ctx* c, const iftwt_rg_params* p) {
  if (c->nrm_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "normals are not computed (call iftwt_normals first)");
  if (p->number_of_neighbours != c->knn_k) {
    // region growing wants its own neighbourhood size: recompute the lists on the same grid
    TRY(knn_run(c, p->number_of_neighbours));
  }
  StageTimer st(c, IFTWT_STAGE_SEEDS);
  const u32 n = (u32)c->n;
  const int k = c->knn_k;
  const size_t ne = (size_t)n * k;
  const float4* nrm = c->nrm.as<float4>();
  const float4* spts = c->spts.as<float4>();
  u32* sc = (u32*)c->scalars.as<int>();
  u32* hs = (u32*)c->h_scalars;

  // the exact parallel form needs the curvature gate to be inactive
  // (curvature = |lambda0 / trace| <= 1/3 for a PSD covariance; NaN never compares greater)
  if (!(p->curvature_threshold >= 0.34f))
    return ctx_fail(c, IFTWT_ERR_UNSUPPORTED,
                    "curvature_threshold %.3f can fire; the GPU restatement of RegionGrowing is exact only "
                    "when it cannot (threshold >= 0.34, reference uses 1.0)", p->curvature_threshold);
  const float cos_thr = cosf(p->smoothness_threshold);

  // 1. the (curvature, index) order lives in 64-bit keys computed on the fly (rg_key)
  ENSURE(c, c->keys, (size_t)n * 8);
  ENSURE(c, c->keys_alt, (size_t)n * 8);
  ENSURE(c, c->vals_alt, (size_t)n * 4);
  ENSURE(c, c->rg_rank, (size_t)n * 4);
  ENSURE(c, c->rg_parent, (size_t)n * 4);
  ENSURE(c, c->rg_label, (size_t)n * 8);
  ENSURE(c, c->rg_aux, (size_t)n * 4 * 4);
  u32* parent = c->rg_parent.as<u32>();
  u64* L = c->rg_label.as<u64>();
  u32* seg_of = c->rg_aux.as<u32>();
  u32* size_at_seed = seg_of + n;
  u32* keep = size_at_seed + n;
  u32* kept_id = keep + n;
  // 2. mutual smooth arcs -> union-find; one-way smooth arcs -> list (worst-case capacity).
  //    Already done when iftwt_segment ran the fused graph + region-growing pass.
  u32* d_cnt = sc + 56;
  u32* arc_u = nullptr;
  u32* arc_v = nullptr;
  if (c->rg_pre_valid && c->rg_pre_k == k && c->rg_pre_cos == cos_thr) {
    arc_u = c->rg_aux2.as<u32>();
    arc_v = arc_u + (ne + 1);
  } else {
    CU_TRY(c, cudaMemsetAsync(d_cnt, 0, 16, c->stream));
    ENSURE(c, c->rg_aux2, (ne + 1) * 8);
    arc_u = c->rg_aux2.as<u32>();
    arc_v = arc_u + (ne + 1);
    if (c->mask_k != c->knn_k) TRY(knn_mutual_run(c, false));
    ENSURE(c, c->rg_masks, (size_t)n * 16);
    u64* hook_mask = c->rg_masks.as<u64>();
    u64* list_mask = hook_mask + n;
    const bool generic = getenv("IFTWT_RGA_GENERIC") && atoi(getenv("IFTWT_RGA_GENERIC")) != 0;
#define RGA_LAUNCH_P2N(KK)                                                                                         \
  LAUNCH(c, (rg_classify_pow2_kernel<false, 4, 4, KK>), div_up(n, RGA_WARPS * 32), RGA_WARPS * 32, 0,             \
         c->nbr.as<u32>(), (const float*)nullptr, (const u64*)nullptr, (const float4*)nullptr,                    \
         c->ow_mask.as<u64>(), nrm, n, cos_thr, hook_mask, list_mask, (u32*)nullptr, (unsigned long long*)nullptr, \
         parent)
    if (!generic && k == 32) {
      RGA_LAUNCH_P2N(32);
    } else if (!generic && k == 16) {
      RGA_LAUNCH_P2N(16);
    } else {
    CU_TRY(c, cudaFuncSetAttribute(rg_classify_kernel<false, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)rga_smem_bytes(k)));
    LAUNCH(c, (rg_classify_kernel<false, 4, 4>), div_up(n, RGA_WARPS * 32), RGA_WARPS * 32, rga_smem_bytes(k),
           c->nbr.as<u32>(), (const float*)nullptr, (const u64*)nullptr, (const float4*)nullptr, c->ow_mask.as<u64>(),
           nrm, n, k, cos_thr, hook_mask, list_mask, (u32*)nullptr, (unsigned long long*)nullptr, parent);
    }
#undef RGA_LAUNCH_P2N
    LAUNCH(c, rg_hook_kernel, div_up(n, 256), 256, 0, c->nbr.as<u32>(), hook_mask, list_mask, n, k, parent, d_cnt,
           arc_u, arc_v, hook_jumps());
    CHECK_LAUNCH(c);
  }
  c->rg_pre_valid = false;  // the union-find is consumed (flattened) below
  TRY(host_fetch(c, hs + 56, d_cnt, 4));
  // 3. component min rank, then label-min over one-way arcs
  CU_TRY(c, cudaMemsetAsync(L, 0xFF, (size_t)n * 8, c->stream));
  u32* root = c->vals_alt.as<u32>();  // free again once the ranks are scattered
  LAUNCH(c, rg_flatten_kernel, div_up(n, 256), 256, 0, parent, nrm, spts, n, root, L);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const u32 n_oneway = hs[56];
  u64 sweeps = 0;
  if (n_oneway) {
    // the compact list goes behind the full one when it fits (it does unless more than half of
    // all arcs are one-way and smooth), else to the sort scratch
    u32* d_m = sc + 57;
    u32 *cu, *cv;
    if ((size_t)n_oneway * 2 <= ne + 1) {
      cu = arc_u + n_oneway;
      cv = arc_v + n_oneway;
    } else {
      ENSURE(c, c->keys, (size_t)n_oneway * 4);
      ENSURE(c, c->keys_alt, (size_t)n_oneway * 4);
      cu = c->keys.as<u32>();
      cv = c->keys_alt.as<u32>();
    }
    CU_TRY(c, cudaMemsetAsync(d_m, 0, 4, c->stream));
    LAUNCH(c, rg_arcs_to_roots_kernel, div_up(n_oneway, 256), 256, 0, root, n_oneway, arc_u, arc_v, cu, cv, d_m);
    TRY(host_fetch(c, hs + 57, d_m, 4));
    u32* d_changed = sc + 60;
    u32 grid_arcs = n_oneway;  // until the compact count has come back with the first sync
    while (true) {
      CU_TRY(c, cudaMemsetAsync(d_changed, 0, 4, c->stream));
      for (int b = 0; b < 2; ++b) {
        LAUNCH(c, rg_sweep_kernel, div_up(grid_arcs ? grid_arcs : 1, 256), 256, 0, cu, cv, d_m, L, d_changed);
        sweeps += SWEEP_ITERS;
      }
      CHECK_LAUNCH(c);
      TRY(host_fetch(c, hs + 60, d_changed, 4));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
      grid_arcs = hs[57];
      if (hs[60] == 0) break;
      if (sweeps > 400000) return ctx_fail(c, IFTWT_ERR_CUDA, "region growing did not converge");
    }
  }
  c->stats.rg_sweeps = sweeps;
  // 4. sizes, filter, emission order (ascending key of the seed point = PCL's)
  CU_TRY(c, cudaMemsetAsync(size_at_seed, 0, (size_t)n * 4, c->stream));
  CU_TRY(c, cudaMemsetAsync(sc + 61, 0, 8, c->stream));  // [61] segments, [62] kept
  LAUNCH(c, rg_sizes_kernel, div_up(n, SIZES_THREADS), SIZES_THREADS, 0, root, L, c->inv.as<u32>(), n, seg_of,
         size_at_seed);
  u64* kept_key = c->keys.as<u64>();
  u64* kept_key_sorted = c->keys_alt.as<u64>();
  u32* kept_vertex = c->rg_rank.as<u32>();
  LAUNCH(c, rg_keep_flags_kernel, div_up(n, 256), 256, 0, size_at_seed, nrm, spts, n,
         (u32)max(p->min_cluster_size, 0), (u32)max(p->max_cluster_size, 0), keep, sc + 61, sc + 62, kept_key,
         kept_vertex);
  CHECK_LAUNCH(c);
  TRY(host_fetch(c, hs + 61, sc + 61, 8));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const u32 n_kept = hs[62];
  c->stats.rg_segments = hs[61];
  ENSURE(c, c->out_b, ((size_t)n_kept + 1) * 4 * 3 + (size_t)n_kept * 16 + 64);
  u32* ksize = c->out_b.as<u32>();
  u32* kstart = ksize + (n_kept + 1);
  u32* sid = kstart + (n_kept + 1);
  float4* centroid = (float4*)(((uintptr_t)(sid + (n_kept + 1)) + 15) & ~(uintptr_t)15);
  if (n_kept) {
    u32* kept_vertex_sorted = c->vals_alt.as<u32>();  // root[] is no longer needed
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, kept_key, kept_key_sorted, kept_vertex, kept_vertex_sorted,
                                              (int)n_kept, 0, 64, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, kept_key, kept_key_sorted, kept_vertex,
                                              kept_vertex_sorted, (int)n_kept, 0, 64, c->stream));
    c->stats.kernel_launches += 2;
    LAUNCH(c, rg_kept_ids_kernel, div_up(n_kept, 256), 256, 0, kept_vertex_sorted, size_at_seed, n_kept, kept_id,
           ksize);
    CHECK_LAUNCH(c);
  }
  // point labels (original order) + member lists
  ENSURE(c, c->out_a, (size_t)n * 4);
  int* point_label = c->out_a.as<int>();
  u32* mkeys = c->keys.as<u32>();
  u32* mkeys_sorted = c->keys_alt.as<u32>();
  u32* mvals = c->keys.as<u32>() + n;          // keys / keys_alt hold 8 bytes per point
  u32* members = c->keys_alt.as<u32>() + n;
  LAUNCH(c, rg_member_keys_kernel, div_up(n, 256), 256, 0, seg_of, keep, kept_id, c->inv.as<u32>(), n, n_kept,
         mkeys, mvals, point_label);
  CHECK_LAUNCH(c);
  c->n_seeds = n_kept;
  ENSURE(c, c->seeds, ((size_t)n_kept + 1) * 4);
  if (n_kept) {
    int kbits = 1;
    while ((1u << kbits) < n_kept + 1) ++kbits;
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, mkeys, mkeys_sorted, mvals, members, (int)n, 0, kbits,
                                              c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, mkeys, mkeys_sorted, mvals, members, (int)n, 0, kbits,
                                              c->stream));
    c->stats.kernel_launches += 1 + (kbits + 7) / 8;
    // kept cluster starts
    size_t tb2 = 0;
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(nullptr, tb2, ksize, kstart, (int)n_kept, c->stream));
    ENSURE(c, c->tmp, tb2);
    CU_TRY(c, cub::DeviceScan::ExclusiveSum(c->tmp.p, tb2, ksize, kstart, (int)n_kept, c->stream));
    c->stats.kernel_launches += 2;
    LAUNCH(c, rg_centroid_kernel, div_up(n_kept, CEN_WARPS), CEN_WARPS * 32, 0, members, kstart, ksize, n_kept,
           c->pts_in.as<float4>(), centroid);
    CHECK_LAUNCH(c);
    // 5. snap to the nearest input point
    TRY(knn_query_run(c, centroid, n_kept, 1, sid, nullptr));
    LAUNCH(c, rg_seed_ids_kernel, div_up(n_kept, 256), 256, 0, sid, c->perm.as<u32>(), n_kept,
           c->seeds.as<u32>());
    CHECK_LAUNCH(c);
  }
  c->seeds_valid = true;
  c->ift_valid = false;
  c->stats.n_seeds = n_kept;
  return IFTWT_OK;
}
