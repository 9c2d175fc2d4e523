// shard.cu — Morton-range sharded k-NN / normals across the GPUs of one node (BASELINE config 4,
// SURVEY 8(e)): partition, owner routing, halo selection and the row export are CUDA kernels, the
// exchange is NCCL (grouped ncclSend / ncclRecv over NVLink), all behind the C ABI of iftwt.h.
//
// The reference has nothing to match here: the loops being sharded are the per-point k-NN /
// normal-estimation loops of graph.hpp:245 and main.cpp:102-109; the spec is the north star.
//
// One exchange step before the compute, none during it:
//   A. every rank: bounding box of its slice (+ an upper bound of the k-th neighbour distance when the
//      caller gave no halo radius)                                   -> ncclAllGather of one 64-byte record
//   B. coarse grid of side s = halo radius over the global box; histogram of the top bits of the
//      coarse Morton key                                             -> ncclAllReduce (<= 2^18 bins)
//      -> contiguous, balanced Morton ranges, one per rank (owner_of_bin, computed identically everywhere)
//   C. ONE pass decides, per point, its owner and every other rank that owns one of the 26 coarse cells
//      around its own (the halo destinations), and a STABLE multi-split packs the send buffer by
//      destination.  Stable matters: slices are contiguous ascending global-id ranges, so what a rank
//      receives — source after source — is already in ascending global id order, owned and halo points
//      interleaved.  Local index order = global id order, which is what makes distance ties break as on
//      one GPU, without any sort.                                    -> ncclAllGather of the send counts
//   D. grouped ncclSend / ncclRecv of the points and their ids       (the only bulk transfer)
//   E. the single-GPU kernels (grid, k-NN, normals) on owned + halo points
//   F. rows of the owned points translated to global ids; exactness check: everything within s of an
//      owned point is present, so a row is the global k-NN iff its k-th distance is <= s
//                                                                    -> ncclAllReduce of the violation count;
//      on a violation the exchange is redone with 2 s.
//
// The IFT does not shard (global ordered propagation): iftwt_sharded_segment gathers rows and normals
// to one rank (grouped ncclSend / ncclRecv), installs them in that rank's ctx and runs graph / seeds /
// IFT there.
//
// Transport.  One process per GPU (iftwt_comm_create: ncclCommInitRank) or one process driving several
// GPUs (iftwt_comm_create_all: ncclCommInitAll).  When the device list of iftwt_comm_create_all names
// the same device more than once NCCL cannot be used (it refuses duplicate devices): the shards then
// time-share that GPU and the collectives are device-to-device copies.  That mode exists so the whole
// path — every kernel in this file — runs in the single-GPU test-suite.
//
// libnccl is opened with dlopen on first use: libiftwt.so itself does not link against it.
#include <cub/cub.cuh>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <new>
#include <string>
#include <vector>

#include "common.cuh"

#define SHARD_MAX_RANKS 8
#define GID_OWNED 0x80000000u
#define GID_MASK 0x7FFFFFFFu
#define SID_NONE 0xFFFFFFFFu
#define PACK_PPT 4                      // consecutive points per thread in the route / pack kernels
#define PACK_TILE (256 * PACK_PPT)      // points per block
#define SHARD_HIST_BITS 18
#define SHARD_SAMPLE 4096               // queries of the halo-radius estimate

namespace {

// ---- NCCL through dlopen ---------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};
NcclApi g_nccl;
bool nccl_load() {
  if (g_nccl.h) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    g_nccl.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?");
    return false;
  }
#define NCCL_SYM(field, name)                                  \
  *(void**)(&g_nccl.field) = dlsym(h, name);                   \
  if (!g_nccl.field) {                                         \
    g_nccl.err = std::string("libnccl lacks ") + name;         \
    return false;                                              \
  }
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  NCCL_SYM(CommInitRank, "ncclCommInitRank");
  NCCL_SYM(CommInitAll, "ncclCommInitAll");
  NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(AllGather, "ncclAllGather");
  NCCL_SYM(AllReduce, "ncclAllReduce");
  NCCL_SYM(Send, "ncclSend");
  NCCL_SYM(Recv, "ncclRecv");
  NCCL_SYM(GroupStart, "ncclGroupStart");
  NCCL_SYM(GroupEnd, "ncclGroupEnd");
  NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
  g_nccl.h = h;
  return true;
}

// ---- the coarse grid ---------------------------------------------------------------------------
struct CoarseGrid {
  double ox, oy, oz, inv_s;
  int dx, dy, dz;
  int shift;   // histogram bin = coarse Morton key >> shift
  int hbits;   // cells sharing (ix >> hbits, iy >> hbits, iz >> hbits) share their bin
  u32 nbins;
};
__device__ __forceinline__ void coarse_cell(const CoarseGrid& g, const float4 p, int& ix, int& iy, int& iz) {
  ix = min(g.dx - 1, max(0, (int)floor(((double)p.x - g.ox) * g.inv_s)));
  iy = min(g.dy - 1, max(0, (int)floor(((double)p.y - g.oy) * g.inv_s)));
  iz = min(g.dz - 1, max(0, (int)floor(((double)p.z - g.oz) * g.inv_s)));
}
__device__ __forceinline__ int f2ord_i(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
inline float ord2f_h(int i) {
  int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
  float f;
  memcpy(&f, &j, 4);
  return f;
}

// the 64-byte record every rank contributes to phase A
struct ShardMeta {
  int lo[3], hi[3];      // bounding box of the slice, order-preserving int image of the floats
  u32 bad;               // non-finite points
  u32 n;                 // points in the slice
  u32 gid0;              // global id of its first point
  u32 kth_bits;          // float bits: upper bound estimate of the k-th neighbour distance^2 (0 = none)
  u32 epoch;             // bumped whenever one of the rank's peer-writable buffers is (re)allocated
  u32 pad[5];
};
static_assert(sizeof(ShardMeta) == 64, "ShardMeta is one 64-byte record");

__global__ void shard_meta_init_kernel(ShardMeta* m, u32 n, u32 gid0, u32 epoch) {
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      m->lo[a] = 0x7FFFFFFF;
      m->hi[a] = (int)0x80000000;
    }
    m->bad = 0;
    m->n = n;
    m->gid0 = gid0;
    m->kth_bits = 0;
    m->epoch = epoch;
    for (int a = 0; a < 5; ++a) m->pad[a] = 0;
  }
}
__global__ void shard_bbox_kernel(const float4* __restrict__ pts, u32 n, ShardMeta* __restrict__ m) {
  int mn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
  int mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  u32 bad = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      ++bad;
      continue;
    }
    const int a = f2ord_i(p.x), b = f2ord_i(p.y), c = f2ord_i(p.z);
    mn[0] = min(mn[0], a); mx[0] = max(mx[0], a);
    mn[1] = min(mn[1], b); mx[1] = max(mx[1], b);
    mn[2] = min(mn[2], c); mx[2] = max(mx[2], c);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&m->lo[a], mn[a]);
      atomicMax(&m->hi[a], mx[a]);
    }
    if (bad) atomicAdd(&m->bad, bad);
  }
}
// halo-radius estimate: SHARD_SAMPLE evenly spaced points of the slice are queried against the slice
// itself; the k-th neighbour inside a SUBSET of the cloud is never nearer than the k-th neighbour in
// the cloud, so the largest of these distances bounds the sampled points' true k-th distance from above
__global__ void shard_sample_kernel(const float4* __restrict__ pts, u32 n, u32 m, float4* __restrict__ q) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  q[i] = pts[(size_t)i * (n / m)];
}
__global__ void shard_kth_max_kernel(const float* __restrict__ d2, u32 m, int k, ShardMeta* __restrict__ meta) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (i < m) {
    // the last finite entry of the row (rows are ascending, +inf padded when the slice holds < k points)
    for (int j = k - 1; j >= 0; --j) {
      const float d = d2[(size_t)i * k + j];
      if (isfinite(d)) {
        v = d;
        break;
      }
    }
  }
  const u32 b = __reduce_max_sync(0xffffffffu, __float_as_uint(v));  // non-negative floats order as uints
  if ((threadIdx.x & 31) == 0 && b) atomicMax(&meta->kth_bits, b);
}

// A slice that is a spatial tile falls into a few dozen bins: one global atomic per warp and bin serialised on
// those few addresses (2 ms for 12.5 M points).  Each block therefore counts its 1,024 points in a small
// shared-memory table first (open addressing on the bin) and flushes one atomic per distinct bin.
#define HIST_SLOTS 2048
__global__ void __launch_bounds__(256)
shard_hist_kernel(const float4* __restrict__ pts, u32 n, CoarseGrid g, u32* __restrict__ hist) {
  __shared__ u32 s_key[HIST_SLOTS];
  __shared__ u32 s_cnt[HIST_SLOTS];
  for (int t = threadIdx.x; t < HIST_SLOTS; t += 256) {
    s_key[t] = 0xFFFFFFFFu;
    s_cnt[t] = 0;
  }
  __syncthreads();
  const u32 i0 = blockIdx.x * PACK_TILE + threadIdx.x;
#pragma unroll
  for (int q = 0; q < PACK_PPT; ++q) {
    const u32 i = i0 + q * 256;
    if (i >= n) break;
    int ix, iy, iz;
    coarse_cell(g, pts[i], ix, iy, iz);
    const u32 bin = (u32)(morton3((u32)ix, (u32)iy, (u32)iz) >> g.shift);
    u32 slot = (bin * 2654435761u) >> 21;  // 11 bits
    while (true) {
      const u32 prev = atomicCAS(&s_key[slot], 0xFFFFFFFFu, bin);
      if (prev == 0xFFFFFFFFu || prev == bin) {
        atomicAdd(&s_cnt[slot], 1u);
        break;
      }
      slot = (slot + 1) & (HIST_SLOTS - 1);  // at most 1,024 distinct bins per block: the table never fills
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < HIST_SLOTS; t += 256)
    if (s_cnt[t]) atomicAdd(&hist[s_key[t]], s_cnt[t]);
}
// owner_of_bin[b] = min(W - 1, points_before_b * W / total): contiguous Morton ranges, balanced to the
// granularity of a bin; every rank computes the same table from the same histogram.  `before` is the exclusive
// scan of the histogram (CUB); a single-block scan of the 2^18 bins took 0.57 ms — a quarter of the partition
// phase (profiles/r02_shard_kernels_ncu_full.txt).
__global__ void shard_owner_kernel(const u32* __restrict__ before, u32 nbins, unsigned long long total, int W,
                                   unsigned char* __restrict__ owner) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const unsigned long long o = total ? (unsigned long long)before[b] * (unsigned long long)W / total : 0ull;
  owner[b] = (unsigned char)min(o, (unsigned long long)(W - 1));
}

// route word of a point: bits 0..7 destination mask (owner included), bits 16..19 owner
__device__ __forceinline__ u32 route_of(const CoarseGrid& g, const unsigned char* __restrict__ owner_of_bin,
                                        const float4 p) {
  int ix, iy, iz;
  coarse_cell(g, p, ix, iy, iz);
  // shift = 3 hbits, so the bin of a cell is the Morton code of (ix >> h, iy >> h, iz >> h): the 27 cells of
  // the block fall into at most 2 x 2 x 2 bins — the own one and, per axis, the one a step across the nearer
  // face of the bin-aligned box leads to, if the cell touches that face (and the grid continues there).
  // (27 cell lookups per point near a face were 65 % issue-active and 1 ms per 12.5 M points.)
  const int h = g.hbits;
  const int bx = ix >> h, by = iy >> h, bz = iz >> h;
  const u32 own = owner_of_bin[(u32)morton3((u32)bx, (u32)by, (u32)bz)];
  u32 mask = 1u << own;
  const int m = (1 << h) - 1;
  // candidate bin offsets per axis: 0 always; -1 / +1 when the neighbouring CELL lies in the neighbouring bin
  int ox[3], oy[3], oz[3], nx = 1, ny = 1, nz = 1;
  ox[0] = oy[0] = oz[0] = 0;
  if ((ix & m) == 0 && ix > 0) ox[nx++] = -1;
  if ((ix & m) == m && ix < g.dx - 1) ox[nx++] = 1;
  if ((iy & m) == 0 && iy > 0) oy[ny++] = -1;
  if ((iy & m) == m && iy < g.dy - 1) oy[ny++] = 1;
  if ((iz & m) == 0 && iz > 0) oz[nz++] = -1;
  if ((iz & m) == m && iz < g.dz - 1) oz[nz++] = 1;
  if (nx + ny + nz > 3) {
    for (int a = 0; a < nz; ++a)
      for (int b = 0; b < ny; ++b)
        for (int c = 0; c < nx; ++c) {
          if ((a | b | c) == 0) continue;
          mask |= 1u << owner_of_bin[(u32)morton3((u32)(bx + ox[c]), (u32)(by + oy[b]), (u32)(bz + oz[a]))];
        }
  }
  return mask | (own << 16);
}
// pass 1 of the stable multi-split: route words + per-block, per-destination counts.
// blk_cnt is destination-major ([d][block]) so that ONE exclusive scan over it yields the position of
// every (destination, block) run inside a send buffer that is the concatenation of the destinations.
__global__ void __launch_bounds__(256)
shard_route_kernel(const float4* __restrict__ pts, u32 n, CoarseGrid g,
                   const unsigned char* __restrict__ owner_of_bin, int W, u32* __restrict__ route,
                   u32* __restrict__ blk_cnt, u32 nblk, u32* __restrict__ owned_cnt) {
  __shared__ u32 s_cnt[SHARD_MAX_RANKS][8];  // per warp: sent to d (low half) | of which owned by d (high half)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 i0 = blockIdx.x * PACK_TILE + threadIdx.x * PACK_PPT;
  u32 r[PACK_PPT];
#pragma unroll
  for (int q = 0; q < PACK_PPT; ++q) {
    r[q] = 0xFFFF0000u;  // no point: no destination, an owner nobody is
    if (i0 + q < n) {
      r[q] = route_of(g, owner_of_bin, pts[i0 + q]);
      route[i0 + q] = r[q];
    }
  }
#pragma unroll
  for (int d = 0; d < SHARD_MAX_RANKS; ++d) {
    if (d >= W) break;
    u32 c = 0;
#pragma unroll
    for (int q = 0; q < PACK_PPT; ++q) c += ((r[q] >> d) & 1u) + ((r[q] >> 16) == (u32)d ? 0x10000u : 0u);
    c = __reduce_add_sync(0xffffffffu, c);   // <= 128 per half: no carry between the halves
    if (lane == 0) s_cnt[d][warp] = c;
  }
  __syncthreads();
  if (threadIdx.x < (u32)W) {
    u32 t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_cnt[threadIdx.x][w];
    blk_cnt[(size_t)threadIdx.x * nblk + blockIdx.x] = t & 0xFFFFu;
    if (t >> 16) atomicAdd(&owned_cnt[threadIdx.x], t >> 16);
  }
}
// pass 2: every point is written once per destination, at a position that preserves the input order
// inside each destination's run
// Where the run of destination d starts: in the local send buffer (NCCL exchange: pts[d] = the send buffer,
// base[d] = first record of d's run) or — peer-to-peer exchange — straight in rank d's receive buffer over
// NVLink (pts[d] / gid[d] = that buffer as mapped on this device, base[d] = where this source's block starts
// in it): the pack kernel IS the exchange then, no send buffer, no NCCL transfer.
struct PackDst {
  float4* pts[SHARD_MAX_RANKS];
  u32* gid[SHARD_MAX_RANKS];
  u32 base[SHARD_MAX_RANKS];
};
__global__ void __launch_bounds__(256)
shard_pack_kernel(const float4* __restrict__ pts, u32 n, u32 gid0, const u32* __restrict__ route,
                  const u32* __restrict__ blk_off, u32 nblk, int W, PackDst dst) {
  __shared__ u32 s_cnt[SHARD_MAX_RANKS][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 i0 = blockIdx.x * PACK_TILE + threadIdx.x * PACK_PPT;
  u32 r[PACK_PPT];
#pragma unroll
  for (int q = 0; q < PACK_PPT; ++q) r[q] = i0 + q < n ? route[i0 + q] : 0u;
  u32 excl[SHARD_MAX_RANKS];
#pragma unroll
  for (int d = 0; d < SHARD_MAX_RANKS; ++d) {
    excl[d] = 0;
    if (d >= W) continue;
    u32 c = 0;
#pragma unroll
    for (int q = 0; q < PACK_PPT; ++q) c += (r[q] >> d) & 1u;
    u32 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    excl[d] = incl - c;
    if (lane == 31) s_cnt[d][warp] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < SHARD_MAX_RANKS; ++d) {
    if (d >= W) continue;
    // position inside d's run (the scan is over the destination-major count array) + where the run goes
    u32 pos = dst.base[d] + (blk_off[(size_t)d * nblk + blockIdx.x] - blk_off[(size_t)d * nblk]) + excl[d];
    for (int w = 0; w < warp; ++w) pos += s_cnt[d][w];
    float4* __restrict__ dp = dst.pts[d];
    u32* __restrict__ dg = dst.gid[d];
#pragma unroll
    for (int q = 0; q < PACK_PPT; ++q) {
      if ((r[q] >> d) & 1u) {
        dp[pos] = pts[i0 + q];
        dg[pos] = (gid0 + i0 + q) | ((r[q] >> 16) == (u32)d ? GID_OWNED : 0u);
        ++pos;
      }
    }
  }
}
// send_cnt[d], send_off[d] from the scanned block counts (blk_off has W * nblk + 1 entries)
__global__ void shard_counts_kernel(const u32* __restrict__ blk_off, u32 nblk, int W, u32* __restrict__ out) {
  const int d = threadIdx.x;
  if (d < W) {
    const u32 a = blk_off[(size_t)d * nblk], b = blk_off[(size_t)(d + 1) * nblk];
    out[d] = b - a;                       // points this rank sends to d (owned + halo)
  }
}

// ---- export ------------------------------------------------------------------------------------
__global__ void shard_owned_flag_kernel(const u32* __restrict__ gid, u32 m, u32* __restrict__ flag) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) flag[i] = gid[i] >> 31;
}
__global__ void shard_owned_list_kernel(const u32* __restrict__ gid, const u32* __restrict__ pos, u32 m,
                                        u32* __restrict__ owned_idx) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m && (gid[i] >> 31)) owned_idx[pos[i]] = i;
}
// per owned point: id, coordinates, normal, and the exactness test of its row
__global__ void shard_export_points_kernel(const u32* __restrict__ owned_idx, u32 n_owned, const u32* __restrict__ gid,
                                           const float4* __restrict__ pts, const u32* __restrict__ inv,
                                           const float4* __restrict__ nrm, const u64* __restrict__ kth,
                                           float s2_limit, int have_all, u32* __restrict__ o_gid,
                                           float4* __restrict__ o_pts, float4* __restrict__ o_nrm,
                                           u32* __restrict__ viol /* [0] violations, [1] max k-th d2 bits */) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  u32 kb = 0;
  if (j < n_owned) {
    const u32 i = owned_idx[j];
    const u32 s = inv[i];
    o_gid[j] = gid[i] & GID_MASK;
    o_pts[j] = pts[i];
    if (o_nrm) o_nrm[j] = nrm[s];
    const u64 key = kth[s];
    if (key != 0xFFFFFFFFFFFFFFFFull) kb = (u32)(key >> 32);
    // fewer than k points here although the cloud holds more, or a k-th neighbour beyond the halo
    if (!have_all) bad = key == 0xFFFFFFFFFFFFFFFFull || !(__uint_as_float(kb) <= s2_limit);
  }
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  kb = __reduce_max_sync(0xffffffffu, kb);  // non-negative floats order as unsigned ints
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicAdd(viol, (u32)__popc(m));
    if (kb > viol[1]) atomicMax(viol + 1, kb);
  }
}
// global id of every point in the ctx's sorted (Morton) space: the row export then needs ONE gather per
// entry, and a Morton-local one (gid[perm[sid]] was two, the second at random: 3.1 ms for 12.5 M rows)
__global__ void shard_gid_sorted_kernel(const u32* __restrict__ gid, const u32* __restrict__ perm, u32 m,
                                        u32* __restrict__ gid_sorted) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) gid_sorted[s] = gid[perm[s]] & GID_MASK;
}
// per row entry: neighbour translated to its global id
__global__ void shard_export_rows_kernel(const u32* __restrict__ owned_idx, u32 n_owned, int k,
                                         const u32* __restrict__ gid_sorted, const u32* __restrict__ inv,
                                         const u32* __restrict__ nbr, const float* __restrict__ nd2,
                                         int* __restrict__ o_knn, float* __restrict__ o_d2) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n_owned * k) return;
  const u32 j = (u32)(e / (u32)k);
  const int col = (int)(e - (size_t)j * k);
  const size_t src = (size_t)inv[owned_idx[j]] * k + col;
  const u32 sid = nbr[src];
  o_knn[e] = sid == SID_NONE ? -1 : (int)gid_sorted[sid];
  o_d2[e] = nd2[src];
}
// the same, four entries per thread (k a multiple of 4: rows are 16-byte aligned on both sides): 16-byte loads and
// stores, four gathers in flight per thread (one entry per thread moved 274 MB per million rows at 2.3 TB/s)
__global__ void shard_export_rows4_kernel(const u32* __restrict__ owned_idx, u32 n_owned, int k4 /* k / 4 */,
                                          const u32* __restrict__ gid_sorted, const u32* __restrict__ inv,
                                          const uint4* __restrict__ nbr4, const float4* __restrict__ nd24,
                                          int4* __restrict__ o_knn4, float4* __restrict__ o_d24) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n_owned * k4) return;
  const u32 j = (u32)(e / (u32)k4);
  const int col = (int)(e - (size_t)j * k4);
  const size_t src = (size_t)inv[owned_idx[j]] * k4 + col;
  const uint4 v = nbr4[src];
  const float4 d = nd24[src];
  int4 o;
  o.x = v.x == SID_NONE ? -1 : (int)gid_sorted[v.x];
  o.y = v.y == SID_NONE ? -1 : (int)gid_sorted[v.y];
  o.z = v.z == SID_NONE ? -1 : (int)gid_sorted[v.z];
  o.w = v.w == SID_NONE ? -1 : (int)gid_sorted[v.w];
  o_knn4[e] = o;
  o_d24[e] = d;
}

// ---- gather to the root (iftwt_sharded_segment) ------------------------------------------------
__global__ void shard_pos_of_gid_kernel(const u32* __restrict__ gid, u32 n, u32* __restrict__ pos_of_gid,
                                        const float4* __restrict__ pts, float4* __restrict__ pts_by_gid) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u32 g = gid[p];
  pos_of_gid[g] = p;
  pts_by_gid[g] = pts[p];
}
// gathered rows (global ids, in arrival order) -> the root ctx's sorted-space lists
__global__ void shard_install_kernel(const u32* __restrict__ pos_of_gid, const int* __restrict__ knn,
                                     const float* __restrict__ d2, const float4* __restrict__ nrm_in,
                                     const u32* __restrict__ perm, const u32* __restrict__ inv, u32 n, int k,
                                     u32* __restrict__ nbr, float* __restrict__ nd2, u64* __restrict__ kth,
                                     float4* __restrict__ nrm) {
  const u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n) return;
  const u32 p = pos_of_gid[perm[s]];
  for (int j = lane; j < k; j += 32) {
    const int v = knn[(size_t)p * k + j];
    const float d = d2[(size_t)p * k + j];
    nbr[(size_t)s * k + j] = v < 0 ? SID_NONE : inv[v];
    nd2[(size_t)s * k + j] = d;
    if (j == k - 1) kth[s] = v < 0 ? 0xFFFFFFFFFFFFFFFFull : (((u64)__float_as_uint(d) << 32) | (u64)(u32)v);
  }
  if (lane == 0 && nrm) nrm[s] = nrm_in[p];
}
__global__ void shard_sum_u32_kernel(u32* __restrict__ dst, const u32* __restrict__ src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

}  // namespace

// ---- the communicator --------------------------------------------------------------------------
// Buffers that PEERS write into directly over NVLink (cudaIpc mappings between processes, peer access inside
// one): the receive buffers of the exchange and, on the root of iftwt_sharded_segment, the gathered rows.  A
// buffer a peer may still have mapped is never freed before the comm is destroyed: growing one allocates a new
// one, retires the old one to the graveyard and bumps the shard's epoch, which every rank sees in phase A of the
// next call and answers with a fresh exchange of handles.
enum { PB_RX_PTS = 0, PB_RX_GID, PB_G_GID, PB_G_PTS, PB_G_KNN, PB_G_D2, PB_G_NRM, PB_COUNT };
struct PeerBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct PeerRecord {  // what a rank tells the others about its peer-writable buffers
  cudaIpcMemHandle_t handle[PB_COUNT];
  unsigned long long raw[PB_COUNT];  // the owner's own pointer (what a shard of the same process uses)
  unsigned long long cap[PB_COUNT];
  u32 epoch;
  int pid;
  int device;
  u32 pad;
};
struct PeerView {  // a rank's buffers as THIS shard can address them
  void* p[PB_COUNT] = {};
  size_t cap[PB_COUNT] = {};
  u32 epoch = 0xFFFFFFFFu;
  bool ok = false;
};
struct IpcMapping {
  cudaIpcMemHandle_t handle;
  void* ptr;
};
struct Shard {
  int rank = 0;
  int device = 0;
  iftwt_ctx* ctx = nullptr;       // local compute; its stream carries everything this shard does
  ncclComm_t nccl = nullptr;
  void* h_pin = nullptr;          // pinned scratch (4 KB)
  DevBuf meta, meta_all, hist, hist_scan, owner, route, blk, blk_off, send_pts, send_gid, cnt, cnt_all, flag, pos,
      owned_idx, small, sample_q, sample_d2, sample_idx, peer_rec, peer_rec_all, in_pts;
  DevBuf o_gid, o_pts, o_knn, o_d2, o_nrm;             // results of iftwt_sharded_knn_normals
  DevBuf g_pos;                                        // root: where each gathered row sits
  PeerBuf pb[PB_COUNT];                                // rx_* : received points / ids; g_* : root's gathered rows
  u32 epoch = 0;
  std::vector<void*> graveyard;
  PeerView view[SHARD_MAX_RANKS];
  std::vector<IpcMapping> mapped;
  cudaEvent_t ev[6] = {};
  iftwt_shard_out out{};
  size_t m_total = 0, n_owned = 0;
  bool exported_to_root = false;  // the last export wrote straight into the root's gathered-row buffers
};
struct iftwt_comm {
  int nranks = 0;
  bool loopback = false;          // every shard on one device, collectives are device-to-device copies
  bool p2p = true;                // peers' buffers are addressable (settled by peers_refresh)
  std::vector<Shard> sh;          // the LOCAL shards (one in the process-per-GPU mode)
  std::string err;
};

static thread_local std::string g_comm_err;

namespace {

int comm_fail(iftwt_comm* cm, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (cm) cm->err = buf;
  else g_comm_err = buf;
  return code;
}
#define CM_CU(cm, expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return comm_fail((cm), IFTWT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                     \
  } while (0)
#define CM_NCCL(cm, expr)                                                                       \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess)                                                                      \
      return comm_fail((cm), IFTWT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), \
                       __FILE__, __LINE__);                                                     \
  } while (0)
// inside ncclGroupStart / ncclGroupEnd: a failure closes the group before it is reported
#define CM_NCCL_G(cm, expr)                                                                     \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess) {                                                                    \
      g_nccl.GroupEnd();                                                                        \
      return comm_fail((cm), IFTWT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), \
                       __FILE__, __LINE__);                                                     \
    }                                                                                           \
  } while (0)
// a stage function of the shard's ctx failed: its message becomes the comm's
#define CM_CTX(cm, s, expr)                                                       \
  do {                                                                            \
    int _r = (expr);                                                              \
    if (_r != IFTWT_OK) {                                                         \
      (cm)->err = "rank " + std::to_string((s).rank) + ": " + (s).ctx->err;       \
      return _r;                                                                  \
    }                                                                             \
  } while (0)
#define SH_ENSURE(cm, s, buf, bytes) CM_CTX(cm, s, dev_ensure((s).ctx, (buf), (bytes)))
#define SH_LAUNCH(s, kernel, grid, block, ...) LAUNCH((s).ctx, kernel, grid, block, 0, __VA_ARGS__)

int sync_all(iftwt_comm* cm) {
  for (Shard& s : cm->sh) {
    CM_CU(cm, cudaSetDevice(s.device));
    CM_CU(cm, cudaStreamSynchronize(s.ctx->stream));
  }
  return IFTWT_OK;
}
// ---- collectives over the local shards ----------------------------------------------------------
// send / recv are per-shard device pointers, indexed like cm->sh
int coll_allgather(iftwt_comm* cm, const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes) {
  if (cm->loopback) {
    TRY(sync_all(cm));
    for (size_t a = 0; a < cm->sh.size(); ++a)
      for (size_t b = 0; b < cm->sh.size(); ++b)
        CM_CU(cm, cudaMemcpyAsync((char*)recv[a] + (size_t)cm->sh[b].rank * bytes, send[b], bytes,
                                  cudaMemcpyDeviceToDevice, cm->sh[a].ctx->stream));
    return IFTWT_OK;
  }
  CM_NCCL(cm, g_nccl.GroupStart());
  for (size_t a = 0; a < cm->sh.size(); ++a) {
    CM_CU(cm, cudaSetDevice(cm->sh[a].device));
    CM_NCCL_G(cm, g_nccl.AllGather(send[a], recv[a], bytes, ncclUint8, cm->sh[a].nccl, cm->sh[a].ctx->stream));
  }
  CM_NCCL(cm, g_nccl.GroupEnd());
  return IFTWT_OK;
}
int coll_allreduce_sum_u32(iftwt_comm* cm, const std::vector<u32*>& buf, size_t count) {
  if (cm->loopback) {
    TRY(sync_all(cm));
    cudaStream_t st = cm->sh[0].ctx->stream;
    for (size_t b = 1; b < cm->sh.size(); ++b)
      shard_sum_u32_kernel<<<div_up(count, 256), 256, 0, st>>>(buf[0], buf[b], count);
    for (size_t b = 1; b < cm->sh.size(); ++b)
      CM_CU(cm, cudaMemcpyAsync(buf[b], buf[0], count * 4, cudaMemcpyDeviceToDevice, st));
    CM_CU(cm, cudaStreamSynchronize(st));
    return IFTWT_OK;
  }
  CM_NCCL(cm, g_nccl.GroupStart());
  for (size_t a = 0; a < cm->sh.size(); ++a) {
    CM_CU(cm, cudaSetDevice(cm->sh[a].device));
    CM_NCCL_G(cm, g_nccl.AllReduce(buf[a], buf[a], count, ncclUint32, ncclSum, cm->sh[a].nccl, cm->sh[a].ctx->stream));
  }
  CM_NCCL(cm, g_nccl.GroupEnd());
  return IFTWT_OK;
}
// One payload of an all-to-all-v: rank a sends cnt[a][b] elements, starting at element send_off[a][b]
// of its send buffer, to rank b, which stores them at element recv_off[b][a] of its receive buffer.
struct A2APayload {
  std::vector<const void*> send;  // per local shard
  std::vector<void*> recv;
  size_t elem;
};
int coll_alltoallv(iftwt_comm* cm, const std::vector<A2APayload>& pay, const std::vector<u32>& M /*[W][W]*/,
                   const std::vector<std::vector<size_t>>& send_off, const std::vector<std::vector<size_t>>& recv_off) {
  const int W = cm->nranks;
  if (cm->loopback) {
    TRY(sync_all(cm));
    for (size_t a = 0; a < cm->sh.size(); ++a)
      for (size_t b = 0; b < cm->sh.size(); ++b) {
        const int ra = cm->sh[a].rank, rb = cm->sh[b].rank;
        const size_t c = M[(size_t)ra * W + rb];
        if (!c) continue;
        for (const A2APayload& p : pay)
          CM_CU(cm, cudaMemcpyAsync((char*)p.recv[b] + recv_off[b][ra] * p.elem,
                                    (const char*)p.send[a] + send_off[a][rb] * p.elem, c * p.elem,
                                    cudaMemcpyDeviceToDevice, cm->sh[b].ctx->stream));
      }
    return IFTWT_OK;
  }
  CM_NCCL(cm, g_nccl.GroupStart());
  for (size_t a = 0; a < cm->sh.size(); ++a) {
    Shard& s = cm->sh[a];
    CM_CU(cm, cudaSetDevice(s.device));
    for (int r = 0; r < W; ++r) {
      const size_t cs = M[(size_t)s.rank * W + r], cr = M[(size_t)r * W + s.rank];
      for (const A2APayload& p : pay) {
        if (r == s.rank) {  // own share: a copy on the shard's stream
          if (cs)
            CM_CU(cm, cudaMemcpyAsync((char*)p.recv[a] + recv_off[a][r] * p.elem,
                                      (const char*)p.send[a] + send_off[a][r] * p.elem, cs * p.elem,
                                      cudaMemcpyDeviceToDevice, s.ctx->stream));
          continue;
        }
        if (cs)
          CM_NCCL_G(cm, g_nccl.Send((const char*)p.send[a] + send_off[a][r] * p.elem, cs * p.elem, ncclUint8, r, s.nccl,
                                  s.ctx->stream));
        if (cr)
          CM_NCCL_G(cm, g_nccl.Recv((char*)p.recv[a] + recv_off[a][r] * p.elem, cr * p.elem, ncclUint8, r, s.nccl,
                                  s.ctx->stream));
      }
    }
  }
  CM_NCCL(cm, g_nccl.GroupEnd());
  return IFTWT_OK;
}

int shard_init(iftwt_comm* cm, Shard& s, int rank, int device) {
  s.rank = rank;
  s.device = device;
  int rc = iftwt_create(&s.ctx, device);
  if (rc != IFTWT_OK) return comm_fail(cm, rc, "rank %d: %s", rank, iftwt_last_error(nullptr));
  CM_CU(cm, cudaSetDevice(device));
  CM_CU(cm, cudaMallocHost(&s.h_pin, 4096));
  for (cudaEvent_t& e : s.ev) CM_CU(cm, cudaEventCreate(&e));
  return IFTWT_OK;
}
void shard_free(Shard& s) {
  if (!s.ctx) return;
  cudaSetDevice(s.device);
  cudaStreamSynchronize(s.ctx->stream);
  DevBuf* bufs[] = {&s.meta, &s.meta_all, &s.hist, &s.hist_scan, &s.owner, &s.route, &s.blk, &s.blk_off, &s.send_pts, &s.send_gid,
                    &s.cnt, &s.cnt_all, &s.flag, &s.pos, &s.owned_idx, &s.small, &s.sample_q, &s.sample_d2,
                    &s.sample_idx, &s.peer_rec, &s.peer_rec_all, &s.in_pts, &s.o_gid, &s.o_pts, &s.o_knn, &s.o_d2, &s.o_nrm, &s.g_pos};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (IpcMapping& m : s.mapped) cudaIpcCloseMemHandle(m.ptr);
  s.mapped.clear();
  for (PeerBuf& b : s.pb)
    if (b.p) cudaFree(b.p);
  for (void* g : s.graveyard) cudaFree(g);
  s.graveyard.clear();
  if (s.h_pin) cudaFreeHost(s.h_pin);
  for (cudaEvent_t& e : s.ev)
    if (e) cudaEventDestroy(e);
  if (s.nccl && g_nccl.h) g_nccl.CommDestroy(s.nccl);
  iftwt_destroy(s.ctx);
  s.ctx = nullptr;
}

// the ctx of a shard takes the first m points of its pts_in buffer as its cloud
void ctx_adopt_cloud(iftwt_ctx* c, size_t m) {
  ctx_invalidate(c);
  c->n = m;
  c->stats.n_points = m;
}

// grow-only; the old buffer is retired, not freed (a peer may have it mapped)
int peer_ensure(iftwt_comm* cm, Shard& s, int which, size_t bytes) {
  PeerBuf& b = s.pb[which];
  if (bytes <= b.cap) return IFTWT_OK;
  const size_t want = bytes + bytes / 4 + 4096;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) return comm_fail(cm, IFTWT_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  if (b.p) s.graveyard.push_back(b.p);
  b.p = p;
  b.cap = want;
  ++s.epoch;
  return IFTWT_OK;
}

// Collective: every rank publishes handles / pointers / capacities of its peer-writable buffers, every shard
// maps what it does not share an address space with.  Called when some rank's epoch moved (phase A).
int peers_refresh(iftwt_comm* cm) {
  const int W = cm->nranks;
  const size_t L = cm->sh.size();
  std::vector<const void*> sendp(L);
  std::vector<void*> recvp(L);
  std::vector<PeerRecord> mine(L);
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    PeerRecord& r = mine[l];
    memset(&r, 0, sizeof r);
    r.epoch = s.epoch;
    r.pid = (int)getpid();
    r.device = s.device;
    for (int b = 0; b < PB_COUNT; ++b) {
      r.raw[b] = (unsigned long long)(uintptr_t)s.pb[b].p;
      r.cap[b] = s.pb[b].cap;
      if (s.pb[b].p && (int)L < W) {  // other processes exist: they need a handle
        if (cudaIpcGetMemHandle(&r.handle[b], s.pb[b].p) != cudaSuccess) {
          cudaGetLastError();
          r.raw[b] = 0;  // cannot be shared: peers see no buffer here and keep to NCCL
          r.cap[b] = 0;
        }
      }
    }
    SH_ENSURE(cm, s, s.peer_rec, sizeof(PeerRecord));
    SH_ENSURE(cm, s, s.peer_rec_all, sizeof(PeerRecord) * W);
    CM_CU(cm, cudaMemcpyAsync(s.peer_rec.p, &r, sizeof r, cudaMemcpyHostToDevice, s.ctx->stream));
    sendp[l] = s.peer_rec.p;
    recvp[l] = s.peer_rec_all.p;
  }
  TRY(sync_all(cm));  // `mine` is pageable: the copies above have read it
  TRY(coll_allgather(cm, sendp, recvp, sizeof(PeerRecord)));
  std::vector<PeerRecord> all(W);
  {
    Shard& s = cm->sh[0];
    CM_CU(cm, cudaSetDevice(s.device));
    CM_CU(cm, cudaMemcpyAsync(all.data(), s.peer_rec_all.p, sizeof(PeerRecord) * W, cudaMemcpyDeviceToHost, s.ctx->stream));
    TRY(sync_all(cm));
  }
  u32 ok_local = 1;
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    for (int r = 0; r < W; ++r) {
      PeerView& v = s.view[r];
      const PeerRecord& rec = all[r];
      const Shard* local = nullptr;
      for (const Shard& t : cm->sh)
        if (t.rank == r) local = &t;
      v.ok = true;
      for (int b = 0; b < PB_COUNT; ++b) {
        v.p[b] = nullptr;
        v.cap[b] = 0;
        if (!rec.raw[b]) continue;
        if (local) {
          if (local->device != s.device) {  // same process, another GPU: peer access, once
            cudaError_t e = cudaDeviceEnablePeerAccess(local->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
              cudaGetLastError();
              v.ok = false;
              continue;
            }
            cudaGetLastError();
          }
          v.p[b] = (void*)(uintptr_t)rec.raw[b];
        } else {
          void* ptr = nullptr;
          for (const IpcMapping& m : s.mapped)
            if (memcmp(&m.handle, &rec.handle[b], sizeof(cudaIpcMemHandle_t)) == 0) ptr = m.ptr;
          if (!ptr) {
            if (cudaIpcOpenMemHandle(&ptr, rec.handle[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
              cudaGetLastError();
              v.ok = false;
              continue;
            }
            IpcMapping m;
            m.handle = rec.handle[b];
            m.ptr = ptr;
            s.mapped.push_back(m);
          }
          v.p[b] = ptr;
        }
        v.cap[b] = (size_t)rec.cap[b];
      }
      v.epoch = rec.epoch;
      if (!v.ok) ok_local = 0;
    }
  }
  // every rank must take the same path afterwards: peer-to-peer only if every shard could map every buffer
  std::vector<u32*> okp(L);
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    SH_ENSURE(cm, s, s.small, 64);
    *(u32*)s.h_pin = ok_local ? 0u : 1u;  // summed: 0 = nobody failed
    CM_CU(cm, cudaMemcpyAsync(s.small.p, s.h_pin, 4, cudaMemcpyHostToDevice, s.ctx->stream));
    okp[l] = s.small.as<u32>();
  }
  TRY(coll_allreduce_sum_u32(cm, okp, 1));
  {
    Shard& s = cm->sh[0];
    CM_CU(cm, cudaSetDevice(s.device));
    CM_CU(cm, cudaMemcpyAsync(s.h_pin, s.small.p, 4, cudaMemcpyDeviceToHost, s.ctx->stream));
    TRY(sync_all(cm));
    if (*(u32*)s.h_pin != 0) cm->p2p = false;  // for good: the exchange and the gather stay on NCCL
  }
  return IFTWT_OK;
}

}  // namespace

// =================================================================================================
// gather_root >= 0: the caller is iftwt_sharded_segment — when the root's gathered-row buffers are addressable
// and large enough, the export kernels write the rows straight into them (export and gather fused).
static int sharded_knn_normals_impl(iftwt_comm* cm, const iftwt_shard_in* in, int k, int want_normals,
                                    double halo_radius, iftwt_shard_out* out, int gather_root) {
  const int W = cm->nranks;
  const size_t L = cm->sh.size();
  if (!in || !out) return comm_fail(cm, IFTWT_ERR_INVALID, "null shard descriptors");
  if (k < 1 || k > 64) return comm_fail(cm, IFTWT_ERR_INVALID, "k must be in [1, 64], got %d", k);
  for (size_t l = 0; l < L; ++l) {
    if (in[l].n >= 0x7FFFFFF0ull || in[l].gid0 + in[l].n > (uint64_t)GID_MASK)
      return comm_fail(cm, IFTWT_ERR_UNSUPPORTED, "sharded clouds are limited to 2^31 - 1 points");
    if (in[l].n && !in[l].d_points) return comm_fail(cm, IFTWT_ERR_INVALID, "null slice");
  }
  // ---- phase A: bounding boxes, slice table, halo-radius estimate ----
  const bool estimate = !(halo_radius > 0.0);
  std::vector<const void*> sendp(L);
  std::vector<void*> recvp(L);
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    iftwt_ctx* c = s.ctx;
    SH_ENSURE(cm, s, c->scalars, 4096);
    SH_ENSURE(cm, s, s.meta, sizeof(ShardMeta));
    SH_ENSURE(cm, s, s.meta_all, sizeof(ShardMeta) * W);
    cudaEventRecord(s.ev[0], c->stream);
    const u32 n = (u32)in[l].n;
    const float4* pts = (const float4*)in[l].d_points;
    SH_LAUNCH(s, shard_meta_init_kernel, 1, 32, (ShardMeta*)s.meta.p, n, (u32)in[l].gid0, s.epoch);
    if (n) SH_LAUNCH(s, shard_bbox_kernel, c->sm_count * 8, 256, pts, n, (ShardMeta*)s.meta.p);
    if (estimate && n) {
      // the slice as the ctx's cloud, SHARD_SAMPLE of its points as external queries (general k-NN path)
      SH_ENSURE(cm, s, c->pts_in, (size_t)n * 16);
      CM_CU(cm, cudaMemcpyAsync(c->pts_in.p, pts, (size_t)n * 16, cudaMemcpyDeviceToDevice, c->stream));
      ctx_adopt_cloud(c, n);
      const u32 m = n < SHARD_SAMPLE ? n : SHARD_SAMPLE;
      SH_ENSURE(cm, s, s.sample_q, (size_t)m * 16);
      SH_ENSURE(cm, s, s.sample_d2, (size_t)m * k * 4);
      SH_ENSURE(cm, s, s.sample_idx, (size_t)m * k * 4);
      SH_LAUNCH(s, shard_sample_kernel, div_up(m, 256), 256, pts, n, m, s.sample_q.as<float4>());
      int rc = grid_build(c, k);
      if (rc == IFTWT_ERR_NONFINITE) {
        // reported below, from the gathered records, by every rank alike
      } else {
        CM_CTX(cm, s, rc);
        CM_CTX(cm, s, knn_query_run(c, s.sample_q.as<float4>(), m, k, s.sample_idx.as<u32>(), s.sample_d2.as<float>()));
        SH_LAUNCH(s, shard_kth_max_kernel, div_up(m, 256), 256, (const float*)s.sample_d2.as<float>(), m, k,
                  (ShardMeta*)s.meta.p);
      }
    }
    sendp[l] = s.meta.p;
    recvp[l] = s.meta_all.p;
  }
  TRY(coll_allgather(cm, sendp, recvp, sizeof(ShardMeta)));
  std::vector<ShardMeta> metas(W);
  {
    Shard& s = cm->sh[0];
    CM_CU(cm, cudaSetDevice(s.device));
    CM_CU(cm, cudaMemcpyAsync(metas.data(), s.meta_all.p, sizeof(ShardMeta) * W, cudaMemcpyDeviceToHost, s.ctx->stream));
    TRY(sync_all(cm));
  }
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint64_t n_total = 0, bad = 0;
  float kth2 = 0.f;
  for (int r = 0; r < W; ++r) {
    const ShardMeta& m = metas[r];
    if ((uint64_t)m.gid0 != n_total)
      return comm_fail(cm, IFTWT_ERR_INVALID,
                       "slices must be contiguous global-id ranges, ascending with the rank: rank %d starts at %u, "
                       "expected %llu", r, m.gid0, (unsigned long long)n_total);
    n_total += m.n;
    bad += m.bad;
    if (m.n > m.bad)
      for (int a = 0; a < 3; ++a) {
        lo[a] = fmin(lo[a], (double)ord2f_h(m.lo[a]));
        hi[a] = fmax(hi[a], (double)ord2f_h(m.hi[a]));
      }
    float f;
    memcpy(&f, &m.kth_bits, 4);
    kth2 = fmaxf(kth2, f);
  }
  if (bad) return comm_fail(cm, IFTWT_ERR_NONFINITE, "the cloud holds %llu non-finite points", (unsigned long long)bad);
  if (n_total == 0) return comm_fail(cm, IFTWT_ERR_INVALID, "empty cloud");
  {  // some rank (re)allocated a peer-writable buffer since the views were made: every rank sees it here
    static const bool p2p_env = !(getenv("IFTWT_SHARD_P2P") && atoi(getenv("IFTWT_SHARD_P2P")) == 0);
    if (!p2p_env) cm->p2p = false;
    bool stale = false;
    for (int r = 0; r < W; ++r) stale = stale || cm->sh[0].view[r].epoch != metas[r].epoch;
    if (cm->p2p && stale) TRY(peers_refresh(cm));
  }
  if (n_total > (uint64_t)GID_MASK) return comm_fail(cm, IFTWT_ERR_UNSUPPORTED, "more than 2^31 - 1 points");
  double ext = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  if (!(ext > 0.0)) ext = 1.0;
  double s_rad = halo_radius;
  // The sample bounds the k-th distance of the SAMPLED points only; the cloud's largest one (a point on a
  // sparse rim) was 1.36 x the sample's on the facade.  A wider halo costs little (0.2 % more points at
  // 12.5 M points per GPU), a violated check costs a whole second exchange: twice the sample's maximum.
  if (estimate) s_rad = kth2 > 0.f ? 2.0 * sqrt((double)kth2) : ext;  // no estimate (tiny slices): everything everywhere
  for (size_t l = 0; l < L; ++l) cudaEventRecord(cm->sh[l].ev[1], cm->sh[l].ctx->stream);

  int attempts = 0;
  float kth_max = 0.f;
  bool p2p_last_x = false, p2p_call_ok = true;
  while (true) {
    ++attempts;
    if (attempts > 40) return comm_fail(cm, IFTWT_ERR_CUDA, "sharded k-NN: halo radius did not converge");
    if (attempts > 1)
      for (size_t l = 0; l < L; ++l) cudaEventRecord(cm->sh[l].ev[1], cm->sh[l].ctx->stream);
    // ---- phase B: coarse grid, histogram, owner table ----
    s_rad = fmax(s_rad, ext / 2097000.0);  // at most 2^21 cells per axis
    CoarseGrid g;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.inv_s = 1.0 / s_rad;
    g.dx = (int)floor((hi[0] - lo[0]) / s_rad) + 1;
    g.dy = (int)floor((hi[1] - lo[1]) / s_rad) + 1;
    g.dz = (int)floor((hi[2] - lo[2]) / s_rad) + 1;
    int maxdim = g.dx > g.dy ? g.dx : g.dy;
    if (g.dz > maxdim) maxdim = g.dz;
    int bits = 1;
    while ((1 << bits) < maxdim) ++bits;
    g.shift = 3 * bits > SHARD_HIST_BITS ? 3 * bits - SHARD_HIST_BITS : 0;
    g.hbits = g.shift / 3;
    g.nbins = 1u << (3 * bits - g.shift);
    const bool single_cell = g.dx == 1 && g.dy == 1 && g.dz == 1;
    std::vector<u32*> histp(L);
    for (size_t l = 0; l < L; ++l) {
      Shard& s = cm->sh[l];
      CM_CU(cm, cudaSetDevice(s.device));
      SH_ENSURE(cm, s, s.hist, (size_t)g.nbins * 4);
      SH_ENSURE(cm, s, s.owner, (size_t)g.nbins);
      CM_CU(cm, cudaMemsetAsync(s.hist.p, 0, (size_t)g.nbins * 4, s.ctx->stream));
      const u32 n = (u32)in[l].n;
      if (n)
        SH_LAUNCH(s, shard_hist_kernel, div_up(n, PACK_TILE), 256, (const float4*)in[l].d_points, n, g, s.hist.as<u32>());
      histp[l] = s.hist.as<u32>();
    }
    TRY(coll_allreduce_sum_u32(cm, histp, g.nbins));
    // ---- phase C: route, stable multi-split, counts ----
    std::vector<u32> nblk(L);
    for (size_t l = 0; l < L; ++l) {
      Shard& s = cm->sh[l];
      CM_CU(cm, cudaSetDevice(s.device));
      const u32 n = (u32)in[l].n;
      const float4* pts = (const float4*)in[l].d_points;
      {
        SH_ENSURE(cm, s, s.hist_scan, (size_t)g.nbins * 4);
        size_t tb = 0;
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(nullptr, tb, s.hist.as<u32>(), s.hist_scan.as<u32>(), (int)g.nbins,
                                                s.ctx->stream));
        SH_ENSURE(cm, s, s.ctx->tmp, tb);
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(s.ctx->tmp.p, tb, s.hist.as<u32>(), s.hist_scan.as<u32>(), (int)g.nbins,
                                                s.ctx->stream));
        s.ctx->stats.kernel_launches += 2;
        SH_LAUNCH(s, shard_owner_kernel, div_up(g.nbins, 256), 256, (const u32*)s.hist_scan.as<u32>(), g.nbins,
                  (unsigned long long)n_total, W, s.owner.as<unsigned char>());
      }
      nblk[l] = n ? div_up(n, PACK_TILE) : 0;
      const size_t nb = nblk[l];
      SH_ENSURE(cm, s, s.route, ((size_t)n + 1) * 4);
      SH_ENSURE(cm, s, s.blk, ((size_t)W * nb + 2) * 4);
      SH_ENSURE(cm, s, s.blk_off, ((size_t)W * nb + 2) * 4);
      SH_ENSURE(cm, s, s.cnt, 2 * SHARD_MAX_RANKS * 4);
      SH_ENSURE(cm, s, s.cnt_all, (size_t)W * 2 * SHARD_MAX_RANKS * 4);
      CM_CU(cm, cudaMemsetAsync(s.cnt.p, 0, 2 * SHARD_MAX_RANKS * 4, s.ctx->stream));
      CM_CU(cm, cudaMemsetAsync((u32*)s.blk.p + (size_t)W * nb, 0, 8, s.ctx->stream));
      if (n) {
        SH_LAUNCH(s, shard_route_kernel, (unsigned)nb, 256, pts, n, g, (const unsigned char*)s.owner.p, W,
                  s.route.as<u32>(), s.blk.as<u32>(), (u32)nb, s.cnt.as<u32>() + SHARD_MAX_RANKS);
        size_t tb = 0;
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(nullptr, tb, s.blk.as<u32>(), s.blk_off.as<u32>(), (int)(W * nb + 1),
                                                s.ctx->stream));
        SH_ENSURE(cm, s, s.ctx->tmp, tb);
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(s.ctx->tmp.p, tb, s.blk.as<u32>(), s.blk_off.as<u32>(),
                                                (int)(W * nb + 1), s.ctx->stream));
        s.ctx->stats.kernel_launches += 2;
        SH_LAUNCH(s, shard_counts_kernel, 1, 32, (const u32*)s.blk_off.as<u32>(), (u32)nb, W, s.cnt.as<u32>());
      }
      sendp[l] = s.cnt.p;
      recvp[l] = s.cnt_all.p;
    }
    TRY(coll_allgather(cm, sendp, recvp, 2 * SHARD_MAX_RANKS * 4));
    std::vector<u32> call((size_t)W * 2 * SHARD_MAX_RANKS);
    {
      Shard& s = cm->sh[0];
      CM_CU(cm, cudaSetDevice(s.device));
      CM_CU(cm, cudaMemcpyAsync(call.data(), s.cnt_all.p, call.size() * 4, cudaMemcpyDeviceToHost, s.ctx->stream));
      TRY(sync_all(cm));
    }
    std::vector<u32> M((size_t)W * W), MO((size_t)W * W);  // sent points / of which owned by the destination
    for (int a = 0; a < W; ++a)
      for (int b = 0; b < W; ++b) {
        M[(size_t)a * W + b] = call[(size_t)a * 2 * SHARD_MAX_RANKS + b];
        MO[(size_t)a * W + b] = call[(size_t)a * 2 * SHARD_MAX_RANKS + SHARD_MAX_RANKS + b];
      }
    // ---- phase D: the exchange ----
    std::vector<std::vector<size_t>> send_off(L, std::vector<size_t>(W)), recv_off(L, std::vector<size_t>(W));
    std::vector<size_t> need(W, 0), owned_of(W, 0);   // points rank d receives / owns
    for (int d = 0; d < W; ++d)
      for (int r = 0; r < W; ++r) {
        need[d] += M[(size_t)r * W + d];
        owned_of[d] += MO[(size_t)r * W + d];
      }
    // peer to peer iff every rank's receive buffers are mapped everywhere and large enough — decided on
    // numbers every rank holds identically (the counts matrix, the capacities published at the last refresh)
    // (an attempt that had to fall back to NCCL grows receive buffers, i.e. moves pointers the views of this call
    // do not know yet: the rest of the call stays on NCCL)
    bool p2p_x = cm->p2p && p2p_call_ok;
    for (int d = 0; d < W && p2p_x; ++d) {
      const PeerView& v = cm->sh[0].view[d];
      p2p_x = v.ok && v.epoch == metas[d].epoch && v.p[PB_RX_PTS] && v.p[PB_RX_GID] &&
              (need[d] + 1) * 16 <= v.cap[PB_RX_PTS] && (need[d] + 1) * 4 <= v.cap[PB_RX_GID];
    }
    A2APayload pp, pg;
    pp.elem = 16;
    pg.elem = 4;
    pp.send.resize(L); pp.recv.resize(L); pg.send.resize(L); pg.recv.resize(L);
    for (size_t l = 0; l < L; ++l) {
      Shard& s = cm->sh[l];
      CM_CU(cm, cudaSetDevice(s.device));
      size_t so = 0, ro = 0, sent_halo = 0;
      for (int r = 0; r < W; ++r) {
        send_off[l][r] = so;
        recv_off[l][r] = ro;
        so += M[(size_t)s.rank * W + r];
        ro += M[(size_t)r * W + s.rank];
        if (r != s.rank) sent_halo += M[(size_t)s.rank * W + r] - MO[(size_t)s.rank * W + r];
      }
      s.m_total = ro;
      s.n_owned = owned_of[s.rank];
      s.out.halo_bytes_sent = (uint64_t)sent_halo * 20u;
      s.out.exchange_bytes_sent = (uint64_t)(so - M[(size_t)s.rank * W + s.rank]) * 20u;
      if (ro >= 0x7FFFFFF0ull) return comm_fail(cm, IFTWT_ERR_UNSUPPORTED, "a shard would hold 2^31 points");
      PackDst dst;
      memset(&dst, 0, sizeof dst);
      if (p2p_x) {
        for (int d = 0; d < W; ++d) {
          dst.pts[d] = (float4*)s.view[d].p[PB_RX_PTS];
          dst.gid[d] = (u32*)s.view[d].p[PB_RX_GID];
          size_t base = 0;  // where this source's block starts in d's receive buffer
          for (int r = 0; r < s.rank; ++r) base += M[(size_t)r * W + d];
          dst.base[d] = (u32)base;
        }
      } else {
        TRY(peer_ensure(cm, s, PB_RX_PTS, (ro + 1) * 16));
        TRY(peer_ensure(cm, s, PB_RX_GID, (ro + 1) * 4));
        SH_ENSURE(cm, s, s.send_pts, (so + 1) * 16);
        SH_ENSURE(cm, s, s.send_gid, (so + 1) * 4);
        for (int d = 0; d < W; ++d) {
          dst.pts[d] = s.send_pts.as<float4>();
          dst.gid[d] = s.send_gid.as<u32>();
          dst.base[d] = (u32)send_off[l][d];
        }
      }
      const u32 n = (u32)in[l].n;
      if (n)
        SH_LAUNCH(s, shard_pack_kernel, nblk[l], 256, (const float4*)in[l].d_points, n, (u32)in[l].gid0,
                  (const u32*)s.route.as<u32>(), (const u32*)s.blk_off.as<u32>(), nblk[l], W, dst);
      cudaEventRecord(s.ev[2], s.ctx->stream);
      pp.send[l] = s.send_pts.p; pp.recv[l] = s.pb[PB_RX_PTS].p;
      pg.send[l] = s.send_gid.p; pg.recv[l] = s.pb[PB_RX_GID].p;
      sendp[l] = s.cnt.p;  // the barrier of the peer-to-peer path: 4 bytes from everybody
      recvp[l] = s.cnt_all.p;
    }
    p2p_last_x = p2p_x;
    if (!p2p_x) p2p_call_ok = false;
    if (p2p_x) {
      // every pack kernel has written into its peers; a rank may read its receive buffer once all of them are
      // done: a 4-byte all-gather, stream-ordered behind the pack kernel on every rank
      TRY(coll_allgather(cm, sendp, recvp, 4));
    } else {
      TRY(coll_alltoallv(cm, {pp, pg}, M, send_off, recv_off));
    }
    if (cm->loopback) TRY(sync_all(cm));
    // the rows go straight into the root's gathered-row buffers (iftwt_sharded_segment) when those are mapped
    // everywhere and hold the whole cloud: export and gather are one kernel then, 164 B per point over NVLink
    // while the export runs instead of a second pass through NCCL
    bool p2p_g = gather_root >= 0 && cm->p2p && want_normals;
    if (p2p_g) {
      const PeerView& v = cm->sh[0].view[gather_root];
      p2p_g = v.ok && v.epoch == metas[gather_root].epoch && v.p[PB_G_GID] && v.p[PB_G_PTS] && v.p[PB_G_KNN] &&
              v.p[PB_G_D2] && v.p[PB_G_NRM] && (n_total + 1) * 4 <= v.cap[PB_G_GID] &&
              (n_total + 1) * 16 <= v.cap[PB_G_PTS] && (n_total * k + 1) * 4 <= v.cap[PB_G_KNN] &&
              (n_total * k + 1) * 4 <= v.cap[PB_G_D2] && (n_total + 1) * 16 <= v.cap[PB_G_NRM];
    }
    // ---- phase E: the single-GPU kernels on owned + halo points ----
    std::vector<u32*> violp(L);
    for (size_t l = 0; l < L; ++l) {
      Shard& s = cm->sh[l];
      CM_CU(cm, cudaSetDevice(s.device));
      iftwt_ctx* c = s.ctx;
      cudaEventRecord(s.ev[3], c->stream);
      const size_t m = s.m_total;
      SH_ENSURE(cm, s, c->pts_in, (m + 1) * 16);
      if (m)  // the receive buffer stays what peers write into; the ctx works on its own copy (0.06 ms per 12.5 M points)
        CM_CU(cm, cudaMemcpyAsync(c->pts_in.p, s.pb[PB_RX_PTS].p, m * 16, cudaMemcpyDeviceToDevice, c->stream));
      ctx_adopt_cloud(c, m);
      if (m) {
        CM_CTX(cm, s, knn_run(c, k));
        if (want_normals) CM_CTX(cm, s, normals_run(c));
      }
      cudaEventRecord(s.ev[4], c->stream);
      // ---- phase F: export the owned rows, count the rows the halo cannot vouch for ----
      const size_t no = s.n_owned;
      SH_ENSURE(cm, s, s.small, 64);
      CM_CU(cm, cudaMemsetAsync(s.small.p, 0, 64, c->stream));
      u32* d_gid;
      float4 *d_pts, *d_nrm;
      int* d_knn;
      float* d_d2;
      s.exported_to_root = p2p_g;
      if (p2p_g) {
        size_t row_base = 0;  // rows arrive in rank order, as the NCCL gather would deliver them
        for (int r = 0; r < s.rank; ++r) row_base += owned_of[r];
        const PeerView& v = s.view[gather_root];
        d_gid = (u32*)v.p[PB_G_GID] + row_base;
        d_pts = (float4*)v.p[PB_G_PTS] + row_base;
        d_nrm = (float4*)v.p[PB_G_NRM] + row_base;
        d_knn = (int*)v.p[PB_G_KNN] + row_base * k;
        d_d2 = (float*)v.p[PB_G_D2] + row_base * k;
      } else {
        SH_ENSURE(cm, s, s.o_gid, (no + 1) * 4);
        SH_ENSURE(cm, s, s.o_pts, (no + 1) * 16);
        SH_ENSURE(cm, s, s.o_knn, (no * k + 1) * 4);
        SH_ENSURE(cm, s, s.o_d2, (no * k + 1) * 4);
        if (want_normals) SH_ENSURE(cm, s, s.o_nrm, (no + 1) * 16);
        d_gid = s.o_gid.as<u32>();
        d_pts = s.o_pts.as<float4>();
        d_nrm = want_normals ? s.o_nrm.as<float4>() : nullptr;
        d_knn = s.o_knn.as<int>();
        d_d2 = s.o_d2.as<float>();
      }
      if (m && no) {
        SH_ENSURE(cm, s, s.flag, (m + 1) * 4);
        SH_ENSURE(cm, s, s.pos, (m + 1) * 4);
        SH_ENSURE(cm, s, s.owned_idx, (no + 1) * 4);
        SH_LAUNCH(s, shard_owned_flag_kernel, div_up(m, 256), 256, (const u32*)s.pb[PB_RX_GID].p, (u32)m,
                  s.flag.as<u32>());
        size_t tb = 0;
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(nullptr, tb, s.flag.as<u32>(), s.pos.as<u32>(), (int)m, c->stream));
        SH_ENSURE(cm, s, c->tmp, tb);
        CM_CU(cm, cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, s.flag.as<u32>(), s.pos.as<u32>(), (int)m, c->stream));
        c->stats.kernel_launches += 2;
        SH_LAUNCH(s, shard_owned_list_kernel, div_up(m, 256), 256, (const u32*)s.pb[PB_RX_GID].p,
                  (const u32*)s.pos.as<u32>(), (u32)m, s.owned_idx.as<u32>());
        const double lim = s_rad * (1.0 - 1e-6);
        const float s2_limit = (float)fmin(lim * lim, 3.0e38);
        // a shard that holds the whole cloud has exact rows whatever the radius
        const int have_all = (uint64_t)m == n_total ? 1 : 0;
        SH_LAUNCH(s, shard_export_points_kernel, div_up(no, 256), 256, (const u32*)s.owned_idx.as<u32>(), (u32)no,
                  (const u32*)s.pb[PB_RX_GID].p, (const float4*)c->pts_in.as<float4>(), (const u32*)c->inv.as<u32>(),
                  (const float4*)(want_normals ? c->nrm.as<float4>() : nullptr), (const u64*)c->kth.as<u64>(),
                  s2_limit, have_all, d_gid, d_pts, d_nrm, s.small.as<u32>());
        u32* gid_sorted = s.flag.as<u32>();  // the flags are consumed: their buffer is reused
        SH_LAUNCH(s, shard_gid_sorted_kernel, div_up(m, 256), 256, (const u32*)s.pb[PB_RX_GID].p,
                  (const u32*)c->perm.as<u32>(), (u32)m, gid_sorted);
        if (k % 4 == 0)
          SH_LAUNCH(s, shard_export_rows4_kernel, div_up(no * (k / 4), 256), 256, (const u32*)s.owned_idx.as<u32>(),
                    (u32)no, k / 4, (const u32*)gid_sorted, (const u32*)c->inv.as<u32>(), (const uint4*)c->nbr.as<uint4>(),
                    (const float4*)c->nd2.as<float4>(), (int4*)d_knn, (float4*)d_d2);
        else
          SH_LAUNCH(s, shard_export_rows_kernel, div_up(no * k, 256), 256, (const u32*)s.owned_idx.as<u32>(), (u32)no, k,
                    (const u32*)gid_sorted, (const u32*)c->inv.as<u32>(), (const u32*)c->nbr.as<u32>(),
                    (const float*)c->nd2.as<float>(), d_knn, d_d2);
      }
      CM_CU(cm, cudaGetLastError());
      cudaEventRecord(s.ev[5], c->stream);
      violp[l] = s.small.as<u32>();
    }
    // {violations, largest k-th distance} of every rank
    for (size_t l = 0; l < L; ++l) {
      sendp[l] = violp[l];
      recvp[l] = cm->sh[l].cnt_all.p;
    }
    TRY(coll_allgather(cm, sendp, recvp, 8));
    u32 viol = 0;
    {
      Shard& s = cm->sh[0];
      CM_CU(cm, cudaSetDevice(s.device));
      CM_CU(cm, cudaMemcpyAsync(s.h_pin, s.cnt_all.p, (size_t)W * 8, cudaMemcpyDeviceToHost, s.ctx->stream));
      TRY(sync_all(cm));
      const u32* hv = (const u32*)s.h_pin;
      kth_max = 0.f;
      for (int r = 0; r < W; ++r) {
        viol += hv[2 * r];
        float f;
        memcpy(&f, &hv[2 * r + 1], 4);
        kth_max = fmaxf(kth_max, f);
      }
    }
    if (viol == 0 || single_cell) break;
    s_rad *= 2.0;
  }
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    iftwt_shard_out& o = s.out;
    o.n_owned = s.n_owned;
    o.n_halo = s.m_total - s.n_owned;
    o.n_total = n_total;
    o.k = k;
    // (rows that went straight to the root of iftwt_sharded_segment are not kept here)
    o.d_gid = s.exported_to_root ? nullptr : s.o_gid.as<uint32_t>();
    o.d_points = s.exported_to_root ? nullptr : s.o_pts.as<float>();
    o.d_knn = s.exported_to_root ? nullptr : s.o_knn.as<int32_t>();
    o.d_d2 = s.exported_to_root ? nullptr : s.o_d2.as<float>();
    o.d_normals = (want_normals && !s.exported_to_root) ? s.o_nrm.as<float>() : nullptr;
    o.peer_to_peer = (p2p_last_x ? 1 : 0) | (s.exported_to_root ? 2 : 0);
    o.halo_radius = s_rad;
    o.kth_distance_max = sqrt((double)kth_max);
    o.attempts = attempts;
    cudaSetDevice(s.device);
    cudaEventElapsedTime(&o.ms_estimate, s.ev[0], s.ev[1]);
    cudaEventElapsedTime(&o.ms_partition, s.ev[1], s.ev[2]);
    cudaEventElapsedTime(&o.ms_exchange, s.ev[2], s.ev[3]);
    cudaEventElapsedTime(&o.ms_compute, s.ev[3], s.ev[4]);
    cudaEventElapsedTime(&o.ms_export, s.ev[4], s.ev[5]);
    o.knn_cell_ms = 0.f;
    if (s.ctx->ev_set[IFTWT_STAGE_KNN_CELL])
      cudaEventElapsedTime(&o.knn_cell_ms, s.ctx->ev0[IFTWT_STAGE_KNN_CELL], s.ctx->ev1[IFTWT_STAGE_KNN_CELL]);
    out[l] = o;
  }
  return IFTWT_OK;
}

// rows, points and normals of every rank -> the root's ctx (graph / seeds / IFT run there)
static int gather_to_root(iftwt_comm* cm, int root, int k) {
  const int W = cm->nranks;
  const size_t L = cm->sh.size();
  // owned counts of every rank
  std::vector<const void*> sendp(L);
  std::vector<void*> recvp(L);
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    *(u32*)s.h_pin = (u32)s.n_owned;
    CM_CU(cm, cudaMemcpyAsync(s.cnt.p, s.h_pin, 4, cudaMemcpyHostToDevice, s.ctx->stream));
    sendp[l] = s.cnt.p;
    recvp[l] = s.cnt_all.p;
  }
  TRY(coll_allgather(cm, sendp, recvp, 4));
  std::vector<u32> owned(W);
  {
    Shard& s = cm->sh[0];
    CM_CU(cm, cudaSetDevice(s.device));
    CM_CU(cm, cudaMemcpyAsync(owned.data(), s.cnt_all.p, (size_t)W * 4, cudaMemcpyDeviceToHost, s.ctx->stream));
    TRY(sync_all(cm));
  }
  std::vector<u32> M((size_t)W * W, 0);
  size_t n_total = 0;
  for (int r = 0; r < W; ++r) {
    M[(size_t)r * W + root] = owned[r];
    n_total += owned[r];
  }
  std::vector<std::vector<size_t>> send_off(L, std::vector<size_t>(W, 0)), recv_off(L, std::vector<size_t>(W, 0));
  A2APayload pg, pp, pk, pd, pn;
  pg.elem = 4; pp.elem = 16; pk.elem = (size_t)k * 4; pd.elem = (size_t)k * 4; pn.elem = 16;
  for (A2APayload* p : {&pg, &pp, &pk, &pd, &pn}) {
    p->send.resize(L);
    p->recv.resize(L);
  }
  for (size_t l = 0; l < L; ++l) {
    Shard& s = cm->sh[l];
    CM_CU(cm, cudaSetDevice(s.device));
    if (s.rank == root) {
      size_t ro = 0;
      for (int r = 0; r < W; ++r) {
        recv_off[l][r] = ro;
        ro += owned[r];
      }
      // peer-writable: the next call's export kernels write here directly (the epoch moves, the peers re-map)
      TRY(peer_ensure(cm, s, PB_G_GID, (n_total + 1) * 4));
      TRY(peer_ensure(cm, s, PB_G_PTS, (n_total + 1) * 16));
      TRY(peer_ensure(cm, s, PB_G_KNN, (n_total * k + 1) * 4));
      TRY(peer_ensure(cm, s, PB_G_D2, (n_total * k + 1) * 4));
      TRY(peer_ensure(cm, s, PB_G_NRM, (n_total + 1) * 16));
    }
    pg.send[l] = s.o_gid.p; pg.recv[l] = s.pb[PB_G_GID].p;
    pp.send[l] = s.o_pts.p; pp.recv[l] = s.pb[PB_G_PTS].p;
    pk.send[l] = s.o_knn.p; pk.recv[l] = s.pb[PB_G_KNN].p;
    pd.send[l] = s.o_d2.p;  pd.recv[l] = s.pb[PB_G_D2].p;
    pn.send[l] = s.o_nrm.p; pn.recv[l] = s.pb[PB_G_NRM].p;
  }
  TRY(coll_alltoallv(cm, {pg, pp, pk, pd, pn}, M, send_off, recv_off));
  if (cm->loopback) TRY(sync_all(cm));
  return IFTWT_OK;
}

extern "C" {

const char* iftwt_comm_last_error(const iftwt_comm* cm) { return cm ? cm->err.c_str() : g_comm_err.c_str(); }

int iftwt_comm_unique_id(void* id128) {
  if (!id128) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null id buffer");
  if (!nccl_load()) return comm_fail(nullptr, IFTWT_ERR_CUDA, "%s", g_nccl.err.c_str());
  static_assert(sizeof(ncclUniqueId) <= IFTWT_COMM_ID_BYTES, "id buffer too small");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return comm_fail(nullptr, IFTWT_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memset(id128, 0, IFTWT_COMM_ID_BYTES);
  memcpy(id128, &id, sizeof id);
  return IFTWT_OK;
}

int iftwt_comm_create(iftwt_comm** out, const void* id128, int nranks, int rank, int device) {
  if (!out || !id128) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (nranks < 1 || nranks > SHARD_MAX_RANKS || rank < 0 || rank >= nranks)
    return comm_fail(nullptr, IFTWT_ERR_INVALID, "rank %d of %d: at most %d ranks (one NVSwitch node)", rank, nranks,
                     SHARD_MAX_RANKS);
  if (!nccl_load()) return comm_fail(nullptr, IFTWT_ERR_CUDA, "%s", g_nccl.err.c_str());
  iftwt_comm* cm = new (std::nothrow) iftwt_comm();
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_CUDA, "out of host memory");
  cm->nranks = nranks;
  cm->sh.resize(1);
  int rc = shard_init(cm, cm->sh[0], rank, device);
  if (rc == IFTWT_OK) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&cm->sh[0].nccl, nranks, id, rank);
    if (r != ncclSuccess) rc = comm_fail(cm, IFTWT_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  }
  if (rc != IFTWT_OK) {
    g_comm_err = cm->err;
    iftwt_comm_destroy(cm);
    return rc;
  }
  *out = cm;
  return IFTWT_OK;
}

int iftwt_comm_create_all(iftwt_comm** out, int ndev, const int* devices) {
  if (!out || !devices) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (ndev < 1 || ndev > SHARD_MAX_RANKS)
    return comm_fail(nullptr, IFTWT_ERR_INVALID, "%d shards: at most %d (one NVSwitch node)", ndev, SHARD_MAX_RANKS);
  bool dup = false;
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < a; ++b) dup = dup || devices[a] == devices[b];
  if (dup)
    for (int a = 1; a < ndev; ++a)
      if (devices[a] != devices[0])
        return comm_fail(nullptr, IFTWT_ERR_INVALID, "shards time-sharing a GPU must all be on the same one");
  iftwt_comm* cm = new (std::nothrow) iftwt_comm();
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_CUDA, "out of host memory");
  cm->nranks = ndev;
  cm->loopback = dup || ndev == 1;
  cm->sh.resize(ndev);
  int rc = IFTWT_OK;
  for (int a = 0; a < ndev && rc == IFTWT_OK; ++a) rc = shard_init(cm, cm->sh[a], a, devices[a]);
  if (rc == IFTWT_OK && !cm->loopback) {
    if (!nccl_load()) rc = comm_fail(cm, IFTWT_ERR_CUDA, "%s", g_nccl.err.c_str());
    if (rc == IFTWT_OK) {
      std::vector<ncclComm_t> comms(ndev);
      ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, devices);
      if (r != ncclSuccess) rc = comm_fail(cm, IFTWT_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(r));
      else
        for (int a = 0; a < ndev; ++a) cm->sh[a].nccl = comms[a];
    }
  }
  if (rc != IFTWT_OK) {
    g_comm_err = cm->err;
    iftwt_comm_destroy(cm);
    return rc;
  }
  *out = cm;
  return IFTWT_OK;
}

void iftwt_comm_destroy(iftwt_comm* cm) {
  if (!cm) return;
  for (Shard& s : cm->sh) shard_free(s);
  delete cm;
}

int iftwt_comm_info(const iftwt_comm* cm, int* nranks, int* n_local, int* first_local_rank) {
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null comm");
  if (nranks) *nranks = cm->nranks;
  if (n_local) *n_local = (int)cm->sh.size();
  if (first_local_rank) *first_local_rank = cm->sh.empty() ? 0 : cm->sh[0].rank;
  return IFTWT_OK;
}

iftwt_ctx* iftwt_comm_ctx(iftwt_comm* cm, int local_index) {
  if (!cm || local_index < 0 || (size_t)local_index >= cm->sh.size()) return nullptr;
  return cm->sh[local_index].ctx;
}

int iftwt_comm_upload(iftwt_comm* cm, int local_index, const void* host_points, size_t n, uint64_t gid0,
                      iftwt_shard_in* in) {
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null comm");
  if (local_index < 0 || (size_t)local_index >= cm->sh.size() || !in || (n && !host_points))
    return comm_fail(cm, IFTWT_ERR_INVALID, "bad upload (local shard %d of %zu)", local_index, cm->sh.size());
  Shard& s = cm->sh[(size_t)local_index];
  CM_CU(cm, cudaSetDevice(s.device));
  SH_ENSURE(cm, s, s.in_pts, (n + 1) * 16);
  if (n) CM_CU(cm, cudaMemcpyAsync(s.in_pts.p, host_points, n * 16, cudaMemcpyHostToDevice, s.ctx->stream));
  CM_CU(cm, cudaStreamSynchronize(s.ctx->stream));
  in->d_points = s.in_pts.p;
  in->n = n;
  in->gid0 = gid0;
  return IFTWT_OK;
}

int iftwt_sharded_knn_normals(iftwt_comm* cm, const iftwt_shard_in* in, int k, int want_normals, double halo_radius,
                              iftwt_shard_out* out) {
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null comm");
  return sharded_knn_normals_impl(cm, in, k, want_normals, halo_radius, out, -1);
}

int iftwt_sharded_segment(iftwt_comm* cm, const iftwt_shard_in* in, const iftwt_segment_params* p, double halo_radius,
                          int root, uint32_t* cost, uint32_t* label, size_t* n_seeds, iftwt_shard_out* out) {
  if (!cm) return comm_fail(nullptr, IFTWT_ERR_INVALID, "null comm");
  if (!p) return comm_fail(cm, IFTWT_ERR_INVALID, "null params");
  if (root < 0 || root >= cm->nranks) return comm_fail(cm, IFTWT_ERR_INVALID, "root %d of %d ranks", root, cm->nranks);
  if (p->weights != 0 && p->weights != 1) return comm_fail(cm, IFTWT_ERR_INVALID, "weights must be 0 or 1");
  if (p->k_normals != p->k_graph || p->rg.number_of_neighbours != p->k_graph || !(p->rg.curvature_threshold >= 0.34f))
    return comm_fail(cm, IFTWT_ERR_UNSUPPORTED,
                     "the sharded pipeline shares one neighbourhood size between graph, normals and region growing "
                     "(and needs the curvature gate inactive, threshold >= 0.34)");
  const int k = p->k_graph;
  TRY(sharded_knn_normals_impl(cm, in, k, 1, halo_radius, out, root));
  if (!cm->sh[0].exported_to_root) TRY(gather_to_root(cm, root, k));  // else: the export kernels wrote them there
  for (Shard& s : cm->sh) {
    if (s.rank != root) continue;
    CM_CU(cm, cudaSetDevice(s.device));
    iftwt_ctx* c = s.ctx;
    const size_t n = out[&s - cm->sh.data()].n_total;
    // the cloud in global-id order + where each gathered row sits
    SH_ENSURE(cm, s, s.g_pos, (n + 1) * 4);
    SH_ENSURE(cm, s, c->pts_in, (n + 1) * 16);
    SH_LAUNCH(s, shard_pos_of_gid_kernel, div_up(n, 256), 256, (const u32*)s.pb[PB_G_GID].p, (u32)n, s.g_pos.as<u32>(),
              (const float4*)s.pb[PB_G_PTS].p, c->pts_in.as<float4>());
    ctx_adopt_cloud(c, n);
    CM_CTX(cm, s, grid_build(c, k));
    SH_ENSURE(cm, s, c->nbr, n * k * 4);
    SH_ENSURE(cm, s, c->nd2, n * k * 4);
    SH_ENSURE(cm, s, c->kth, n * 8);
    SH_ENSURE(cm, s, c->fb_list, n * 8);
    SH_ENSURE(cm, s, c->nrm, n * 16);
    SH_LAUNCH(s, shard_install_kernel, div_up(n * 32, 256), 256, (const u32*)s.g_pos.as<u32>(),
              (const int*)s.pb[PB_G_KNN].p, (const float*)s.pb[PB_G_D2].p, (const float4*)s.pb[PB_G_NRM].p,
              (const u32*)c->perm.as<u32>(), (const u32*)c->inv.as<u32>(), (u32)n, k, c->nbr.as<u32>(),
              c->nd2.as<float>(), c->kth.as<u64>(), c->nrm.as<float4>());
    CM_CU(cm, cudaGetLastError());
    c->knn_k = k;   // lists and normals are installed, not recomputed
    c->nrm_k = k;
    ((u32*)c->h_scalars)[48] = 0;
    CM_CTX(cm, s, fused_graph_rg_run(c, &p->rg));
    CM_CTX(cm, s, graph_build(c));
    CM_CTX(cm, s, seeds_run(c, &p->rg));
    if (n_seeds) *n_seeds = c->n_seeds;
    const float* d_w = nullptr;
    if (p->weights == 1) {
      SH_ENSURE(cm, s, c->out_b, n * 4);
      CM_CTX(cm, s, morph_run(c, IFTWT_MORPH_GRADIENT, c->out_b.as<float>()));
      d_w = c->out_b.as<float>();
    }
    CM_CTX(cm, s, ift_run(c, d_w, c->seeds.as<u32>(), c->n_seeds));
    if (cost) CM_CU(cm, cudaMemcpyAsync(cost, c->out_a.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (label) CM_CU(cm, cudaMemcpyAsync(label, c->out_a.as<u32>() + n, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CM_CU(cm, cudaStreamSynchronize(c->stream));
  }
  return IFTWT_OK;
}

}  // extern "C"
