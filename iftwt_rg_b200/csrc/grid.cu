// grid.cu — K0: bounding box, cell-size estimate, Morton sort, cell table, cell hash.
//
// Replaces the PCL KdTree / octree builds of graph.hpp:39 and graph.hpp:297-299.
// The cloud is binned into a uniform grid whose cell is sized so that the 3x3x3
// block around a query holds its k nearest neighbours (k-NN fast path, knn.cu);
// points are sorted by the Morton code of their cell (x = most significant axis
// bit) so every cell is one contiguous run of the sorted array and spatially
// close points are close in memory for every later gather.
#include <cub/cub.cuh>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

int ctx_fail(iftwt_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int dev_ensure(iftwt_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return IFTWT_OK;
  size_t want = bytes + bytes / 8 + 256;
  if (b.p) {
    cudaStreamSynchronize(c->stream);
    cudaFree(b.p);
    c->stats.device_bytes -= b.cap;
    b.p = nullptr;
    b.cap = 0;
  }
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return ctx_fail(c, IFTWT_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  c->stats.device_bytes += want;
  return IFTWT_OK;
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
static inline float ord2f(int i) {
  int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
  float f;
  memcpy(&f, &j, 4);
  return f;
}

// scalars layout (ints): [0..2] min xyz, [3..5] max xyz, [6] non-finite count
__global__ void bbox_kernel(const float4* __restrict__ pts, u32 n, int* __restrict__ sc) {
  int mn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
  int mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  int bad = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      bad++;
      continue;
    }
    int a = f2ord(p.x), b = f2ord(p.y), cc = f2ord(p.z);
    mn[0] = min(mn[0], a); mx[0] = max(mx[0], a);
    mn[1] = min(mn[1], b); mx[1] = max(mx[1], b);
    mn[2] = min(mn[2], cc); mx[2] = max(mx[2], cc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&sc[a], mn[a]);
      atomicMax(&sc[3 + a], mx[a]);
    }
    if (bad) atomicAdd(&sc[6], bad);
  }
}

// 16 bits / axis Morton key over the bounding cube, for the occupancy-per-level estimate
// (of every `sub`-th point: the estimate is statistical, the full cloud is not needed)
__global__ void fine_key_kernel(const float4* __restrict__ pts, u32 m, u32 sub, double ox, double oy, double oz,
                                double scale, u64* __restrict__ keys) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float4 p = pts[(size_t)i * sub];
  u32 ix = (u32)fmin(65535.0, fmax(0.0, ((double)p.x - ox) * scale));
  u32 iy = (u32)fmin(65535.0, fmax(0.0, ((double)p.y - oy) * scale));
  u32 iz = (u32)fmin(65535.0, fmax(0.0, ((double)p.z - oz) * scale));
  keys[i] = morton3(ix, iy, iz);
}

__device__ __forceinline__ u32 lower_bound_keys(const u64* __restrict__ a, u32 n, u64 v) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
// Point-weighted cell occupancy at every octree level of the bounding cube: for a
// sample of points, the number of points sharing their level-l cell (a Morton-prefix
// range of the sorted keys).  Unlike N / #occupied cells this is not dragged down by
// the partially filled cells along surface boundaries, so it measures the density a
// query actually sees.  sums[l] accumulates the run lengths, l = 0..16.
// hist[l][b] additionally counts, for every 8th sample, the run length in OCC_BINS_PER_OCTAVE bins per octave:
// the host reads a low percentile of the occupancy from it (density spread of the cloud, see grid_build).
#define OCC_BINS_PER_OCTAVE 8
#define OCC_BINS 256
__global__ void occupancy_kernel(const u64* __restrict__ keys, u32 n, u32 stride, u32 n_samples,
                                 unsigned long long* __restrict__ sums, u32* __restrict__ hist) {
  __shared__ unsigned long long sh[17];
  if (threadIdx.x < 17) sh[threadIdx.x] = 0ull;
  __syncthreads();
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_samples) {
    u64 key = keys[(size_t)s * stride];
    u32 lo = 0, hi = n;
    for (int l = 1; l <= 16; ++l) {
      int sh_bits = 3 * (16 - l);
      u64 pfx = key >> sh_bits;
      // the level-l range nests inside the level-(l-1) range
      u32 a = lo + lower_bound_keys(keys + lo, hi - lo, pfx << sh_bits);
      u32 b = lo + lower_bound_keys(keys + lo, hi - lo, (pfx + 1) << sh_bits);
      lo = a;
      hi = b;
      atomicAdd(&sh[l], (unsigned long long)(hi - lo));
      if ((s & 7u) == 0u && hi > lo) {
        const int bin = min(OCC_BINS - 1, (int)((float)OCC_BINS_PER_OCTAVE * log2f((float)(hi - lo))));
        atomicAdd(&hist[l * OCC_BINS + bin], 1u);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x < 17 && sh[threadIdx.x]) atomicAdd(&sums[threadIdx.x], sh[threadIdx.x]);
}

// KeyT = u32 when the Morton code of a cell fits 32 bits (a third less sort traffic), else u64
template <class KeyT>
__global__ void cell_key_kernel(const float4* __restrict__ pts, u32 n, double ox, double oy, double oz,
                                double inv_cell, int dx, int dy, int dz, KeyT* __restrict__ keys,
                                u32* __restrict__ vals) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int ix = min(dx - 1, max(0, (int)floor(((double)p.x - ox) * inv_cell)));
  int iy = min(dy - 1, max(0, (int)floor(((double)p.y - oy) * inv_cell)));
  int iz = min(dz - 1, max(0, (int)floor(((double)p.z - oz) * inv_cell)));
  keys[i] = (KeyT)morton3((u32)ix, (u32)iy, (u32)iz);
  vals[i] = i;
}
__global__ void widen_keys_kernel(const u32* __restrict__ k32, u32 m, u64* __restrict__ k64) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) k64[i] = (u64)k32[i];
}

// sorted gather: float4 {x,y,z,bits(orig)} + inverse permutation + per-point
// completeness bound of the one-ring search: every point outside the 3x3x3 cell
// block of a query is farther than (1 + min(f, 1-f)) * cell, f = fractional
// position inside the own cell (same binary64 expression as the binning).
__global__ void gather_sorted_kernel(const float4* __restrict__ pts, const u32* __restrict__ perm, u32 n,
                                     double ox, double oy, double oz, double inv_cell, double cell,
                                     int dx, int dy, int dz, float4* __restrict__ spts,
                                     u32* __restrict__ inv, float* __restrict__ qb2) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  u32 o = perm[j];
  float4 p = pts[o];
  double vx = ((double)p.x - ox) * inv_cell, vy = ((double)p.y - oy) * inv_cell,
         vz = ((double)p.z - oz) * inv_cell;
  double fx = vx - (double)min(dx - 1, max(0, (int)floor(vx)));
  double fy = vy - (double)min(dy - 1, max(0, (int)floor(vy)));
  double fz = vz - (double)min(dz - 1, max(0, (int)floor(vz)));
  double fm = fmin(fmin(fmin(fx, 1.0 - fx), fmin(fy, 1.0 - fy)), fmin(fz, 1.0 - fz));
  fm = fmax(fm, 0.0);
  double b = (1.0 + fm) * cell * (1.0 - 1e-6);
  qb2[j] = __double2float_rd(b * b);
  p.w = __uint_as_float(o);
  spts[j] = p;
  inv[o] = j;
}

__global__ void hash_build_kernel(const u64* __restrict__ cell_key, u32 n_cells, u64* __restrict__ hkeys,
                                  u32* __restrict__ hvals, u32 hmask) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  u64 key = cell_key[i];
  u32 slot = hash64(key) & hmask;
  while (true) {
    u64 prev = atomicCAS((unsigned long long*)&hkeys[slot], (unsigned long long)IFTWT_EMPTY_KEY,
                         (unsigned long long)key);
    if (prev == IFTWT_EMPTY_KEY || prev == key) {
      hvals[slot] = i;
      return;
    }
    slot = (slot + 1) & hmask;
  }
}

__global__ void set_u32_kernel(u32* p, u32 v) { *p = v; }
// the device scratch words of a grid build: bbox accumulators (ordered ints) + zeroed counters
__global__ void scalars_reset_kernel(int* sc) {
  const int i = threadIdx.x;
  sc[i] = i < 3 ? 0x7FFFFFFF : (i < 6 ? (int)0x80000000 : 0);
}

// ---------------------------------------------------------------------------
static double env_double(const char* name, double dflt) {
  const char* s = getenv(name);
  return s ? atof(s) : dflt;
}

int grid_build(iftwt_ctx* c, int k) {
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  // a re-grid re-sorts the cloud, which would orphan every sorted-space result; the
  // k-NN general path keeps any k exact on any grid, so keep the grid once results hang off it
  if (c->grid_valid && (c->grid_k == k || c->graph_k || c->nrm_k || c->seeds_valid)) return IFTWT_OK;
  StageTimer st(c, IFTWT_STAGE_GRID);
  const u32 n = (u32)c->n;
  const float4* pts = c->pts_in.as<float4>();
  ENSURE(c, c->scalars, 4096);
  int* sc = c->scalars.as<int>();
  int* hs = (int*)c->h_scalars;

  // --- bounding box ---
  LAUNCH(c, scalars_reset_kernel, 1, 128, 0, sc);
  LAUNCH(c, bbox_kernel, c->sm_count * 8, 256, 0, pts, n, sc);
  CHECK_LAUNCH(c);
  TRY(host_fetch(c, hs, sc, 8 * sizeof(int)));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (hs[6] != 0)
    return ctx_fail(c, IFTWT_ERR_NONFINITE, "cloud holds %d non-finite points", hs[6]);
  double lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = ord2f(hs[a]);
    hi[a] = ord2f(hs[3 + a]);
  }
  for (int a = 0; a < 3; ++a) {
    c->bb_lo[a] = lo[a];
    c->bb_hi[a] = hi[a];
  }
  double ext = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  if (!(ext > 0.0)) ext = 1.0;

  // --- cell size: points-per-occupied-cell at every octree level of the bounding cube ---
  double cell = env_double("IFTWT_CELL_SIZE", 0.0);
  if (!(cell > 0.0)) {
    ENSURE(c, c->keys, (size_t)n * 8);
    ENSURE(c, c->keys_alt, (size_t)n * 8);
    u64* k0 = c->keys.as<u64>();
    u64* k1 = c->keys_alt.as<u64>();
    double side = ext * (1.0 + 1e-6);
    // every sub-th point is enough: a cell holding m points shows a sampled point (m - 1) / sub
    // sampled companions on average, so m is estimated by 1 + sub * (run - 1).  (Sorting all
    // 10 M keys for this estimate cost 0.6 ms.)
    // (clouds below 2 M points used to sort every key: a 2 M-point cloud of config 5 paid 0.15 ms for it; the
    // sample stays above 2^18 points, four per sampled cell range at the coarsest level that matters)
    const u32 sub = n >= (1u << 21) ? 8u : (n >= (1u << 19) ? n >> 18 : 1u);
    const u32 m = (n + sub - 1) / sub;
    LAUNCH(c, fine_key_kernel, div_up(m, 256), 256, 0, pts, m, sub, lo[0], lo[1], lo[2], 65536.0 / side, k0);
    CHECK_LAUNCH(c);
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, (int)m, 0, 48, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortKeys(c->tmp.p, tb, k0, k1, (int)m, 0, 48, c->stream));
    c->stats.kernel_launches += 8;  // CUB onesweep: histogram + 6 passes (+ scan)
    unsigned long long* sums = (unsigned long long*)(sc + 80);  // 17 x u64, 8-byte aligned
    u32 stride = m / 65536u;
    if (stride == 0) stride = 1;
    u32 n_samples = (m + stride - 1) / stride;
    ENSURE(c, c->occ_hist, 17 * OCC_BINS * 4);
    CU_TRY(c, cudaMemsetAsync(c->occ_hist.p, 0, 17 * OCC_BINS * 4, c->stream));
    LAUNCH(c, occupancy_kernel, div_up(n_samples, 128), 128, 0, k1, m, stride, n_samples, sums, c->occ_hist.as<u32>());
    CHECK_LAUNCH(c);
    TRY(host_fetch(c, hs + 80, sums, 17 * 8));
    TRY(host_fetch(c, c->h_hist, c->occ_hist.p, 17 * OCC_BINS * 4));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const unsigned long long* hh = (const unsigned long long*)(hs + 80);
    double ppc[17];
    ppc[0] = (double)n;
    for (int l = 1; l <= 16; ++l)
      ppc[l] = fmax(1.0, 1.0 + (double)sub * ((double)hh[l] / (double)n_samples - 1.0));
    // a Poisson-filled cell of mean occupancy m is seen by its own points as m + 1;
    // m = 0.45 k puts the k-th neighbour at about cell / 1.2 on a surface (swept on B200)
    double beta = env_double("IFTWT_PPC_PER_K", 0.45);
    double target = beta * (double)(k < 2 ? 2 : k) + 1.0;
    c->stats_density_spread = 1.0;
    if (!getenv("IFTWT_PPC_PER_K") && env_double("IFTWT_PPC_ADAPT", 1.0) != 0.0 && n_samples >= 8192) {
      // Density spread.  The mean above is carried by the DENSE parts of a cloud; where the density is a
      // third of it (the ground plane of BASELINE configs 2 / 5 against its cylinder and sphere) the k-th
      // neighbour lies outside the 3x3x3 block and the query leaves the fast path: 17.7 % of cfg5's queries,
      // 0.9 of its 2.0 ms of k-NN.  One level above the cell scale — where a cell holds enough sampled points
      // for the count to mean something — the ratio of the mean occupancy to its 25th percentile (both
      // point-weighted), with the Poisson scatter of the sample taken out, is ~1.05 on the facade, ~1.8 on the
      // primitives scene.  At 1.25 or more the cell is sized for the sparse quarter instead: the occupancy
      // target grows by that ratio (at most 2).  Swept on B200 (cfg5): beta 0.45 -> 5.35 ms per cloud, 0.7 -> 4.43, 0.85 -> 4.33,
      // 1.0 -> 4.46; cfg3 keeps 0.45.  Any cell size gives the same, exact rows.
      int l = 0;
      while (l < 16 && ppc[l + 1] >= target) ++l;
      int L2 = l - 1;
      // sampled points per cell at that level: below 16 the 25th percentile is mostly Poisson noise — one level up
      if (L2 > 1 && (ppc[L2] - 1.0) / (double)sub + 1.0 < 16.0) --L2;
      if (L2 >= 1 && ppc[L2] >= 8.0) {
        const u32* hh2 = (const u32*)c->h_hist + (size_t)L2 * OCC_BINS;
        unsigned long long tot = 0;
        for (int b = 0; b < OCC_BINS; ++b) tot += hh2[b];
        if (tot >= 1024) {
          unsigned long long acc = 0;
          int b25 = 0;
          for (; b25 < OCC_BINS; ++b25) {
            acc += hh2[b25];
            if (4 * acc >= tot) break;
          }
          const double run25 = exp2(((double)b25 + 0.5) / (double)OCC_BINS_PER_OCTAVE);
          const double occ25 = fmax(1.0, 1.0 + (double)sub * (run25 - 1.0));
          // a uniform cloud already shows mean / p25 = 1 / (1 - 0.674 / sqrt(lambda)) from the Poisson scatter of
          // the lambda sampled points per cell: taken out, so that the ratio measures the cloud, not the sample
          const double lambda = fmax(1.0, (ppc[L2] - 1.0) / (double)sub + 1.0);
          const double spread = ppc[L2] / occ25 * fmax(0.5, 1.0 - 0.674 / sqrt(lambda));
          c->stats_density_spread = spread;
          if (spread >= 1.25) {
            beta *= fmin(spread, 2.0);
            target = beta * (double)(k < 2 ? 2 : k) + 1.0;
          }
        }
      }
    }
    if (ppc[0] <= target) {
      cell = side;
    } else {
      int l = 0;
      while (l < 16 && ppc[l + 1] >= target) ++l;
      double s_l = side / (double)(1 << l);
      if (l == 16) {
        cell = s_l;
      } else {
        double D = log2(ppc[l] / ppc[l + 1]);
        D = fmin(3.0, fmax(1.0, D));
        cell = s_l * exp2(-log2(ppc[l] / target) / D);
      }
    }
  }
  cell = fmax(cell, ext / 1048000.0);
  int dim[3];
  for (int a = 0; a < 3; ++a) dim[a] = (int)floor((hi[a] - lo[a]) / cell) + 1;

  // --- Morton sort by cell ---
  int maxdim = dim[0] > dim[1] ? dim[0] : dim[1];
  if (dim[2] > maxdim) maxdim = dim[2];
  int bits = 1;
  while ((1 << bits) < maxdim) ++bits;
  ENSURE(c, c->keys, (size_t)n * 8);
  ENSURE(c, c->keys_alt, (size_t)n * 8);
  ENSURE(c, c->vals_alt, (size_t)n * 4);
  ENSURE(c, c->perm, (size_t)n * 4);
  ENSURE(c, c->inv, (size_t)n * 4);
  ENSURE(c, c->spts, (size_t)n * 16);
  u64* k0 = c->keys.as<u64>();
  u64* k1 = c->keys_alt.as<u64>();
  u32* v0 = c->vals_alt.as<u32>();
  u32* perm = c->perm.as<u32>();
  const bool narrow = 3 * bits <= 32;  // cfg3: 30 bits
  u32* k0n = (u32*)k0;
  u32* k1n = (u32*)k1;
  if (narrow) {
    LAUNCH(c, cell_key_kernel<u32>, div_up(n, 256), 256, 0, pts, n, lo[0], lo[1], lo[2], 1.0 / cell, dim[0],
           dim[1], dim[2], k0n, v0);
    CHECK_LAUNCH(c);
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, k0n, k1n, v0, perm, (int)n, 0, 3 * bits, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k0n, k1n, v0, perm, (int)n, 0, 3 * bits, c->stream));
  } else {
    LAUNCH(c, cell_key_kernel<u64>, div_up(n, 256), 256, 0, pts, n, lo[0], lo[1], lo[2], 1.0 / cell, dim[0],
           dim[1], dim[2], k0, v0);
    CHECK_LAUNCH(c);
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, perm, (int)n, 0, 3 * bits, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k0, k1, v0, perm, (int)n, 0, 3 * bits, c->stream));
  }
  c->stats.kernel_launches += 2 + (3 * bits + 7) / 8;
  ENSURE(c, c->qb2, (size_t)n * 4);
  LAUNCH(c, gather_sorted_kernel, div_up(n, 256), 256, 0, pts, perm, n, lo[0], lo[1], lo[2], 1.0 / cell,
         cell, dim[0], dim[1], dim[2], c->spts.as<float4>(), c->inv.as<u32>(), c->qb2.as<float>());
  CHECK_LAUNCH(c);

  // --- occupied cells: run-length encode the sorted keys ---
  ENSURE(c, c->cell_key, (size_t)n * 8);
  ENSURE(c, c->cell_cnt, (size_t)n * 4);
  ENSURE(c, c->cell_start, ((size_t)n + 1) * 4);
  u32* d_nruns = (u32*)(sc + 40);
  if (narrow) {  // unique 32-bit keys land in the (now free) unsorted key buffer and are widened below
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRunLengthEncode::Encode(nullptr, tb, k1n, k0n, c->cell_cnt.as<u32>(), d_nruns, (int)n,
                                                 c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRunLengthEncode::Encode(c->tmp.p, tb, k1n, k0n, c->cell_cnt.as<u32>(), d_nruns, (int)n,
                                                 c->stream));
    c->stats.kernel_launches += 2;
  } else {
    size_t tb = 0;
    CU_TRY(c, cub::DeviceRunLengthEncode::Encode(nullptr, tb, k1, c->cell_key.as<u64>(),
                                                 c->cell_cnt.as<u32>(), d_nruns, (int)n, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceRunLengthEncode::Encode(c->tmp.p, tb, k1, c->cell_key.as<u64>(),
                                                 c->cell_cnt.as<u32>(), d_nruns, (int)n, c->stream));
    c->stats.kernel_launches += 2;
  }
  TRY(host_fetch(c, hs + 40, d_nruns, sizeof(u32)));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  u32 n_cells = (u32)hs[40];
  if (n_cells == 0 || n_cells > n) return ctx_fail(c, IFTWT_ERR_CUDA, "cell table build failed");
  if (narrow) {
    LAUNCH(c, widen_keys_kernel, div_up(n_cells, 256), 256, 0, (const u32*)k0n, n_cells, c->cell_key.as<u64>());
    CHECK_LAUNCH(c);
  }
  LAUNCH(c, set_u32_kernel, 1, 1, 0, c->cell_start.as<u32>(), 0u);
  {
    size_t tb = 0;
    CU_TRY(c, cub::DeviceScan::InclusiveSum(nullptr, tb, c->cell_cnt.as<u32>(), c->cell_start.as<u32>() + 1,
                                            (int)n_cells, c->stream));
    ENSURE(c, c->tmp, tb);
    CU_TRY(c, cub::DeviceScan::InclusiveSum(c->tmp.p, tb, c->cell_cnt.as<u32>(),
                                            c->cell_start.as<u32>() + 1, (int)n_cells, c->stream));
    c->stats.kernel_launches += 2;
  }
  // --- cell hash ---
  u32 hsize = 1024;
  while (hsize < 2 * n_cells) hsize <<= 1;
  ENSURE(c, c->hkeys, (size_t)hsize * 8);
  ENSURE(c, c->hvals, (size_t)hsize * 4);
  CU_TRY(c, cudaMemsetAsync(c->hkeys.p, 0xFF, (size_t)hsize * 8, c->stream));
  LAUNCH(c, hash_build_kernel, div_up(n_cells, 256), 256, 0, c->cell_key.as<u64>(), n_cells,
         c->hkeys.as<u64>(), c->hvals.as<u32>(), hsize - 1);
  CHECK_LAUNCH(c);

  GridView& g = c->grid;
  g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
  g.cell = cell;
  g.inv_cell = 1.0 / cell;
  g.dimx = dim[0]; g.dimy = dim[1]; g.dimz = dim[2];
  g.n_points = n;
  g.n_cells = n_cells;
  g.cell_key = c->cell_key.as<u64>();
  g.cell_start = c->cell_start.as<u32>();
  g.hkeys = c->hkeys.as<u64>();
  g.hvals = c->hvals.as<u32>();
  g.hmask = hsize - 1;
  c->grid_valid = true;
  c->grid_k = k;
  c->stats.n_cells = n_cells;
  c->stats.cell_size = cell;
  // everything derived from an older grid is stale
  c->knn_k = 0;
  c->graph_k = 0;
  c->graph_kind = 0;
  c->graph_shares_nbr = false;
  c->mask_k = 0;
  c->nrm_k = 0;
  c->seeds_valid = false;
  c->ift_valid = false;
  return IFTWT_OK;
}
