// common.cuh — context, device buffers, launch bookkeeping shared by every stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/iftwt.h"

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define IFTWT_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

// Grow-only device buffer owned by a ctx.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T>
  T* as() const { return (T*)p; }
};

// Uniform grid over the Morton-sorted cloud (device view, passed by value).
struct GridView {
  double ox, oy, oz;   // grid origin (bbox min)
  double inv_cell;     // 1 / cell size
  double cell;         // cell size
  int dimx, dimy, dimz;
  u32 n_points;
  u32 n_cells;
  const u64* cell_key;    // [n_cells] Morton key of each occupied cell, ascending
  const u32* cell_start;  // [n_cells + 1] first sorted point of each cell
  const u64* hkeys;       // open-addressing hash: cell key -> cell index
  const u32* hvals;
  u32 hmask;
};

struct iftwt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;

  // --- cloud ---
  size_t n = 0;
  DevBuf pts_in;   // float4 [n] original order {x,y,z,intensity}
  DevBuf stage_in; // raw host records when the input is not packed float4
  DevBuf spts;     // float4 [n] Morton order {x,y,z,bits(orig index)}
  DevBuf perm;     // u32 [n] sorted -> orig
  DevBuf inv;      // u32 [n] orig -> sorted
  DevBuf keys, keys_alt, vals_alt;  // sort scratch
  DevBuf tmp;      // CUB temp storage
  DevBuf scalars;  // small device scratch (counters, bbox, histograms)
  void* h_scalars = nullptr;  // pinned mirror

  // --- grid ---
  double bb_lo[3] = {0, 0, 0}, bb_hi[3] = {0, 0, 0};  // float bounding box of the cloud
  bool grid_valid = false;
  int grid_k = 0;
  GridView grid{};
  DevBuf cell_key, cell_cnt, cell_start, hkeys, hvals;
  DevBuf qb2;      // float [n] squared one-ring completeness bound per sorted point
  DevBuf occ_hist; // u32 [17][256] run-length histogram of the occupancy estimate (grid.cu)
  void* h_hist = nullptr;  // pinned mirror
  double stats_density_spread = 1.0;  // mean / 25th-percentile occupancy of the last grid build

  // --- knn ---
  int knn_k = 0;   // 0 = none
  DevBuf nbr;      // u32 [n*k] sorted space
  DevBuf nd2;      // float [n*k]
  DevBuf kth;      // u64 [n] key of the k-th neighbour (d2 bits << 32 | orig idx)
  DevBuf fb_list;  // u32 [n] fallback query list
  // --- graph ---
  int graph_k = 0;
  int graph_kind = 0;             // 0 none, 1 symmetric k-NN (builder B), 2 voxel-26 (builder A)
  size_t n_arcs = 0;         // forward (valid, non-self) + reverse arcs
  size_t n_rev = 0;
  bool graph_shares_nbr = false;  // forward rows alias `nbr` (else they live in g_fwd)
  DevBuf g_off, g_adj, g_deg;     // reverse-arc CSR: offsets, targets, degree / fill cursor
  DevBuf g_fwd;                   // private copy of the forward rows once nbr is recomputed
  int mask_k = 0;                 // ow_mask describes the current k-NN lists when == knn_k
  DevBuf ow_mask;                 // u64 [n] bit j: arc u -> nbr[u][j] is one-way
  int rdeg_k = 0;                 // g_deg holds reverse degrees + n_fwd of the current lists when == knn_k
  bool rg_pre_valid = false;      // the fused pass left the region-growing union-find + arc list
  int rg_pre_k = 0;
  float rg_pre_cos = 0.f;
  // --- normals ---
  int nrm_k = 0;
  DevBuf nrm;  // float4 [n] sorted space {nx,ny,nz,curv}
  // --- seeds ---
  DevBuf rg_rank, rg_parent, rg_label, rg_aux, rg_aux2, rg_masks;
  DevBuf seeds;  // u32 [n_seeds] orig indices
  size_t n_seeds = 0;
  bool seeds_valid = false;
  // --- ift ---
  DevBuf ift_w, ift_byw;
  DevBuf ift_node; // u64 [n] union-find link | result | weight in one word (ift.cu)
  DevBuf ift_res;  // u64 [n] resolved (cost << 32 | label) per sorted vertex
  DevBuf ift_wf;   // float [n] the weights the IFT ran on, sorted space
  // --- hierarchy ---
  DevBuf h_keys, h_vals, h_pairs, h_seeds, h_info;
  size_t n_pairs = 0, n_regions = 0, pairs_vals_offset = 0;
  bool rag_valid = false;
  bool ift_valid = false;
  bool ift_res_valid = false;  // ift_res holds the words of the current run (built on demand: ift_res_ensure)
  // --- voxel graph (builder A) ---
  DevBuf vx_keys, vx_rows, vx_cent, vx_hk, vx_hv;
  double voxel_size = 0.0;
  // --- export scratch ---
  DevBuf out_a, out_b;

  // --- instrumentation ---
  cudaEvent_t ev0[IFTWT_STAGE_COUNT], ev1[IFTWT_STAGE_COUNT];
  bool ev_set[IFTWT_STAGE_COUNT];
  iftwt_stats stats{};
};

int ctx_fail(iftwt_ctx* c, int code, const char* fmt, ...);
int dev_ensure(iftwt_ctx* c, DevBuf& b, size_t bytes);
// device words -> the ctx's mapped pinned mirrors (h_scalars / h_hist), written by a kernel: no copy engine
int host_fetch(iftwt_ctx* c, void* h_dst, const void* d_src, size_t bytes);

#define CU_TRY(ctx, expr)                                                             \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ctx_fail((ctx), IFTWT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,          \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)
#define TRY(expr)              \
  do {                         \
    int _r = (expr);           \
    if (_r != IFTWT_OK) return _r; \
  } while (0)
#define ENSURE(ctx, buf, bytes) TRY(dev_ensure((ctx), (buf), (bytes)))

// every kernel launch goes through this so gpu_launches is a count, not a guess
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
  do {                                                                      \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
    (ctx)->stats.kernel_launches++;                                         \
  } while (0)
#define CHECK_LAUNCH(ctx) CU_TRY(ctx, cudaGetLastError())
// Programmatic dependent launch: the kernel may become resident while its predecessor in the
// stream is still running; everything it does before pdl_wait() must not depend on it.
#define LAUNCH_PDL(ctx, kernel, grid, block, smem, ...)                               \
  do {                                                                                \
    cudaLaunchConfig_t cfg_ = {};                                                     \
    cfg_.gridDim = dim3(grid);                                                        \
    cfg_.blockDim = dim3(block);                                                      \
    cfg_.dynamicSmemBytes = (smem);                                                   \
    cfg_.stream = (ctx)->stream;                                                      \
    cudaLaunchAttribute at_[1];                                                       \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                   \
    at_[0].val.programmaticStreamSerializationAllowed = 1;                            \
    cfg_.attrs = at_;                                                                 \
    cfg_.numAttrs = 1;                                                                \
    cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                   \
    (ctx)->stats.kernel_launches++;                                                   \
  } while (0)
#ifdef __CUDACC__
// lets the next kernel of the stream start its independent prologue
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
// returns when every earlier kernel of the stream has completed and its writes are visible
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

struct StageTimer {
  iftwt_ctx* c;
  int s;
  StageTimer(iftwt_ctx* c_, int s_) : c(c_), s(s_) { cudaEventRecord(c->ev0[s], c->stream); }
  ~StageTimer() {
    cudaEventRecord(c->ev1[s], c->stream);
    c->ev_set[s] = true;
  }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- stage entry points (host side), one per .cu ----
int grid_build(iftwt_ctx* c, int k);                                       // grid.cu
int knn_run(iftwt_ctx* c, int k);                                          // knn.cu
int knn_query_run(iftwt_ctx* c, const float4* d_q, size_t m, int k, u32* d_idx_sorted, float* d_d2);  // knn.cu
int graph_build(iftwt_ctx* c);                                             // graph.cu
int knn_mutual_run(iftwt_ctx* c, bool for_graph);
int graph_detach_rows(iftwt_ctx* c);
int cloud_resolution_run(iftwt_ctx* c, double overlap, double* res);            // voxel.cu
int voxel_graph_run(iftwt_ctx* c, double resolution, int k_hint);
int morph_run(iftwt_ctx* c, int op, float* d_out_orig);
int regmin_run(iftwt_ctx* c);
int morph_rgb_run(iftwt_ctx* c, int op, u32* d_out_orig);
int rgb_to_gray_run(iftwt_ctx* c, int otsu, double* threshold);
int region_adjacency_run(iftwt_ctx* c);                                     // hierarchy.cu
int region_info_run(iftwt_ctx* c);
int cluster_run(iftwt_ctx* c, int num_regions, u32* d_cluster_orig, size_t* n_clusters);
void ctx_invalidate(iftwt_ctx* c);
int graph_export(iftwt_ctx* c, u32* h_off, u32* h_adj, size_t cap, size_t* n_arcs);
int normals_run(iftwt_ctx* c);                                             // normals.cu
int seeds_run(iftwt_ctx* c, const iftwt_rg_params* p);                     // seeds.cu
int fused_graph_rg_run(iftwt_ctx* c, const iftwt_rg_params* p);
int ift_run(iftwt_ctx* c, const float* d_w_orig /*or null*/, const u32* d_seeds_orig, size_t ns);  // ift.cu
int ift_res_ensure(iftwt_ctx* c);

// ---- device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ u64 morton_spread(u32 v) {  // 21 bits -> every third bit
  u64 x = v & 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ u32 morton_compact(u64 x) {
  x &= 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10C30C30C30C30C3ull;
  x = (x ^ (x >> 4)) & 0x100F00F00F00F00Full;
  x = (x ^ (x >> 8)) & 0x1F0000FF0000FFull;
  x = (x ^ (x >> 16)) & 0x1F00000000FFFFull;
  x = (x ^ (x >> 32)) & 0x1FFFFFull;
  return (u32)x;
}
// x is the most significant axis bit of every triple
__device__ __forceinline__ u64 morton3(u32 x, u32 y, u32 z) {
  return (morton_spread(x) << 2) | (morton_spread(y) << 1) | morton_spread(z);
}
__device__ __forceinline__ void demorton3(u64 k, u32& x, u32& y, u32& z) {
  x = morton_compact(k >> 2);
  y = morton_compact(k >> 1);
  z = morton_compact(k);
}
__device__ __forceinline__ u32 hash64(u64 k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (u32)k;
}
// returns cell index or 0xFFFFFFFF
__device__ __forceinline__ u32 grid_lookup(const GridView& g, u64 key) {
  u32 slot = hash64(key) & g.hmask;
  while (true) {
    u64 k = g.hkeys[slot];
    if (k == key) return g.hvals[slot];
    if (k == IFTWT_EMPTY_KEY) return 0xFFFFFFFFu;
    slot = (slot + 1) & g.hmask;
  }
}
// FLANN L2_Simple<float> accumulation order, no contraction (oracle: dist2()).
__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}
#endif
