// capi.cu — the extern "C" boundary declared in include/iftwt.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

static thread_local std::string g_err;  // for failures before a ctx exists

namespace {

// strided host records -> packed float4 {x,y,z,intensity}
__global__ void repack_kernel(const unsigned char* __restrict__ raw, size_t stride, u32 n, int xyz_off, int i_off,
                              float4* __restrict__ out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* r = raw + (size_t)i * stride;
  const float* xyz = (const float*)(r + xyz_off);
  float w = i_off >= 0 ? *(const float*)(r + i_off) : 0.f;
  out[i] = make_float4(xyz[0], xyz[1], xyz[2], w);
}
__global__ void pack_queries_kernel(const unsigned char* __restrict__ raw, size_t stride, u32 m,
                                    float4* __restrict__ out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float* xyz = (const float*)(raw + (size_t)i * stride);
  out[i] = make_float4(xyz[0], xyz[1], xyz[2], 0.f);
}
// k-NN rows: sorted space -> original space
__global__ void export_knn_kernel(const u32* __restrict__ nbr, const float* __restrict__ nd2,
                                  const u32* __restrict__ inv, const u32* __restrict__ perm, size_t n_entries,
                                  int k, int* __restrict__ idx_o, float* __restrict__ d2_o) {
  size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_entries) return;
  u32 i = (u32)(a / k);
  int j = (int)(a - (size_t)i * k);
  size_t src = (size_t)inv[i] * k + j;
  u32 s = nbr[src];
  idx_o[a] = s == SID_NONE ? -1 : (int)perm[s];
  if (d2_o) d2_o[a] = nd2[src];
}
__global__ void export_query_kernel(const u32* __restrict__ sid, const u32* __restrict__ perm, size_t n_entries,
                                    int* __restrict__ idx_o) {
  size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_entries) return;
  u32 s = sid[a];
  idx_o[a] = s == SID_NONE ? -1 : (int)perm[s];
}
__global__ void export_normals_kernel(const float4* __restrict__ nrm, const u32* __restrict__ inv, u32 n,
                                      float4* __restrict__ out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = nrm[inv[i]];
}


__global__ void set_intensity_kernel(const float* __restrict__ v, u32 n, float4* __restrict__ pts) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pts[i].w = v[i];
}

// externally computed rows (original ids) -> the ctx's sorted-space lists
__global__ void install_rows_kernel(const int* __restrict__ idx_o, const float* __restrict__ d2_o,
                                    const u32* __restrict__ perm, const u32* __restrict__ inv, u32 n, int k,
                                    u32* __restrict__ nbr, float* __restrict__ nd2, u64* __restrict__ kth) {
  const u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n) return;
  const u32 o = perm[s];
  for (int j = lane; j < k; j += 32) {
    int v = idx_o[(size_t)o * k + j];
    float d = d2_o[(size_t)o * k + j];
    nbr[(size_t)s * k + j] = v < 0 ? SID_NONE : inv[v];
    nd2[(size_t)s * k + j] = d;
    if (j == k - 1) kth[s] = v < 0 ? 0xFFFFFFFFFFFFFFFFull : (((u64)__float_as_uint(d) << 32) | (u64)(u32)v);
  }
}

int check_ctx(iftwt_ctx* c) {
  if (!c) {
    g_err = "null ctx";
    return IFTWT_ERR_INVALID;
  }
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return ctx_fail(c, IFTWT_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return IFTWT_OK;
}

}  // namespace

void ctx_invalidate(iftwt_ctx* c) {
  c->graph_kind = 0;
  c->grid_valid = false;
  c->grid_k = 0;
  c->knn_k = 0;
  c->graph_k = 0;
  c->graph_shares_nbr = false;
  c->mask_k = 0;
  c->rdeg_k = 0;
  c->rg_pre_valid = false;
  c->nrm_k = 0;
  c->seeds_valid = false;
  c->ift_valid = false;
  c->n_seeds = 0;
  c->n_arcs = 0;
  c->n_rev = 0;
  c->rag_valid = false;
}

// Small device -> host reads (counters, the level table) are written by a kernel straight into the
// ctx's mapped pinned mirrors instead of going through cudaMemcpyAsync: a 4-byte copy queues on the same
// copy engine as the 80 MB result download (or the 160 MB upload) of ANOTHER ctx and waits for it, which
// stalls this ctx's host thread for up to a millisecond at every round trip when several clouds are in
// flight.  IFTWT_FETCH=copy restores the copies (A/B measurements).
__global__ void fetch_words_kernel(u32* __restrict__ dst, const u32* __restrict__ src, u32 nw) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
int host_fetch(iftwt_ctx* c, void* h_dst, const void* d_src, size_t bytes) {
  static const bool by_copy = getenv("IFTWT_FETCH") && !strcmp(getenv("IFTWT_FETCH"), "copy");
  if (by_copy || (bytes & 3))
    CU_TRY(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  else {
    const u32 nw = (u32)(bytes / 4);
    LAUNCH(c, fetch_words_kernel, nw > 1024 ? 4 : 1, nw >= 256 ? 256 : 32, 0, (u32*)h_dst, (const u32*)d_src, nw);
    CHECK_LAUNCH(c);
  }
  return IFTWT_OK;
}

extern "C" {

int iftwt_version(void) { return IFTWT_VERSION; }

const char* iftwt_last_error(const iftwt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int iftwt_create(iftwt_ctx** out, int device) {
  if (!out) {
    g_err = "null out pointer";
    return IFTWT_ERR_INVALID;
  }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return IFTWT_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_err = "device ordinal out of range";
    return IFTWT_ERR_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_err = cudaGetErrorString(e);
    return IFTWT_ERR_CUDA;
  }
  if (prop.major != 10) {
    char b[160];
    snprintf(b, sizeof b, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
             prop.minor);
    g_err = b;
    return IFTWT_ERR_CUDA;
  }
  iftwt_ctx* c = new (std::nothrow) iftwt_ctx();
  if (!c) {
    g_err = "out of host memory";
    return IFTWT_ERR_CUDA;
  }
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaHostAlloc(&c->h_scalars, 4096, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostAlloc(&c->h_hist, 17 * 256 * 4, cudaHostAllocMapped) != cudaSuccess) {
    g_err = std::string("ctx setup failed: ") + cudaGetErrorString(cudaGetLastError());
    delete c;
    return IFTWT_ERR_CUDA;
  }
  for (int s = 0; s < IFTWT_STAGE_COUNT; ++s) {
    cudaEventCreate(&c->ev0[s]);
    cudaEventCreate(&c->ev1[s]);
    c->ev_set[s] = false;
  }
  memset(&c->stats, 0, sizeof c->stats);
  *out = c;
  return IFTWT_OK;
}

void iftwt_destroy(iftwt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->pts_in, &c->stage_in, &c->spts, &c->perm, &c->inv, &c->keys, &c->keys_alt, &c->vals_alt,
                    &c->tmp, &c->scalars, &c->cell_key, &c->cell_cnt, &c->cell_start, &c->hkeys, &c->hvals,
                    &c->qb2, &c->nbr, &c->nd2, &c->kth, &c->fb_list, &c->g_off, &c->g_adj, &c->g_deg, &c->g_fwd,
                    &c->ow_mask, &c->nrm, &c->vx_keys, &c->vx_rows, &c->vx_cent, &c->vx_hk, &c->vx_hv, &c->ift_res, &c->ift_wf, &c->ift_node, &c->h_keys, &c->h_vals, &c->h_pairs, &c->h_seeds, &c->h_info,
                    &c->rg_rank, &c->rg_parent, &c->rg_label, &c->rg_aux, &c->rg_aux2, &c->rg_masks, &c->seeds, &c->ift_w,
                    &c->ift_byw, &c->out_a, &c->out_b, &c->occ_hist};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  if (c->h_hist) cudaFreeHost(c->h_hist);
  for (int s = 0; s < IFTWT_STAGE_COUNT; ++s) {
    cudaEventDestroy(c->ev0[s]);
    cudaEventDestroy(c->ev1[s]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}


int iftwt_set_cloud(iftwt_ctx* c, const void* points, size_t stride, size_t n, int xyz_off, int i_off) {
  TRY(check_ctx(c));
  if (!points || n == 0 || n >= 0x7FFFFFF0ull || stride < 12 || xyz_off < 0 || (size_t)xyz_off + 12 > stride ||
      (i_off >= 0 && (size_t)i_off + 4 > stride) || (xyz_off & 3) || (i_off >= 0 && (i_off & 3)) || (stride & 3))
    return ctx_fail(c, IFTWT_ERR_INVALID, "bad cloud description (n=%zu stride=%zu xyz_off=%d i_off=%d)", n,
                    stride, xyz_off, i_off);
  ctx_invalidate(c);
  ENSURE(c, c->scalars, 4096);
  ENSURE(c, c->pts_in, n * 16);
  StageTimer st(c, IFTWT_STAGE_UPLOAD);
  if (stride == 16 && xyz_off == 0 && i_off == 12) {
    CU_TRY(c, cudaMemcpyAsync(c->pts_in.p, points, n * 16, cudaMemcpyHostToDevice, c->stream));
  } else {
    ENSURE(c, c->stage_in, n * stride);
    CU_TRY(c, cudaMemcpyAsync(c->stage_in.p, points, n * stride, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, repack_kernel, div_up(n, 256), 256, 0, c->stage_in.as<unsigned char>(), stride, (u32)n, xyz_off,
           i_off, c->pts_in.as<float4>());
    CHECK_LAUNCH(c);
  }
  c->n = n;
  c->stats.n_points = n;
  return IFTWT_OK;
}

int iftwt_set_cloud_device(iftwt_ctx* c, const void* device_float4, size_t n) {
  TRY(check_ctx(c));
  if (!device_float4 || n == 0 || n >= 0x7FFFFFF0ull) return ctx_fail(c, IFTWT_ERR_INVALID, "bad device cloud");
  ctx_invalidate(c);
  ENSURE(c, c->scalars, 4096);
  ENSURE(c, c->pts_in, n * 16);
  StageTimer st(c, IFTWT_STAGE_UPLOAD);
  CU_TRY(c, cudaMemcpyAsync(c->pts_in.p, device_float4, n * 16, cudaMemcpyDeviceToDevice, c->stream));
  c->n = n;
  c->stats.n_points = n;
  return IFTWT_OK;
}

static void note_fallback(iftwt_ctx* c) { c->stats.knn_fallback = ((u32*)c->h_scalars)[48]; }

int iftwt_knn(iftwt_ctx* c, int k, int32_t* idx, float* d2) {
  TRY(check_ctx(c));
  TRY(knn_run(c, k));
  const size_t ne = c->n * (size_t)k;
  if (idx) {
    StageTimer st(c, IFTWT_STAGE_DOWNLOAD);
    ENSURE(c, c->out_a, ne * 4);
    if (d2) ENSURE(c, c->out_b, ne * 4);
    LAUNCH(c, export_knn_kernel, div_up(ne, 256), 256, 0, c->nbr.as<u32>(), c->nd2.as<float>(), c->inv.as<u32>(),
           c->perm.as<u32>(), ne, k, c->out_a.as<int>(), d2 ? c->out_b.as<float>() : (float*)nullptr);
    CHECK_LAUNCH(c);
    CU_TRY(c, cudaMemcpyAsync(idx, c->out_a.p, ne * 4, cudaMemcpyDeviceToHost, c->stream));
    if (d2) CU_TRY(c, cudaMemcpyAsync(d2, c->out_b.p, ne * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  note_fallback(c);
  return IFTWT_OK;
}

int iftwt_knn_query(iftwt_ctx* c, const void* queries, size_t stride, size_t m, int k, int32_t* idx, float* d2) {
  TRY(check_ctx(c));
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  if (!queries || stride < 12 || (stride & 3) || m >= 0x7FFFFFF0ull)
    return ctx_fail(c, IFTWT_ERR_INVALID, "bad query description");
  if (m == 0) return IFTWT_OK;
  if (!c->grid_valid) TRY(grid_build(c, k < 8 ? 8 : k));
  const size_t ne = m * (size_t)k;
  ENSURE(c, c->stage_in, m * stride + m * 16 + 16);
  unsigned char* raw = c->stage_in.as<unsigned char>() + ((m * 16 + 15) & ~(size_t)15);
  float4* q4 = c->stage_in.as<float4>();
  CU_TRY(c, cudaMemcpyAsync(raw, queries, m * stride, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, pack_queries_kernel, div_up(m, 256), 256, 0, raw, stride, (u32)m, q4);
  ENSURE(c, c->out_a, ne * 4 * 2);
  ENSURE(c, c->out_b, ne * 4);
  u32* sid = c->out_a.as<u32>();
  int* idx_o = (int*)(sid + ne);
  TRY(knn_query_run(c, q4, m, k, sid, c->out_b.as<float>()));
  LAUNCH(c, export_query_kernel, div_up(ne, 256), 256, 0, sid, c->perm.as<u32>(), ne, idx_o);
  CHECK_LAUNCH(c);
  if (idx) CU_TRY(c, cudaMemcpyAsync(idx, idx_o, ne * 4, cudaMemcpyDeviceToHost, c->stream));
  if (d2) CU_TRY(c, cudaMemcpyAsync(d2, c->out_b.p, ne * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_knn_graph(iftwt_ctx* c, int k, uint32_t* off, uint32_t* adj, size_t cap, size_t* n_arcs) {
  TRY(check_ctx(c));
  TRY(knn_run(c, k));
  TRY(graph_build(c));
  if (off || adj) {
    TRY(graph_export(c, off, adj, cap, n_arcs));
  } else {
    if (n_arcs) *n_arcs = c->n_arcs;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
  }
  note_fallback(c);
  return IFTWT_OK;
}

int iftwt_normals(iftwt_ctx* c, int k, void* out, size_t stride, int normal_off, int curv_off) {
  TRY(check_ctx(c));
  TRY(knn_run(c, k));
  TRY(normals_run(c));
  if (out) {
    if (stride < 16 || normal_off < 0 || curv_off < 0 || (size_t)normal_off + 12 > stride ||
        (size_t)curv_off + 4 > stride)
      return ctx_fail(c, IFTWT_ERR_INVALID, "bad normal record description");
    StageTimer st(c, IFTWT_STAGE_DOWNLOAD);
    const u32 n = (u32)c->n;
    ENSURE(c, c->out_a, (size_t)n * 16);
    LAUNCH(c, export_normals_kernel, div_up(n, 256), 256, 0, c->nrm.as<float4>(), c->inv.as<u32>(), n,
           c->out_a.as<float4>());
    CHECK_LAUNCH(c);
    if (stride == 16 && normal_off == 0 && curv_off == 12) {
      CU_TRY(c, cudaMemcpyAsync(out, c->out_a.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    } else {
      // two strided copies: xyz (12 bytes) and curvature (4 bytes) per record
      CU_TRY(c, cudaMemcpy2DAsync((char*)out + normal_off, stride, c->out_a.p, 16, 12, n, cudaMemcpyDeviceToHost,
                                  c->stream));
      CU_TRY(c, cudaMemcpy2DAsync((char*)out + curv_off, stride, (char*)c->out_a.p + 12, 16, 4, n,
                                  cudaMemcpyDeviceToHost, c->stream));
    }
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  note_fallback(c);
  return IFTWT_OK;
}

int iftwt_region_seeds(iftwt_ctx* c, const iftwt_rg_params* p, int32_t* point_label, int32_t* seed_index,
                       size_t seed_cap, size_t* n_seeds) {
  TRY(check_ctx(c));
  if (!p) return ctx_fail(c, IFTWT_ERR_INVALID, "null params");
  TRY(seeds_run(c, p));
  if (n_seeds) *n_seeds = c->n_seeds;
  if (point_label)
    CU_TRY(c, cudaMemcpyAsync(point_label, c->out_a.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
  int rc = IFTWT_OK;
  if (seed_index) {
    if (seed_cap < c->n_seeds)
      rc = ctx_fail(c, IFTWT_ERR_CAPACITY, "seed buffer needs %zu entries, caller gave %zu", c->n_seeds, seed_cap);
    else if (c->n_seeds)
      CU_TRY(c, cudaMemcpyAsync(seed_index, c->seeds.p, c->n_seeds * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return rc;
}

int iftwt_ift(iftwt_ctx* c, const float* weights, const int32_t* seeds, size_t n_seeds, uint32_t* cost,
              uint32_t* label) {
  TRY(check_ctx(c));
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built (call iftwt_knn_graph first)");
  const size_t n = c->n;
  const float* d_w = nullptr;
  if (weights) {
    ENSURE(c, c->out_b, n * 4);
    CU_TRY(c, cudaMemcpyAsync(c->out_b.p, weights, n * 4, cudaMemcpyHostToDevice, c->stream));
    d_w = c->out_b.as<float>();
  }
  const u32* d_seeds;
  size_t ns;
  if (seeds) {
    ENSURE(c, c->seeds, (n_seeds + 1) * 4);
    if (n_seeds)
      CU_TRY(c, cudaMemcpyAsync(c->seeds.p, seeds, n_seeds * 4, cudaMemcpyHostToDevice, c->stream));
    c->n_seeds = n_seeds;
    c->seeds_valid = true;
    d_seeds = c->seeds.as<u32>();
    ns = n_seeds;
  } else {
    if (!c->seeds_valid) return ctx_fail(c, IFTWT_ERR_STATE, "no seeds on the device (call iftwt_region_seeds)");
    d_seeds = c->seeds.as<u32>();
    ns = c->n_seeds;
  }
  TRY(ift_run(c, d_w, d_seeds, ns));
  {
    StageTimer st(c, IFTWT_STAGE_DOWNLOAD);
    if (cost) CU_TRY(c, cudaMemcpyAsync(cost, c->out_a.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (label)
      CU_TRY(c, cudaMemcpyAsync(label, c->out_a.as<u32>() + n, n * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_segment(iftwt_ctx* c, const iftwt_segment_params* p, uint32_t* cost, uint32_t* label, size_t* n_seeds) {
  TRY(check_ctx(c));
  if (!p) return ctx_fail(c, IFTWT_ERR_INVALID, "null params");
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  if (p->weights != 0 && p->weights != 1) return ctx_fail(c, IFTWT_ERR_INVALID, "weights must be 0 or 1");
  int kmax = p->k_graph;
  if (p->k_normals > kmax) kmax = p->k_normals;
  if (p->rg.number_of_neighbours > kmax) kmax = p->rg.number_of_neighbours;
  TRY(grid_build(c, kmax));
  TRY(knn_run(c, p->k_graph));
  if (p->k_normals == p->k_graph && p->rg.number_of_neighbours == p->k_graph && p->rg.curvature_threshold >= 0.34f) {
    // one neighbourhood size: graph mutuality and the region-growing arcs share one pass over the rows
    TRY(normals_run(c));
    TRY(fused_graph_rg_run(c, &p->rg));
    TRY(graph_build(c));
    TRY(seeds_run(c, &p->rg));
  } else {
    // graph first (its lists may be overwritten by the other neighbourhood sizes)
    TRY(graph_build(c));
    TRY(knn_run(c, p->k_normals));
    TRY(normals_run(c));
    TRY(seeds_run(c, &p->rg));
  }
  if (n_seeds) *n_seeds = c->n_seeds;
  const size_t n = c->n;
  const float* d_w = nullptr;
  if (p->weights == 1) {  // morphological gradient of the intensity on the graph (graph.hpp:147-163)
    ENSURE(c, c->out_b, n * 4);
    TRY(morph_run(c, IFTWT_MORPH_GRADIENT, c->out_b.as<float>()));
    d_w = c->out_b.as<float>();
  }
  TRY(ift_run(c, d_w, c->seeds.as<u32>(), c->n_seeds));
  {
    StageTimer st(c, IFTWT_STAGE_DOWNLOAD);
    if (cost) CU_TRY(c, cudaMemcpyAsync(cost, c->out_a.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (label)
      CU_TRY(c, cudaMemcpyAsync(label, c->out_a.as<u32>() + n, n * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  note_fallback(c);
  return IFTWT_OK;
}


int iftwt_cloud_resolution(iftwt_ctx* c, double overlap, double* voxel_size) {
  TRY(check_ctx(c));
  if (!voxel_size) return ctx_fail(c, IFTWT_ERR_INVALID, "null out");
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  return cloud_resolution_run(c, overlap, voxel_size);
}

int iftwt_voxel_graph(iftwt_ctx* c, double voxel_size, size_t* n_vertices) {
  TRY(check_ctx(c));
  TRY(voxel_graph_run(c, voxel_size, 30));
  if (n_vertices) *n_vertices = c->n;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_get_cloud(iftwt_ctx* c, float* xyzi, size_t cap, size_t* n_points) {
  TRY(check_ctx(c));
  if (n_points) *n_points = c->n;
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  if (xyzi) {
    if (cap < c->n) return ctx_fail(c, IFTWT_ERR_CAPACITY, "cloud needs %zu points, caller gave %zu", c->n, cap);
    CU_TRY(c, cudaMemcpyAsync(xyzi, c->pts_in.p, c->n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
  }
  return IFTWT_OK;
}

int iftwt_graph_csr(iftwt_ctx* c, uint32_t* off, uint32_t* adj, size_t cap, size_t* n_arcs) {
  TRY(check_ctx(c));
  if (c->graph_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "graph is not built");
  return graph_export(c, off, adj, cap, n_arcs);
}

int iftwt_morph(iftwt_ctx* c, int op, float* out) {
  TRY(check_ctx(c));
  if (!out) return ctx_fail(c, IFTWT_ERR_INVALID, "null out");
  ENSURE(c, c->out_a, c->n * 4);
  TRY(morph_run(c, op, c->out_a.as<float>()));
  CU_TRY(c, cudaMemcpyAsync(out, c->out_a.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_set_intensity(iftwt_ctx* c, const float* values) {
  TRY(check_ctx(c));
  if (!values || c->n == 0) return ctx_fail(c, IFTWT_ERR_INVALID, "null values or no cloud");
  ENSURE(c, c->out_b, c->n * 4);
  CU_TRY(c, cudaMemcpyAsync(c->out_b.p, values, c->n * 4, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, set_intensity_kernel, div_up(c->n, 256), 256, 0, c->out_b.as<float>(), (u32)c->n, c->pts_in.as<float4>());
  CHECK_LAUNCH(c);
  c->ift_valid = false;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_regmin_seeds(iftwt_ctx* c, int32_t* seeds, size_t cap, size_t* n_seeds) {
  TRY(check_ctx(c));
  TRY(regmin_run(c));
  if (n_seeds) *n_seeds = c->n_seeds;
  if (seeds) {
    if (cap < c->n_seeds)
      return ctx_fail(c, IFTWT_ERR_CAPACITY, "seed buffer needs %zu entries, caller gave %zu", c->n_seeds, cap);
    if (c->n_seeds) CU_TRY(c, cudaMemcpyAsync(seeds, c->seeds.p, c->n_seeds * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
  }
  return IFTWT_OK;
}

int iftwt_region_adjacency(iftwt_ctx* c, uint32_t* a, uint32_t* b, uint32_t* saddle, size_t cap, size_t* n_pairs) {
  TRY(check_ctx(c));
  TRY(region_adjacency_run(c));
  const size_t m = c->n_pairs;
  if (n_pairs) *n_pairs = m;
  if (a || b || saddle) {
    if (cap < m) return ctx_fail(c, IFTWT_ERR_CAPACITY, "pair buffers need %zu entries, caller gave %zu", m, cap);
    if (m) {
      std::vector<u64> keys(m);
      CU_TRY(c, cudaMemcpyAsync(keys.data(), c->h_pairs.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
      if (saddle)
        CU_TRY(c, cudaMemcpyAsync(saddle, (char*)c->h_pairs.p + c->pairs_vals_offset, m * 4, cudaMemcpyDeviceToHost,
                                  c->stream));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
      for (size_t i = 0; i < m; ++i) {
        if (a) a[i] = (uint32_t)(keys[i] >> 32);
        if (b) b[i] = (uint32_t)keys[i];
      }
    }
  }
  return IFTWT_OK;
}

int iftwt_region_info(iftwt_ctx* c, uint32_t* seed, uint32_t* area, float* height, size_t cap, size_t* n_regions) {
  TRY(check_ctx(c));
  TRY(region_info_run(c));
  const size_t nr = c->n_regions;
  if (n_regions) *n_regions = nr;
  if (seed || area || height) {
    if (cap < nr) return ctx_fail(c, IFTWT_ERR_CAPACITY, "region buffers need %zu entries, caller gave %zu", nr, cap);
    if (nr) {
      if (seed)
        CU_TRY(c, cudaMemcpyAsync(seed, c->h_seeds.as<u32>() + (c->n_seeds + 1), nr * 4, cudaMemcpyDeviceToHost,
                                  c->stream));
      if (area) CU_TRY(c, cudaMemcpyAsync(area, c->h_info.p, nr * 4, cudaMemcpyDeviceToHost, c->stream));
      if (height)
        CU_TRY(c, cudaMemcpyAsync(height, c->h_info.as<u32>() + (nr + 1), nr * 4, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_cluster(iftwt_ctx* c, int num_regions, uint32_t* cluster_of, size_t* n_clusters) {
  TRY(check_ctx(c));
  if (!c->ift_valid) return ctx_fail(c, IFTWT_ERR_STATE, "no IFT result on the device");
  ENSURE(c, c->out_b, c->n * 4);
  TRY(cluster_run(c, num_regions, c->out_b.as<u32>(), n_clusters));
  if (cluster_of) CU_TRY(c, cudaMemcpyAsync(cluster_of, c->out_b.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_knn_device(iftwt_ctx* c, int k, int32_t* d_idx, float* d_d2) {
  TRY(check_ctx(c));
  if (!d_idx) return ctx_fail(c, IFTWT_ERR_INVALID, "null device output");
  TRY(knn_run(c, k));
  const size_t ne = c->n * (size_t)k;
  LAUNCH(c, export_knn_kernel, div_up(ne, 256), 256, 0, c->nbr.as<u32>(), c->nd2.as<float>(), c->inv.as<u32>(),
         c->perm.as<u32>(), ne, k, (int*)d_idx, d_d2);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  note_fallback(c);
  return IFTWT_OK;
}

int iftwt_normals_device(iftwt_ctx* c, int k, float* d_out4) {
  TRY(check_ctx(c));
  if (!d_out4) return ctx_fail(c, IFTWT_ERR_INVALID, "null device output");
  TRY(knn_run(c, k));
  TRY(normals_run(c));
  LAUNCH(c, export_normals_kernel, div_up(c->n, 256), 256, 0, c->nrm.as<float4>(), c->inv.as<u32>(), (u32)c->n,
         (float4*)d_out4);
  CHECK_LAUNCH(c);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_set_knn_rows_device(iftwt_ctx* c, int k, const int32_t* d_idx, const float* d_d2) {
  TRY(check_ctx(c));
  if (c->n == 0) return ctx_fail(c, IFTWT_ERR_STATE, "no cloud set");
  if (!d_idx || !d_d2 || k < 1 || k > 64) return ctx_fail(c, IFTWT_ERR_INVALID, "bad rows");
  TRY(grid_build(c, k));
  TRY(graph_detach_rows(c));
  const size_t n = c->n;
  ENSURE(c, c->nbr, n * k * 4);
  ENSURE(c, c->nd2, n * k * 4);
  ENSURE(c, c->kth, n * 8);
  ENSURE(c, c->fb_list, n * 4);
  LAUNCH(c, install_rows_kernel, div_up(n * 32, 256), 256, 0, (const int*)d_idx, d_d2, c->perm.as<u32>(),
         c->inv.as<u32>(), (u32)n, k, c->nbr.as<u32>(), c->nd2.as<float>(), c->kth.as<u64>());
  CHECK_LAUNCH(c);
  c->knn_k = k;
  c->mask_k = 0;
  if (c->graph_kind == 1) {
    c->graph_k = 0;
    c->graph_kind = 0;
  }
  c->nrm_k = 0;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_morph_rgb(iftwt_ctx* c, int op, uint32_t* rgb_out) {
  TRY(check_ctx(c));
  if (!rgb_out) return ctx_fail(c, IFTWT_ERR_INVALID, "null out");
  ENSURE(c, c->out_a, c->n * 4);
  TRY(morph_rgb_run(c, op, c->out_a.as<u32>()));
  CU_TRY(c, cudaMemcpyAsync(rgb_out, c->out_a.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_rgb_to_gray(iftwt_ctx* c, int otsu, double* threshold) {
  TRY(check_ctx(c));
  TRY(rgb_to_gray_run(c, otsu, threshold));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

int iftwt_stage_ms(iftwt_ctx* c, int stage, float* ms) {
  TRY(check_ctx(c));
  if (stage < 0 || stage >= IFTWT_STAGE_COUNT || !ms) return ctx_fail(c, IFTWT_ERR_INVALID, "bad stage");
  if (!c->ev_set[stage]) {
    *ms = 0.f;
    return IFTWT_OK;
  }
  CU_TRY(c, cudaEventSynchronize(c->ev1[stage]));
  CU_TRY(c, cudaEventElapsedTime(ms, c->ev0[stage], c->ev1[stage]));
  return IFTWT_OK;
}

int iftwt_get_stats(iftwt_ctx* c, iftwt_stats* out) {
  TRY(check_ctx(c));
  if (!out) return ctx_fail(c, IFTWT_ERR_INVALID, "null out");
  *out = c->stats;
  out->density_spread = c->stats_density_spread;
  return IFTWT_OK;
}

int iftwt_reset_counters(iftwt_ctx* c) {
  TRY(check_ctx(c));
  c->stats.kernel_launches = 0;
  for (int s = 0; s < IFTWT_STAGE_COUNT; ++s) c->ev_set[s] = false;
  return IFTWT_OK;
}

int iftwt_synchronize(iftwt_ctx* c) {
  TRY(check_ctx(c));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return IFTWT_OK;
}

void* iftwt_stream(iftwt_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"

// ---- batch of clouds on one GPU (BASELINE config 5) ------------------------------------------------
// A 2 M-point cloud does not fill a B200 for long: 162 of its IFT level launches, the label-min sweeps of
// the seed stage and every host round trip leave most SMs idle.  A batch is therefore run on several
// LANES — one ctx, one stream and one host thread each, independent by the ABI's contract — that pull
// clouds from one queue: the latency-bound stretches of one cloud are filled by the kernels of the
// others.  Results are those of iftwt_segment on each cloud by construction.
// Host clouds: every lane owns an upload and a download stream next to its ctx's compute stream.  The
// lane's NEXT cloud is copied into the second of two staging buffers while the current one is being
// segmented, and the results leave through a result buffer of the lane's own while the next cloud is
// already running — the copies never sit in front of a kernel in the compute stream, so two lanes keep
// the SMs as busy as with resident clouds (three ctxs that copy in their own streams, as the bench
// drove them before, pay for the third cloud's share of the caches: 16.4 ms per 10 M-point cloud
// against 15.4 with two).
struct LaneIO {
  cudaStream_t up = nullptr, down = nullptr;
  void* in[2] = {nullptr, nullptr};
  size_t in_cap[2] = {0, 0};
  void* res = nullptr;
  size_t res_cap = 0;
  cudaEvent_t up_done[2] = {nullptr, nullptr}, res_ready = nullptr, down_done = nullptr;
  bool ready = false;
};
struct iftwt_batch {
  int device = 0;
  std::vector<iftwt_ctx*> lanes;
  std::vector<LaneIO> io;
  std::string err;
};

static cudaError_t lane_io_setup(LaneIO& io) {
  if (io.ready) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&io.up, cudaStreamNonBlocking)) != cudaSuccess) return e;
  if ((e = cudaStreamCreateWithFlags(&io.down, cudaStreamNonBlocking)) != cudaSuccess) return e;
  cudaEvent_t* evs[] = {&io.up_done[0], &io.up_done[1], &io.res_ready, &io.down_done};
  for (cudaEvent_t* ev : evs)
    if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
  io.ready = true;
  return cudaSuccess;
}
static void lane_io_free(LaneIO& io) {
  if (io.up) cudaStreamSynchronize(io.up);
  if (io.down) cudaStreamSynchronize(io.down);
  for (int s = 0; s < 2; ++s)
    if (io.in[s]) cudaFree(io.in[s]);
  if (io.res) cudaFree(io.res);
  cudaEvent_t evs[] = {io.up_done[0], io.up_done[1], io.res_ready, io.down_done};
  for (cudaEvent_t ev : evs)
    if (ev) cudaEventDestroy(ev);
  if (io.up) cudaStreamDestroy(io.up);
  if (io.down) cudaStreamDestroy(io.down);
  io = LaneIO();
}
static cudaError_t grow(void*& p, size_t& cap, size_t bytes) {
  if (cap >= bytes) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  cudaError_t e = cudaMalloc(&p, bytes + bytes / 8);
  if (e == cudaSuccess) cap = bytes + bytes / 8;
  return e;
}

extern "C" {

int iftwt_batch_create(iftwt_batch** out, int device, int lanes) {
  if (!out) {
    g_err = "null out pointer";
    return IFTWT_ERR_INVALID;
  }
  *out = nullptr;
  if (lanes < 1 || lanes > 16) {
    g_err = "lanes must be in [1, 16]";
    return IFTWT_ERR_INVALID;
  }
  iftwt_batch* b = new (std::nothrow) iftwt_batch();
  if (!b) {
    g_err = "out of host memory";
    return IFTWT_ERR_CUDA;
  }
  b->device = device;
  for (int l = 0; l < lanes; ++l) {
    iftwt_ctx* c = nullptr;
    int rc = iftwt_create(&c, device);
    if (rc != IFTWT_OK) {
      iftwt_batch_destroy(b);
      return rc;
    }
    b->lanes.push_back(c);
  }
  b->io.resize(b->lanes.size());
  *out = b;
  return IFTWT_OK;
}

void iftwt_batch_destroy(iftwt_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (LaneIO& io : b->io) lane_io_free(io);
  for (iftwt_ctx* c : b->lanes) iftwt_destroy(c);
  delete b;
}

const char* iftwt_batch_last_error(const iftwt_batch* b) { return b ? b->err.c_str() : g_err.c_str(); }

iftwt_ctx* iftwt_batch_ctx(iftwt_batch* b, int lane) {
  if (!b || lane < 0 || (size_t)lane >= b->lanes.size()) return nullptr;
  return b->lanes[lane];
}

int iftwt_batch_segment(iftwt_batch* b, const void* const* clouds, const size_t* n, size_t B, int clouds_on_device,
                        const iftwt_segment_params* params, uint32_t* const* cost, uint32_t* const* label,
                        size_t* n_seeds) {
  if (!b) {
    g_err = "null batch";
    return IFTWT_ERR_INVALID;
  }
  if (!clouds || !n || !params) {
    b->err = "null argument";
    return IFTWT_ERR_INVALID;
  }
  std::atomic<size_t> next(0);
  std::atomic<int> first_rc(IFTWT_OK);
  std::vector<std::string> errs(b->lanes.size());
  auto fail = [&](size_t lane, size_t i, int rc, const std::string& what) {
    errs[lane] = "cloud " + std::to_string(i) + ": " + what;
    int expected = IFTWT_OK;
    first_rc.compare_exchange_strong(expected, rc);
  };
  auto work_device = [&](size_t lane) {
    iftwt_ctx* c = b->lanes[lane];
    while (first_rc.load() == IFTWT_OK) {
      const size_t i = next.fetch_add(1);
      if (i >= B) break;
      int rc = iftwt_set_cloud_device(c, clouds[i], n[i]);
      size_t ns = 0;
      if (rc == IFTWT_OK)
        rc = iftwt_segment(c, params, cost ? cost[i] : nullptr, label ? label[i] : nullptr, &ns);
      if (rc != IFTWT_OK) {
        fail(lane, i, rc, c->err);
        break;
      }
      if (n_seeds) n_seeds[i] = ns;
    }
  };
  // host clouds: uploads and downloads in the lane's own copy streams (LaneIO above)
  auto work_host = [&](size_t lane) {
    iftwt_ctx* c = b->lanes[lane];
    LaneIO& io = b->io[lane];
    cudaSetDevice(b->device);
#define LANE_CU(i_, expr)                                                            \
  {                                                                                  \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      fail(lane, (i_), IFTWT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
      break;                                                                         \
    }                                                                                \
  }
    bool down_pending = false;
    do {
      LANE_CU(0, lane_io_setup(io));
      size_t i = next.fetch_add(1);
      if (i >= B) break;
      if (!clouds[i] || n[i] == 0) {
        fail(lane, i, IFTWT_ERR_INVALID, "bad cloud");
        break;
      }
      LANE_CU(i, grow(io.in[0], io.in_cap[0], n[i] * 16));
      LANE_CU(i, cudaMemcpyAsync(io.in[0], clouds[i], n[i] * 16, cudaMemcpyHostToDevice, io.up));
      LANE_CU(i, cudaEventRecord(io.up_done[0], io.up));
      int cur = 0;
      while (true) {
        // the lane's next cloud starts moving now (the staging buffer it goes to was last read by a copy of
        // the compute stream that iftwt_segment has long synchronised on)
        size_t inext = B;
        if (first_rc.load() == IFTWT_OK) inext = next.fetch_add(1);
        if (inext < B) {
          if (!clouds[inext] || n[inext] == 0) {
            fail(lane, inext, IFTWT_ERR_INVALID, "bad cloud");
            break;
          }
          LANE_CU(inext, grow(io.in[cur ^ 1], io.in_cap[cur ^ 1], n[inext] * 16));
          LANE_CU(inext, cudaMemcpyAsync(io.in[cur ^ 1], clouds[inext], n[inext] * 16, cudaMemcpyHostToDevice, io.up));
          LANE_CU(inext, cudaEventRecord(io.up_done[cur ^ 1], io.up));
        }
        LANE_CU(i, cudaStreamWaitEvent(c->stream, io.up_done[cur], 0));
        int rc = iftwt_set_cloud_device(c, io.in[cur], n[i]);
        size_t ns = 0;
        if (rc == IFTWT_OK) rc = iftwt_segment(c, params, nullptr, nullptr, &ns);
        if (rc != IFTWT_OK) {
          fail(lane, i, rc, c->err);
          break;
        }
        if (n_seeds) n_seeds[i] = ns;
        uint32_t* co = cost ? cost[i] : nullptr;
        uint32_t* la = label ? label[i] : nullptr;
        if (co || la) {
          if (io.res_cap < n[i] * 8) {
            if (down_pending) LANE_CU(i, cudaStreamSynchronize(io.down));
            LANE_CU(i, grow(io.res, io.res_cap, n[i] * 8));
          }
          // cost | label leave the ctx (its buffers belong to the next cloud from here on) ...
          if (down_pending) LANE_CU(i, cudaStreamWaitEvent(c->stream, io.down_done, 0));
          LANE_CU(i, cudaMemcpyAsync(io.res, c->out_a.p, n[i] * 8, cudaMemcpyDeviceToDevice, c->stream));
          LANE_CU(i, cudaEventRecord(io.res_ready, c->stream));
          // ... and go home in the download stream while the next cloud runs
          LANE_CU(i, cudaStreamWaitEvent(io.down, io.res_ready, 0));
          if (co) LANE_CU(i, cudaMemcpyAsync(co, io.res, n[i] * 4, cudaMemcpyDeviceToHost, io.down));
          if (la)
            LANE_CU(i, cudaMemcpyAsync(la, (const uint32_t*)io.res + n[i], n[i] * 4, cudaMemcpyDeviceToHost, io.down));
          LANE_CU(i, cudaEventRecord(io.down_done, io.down));
          down_pending = true;
        }
        if (inext >= B) break;
        i = inext;
        cur ^= 1;
      }
    } while (false);
#undef LANE_CU
    // the caller's buffers are his again when this returns, whatever happened
    if (io.up) cudaStreamSynchronize(io.up);
    if (io.down) cudaStreamSynchronize(io.down);
  };
  auto work = [&](size_t lane) {
    if (clouds_on_device) work_device(lane);
    else work_host(lane);
  };
  std::vector<std::thread> th;
  for (size_t l = 1; l < b->lanes.size() && l < B; ++l) th.emplace_back(work, l);
  work(0);
  for (std::thread& t : th) t.join();
  if (first_rc.load() != IFTWT_OK) {
    for (const std::string& e : errs)
      if (!e.empty()) {
        b->err = e;
        break;
      }
    return first_rc.load();
  }
  return IFTWT_OK;
}

}  // extern "C"
