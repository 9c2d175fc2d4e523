/* iftwt.h — C ABI of the B200-native IFT-WT segmentation hot path.
 *
 * Drop-in boundary for enemy537/IFTWT_RG.  The reference has no plugin / FFI
 * interface of its own: its boundary is the set of C++ constructors and getters
 * that the pipeline block main.cpp:91-129 calls.  Each entry point below names
 * the reference interface it replaces (file:line in /root/reference); the
 * header-only adapters in include/iftwt_adapters.hpp keep the reference's class
 * and method names on top of this ABI.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, returns int (0 = IFTWT_OK, < 0 =
 *     error), never throws.  iftwt_last_error(ctx) gives the message.
 *   - every index in and out of this ABI is an ORIGINAL point index (the order
 *     of the cloud passed to iftwt_set_cloud); the Morton order used on the
 *     device never leaks.
 *   - the caller owns every host buffer.  An output pointer may be NULL: the
 *     stage still runs and its result stays device resident for the next stage.
 *   - a ctx owns one CUDA device, one stream and all device memory; it is not
 *     re-entrant.  Different ctxs are independent (batch data-parallelism uses
 *     one ctx per GPU / per process).
 *   - there is no CPU fallback: every entry point fails with IFTWT_ERR_CUDA if
 *     no sm_100 device is usable.
 */
#ifndef IFTWT_H
#define IFTWT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFTWT_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define IFTWT_API __attribute__((visibility("default")))
#else
#define IFTWT_API
#endif

enum {
  IFTWT_OK = 0,
  IFTWT_ERR_INVALID = -1,     /* bad argument */
  IFTWT_ERR_CUDA = -2,        /* CUDA runtime failure / no device */
  IFTWT_ERR_STATE = -3,       /* stage called before its inputs exist */
  IFTWT_ERR_NONFINITE = -4,   /* cloud holds NaN / Inf coordinates */
  IFTWT_ERR_UNSUPPORTED = -5, /* parameter combination the GPU path cannot restate exactly */
  IFTWT_ERR_CAPACITY = -6,    /* caller buffer too small (needed size is reported) */
  IFTWT_ERR_WEIGHTS = -7      /* IFT weights are negative or not integer valued (App. B #6) */
};

/* ift.hpp:1-2 */
#define IFTWT_MAX_COST 500u
#define IFTWT_MIN_COST 0u

typedef struct iftwt_ctx iftwt_ctx;

/* pcl::RegionGrowing setters used at main.cpp:111-119. */
typedef struct iftwt_rg_params {
  int min_cluster_size;       /* setMinClusterSize       (main.cpp:112, 50)      */
  int max_cluster_size;       /* setMaxClusterSize       (main.cpp:113, 10000)   */
  int number_of_neighbours;   /* setNumberOfNeighbours   (main.cpp:115, 50)      */
  float smoothness_threshold; /* setSmoothnessThreshold  (main.cpp:118, 3 deg in rad) */
  float curvature_threshold;  /* setCurvatureThreshold   (main.cpp:119, 1.0)     */
} iftwt_rg_params;

/* Whole pipeline of main.cpp:97-123 in one call (k-NN graph builder (B)). */
typedef struct iftwt_segment_params {
  int k_graph;   /* neighbours of the symmetric k-NN graph the IFT runs on            */
  int k_normals; /* setKSearch of the normal estimation (main.cpp:108, 30)            */
  iftwt_rg_params rg;
  int weights;   /* 0: vertex intensity (what ift.hpp:198 reads), 1: morphological
                    gradient dilate-erode of the intensity on the graph (graph.hpp:147-163) */
} iftwt_segment_params;

/* Stage ids for iftwt_stage_ms(). */
enum {
  IFTWT_STAGE_UPLOAD = 0,  /* host -> device copy of the cloud                */
  IFTWT_STAGE_GRID = 1,    /* bbox, cell-size estimate, Morton sort, cell hash */
  IFTWT_STAGE_KNN = 2,     /* k-NN kernels                                     */
  IFTWT_STAGE_GRAPH = 3,   /* symmetric CSR build                              */
  IFTWT_STAGE_NORMALS = 4, /* covariance + eigen kernel                        */
  IFTWT_STAGE_SEEDS = 5,   /* region growing, centroids, snap                  */
  IFTWT_STAGE_IFT = 6,     /* watershed                                        */
  IFTWT_STAGE_DOWNLOAD = 7,
  IFTWT_STAGE_KNN_CELL = 8, /* the k-NN fast-path kernel alone (the roofline kernel)     */
  IFTWT_STAGE_GRAPH_RG = 9, /* iftwt_segment's fused graph-mutuality + region-growing arcs pass */
  IFTWT_STAGE_COUNT = 10
};

typedef struct iftwt_stats {
  uint64_t n_points;
  uint64_t n_cells;          /* occupied grid cells                                   */
  double cell_size;
  uint64_t knn_fallback;     /* queries that left the one-ring fast path              */
  uint64_t n_arcs;           /* directed arcs of the last graph                       */
  uint64_t rg_segments;      /* region-growing segments before the size filter        */
  uint64_t rg_sweeps;        /* label-min sweeps over one-way arcs                    */
  uint64_t n_seeds;          /* distinct IFT seeds                                    */
  uint64_t ift_levels;       /* non-empty cost levels processed                       */
  uint64_t kernel_launches;  /* kernels launched by this ctx since iftwt_reset_counters */
  uint64_t device_bytes;     /* device memory currently held                          */
  double density_spread;     /* mean / 25th-percentile cell occupancy one level above the cell scale
                                (1.2 on a uniformly sampled surface; >= 1.35 enlarges the cell)  */
} iftwt_stats;

IFTWT_API int iftwt_version(void);

/* Lifetime.  `device` is a CUDA ordinal.  Replaces nothing in the reference
 * (it has no context object; everything is process-global). */
IFTWT_API int iftwt_create(iftwt_ctx** out, int device);
IFTWT_API void iftwt_destroy(iftwt_ctx* ctx);
IFTWT_API const char* iftwt_last_error(const iftwt_ctx* ctx); /* valid until the next call on ctx; ctx may be NULL */

/* Input cloud.  Replaces Graph::Graph(CloudI::Ptr) / setInputCloud
 * (graph.hpp:34-57) and NormalEstimationOMP::setInputCloud (main.cpp:107).
 * `points` is an array of `n` records `stride_bytes` apart; xyz are three
 * consecutive floats at byte offset `xyz_offset`, intensity one float at
 * `intensity_offset` (or -1: intensity 0).  This accepts pcl::PointXYZI (stride
 * 32, offsets 0 / 16), pcl::PointXYZ (16, 0, -1) and packed float4 (16, 0, 12)
 * without repacking on the host. */
IFTWT_API int iftwt_set_cloud(iftwt_ctx* ctx, const void* points, size_t stride_bytes, size_t n,
                    int xyz_offset, int intensity_offset);
/* Same, from a device-resident packed float4 {x,y,z,intensity} array on ctx's device. */
IFTWT_API int iftwt_set_cloud_device(iftwt_ctx* ctx, const void* device_float4, size_t n);

/* Exact k-NN of every cloud point (self included, first).  Replaces
 * pcl::search::KdTree::nearestKSearch(index, k, ...) as called at graph.hpp:245
 * and inside PCL at main.cpp:108 / 115.  Rows are in canonical (d2, index)
 * order; missing neighbours (n < k) are -1 / +inf.  idx: n*k, d2: n*k or NULL. */
IFTWT_API int iftwt_knn(iftwt_ctx* ctx, int k, int32_t* idx, float* d2);

/* Exact k-NN of arbitrary query points against the cloud.  Replaces
 * Graph::find_p (graph.hpp:260-266), flat_point.hpp:74 and ift.hpp:135.
 * queries: m records of 3 floats, `stride_bytes` apart. */
IFTWT_API int iftwt_knn_query(iftwt_ctx* ctx, const void* queries, size_t stride_bytes, size_t m, int k,
                    int32_t* idx, float* d2);

/* Symmetric k-NN graph as CSR (rows ascending, no self loops, no duplicates):
 * graph builder (B).  Replaces Graph::getAdjacencyList (graph.hpp:63-66) for the
 * BASELINE k-NN configurations.  off: n+1 entries.  adj: `adj_capacity` entries
 * or NULL.  *n_arcs receives the arc count (also on IFTWT_ERR_CAPACITY). */
IFTWT_API int iftwt_knn_graph(iftwt_ctx* ctx, int k, uint32_t* off, uint32_t* adj, size_t adj_capacity,
                    size_t* n_arcs);

/* Normals + curvature.  Replaces NormalEstimationOMP::{setKSearch, compute}
 * (main.cpp:104-109).  Writes, per point, normal xyz at byte offset
 * `normal_offset` and curvature at `curvature_offset` of records `stride_bytes`
 * apart (pcl::Normal: 32, 0, 16; packed float4: 16, 0, 12).  out may be NULL. */
IFTWT_API int iftwt_normals(iftwt_ctx* ctx, int k, void* out, size_t stride_bytes, int normal_offset,
                  int curvature_offset);

/* Region-growing seeds.  Replaces FlatPoint::getCentroidsCloud
 * (flat_point.hpp:50-56 -> 65-91 -> 23-42 -> pcl::RegionGrowing::extract).
 * Needs iftwt_normals first.  point_label[n] (or NULL) = index of the kept
 * cluster, PCL emission order, or -1.  seed_index[seed_capacity] (or NULL) = the
 * input point nearest to each kept cluster's centroid. *n_seeds = kept clusters. */
IFTWT_API int iftwt_region_seeds(iftwt_ctx* ctx, const iftwt_rg_params* params, int32_t* point_label,
                       int32_t* seed_index, size_t seed_capacity, size_t* n_seeds);

/* IFT watershed with the f_max path cost on the current graph.  Replaces
 * IFT_PCD::IFT_PCD(Graph&, Cloud::Ptr) (ift.hpp:31-37: copy_graph, add_seeds,
 * compute_watershed) and getRoots (ift.hpp:48-50).
 *   weights: n integer-valued floats in original order, or NULL = intensities.
 *   seeds:   n_seeds vertex indices, or NULL = the seeds left on the device by
 *            iftwt_region_seeds.
 *   cost[n]  (or NULL): final path cost, IFTWT_MAX_COST where unreachable.
 *   label[n] (or NULL): seed vertex that conquered the vertex (root_m,
 *            ift.hpp:204); unreachable vertices keep label = self.
 * Ties are broken by the "min seed" policy (policy C). */
IFTWT_API int iftwt_ift(iftwt_ctx* ctx, const float* weights, const int32_t* seeds, size_t n_seeds,
              uint32_t* cost, uint32_t* label);

/* main.cpp:97-123 in one call: k-NN graph, normals, region-growing seeds, IFT.
 * Everything stays on the device between stages. */
IFTWT_API int iftwt_segment(iftwt_ctx* ctx, const iftwt_segment_params* params, uint32_t* cost,
                  uint32_t* label, size_t* n_seeds);


/* ---- graph builder (A): the reference's own voxel-26 graph ------------------------- */

/* Graph::computeCloudResolution (graph.hpp:228-259): mean distance to the 2nd nearest
 * neighbour times `overlap` (1.01 at graph.hpp:34, 5.0 at main.cpp:87).  The mean is the
 * exact sum of the float terms, rounded once (order independent). */
IFTWT_API int iftwt_cloud_resolution(iftwt_ctx* ctx, double overlap, double* voxel_size);

/* Graph::initialize (graph.hpp:293-308): occupied voxels of side `voxel_size` (PCL
 * OctreePointCloudAdjacency), vertices = voxel centres carrying the intensity of their
 * nearest input point (optimizeGraph, graph.hpp:269-284), edges = occupied 26-neighbours.
 * As in the reference (graph.hpp:307) the vertices REPLACE the ctx's cloud: every later
 * call (iftwt_get_cloud, iftwt_normals, iftwt_region_seeds, iftwt_ift ...) works on the
 * graph's vertices, numbered in ascending Morton order of the octree key. */
IFTWT_API int iftwt_voxel_graph(iftwt_ctx* ctx, double voxel_size, size_t* n_vertices);

/* Graph::getCloud (graph.hpp:177-180): the current cloud as packed float4 {x,y,z,intensity}. */
IFTWT_API int iftwt_get_cloud(iftwt_ctx* ctx, float* xyzi, size_t capacity_points, size_t* n_points);

/* Graph::getAdjacencyList (graph.hpp:63-66) for whichever graph is current, as CSR in
 * original vertex ids, rows ascending, no self loops. */
IFTWT_API int iftwt_graph_csr(iftwt_ctx* ctx, uint32_t* off, uint32_t* adj, size_t adj_capacity, size_t* n_arcs);

/* Graph::morph_* (graph.hpp:131-163, 323-371) over the closed neighbourhood.
 * op: 1 bin_erode, 2 bin_dilate, 3 dilate, 4 erode (enum MORPH, graph.hpp:13), 5 gradient. */
enum { IFTWT_MORPH_BIN_ERODE = 1, IFTWT_MORPH_BIN_DILATE = 2, IFTWT_MORPH_DILATE = 3, IFTWT_MORPH_ERODE = 4,
       IFTWT_MORPH_GRADIENT = 5 };
IFTWT_API int iftwt_morph(iftwt_ctx* ctx, int op, float* out);

/* Graph::pc_to_g / morph_erode_gradient write-back (graph.hpp:164-176, 203-208): replace the
 * vertex intensities (n floats, original order). */
IFTWT_API int iftwt_set_intensity(iftwt_ctx* ctx, const float* values);

/* Graph::getRegmin (graph.hpp:209-226), the seeds of IFT_PCD(Graph&) (ift.hpp:38-44):
 * vertices whose value is 0 or the minimum of their closed neighbourhood.  Leaves them on
 * the device for iftwt_ift(seeds = NULL). */
IFTWT_API int iftwt_regmin_seeds(iftwt_ctx* ctx, int32_t* seeds, size_t capacity, size_t* n_seeds);

/* ---- region hierarchy (ift.hpp:5-23, 207-243; cluster.hpp:23-229) ---------------------
 * NOT a reproduction of the reference's getMST() / r_info, and not claimed as one.  The reference
 * updates r_info inside the relaxation branch of its heap loop (ift.hpp:207-209: once per cost
 * improvement, not once per vertex), records an edge at a region's FIRST contact with another
 * (guarded by `status`, ift.hpp:220-222) with the volumes the two regions happen to have at that
 * moment, and merges their bookkeeping on the spot (ift.hpp:226-240).  Every one of those values is a
 * function of the sequential relaxation order — which itself depends on container address order
 * (globals.hpp:37-38) — so no order-free computation reproduces them; Cluster (cluster.hpp) consumes
 * that map and has undefined behaviour of its own (cluster.hpp:164, 187).  The literal procedure is
 * restated in the oracle and checked bit for bit against the reference's own code
 * (tests/test_ref_ift.py); it is test infrastructure.  What the product provides instead
 *  is the classical, order-free watershed hierarchy on the same IFT result: the region
 * adjacency graph with saddle heights, per-region area and height, and a single-linkage merge along
 * its minimum spanning forest down to num_regions.  All three calls need a finished iftwt_ift /
 * iftwt_segment. */

/* Region adjacency: pairs a < b of seed ids whose regions touch, saddle = min over the
 * touching arcs of max(cost[u], cost[v]).  Sorted by (a, b).  The product's stand-in for the edge
 * list getMST returns (ift.hpp:45-47) — a different, order-free object (see above). */
IFTWT_API int iftwt_region_adjacency(iftwt_ctx* ctx, uint32_t* a, uint32_t* b, uint32_t* saddle, size_t capacity,
                           size_t* n_pairs);
/* Region attributes, sorted by seed id: area = conquered vertices, height = max vertex
 * value (the order-free part of r_info, ift.hpp:5-23; `volume` is order dependent and not provided). */
IFTWT_API int iftwt_region_info(iftwt_ctx* ctx, uint32_t* seed, uint32_t* area, float* height, size_t capacity,
                      size_t* n_regions);
/* Cluster(mst, root_m, g_v, num_regions) + getLabelCloud (cluster.hpp:25-81): merge the
 * regions down to num_regions; cluster_of[n] = smallest seed id of the vertex's cluster,
 * 0xFFFFFFFF for vertices the IFT never conquered. */
IFTWT_API int iftwt_cluster(iftwt_ctx* ctx, int num_regions, uint32_t* cluster_of, size_t* n_clusters);

/* ---- device-pointer variants (multi-GPU plumbing: the Morton-range sharded k-NN of
 * iftwt_rg_b200/sharded.py keeps everything on the device between NCCL exchanges) -------- */

/* iftwt_knn / iftwt_normals with DEVICE output buffers on ctx's device (rows in original
 * index order; d_d2 may be NULL; d_out4 is packed float4 {nx,ny,nz,curvature}). */
IFTWT_API int iftwt_knn_device(iftwt_ctx* ctx, int k, int32_t* d_idx, float* d_d2);
IFTWT_API int iftwt_normals_device(iftwt_ctx* ctx, int k, float* d_out4);
/* Install externally computed k-NN rows (DEVICE buffers, original indices, canonical order,
 * -1 / +inf padded) as the ctx's current lists: the bridge from a sharded k-NN to the
 * single-GPU graph / seeds / IFT stages.  The cloud must be set. */
IFTWT_API int iftwt_set_knn_rows_device(iftwt_ctx* ctx, int k, const int32_t* d_idx, const float* d_d2);

/* ---- a batch of clouds on one GPU (BASELINE config 5) ----------------------------------------
 * The reference segments one cloud per process run (main.cpp:78-148).  A batch object owns `lanes`
 * ctxs on one device — one stream and one host thread each — that pull clouds from one queue, so the
 * latency-bound stretches of one cloud (small IFT levels, seed sweeps, host round trips) are filled by
 * the kernels of the others.  Results are exactly those of iftwt_segment on each cloud. */
typedef struct iftwt_batch iftwt_batch;
IFTWT_API int iftwt_batch_create(iftwt_batch** out, int device, int lanes);
IFTWT_API void iftwt_batch_destroy(iftwt_batch* batch);
IFTWT_API const char* iftwt_batch_last_error(const iftwt_batch* batch); /* batch may be NULL */
IFTWT_API iftwt_ctx* iftwt_batch_ctx(iftwt_batch* batch, int lane);     /* e.g. for iftwt_get_stats */
/* clouds[i]: n[i] packed float4 {x,y,z,intensity} records — DEVICE pointers when clouds_on_device != 0,
 * host pointers otherwise (pinned memory lets the copies run beside the kernels: every lane uploads its next
 * cloud and downloads its previous results in copy streams of its own while the current cloud is segmented;
 * pageable memory works, the copies then block the lane's host thread).  cost / label: B host pointers (each n[i] entries) or NULL; entries may be NULL.
 * n_seeds: B entries or NULL.  Returns the first error of any cloud. */
IFTWT_API int iftwt_batch_segment(iftwt_batch* batch, const void* const* clouds, const size_t* n, size_t B,
                                  int clouds_on_device, const iftwt_segment_params* params, uint32_t* const* cost,
                                  uint32_t* const* label, size_t* n_seeds);

/* ---- Morton-range sharded k-NN / normals across the GPUs of one node (BASELINE config 4) ------
 * The reference has no multi-device path; what is sharded are its per-point loops — nearestKSearch
 * (graph.hpp:245) and NormalEstimationOMP::compute (main.cpp:102-109).  Points are partitioned by
 * contiguous ranges of a coarse Morton key, one range per rank; every rank also receives the halo
 * (the points of the 26 coarse cells around each of its own) from its peers over NVLink (grouped
 * ncclSend / ncclRecv), runs the single-GPU kernels, and keeps the rows of the points it owns,
 * translated to GLOBAL point ids.  The rows are bit-identical to what one GPU computes on the
 * whole cloud (ties break by global id).  From a comm's second call on the exchange is written by
 * the partition kernel straight into the peers' receive buffers (cudaIpc / peer access over NVLink)
 * and the rows of iftwt_sharded_segment straight into the root's buffers; NCCL then only carries
 * the small all-gathers / all-reduces that double as barriers.  The IFT does not shard: iftwt_sharded_segment gathers
 * the rows to one rank and runs graph / seeds / IFT there.  (csrc/shard.cu) */
typedef struct iftwt_comm iftwt_comm;

#define IFTWT_COMM_ID_BYTES 128
#define IFTWT_COMM_MAX_RANKS 8

/* One process per GPU: rank 0 calls iftwt_comm_unique_id (ncclGetUniqueId) and hands the 128 bytes to
 * every rank over any side channel; every rank then calls iftwt_comm_create (ncclCommInitRank). */
IFTWT_API int iftwt_comm_unique_id(void* id128);
IFTWT_API int iftwt_comm_create(iftwt_comm** out, const void* id128, int nranks, int rank, int device);
/* One process driving ndev GPUs (ncclCommInitAll); shard i is rank i on devices[i].  Naming the
 * same device ndev times gives ndev shards that time-share that GPU, with device-to-device copies
 * as transport (NCCL refuses duplicate devices): the whole path on a single GPU, for testing. */
IFTWT_API int iftwt_comm_create_all(iftwt_comm** out, int ndev, const int* devices);
IFTWT_API void iftwt_comm_destroy(iftwt_comm* comm);
IFTWT_API const char* iftwt_comm_last_error(const iftwt_comm* comm); /* comm may be NULL */
IFTWT_API int iftwt_comm_info(const iftwt_comm* comm, int* nranks, int* n_local, int* first_local_rank);
/* The ctx of local shard `local_index` (owned by the comm).  After iftwt_sharded_segment the root's
 * ctx holds the whole cloud, its graph and the IFT result: iftwt_region_adjacency, iftwt_cluster,
 * iftwt_graph_csr ... apply to it. */
IFTWT_API iftwt_ctx* iftwt_comm_ctx(iftwt_comm* comm, int local_index);

/* A rank's slice of the cloud: DEVICE packed float4 {x,y,z,intensity} on the shard's device.  The
 * slices of ranks 0, 1, ... are consecutive ranges of global point ids (gid0 of rank r+1 =
 * gid0 + n of rank r, rank 0 starts at 0); at most 2^31 - 1 points in all. */
typedef struct iftwt_shard_in {
  const void* d_points;
  size_t n;
  uint64_t gid0;
} iftwt_shard_in;

/* For hosts without CUDA of their own (the reference's main.cpp holds its cloud in a std::vector): copies
 * `n` packed float4 {x,y,z,intensity} records from HOST memory into a device buffer the comm owns for local
 * shard `local_index`, and fills *in for the calls below.  The buffer lives until the next upload to that shard. */
IFTWT_API int iftwt_comm_upload(iftwt_comm* comm, int local_index, const void* host_points, size_t n, uint64_t gid0,
                                iftwt_shard_in* in);

/* Result of one shard.  Device pointers are owned by the comm and valid until its next call. */
typedef struct iftwt_shard_out {
  size_t n_owned;            /* points of the Morton range this rank owns                      */
  size_t n_halo;             /* halo points received on top of them                            */
  size_t n_total;            /* points of the whole cloud                                      */
  int k;
  int attempts;              /* exchanges run (the halo radius doubles after a violated check) */
  const uint32_t* d_gid;     /* [n_owned] global ids, ascending                                */
  const float* d_points;     /* [n_owned] float4                                               */
  const int32_t* d_knn;      /* [n_owned * k] global ids of the neighbours, canonical (d2, id) order, -1 padded */
  const float* d_d2;         /* [n_owned * k]                                                  */
  const float* d_normals;    /* [n_owned] float4 {nx,ny,nz,curvature}, NULL when not requested */
  double halo_radius;        /* the radius the accepted exchange used                          */
  double kth_distance_max;   /* largest k-th neighbour distance over the WHOLE cloud: the smallest radius that
                                passes the check (a caller that segments similar clouds repeatedly passes a
                                little more than this to its next call)                        */
  uint64_t halo_bytes_sent;      /* halo copies this rank sent to its peers (points + ids)     */
  uint64_t exchange_bytes_sent;  /* everything it sent (owner redistribution + halo)           */
  float ms_estimate;         /* CUDA-event times of the last accepted exchange: halo-radius estimate, */
  float ms_partition;        /* bbox / histogram / owner table / route / pack,                 */
  float ms_exchange;         /* the grouped send / recv,                                       */
  float ms_compute;          /* grid + k-NN + normals on owned + halo points,                  */
  float ms_export;           /* rows to global ids + exactness check                           */
  float knn_cell_ms;         /* the k-NN fast-path kernel inside ms_compute                    */
  int peer_to_peer;          /* bit 0: the exchange was written by the pack kernel straight into the peers' receive
                                buffers over NVLink (cudaIpc / peer access) instead of ncclSend / ncclRecv; bit 1:
                                the rows went straight into the root's buffers (iftwt_sharded_segment) — the d_*
                                pointers above are NULL then.  Both need one earlier call on the comm (the first
                                one sizes the buffers and runs over NCCL); IFTWT_SHARD_P2P=0 keeps everything on NCCL */
} iftwt_shard_out;

/* `in` / `out`: one entry per LOCAL shard (1 in the process-per-GPU mode).  halo_radius <= 0: estimated
 * (the k-th neighbour distance inside a slice bounds the cloud's from above).  Exactness does not
 * depend on the radius: a row whose k-th distance exceeds it is detected (ncclAllReduce) and the
 * exchange is redone with twice the radius.  Collective: every rank of the comm must call it. */
IFTWT_API int iftwt_sharded_knn_normals(iftwt_comm* comm, const iftwt_shard_in* in, int k, int want_normals,
                                        double halo_radius, iftwt_shard_out* out);
/* BASELINE config 4 in one call: sharded k-NN + normals, rows / normals / points gathered to rank
 * `root` (grouped ncclSend / ncclRecv), then graph, region-growing seeds and the IFT on root's GPU.
 * cost / label (HOST, n_total entries in global id order, may be NULL) and *n_seeds are written by the
 * process that holds `root` only.  Needs k_normals == k_graph == rg.number_of_neighbours. */
IFTWT_API int iftwt_sharded_segment(iftwt_comm* comm, const iftwt_shard_in* in, const iftwt_segment_params* params,
                                    double halo_radius, int root, uint32_t* cost, uint32_t* label, size_t* n_seeds,
                                    iftwt_shard_out* out);

/* ---- colour path (graco.hpp, main.cpp:16-89): the cloud's 4th channel carries PCL's packed
 * rgb float (pcl::PointXYZRGB: iftwt_set_cloud(points, 32, n, 0, 16)) -------------------- */

/* GraCo::morph_erode / morph_dilate (graco.hpp:301-351) on the current graph: per vertex the
 * colour (0x00RRGGBB) of the closed-neighbourhood vertex with the smallest (erode, op 4) or
 * largest (dilate, op 3) r+g+b; ties go to the smallest vertex id. */
IFTWT_API int iftwt_morph_rgb(iftwt_ctx* ctx, int op, uint32_t* rgb_out);
/* cloudGray / cloudRGB2GRAY (main.cpp:16-76): replace the packed colour by its 8-bit grey
 * (cv::cvtColor RGB2GRAY); with otsu != 0 binarise at the Otsu threshold (main.cpp:44-51),
 * returned in *threshold.  The cloud becomes an XYZI cloud for the stages that follow. */
IFTWT_API int iftwt_rgb_to_gray(iftwt_ctx* ctx, int otsu, double* threshold);

/* Instrumentation (the reference's is one std::time pair, ift.hpp:172-173). */
IFTWT_API int iftwt_stage_ms(iftwt_ctx* ctx, int stage, float* ms); /* CUDA-event time of the last run of a stage */
IFTWT_API int iftwt_get_stats(iftwt_ctx* ctx, iftwt_stats* out);
IFTWT_API int iftwt_reset_counters(iftwt_ctx* ctx);
IFTWT_API int iftwt_synchronize(iftwt_ctx* ctx);
IFTWT_API void* iftwt_stream(iftwt_ctx* ctx); /* the ctx's cudaStream_t, for callers that record their own events */

#ifdef __cplusplus
}
#endif
#endif /* IFTWT_H */
