/*
 * dsm.h — C ABI of the B200-native DenseSurfelMapping per-frame hot path.
 *
 * Drop-in boundary: this library replaces exactly one call site of the reference,
 *
 *     SurfelMap::fuse_map -> fusion_functions.fuse_initialize_map(...)   surfel_map.cpp:1066-1073
 *     SurfelMap::SurfelMap -> fusion_functions.initialize(...)            surfel_map.cpp:53
 *
 * i.e. the public surface of `class FusionFunctions` (fusion_functions.h:84-94).  Everything
 * behind it (superpixel extraction, back-projection + normals, robust per-superpixel plane
 * fit, surfel associate / weighted-fuse / initialise; fusion_functions.cpp:30-975) runs as
 * hand-written sm_100a CUDA kernels.  There is NO CPU fallback: every entry point returns
 * DSM_E_CUDA / DSM_E_NODEVICE when no B200-class device is usable.
 *
 * Conventions: extern "C", plain pointers and sizes, int return (0 = DSM_OK, negative = error),
 * no exceptions cross the boundary, one context per GPU, calls on one context are serialised
 * by the caller (the reference is single-threaded at this call site too: ros_node.cpp:38-41).
 * Element layouts are byte-identical to the reference PODs (elements.h:5-31).
 *
 * Numerical contract: superpixel labels and the clustering state (seed x, y, intensity, Huber mean depth, stable flag)
 * are bit-identical to the serialised reference; the plane fit and everything downstream of it (seed normal / position /
 * view_cos / size, surfel geometry) agree within 1e-4 (norm-based) but are NOT bit-identical: their float / fp64 sums are
 * accumulated lane-partially and tree-reduced instead of sequentially (fusion_functions.cpp:117-119, :849-860).  A frame's
 * results do not depend on the batch size, on chunking or on which entry point delivered it: every entry point runs the
 * same kernels.
 *
 * INTEGRATION.md shows the reference-side binding (a FusionFunctions-compatible C++ adapter,
 * include/dsm_fusion_functions.hpp, plus the ctypes stub used by the tests).
 */
#ifndef DSM_H
#define DSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSM_VERSION 200 /* 0.2.0: tile schedule (TMA-staged kernels), multi-GPU gather in the C ABI, runtime constant sets */

/* ---- error codes ---- */
#define DSM_OK 0
#define DSM_E_INVALID (-1)   /* bad argument (null pointer, negative count, batch too large ...) */
#define DSM_E_SHAPE (-2)     /* unsupported image shape: W%8 > 4 or H%8 > 4 (reference UB, fusion_functions.cpp:408-451) */
#define DSM_E_NODEVICE (-3)  /* no CUDA device / device is not sm_100 */
#define DSM_E_CUDA (-4)      /* a CUDA runtime call or kernel failed; see dsm_last_error() */
#define DSM_E_NOMEM (-5)     /* device or host allocation failed */
#define DSM_E_CAPACITY (-6)  /* more surfels than the context was created for */
#define DSM_E_STATE (-7)     /* call sequence error (e.g. download before run) */
#define DSM_E_NCCL (-8)      /* NCCL unavailable or a collective failed */
#define DSM_E_IO (-9)        /* a file could not be opened or written (dsm_write_*) */

/* ---- element types: byte-identical to the reference ---- */

/* reference: elements.h:22-31 `struct SurfelElement` (44 bytes) */
typedef struct dsm_surfel_t
{
    float px, py, pz;   /* world position */
    float nx, ny, nz;   /* world normal */
    float size;         /* radius, metres */
    float color;        /* gray 0..255 stored as float */
    float weight;
    int32_t update_times; /* 0 == dead */
    int32_t last_update;  /* reference KEYFRAME index of the last fuse (surfel_map.cpp:145,161) */
} dsm_surfel_t;

/* reference: elements.h:5-20 `struct Superpixel_seed` (60 bytes) — parity/debug readback only */
typedef struct dsm_seed_t
{
    float x, y;
    float size;
    float norm_x, norm_y, norm_z;
    float posi_x, posi_y, posi_z;
    float view_cos;
    float mean_depth;
    float mean_intensity;
    uint8_t fused;
    uint8_t stable;
    uint8_t _pad[2];
    float min_eigen_value; /* debug fields of the reference; always 0 here */
    float max_eigen_value;
} dsm_seed_t;

/* ---- context parameters ---- */
/* The first eight fields are the arguments of FusionFunctions::initialize
 * (fusion_functions.h:84-87, fusion_functions.cpp:7-28).  The rest size the GPU-resident
 * buffers.  SP_SIZE 8, ITERATION_NUM 3 and MAX_ANGLE_COS 0.1 (fusion_functions.h:8-11) are compiled into the kernels;
 * the four constants the reference ships in two sets ("drive" / "RGBD", fusion_functions.h:12-21) are run-time
 * values: a new context holds the drive set, dsm_set_constants() switches. */
typedef struct dsm_params
{
    int32_t width, height;
    float fx, fy, cx, cy;
    float fuse_far, fuse_near;
    int32_t max_batch;         /* frames processed per dsm_batch_* call (>= 1) */
    int32_t max_local_surfels; /* capacity of the local-surfel pool, summed over the batch */
} dsm_params;

typedef struct dsm_ctx dsm_ctx;

/* ---- lifetime ---- */
int dsm_version(void);
const char *dsm_strerror(int code);
/* last CUDA/NCCL error text recorded on this context (never NULL) */
const char *dsm_last_error(const dsm_ctx *ctx);

/* Replaces FusionFunctions::initialize (fusion_functions.cpp:7-28).  `device` is the CUDA
 * ordinal.  `cuda_stream` may be NULL (the context creates its own non-blocking stream) or a
 * cudaStream_t owned by the caller on which all work of this context is enqueued. */
int dsm_create(const dsm_params *params, int device, void *cuda_stream, dsm_ctx **out);
void dsm_destroy(dsm_ctx *ctx);
int dsm_num_seeds(const dsm_ctx *ctx); /* S = (W/8)*(H/8) */

/* The constants the reference selects at compile time by (un)commenting one of two #define blocks (fusion_functions.h:12-21):
 * HUBER_RANGE (Huber-Newton mean depth :540-547, plane-fit inliers :850 and robust weights :134-169), and BASELINE,
 * DISPARITY_ERROR, MIN_TOLERATE_DIFF (depth tolerance of the surfel association :252-253).  Applies to every later call on
 * the context; a fresh context holds DSM_CONSTANTS_DRIVE. */
typedef struct dsm_constants
{
    double huber_range, baseline, disparity_error, min_tolerate_diff;
} dsm_constants;
#define DSM_CONSTANTS_DRIVE {0.4, 0.5, 4.0, 0.1}   /* fusion_functions.h:13-16, the set the reference ships enabled (KITTI) */
#define DSM_CONSTANTS_RGBD {0.05, 0.08, 1.0, 0.05} /* fusion_functions.h:18-21, the commented RGB-D / VINS set */
int dsm_set_constants(dsm_ctx *ctx, const dsm_constants *constants);

/* dsm_batch_run executes a resident batch as `sub_batches` (1..4) groups of frames on concurrent CUDA streams; results do
 * not depend on it (frames are independent).  Default 2: most kernels of the schedule are issue- or latency-bound, two
 * concurrent half-batches fill each other's tails.  1 = the whole batch as one launch sequence (what per-kernel timing
 * with dsm_profile_* should use: concurrent kernels share the SMs and stretch each other's event-pair durations). */
int dsm_set_concurrency(dsm_ctx *ctx, int sub_batches);

/* ---- reference-identical single-frame call ----
 * Replaces FusionFunctions::fuse_initialize_map (fusion_functions.cpp:30-83) with the same
 * semantics: `gray` is CV_8UC1 (row pitch gray_pitch bytes), `depth` CV_32FC1 metres (row
 * pitch depth_pitch bytes), `pose_colmajor` = Eigen::Matrix4f T_world<-cam memory order,
 * `local[n_local]` is updated IN PLACE and never resized (dead surfels are flagged
 * update_times == 0, the caller compacts: surfel_map.cpp:1077-1109), `new_out` receives the
 * newly initialised surfels in seed-index order (the reference clears and push_backs,
 * fusion_functions.cpp:320,359); at most new_cap are written, *n_new is the full count.
 * Host pointers; uploads, runs all kernels, downloads, synchronises.
 * Input domain: depth values are metres, <= 0.01 means "no measurement"; valid depths must be >= 0.02 m
 * (below that the reference itself indexes out of bounds). */
int dsm_fuse_frame(dsm_ctx *ctx, int reference_frame_index,
                   const uint8_t *gray, size_t gray_pitch,
                   const float *depth, size_t depth_pitch,
                   const float pose_colmajor[16],
                   dsm_surfel_t *local, int n_local,
                   dsm_surfel_t *new_out, int new_cap, int *n_new);

/* ---- batch of independent frames (BASELINE configs 3/4) ----
 * Frame b has its own image pair, pose, reference index and its own slice of the local pool:
 * local[local_offsets[b] .. local_offsets[b+1]).  gray/depth are tightly packed
 * [n_frames][H][W].  new_out is [n_frames][S] (S = dsm_num_seeds), n_new is [n_frames].
 * The three stages are exposed separately so that a caller (and bench.py) can keep inputs
 * resident in HBM and time the kernels alone; dsm_fuse_batch() chains them and synchronises. */
int dsm_batch_upload(dsm_ctx *ctx, int n_frames, const int32_t *reference_frame_index,
                     const uint8_t *gray, const float *depth, const float *poses_colmajor,
                     const dsm_surfel_t *local, const int32_t *local_offsets);
int dsm_batch_run(dsm_ctx *ctx);     /* enqueue K0..K6 for the uploaded batch; asynchronous */
int dsm_batch_download(dsm_ctx *ctx, dsm_surfel_t *local_out, dsm_surfel_t *new_out, int32_t *n_new);
int dsm_sync(dsm_ctx *ctx);          /* wait for the context's stream */
int dsm_fuse_batch(dsm_ctx *ctx, int n_frames, const int32_t *reference_frame_index,
                   const uint8_t *gray, const float *depth, const float *poses_colmajor,
                   dsm_surfel_t *local, const int32_t *local_offsets,
                   dsm_surfel_t *new_out, int32_t *n_new);
/* The same call split in two so that a caller can double-buffer: dsm_fuse_batch_async() enqueues the copies
 * and kernels and returns; every host buffer must stay valid and untouched until dsm_batch_wait() returns.
 * One batch per context may be in flight; with two contexts the H2D copies of batch k+1 overlap the kernels
 * of batch k completely (what bench.py's e2e leg does). */
int dsm_fuse_batch_async(dsm_ctx *ctx, int n_frames, const int32_t *reference_frame_index,
                         const uint8_t *gray, const float *depth, const float *poses_colmajor,
                         dsm_surfel_t *local, const int32_t *local_offsets,
                         dsm_surfel_t *new_out, int32_t *n_new);
int dsm_batch_wait(dsm_ctx *ctx);
/* Re-upload only the local pool (the kernels update it in place, so a benchmark that replays
 * the same batch must restore it between steps). Device-to-device from an internal snapshot
 * taken at the last dsm_batch_upload. */
int dsm_batch_restore_pool(dsm_ctx *ctx);

/* ---- GPU-resident pool for a sequential stream ----
 * The pool of ONE stream lives on the device between frames (frame slot 0 of the context), so per
 * frame only the image pair and the pose cross PCIe.  Semantics mirror what SurfelMap does around
 * the hot-path call:
 *   dsm_pool_upload            local_surfels := given array
 *   dsm_fuse_frame_resident    fuse_initialize_map on the resident pool, THEN the post-step of
 *                              SurfelMap::fuse_map (surfel_map.cpp:1077-1109): dead surfels dropped,
 *                              new surfels added.  The resulting pool equals the reference's as a
 *                              SET (the reference's slot recycling order is not reproduced).
 *                              *n_new (optional) = number of new surfels; passing it synchronises.
 *   dsm_pool_transform         warp_active_surfels_cpu_kernel (surfel_map.cpp:750-789):
 *                              p <- W p, n <- R_W n for every pool surfel after a loop closure
 *                              (W = T_loop * T_cam^-1 computed by the caller in fp64, cast to f32).
 *   dsm_pool_retire            the removal half of SurfelMap::move_add_surfels (surfel_map.cpp:1479-1497):
 *                              the live surfels whose last_update == keyframe_index are copied out in
 *                              pool order (they become that pose's attached_surfels) and flagged dead
 *                              (update_times = 0; the next fuse post-step drops them).  Counts first: if more than
 *                              `cap` surfels match, nothing is flagged, *n_out is the number needed and the call
 *                              returns DSM_E_CAPACITY.  Synchronises.
 *   dsm_pool_append            the insertion half (surfel_map.cpp:1583-1587): surfels of poses that
 *                              re-enter the drift-free set are appended to the pool.  Synchronises.
 *   dsm_pool_size / dsm_pool_download   read back (synchronise). */
int dsm_pool_upload(dsm_ctx *ctx, const dsm_surfel_t *local, int n_local);
int dsm_fuse_frame_resident(dsm_ctx *ctx, int reference_frame_index,
                            const uint8_t *gray, size_t gray_pitch, const float *depth, size_t depth_pitch,
                            const float pose_colmajor[16], int *n_new);
/* n consecutive frames of the SAME stream in one call (tightly packed [n][H][W] images, [n][16] poses): results are
 * those of n calls of dsm_fuse_frame_resident (same kernels, same order on the pool), but the pose- and pool-independent stages (superpixels, normals, plane fit) of all n frames run as one
 * batch; only fuse / initialise / compaction run frame by frame.  Trades
 * n-1 frames of latency for throughput (offline sequences, or a node that lags behind its camera).  n <= max_batch;
 * with 2n <= max_batch consecutive runs use alternate halves of the frame slots: the copy and the batched stages of one
 * run overlap the frame-by-frame stages of the previous one.  n_new (optional, [n]) = new surfels per frame; passing it
 * synchronises. */
int dsm_fuse_stream_resident(dsm_ctx *ctx, int n_frames, const int32_t *reference_frame_index,
                             const uint8_t *gray, const float *depth, const float *poses_colmajor, int32_t *n_new);
int dsm_pool_transform(dsm_ctx *ctx, const float W_colmajor[16]);
int dsm_pool_retire(dsm_ctx *ctx, int keyframe_index, dsm_surfel_t *out, int cap, int *n_out);
int dsm_pool_append(dsm_ctx *ctx, const dsm_surfel_t *surfels, int n);
int dsm_pool_size(dsm_ctx *ctx, int *n_local);
int dsm_pool_download(dsm_ctx *ctx, dsm_surfel_t *out, int cap, int *n_local);

/* ---- output side: point clouds and files ----
 * pcl::PointXYZI as the reference publishes and saves it, without PCL's SSE padding: 16 bytes. */
typedef struct dsm_point_t
{
    float x, y, z;
    float intensity; /* SurfelElement::color, gray 0-255 */
} dsm_point_t;
/*   dsm_pool_export_cloud      the cloud builders of SurfelMap over local_surfels: every pool surfel with
 *                              update_times >= min_update_times becomes one point {px, py, pz, color}, in pool
 *                              order like the serial push_back loops.  min_update_times = 5:
 *                              publish_active_pointcloud (surfel_map.cpp:1398-1417), the local part of
 *                              publish_all_pointcloud (:1419-1454) and save_cloud (:1153-1173); = 1: the local part
 *                              of publish_neighbor_pointcloud (:1284-1300, "update_times == 0 -> skip").  Filter and
 *                              compaction run on the device; only the selected 16-byte points cross PCIe.
 *                              *n_out = number selected (may exceed cap; min(cap, *n_out) points are written).
 *   dsm_pool_export_surfels    the same selection returning whole surfels (what save_mesh iterates, :1241-1247).
 *   Both synchronise and leave the pool untouched.
 *   dsm_write_pcd              pcl::io::savePCDFile(name, cloud) of surfel_map.cpp:1171 (PCD v0.7, fields
 *                              x y z intensity; binary = 0 is what the reference writes).  Host only.
 *   dsm_write_ply_mesh         SurfelMap::save_mesh (:1229-1280): one hexagon (6 vertices, 4 triangles) per
 *                              surfel via push_a_surfel (:1175-1226), ASCII PLY.  The caller passes the attached
 *                              surfels of every pose followed by dsm_pool_export_surfels(ctx, 5, ...).  Host only.
 *   dsm_mesh_vertices          push_a_surfel alone: 36 floats (6 x {x, y, z, c, c, c}) per surfel. */
int dsm_pool_export_cloud(dsm_ctx *ctx, int min_update_times, dsm_point_t *out, int cap, int *n_out);
int dsm_pool_export_surfels(dsm_ctx *ctx, int min_update_times, dsm_surfel_t *out, int cap, int *n_out);
int dsm_write_pcd(const char *path, const dsm_point_t *points, size_t n, int binary);
int dsm_write_ply_mesh(const char *path, const dsm_surfel_t *surfels, size_t n);
int dsm_mesh_vertices(const dsm_surfel_t *surfels, size_t n, float *vertices36);

/* ---- inactive store ----
 * What SurfelMap keeps per pose outside the drift-free window -- PoseElement::attached_surfels and the mirror
 * inactive_pointcloud (surfel_map.h:36-46) -- resident on the device next to the local pool, so that
 * move_add_surfels and the loop-closure warp of the inactive surfels never cross PCIe.  Host orchestration over the
 * pool kernels; segments are kept dense in retirement order.
 *   dsm_inactive_reserve      capacity in surfels (once, while the store is empty)
 *   dsm_inactive_retire       move_add_surfels removal loop (surfel_map.cpp:1479-1497): local surfels with
 *                             update_times > 0 and last_update == keyframe_index become that pose's segment (pool
 *                             order) and are flagged dead in the pool.  Synchronises; *n_moved optional.
 *   dsm_inactive_reactivate   insertion (surfel_map.cpp:1583-1587): the pose's segment returns to the end of the pool
 *   dsm_inactive_transform    warp_inactive_surfels_cpu_kernel for one pose (surfel_map.cpp:704-733), asynchronous
 *   dsm_inactive_export_cloud inactive_pointcloud as the node publishes / saves it (store order)
 *   dsm_inactive_download     attached_surfels of one pose (keyframe_index >= 0) or of all poses (< 0; save_mesh)
 *   dsm_inactive_size         surfels / segments currently stored */
int dsm_inactive_reserve(dsm_ctx *ctx, int max_inactive_surfels);
int dsm_inactive_retire(dsm_ctx *ctx, int keyframe_index, int *n_moved);
int dsm_inactive_reactivate(dsm_ctx *ctx, int keyframe_index, int *n_moved);
int dsm_inactive_transform(dsm_ctx *ctx, int keyframe_index, const float W_colmajor[16]);
int dsm_inactive_export_cloud(dsm_ctx *ctx, dsm_point_t *out, int cap, int *n_out);
int dsm_inactive_download(dsm_ctx *ctx, int keyframe_index, dsm_surfel_t *out, int cap, int *n_out);
int dsm_inactive_size(dsm_ctx *ctx, int *n_surfels, int *n_segments);

/* ---- multi-GPU: frames of a batch sharded across GPUs, ONE gather of the surfel deltas at the end ----
 * One context per GPU (one process or thread each).  The reference has no counterpart (fuse_initialize_map is
 * single-process, fusion_functions.cpp:30-83); what is gathered is what SurfelMap::fuse_map consumes after the call
 * (surfel_map.cpp:1077-1109): every frame's new surfels and updated local surfels.
 *   dsm_comm_unique_id   rank 0 creates the 128-byte NCCL id and hands it to the other ranks out of band
 *   dsm_comm_init        collective: joins the communicator (ncclCommInitRank); NCCL is dlopen'ed ("libnccl.so.2": inside a
 *                        PyTorch process the copy torch already loaded), DSM_E_NCCL if it is missing
 *   dsm_gather_deltas    collective, after dsm_batch_run: packs ONLY the valid records of this rank's batch
 *                          int32 'DSMD', n_frames, n_new_total, n_pool_total, n_new[n_frames], pool_ofs[n_frames+1], pad to 16 B,
 *                          dsm_surfel_t new[n_new_total] (frame by frame, seed-index order), dsm_surfel_t pool[n_pool_total]
 *                        and moves it to `root` on a side stream.  Ranks of one node write their payload straight into
 *                        their slot of the root's receive buffer over NVLink peer memory (CUDA IPC mapping made at the first
 *                        gather; counts stay on the device, nothing waits on the host, exactly the valid bytes travel);
 *                        without a peer mapping (or with DSM_GATHER_NCCL=1) the counts go to the host and the payloads move
 *                        with ncclAllGather (sizes) + one grouped ncclSend/ncclRecv.  Collective: every rank of the
 *                        communicator calls it with the same root, in the same order.  Returns before the transfer has
 *                        finished; the next dsm_batch_run may be enqueued at once.  A new gather overwrites the previous one.
 *   dsm_gather_wait      root: waits until every rank's payload has arrived; other ranks: until their own has left
 *   dsm_gathered_*       root only: the payload of every rank, on the device (zero copy; payload r occupies
 *                        [rank_offsets[r], rank_offsets[r] + dsm_gathered_rank_bytes(r))) or copied to host memory */
#define DSM_COMM_ID_BYTES 128
int dsm_comm_unique_id(void *id_out128);
int dsm_comm_init(dsm_ctx *ctx, const void *id128, int rank, int nranks);
int dsm_comm_destroy(dsm_ctx *ctx);
int dsm_gather_deltas(dsm_ctx *ctx, int root);
int dsm_gather_wait(dsm_ctx *ctx);
int dsm_gathered_device(dsm_ctx *ctx, void **dev_ptr, size_t *rank_offsets /* [nranks + 1] */);
int dsm_gathered_rank_bytes(dsm_ctx *ctx, int rank, size_t *bytes);
int dsm_gathered_download(dsm_ctx *ctx, int rank, void *host_out, size_t cap);

/* ---- parity / debug readback (what the reference keeps private: fusion_functions.h:34-37) ---- */
int dsm_get_labels(dsm_ctx *ctx, int frame, int32_t *labels_hw);   /* superpixel_index, [H][W] */
int dsm_get_seeds(dsm_ctx *ctx, int frame, dsm_seed_t *seeds_s);   /* superpixel_seeds, [S] */

/* debug: enqueue only the first n kernels of the schedule on the next dsm_batch_run (n <= 0: all).
 * Lets the parity tests localise a mismatch to one pass; never used on the product path. */
int dsm_debug_stop_after(dsm_ctx *ctx, int n_kernels);
/* debug: number of "cannot happen" events the kernels counted in the last batch (a non-stable seed
 * without members); the parity tests assert it is 0. */
int dsm_debug_invariant_violations(dsm_ctx *ctx, int *count);

/* ---- measurement hooks ----
 * Per-kernel CUDA-event timing on the context's stream.  mask selects kernels (bit k = kernel
 * id k, see dsm_kernel_name); 0 disables.  Accumulates until dsm_profile_reset(). */
#define DSM_NUM_KERNELS 10
int dsm_profile_enable(dsm_ctx *ctx, uint32_t kernel_mask);
int dsm_profile_reset(dsm_ctx *ctx);
/* resolves pending events (synchronises); ms_total/launches are [DSM_NUM_KERNELS] */
int dsm_profile_read(dsm_ctx *ctx, float *ms_total, int32_t *launches);
const char *dsm_kernel_name(int kernel_id);
/* raw device pointers for zero-copy interop (e.g. the multi-GPU gather); which: see DSM_BUF_* */
#define DSM_BUF_NEW_SURFELS 0 /* dsm_surfel_t [max_batch][S] */
#define DSM_BUF_NEW_COUNTS 1  /* int32 [max_batch] */
#define DSM_BUF_LOCAL 2       /* dsm_surfel_t [max_local_surfels] */
int dsm_device_buffer(dsm_ctx *ctx, int which, void **dev_ptr, size_t *bytes);

#ifdef __cplusplus
}
#endif
#endif /* DSM_H */
