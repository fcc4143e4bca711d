// Device-side data layout shared by the kernels (dsm_kernels.cu) and the host driver (dsm_capi.cu).
//
// HBM layout per context (B = max_batch frames, P = H*Wp pitched pixels, S = seeds per frame):
//   gray    u8  [B][H][Wp]      Wp = W rounded up to 16 so every row starts 16-byte aligned
//   depth   f32 [B][H][Wp]      (vector loads, TMA boxes)
//   invd    f32 [B][H][Wp]      (float)(1.0 / (double)depth) for depth > 0.01, else 0 (:404-405): what the assign passes read
//   labels  i32 [B][H][Wp]      superpixel_index of the reference (fusion_functions.h:37)
//   code    u8  [B][H][Wp]      the same label as the index of the winning candidate among the pixel's 2x2 candidate seeds:
//                               what the window gathers test (1 B/px, a byte compare against a position-only pattern)
//   seed    float4 [B][S]       (x, y, mean_intensity, mean_depth)  -- the clustering state
//   seed_hl float2 [B][S]       1.0 / (double)mean_depth as hi + lo floats (fp32 cost filter of the assign pass)
//   inv_md  f64 [B][S]          1.0 / (double)mean_depth (exact path of the assign pass, :380)
//   tstable i32 [B][S]          stable flag + raster time stamp: INT_MAX = stable,
//                               -1 = unstable, k >= 0 = un-stabled when raster scan reached k
//   usum    int4 [B][S], und i32 [B][S]   update_seeds integer sums / depth-list lengths
//   dlist   f32 [B][S][232]     member depths in raster order (k_gather -> k_newton2), one contiguous list per seed
//   done    i32 [B]             frame-completion tickets of the assign pass
//   errflag i32 [B]             invariant violations (must stay 0)
//   kx, ky  f32 [Wp+16], [H+16] back-projection factors, computed once per context
//   qlist   f32 [B][S][3][232]  centred plane-fit inlier points (k_plane_gather -> k_gn_solve fallback passes)
//   hrec    f64 [B][24][S]      first Gauss-Newton pass sums of the plane fit, field-major
//   pfsum   float4 [B][S][2]    per-seed plane-fit summary
//   plane   float4 [B][S][3]    (n.xyz, view_cos) (posi.xyz, mean_depth) (size, I, x, y)
//   fused   i32 [B][S]          Superpixel_seed::fused
//   list    int2 [B][P]         deferred pixels of the stable relaxation: (pitched raster index, winner)
//   nlist   i32 [B]
//   pool    dsm_surfel_t [max_local]   AoS, ABI layout (44 B) so upload/download are plain copies
//   poolofs i32 [B+1]
//   newsurf dsm_surfel_t [B][S], nnew i32 [B]
//   pose/ipose f32 [B][16] column-major, refidx i32 [B]
#pragma once
#include <cstdint>
#include <cuda.h> // CUtensorMap (types only; the encoder is resolved through the runtime, no libcuda link dependency)
#include <cuda_runtime.h>
#include "../../include/dsm.h"

#define DSM_SP 8
#define DSM_STABLE 0x7fffffff

struct DsmDev
{
    // geometry
    int W, H, Wp, spw, sph, S, B;
    int frame0; // first frame slot this launch works on (chunked copy/compute overlap)
    int Sp; // S rounded up to 32: row stride of the [k][seed] scratch lists
    float fx, fy, cx, cy, fuse_far, fuse_near, camera_f;
    // algorithm constants (dsm_set_constants; fusion_functions.h:12-21).  huber_hi = smallest float >= huber: for every float x,
    // (double)x < huber  <=>  x < huber_hi  and  (double)x >= huber  <=>  x >= huber_hi (the float-vs-double-literal compares)
    double huber, baseline, disparity_error, min_tolerate_diff;
    float huber_hi;
    // per-frame strides in elements
    size_t px_stride; // H*Wp
    // buffers
    const uint8_t *gray;
    const float *depth;
    int32_t *labels;
    float4 *seed;
    double *inv_md;
    int32_t *tstable;
    int4 *usum;         // update_seeds integer sums per seed: (count, sum x, sum y, sum intensity)
    int32_t *und;       // number of member depths > 0.1
    float *dlist;       // [B][S][232] member depths in raster order, one contiguous list per seed
    int32_t *errflag;   // [B] count of "impossible" events (non-stable seed without members)
    float4 *plane;
    const float *kx;    // [Wp+16]: ((float)u - cx) / fx, the per-column factor of back_project (:94)
    const float *ky;    // [H+16]:  ((float)v - cy) / fy
    float *qlist;       // [B][S][3][232] centred plane-fit inlier points
    float4 *pfsum;      // [B][S][2]: (sum n.xyz, max_dist) (mean.xyz, inlier count or 0 if rejected)
    int32_t *fused;
    int2 *list;
    int32_t *nlist;
    dsm_surfel_t *pool;
    const int32_t *poolofs;
    dsm_surfel_t *newsurf;
    int32_t *nnew;
    const float *pose;
    const float *ipose;
    const int32_t *refidx;
    int max_pool_per_frame; // largest per-frame pool slice in this batch (grid sizing)
    // tile path (dsm_tile.cu)
    uint8_t *code;      // [B][H][Wp] label of every pixel as the index (0..3) of the winning candidate among its 2x2 candidate seeds (mirror of labels)
    float *invd;        // [B][H][Wp] (float)(1.0 / (double)depth) for depth > 0.01, else 0 (:404-405), written by the first assign pass
    float2 *seed_hl;    // [B][S] 1.0 / (double)mean_depth split into two floats (hi, lo) for the filtered assign pass
    int32_t *hardq;     // [B][S] seeds whose Huber-Newton needs the entry-by-entry classification (k_newton2 -> k_newton_hard)
    int32_t *nhard;     // [B] queue lengths
    int32_t *done;      // [B] frame-completion tickets of the assign pass (the last CTA of a frame runs the stable relaxation)
    double *hrec;       // [B][24][S] plane fit, first residual pass: H = sum 2 q q^T (9), the same over out-of-range points (10), their clamped gradient (4), packed (margin, qmax2)
};

// TMA descriptors (cuTensorMapEncodeTiled) of the three per-pixel arrays as [B][H][Wp] tensors; box = one seed tile plus halo
struct DsmMaps
{
    CUtensorMap cod, dep, gry; // label codes (u8), depth (f32), gray (u8)
};
// seed tile of the tile kernels: 8 x 4 superpixels = 64 x 32 pixels, plus a 4-pixel halo (the 16 x 16 windows of
// update_seeds_kernel :481-489 and calculate_sp_depth_norms_kernel :806-811 overlap their neighbours by 8)
#define DSM_TILE_SX 8
#define DSM_TILE_SY 4
#define DSM_TILE_W 76  // box width in elements: 72 used; row pitch 76 words = 12 mod 32, so a quarter-warp reading 16 bytes per lane from 8 consecutive rows hits 32 distinct banks
#define DSM_TILE_H 41  // box height: 40 rows + the down neighbour of the last row (pixel normals)
// gray (u8) box: TMA needs the box start 16-byte aligned in global memory, which 64 bx - 4 is for the 4-byte arrays but
// not for bytes, so the gray box starts DSM_TILE_GX = 12 pixels further left (64 bx - 16) and is 112 bytes wide
#define DSM_TILE_GW 112
#define DSM_TILE_GX 12

enum DsmKernelId
{
    DSM_K_SEED_INIT = 0,
    DSM_K_ASSIGN_FIRST = 1,
    DSM_K_ASSIGN = 2,
    DSM_K_GATHER = 3,       // window gather of update_seeds_kernel
    DSM_K_NEWTON = 4,       // Huber-Newton solve of update_seeds_kernel
    DSM_K_PLANE_GATHER = 5, // pixel normals + plane-fit gather
    DSM_K_PLANE_SOLVE = 6,  // plane-fit solver
    DSM_K_FUSE = 7,
    DSM_K_INIT_SURFELS = 8,
    DSM_K_REPACK = 9,
};

// launchers (dsm_kernels.cu); nb = frames in this batch
void dsm_launch_repack(const DsmDev &d, int nb, const uint8_t *gray_packed, const float *depth_packed, cudaStream_t s);
void dsm_launch_seed_init(const DsmDev &d, int nb, cudaStream_t s);
void dsm_launch_fuse(const DsmDev &d, int nb, cudaStream_t s);
void dsm_launch_init_surfels(const DsmDev &d, int nb, cudaStream_t s);
void dsm_launch_seeds_export(const DsmDev &d, int frame, dsm_seed_t *out_dev, int raw_md, cudaStream_t s);
void dsm_launch_pool_compact(const DsmDev &d, int frame, int upper, int *blkcnt, int *blkofs, int *newofs, dsm_surfel_t *dst, int32_t *ofs_out, cudaStream_t s);
void dsm_launch_pool_transform(const DsmDev &d, int frame, int upper, const float *Wm_dev, cudaStream_t s);
void dsm_launch_pool_export(const DsmDev &d, int frame, int upper, int mode, int key, bool as_cloud, int *blkcnt, int *blkofs, int *newofs, void *dst, cudaStream_t s);
void dsm_launch_set2(int32_t *p, int a, int b, cudaStream_t s);
// tile path (dsm_tile.cu)
int dsm_tile_setup(); // raises the dynamic shared-memory limits of the tile kernels (once per process and device)
void dsm_launch_assign2(const DsmDev &d, int nb, bool first, cudaStream_t s);
void dsm_launch_gather(const DsmDev &d, const DsmMaps &m, int nb, cudaStream_t s);
void dsm_launch_newton2(const DsmDev &d, int nb, cudaStream_t s);
void dsm_launch_plane_gather(const DsmDev &d, const DsmMaps &m, int nb, cudaStream_t s);
void dsm_launch_gn_solve(const DsmDev &d, int nb, cudaStream_t s);
void dsm_launch_pool_retire(const DsmDev &d, int frame, int upper, int key, int *blkcnt, int *blkofs, int *newofs, dsm_surfel_t *dst, cudaStream_t s);
