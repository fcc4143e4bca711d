// Private to the library: the context behind the opaque `dsm_ctx` of include/dsm.h, shared by the host-side
// translation units (dsm_capi.cu: single-GPU entry points; dsm_comm.cu: multi-GPU gather over NCCL).
#pragma once
#include "dsm_device.cuh"
#include <cstdio>
#include <vector>

struct DsmComm; // multi-GPU state (dsm_comm.cu), nullptr until dsm_comm_init

struct ProfRec
{
    int id;
    cudaEvent_t e0, e1;
};

// A captured kernel schedule, replayable while every launch parameter is unchanged.
struct GraphEntry
{
    int f0, nf, maxper, compact;
    const void *pool, *poolofs, *alt;
    cudaGraphExec_t exec;
    int phase;
};

struct dsm_ctx
{
    dsm_params p;
    int device;
    cudaStream_t stream;
    bool own_stream;
    DsmComm *comm;
    DsmDev d;
    DsmMaps maps;  // TMA descriptors of labels / depth / gray for the tile kernels
    int S, Wp;
    size_t px;    // pitched pixels per frame
    int nb;       // frames in the current batch
    int n_pool;   // local surfels in the current batch
    bool uploaded, ran;
    bool in_flight; // a dsm_fuse_batch_async batch has been enqueued and not yet waited for
    int stop_after; // debug: number of kernels to enqueue (<= 0: all)
    std::vector<GraphEntry> graphs; // CUDA-graph cache of the kernel schedule (launch-bound single-frame / small-chunk runs)
    bool use_graphs;
    // raw allocations (non-const views of what DsmDev holds)
    uint8_t *gray;
    float *depth;
    dsm_surfel_t *pool_snap;
    int32_t *poolofs, *refidx;
    float *pose, *ipose;
    dsm_seed_t *seed_export;
    float *kx, *ky;
    // end-to-end pipeline (dsm_fuse_batch): packed staging + copy streams + per-chunk events
    uint8_t *gray_packed; // [B][H][W]
    float *depth_packed;  // [B][H][W]
    cudaStream_t s_h2d, s_d2h, s_comp[4];
    cudaStream_t s_hi;   // highest priority: the frame-by-frame (latency-bound) phase of dsm_fuse_stream_resident
    cudaEvent_t ev_p2;   // end of that phase, joined into the main stream
    // resident pool (stream mode)
    int res_upper;      // host-side upper bound of the resident pool size (exact after a sync)
    bool res_active;
    int res_frame;      // frames fused in resident mode so far (selects the frame slot)
    int32_t *res_ofs;   // device [2]: {0, resident pool size}
    int *blkcnt, *blkofs, *newofs;
    float *wmat;        // device copy of the 4x4 of dsm_pool_transform
    // inactive store (EXPERIMENTAL, dsm_inactive_*): the attached_surfels of every pose outside the drift-free window,
    // dense on the device in retirement order; the (keyframe, offset, count) segment list lives on the host
    dsm_surfel_t *inact;
    int inact_cap, inact_size;
    int32_t *inact_ofs; // device [2]
    struct InactSeg
    {
        int kf, ofs, cnt;
    };
    std::vector<InactSeg> inact_segs;
    cudaEvent_t ev_h2d[8], ev_done[8], ev_start;
    cudaEvent_t ev_fork, ev_join[4]; // dsm_batch_run: fork / join of the concurrent sub-batches
    bool single_pending;             // a dsm_fuse_frame_resident call since the last dsm_fuse_stream_resident (they share frame slots 0 / 1)
    int run_split;                   // sub-batches per dsm_batch_run (dsm_set_concurrency)
    // pinned host staging for the small per-batch tables
    float *h_pose; // [B][32]: pose then inverse
    int32_t *h_ofs;
    int32_t *h_ref;
    // profiling
    uint32_t prof_mask;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> ev_free;
    float prof_ms[DSM_NUM_KERNELS];
    int32_t prof_n[DSM_NUM_KERNELS];
    char err[512];
};

#define CK(call)                                                                                         \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess)                                                                           \
        {                                                                                                \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return DSM_E_CUDA;                                                                           \
        }                                                                                                \
    } while (0)

