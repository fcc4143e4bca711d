// Host side of the C ABI declared in include/dsm.h: context, HBM allocation, the per-frame
// kernel schedule (mirrors FusionFunctions::fuse_initialize_map / generate_super_pixels,
// fusion_functions.cpp:30-83, :960-975) and the copies either side of it.
// No CPU implementation of any phase lives here: if CUDA is unavailable every call fails.
#include "dsm_ctx.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <new>
#include <vector>

static const char *kKernelNames[DSM_NUM_KERNELS] = {"seed_init", "slic_assign_first", "slic_assign", "slic_gather", "slic_newton",
                                                     "plane_gather", "plane_solve", "surfel_fuse", "surfel_init", "repack"};

extern "C" int dsm_version(void) { return DSM_VERSION; }

extern "C" const char *dsm_strerror(int code)
{
    switch (code)
    {
    case DSM_OK: return "ok";
    case DSM_E_INVALID: return "invalid argument";
    case DSM_E_SHAPE: return "unsupported image shape (W%8 or H%8 > 4, or smaller than 24x24)";
    case DSM_E_NODEVICE: return "no usable CUDA device (sm_100 required)";
    case DSM_E_CUDA: return "CUDA error";
    case DSM_E_NOMEM: return "out of memory";
    case DSM_E_CAPACITY: return "surfel capacity exceeded";
    case DSM_E_STATE: return "invalid call sequence";
    case DSM_E_NCCL: return "NCCL error";
    case DSM_E_IO: return "file could not be opened or written";
    default: return "unknown error";
    }
}

extern "C" const char *dsm_last_error(const dsm_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" const char *dsm_kernel_name(int id) { return (id >= 0 && id < DSM_NUM_KERNELS) ? kKernelNames[id] : "?"; }
extern "C" int dsm_num_seeds(const dsm_ctx *ctx) { return ctx ? ctx->S : DSM_E_INVALID; }

template <typename T>
static cudaError_t dmalloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

extern "C" void dsm_destroy(dsm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    dsm_comm_destroy(ctx);
    // an asynchronous batch may still be in flight on any of the context's streams
    if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
    for (int i = 0; i < 4; i++)
        if (ctx->s_comp[i]) cudaStreamSynchronize(ctx->s_comp[i]);
    if (ctx->s_hi) cudaStreamSynchronize(ctx->s_hi);
    if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &r : ctx->prof_pending)
    {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    for (auto &e : ctx->ev_free) cudaEventDestroy(e);
    for (auto &g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    DsmDev &d = ctx->d;
    cudaFree(ctx->gray);
    cudaFree(ctx->depth);
    cudaFree(d.labels);
    cudaFree(d.seed);
    cudaFree(d.inv_md);
    cudaFree(d.invd);
    cudaFree(d.code);
    cudaFree(d.seed_hl);
    cudaFree(d.done);
    cudaFree(d.hardq);
    cudaFree(d.nhard);
    cudaFree(d.hrec);
    cudaFree(d.tstable);
    cudaFree(d.usum);
    cudaFree(d.und);
    cudaFree(d.dlist);
    cudaFree(d.errflag);
    cudaFree(d.qlist);
    cudaFree(d.pfsum);
    cudaFree(d.plane);
    cudaFree(d.fused);
    cudaFree(d.list);
    cudaFree(d.nlist);
    cudaFree(d.pool);
    cudaFree(ctx->pool_snap);
    cudaFree(ctx->poolofs);
    cudaFree(d.newsurf);
    cudaFree(d.nnew);
    cudaFree(ctx->pose);
    cudaFree(ctx->ipose);
    cudaFree(ctx->refidx);
    cudaFree(ctx->seed_export);
    cudaFree(ctx->kx);
    cudaFree(ctx->ky);
    cudaFree(ctx->gray_packed);
    cudaFree(ctx->blkcnt);
    cudaFree(ctx->blkofs);
    cudaFree(ctx->newofs);
    cudaFree(ctx->wmat);
    cudaFree(ctx->inact);
    cudaFree(ctx->inact_ofs);
    cudaFree(ctx->res_ofs);
    cudaFree(ctx->depth_packed);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (int i = 0; i < 4; i++)
        if (ctx->s_comp[i]) cudaStreamDestroy(ctx->s_comp[i]);
    if (ctx->s_hi) cudaStreamDestroy(ctx->s_hi);
    if (ctx->ev_p2) cudaEventDestroy(ctx->ev_p2);
    for (int i = 0; i < 8; i++)
    {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int i = 0; i < 4; i++)
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    cudaFreeHost(ctx->h_pose);
    cudaFreeHost(ctx->h_ofs);
    cudaFreeHost(ctx->h_ref);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// TMA descriptors for the tile kernels: each per-pixel array as a [B][H][Wp] tensor, box = one 8x4-seed tile plus
// halo (DSM_TILE_W x DSM_TILE_H elements; out-of-image parts of a box are zero-filled by the copy engine).
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so the library keeps linking
// against cudart only.
// smallest float >= c: (double)x < c  <=>  x < ceil_float(c) for every float x
static float ceil_float(double c)
{
    float f = (float)c;
    if ((double)f < c) f = nextafterf(f, INFINITY);
    return f;
}
static void apply_constants(DsmDev &d, const dsm_constants &k)
{
    d.huber = k.huber_range;
    d.huber_hi = ceil_float(k.huber_range);
    d.baseline = k.baseline;
    d.disparity_error = k.disparity_error;
    d.min_tolerate_diff = k.min_tolerate_diff;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tile_maps(dsm_ctx *ctx)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
    {
        snprintf(ctx->err, sizeof(ctx->err), "cuTensorMapEncodeTiled is not available in this driver");
        return DSM_E_CUDA;
    }
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    const DsmDev &d = ctx->d;
    struct
    {
        CUtensorMap *map;
        void *base;
        CUtensorMapDataType type;
        unsigned esize, boxw;
    } specs[3] = {{&ctx->maps.cod, d.code, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, DSM_TILE_GW},
                  {&ctx->maps.dep, ctx->depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, DSM_TILE_W},
                  {&ctx->maps.gry, ctx->gray, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, DSM_TILE_GW}};
    for (auto &sp : specs)
    {
        const cuuint64_t dims[3] = {(cuuint64_t)d.Wp, (cuuint64_t)d.H, (cuuint64_t)d.B};
        const cuuint64_t strides[2] = {(cuuint64_t)d.Wp * sp.esize, (cuuint64_t)d.px_stride * sp.esize}; // bytes, multiples of 16 (Wp % 16 == 0)
        const cuuint32_t box[3] = {sp.boxw, DSM_TILE_H, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(sp.map, sp.type, 3, sp.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            snprintf(ctx->err, sizeof(ctx->err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return DSM_E_CUDA;
        }
    }
    return DSM_OK;
}

extern "C" int dsm_create(const dsm_params *params, int device, void *cuda_stream, dsm_ctx **out)
{
    if (!params || !out) return DSM_E_INVALID;
    *out = nullptr;
    const int W = params->width, H = params->height;
    if (params->max_batch < 1 || params->max_local_surfels < 0) return DSM_E_INVALID;
    if (W < 3 * DSM_SP || H < 3 * DSM_SP || W % DSM_SP > 4 || H % DSM_SP > 4) return DSM_E_SHAPE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DSM_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DSM_E_NODEVICE;
    if (prop.major != 10) return DSM_E_NODEVICE; // the only code in this library is sm_100a SASS
    if (cudaSetDevice(device) != cudaSuccess) return DSM_E_NODEVICE;

    dsm_ctx *ctx = new (std::nothrow) dsm_ctx();
    if (!ctx) return DSM_E_NOMEM;
    memset(&ctx->d, 0, sizeof(ctx->d));
    ctx->p = *params;
    ctx->comm = nullptr;
    ctx->device = device;
    ctx->err[0] = 0;
    ctx->prof_mask = 0;
    memset(ctx->prof_ms, 0, sizeof(ctx->prof_ms));
    memset(ctx->prof_n, 0, sizeof(ctx->prof_n));
    ctx->nb = 0;
    ctx->n_pool = 0;
    ctx->uploaded = ctx->ran = false;
    ctx->in_flight = false;
    ctx->stop_after = 0;
    {
        const char *e = getenv("DSM_GRAPHS");
        ctx->use_graphs = !(e && e[0] == '0');
    }
    ctx->s_h2d = ctx->s_d2h = nullptr;
    for (int i = 0; i < 4; i++) ctx->s_comp[i] = nullptr;
    ctx->ev_start = nullptr;
    ctx->ev_fork = nullptr;
    ctx->s_hi = nullptr;
    ctx->ev_p2 = nullptr;
    for (int i = 0; i < 4; i++) ctx->ev_join[i] = nullptr;
    ctx->run_split = 2;
    ctx->single_pending = false;
    for (int i = 0; i < 8; i++) ctx->ev_h2d[i] = ctx->ev_done[i] = nullptr;
    ctx->gray_packed = nullptr;
    ctx->res_upper = 0;
    ctx->res_active = false;
    ctx->res_frame = 0;
    ctx->res_ofs = nullptr;
    ctx->blkcnt = ctx->blkofs = ctx->newofs = nullptr;
    ctx->wmat = nullptr;
    ctx->inact = nullptr, ctx->inact_ofs = nullptr, ctx->inact_cap = ctx->inact_size = 0;
    ctx->depth_packed = nullptr;
    ctx->own_stream = (cuda_stream == nullptr);
    ctx->stream = (cudaStream_t)cuda_stream;
    if (ctx->own_stream && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete ctx;
        return DSM_E_CUDA;
    }
    const int B = params->max_batch;
    const int Wp = (W + 15) / 16 * 16;
    const int spw = W / DSM_SP, sph = H / DSM_SP, S = spw * sph;
    const size_t px = (size_t)H * Wp;
    ctx->S = S;
    ctx->Wp = Wp;
    ctx->px = px;
    DsmDev &d = ctx->d;
    d.W = W, d.H = H, d.Wp = Wp, d.spw = spw, d.sph = sph, d.S = S, d.B = B;
    d.Sp = (S + 31) / 32 * 32;
    d.fx = params->fx, d.fy = params->fy, d.cx = params->cx, d.cy = params->cy;
    d.fuse_far = params->fuse_far, d.fuse_near = params->fuse_near;
    d.camera_f = (float)((fabs((double)params->fx) + fabs((double)params->fy)) / 2.0); // (:250)
    {
        const dsm_constants drive = DSM_CONSTANTS_DRIVE;
        apply_constants(d, drive);
    }
    d.px_stride = px;
    const size_t npool = (size_t)(params->max_local_surfels > 0 ? params->max_local_surfels : 1);
    cudaError_t e = cudaSuccess;
    // +64 bytes/elements of slack so the 16-byte vector accesses on the last pitched row stay in bounds
#define ALLOC(ptr, n)                        \
    if (e == cudaSuccess) e = dmalloc(&(ptr), (size_t)(n))
    ALLOC(ctx->gray, B * px + 64);
    ALLOC(ctx->depth, B * px + 64);
    ALLOC(d.labels, B * px + 64);
    ALLOC(d.seed, (size_t)B * S);
    ALLOC(d.inv_md, (size_t)B * S);
    ALLOC(d.invd, B * px + 64);
    ALLOC(d.code, B * px + 64);
    ALLOC(d.seed_hl, (size_t)B * S);
    ALLOC(d.done, (size_t)B);
    ALLOC(d.hardq, (size_t)B * S);
    ALLOC(d.nhard, (size_t)B);
    ALLOC(d.hrec, (size_t)B * S * 24);
    ALLOC(d.tstable, (size_t)B * S);
    ALLOC(d.usum, (size_t)B * S);
    ALLOC(d.und, (size_t)B * S);
    ALLOC(d.dlist, (size_t)B * S * 232);
    ALLOC(d.errflag, (size_t)B);
    ALLOC(d.qlist, (size_t)B * S * 3 * 232);
    ALLOC(d.pfsum, (size_t)B * S * 2);
    ALLOC(d.plane, (size_t)B * S * 3);
    ALLOC(d.fused, (size_t)B * S);
    ALLOC(d.list, B * px);
    ALLOC(d.nlist, (size_t)B);
    ALLOC(d.pool, npool);
    ALLOC(ctx->pool_snap, npool);
    ALLOC(ctx->poolofs, (size_t)B + 1);
    ALLOC(d.newsurf, (size_t)B * S);
    ALLOC(d.nnew, (size_t)B);
    ALLOC(ctx->pose, (size_t)B * 16);
    ALLOC(ctx->ipose, (size_t)B * 16);
    ALLOC(ctx->refidx, (size_t)B);
    ALLOC(ctx->seed_export, (size_t)S);
    ALLOC(ctx->kx, (size_t)Wp + 16);
    ALLOC(ctx->ky, (size_t)H + 16);
    ALLOC(ctx->gray_packed, (size_t)B * H * W + 64);
    ALLOC(ctx->depth_packed, (size_t)B * H * W + 64);
    ALLOC(ctx->blkcnt, npool / 256 + 64);
    ALLOC(ctx->blkofs, npool / 256 + 64);
    ALLOC(ctx->newofs, 2);
    ALLOC(ctx->wmat, 16);
    ALLOC(ctx->res_ofs, 2);
#undef ALLOC
    d.frame0 = 0;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&ctx->s_comp[i], cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++)
    {
        e = cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_p2, cudaEventDisableTiming);
    if (e == cudaSuccess)
    {
        int least = 0, greatest = 0;
        e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->s_hi, cudaStreamNonBlocking, greatest);
    }
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    if (e == cudaSuccess)
    { // back-projection factor tables: the same float ops as back_project (fusion_functions.cpp:94-95)
        std::vector<float> hx((size_t)Wp + 16), hy((size_t)H + 16);
        for (int u = 0; u < Wp + 16; u++) hx[u] = ((float)u - params->cx) / params->fx;
        for (int v = 0; v < H + 16; v++) hy[v] = ((float)v - params->cy) / params->fy;
        e = cudaMemcpy(ctx->kx, hx.data(), hx.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ctx->ky, hy.data(), hy.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_pose, (size_t)2 * B * 32 * sizeof(float)); // second half: dsm_fuse_stream_resident's own tables
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_ofs, ((size_t)B + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_ref, (size_t)2 * B * sizeof(int32_t));
    if (e != cudaSuccess)
    {
        dsm_destroy(ctx);
        return e == cudaErrorMemoryAllocation ? DSM_E_NOMEM : DSM_E_CUDA;
    }
    d.gray = ctx->gray;
    d.depth = ctx->depth;
    d.poolofs = ctx->poolofs;
    d.pose = ctx->pose;
    d.ipose = ctx->ipose;
    d.refidx = ctx->refidx;
    d.kx = ctx->kx;
    d.ky = ctx->ky;
    d.max_pool_per_frame = 0;
    if (make_tile_maps(ctx) != DSM_OK || dsm_tile_setup() != 0)
    {
        dsm_destroy(ctx);
        return DSM_E_CUDA;
    }
    *out = ctx;
    return DSM_OK;
}

// general 4x4 float inverse (adjugate / determinant) of a column-major matrix: stands in for
// Eigen's `pose.inverse()` (fusion_functions.cpp:59); same formula as the oracle's Eigen stand-in.
static void inverse4f(const float *m, float *out)
{
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float inv_det = 1.0f / det;
    for (int i = 0; i < 16; i++) out[i] = inv[i] * inv_det;
}

// ---- profiling helpers ----
static cudaEvent_t prof_event(dsm_ctx *ctx)
{
    if (!ctx->ev_free.empty())
    {
        cudaEvent_t e = ctx->ev_free.back();
        ctx->ev_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope
{
    dsm_ctx *ctx;
    int id;
    cudaEvent_t e0, e1;
    bool on;
    cudaStream_t st;
    ProfScope(dsm_ctx *c, int k, cudaStream_t stream = nullptr) : ctx(c), id(k), on((c->prof_mask >> k) & 1u), st(stream ? stream : c->stream)
    {
        if (on)
        {
            e0 = prof_event(ctx);
            e1 = prof_event(ctx);
            cudaEventRecord(e0, st);
        }
    }
    ~ProfScope()
    {
        if (on)
        {
            cudaEventRecord(e1, st);
            ctx->prof_pending.push_back(ProfRec{id, e0, e1});
        }
    }
};

extern "C" int dsm_profile_enable(dsm_ctx *ctx, uint32_t mask)
{
    if (!ctx) return DSM_E_INVALID;
    ctx->prof_mask = mask;
    return DSM_OK;
}
static int prof_drain(dsm_ctx *ctx, bool accumulate)
{
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_pending)
    {
        if (accumulate)
        {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
            ctx->prof_ms[r.id] += ms;
            ctx->prof_n[r.id] += 1;
        }
        ctx->ev_free.push_back(r.e0);
        ctx->ev_free.push_back(r.e1);
    }
    ctx->prof_pending.clear();
    return DSM_OK;
}
extern "C" int dsm_profile_reset(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    int rc = prof_drain(ctx, false);
    memset(ctx->prof_ms, 0, sizeof(ctx->prof_ms));
    memset(ctx->prof_n, 0, sizeof(ctx->prof_n));
    return rc;
}
extern "C" int dsm_profile_read(dsm_ctx *ctx, float *ms_total, int32_t *launches)
{
    if (!ctx || !ms_total || !launches) return DSM_E_INVALID;
    int rc = prof_drain(ctx, true);
    if (rc != DSM_OK) return rc;
    memcpy(ms_total, ctx->prof_ms, sizeof(ctx->prof_ms));
    memcpy(launches, ctx->prof_n, sizeof(ctx->prof_n));
    return DSM_OK;
}

// ---- batch stages ----
static int upload_tables(dsm_ctx *ctx, int n, const int32_t *ref, const float *poses, const int32_t *ofs, int n_local_single)
{
    int total = 0, maxper = 0;
    for (int b = 0; b < n; b++)
    {
        memcpy(ctx->h_pose + (size_t)b * 32, poses + (size_t)b * 16, 16 * sizeof(float));
        inverse4f(poses + (size_t)b * 16, ctx->h_pose + (size_t)b * 32 + 16);
        ctx->h_ref[b] = ref[b];
    }
    if (ofs)
    {
        if (ofs[0] != 0) return DSM_E_INVALID;
        for (int b = 0; b <= n; b++) ctx->h_ofs[b] = ofs[b];
        for (int b = 0; b < n; b++)
        {
            int c = ofs[b + 1] - ofs[b];
            if (c < 0) return DSM_E_INVALID;
            if (c > maxper) maxper = c;
        }
        total = ofs[n];
    }
    else
    {
        ctx->h_ofs[0] = 0;
        for (int b = 1; b <= n; b++) ctx->h_ofs[b] = n_local_single;
        total = maxper = n_local_single;
    }
    if (total > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    ctx->n_pool = total;
    ctx->d.max_pool_per_frame = maxper;
    // pose and inverse interleaved on the host: two strided copies
    CK(cudaMemcpy2DAsync(ctx->pose, 16 * sizeof(float), ctx->h_pose, 32 * sizeof(float), 16 * sizeof(float), n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(ctx->ipose, 16 * sizeof(float), ctx->h_pose + 16, 32 * sizeof(float), 16 * sizeof(float), n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->poolofs, ctx->h_ofs, ((size_t)n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->refidx, ctx->h_ref, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    return DSM_OK;
}

extern "C" int dsm_batch_upload(dsm_ctx *ctx, int n, const int32_t *ref, const uint8_t *gray, const float *depth,
                                const float *poses, const dsm_surfel_t *local, const int32_t *ofs)
{
    if (!ctx || !ref || !gray || !depth || !poses || !ofs) return DSM_E_INVALID;
    if (n < 1 || n > ctx->p.max_batch) return DSM_E_INVALID;
    if (ofs[n] > 0 && !local) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    // the pinned tables are reused: make sure the previous batch's copies have drained
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = upload_tables(ctx, n, ref, poses, ofs, 0);
    if (rc != DSM_OK) return rc;
    const int W = ctx->p.width, H = ctx->p.height;
    // [n][H][W] packed: one contiguous copy each, then the repack kernel lays out the pitched format
    CK(cudaMemcpyAsync(ctx->gray_packed, gray, (size_t)n * H * W, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->depth_packed, depth, (size_t)n * H * W * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    {
        DsmDev d = ctx->d;
        d.frame0 = 0;
        ProfScope p(ctx, DSM_K_REPACK);
        dsm_launch_repack(d, n, ctx->gray_packed, ctx->depth_packed, ctx->stream);
    }
    if (ctx->n_pool > 0)
    {
        CK(cudaMemcpyAsync(ctx->d.pool, local, (size_t)ctx->n_pool * sizeof(dsm_surfel_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->pool_snap, ctx->d.pool, (size_t)ctx->n_pool * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->nb = n;
    ctx->uploaded = true;
    ctx->ran = false;
    ctx->res_active = false;
    return DSM_OK;
}

extern "C" int dsm_batch_restore_pool(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->uploaded) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_pool > 0)
        CK(cudaMemcpyAsync(ctx->d.pool, ctx->pool_snap, (size_t)ctx->n_pool * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return DSM_OK;
}

// The per-frame schedule: generate_super_pixels (:960-975) then fuse (:58-71) then initialise (:79),
// enqueued for the frame slots [f0, f0 + nf).
// phase bit 0: the pose- and pool-independent part (superpixels, pixel normals, plane fit); bit 1: fuse + initialise.
static int launch_schedule(dsm_ctx *ctx, int f0, int nf, int max_pool_per_frame, cudaStream_t st, int phase = 3)
{
    DsmDev d = ctx->d;
    d.frame0 = f0;
    d.max_pool_per_frame = max_pool_per_frame;
    const int nb = nf;
    int budget = ctx->stop_after > 0 ? ctx->stop_after : 1 << 30;
#define STEP(ID, CALL)                  \
    if (budget-- > 0)                   \
    {                                   \
        ProfScope p(ctx, ID, st);       \
        CALL;                           \
    }
    if (phase & 1)
    { // generate_super_pixels (:960-975): seed init, 3 x (assign [+ stable relaxation], gather, Newton), plane fit
        STEP(DSM_K_SEED_INIT, dsm_launch_seed_init(d, nb, st));
        for (int it = 0; it < 3; it++) // ITERATION_NUM (fusion_functions.h:8)
        {
            STEP(it == 0 ? DSM_K_ASSIGN_FIRST : DSM_K_ASSIGN, dsm_launch_assign2(d, nb, it == 0, st));
            STEP(DSM_K_GATHER, dsm_launch_gather(d, ctx->maps, nb, st));
            STEP(DSM_K_NEWTON, dsm_launch_newton2(d, nb, st));
        }
        STEP(DSM_K_PLANE_GATHER, dsm_launch_plane_gather(d, ctx->maps, nb, st));
        STEP(DSM_K_PLANE_SOLVE, dsm_launch_gn_solve(d, nb, st));
    }
    if (phase & 2)
    {
    if (max_pool_per_frame > 0)
    {
        STEP(DSM_K_FUSE, dsm_launch_fuse(d, nb, st));
    }
    STEP(DSM_K_INIT_SURFELS, dsm_launch_init_surfels(d, nb, st));
    }
#undef STEP
    CK(cudaGetLastError());
    return DSM_OK;
}

// Enqueue the schedule, optionally followed by the resident-pool post-step (compaction into `alt`).
// The 17-21 launches are captured once per distinct parameter set into a CUDA graph and replayed: the
// per-launch CPU cost and inter-kernel gaps matter for single frames and small chunks.  Profiling
// (event pairs around kernels) and the debug kernel budget use plain launches.
static int enqueue_schedule(dsm_ctx *ctx, int f0, int nf, int maxper, cudaStream_t st, bool compact = false, int phase = 3)
{
    // grids over the pool are sized from `maxper` rounded up to 4 Ki surfels (the kernels take the true
    // ranges from poolofs on the device, surplus blocks exit at once), so that callers whose pool size
    // changes a little from call to call keep hitting the same captured graph
    if (maxper > 0) maxper = (maxper + 4095) & ~4095;
    auto plain = [&]() -> int
    {
        int rc = launch_schedule(ctx, f0, nf, maxper, st, phase);
        if (rc != DSM_OK) return rc;
        if (compact)
        {
            DsmDev d = ctx->d;
            d.frame0 = f0;
            dsm_launch_pool_compact(d, f0, maxper, ctx->blkcnt, ctx->blkofs, ctx->newofs, ctx->pool_snap, ctx->res_ofs, st);
        }
        return DSM_OK;
    };
    if (!ctx->use_graphs || ctx->prof_mask != 0 || ctx->stop_after > 0) return plain();
    for (auto &g : ctx->graphs)
        if (g.f0 == f0 && g.nf == nf && g.maxper == maxper && g.compact == (int)compact && g.pool == ctx->d.pool &&
            g.poolofs == ctx->d.poolofs && g.alt == ctx->pool_snap && g.phase == phase)
        {
            CK(cudaGraphLaunch(g.exec, st));
            return DSM_OK;
        }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    {
        cudaGetLastError();
        return plain();
    }
    int rc = plain();
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != DSM_OK || ce != cudaSuccess || !graph)
    {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        ctx->use_graphs = false; // capture is not possible in this environment: fall back to plain launches
        return rc != DSM_OK ? rc : plain();
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess || !exec)
    {
        cudaGetLastError();
        ctx->use_graphs = false;
        return plain();
    }
    if (ctx->graphs.size() >= 64)
    { // bounded cache
        cudaGraphExecDestroy(ctx->graphs.front().exec);
        ctx->graphs.erase(ctx->graphs.begin());
    }
    ctx->graphs.push_back(GraphEntry{f0, nf, maxper, (int)compact, ctx->d.pool, ctx->d.poolofs, ctx->pool_snap, exec, phase});
    CK(cudaGraphLaunch(exec, st));
    return DSM_OK;
}

extern "C" int dsm_batch_run(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->uploaded) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    // The frames of a batch are independent, and most kernels of the schedule are issue- or latency-bound rather than
    // DRAM-bound: two sub-batches run CONCURRENTLY on two streams fill each other's tails and stalls (measured: 32 frames
    // as 2 x 16 concurrent take less time than 1 x 32).  Running sub-batches one after the other to stay inside L2 was
    // also measured and is slower (8 / 16 frames per pass: 1.70 / 1.29 ms per 32 frames against 1.08 ms); profiles/README.md.
    const int parts = (ctx->stop_after > 0 || ctx->nb < 8) ? 1 : (ctx->run_split < ctx->nb ? ctx->run_split : ctx->nb);
    if (parts <= 1)
    {
        int rc = enqueue_schedule(ctx, 0, ctx->nb, ctx->d.max_pool_per_frame, ctx->stream);
        if (rc != DSM_OK) return rc;
    }
    else
    {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        int f0 = 0;
        for (int i = 0; i < parts; i++)
        {
            const int nf = (ctx->nb - f0 + (parts - i) - 1) / (parts - i);
            CK(cudaStreamWaitEvent(ctx->s_comp[i], ctx->ev_fork, 0));
            int rc = enqueue_schedule(ctx, f0, nf, ctx->d.max_pool_per_frame, ctx->s_comp[i]);
            if (rc != DSM_OK) return rc;
            CK(cudaEventRecord(ctx->ev_join[i], ctx->s_comp[i]));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
            f0 += nf;
        }
    }
    ctx->ran = true;
    return DSM_OK;
}

extern "C" int dsm_set_concurrency(dsm_ctx *ctx, int sub_batches)
{
    if (!ctx || sub_batches < 1 || sub_batches > 4) return DSM_E_INVALID;
    ctx->run_split = sub_batches;
    return DSM_OK;
}

extern "C" int dsm_set_constants(dsm_ctx *ctx, const dsm_constants *k)
{
    if (!ctx || !k) return DSM_E_INVALID;
    if (!(k->huber_range > 0.0) || !(k->baseline > 0.0) || !(k->disparity_error > 0.0) || !(k->min_tolerate_diff >= 0.0)) return DSM_E_INVALID;
    if (ctx->in_flight) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &g : ctx->graphs) cudaGraphExecDestroy(g.exec); // captured schedules embed the kernel parameters
    ctx->graphs.clear();
    apply_constants(ctx->d, *k);
    return DSM_OK;
}

extern "C" int dsm_debug_stop_after(dsm_ctx *ctx, int n)
{
    if (!ctx) return DSM_E_INVALID;
    ctx->stop_after = n;
    return DSM_OK;
}

extern "C" int dsm_debug_invariant_violations(dsm_ctx *ctx, int *count)
{
    if (!ctx || !count) return DSM_E_INVALID;
    if (!ctx->ran) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    std::vector<int32_t> h((size_t)ctx->nb);
    CK(cudaMemcpyAsync(h.data(), ctx->d.errflag, (size_t)ctx->nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int c = 0;
    for (int v : h) c += v;
    *count = c;
    return DSM_OK;
}

extern "C" int dsm_batch_download(dsm_ctx *ctx, dsm_surfel_t *local_out, dsm_surfel_t *new_out, int32_t *n_new)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->ran) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    if (local_out && ctx->n_pool > 0)
        CK(cudaMemcpyAsync(local_out, ctx->d.pool, (size_t)ctx->n_pool * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (new_out)
        CK(cudaMemcpyAsync(new_out, ctx->d.newsurf, (size_t)ctx->nb * ctx->S * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_new)
        CK(cudaMemcpyAsync(n_new, ctx->d.nnew, (size_t)ctx->nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return DSM_OK;
}

extern "C" int dsm_sync(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return DSM_OK;
}

// End-to-end batch call with host buffers (dsm_fuse_batch = dsm_fuse_batch_async + dsm_batch_wait).  The batch is cut into chunks of frames; chunk c's H2D
// (copy stream), kernels (compute stream) and D2H (second copy stream) overlap with the neighbouring
// chunks, so the call costs about max(PCIe time, kernel time) instead of their sum.  Host buffers
// should be pinned for the copies to be truly asynchronous (pageable memory works, staged by the driver).
extern "C" int dsm_batch_wait(dsm_ctx *ctx)
{
    if (!ctx) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_h2d));
    for (int i = 0; i < 4; i++) CK(cudaStreamSynchronize(ctx->s_comp[i]));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->s_d2h));
    CK(cudaGetLastError());
    ctx->in_flight = false;
    return DSM_OK;
}

extern "C" int dsm_fuse_batch_async(dsm_ctx *ctx, int n, const int32_t *ref, const uint8_t *gray, const float *depth,
                                    const float *poses, dsm_surfel_t *local, const int32_t *ofs,
                                    dsm_surfel_t *new_out, int32_t *n_new)
{
    if (!ctx || !ref || !gray || !depth || !poses || !ofs) return DSM_E_INVALID;
    if (n < 1 || n > ctx->p.max_batch) return DSM_E_INVALID;
    if (ofs[n] > 0 && !local) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->in_flight)
    { // one batch per context at a time: finish the previous one first
        int rcw = dsm_batch_wait(ctx);
        if (rcw != DSM_OK) return rcw;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = upload_tables(ctx, n, ref, poses, ofs, 0); // small tables, on the compute stream
    if (rc != DSM_OK) return rc;
    const int W = ctx->p.width, H = ctx->p.height, S = ctx->S;
    const size_t fpx = (size_t)H * W;
    int nchunks = n >= 16 ? 4 : (n >= 4 ? 2 : 1);
    if (const char *e = getenv("DSM_E2E_CHUNKS"))
    { // tuning knob for experiments (1..8)
        const int v = atoi(e);
        if (v >= 1 && v <= 8 && v <= n) nchunks = v;
    }
    int csize[8]; // even chunks (a smaller first chunk was tried: no gain, the kernels of small chunks are less efficient)
    {
        int rest = n, left = nchunks;
        for (int c = 0; c < nchunks; c++)
        {
            csize[c] = (rest + left - 1) / left;
            rest -= csize[c];
            left--;
        }
    }
    // copies must not start before earlier work on the compute stream (previous users of the buffers) is done
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_start, 0));
    for (int i = 0; i < 4; i++) CK(cudaStreamWaitEvent(ctx->s_comp[i], ctx->ev_start, 0));
    int nc = 0;
    for (int f0 = 0; f0 < n; f0 += csize[nc], nc++)
    {
        const int nf = csize[nc];
        if (nf <= 0) break;
        const int p0 = ofs[f0], p1 = ofs[f0 + nf];
        int maxper = 0;
        for (int b = f0; b < f0 + nf; b++) maxper = (ofs[b + 1] - ofs[b] > maxper) ? ofs[b + 1] - ofs[b] : maxper;
        // H2D of this chunk
        CK(cudaMemcpyAsync(ctx->gray_packed + (size_t)f0 * fpx, gray + (size_t)f0 * fpx, (size_t)nf * fpx, cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaMemcpyAsync(ctx->depth_packed + (size_t)f0 * fpx, depth + (size_t)f0 * fpx, (size_t)nf * fpx * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
        if (p1 > p0)
            CK(cudaMemcpyAsync(ctx->d.pool + p0, local + p0, (size_t)(p1 - p0) * sizeof(dsm_surfel_t), cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaEventRecord(ctx->ev_h2d[nc], ctx->s_h2d));
        // kernels: each chunk on its own compute stream, so the latency-bound per-seed kernels of one
        // chunk overlap the wide kernels of its neighbours instead of leaving SMs idle
        cudaStream_t cs = nchunks > 1 ? ctx->s_comp[nc & 3] : ctx->stream;
        CK(cudaStreamWaitEvent(cs, ctx->ev_h2d[nc], 0));
        {
            DsmDev d = ctx->d;
            d.frame0 = f0;
            ProfScope p(ctx, DSM_K_REPACK, cs);
            dsm_launch_repack(d, nf, ctx->gray_packed + (size_t)f0 * fpx, ctx->depth_packed + (size_t)f0 * fpx, cs);
        }
        rc = enqueue_schedule(ctx, f0, nf, maxper, cs);
        if (rc != DSM_OK) return rc;
        CK(cudaEventRecord(ctx->ev_done[nc], cs));
        // D2H of this chunk
        CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[nc], 0));
        if (p1 > p0)
            CK(cudaMemcpyAsync(local + p0, ctx->d.pool + p0, (size_t)(p1 - p0) * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (new_out)
            CK(cudaMemcpyAsync(new_out + (size_t)f0 * S, ctx->d.newsurf + (size_t)f0 * S, (size_t)nf * S * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (n_new)
            CK(cudaMemcpyAsync(n_new + f0, ctx->d.nnew + f0, (size_t)nf * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
    ctx->nb = n;
    ctx->uploaded = true;
    ctx->ran = true;
    ctx->in_flight = true;
    ctx->res_active = false; // the pool buffer now holds this batch's slices: a resident pool (if any) is gone
    CK(cudaGetLastError());
    return DSM_OK;
}

extern "C" int dsm_fuse_batch(dsm_ctx *ctx, int n, const int32_t *ref, const uint8_t *gray, const float *depth,
                              const float *poses, dsm_surfel_t *local, const int32_t *ofs,
                              dsm_surfel_t *new_out, int32_t *n_new)
{
    int rc = dsm_fuse_batch_async(ctx, n, ref, gray, depth, poses, local, ofs, new_out, n_new);
    if (rc != DSM_OK) return rc;
    return dsm_batch_wait(ctx);
}

extern "C" int dsm_fuse_frame(dsm_ctx *ctx, int ref_idx, const uint8_t *gray, size_t gray_pitch,
                              const float *depth, size_t depth_pitch, const float pose[16],
                              dsm_surfel_t *local, int n_local, dsm_surfel_t *new_out, int new_cap, int *n_new)
{
    if (!ctx || !gray || !depth || !pose || n_local < 0 || new_cap < 0 || !n_new) return DSM_E_INVALID;
    if (n_local > 0 && !local) return DSM_E_INVALID;
    if (new_cap > 0 && !new_out) return DSM_E_INVALID;
    const int W = ctx->p.width, H = ctx->p.height;
    if (gray_pitch < (size_t)W || depth_pitch < (size_t)W * 4) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->in_flight)
    { // an asynchronous batch still owns the buffers: finish it first
        int rcw = dsm_batch_wait(ctx);
        if (rcw != DSM_OK) return rcw;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->res_active = false; // the caller's surfels replace whatever the pool buffer held (a resident pool included)
    int32_t ref = ref_idx;
    int rc = upload_tables(ctx, 1, &ref, pose, nullptr, n_local);
    if (rc != DSM_OK) return rc;
    CK(cudaMemcpy2DAsync(ctx->gray, ctx->Wp, gray, gray_pitch, W, H, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(ctx->depth, (size_t)ctx->Wp * 4, depth, depth_pitch, (size_t)W * 4, H, cudaMemcpyHostToDevice, ctx->stream));
    if (n_local > 0)
        CK(cudaMemcpyAsync(ctx->d.pool, local, (size_t)n_local * sizeof(dsm_surfel_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->nb = 1;
    ctx->uploaded = true;
    rc = dsm_batch_run(ctx);
    if (rc != DSM_OK) return rc;
    if (n_local > 0)
        CK(cudaMemcpyAsync(local, ctx->d.pool, (size_t)n_local * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
    int32_t cnt = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->d.nnew, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_new = cnt;
    const int ncopy = cnt < new_cap ? cnt : new_cap;
    if (ncopy > 0)
    {
        CK(cudaMemcpyAsync(new_out, ctx->d.newsurf, (size_t)ncopy * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaGetLastError());
    return DSM_OK;
}

// ---- GPU-resident pool (stream mode) ----
// The resident pool is described by res_ofs = {0, n} on the device.  Frames alternate between two
// frame slots (when max_batch >= 2) so that the H2D copy of frame t+1 overlaps the kernels of frame t.
static DsmDev resident_view(dsm_ctx *ctx, int slot)
{
    DsmDev d = ctx->d;
    d.frame0 = slot;
    d.poolofs = ctx->res_ofs - slot; // kernels read poolofs[b], poolofs[b+1] with b == slot
    return d;
}

extern "C" int dsm_pool_upload(dsm_ctx *ctx, const dsm_surfel_t *local, int n)
{
    if (!ctx || n < 0 || (n > 0 && !local)) return DSM_E_INVALID;
    if (n > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_h2d));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n > 0) CK(cudaMemcpyAsync(ctx->d.pool, local, (size_t)n * sizeof(dsm_surfel_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h_ofs[0] = 0;
    ctx->h_ofs[1] = n;
    CK(cudaMemcpyAsync(ctx->res_ofs, ctx->h_ofs, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // `local` may be pageable / reused by the caller
    ctx->res_upper = n;
    ctx->res_active = true;
    ctx->n_pool = n;
    return DSM_OK;
}

extern "C" int dsm_pool_size(dsm_ctx *ctx, int *n)
{
    if (!ctx || !n) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int32_t h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->res_ofs, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n = h[1];
    ctx->res_upper = h[1];
    return DSM_OK;
}

extern "C" int dsm_pool_download(dsm_ctx *ctx, dsm_surfel_t *out, int cap, int *n)
{
    if (!ctx || !n || cap < 0 || (cap > 0 && !out)) return DSM_E_INVALID;
    int rc = dsm_pool_size(ctx, n);
    if (rc != DSM_OK) return rc;
    const int c = *n < cap ? *n : cap;
    if (c > 0)
    {
        CK(cudaMemcpyAsync(out, ctx->d.pool, (size_t)c * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DSM_OK;
}

extern "C" int dsm_pool_retire(dsm_ctx *ctx, int kf, dsm_surfel_t *out, int cap, int *n_out)
{
    if (!ctx || !n_out || cap < 0 || (cap > 0 && !out)) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    // count first: nothing is flagged dead unless every retired surfel fits the caller's buffer
    int32_t h[2] = {0, 0};
    dsm_launch_pool_retire(resident_view(ctx, 0), 0, ctx->res_upper, kf, ctx->blkcnt, ctx->blkofs, ctx->newofs, nullptr, ctx->stream);
    CK(cudaMemcpyAsync(h, ctx->newofs, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    *n_out = h[0];
    if (h[0] > cap) return DSM_E_CAPACITY; // pool untouched; call again with cap >= *n_out
    // the alternate pool buffer is free between frames: use it as the ordered output
    dsm_launch_pool_retire(resident_view(ctx, 0), 0, ctx->res_upper, kf, ctx->blkcnt, ctx->blkofs, ctx->newofs, ctx->pool_snap, ctx->stream);
    CK(cudaGetLastError());
    const int c = h[0] < cap ? h[0] : cap;
    if (c > 0)
    {
        CK(cudaMemcpyAsync(out, ctx->pool_snap, (size_t)c * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DSM_OK;
}

// publish_*_pointcloud / save_cloud / save_mesh filters on the resident pool (see include/dsm.h)
static int pool_export(dsm_ctx *ctx, int min_ut, bool as_cloud, void *out, int cap, int *n_out)
{
    if (!ctx || !n_out || cap < 0 || (cap > 0 && !out)) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    // the alternate pool buffer is free between frames: ordered output (16 or 44 bytes per selected surfel)
    dsm_launch_pool_export(resident_view(ctx, 0), 0, ctx->res_upper, 2, min_ut, as_cloud, ctx->blkcnt, ctx->blkofs, ctx->newofs,
                           ctx->pool_snap, ctx->stream);
    int32_t h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->newofs, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    *n_out = h[0];
    const int c = h[0] < cap ? h[0] : cap;
    if (c > 0)
    {
        const size_t rec = as_cloud ? sizeof(dsm_point_t) : sizeof(dsm_surfel_t);
        CK(cudaMemcpyAsync(out, ctx->pool_snap, (size_t)c * rec, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DSM_OK;
}
extern "C" int dsm_pool_export_cloud(dsm_ctx *ctx, int min_update_times, dsm_point_t *out, int cap, int *n_out)
{
    return pool_export(ctx, min_update_times, true, out, cap, n_out);
}
extern "C" int dsm_pool_export_surfels(dsm_ctx *ctx, int min_update_times, dsm_surfel_t *out, int cap, int *n_out)
{
    return pool_export(ctx, min_update_times, false, out, cap, n_out);
}

extern "C" int dsm_pool_append(dsm_ctx *ctx, const dsm_surfel_t *surfels, int n)
{
    if (!ctx || n < 0 || (n > 0 && !surfels)) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    int cur = 0;
    int rc = dsm_pool_size(ctx, &cur);
    if (rc != DSM_OK) return rc;
    if (cur + n > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    if (n == 0) return DSM_OK;
    CK(cudaMemcpyAsync(ctx->d.pool + cur, surfels, (size_t)n * sizeof(dsm_surfel_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h_ofs[0] = 0;
    ctx->h_ofs[1] = cur + n;
    CK(cudaMemcpyAsync(ctx->res_ofs, ctx->h_ofs, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->res_upper = cur + n;
    return DSM_OK;
}

extern "C" int dsm_pool_transform(dsm_ctx *ctx, const float Wm[16])
{
    if (!ctx || !Wm) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->wmat, Wm, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dsm_launch_pool_transform(resident_view(ctx, 0), 0, ctx->res_upper, ctx->wmat, ctx->stream);
    CK(cudaGetLastError());
    return DSM_OK;
}

extern "C" int dsm_fuse_frame_resident(dsm_ctx *ctx, int ref_idx, const uint8_t *gray, size_t gray_pitch,
                                       const float *depth, size_t depth_pitch, const float pose[16], int *n_new)
{
    if (!ctx || !gray || !depth || !pose) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    const int W = ctx->p.width, H = ctx->p.height;
    if (gray_pitch < (size_t)W || depth_pitch < (size_t)W * 4) return DSM_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->res_upper + ctx->S > ctx->p.max_local_surfels)
    { // the bound may be loose: fetch the exact size before giving up
        int n = 0;
        int rc = dsm_pool_size(ctx, &n);
        if (rc != DSM_OK) return rc;
        if (n + ctx->S > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    }
    const int nslots = ctx->p.max_batch >= 2 ? 2 : 1;
    const int slot = ctx->res_frame % nslots;
    ctx->res_frame++;
    const size_t fpx = (size_t)H * W;
    // this slot's pinned tables / image buffers were last used two frames ago: make sure that copy and
    // those kernels are done (normally long finished, so these do not stall)
    CK(cudaEventSynchronize(ctx->ev_h2d[slot]));
    CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[slot], 0));
    memcpy(ctx->h_pose + (size_t)slot * 32, pose, 16 * sizeof(float));
    inverse4f(pose, ctx->h_pose + (size_t)slot * 32 + 16);
    ctx->h_ref[slot] = ref_idx;
    CK(cudaMemcpyAsync(ctx->pose + slot * 16, ctx->h_pose + (size_t)slot * 32, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaMemcpyAsync(ctx->ipose + slot * 16, ctx->h_pose + (size_t)slot * 32 + 16, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaMemcpyAsync(ctx->refidx + slot, ctx->h_ref + slot, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_h2d));
    const bool packed = gray_pitch == (size_t)W && depth_pitch == (size_t)W * 4;
    if (packed)
    { // one contiguous copy per image, repacked to the pitched layout on the device
        CK(cudaMemcpyAsync(ctx->gray_packed + (size_t)slot * fpx, gray, fpx, cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaMemcpyAsync(ctx->depth_packed + (size_t)slot * fpx, depth, fpx * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
    }
    else
    {
        CK(cudaMemcpy2DAsync(ctx->gray + (size_t)slot * ctx->px, ctx->Wp, gray, gray_pitch, W, H, cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaMemcpy2DAsync(ctx->depth + (size_t)slot * ctx->px, (size_t)ctx->Wp * 4, depth, depth_pitch, (size_t)W * 4, H, cudaMemcpyHostToDevice, ctx->s_h2d));
    }
    CK(cudaEventRecord(ctx->ev_h2d[slot], ctx->s_h2d));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[slot], 0));
    const DsmDev dv = resident_view(ctx, slot);
    if (packed)
    {
        ProfScope p(ctx, DSM_K_REPACK);
        dsm_launch_repack(dv, 1, ctx->gray_packed + (size_t)slot * fpx, ctx->depth_packed + (size_t)slot * fpx, ctx->stream);
    }
    ctx->nb = slot + 1;
    ctx->uploaded = true;
    {
        // enqueue_schedule takes its view from ctx->d: temporarily present the resident pool table
        const int32_t *saved = ctx->d.poolofs;
        ctx->d.poolofs = dv.poolofs;
        // + post-step of SurfelMap::fuse_map on the device: compact into the alternate buffer (then swap)
        // grid sizes use the bound rounded up to 64 Ki surfels so that the captured graph stays valid while
        // the pool grows (surplus blocks exit at once: the kernels read the true range from the device)
        int upq = (ctx->res_upper + 65535) / 65536 * 65536;
        if (upq > ctx->p.max_local_surfels) upq = ctx->p.max_local_surfels;
        int rc = enqueue_schedule(ctx, slot, 1, upq, ctx->stream, true);
        ctx->d.poolofs = saved;
        if (rc != DSM_OK) return rc;
    }
    ctx->ran = true;
    CK(cudaEventRecord(ctx->ev_done[slot], ctx->stream)); // (the compaction's last kernel stored the new pool size in res_ofs[1])
    ctx->single_pending = true;
    {
        dsm_surfel_t *t = ctx->d.pool;
        ctx->d.pool = ctx->pool_snap;
        ctx->pool_snap = t;
    }
    ctx->res_upper += ctx->S; // every seed can add at most one surfel; dsm_pool_size() tightens it
    CK(cudaGetLastError());
    // the caller may reuse its image buffers as soon as we return: wait for THIS frame's copy only
    // (the kernels keep running and overlap the next call's copy)
    CK(cudaEventSynchronize(ctx->ev_h2d[slot]));
    if (n_new)
    {
        int32_t c = 0;
        CK(cudaMemcpyAsync(&c, ctx->d.nnew + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *n_new = c;
    }
    return DSM_OK;
}

// ---- inactive store (EXPERIMENTAL) ----
// What SurfelMap keeps per pose in attached_surfels / inactive_pointcloud (surfel_map.cpp:1479-1497, :1583-1587,
// :681-748), resident on the device.  Only host orchestration over the pool kernels: retirement is the ordered
// compaction of dsm_pool_retire writing straight into the store, the per-pose warp is k_pool_transform over one
// segment, the inactive cloud is k_pool_scatter_cloud over the store.
static DsmDev inactive_view(dsm_ctx *ctx)
{
    DsmDev d = ctx->d;
    d.frame0 = 0;
    d.pool = ctx->inact;
    d.poolofs = ctx->inact_ofs; // {first, last} of the range a call works on, set in stream order by k_set2
    return d;
}

extern "C" int dsm_inactive_reserve(dsm_ctx *ctx, int max_inactive_surfels)
{
    if (!ctx || max_inactive_surfels < 1) return DSM_E_INVALID;
    if (ctx->inact_size > 0) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->inact);
    ctx->inact = nullptr;
    ctx->inact_cap = 0;
    if (!ctx->inact_ofs && dmalloc(&ctx->inact_ofs, 2) != cudaSuccess) return DSM_E_NOMEM;
    if (dmalloc(&ctx->inact, (size_t)max_inactive_surfels) != cudaSuccess)
    {
        cudaGetLastError();
        return DSM_E_NOMEM;
    }
    ctx->inact_cap = max_inactive_surfels;
    ctx->inact_segs.clear();
    return DSM_OK;
}

extern "C" int dsm_inactive_size(dsm_ctx *ctx, int *n_surfels, int *n_segments)
{
    if (!ctx) return DSM_E_INVALID;
    if (n_surfels) *n_surfels = ctx->inact_size;
    if (n_segments) *n_segments = (int)ctx->inact_segs.size();
    return DSM_OK;
}

// move_add_surfels, removal loop: local surfels last updated by keyframe kf -> a new segment of the store
extern "C" int dsm_inactive_retire(dsm_ctx *ctx, int kf, int *n_moved)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->res_active || !ctx->inact) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int cur = 0;
    int rc = dsm_pool_size(ctx, &cur); // exact pool size: bounds the segment
    if (rc != DSM_OK) return rc;
    if (ctx->inact_size + cur > ctx->inact_cap) return DSM_E_CAPACITY;
    dsm_launch_pool_retire(resident_view(ctx, 0), 0, ctx->res_upper, kf, ctx->blkcnt, ctx->blkofs, ctx->newofs, ctx->inact + ctx->inact_size, ctx->stream);
    int32_t h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->newofs, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (h[0] > 0)
    {
        ctx->inact_segs.push_back(dsm_ctx::InactSeg{kf, ctx->inact_size, h[0]});
        ctx->inact_size += h[0];
    }
    if (n_moved) *n_moved = h[0];
    return DSM_OK;
}

// move_add_surfels, insertion: the pose re-enters the drift-free window, its surfels go back to the end of the pool
extern "C" int dsm_inactive_reactivate(dsm_ctx *ctx, int kf, int *n_moved)
{
    if (!ctx) return DSM_E_INVALID;
    if (!ctx->res_active || !ctx->inact) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int total = 0;
    for (const auto &sg : ctx->inact_segs)
        if (sg.kf == kf) total += sg.cnt;
    if (n_moved) *n_moved = total;
    if (total == 0) return DSM_OK;
    int cur = 0;
    int rc = dsm_pool_size(ctx, &cur);
    if (rc != DSM_OK) return rc;
    if (cur + total > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    cudaStream_t st = ctx->stream;
    const int npool = ctx->p.max_local_surfels > 0 ? ctx->p.max_local_surfels : 1;
    for (size_t i = 0; i < ctx->inact_segs.size();)
    {
        const dsm_ctx::InactSeg sg = ctx->inact_segs[i];
        if (sg.kf != kf)
        {
            i++;
            continue;
        }
        CK(cudaMemcpyAsync(ctx->d.pool + cur, ctx->inact + sg.ofs, (size_t)sg.cnt * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, st));
        cur += sg.cnt;
        // close the gap: move the tail down through the alternate pool buffer (free between frames), chunk by chunk
        int src = sg.ofs + sg.cnt;
        const int end = ctx->inact_size;
        while (src < end)
        {
            const int len = (end - src) < npool ? (end - src) : npool;
            CK(cudaMemcpyAsync(ctx->pool_snap, ctx->inact + src, (size_t)len * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(ctx->inact + src - sg.cnt, ctx->pool_snap, (size_t)len * sizeof(dsm_surfel_t), cudaMemcpyDeviceToDevice, st));
            src += len;
        }
        ctx->inact_size -= sg.cnt;
        ctx->inact_segs.erase(ctx->inact_segs.begin() + (long)i);
        for (size_t j = i; j < ctx->inact_segs.size(); j++) ctx->inact_segs[j].ofs -= sg.cnt;
    }
    dsm_launch_set2(ctx->res_ofs, 0, cur, st);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->res_upper = cur;
    return DSM_OK;
}

// warp_inactive_surfels_cpu_kernel for ONE pose (surfel_map.cpp:704-733): p <- W p, n <- R_W n over its segment(s),
// W = (T_loop * T_cam^-1) computed by the caller in fp64 and cast to float, column-major.  Asynchronous.
extern "C" int dsm_inactive_transform(dsm_ctx *ctx, int kf, const float Wm[16])
{
    if (!ctx || !Wm) return DSM_E_INVALID;
    if (!ctx->inact) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    bool first = true;
    for (const auto &sg : ctx->inact_segs)
    {
        if (sg.kf != kf) continue;
        if (first) CK(cudaMemcpyAsync(ctx->wmat, Wm, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        first = false;
        dsm_launch_set2(ctx->inact_ofs, sg.ofs, sg.ofs + sg.cnt, ctx->stream);
        dsm_launch_pool_transform(inactive_view(ctx), 0, sg.cnt, ctx->wmat, ctx->stream);
    }
    CK(cudaGetLastError());
    return DSM_OK;
}

// attached_surfels of one pose (kf >= 0) or of every pose in store order (kf < 0), e.g. for save_mesh
extern "C" int dsm_inactive_download(dsm_ctx *ctx, int kf, dsm_surfel_t *out, int cap, int *n_out)
{
    if (!ctx || !n_out || cap < 0 || (cap > 0 && !out)) return DSM_E_INVALID;
    if (!ctx->inact) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int n = 0;
    for (const auto &sg : ctx->inact_segs)
    {
        if (kf >= 0 && sg.kf != kf) continue;
        const int c = (n + sg.cnt <= cap) ? sg.cnt : (cap > n ? cap - n : 0);
        if (c > 0) CK(cudaMemcpyAsync(out + n, ctx->inact + sg.ofs, (size_t)c * sizeof(dsm_surfel_t), cudaMemcpyDeviceToHost, ctx->stream));
        n += sg.cnt;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return DSM_OK;
}

// inactive_pointcloud (publish_inactive_pointcloud surfel_map.cpp:1385-1396, the tail of publish_all_pointcloud and
// save_cloud): one PointXYZI per stored surfel, store order
extern "C" int dsm_inactive_export_cloud(dsm_ctx *ctx, dsm_point_t *out, int cap, int *n_out)
{
    if (!ctx || !n_out || cap < 0 || (cap > 0 && !out)) return DSM_E_INVALID;
    if (!ctx->inact) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    *n_out = ctx->inact_size;
    const int want = ctx->inact_size < cap ? ctx->inact_size : cap;
    const int npool = ctx->p.max_local_surfels > 0 ? ctx->p.max_local_surfels : 1;
    const DsmDev d = inactive_view(ctx);
    for (int c0 = 0; c0 < want; c0 += npool)
    { // chunks sized for the block tables and the alternate pool buffer (16-byte points into 44-byte slots)
        const int len = (want - c0) < npool ? (want - c0) : npool;
        dsm_launch_set2(ctx->inact_ofs, c0, c0 + len, ctx->stream);
        dsm_launch_pool_export(d, 0, len, 2, INT32_MIN, true, ctx->blkcnt, ctx->blkofs, ctx->newofs, ctx->pool_snap, ctx->stream);
        CK(cudaMemcpyAsync(out + c0, ctx->pool_snap, (size_t)len * sizeof(dsm_point_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return DSM_OK;
}

// A run of n consecutive frames of ONE sequential stream on the resident pool.  Results are identical to n calls of
// dsm_fuse_frame_resident: superpixels, normals and plane fits do not depend on the pose or the pool, so they run
// for all n frames as one batch (full-width launches instead of n latency-bound single-frame ones); only the fuse /
// initialise / compaction steps, which carry the pool from frame t to frame t+1, run frame by frame.
extern "C" int dsm_fuse_stream_resident(dsm_ctx *ctx, int n, const int32_t *ref_idx, const uint8_t *gray, const float *depth,
                                        const float *poses, int32_t *n_new)
{
    if (!ctx || !ref_idx || !gray || !depth || !poses) return DSM_E_INVALID;
    if (n < 1 || n > ctx->p.max_batch) return DSM_E_INVALID;
    if (!ctx->res_active) return DSM_E_STATE;
    if (ctx->in_flight) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    if (ctx->res_upper + n * ctx->S > ctx->p.max_local_surfels)
    { // the bound may be loose: fetch the exact size before giving up
        int cur = 0;
        int rc = dsm_pool_size(ctx, &cur);
        if (rc != DSM_OK) return rc;
        if (cur + n * ctx->S > ctx->p.max_local_surfels) return DSM_E_CAPACITY;
    }
    const int W = ctx->p.width, H = ctx->p.height;
    const size_t fpx = (size_t)H * W;
    // two staging halves when they fit: the H2D copy of this run overlaps the kernels of the previous one
    const bool dbl = 2 * n <= ctx->p.max_batch;
    const int par = dbl ? (ctx->res_frame & 1) : 0;
    ctx->res_frame++;
    const int e = 6 + par; // events 6/7: this half's staging buffers and pinned tables are free again
    CK(cudaEventSynchronize(ctx->ev_done[e]));
    if (!dbl) CK(cudaEventSynchronize(ctx->ev_done[6 + (1 - par)]));
    // single-frame calls (dsm_fuse_frame_resident) may still be running out of the same staging buffers and frame slots
    // 0 / 1 (only then: the events are also recorded at the end of every run of this function, and waiting for the
    // previous run here would serialise this run's copies behind it)
    const bool after_single = ctx->single_pending;
    ctx->single_pending = false;
    if (after_single)
    {
        CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[0], 0));
        CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[1], 0));
    }
    const size_t so = (size_t)par * (size_t)(ctx->p.max_batch / 2); // fixed halves: runs of different length never overlap
    // pinned tables of this mode live in the second half of h_pose / h_ref (the first half belongs to the single-frame
    // and batch calls, which only wait for their own copies before rewriting it); this half's previous table copy ran
    // before ev_done[e], which has been waited for above
    const size_t ho = (size_t)ctx->p.max_batch + so;
    for (int t = 0; t < n; t++)
    {
        memcpy(ctx->h_pose + (ho + t) * 32, poses + (size_t)t * 16, 16 * sizeof(float));
        inverse4f(poses + (size_t)t * 16, ctx->h_pose + (ho + t) * 32 + 16);
        ctx->h_ref[ho + t] = ref_idx[t];
    }
    uint8_t *gp = ctx->gray_packed + so * fpx;
    float *dp = ctx->depth_packed + so * fpx;
    CK(cudaMemcpyAsync(gp, gray, (size_t)n * fpx, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaMemcpyAsync(dp, depth, (size_t)n * fpx * sizeof(float), cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(cudaEventRecord(ctx->ev_h2d[e], ctx->s_h2d));
    cudaStream_t st = ctx->stream;
    // Phase 1 (superpixels, normals, plane fits of all n frames; never touches the pool) runs on a side stream and in this
    // half's own frame slots [so, so + n): it overlaps the frame-by-frame phase 2 of the PREVIOUS run, which owns the
    // other half's slots and the main stream.  (With a single half, so == 0 and everything is ordered on the main stream.)
    // Phase 2 is a chain of small launches per frame: next to phase 1's full-width grids its CTAs would queue behind a
    // wave of theirs at every launch, so it runs on the context's highest-priority stream (forked from / joined into the
    // main stream, which keeps the order with every other call on the context).
    cudaStream_t s1 = dbl ? ctx->s_comp[par] : st;
    cudaStream_t s2 = dbl ? ctx->s_hi : st;
    const int slot0 = (int)so;
    if (dbl)
    { // this half's slots and device tables were last read by the phase 2 of the run before the previous one
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
        CK(cudaStreamWaitEvent(s1, ctx->ev_done[e], 0));
    }
    CK(cudaMemcpy2DAsync(ctx->pose + (size_t)slot0 * 16, 16 * sizeof(float), ctx->h_pose + ho * 32, 32 * sizeof(float), 16 * sizeof(float), n, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpy2DAsync(ctx->ipose + (size_t)slot0 * 16, 16 * sizeof(float), ctx->h_pose + ho * 32 + 16, 32 * sizeof(float), 16 * sizeof(float), n, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(ctx->refidx + slot0, ctx->h_ref + ho, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s1));
    CK(cudaStreamWaitEvent(s1, ctx->ev_h2d[e], 0));
    {
        DsmDev d = ctx->d;
        d.frame0 = slot0;
        ProfScope p(ctx, DSM_K_REPACK, s1);
        dsm_launch_repack(d, n, gp, dp, s1);
    }
    ctx->nb = slot0 + n;
    ctx->uploaded = true;
    const int32_t *saved = ctx->d.poolofs;
    // phase 1: all n frames at once
    int rc = enqueue_schedule(ctx, slot0, n, 0, s1, false, 1);
    if (dbl && rc == DSM_OK)
    {
        CK(cudaEventRecord(ctx->ev_join[par], s1));
        CK(cudaStreamWaitEvent(s2, ctx->ev_join[par], 0));
    }
    // phase 2: frame by frame on the resident pool, compaction into the alternate buffer, swap
    for (int t = 0; t < n && rc == DSM_OK; t++)
    {
        ctx->d.poolofs = ctx->res_ofs - (slot0 + t); // kernels read poolofs[b], poolofs[b+1] with b == slot0 + t
        int upq = (ctx->res_upper + 65535) / 65536 * 65536;
        if (upq > ctx->p.max_local_surfels) upq = ctx->p.max_local_surfels;
        rc = enqueue_schedule(ctx, slot0 + t, 1, upq, s2, true, 2);
        if (rc != DSM_OK) break;
        dsm_surfel_t *tmp = ctx->d.pool;
        ctx->d.pool = ctx->pool_snap;
        ctx->pool_snap = tmp;
        ctx->res_upper += ctx->S;
    }
    ctx->d.poolofs = saved;
    if (rc != DSM_OK) return rc;
    ctx->ran = true;
    if (dbl)
    {
        CK(cudaEventRecord(ctx->ev_p2, s2));
        CK(cudaStreamWaitEvent(st, ctx->ev_p2, 0));
    }
    CK(cudaEventRecord(ctx->ev_done[e], st));
    CK(cudaEventRecord(ctx->ev_done[0], st)); // dsm_fuse_frame_resident waits on these before reusing slot 0 / 1
    CK(cudaEventRecord(ctx->ev_done[1], st));
    CK(cudaGetLastError());
    // the caller may reuse its image buffers as soon as we return
    CK(cudaEventSynchronize(ctx->ev_h2d[e]));
    if (n_new)
    {
        CK(cudaMemcpyAsync(n_new, ctx->d.nnew + slot0, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return DSM_OK;
}

// ---- parity / debug readback ----
extern "C" int dsm_get_labels(dsm_ctx *ctx, int frame, int32_t *labels_hw)
{
    if (!ctx || !labels_hw || frame < 0 || frame >= ctx->p.max_batch) return DSM_E_INVALID;
    if (!ctx->ran) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    const int W = ctx->p.width, H = ctx->p.height;
    CK(cudaMemcpy2DAsync(labels_hw, (size_t)W * 4, ctx->d.labels + (size_t)frame * ctx->px, (size_t)ctx->Wp * 4,
                         (size_t)W * 4, H, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSM_OK;
}

extern "C" int dsm_get_seeds(dsm_ctx *ctx, int frame, dsm_seed_t *seeds)
{
    if (!ctx || !seeds || frame < 0 || frame >= ctx->p.max_batch) return DSM_E_INVALID;
    if (!ctx->ran) return DSM_E_STATE;
    CK(cudaSetDevice(ctx->device));
    const int plane_done = 12; // kernels of the schedule up to and including the plane fit
    dsm_launch_seeds_export(ctx->d, frame, ctx->seed_export, (ctx->stop_after > 0 && ctx->stop_after < plane_done) ? 1 : 0, ctx->stream);
    CK(cudaMemcpyAsync(seeds, ctx->seed_export, (size_t)ctx->S * sizeof(dsm_seed_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return DSM_OK;
}

extern "C" int dsm_device_buffer(dsm_ctx *ctx, int which, void **dev_ptr, size_t *bytes)
{
    if (!ctx || !dev_ptr || !bytes) return DSM_E_INVALID;
    switch (which)
    {
    case DSM_BUF_NEW_SURFELS:
        *dev_ptr = ctx->d.newsurf;
        *bytes = (size_t)ctx->p.max_batch * ctx->S * sizeof(dsm_surfel_t);
        return DSM_OK;
    case DSM_BUF_NEW_COUNTS:
        *dev_ptr = ctx->d.nnew;
        *bytes = (size_t)ctx->p.max_batch * sizeof(int32_t);
        return DSM_OK;
    case DSM_BUF_LOCAL:
        *dev_ptr = ctx->d.pool;
        *bytes = (size_t)(ctx->p.max_local_surfels > 0 ? ctx->p.max_local_surfels : 1) * sizeof(dsm_surfel_t);
        return DSM_OK;
    default:
        return DSM_E_INVALID;
    }
}
