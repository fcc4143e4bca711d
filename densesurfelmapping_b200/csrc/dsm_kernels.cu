// Hand-written sm_100a kernels for the DenseSurfelMapping per-frame hot path: input repack, seed initialisation,
// surfel fuse / initialise and the GPU-resident pool kernels.  The superpixel and plane-fit kernels live in dsm_tile.cu.
//
// Every kernel takes the frame index from the grid so a batch of independent frames is one launch.  No tensor cores:
// there is no dense contraction anywhere on this path (largest "matrix" is a 4x4 fp64 solve per superpixel).
//
// EXACTNESS CONTRACT: superpixel labels must be bit-identical to the
// serialised reference.  Therefore this file is compiled with -fmad=false (the reference is
// built for baseline x86-64: SSE2, no FMA contraction), keeps IEEE division / sqrt
// (-prec-div/-prec-sqrt defaults, no fast-math, no FTZ) and spells out every float<->double
// promotion exactly where the reference's C++ expressions have them.  All citations
// ":NNN" refer to /root/reference/surfel_fusion/src/fusion_functions.cpp.
#include "dsm_exact.cuh"
#include <cuda_pipeline.h>

// -------------------------------------------------------------------------------------------
// Kin  repack — the caller's tightly packed [n][H][W] gray / depth arrive by ONE contiguous H2D copy
// each (a strided 2-D copy of 1226-byte rows runs at less than half the PCIe rate) and are laid
// out here into the 16-byte-aligned pitched device format every other kernel relies on.
// One thread per 4 output pixels; reads are unaligned scalars (coalesced), writes are vectors.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_repack(const __grid_constant__ DsmDev d, const uint8_t *gray_packed, const float *depth_packed)
{
    const int b = d.frame0 + blockIdx.z; // device frame slot; the packed source holds this chunk's frames from 0
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x4 >= d.Wp || y >= d.H) return;
    const size_t src = ((size_t)blockIdx.z * d.H + y) * d.W + x4;
    const size_t dst = (size_t)b * d.px_stride + (size_t)y * d.Wp + x4;
    uchar4 g = make_uchar4(0, 0, 0, 0);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x4 + 3 < d.W)
    {
        g = make_uchar4(gray_packed[src], gray_packed[src + 1], gray_packed[src + 2], gray_packed[src + 3]);
        z = make_float4(depth_packed[src], depth_packed[src + 1], depth_packed[src + 2], depth_packed[src + 3]);
    }
    else
    {
        if (x4 < d.W) g.x = gray_packed[src], z.x = depth_packed[src];
        if (x4 + 1 < d.W) g.y = gray_packed[src + 1], z.y = depth_packed[src + 1];
        if (x4 + 2 < d.W) g.z = gray_packed[src + 2], z.z = depth_packed[src + 2];
    }
    *reinterpret_cast<uchar4 *>(const_cast<uint8_t *>(d.gray) + dst) = g;
    *reinterpret_cast<float4 *>(const_cast<float *>(d.depth) + dst) = z;
}

// -------------------------------------------------------------------------------------------
// K0  seed_init   — initialize_seeds_kernel (:577-629) + the per-frame clears (:963-965)
// one thread per seed
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_seed_init(const __grid_constant__ DsmDev d)
{
    pdl_enter();
    const int b = d.frame0 + blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        d.nlist[b] = 0;
        d.nnew[b] = 0;
        d.errflag[b] = 0;
        d.done[b] = 0;
    }
    const int W = d.W, H = d.H, Wp = d.Wp;
    const uint8_t *gray = d.gray + (size_t)b * d.px_stride;
    const float *depth = d.depth + (size_t)b * d.px_stride;
    const bool live = s < d.S;
    const int sp_x = live ? s % d.spw : 0, sp_y = live ? s / d.spw : 0;
    int ix = sp_x * DSM_SP + DSM_SP / 2, iy = sp_y * DSM_SP + DSM_SP / 2;
    ix = ix < W - 1 ? ix : W - 1;
    iy = iy < H - 1 ? iy : H - 1;
    float md = live ? depth[iy * Wp + ix] : 1.0f;
    // seeds sitting on a hole: first depth > 0.01 in raster order of the clamped END-EXCLUSIVE window
    // (:602-625).  The warp serves its hole seeds one at a time, 32 window pixels per step.
    unsigned todo = __ballot_sync(FULL, live && (double)md < 0.01);
    while (todo)
    {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int hx = __shfl_sync(FULL, sp_x, src), hy = __shfl_sync(FULL, sp_y, src);
        int xb = hx * DSM_SP + DSM_SP / 2 - DSM_SP, yb = hy * DSM_SP + DSM_SP / 2 - DSM_SP;
        int xe = xb + DSM_SP * 2, ye = yb + DSM_SP * 2;
        xb = xb > 0 ? xb : 0;
        yb = yb > 0 ? yb : 0;
        xe = xe < W - 1 ? xe : W - 1;
        ye = ye < H - 1 ? ye : H - 1;
        const int ww = xe - xb, n = ww * (ye - yb);
        float found = 0.f;
        bool got = false;
        for (int base = 0; base < n && !got; base += 32)
        {
            const int i = base + lane;
            float t = 0.f;
            if (i < n) t = depth[(yb + i / ww) * Wp + xb + i % ww];
            const unsigned hit = __ballot_sync(FULL, (double)t > 0.01);
            if (hit)
            {
                found = __shfl_sync(FULL, t, __ffs(hit) - 1);
                got = true;
            }
        }
        if (got && lane == src) md = found;
    }
    if (!live) return;
    const size_t o = (size_t)b * d.S + s;
    d.seed[o] = make_float4((float)ix, (float)iy, (float)gray[iy * Wp + ix], md);
    d.inv_md[o] = 1.0 / (double)md; // only consumed when md > 0 (:378)
    d.seed_hl[o] = split_inverse(md);
    d.tstable[o] = -1;              // stable = false
    d.fused[o] = 0;                 // fused = false
}

// -------------------------------------------------------------------------------------------
// Staging of a contiguous run of 44-byte surfel records through shared memory for the streaming pool
// kernels (k_fuse, k_pool_transform).  When the run is 16-byte aligned and a multiple of 16 bytes it
// moves as ONE TMA 1-D bulk copy each way (cp.async.bulk: UBLKCP in SASS, completion on an mbarrier for
// the load, a bulk async-group for the store) issued by a single thread -- no per-thread address math,
// no register staging; otherwise (unaligned slice start, ragged tail) coalesced 4-byte accesses.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ bool stage_in(float *sm, const float *g, int cnt, dsm_barrier *bar, int nthreads)
{
    const unsigned bytes = (unsigned)cnt * 44u;
    const bool bulk = (bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(g) & 15u) == 0u);
    if (bulk)
    {
        if (threadIdx.x == 0)
        {
            init(bar, 1);
            cuda::ptx::fence_proxy_async(cuda::ptx::space_shared); // make the initialised barrier visible to the async proxy
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            cuda::device::memcpy_async_tx(sm, g, cuda::aligned_size_t<16>(bytes), *bar);
            (void)cuda::device::barrier_arrive_tx(*bar, 1, bytes);
        }
        // bounded wait: a bulk copy that never completes (a bad pointer would fault first) traps instead of hanging the GPU
        unsigned spin = 0;
        while (!cuda::ptx::mbarrier_try_wait_parity(cuda::device::barrier_native_handle(*bar), 0))
            if (++spin > (1u << 22)) __trap();
    }
    else
    {
        for (int i = threadIdx.x; i < cnt * 11; i += nthreads) sm[i] = g[i];
        __syncthreads();
    }
    return bulk;
}

__device__ __forceinline__ void stage_out(float *g, const float *sm, int cnt, bool bulk, int nthreads)
{
    if (bulk)
    {
        cuda::ptx::fence_proxy_async(cuda::ptx::space_shared); // this thread's smem writes -> async proxy
        __syncthreads();
        if (threadIdx.x == 0)
        {
            cuda::ptx::cp_async_bulk(cuda::ptx::space_global, cuda::ptx::space_shared, g, sm, (unsigned)cnt * 44u);
            cuda::ptx::cp_async_bulk_commit_group();
            cuda::ptx::cp_async_bulk_wait_group_read(cuda::ptx::n32_t<0>()); // smem must stay alive until it has been read
        }
    }
    else
    {
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 11; i += nthreads) g[i] = sm[i];
    }
}

// -------------------------------------------------------------------------------------------
// K5  surfel_fuse — fuse_surfels_kernel (:190-313).  Pure map over the frame's pool slice.
// The AoS pool (44 B/element, ABI layout) is staged through shared memory with fully coalesced
// 4-byte accesses; each thread then works on its element at stride 11 words (conflict-free).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat4_mul(const float *m, float x, float y, float z, float w, float *o)
{ // column-major, column-by-column accumulation (matches the oracle's Eigen stand-in)
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = ((m[i] * x + m[4 + i] * y) + m[8 + i] * z) + m[12 + i] * w;
}
__device__ __forceinline__ void mat3_mul(const float *m, float x, float y, float z, float *o)
{
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = (m[i] * x + m[4 + i] * y) + m[8 + i] * z;
}
__device__ __forceinline__ float get_weight(float depth)
{ // std::min(1.0 / depth / depth, 1.0) (:99-102); std::min(a,b) = (b < a) ? b : a
    const double w = 1.0 / (double)depth / (double)depth;
    return (float)((1.0 < w) ? 1.0 : w);
}

#define FUSE_BLOCK 256
__global__ void __launch_bounds__(FUSE_BLOCK) k_fuse(const __grid_constant__ DsmDev d)
{
    pdl_enter();
    __shared__ alignas(128) float sm[FUSE_BLOCK * 11];
    __shared__ float s_pose[32];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ dsm_barrier bar;
    const int b = d.frame0 + blockIdx.y;
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int first = begin + blockIdx.x * FUSE_BLOCK;
    if (first >= end) return;
    const int cnt = min(FUSE_BLOCK, end - first);
    float *g = reinterpret_cast<float *>(d.pool + first);
    if (threadIdx.x < 16) s_pose[threadIdx.x] = d.pose[b * 16 + threadIdx.x];
    else if (threadIdx.x < 32) s_pose[threadIdx.x] = d.ipose[b * 16 + threadIdx.x - 16];
    const bool bulk = stage_in(sm, g, cnt, &bar, FUSE_BLOCK);
    __syncthreads(); // s_pose
    if (threadIdx.x < cnt)
    {
        float *e = sm + threadIdx.x * 11;
        const float *pose = s_pose, *inv_pose = s_pose + 16;
        const int ref = d.refidx[b];
        const int W = d.W, H = d.H, Wp = d.Wp;
        const size_t fo = (size_t)b * d.px_stride, so = (size_t)b * d.S;
        int update_times = __float_as_int(e[9]);
        const int last_update = __float_as_int(e[10]);
        do
        {
            if (ref - last_update > 5 && update_times < 5)
            { // remove unstable (:207-211)
                e[9] = __int_as_float(0);
                break;
            }
            if (update_times == 0) break;
            float pc[4];
            mat4_mul(inv_pose, e[0], e[1], e[2], 1.0f, pc);
            if (pc[2] < d.fuse_near || pc[2] > d.fuse_far) break;
            float nc[3];
            mat3_mul(inv_pose, e[3], e[4], e[5], nc);
            const float pu = pc[0] * d.fx / pc[2] + d.cx;
            const float pv = pc[1] * d.fy / pc[2] + d.cy;
            const int ui = (int)((double)pu + 0.5), vi = (int)((double)pv + 0.5);
            if (ui < 1 || ui > W - 2 || vi < 1 || vi > H - 2) break;
            if ((double)pc[2] < (double)d.depth[fo + (size_t)vi * Wp + ui] - 1.0)
            { // occluding the measurement (:239-243)
                e[9] = __int_as_float(0);
                break;
            }
            const int sp = d.labels[fo + (size_t)vi * Wp + ui];
            const float4 *pl = d.plane + (so + sp) * 3;
            const float4 r0 = pl[0]; // n, view_cos
            if (r0.x == 0 && r0.y == 0 && r0.z == 0) break;
            if ((double)r0.w < MAX_ANGLE_COS) break;
            const float4 r1 = pl[1]; // posi, mean_depth
            const float4 r2 = pl[2]; // size, I
            float tol = (float)((double)(pc[2] * pc[2]) / (d.baseline * (double)d.camera_f) * d.disparity_error);
            tol = (double)tol < d.min_tolerate_diff ? (float)d.min_tolerate_diff : tol;
            if (pc[2] < r1.w - tol) break;
            if (pc[2] > r1.w + tol) break;
            const float ndc = nc[0] * r0.x + nc[1] * r0.y + nc[2] * r0.z;
            if ((double)ndc < MAX_ANGLE_COS)
            {
                e[9] = __int_as_float(0);
                break;
            }
            const float ow = e[8];
            const float nw = get_weight(r1.w);
            const float sw = ow + nw;
            float pw[4];
            mat4_mul(pose, r1.x, r1.y, r1.z, 1.0f, pw);
            const float fpx = (e[0] * ow + nw * pw[0]) / sw;
            const float fpy = (e[1] * ow + nw * pw[1]) / sw;
            const float fpz = (e[2] * ow + nw * pw[2]) / sw;
            float fnx = nc[0] * ow + nw * r0.x;
            float fny = nc[1] * ow + nw * r0.y;
            float fnz = nc[2] * ow + nw * r0.z;
            const double nl = (double)sqrtf(fnx * fnx + fny * fny + fnz * fnz);
            fnx = (float)((double)fnx / nl);
            fny = (float)((double)fny / nl);
            fnz = (float)((double)fnz / nl);
            float nwd[3];
            mat3_mul(pose, fnx, fny, fnz, nwd);
            e[0] = fpx, e[1] = fpy, e[2] = fpz;
            e[3] = nwd[0], e[4] = nwd[1], e[5] = nwd[2];
            e[8] = sw;
            e[7] = r2.y;
            const float new_size = (float)((double)r2.x * fabs((double)(r1.w / (d.camera_f * r0.w))));
            if (new_size < e[6]) e[6] = new_size;
            e[10] = __int_as_float(ref);
            e[9] = __int_as_float(update_times + 1);
            d.fused[so + sp] = 1; // idempotent multi-writer store (:311)
        } while (0);
    }
    stage_out(g, sm, cnt, bulk, FUSE_BLOCK);
}

// -------------------------------------------------------------------------------------------
// K6  surfel_init — initialize_surfels (:315-361).  new_surfels come out in seed-index order exactly like the
// reference's serial push_back loop, with several CTAs per frame and no inter-CTA communication: CTA j owns
// seeds [1024 j, 1024 j + 1024) and finds its output offset without any inter-CTA communication by re-evaluating
// the cheap emit predicate over the seeds before its range (<= 36 B per seed, L2-resident), then scans its own
// range.  Same predicate, same per-surfel arithmetic, same order: byte-identical output.
__device__ __forceinline__ bool init_emits(const DsmDev &d, size_t so, int s)
{
    const float4 *pl = d.plane + (so + s) * 3;
    const float4 r0 = pl[0];
    const float md = pl[1].w;
    return !(md == 0) && !d.fused[so + s] && !((double)r0.w < MAX_ANGLE_COS) && !(r0.x == 0 && r0.y == 0 && r0.z == 0);
}
__global__ void __launch_bounds__(1024) k_init_surfels(const __grid_constant__ DsmDev d)
{
    pdl_enter();
    __shared__ int s_warp[32];
    __shared__ int s_before[32];
    __shared__ float s_pose[16];
    const int b = d.frame0 + blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t so = (size_t)b * d.S;
    if (threadIdx.x < 16) s_pose[threadIdx.x] = d.pose[b * 16 + threadIdx.x];
    const int first = blockIdx.x * 1024;
    // emitted surfels before this CTA's range
    int before = 0;
    for (int s = threadIdx.x; s < first; s += 1024) before += init_emits(d, so, s) ? 1 : 0;
    before = __reduce_add_sync(FULL, before);
    const int s = first + threadIdx.x;
    const bool e = s < d.S && init_emits(d, so, s);
    const unsigned bal = __ballot_sync(FULL, e);
    if (lane == 0)
    {
        s_warp[warp] = __popc(bal);
        s_before[warp] = before;
    }
    __syncthreads();
    int pos = 0, total = 0;
    for (int w = 0; w < 32; w++)
    {
        const int c = s_warp[w];
        pos += s_before[w] + (w < warp ? c : 0);
        total += s_before[w] + c;
    }
    if (e)
    {
        pos += __popc(bal & ((1u << lane) - 1));
        const float4 *pl = d.plane + (so + s) * 3;
        const float4 r0 = pl[0], r1 = pl[1], r2 = pl[2];
        float pw[4], nw[3];
        mat4_mul(s_pose, r1.x, r1.y, r1.z, 1.0f, pw);
        mat3_mul(s_pose, r0.x, r0.y, r0.z, nw);
        dsm_surfel_t o;
        o.px = pw[0], o.py = pw[1], o.pz = pw[2];
        o.nx = nw[0], o.ny = nw[1], o.nz = nw[2];
        o.size = (float)((double)r2.x * fabs((double)(r1.w / (d.camera_f * r0.w))));
        o.color = r2.y;
        o.weight = get_weight(r1.w);
        o.update_times = 1;
        o.last_update = d.refidx[b];
        d.newsurf[so + pos] = o;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) d.nnew[b] = total;
}

// -------------------------------------------------------------------------------------------
// GPU-resident pool: what SurfelMap does to `local_surfels` between and
// after the hot-path calls, kept on the device so a stream never round-trips its pool.
//
// k_pool_count / k_pool_scan / k_pool_scatter / k_pool_append — the post-fusion step of
//   SurfelMap::fuse_map (surfel_map.cpp:1077-1109): drop surfels with update_times == 0, add the
//   newly initialised ones.  The reference recycles freed slots from the highest index and swaps
//   the rest with the back, which is inherently sequential; here it is a deterministic ordered
//   compaction (live surfels keep their order, new ones follow), so the resulting pool equals the
//   reference's AS A SET (nothing on the path depends on pool order).
// k_pool_transform — warp_active_surfels_cpu_kernel (surfel_map.cpp:750-789): p <- W p, n <- R_W n
//   for every surfel of the pool after a loop closure.  Pure streaming over the 44-byte records,
//   staged through shared memory like k_fuse.
// -------------------------------------------------------------------------------------------
// selection predicate of the pool kernels: mode 0 = live surfels (update_times != 0, fuse_map post-step);
// mode 1 = live surfels last updated by keyframe `key` (move_add_surfels, surfel_map.cpp:1479-1497);
// mode 2 = surfels with update_times >= key (the publish/save filters, surfel_map.cpp:1158, :1243, :1406, :1429)
__device__ __forceinline__ bool pool_pred(const dsm_surfel_t &e, int mode, int key)
{
    if (mode == 2) return e.update_times >= key;
    return mode == 0 ? (e.update_times != 0) : (e.update_times > 0 && e.last_update == key);
}

__global__ void __launch_bounds__(256) k_pool_count(const __grid_constant__ DsmDev d, int b, int *blkcnt, int mode, int key)
{
    pdl_enter();
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int i = begin + blockIdx.x * 256 + threadIdx.x;
    const bool live = i < end && pool_pred(d.pool[i], mode, key);
    const int c = __syncthreads_count(live);
    if (threadIdx.x == 0) blkcnt[blockIdx.x] = (begin + blockIdx.x * 256 < end) ? c : 0;
}

__global__ void __launch_bounds__(1024) k_pool_scan(const __grid_constant__ DsmDev d, int b, const int *blkcnt, int *blkofs, int *newofs)
{
    pdl_enter();
    __shared__ int s_warp[32];
    __shared__ int s_run;
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int nblk = (end - begin + 255) / 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024)
    {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? blkcnt[i] : 0;
        int wtot;
        const int wex = warp_excl_scan(v, lane, wtot);
        if (lane == 31) s_warp[warp] = wtot;
        __syncthreads();
        const int run = s_run;
        int wofs = 0, tot = 0;
        for (int w = 0; w < 32; w++)
        {
            const int c = s_warp[w];
            if (w < warp) wofs += c;
            tot += c;
        }
        if (i < nblk) blkofs[i] = run + wofs + wex;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        newofs[0] = s_run;                // live surfels kept
        newofs[1] = s_run + d.nnew[b];    // new pool size
    }
}

__global__ void __launch_bounds__(256) k_pool_scatter(const __grid_constant__ DsmDev d, int b, const int *blkofs, dsm_surfel_t *dst, int mode, int key)
{
    pdl_enter();
    __shared__ int s_warp[8];
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int i = begin + blockIdx.x * 256 + threadIdx.x;
    if (begin + blockIdx.x * 256 >= end) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    dsm_surfel_t e;
    bool live = false;
    if (i < end)
    {
        e = d.pool[i];
        live = pool_pred(e, mode, key);
        if (live && mode == 1) d.pool[i].update_times = 0; // "delete the surfel from the local point" (surfel_map.cpp:1496)
    }
    const unsigned bal = __ballot_sync(FULL, live);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wofs = 0;
    for (int w = 0; w < warp; w++) wofs += s_warp[w];
    if (live) dst[blkofs[blockIdx.x] + wofs + __popc(bal & ((1u << lane) - 1))] = e;
}

// k_pool_scatter_cloud — the point-cloud builders of SurfelMap (publish_active_pointcloud surfel_map.cpp:1398-1417,
//   publish_all_pointcloud :1419-1454, publish_neighbor_pointcloud :1284-1300, save_cloud :1153-1173): every
//   selected surfel becomes one PointXYZI {px, py, pz, intensity = color}, in pool order like the serial
//   push_back loops.  Same count/scan as the compaction; 44 B read and 16 B written per surfel.
__global__ void __launch_bounds__(256) k_pool_scatter_cloud(const __grid_constant__ DsmDev d, int b, const int *blkofs, float4 *dst, int mode, int key)
{
    __shared__ int s_warp[8];
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int i = begin + blockIdx.x * 256 + threadIdx.x;
    if (begin + blockIdx.x * 256 >= end) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    bool live = false;
    if (i < end)
    {
        const dsm_surfel_t e = d.pool[i];
        live = pool_pred(e, mode, key);
        pt = make_float4(e.px, e.py, e.pz, e.color);
    }
    const unsigned bal = __ballot_sync(FULL, live);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wofs = 0;
    for (int w = 0; w < warp; w++) wofs += s_warp[w];
    if (live) dst[blkofs[blockIdx.x] + wofs + __popc(bal & ((1u << lane) - 1))] = pt;
}

__global__ void __launch_bounds__(256) k_pool_append(const __grid_constant__ DsmDev d, int b, const int *newofs, dsm_surfel_t *dst, int32_t *ofs_out)
{
    pdl_enter();
    const int n = d.nnew[b];
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[newofs[0] + j] = d.newsurf[(size_t)b * d.S + j];
    if (j == 0 && ofs_out) ofs_out[1] = newofs[1]; // nothing in this grid reads the pool's offset table
}

#define XF_BLOCK 256
__global__ void __launch_bounds__(XF_BLOCK) k_pool_transform(const __grid_constant__ DsmDev d, int b, const float *Wm)
{
    __shared__ alignas(128) float sm[XF_BLOCK * 11];
    __shared__ float s_w[16];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ dsm_barrier bar;
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int first = begin + blockIdx.x * XF_BLOCK;
    if (first >= end) return;
    const int cnt = min(XF_BLOCK, end - first);
    float *g = reinterpret_cast<float *>(d.pool + first);
    if (threadIdx.x < 16) s_w[threadIdx.x] = Wm[threadIdx.x];
    const bool bulk = stage_in(sm, g, cnt, &bar, XF_BLOCK);
    __syncthreads(); // s_w
    if (threadIdx.x < cnt)
    {
        float *e = sm + threadIdx.x * 11;
        float pw[4], nw[3];
        mat4_mul(s_w, e[0], e[1], e[2], 1.0f, pw);
        mat3_mul(s_w, e[3], e[4], e[5], nw);
        e[0] = pw[0], e[1] = pw[1], e[2] = pw[2];
        e[3] = nw[0], e[4] = nw[1], e[5] = nw[2];
    }
    stage_out(g, sm, cnt, bulk, XF_BLOCK);
}

// two-entry offset table {a, b} written in stream order (the pool kernels take their range from device memory)
__global__ void k_set2(int32_t *p, int a, int b)
{
    p[0] = a;
    p[1] = b;
}
void dsm_launch_set2(int32_t *p, int a, int b, cudaStream_t s) { k_set2<<<1, 1, 0, s>>>(p, a, b); }

// -------------------------------------------------------------------------------------------
// parity readback: rebuild the reference's 60-byte Superpixel_seed records for one frame
// -------------------------------------------------------------------------------------------
__global__ void k_seeds_export(const __grid_constant__ DsmDev d, int b, dsm_seed_t *out, int raw_md)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.S) return;
    const size_t o = (size_t)b * d.S + s;
    const float4 sd = d.seed[o];
    const float4 *pl = d.plane + o * 3;
    const float4 r0 = pl[0], r1 = pl[1], r2 = pl[2];
    dsm_seed_t e;
    e.x = sd.x, e.y = sd.y;
    e.size = r2.x;
    e.norm_x = r0.x, e.norm_y = r0.y, e.norm_z = r0.z;
    e.posi_x = r1.x, e.posi_y = r1.y, e.posi_z = r1.z;
    e.view_cos = r0.w;
    e.mean_depth = raw_md ? sd.w : r1.w; // raw_md: clustering state before the plane fit (debug staging)
    e.mean_intensity = sd.z;
    e.fused = d.fused[o] ? 1 : 0;
    e.stable = (d.tstable[o] == DSM_STABLE) ? 1 : 0;
    e._pad[0] = e._pad[1] = 0;
    e.min_eigen_value = 0.f;
    e.max_eigen_value = 0.f;
    out[s] = e;
}

// -------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------
void dsm_launch_repack(const DsmDev &d, int nb, const uint8_t *gray_packed, const float *depth_packed, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.Wp + 255) / 256, (d.H + 3) / 4, nb);
    k_repack<<<grid, block, 0, s>>>(d, gray_packed, depth_packed);
}
void dsm_launch_seed_init(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 255) / 256, nb);
    pdl_launch(k_seed_init, grid, dim3(256), 0, s, d);
}
void dsm_launch_fuse(const DsmDev &d, int nb, cudaStream_t s)
{
    if (d.max_pool_per_frame <= 0) return;
    dim3 grid((d.max_pool_per_frame + FUSE_BLOCK - 1) / FUSE_BLOCK, nb);
    pdl_launch(k_fuse, grid, dim3(FUSE_BLOCK), 0, s, d);
}
void dsm_launch_init_surfels(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 1023) / 1024, nb);
    pdl_launch(k_init_surfels, grid, dim3(1024), 0, s, d);
}
void dsm_launch_seeds_export(const DsmDev &d, int frame, dsm_seed_t *out_dev, int raw_md, cudaStream_t s)
{
    k_seeds_export<<<(d.S + 255) / 256, 256, 0, s>>>(d, frame, out_dev, raw_md);
}

// ofs_out (may be null): the resident pool's offset table {0, n}; the last kernel of the chain stores the new size there
void dsm_launch_pool_compact(const DsmDev &d, int frame, int upper, int *blkcnt, int *blkofs, int *newofs, dsm_surfel_t *dst, int32_t *ofs_out, cudaStream_t s)
{
    const int nblk = (upper + 255) / 256;
    if (nblk > 0) pdl_launch(k_pool_count, dim3(nblk), dim3(256), 0, s, d, frame, blkcnt, 0, 0);
    pdl_launch(k_pool_scan, dim3(1), dim3(1024), 0, s, d, frame, (const int *)blkcnt, blkofs, newofs);
    if (nblk > 0) pdl_launch(k_pool_scatter, dim3(nblk), dim3(256), 0, s, d, frame, (const int *)blkofs, dst, 0, 0);
    pdl_launch(k_pool_append, dim3((d.S + 255) / 256), dim3(256), 0, s, d, frame, (const int *)newofs, dst, ofs_out);
}
// move_add_surfels, removal half (surfel_map.cpp:1479-1497): the live surfels whose last_update == key are
// copied in pool order to dst (count in newofs[0]) and flagged dead in the pool
void dsm_launch_pool_retire(const DsmDev &d, int frame, int upper, int key, int *blkcnt, int *blkofs, int *newofs, dsm_surfel_t *dst, cudaStream_t s)
{
    const int nblk = (upper + 255) / 256;
    if (nblk > 0) k_pool_count<<<nblk, 256, 0, s>>>(d, frame, blkcnt, 1, key);
    k_pool_scan<<<1, 1024, 0, s>>>(d, frame, blkcnt, blkofs, newofs);
    if (nblk > 0 && dst) k_pool_scatter<<<nblk, 256, 0, s>>>(d, frame, blkofs, dst, 1, key); // dst == nullptr: count only
}
// publish/save filters (mode/key as pool_pred): as_cloud ? PointXYZI float4 records : whole 44-byte surfels, in pool
// order to dst (count in newofs[0]); the pool itself is not modified
void dsm_launch_pool_export(const DsmDev &d, int frame, int upper, int mode, int key, bool as_cloud, int *blkcnt, int *blkofs, int *newofs, void *dst, cudaStream_t s)
{
    const int nblk = (upper + 255) / 256;
    if (nblk > 0) k_pool_count<<<nblk, 256, 0, s>>>(d, frame, blkcnt, mode, key);
    k_pool_scan<<<1, 1024, 0, s>>>(d, frame, blkcnt, blkofs, newofs);
    if (nblk == 0) return;
    if (as_cloud)
        k_pool_scatter_cloud<<<nblk, 256, 0, s>>>(d, frame, blkofs, static_cast<float4 *>(dst), mode, key);
    else
        k_pool_scatter<<<nblk, 256, 0, s>>>(d, frame, blkofs, static_cast<dsm_surfel_t *>(dst), mode, key);
}
void dsm_launch_pool_transform(const DsmDev &d, int frame, int upper, const float *Wm_dev, cudaStream_t s)
{
    if (upper <= 0) return;
    k_pool_transform<<<(upper + XF_BLOCK - 1) / XF_BLOCK, XF_BLOCK, 0, s>>>(d, frame, Wm_dev);
}
