// Hand-written sm_100a kernels for the DenseSurfelMapping per-frame hot path.
//
// One kernel (family) per reference phase; every kernel takes the frame index from the grid so
// a batch of independent frames is one launch.  No tensor cores: there is no dense contraction
// anywhere on this path (largest "matrix" is a 4x4 fp64 solve per superpixel).
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

// K0', EXPERIMENTAL (variant bit 6, off by default): k_seed_init whose hole search requests the
// whole window at once.  k_seed_init walks the window of a seed that sits on a depth hole 32 pixels per step with an
// early exit, i.e. up to 8 dependent memory round trips per hole seed and warp; on a frame with large invalid regions
// that is the single slowest kernel of a one-frame stream (42 us).  Same result: first valid depth in raster order.
__global__ void __launch_bounds__(256) k_seed_init_wide(const __grid_constant__ DsmDev d)
{
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
        // all (at most 8 x 32 = 256) window pixels are requested at once; the first hit in raster order is the
        // smallest flat index, found with one warp-wide integer minimum
        float tv[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int i = lane + 32 * k;
            tv[k] = (i < n) ? depth[(yb + i / ww) * Wp + xb + i % ww] : 0.f;
        }
        int first = INT_MAX;
        float val = 0.f;
#pragma unroll
        for (int k = 7; k >= 0; k--)
            if ((double)tv[k] > 0.01) first = lane + 32 * k, val = tv[k]; // descending k: the smallest index of this lane wins
        const int wmin = __reduce_min_sync(FULL, first);
        const float found = __shfl_sync(FULL, val, wmin & 31);
        if (wmin != INT_MAX && lane == src) md = found;
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
// K1  slic_assign — update_pixels_kernel (:389-453) + calculate_cost (:364-387)
//
// Each thread owns 4 horizontally consecutive pixels (one uchar4 / float4 / int4 access each).
// Geometry: a pixel with x%8 = r can only pass the |8c+4-x| < 8 test (:418-420) for seed
// columns {b-1,b} (r<4), {b} (r==4) or {b,b+1} (r>4); same for rows.  So at most 2x2 of the
// reference's 3x3 candidates are ever valid, and the 4 pixels of a thread share them.  The
// candidate visiting order is the reference's: dx outer, dy inner, strict '<' (first wins).
//
// `stable` raster semantics: the winner w(p) never depends on the flags, so it
// is computed for every pixel.  Pixels owned by an UNSTABLE seed (tstable < 0) are always
// evaluated by the reference: commit, and stamp the winner with atomicMin(t[w], idx(p)).
// Pixels owned by a seed that was stable when the pass began are appended to a per-frame list;
// k_relax resolves which of them the sequential raster scan would have evaluated.
// In the first iteration every label is 0 and seed 0 is unstable, so everything commits.
// -------------------------------------------------------------------------------------------
// K1', EXPERIMENTAL (variant bit 7, off by default): the assign pass with 4 instead of 7
// float<->double conversions per (pixel, candidate).  The pass is bound by the conversion unit (XU at 74-85 % of peak,
// profiles/r1_final_pipes.csv).  Two of the reference's roundings to float -- (float)nd at :376 and the final (float) at
// :381 -- only exist to be widened again or compared, so they are done without leaving the fp64 pipe:
//     M = 1.5 * 2^(exponent(x) + 29),   rn24(x) = (x + M) - M
// (the fp64 adder performs the round-to-nearest-even to 24 significant bits; M comes from the high word of x with two
// integer ops).  For non-negative x whose float image is zero or normal this equals (double)(float)x bit for bit
// (tools/check_rn24_magic.py: 1e8 random values, exact ties and their neighbours); costs are >= 0 and either 0 or
// >= 1e-7, values >= 1e6 never win, so overflow to inf and the subnormal grid cannot matter.  The costs stay
// float-valued doubles and are compared as doubles, which orders them exactly like the float comparisons.
__device__ __forceinline__ double rn24(double x)
{
    const double M = __hiloint2double((__double2hiint(x) & 0x7ff00000) + 0x01d80000, 0);
    return (x + M) - M;
}
__device__ __forceinline__ bool calc_cost_x(const SeedC &sd, float pix_i, float pix_inv, double pix_inv_d, float fx, float fy,
                                            double &nodepth, double &withdepth)
{
    const float ax = sd.x - fx, ay = sd.y - fy;
    const float dist = ax * ax + ay * ay;
    const float n = dist * 0.0625f; // (:374)
    const float idf = sd.I - pix_i;
    const double a = (double)(idf * idf);
    const double q0 = a * 0.01;
    const double q = __fma_rn(__fma_rn(-q0, 100.0, a), 0.01, q0); // a / 100.0, correctly rounded (see calc_cost)
    const double nr = rn24((double)n + q);                         // (double)(float)nd of (:376), no conversion
    nodepth = nr;
    const bool has = sd.md > 0 && pix_inv > 0; // (:378)
    const float idd = (float)(sd.inv - pix_inv_d);                 // (:380)
    const double wr = rn24(nr + (double)(idd * idd) * 400.0);      // (:381)
    withdepth = has ? wr : nr;
    return has;
}

template <bool FIRST>
__global__ void __launch_bounds__(256, 4) k_assign_x(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.z;
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int lane = threadIdx.x & 31;
    const bool active = (x4 < d.W) && (y < d.H);

    const size_t fo = (size_t)b * d.px_stride;
    const size_t so = (size_t)b * d.S;
    int win[4] = {-1, -1, -1, -1};
    int L[4] = {0, 0, 0, 0};
    if (active)
    {
        const size_t po = fo + (size_t)y * d.Wp + x4;
        const uchar4 g4 = *reinterpret_cast<const uchar4 *>(d.gray + po);
        const float4 z4 = *reinterpret_cast<const float4 *>(d.depth + po);
        if (!FIRST)
        {
            const int4 l4 = *reinterpret_cast<const int4 *>(d.labels + po);
            L[0] = l4.x, L[1] = l4.y, L[2] = l4.z, L[3] = l4.w;
        }
        const float gi[4] = {(float)g4.x, (float)g4.y, (float)g4.z, (float)g4.w};
        const float zi[4] = {z4.x, z4.y, z4.z, z4.w};
        const int bx = x4 >> 3, by = y >> 3, rx0 = x4 & 7, ry = y & 7;
        const int xa = (rx0 == 0) ? bx - 1 : bx, xb = xa + 1;
        const int ya = (ry < 4) ? by - 1 : by, yb = ya + 1;
        const bool vxa = xa >= 0 && xa < d.spw, vxb = xb >= 0 && xb < d.spw;
        const bool vya = ya >= 0 && ya < d.sph, vyb = (ry != 4) && yb >= 0 && yb < d.sph;
        // candidate c = 2*ix + iy  -> (xa,ya) (xa,yb) (xb,ya) (xb,yb): dx outer, dy inner
        SeedC sc[4];
        bool sv[4];
        int sidx[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            const int cx = (c >> 1) ? xb : xa, cy = (c & 1) ? yb : ya;
            sv[c] = ((c >> 1) ? vxb : vxa) && ((c & 1) ? vyb : vya);
            sidx[c] = cy * d.spw + cx;
            const int li = sv[c] ? sidx[c] : 0; // invalid candidates read seed 0 and are masked out below
            const float4 s4 = d.seed[so + li];
            sc[c].x = s4.x, sc[c].y = s4.y, sc[c].I = s4.z, sc[c].md = s4.w;
            sc[c].inv = d.inv_md[so + li];
        }
        const float fy = (float)y;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int x = x4 + i;
            const float my_i = gi[i];
            // (:404-405) my_inv = (float)(1.0 / (double)depth) for depth > 0.01.  53 >= 2*24+2 bits, so the
            // double rounding is innocuous and the IEEE float reciprocal gives the same value.
            // (__frcp_rn is the correctly rounded reciprocal, subnormal results included; -ftz=false)
            const float my_inv = (zi[i] > F_0p01_LO) ? __frcp_rn(zi[i]) : 0.0f;
            const double my_inv_d = (double)my_inv;
            const float fx = (float)x;
            double min_d = 1e6, min_nd = 1e6; // float-valued doubles: same order as the float comparisons
            int idx_d = -1, idx_nd = -1;
            bool all_has_depth = true;
#pragma unroll
            for (int c = 0; c < 4; c++)
            { // branch-free: an invalid candidate gets cost +inf (never < the running minimum) and does not
              // touch all_has_depth; x%8 == 4 sees only its own column (rx0==4, i==0 -> only xa)
                // a warp covers 128 pixels of ONE image row, so the row part of the validity test is warp-uniform: on rows
                // with y % 8 == 4 (and on the border rows) two of the four candidates are skipped without divergence
                if (!((c & 1) ? vyb : vya)) continue;
                const bool valid = sv[c] && ((c >> 1) ? !(rx0 == 4 && i == 0) : true);
                double cnd, cd;
                const bool has = calc_cost_x(sc[c], my_i, my_inv, my_inv_d, fx, fy, cnd, cd);
                all_has_depth &= has || !valid;
                const bool bd = valid && cd < min_d, bn = valid && cnd < min_nd; // validity folded into the compare predicate
                min_d = bd ? cd : min_d;
                idx_d = bd ? sidx[c] : idx_d;
                min_nd = bn ? cnd : min_nd;
                idx_nd = bn ? sidx[c] : idx_nd;
            }
            win[i] = (x < d.W) ? (all_has_depth ? idx_d : idx_nd) : -1;
        }
    }

    if (FIRST)
    {
        if (active)
        {
            const size_t po = fo + (size_t)y * d.Wp + x4;
            *reinterpret_cast<int4 *>(d.labels + po) = make_int4(win[0] < 0 ? 0 : win[0], win[1] < 0 ? 0 : win[1],
                                                                 win[2] < 0 ? 0 : win[2], win[3] < 0 ? 0 : win[3]);
        }
        return;
    }

    // ---- iterations 2..: commit / defer
    int2 ent[4];
    int nent = 0;
    if (active)
    {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            if (x4 + i >= d.W || win[i] < 0) continue;
            const int pidx = y * d.Wp + x4 + i;
            const int ts = d.tstable[so + L[i]];
            if (ts < 0)
            { // owner unstable since the start of the pass: the reference evaluates this pixel
                if (win[i] != L[i])
                {
                    L[i] = win[i];
                    changed = true;
                }
                if (d.tstable[so + win[i]] > pidx) atomicMin(&d.tstable[so + win[i]], pidx); // stable = false at time pidx (:445/:450)
            }
            else
            {
                ent[nent++] = make_int2(pidx, win[i]);
            }
        }
        if (changed)
        {
            const size_t po = fo + (size_t)y * d.Wp + x4;
            *reinterpret_cast<int4 *>(d.labels + po) = make_int4(L[0], L[1], L[2], L[3]);
        }
    }
    // warp-aggregated append of the deferred pixels
    int total;
    const int excl = warp_excl_scan(nent, lane, total);
    if (total > 0)
    {
        int base = 0;
        if (lane == 31) base = atomicAdd(&d.nlist[b], total);
        base = __shfl_sync(FULL, base, 31);
        int2 *list = d.list + fo;
        for (int j = 0; j < nent; j++) list[base + excl + j] = ent[j];
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(256, 4) k_assign(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.z;
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int lane = threadIdx.x & 31;
    const bool active = (x4 < d.W) && (y < d.H);

    const size_t fo = (size_t)b * d.px_stride;
    const size_t so = (size_t)b * d.S;
    int win[4] = {-1, -1, -1, -1};
    int L[4] = {0, 0, 0, 0};
    if (active)
    {
        const size_t po = fo + (size_t)y * d.Wp + x4;
        const uchar4 g4 = *reinterpret_cast<const uchar4 *>(d.gray + po);
        const float4 z4 = *reinterpret_cast<const float4 *>(d.depth + po);
        if (!FIRST)
        {
            const int4 l4 = *reinterpret_cast<const int4 *>(d.labels + po);
            L[0] = l4.x, L[1] = l4.y, L[2] = l4.z, L[3] = l4.w;
        }
        const float gi[4] = {(float)g4.x, (float)g4.y, (float)g4.z, (float)g4.w};
        const float zi[4] = {z4.x, z4.y, z4.z, z4.w};
        const int bx = x4 >> 3, by = y >> 3, rx0 = x4 & 7, ry = y & 7;
        const int xa = (rx0 == 0) ? bx - 1 : bx, xb = xa + 1;
        const int ya = (ry < 4) ? by - 1 : by, yb = ya + 1;
        const bool vxa = xa >= 0 && xa < d.spw, vxb = xb >= 0 && xb < d.spw;
        const bool vya = ya >= 0 && ya < d.sph, vyb = (ry != 4) && yb >= 0 && yb < d.sph;
        // candidate c = 2*ix + iy  -> (xa,ya) (xa,yb) (xb,ya) (xb,yb): dx outer, dy inner
        SeedC sc[4];
        bool sv[4];
        int sidx[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            const int cx = (c >> 1) ? xb : xa, cy = (c & 1) ? yb : ya;
            sv[c] = ((c >> 1) ? vxb : vxa) && ((c & 1) ? vyb : vya);
            sidx[c] = cy * d.spw + cx;
            const int li = sv[c] ? sidx[c] : 0; // invalid candidates read seed 0 and are masked out below
            const float4 s4 = d.seed[so + li];
            sc[c].x = s4.x, sc[c].y = s4.y, sc[c].I = s4.z, sc[c].md = s4.w;
            sc[c].inv = d.inv_md[so + li];
        }
        const float fy = (float)y;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int x = x4 + i;
            const float my_i = gi[i];
            // (:404-405) my_inv = (float)(1.0 / (double)depth) for depth > 0.01.  53 >= 2*24+2 bits, so the
            // double rounding is innocuous and the IEEE float reciprocal gives the same value.
            // (__frcp_rn is the correctly rounded reciprocal, subnormal results included; -ftz=false)
            const float my_inv = (zi[i] > F_0p01_LO) ? __frcp_rn(zi[i]) : 0.0f;
            const double my_inv_d = (double)my_inv;
            const float fx = (float)x;
            float min_d = 1e6f, min_nd = 1e6f;
            int idx_d = -1, idx_nd = -1;
            bool all_has_depth = true;
#pragma unroll
            for (int c = 0; c < 4; c++)
            { // branch-free: an invalid candidate gets cost +inf (never < the running minimum) and does not
              // touch all_has_depth; x%8 == 4 sees only its own column (rx0==4, i==0 -> only xa)
                const bool valid = sv[c] && ((c >> 1) ? !(rx0 == 4 && i == 0) : true);
                float cnd, cd;
                const bool has = calc_cost(sc[c], my_i, my_inv, my_inv_d, fx, fy, cnd, cd);
                cd = valid ? cd : __int_as_float(0x7f800000);
                cnd = valid ? cnd : __int_as_float(0x7f800000);
                all_has_depth &= has || !valid;
                const bool bd = cd < min_d, bn = cnd < min_nd;
                min_d = bd ? cd : min_d;
                idx_d = bd ? sidx[c] : idx_d;
                min_nd = bn ? cnd : min_nd;
                idx_nd = bn ? sidx[c] : idx_nd;
            }
            win[i] = (x < d.W) ? (all_has_depth ? idx_d : idx_nd) : -1;
        }
    }

    if (FIRST)
    {
        if (active)
        {
            const size_t po = fo + (size_t)y * d.Wp + x4;
            *reinterpret_cast<int4 *>(d.labels + po) = make_int4(win[0] < 0 ? 0 : win[0], win[1] < 0 ? 0 : win[1],
                                                                 win[2] < 0 ? 0 : win[2], win[3] < 0 ? 0 : win[3]);
        }
        return;
    }

    // ---- iterations 2..: commit / defer
    int2 ent[4];
    int nent = 0;
    if (active)
    {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            if (x4 + i >= d.W || win[i] < 0) continue;
            const int pidx = y * d.Wp + x4 + i;
            const int ts = d.tstable[so + L[i]];
            if (ts < 0)
            { // owner unstable since the start of the pass: the reference evaluates this pixel
                if (win[i] != L[i])
                {
                    L[i] = win[i];
                    changed = true;
                }
                if (d.tstable[so + win[i]] > pidx) atomicMin(&d.tstable[so + win[i]], pidx); // stable = false at time pidx (:445/:450)
            }
            else
            {
                ent[nent++] = make_int2(pidx, win[i]);
            }
        }
        if (changed)
        {
            const size_t po = fo + (size_t)y * d.Wp + x4;
            *reinterpret_cast<int4 *>(d.labels + po) = make_int4(L[0], L[1], L[2], L[3]);
        }
    }
    // warp-aggregated append of the deferred pixels
    int total;
    const int excl = warp_excl_scan(nent, lane, total);
    if (total > 0)
    {
        int base = 0;
        if (lane == 31) base = atomicAdd(&d.nlist[b], total);
        base = __shfl_sync(FULL, base, 31);
        int2 *list = d.list + fo;
        for (int j = 0; j < nent; j++) list[base + excl + j] = ent[j];
    }
}

// -------------------------------------------------------------------------------------------
// K1r relax — exact resolution of the raster-order `stable` semantics
//
// For a deferred pixel p (owner L(p) was stable at pass start): the sequential scan evaluates p
// iff some earlier-evaluated pixel q < p chose L(p) as its winner, i.e. iff t[L(p)] < idx(p)
// where t[s] = min raster index of an evaluated pixel with winner s.  Jacobi iteration from
// above with atomicMin is monotone and its fixed point is the unique causal solution.
// One CTA per frame; the list is usually tiny (pixels of the few seeds that went stable).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_relax(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.x;
    const int n = d.nlist[b];
    if (n == 0) return;
    const size_t fo = (size_t)b * d.px_stride;
    int2 *list = d.list + fo;
    int32_t *labels = d.labels + fo;
    int32_t *t = d.tstable + (size_t)b * d.S;
    for (;;)
    {
        int changed = 0;
        for (int e = threadIdx.x; e < n; e += blockDim.x)
        {
            const int2 en = list[e];
            if (en.x < 0) continue; // already evaluated
            const int owner = labels[en.x];
            if (__ldcg(&t[owner]) < en.x)
            {
                labels[en.x] = en.y;
                list[e].x = -1;
                if (__ldcg(&t[en.y]) > en.x) atomicMin(&t[en.y], en.x);
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// -------------------------------------------------------------------------------------------
// K2  slic_update — update_seeds_kernel (:468-562), as two kernels
//
// K2a k_gather_depths (warp per seed): scans the seed's clamped 16x16 window (lane = 2*row+half,
//   8 pixels per lane = raster order), reduces the exactly representable integer sums (count,
//   sum x, sum y, sum intensity: all < 2^24, so the reference's float accumulation is exact and
//   order-free) with REDUX, and scan-compacts the member depths (> 0.1) IN RASTER ORDER into a
//   global list laid out [k][seed].
// K2b k_newton (thread per seed): the label-affecting float sums -- sum_depth (:511) and the
//   Huber-Newton sum_a (:536-549) -- are order-sensitive, so one thread walks its
//   seed's list sequentially exactly like the reference; 32 neighbouring seeds read the [k][seed]
//   list with coalesced loads and all seeds of a batch are in flight at once, which hides the
//   dependent-add latency.
// The reference's early `return` for a non-stable seed without members (:516-517) cannot
// fire for a supported shape: the pixel at (8sx+4, 8sy+4) has seed s as its ONLY candidate
// (x%8 == y%8 == 4) and lies inside the counted window, so every seed always owns >= 1 pixel
// (tests/test_oracle.py::test_every_seed_keeps_its_centre_pixel).  The kernel still counts such an
// event in errflag so that the parity tests would expose it.
// -------------------------------------------------------------------------------------------
#define DL_CAP 228 // >= 15*15 possible members

// Lane layout of the two window-scan kernels: the 16x16 window is read in two passes of 8 rows;
// in a pass lane = 4*row + quarter owns 4 consecutive pixels (one 16-byte load per array), so a
// load instruction touches 8 cache lines and (pass, lane, k) lexicographic order == raster order.
// Grid: x = groups of 8 seed columns, y = seed row, z = frame -- no integer division anywhere.
__global__ void __launch_bounds__(256) k_gather_depths(const __grid_constant__ DsmDev d)
{
    // tile[k][seed-in-block]: the 8 warps (one seed each) compact into shared memory, then the block
    // copies the tile out as full 32-byte sectors of the [k][seed] global list (a direct scatter
    // would cost one L2 write request per element)
    __shared__ float tile[DL_CAP * 8];
    __shared__ int s_rows;
    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp_x = blockIdx.x * 8 + warp, sp_y = blockIdx.y;
    const int s = sp_y * d.spw + sp_x;
    if (threadIdx.x == 0) s_rows = 0;
    __syncthreads();
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t so = (size_t)b * d.S;
    // The stable flag is fetched together with the window (not before it): one memory round trip per block
    // instead of two; the few stable seeds just discard what was loaded (:478-479).
    const bool inside = sp_x < d.spw;
    const int tflag = d.tstable[so + (inside ? s : 0)];
    if (inside) // warp-uniform
    {
        const size_t fo = (size_t)b * d.px_stride;
        const int32_t *lab = d.labels + fo; // per-frame bases once; 32-bit element offsets below
        const float *dep = d.depth + fo;
        const uint8_t *gry = d.gray + fo;
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xb = x0 > 0 ? x0 : 0, yb = y0 > 0 ? y0 : 0;
        const int xe = (x0 + 16) < W - 1 ? (x0 + 16) : W - 1; // end-exclusive: last row/col never visited (:488-489)
        const int ye = (y0 + 16) < H - 1 ? (y0 + 16) : H - 1;
        const int xq = x0 + 4 * (lane & 3);
        const bool colin = xq >= 0 && xq < Wp;
        int4 l4[2];
        float4 z4[2];
        uchar4 g4[2];
        int yy[2];
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        { // all six loads are issued before any of them is consumed
            const int y = y0 + 8 * ps + (lane >> 2);
            yy[ps] = y;
            const bool in = colin && y >= yb && y < ye;
            const unsigned po = in ? (unsigned)(y * Wp + xq) : 0u;
            l4[ps] = *reinterpret_cast<const int4 *>(lab + po); // out-of-window lanes read element 0 and are masked below
            z4[ps] = *reinterpret_cast<const float4 *>(dep + po);
            g4[ps] = *reinterpret_cast<const uchar4 *>(gry + po);
            if (!in || tflag == DSM_STABLE) l4[ps] = make_int4(-1, -1, -1, -1);
        }
        unsigned mdm = 0; // bit 4*ps+k: member with depth > 0.1
        int cnt2 = 0;     // member count, pass 0 in the low half-word, pass 1 in the high one
        int sumx = 0, sumy = 0, sumi = 0;
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        {
            const int lk[4] = {l4[ps].x, l4[ps].y, l4[ps].z, l4[ps].w};
            const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
            const int gk[4] = {g4[ps].x, g4[ps].y, g4[ps].z, g4[ps].w};
            int c = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int x = xq + k;
                const bool mem = lk[k] == s && x >= xb && x < xe;
                c += mem ? 1 : 0;
                sumx += mem ? x : 0;
                sumi += mem ? gk[k] : 0;
                if (mem && zk[k] > F_0p1_LO) mdm |= 1u << (4 * ps + k); // (double)depth > 0.1 (:508)
            }
            cnt2 += c;
            sumy += c * yy[ps];
        }
        const int cnt = __reduce_add_sync(FULL, cnt2);
        const int tsx = __reduce_add_sync(FULL, sumx);
        const int tsy = __reduce_add_sync(FULL, sumy);
        const int tsi = __reduce_add_sync(FULL, sumi);
        // one scan for both passes: pass-0 count in the low half-word, pass-1 count in the high one
        const int c2 = __popc(mdm & 0xfu) | (__popc(mdm >> 4) << 16);
        int tot2;
        const int ex2 = warp_excl_scan(c2, lane, tot2);
        const int n0 = tot2 & 0xffff, ndt = n0 + (tot2 >> 16);
        const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile) + 4u * warp;
        unsigned a0 = tbase + 32u * (ex2 & 0xffff), a1 = tbase + 32u * (n0 + (ex2 >> 16)); // 32 bytes per list row
        const float za[4] = {z4[0].x, z4[0].y, z4[0].z, z4[0].w}, zb[4] = {z4[1].x, z4[1].y, z4[1].z, z4[1].w};
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if ((mdm >> k) & 1u)
            {
                sts_f32(a0, za[k]);
                a0 += 32u;
            }
            if ((mdm >> (4 + k)) & 1u)
            {
                sts_f32(a1, zb[k]);
                a1 += 32u;
            }
        }
        if (lane == 0 && tflag != DSM_STABLE)
        {
            d.usum[so + s] = make_int4(cnt, tsx, tsy, tsi);
            d.und[so + s] = ndt;
            atomicMax(&s_rows, ndt);
        }
    }
    __syncthreads();
    const int rows = s_rows;
    const int c = threadIdx.x & 7;
    if (blockIdx.x * 8 + c < d.spw)
    {
        float *dst = d.dlist + (size_t)b * DL_CAP * d.Sp + sp_y * d.spw + blockIdx.x * 8 + c;
        for (int r = threadIdx.x >> 3; r < rows; r += 32) dst[(size_t)r * d.Sp] = tile[r * 8 + c];
    }
}

__global__ void __launch_bounds__(128) k_newton(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) d.nlist[b] = 0; // the deferred-pixel list of this pass is consumed
    if (s >= d.S) return;
    const size_t so = (size_t)b * d.S;
    if (d.tstable[so + s] == DSM_STABLE) return; // untouched by update_seeds
    const int4 su = d.usum[so + s];
    const int n = su.x;
    if (n == 0)
    { // unreachable for supported shapes (see above); recorded, never silently ignored
        atomicAdd(&d.errflag[b], 1);
        d.tstable[so + s] = -1;
        return;
    }
    const float fn = (float)n; // sums below are < 2^24 so the reference's float accumulation is exact
    const float mi = (float)su.w / fn;
    const float mx = (float)su.y / fn;
    const float my = (float)su.z / fn;
    const float4 pre = d.seed[so + s];
    // ::fabs(double): float differences, summed in double, rounded once (:527)
    const float diff = (float)(fabs((double)(pre.z - mi)) + fabs((double)(pre.x - mx)) + fabs((double)(pre.y - my)));
    const bool newstable = diff < F_0p2_HI; // (double)diff < 0.2 (:528)
    const int nd = d.und[so + s];
    float md = 0.0f;
    if (nd > 0)
    {
        const float *dl = d.dlist + (size_t)b * DL_CAP * d.Sp + s;
        const size_t st = (size_t)d.Sp;
        // The adds are a serial dependence chain (that IS the reference's rounding order), but the
        // loads are independent: fetch 8 list entries at a time so 8 requests are in flight per lane.
        // The list stays in registers for the first 32 entries (most seeds' whole first pass).
        float sum_d = 0.0f;
        {
            const float *pp = dl;
            int k = 0;
            for (; k + 8 <= nd; k += 8, pp += 8 * st)
            {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = pp[j * st];
#pragma unroll
                for (int j = 0; j < 8; j++) sum_d += v[j]; // raster order (:511)
            }
            for (; k < nd; k++, pp += st) sum_d += *pp;
        }
        md = sum_d / (float)nd;
        for (int it = 0; it < 5; it++)
        { // damped Huber-Newton (:534-554)
            float sa = 0.0f, sb = 0.0f;
            const float *pp = dl;
            int k = 0;
            for (; k + 8 <= nd; k += 8, pp += 8 * st)
            {
                float r[8];
                bool allin = true;
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    r[j] = md - pp[j * st];
                    allin &= r[j] < F_0p4_HI && r[j] > -F_0p4_HI; // (double)r < 0.4 && (double)r > -0.4
                }
                if (allin)
                { // common case: every residual inside the Huber range -> pure float chain
#pragma unroll
                    for (int j = 0; j < 8; j++) sa += 2 * r[j];
                    sb += 16; // eight exact +2 steps
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        if (r[j] < F_0p4_HI && r[j] > -F_0p4_HI)
                        {
                            sa += 2 * r[j];
                            sb += 2;
                        }
                        else
                            sa = (float)((double)sa + (r[j] > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
                    }
                }
            }
            for (; k < nd; k++, pp += st)
            {
                const float r = md - *pp;
                if (r < F_0p4_HI && r > -F_0p4_HI)
                {
                    sa += 2 * r;
                    sb += 2;
                }
                else
                    sa = (float)((double)sa + (r > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
            }
            const float delta = (float)((double)(-sa) / ((double)sb + 10.0));
            md = md + delta;
            if (delta < F_0p01_HI && delta > -F_0p01_HI) break; // |delta| < 0.01 in double (:552)
        }
    }
    d.seed[so + s] = make_float4(mx, my, mi, md);
    d.inv_md[so + s] = 1.0 / (double)md;
    d.tstable[so + s] = newstable ? DSM_STABLE : -1;
}

// -------------------------------------------------------------------------------------------
// K2a', EXPERIMENTAL (variant bit 1, off by default): k_gather_depths with the window data staged
// through shared memory by TMA.  The direct-load kernel issues six 16-byte loads per lane whose 32 lanes touch 8
// image rows each, i.e. 48 L1 wavefronts per seed window, and neighbouring windows fetch their common 8 columns
// twice.  Here warp 0 fetches the block's whole 16 x 72 strip with 48 1-D bulk copies (cp.async.bulk, completion on
// an mbarrier), and the 8 warps read their windows with conflict-free LDS.128 (row stride 80 words: the two rows
// of a quarter-warp fall into different halves of the 32 banks).  Everything after the loads is the text of
// k_gather_depths, so the results are bit-identical.
// -------------------------------------------------------------------------------------------
#define GT_STRIDE 80   // int32 / float elements per tile row (72 used)
#define GT_GSTRIDE 144 // bytes per gray tile row (96 used; 36 words: 8 rows x 4 quarters hit 32 distinct banks)
__global__ void __launch_bounds__(256) k_gather_depths_tiled(const __grid_constant__ DsmDev d)
{
    // tile[seed-in-block][k]: the 8 warps (one seed each) compact into shared memory, then the block
    // copies the tile out as full 32-byte sectors of the [k][seed] global list
    __shared__ float tile[DL_CAP * 8];
    __shared__ alignas(128) int32_t t_lab[16 * GT_STRIDE];
    __shared__ alignas(128) float t_dep[16 * GT_STRIDE];
    __shared__ alignas(128) uint8_t t_gry[16 * GT_GSTRIDE];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ dsm_barrier bar;
    __shared__ int s_rows;
    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp_x = blockIdx.x * 8 + warp, sp_y = blockIdx.y;
    const int s = sp_y * d.spw + sp_x;
    if (threadIdx.x == 0)
    {
        s_rows = 0;
        init(&bar, 1);
        cuda::ptx::fence_proxy_async(cuda::ptx::space_shared); // make the initialised barrier visible to the async proxy
    }
    __syncthreads();
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t so = (size_t)b * d.S;
    // ---- the block's 16 x 72 pixel strip (8 seed windows of 16 columns, neighbours overlap by 8), one 1-D bulk
    // copy (TMA) per image row and array: labels / depth rows of <= 288 B, gray rows of <= 96 B starting at the
    // 16-byte boundary left of the strip.  Sources, destinations and sizes are multiples of 16 bytes because
    // Wp % 16 == 0 and the strip starts at 64*bx - 4.  Rows above / below the image and columns past the pitch are
    // simply not copied: every consumer below masks them exactly like the direct-load kernel does.
    const int X0 = blockIdx.x * 64 - 4, Y0 = sp_y * DSM_SP - DSM_SP / 2;
    if (warp == 0)
    {
        const size_t fo = (size_t)b * d.px_stride;
        const int ya = Y0 > 0 ? Y0 : 0, yz = (Y0 + 16) < H ? (Y0 + 16) : H;
        const int xs = X0 > 0 ? X0 : 0, xt = (X0 + 72) < Wp ? (X0 + 72) : Wp;
        const int gs = (X0 - 12) > 0 ? (X0 - 12) : 0, gt = (X0 + 84) < Wp ? (X0 + 84) : Wp; // X0 - 12 = 64*bx - 16
        const unsigned nb4 = (unsigned)(xt - xs) * 4u, nbg = (unsigned)(gt - gs);
        if (lane == 0) (void)cuda::device::barrier_arrive_tx(bar, 1, (size_t)(yz - ya) * (2u * nb4 + nbg));
        __syncwarp();
        const int y = Y0 + lane;
        if (lane < 16 && y >= ya && y < yz)
        {
            const size_t ro = fo + (size_t)y * Wp;
            cuda::device::memcpy_async_tx(t_lab + lane * GT_STRIDE + (xs - X0), d.labels + ro + xs, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_dep + lane * GT_STRIDE + (xs - X0), d.depth + ro + xs, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_gry + lane * GT_GSTRIDE + (gs - (X0 - 12)), d.gray + ro + gs, cuda::aligned_size_t<16>(nbg), bar);
        }
    }
    // The stable flag is fetched together with the window (not before it): one memory round trip per block
    // instead of two; the few stable seeds just discard what was loaded (:478-479).
    const bool inside = sp_x < d.spw;
    const int tflag = d.tstable[so + (inside ? s : 0)];
    if (inside) // warp-uniform
    {
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xb = x0 > 0 ? x0 : 0, yb = y0 > 0 ? y0 : 0;
        const int xe = (x0 + 16) < W - 1 ? (x0 + 16) : W - 1; // end-exclusive: last row/col never visited (:488-489)
        const int ye = (y0 + 16) < H - 1 ? (y0 + 16) : H - 1;
        const int xq = x0 + 4 * (lane & 3);
        const bool colin = xq >= 0 && xq < Wp;
        int4 l4[2];
        float4 z4[2];
        uchar4 g4[2];
        int yy[2];
        tile_wait(bar);
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        { // tile row r = 8*ps + lane/4, tile column 8*warp + 4*(lane&3); rows / columns that were not copied hold
          // unspecified data and are masked by `in` exactly like the out-of-window lanes of the direct-load kernel
            const int r = 8 * ps + (lane >> 2);
            const int y = y0 + r;
            yy[ps] = y;
            const bool in = colin && y >= yb && y < ye;
            const int c = 8 * warp + 4 * (lane & 3);
            l4[ps] = *reinterpret_cast<const int4 *>(t_lab + r * GT_STRIDE + c);
            z4[ps] = *reinterpret_cast<const float4 *>(t_dep + r * GT_STRIDE + c);
            g4[ps] = *reinterpret_cast<const uchar4 *>(t_gry + r * GT_GSTRIDE + c + 12);
            if (!in || tflag == DSM_STABLE) l4[ps] = make_int4(-1, -1, -1, -1);
        }
        unsigned mdm = 0; // bit 4*ps+k: member with depth > 0.1
        int cnt2 = 0;     // member count, pass 0 in the low half-word, pass 1 in the high one
        int sumx = 0, sumy = 0, sumi = 0;
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        {
            const int lk[4] = {l4[ps].x, l4[ps].y, l4[ps].z, l4[ps].w};
            const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
            const int gk[4] = {g4[ps].x, g4[ps].y, g4[ps].z, g4[ps].w};
            int c = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int x = xq + k;
                const bool mem = lk[k] == s && x >= xb && x < xe;
                c += mem ? 1 : 0;
                sumx += mem ? x : 0;
                sumi += mem ? gk[k] : 0;
                if (mem && zk[k] > F_0p1_LO) mdm |= 1u << (4 * ps + k); // (double)depth > 0.1 (:508)
            }
            cnt2 += c;
            sumy += c * yy[ps];
        }
        const int cnt = __reduce_add_sync(FULL, cnt2);
        const int tsx = __reduce_add_sync(FULL, sumx);
        const int tsy = __reduce_add_sync(FULL, sumy);
        const int tsi = __reduce_add_sync(FULL, sumi);
        // one scan for both passes: pass-0 count in the low half-word, pass-1 count in the high one
        const int c2 = __popc(mdm & 0xfu) | (__popc(mdm >> 4) << 16);
        int tot2;
        const int ex2 = warp_excl_scan(c2, lane, tot2);
        const int n0 = tot2 & 0xffff, ndt = n0 + (tot2 >> 16);
        // list tile laid out [seed][k] here (the direct-load kernel uses [k][seed]): the lanes of one compaction store
        // write consecutive list positions = consecutive banks; ncu counted 5.3 M bank-conflict wavefronts per launch
        // for the [k][seed] stores (4 banks per warp), about a fifth of the kernel's L1 data-pipe load
        const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile) + 4u * DL_CAP * warp;
        unsigned a0 = tbase + 4u * (ex2 & 0xffff), a1 = tbase + 4u * (n0 + (ex2 >> 16));
        const float za[4] = {z4[0].x, z4[0].y, z4[0].z, z4[0].w}, zb[4] = {z4[1].x, z4[1].y, z4[1].z, z4[1].w};
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if ((mdm >> k) & 1u)
            {
                sts_f32(a0, za[k]);
                a0 += 4u;
            }
            if ((mdm >> (4 + k)) & 1u)
            {
                sts_f32(a1, zb[k]);
                a1 += 4u;
            }
        }
        if (lane == 0 && tflag != DSM_STABLE)
        {
            d.usum[so + s] = make_int4(cnt, tsx, tsy, tsi);
            d.und[so + s] = ndt;
            atomicMax(&s_rows, ndt);
        }
    }
    __syncthreads();
    const int rows = s_rows;
    const int c = threadIdx.x & 7;
    if (blockIdx.x * 8 + c < d.spw)
    {
        float *dst = d.dlist + (size_t)b * DL_CAP * d.Sp + sp_y * d.spw + blockIdx.x * 8 + c;
        for (int r = threadIdx.x >> 3; r < rows; r += 32) dst[(size_t)r * d.Sp] = tile[c * DL_CAP + r]; // DL_CAP % 32 == 4: conflict-free
    }
}

// -------------------------------------------------------------------------------------------
// K2b', EXPERIMENTAL (variant bit 0, off by default): k_newton with the member-depth list
// staged ONCE into shared memory.  k_newton walks its [k][seed] list six times (mean + up to five Huber-
// Newton passes); with 16 CTAs per SM the lists do not fit L1, so every pass comes from L2 again
// (~54 MB per pass per 32-frame launch, the kernel runs at ~5 TB/s of L2 reads).  Here each thread
// copies its own list column global -> shared with 4-byte cp.async (LDGSTS: no register staging, all
// copies of a thread in flight at once), waits for its own copies only (a thread never reads another
// thread's column, so no CTA barrier is needed) and then runs the identical arithmetic, in the identical
// order, out of shared memory ([k][thread] layout: conflict-free).  Entries past NS_ROWS stay in global.
// -------------------------------------------------------------------------------------------
#define NS_ROWS 128
#define NS_THREADS 64
__global__ void __launch_bounds__(NS_THREADS) k_newton_staged(const __grid_constant__ DsmDev d)
{
    __shared__ float col[NS_ROWS * NS_THREADS]; // 32 KB
    const int b = d.frame0 + blockIdx.y;
    const int s = blockIdx.x * NS_THREADS + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) d.nlist[b] = 0; // the deferred-pixel list of this pass is consumed
    if (s >= d.S) return;
    const size_t so = (size_t)b * d.S;
    if (d.tstable[so + s] == DSM_STABLE) return; // untouched by update_seeds
    const int4 su = d.usum[so + s];
    const int n = su.x;
    if (n == 0)
    {
        atomicAdd(&d.errflag[b], 1);
        d.tstable[so + s] = -1;
        return;
    }
    const int nd = d.und[so + s];
    const float *dl = d.dlist + (size_t)b * DL_CAP * d.Sp + s;
    const size_t st = (size_t)d.Sp;
    float *mycol = col + threadIdx.x;
    const int ns = nd < NS_ROWS ? nd : NS_ROWS;
    {
        const float *pp = dl;
        float *q = mycol;
        for (int k = 0; k < ns; k++, pp += st, q += NS_THREADS) __pipeline_memcpy_async(q, pp, 4);
        __pipeline_commit();
    }
    const float fn = (float)n; // sums below are < 2^24 so the reference's float accumulation is exact
    const float mi = (float)su.w / fn;
    const float mx = (float)su.y / fn;
    const float my = (float)su.z / fn;
    const float4 pre = d.seed[so + s];
    const float diff = (float)(fabs((double)(pre.z - mi)) + fabs((double)(pre.x - mx)) + fabs((double)(pre.y - my)));
    const bool newstable = diff < F_0p2_HI; // (double)diff < 0.2 (:528)
    __pipeline_wait_prior(0);
    float md = 0.0f;
    if (nd > 0)
    {
        // The list is walked as two ranges -- [0, ns) from shared memory, [ns, nd) from global memory -- each in
        // groups of 8 like k_newton.  The grouping only batches the loads: every update of sum_d / sa / sb
        // happens in list order, and "sb += 16" equals eight exact "+= 2" steps, so the values are identical.
        auto fsm = [&](int k) -> float { return mycol[k * NS_THREADS]; };
        auto fgl = [&](int k) -> float { return dl[(size_t)k * st]; };
        float sum_d = 0.0f;
        auto mean_range = [&](int k0, int k1, auto get)
        {
            int k = k0;
            for (; k + 8 <= k1; k += 8)
            {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = get(k + j);
#pragma unroll
                for (int j = 0; j < 8; j++) sum_d += v[j]; // raster order (:511)
            }
            for (; k < k1; k++) sum_d += get(k);
        };
        mean_range(0, ns, fsm);
        mean_range(ns, nd, fgl);
        md = sum_d / (float)nd;
        for (int it = 0; it < 5; it++)
        { // damped Huber-Newton (:534-554), expression by expression as in k_newton
            float sa = 0.0f, sb = 0.0f;
            auto newton_range = [&](int k0, int k1, auto get)
            {
                int k = k0;
                for (; k + 8 <= k1; k += 8)
                {
                    float r[8];
                    bool allin = true;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        r[j] = md - get(k + j);
                        allin &= r[j] < F_0p4_HI && r[j] > -F_0p4_HI; // (double)r < 0.4 && (double)r > -0.4
                    }
                    if (allin)
                    {
#pragma unroll
                        for (int j = 0; j < 8; j++) sa += 2 * r[j];
                        sb += 16; // eight exact +2 steps
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 8; j++)
                        {
                            if (r[j] < F_0p4_HI && r[j] > -F_0p4_HI)
                            {
                                sa += 2 * r[j];
                                sb += 2;
                            }
                            else
                                sa = (float)((double)sa + (r[j] > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
                        }
                    }
                }
                for (; k < k1; k++)
                {
                    const float r = md - get(k);
                    if (r < F_0p4_HI && r > -F_0p4_HI)
                    {
                        sa += 2 * r;
                        sb += 2;
                    }
                    else
                        sa = (float)((double)sa + (r > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
                }
            };
            newton_range(0, ns, fsm);
            newton_range(ns, nd, fgl);
            const float delta = (float)((double)(-sa) / ((double)sb + 10.0));
            md = md + delta;
            if (delta < F_0p01_HI && delta > -F_0p01_HI) break; // |delta| < 0.01 in double (:552)
        }
    }
    d.seed[so + s] = make_float4(mx, my, mi, md);
    d.inv_md[so + s] = 1.0 / (double)md;
    d.tstable[so + s] = newstable ? DSM_STABLE : -1;
}

// -------------------------------------------------------------------------------------------
// K3  backproject_normals — calculate_spaces_kernel (:644-662) + calculate_pixels_norms_kernel
// (:664-712), fused: the reference's 24 B/px fp64 space_map is never materialised.
//
// Streaming, pixel-parallel: one thread owns 4 consecutive pixels, reads the depth row and the
// row below with 16-byte loads and writes the three normal planes with 16-byte stores.  The
// per-column / per-row factors (u-cx)/fx and (v-cy)/fy come from two small tables computed once
// per context with the same float ops as back_project (:94-95), so a back-projected point is
// table[u]*d exactly as in the reference.  Zero normal outside rows 1..H-2 / cols 1..W-2 and
// for the skipped pixels, like the reference's pre-zeroed norm_map (:965).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pixel_normals(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.z;
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x4 >= d.Wp || y >= d.H) return;
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t fo = (size_t)b * d.px_stride;
    const size_t po = fo + (size_t)y * Wp + x4;
    float nx[4] = {0.f, 0.f, 0.f, 0.f}, ny[4] = {0.f, 0.f, 0.f, 0.f}, nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 1 && y <= H - 2 && x4 < W)
    {
        const float4 z4 = *reinterpret_cast<const float4 *>(d.depth + po);
        const float4 zd4 = *reinterpret_cast<const float4 *>(d.depth + po + Wp);
        const float zr = (x4 + 4 < Wp) ? d.depth[po + 4] : 0.f;
        const float4 kx4 = *reinterpret_cast<const float4 *>(d.kx + x4);
        const float kxr = d.kx[x4 + 4];
        const float ky0 = d.ky[y], ky1 = d.ky[y + 1];
        const float z[5] = {z4.x, z4.y, z4.z, z4.w, zr};
        const float zd[4] = {zd4.x, zd4.y, zd4.z, zd4.w};
        const float kx[5] = {kx4.x, kx4.y, kx4.z, kx4.w, kxr};
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int x = x4 + i;
            if (x < 1 || x > W - 2) continue;
            const float mz = z[i], rz = z[i + 1], dz = zd[i];
            if (mz < F_0p1_HI || rz < F_0p1_HI || dz < F_0p1_HI) continue; // (double)z < 0.1 (:688)
            const float mx = kx[i] * mz, my = ky0 * mz;
            const float rx = kx[i + 1] * rz - mx, ry = ky0 * rz - my, rzz = rz - mz;
            const float dx = kx[i] * dz - mx, dy = ky1 * dz - my, dzz = dz - mz;
            float cxn = ry * dzz - rzz * dy;
            float cyn = rzz * dx - rx * dzz;
            float czn = rx * dy - ry * dx;
            const float len = sqrtf(cxn * cxn + cyn * cyn + czn * czn);
            cxn /= len;
            cyn /= len;
            czn /= len;
            const float view = (cxn * mx + cyn * my + czn * mz) / sqrtf(mx * mx + my * my + mz * mz);
            if (view > -F_0p1_HI && view < F_0p1_HI) continue; // |view| < 0.1 in double (:706)
            nx[i] = cxn, ny[i] = cyn, nz[i] = czn;
        }
    }
    *reinterpret_cast<float4 *>(d.nrm + po) = make_float4(nx[0], nx[1], nx[2], nx[3]);
    *reinterpret_cast<float4 *>(d.nrm + d.nrm_plane + po) = make_float4(ny[0], ny[1], ny[2], ny[3]);
    *reinterpret_cast<float4 *>(d.nrm + 2 * d.nrm_plane + po) = make_float4(nz[0], nz[1], nz[2], nz[3]);
}

// -------------------------------------------------------------------------------------------
// K4  seed_plane_fit — calculate_sp_depth_norms_kernel (:792-914) + get_huber_norm (:104-188),
// as two kernels
//
// K4a k_gather_points (warp per seed): scans the 16x16 window (lane = 2*row+half, 8 px/lane), counts
//   valid depths (> 0.05), finds max_dist, classifies inliers |mean_depth-d| < 0.4, reduces the
//   inlier normal / position sums with shuffles and scan-compacts the CENTRED inlier points
//   (get_huber_norm centres them first, :111-126) into three global planes laid out [k][seed].
// K4b k_gauss_newton (thread per seed): the five LM-damped Gauss-Newton steps with no cross-lane
//   traffic at all: 32 neighbouring seeds stream their lists with coalesced loads.  Algebra: for
//   the points whose residual is inside the Huber range, sum 2 r q~ = (sum 2 q~ q~^T) theta, i.e. the
//   in-range part of the Jacobian is H theta.  So H over ALL points is accumulated once (first pass),
//   every pass only evaluates the float residual r exactly as (:133) to classify, and the (rare)
//   out-of-range points contribute corrections: H_R = H_all - sum_out 2 q~ q~^T,
//   J = H_R theta + sum_out +-0.4 q~.  Deviation from the reference: J's in-range part is formed in
//   fp64 from H instead of summing float products 2*r*q -- a ~1e-7 relative perturbation, inside the
//   1e-4 budget of this non-label-affecting stage; classification thresholds,
//   promotions and the projection (:884-912) follow the reference expression by expression.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void solve4_spd(const double *h, const double *j, double *u)
{ // h: 10 unique entries xx xy xz xw yy yz yw zz zw ww of an SPD matrix; solves H u = j
    const double a00 = h[0], a01 = h[1], a02 = h[2], a03 = h[3];
    const double i0 = 1.0 / a00;
    const double l10 = a01 * i0, l20 = a02 * i0, l30 = a03 * i0;
    const double a11 = h[4] - l10 * a01, a12 = h[5] - l10 * a02, a13 = h[6] - l10 * a03;
    const double a22p = h[7] - l20 * a02, a23p = h[8] - l20 * a03, a33p = h[9] - l30 * a03;
    const double i1 = 1.0 / a11;
    const double l21 = a12 * i1, l31 = a13 * i1;
    const double a22 = a22p - l21 * a12, a23 = a23p - l21 * a13, a33q = a33p - l31 * a13;
    const double i2 = 1.0 / a22;
    const double l32 = a23 * i2;
    const double a33 = a33q - l32 * a23;
    const double y0 = j[0];
    const double y1 = j[1] - l10 * y0;
    const double y2 = j[2] - l20 * y0 - l21 * y1;
    const double y3 = j[3] - l30 * y0 - l31 * y1 - l32 * y2;
    u[3] = y3 / a33;
    u[2] = y2 * i2 - l32 * u[3];
    u[1] = y1 * i1 - l21 * u[2] - l31 * u[3];
    u[0] = y0 * i0 - l10 * u[1] - l20 * u[2] - l30 * u[3];
}

#define PF_CAP 228 // >= 15*15 possible members of a superpixel

__global__ void __launch_bounds__(256) k_gather_points(const __grid_constant__ DsmDev d)
{
    // tile[plane][k][seed-in-block]: compacted in shared memory, copied out as full 32-byte sectors
    __shared__ float tile[3 * PF_CAP * 8];
    __shared__ int s_rows;
    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp_x = blockIdx.x * 8 + warp, sp_y = blockIdx.y;
    const int s = sp_y * d.spw + sp_x;
    if (threadIdx.x == 0) s_rows = 0;
    __syncthreads();
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t so = (size_t)b * d.S;
    const bool live = sp_x < d.spw;
    // Slots past the last seed column run the same straight-line code on an empty window: every shuffle
    // below sits in convergent code (a shuffle inside a possibly-divergent branch costs ~10 instructions).
    {
        const size_t fo = (size_t)b * d.px_stride;
        const int32_t *lab = d.labels + fo; // per-frame bases once; 32-bit element offsets below
        const float *dep = d.depth + fo;
        const float *nrx = d.nrm + fo, *nry = nrx + d.nrm_plane, *nrz = nry + d.nrm_plane;
        const float4 sd = d.seed[so + (live ? s : 0)]; // x, y, I, mean_depth (Huber mean after the 3 iterations)
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xq = x0 + 4 * (lane & 3);
        const bool colin = live && xq >= 0 && xq < Wp;
        const float4 k4 = *reinterpret_cast<const float4 *>(d.kx + (colin ? xq : 0));
        const float kxv[4] = {k4.x, k4.y, k4.z, k4.w};
        int4 l4[2];
        float4 z4[2];
        float kyv[2];
        int yy[2];
        unsigned pof[2];
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        {
            const int y = y0 + 8 * ps + (lane >> 2);
            yy[ps] = y;
            const bool in = colin && y >= 0 && y < H;
            const unsigned po = in ? (unsigned)(y * Wp + xq) : 0u;
            pof[ps] = po;
            l4[ps] = *reinterpret_cast<const int4 *>(lab + po); // out-of-window lanes read element 0 and are masked below
            z4[ps] = *reinterpret_cast<const float4 *>(dep + po);
            kyv[ps] = d.ky[in ? y : 0];
            if (!in) l4[ps] = make_int4(-1, -1, -1, -1);
        }
        unsigned inl = 0; // bit 4*ps+k: inlier
        int nvalid = 0;
        float maxd = 0.f, snx = 0.f, sny = 0.f, snz = 0.f, spx = 0.f, spy = 0.f, spz = 0.f;
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        {
            const int lk[4] = {l4[ps].x, l4[ps].y, l4[ps].z, l4[ps].w};
            const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
            unsigned mi = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int x = xq + k;
                if (lk[k] != s || x >= W) continue; // window bounded by the flat index only (:816)
                const float xd = (float)x - sd.x, yd = (float)yy[ps] - sd.y;
                const float dist = xd * xd + yd * yd;
                if (dist > maxd) maxd = dist;
                const float mz = zk[k];
                if (!(mz > F_0p05_LO)) continue; // (double)depth > 0.05 (:827)
                nvalid++;
                const float r = sd.w - mz;
                if (r < F_0p4_HI && r > -F_0p4_HI)
                { // inlier (:849-860)
                    mi |= 1u << k;
                    spx += kxv[k] * mz; // back_project in float (:94-96)
                    spy += kyv[ps] * mz;
                    spz += mz;
                }
            }
            if (mi)
            { // pixel normals of this 4-pixel group: three 16-byte loads instead of up to 12 scalar ones
                const float4 a = *reinterpret_cast<const float4 *>(nrx + pof[ps]);
                const float4 bb = *reinterpret_cast<const float4 *>(nry + pof[ps]);
                const float4 c = *reinterpret_cast<const float4 *>(nrz + pof[ps]);
                const float ax[4] = {a.x, a.y, a.z, a.w}, ay[4] = {bb.x, bb.y, bb.z, bb.w}, az[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((mi >> k) & 1u) snx += ax[k], sny += ay[k], snz += az[k];
                inl |= mi << (4 * ps);
            }
        }
        maxd = warp_max_f(maxd);
        nvalid = __reduce_add_sync(FULL, nvalid);
        snx = warp_sum_f(snx), sny = warp_sum_f(sny), snz = warp_sum_f(snz);
        spx = warp_sum_f(spx), spy = warp_sum_f(spy), spz = warp_sum_f(spz);
        const int c2 = __popc(inl & 0xfu) | (__popc(inl >> 4) << 16);
        int tot2;
        const int ex2 = warp_excl_scan(c2, lane, tot2);
        const int n0 = tot2 & 0xffff, ninl = n0 + (tot2 >> 16);
        const bool ok = live && nvalid >= 16 && !((float)ninl / (float)nvalid < F_0p8_HI); // (:841), (double)ratio < 0.8 (:862)
        float4 P0 = make_float4(0.f, 0.f, 0.f, maxd), P1 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        if (ok)
        {
            const float fn = (float)ninl;
            const float mxs = spx / fn, mys = spy / fn, mzs = spz / fn; // (:117-119)
            const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile) + 4u * warp;
#pragma unroll
            for (int ps = 0; ps < 2; ps++)
            {
                unsigned a = tbase + 32u * (ps == 0 ? (ex2 & 0xffff) : n0 + (ex2 >> 16)); // 32 bytes per list row
                const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((inl >> (4 * ps + k)) & 1u)
                    {
                        sts_f32(a, kxv[k] * zk[k] - mxs); // centred points (:121-126)
                        sts_f32(a + 4u * PF_CAP * 8, kyv[ps] * zk[k] - mys);
                        sts_f32(a + 8u * PF_CAP * 8, zk[k] - mzs);
                        a += 32u;
                    }
            }
            P0 = make_float4(snx, sny, snz, maxd);
            P1 = make_float4(mxs, mys, mzs, __int_as_float(ninl));
            if (lane == 0) atomicMax(&s_rows, ninl);
        }
        if (live && lane == 0)
        {
            d.pfsum[(so + s) * 2] = P0;
            d.pfsum[(so + s) * 2 + 1] = P1;
        }
    }
    __syncthreads();
    const int rows = s_rows;
    const unsigned c = threadIdx.x & 7u;
    if (blockIdx.x * 8 + c < (unsigned)d.spw)
    {
        const size_t plane = (size_t)d.B * PF_CAP * d.Sp;
        float *dx = d.qlist + ((size_t)b * PF_CAP * d.Sp + sp_y * d.spw + blockIdx.x * 8 + c);
        float *dy = dx + plane, *dz = dy + plane;
        const unsigned sp = (unsigned)d.Sp;
#pragma unroll 2
        for (unsigned r = threadIdx.x >> 3; r < (unsigned)rows; r += 32u)
        {
            const unsigned t = r * 8u + c, o = r * sp;
            dx[o] = tile[t];
            dy[o] = tile[PF_CAP * 8 + t];
            dz[o] = tile[2 * PF_CAP * 8 + t];
        }
    }
}

// K4a', EXPERIMENTAL (variant bit 2, off by default): k_gather_points with labels, depth and the
// three normal planes of the block's strip staged through shared memory by TMA 1-D bulk copies, like
// k_gather_depths_tiled.  Everything after the loads is the text of k_gather_points: bit-identical results.
__global__ void __launch_bounds__(256) k_gather_points_tiled(const __grid_constant__ DsmDev d)
{
    // tile[plane][seed-in-block][k]: compacted in shared memory, copied out as full 32-byte sectors
    __shared__ float tile[3 * PF_CAP * 8];
    __shared__ alignas(128) int32_t t_lab[16 * GT_STRIDE];
    __shared__ alignas(128) float t_dep[16 * GT_STRIDE];
    __shared__ alignas(128) float t_nrm[3 * 16 * GT_STRIDE];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ dsm_barrier bar;
    __shared__ int s_rows;
    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp_x = blockIdx.x * 8 + warp, sp_y = blockIdx.y;
    const int s = sp_y * d.spw + sp_x;
    if (threadIdx.x == 0)
    {
        s_rows = 0;
        init(&bar, 1);
        cuda::ptx::fence_proxy_async(cuda::ptx::space_shared); // make the initialised barrier visible to the async proxy
    }
    __syncthreads();
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t so = (size_t)b * d.S;
    const bool live = sp_x < d.spw;
    // the block's 16 x 72 strip of labels, depth and the three normal planes: 80 1-D bulk copies (TMA) of <= 288 B
    // (see k_gather_depths_tiled); the normals arrive in the same phase as the labels instead of by a second,
    // dependent round of global loads
    const int X0 = blockIdx.x * 64 - 4, Y0 = sp_y * DSM_SP - DSM_SP / 2;
    if (warp == 0)
    {
        const size_t fo = (size_t)b * d.px_stride;
        const int ya = Y0 > 0 ? Y0 : 0, yz = (Y0 + 16) < H ? (Y0 + 16) : H;
        const int xs = X0 > 0 ? X0 : 0, xt = (X0 + 72) < Wp ? (X0 + 72) : Wp;
        const unsigned nb4 = (unsigned)(xt - xs) * 4u;
        if (lane == 0) (void)cuda::device::barrier_arrive_tx(bar, 1, (size_t)(yz - ya) * (5u * nb4));
        __syncwarp();
        const int y = Y0 + lane;
        if (lane < 16 && y >= ya && y < yz)
        {
            const size_t ro = fo + (size_t)y * Wp + xs;
            const int to = lane * GT_STRIDE + (xs - X0);
            cuda::device::memcpy_async_tx(t_lab + to, d.labels + ro, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_dep + to, d.depth + ro, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_nrm + to, d.nrm + ro, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_nrm + 16 * GT_STRIDE + to, d.nrm + d.nrm_plane + ro, cuda::aligned_size_t<16>(nb4), bar);
            cuda::device::memcpy_async_tx(t_nrm + 32 * GT_STRIDE + to, d.nrm + 2 * d.nrm_plane + ro, cuda::aligned_size_t<16>(nb4), bar);
        }
    }
    // Slots past the last seed column run the same straight-line code on an empty window: every shuffle
    // below sits in convergent code (a shuffle inside a possibly-divergent branch costs ~10 instructions).
    {
        const float4 sd = d.seed[so + (live ? s : 0)]; // x, y, I, mean_depth (Huber mean after the 3 iterations)
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xq = x0 + 4 * (lane & 3);
        const bool colin = live && xq >= 0 && xq < Wp;
        const float4 k4 = *reinterpret_cast<const float4 *>(d.kx + (colin ? xq : 0));
        const float kxv[4] = {k4.x, k4.y, k4.z, k4.w};
        int4 l4[2];
        float4 z4[2];
        float kyv[2];
        int yy[2];
        unsigned pof[2];
        tile_wait(bar);
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        { // tile row 8*ps + lane/4, tile column 8*warp + 4*(lane&3); positions that were not copied are masked by `in`
            const int r = 8 * ps + (lane >> 2);
            const int y = y0 + r;
            yy[ps] = y;
            const bool in = colin && y >= 0 && y < H;
            const unsigned po = (unsigned)(r * GT_STRIDE + 8 * warp + 4 * (lane & 3));
            pof[ps] = po;
            l4[ps] = *reinterpret_cast<const int4 *>(t_lab + po);
            z4[ps] = *reinterpret_cast<const float4 *>(t_dep + po);
            kyv[ps] = d.ky[in ? y : 0];
            if (!in) l4[ps] = make_int4(-1, -1, -1, -1);
        }
        unsigned inl = 0; // bit 4*ps+k: inlier
        int nvalid = 0;
        float maxd = 0.f, snx = 0.f, sny = 0.f, snz = 0.f, spx = 0.f, spy = 0.f, spz = 0.f;
#pragma unroll
        for (int ps = 0; ps < 2; ps++)
        {
            const int lk[4] = {l4[ps].x, l4[ps].y, l4[ps].z, l4[ps].w};
            const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
            unsigned mi = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int x = xq + k;
                if (lk[k] != s || x >= W) continue; // window bounded by the flat index only (:816)
                const float xd = (float)x - sd.x, yd = (float)yy[ps] - sd.y;
                const float dist = xd * xd + yd * yd;
                if (dist > maxd) maxd = dist;
                const float mz = zk[k];
                if (!(mz > F_0p05_LO)) continue; // (double)depth > 0.05 (:827)
                nvalid++;
                const float r = sd.w - mz;
                if (r < F_0p4_HI && r > -F_0p4_HI)
                { // inlier (:849-860)
                    mi |= 1u << k;
                    spx += kxv[k] * mz; // back_project in float (:94-96)
                    spy += kyv[ps] * mz;
                    spz += mz;
                }
            }
            if (mi)
            { // pixel normals of this 4-pixel group: three 16-byte loads instead of up to 12 scalar ones
                const float4 a = *reinterpret_cast<const float4 *>(t_nrm + pof[ps]);
                const float4 bb = *reinterpret_cast<const float4 *>(t_nrm + 16 * GT_STRIDE + pof[ps]);
                const float4 c = *reinterpret_cast<const float4 *>(t_nrm + 32 * GT_STRIDE + pof[ps]);
                const float ax[4] = {a.x, a.y, a.z, a.w}, ay[4] = {bb.x, bb.y, bb.z, bb.w}, az[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((mi >> k) & 1u) snx += ax[k], sny += ay[k], snz += az[k];
                inl |= mi << (4 * ps);
            }
        }
        maxd = warp_max_f(maxd);
        nvalid = __reduce_add_sync(FULL, nvalid);
        snx = warp_sum_f(snx), sny = warp_sum_f(sny), snz = warp_sum_f(snz);
        spx = warp_sum_f(spx), spy = warp_sum_f(spy), spz = warp_sum_f(spz);
        const int c2 = __popc(inl & 0xfu) | (__popc(inl >> 4) << 16);
        int tot2;
        const int ex2 = warp_excl_scan(c2, lane, tot2);
        const int n0 = tot2 & 0xffff, ninl = n0 + (tot2 >> 16);
        const bool ok = live && nvalid >= 16 && !((float)ninl / (float)nvalid < F_0p8_HI); // (:841), (double)ratio < 0.8 (:862)
        float4 P0 = make_float4(0.f, 0.f, 0.f, maxd), P1 = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        if (ok)
        {
            const float fn = (float)ninl;
            const float mxs = spx / fn, mys = spy / fn, mzs = spz / fn; // (:117-119)
            // [plane][seed][k] instead of [plane][k][seed]: conflict-free compaction stores (see k_gather_depths_tiled;
            // ncu: 13.6 M store bank-conflict wavefronts per launch in the direct-load kernel)
            const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile) + 4u * PF_CAP * warp;
#pragma unroll
            for (int ps = 0; ps < 2; ps++)
            {
                unsigned a = tbase + 4u * (ps == 0 ? (ex2 & 0xffff) : n0 + (ex2 >> 16));
                const float zk[4] = {z4[ps].x, z4[ps].y, z4[ps].z, z4[ps].w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((inl >> (4 * ps + k)) & 1u)
                    {
                        sts_f32(a, kxv[k] * zk[k] - mxs); // centred points (:121-126)
                        sts_f32(a + 4u * PF_CAP * 8, kyv[ps] * zk[k] - mys);
                        sts_f32(a + 8u * PF_CAP * 8, zk[k] - mzs);
                        a += 4u;
                    }
            }
            P0 = make_float4(snx, sny, snz, maxd);
            P1 = make_float4(mxs, mys, mzs, __int_as_float(ninl));
            if (lane == 0) atomicMax(&s_rows, ninl);
        }
        if (live && lane == 0)
        {
            d.pfsum[(so + s) * 2] = P0;
            d.pfsum[(so + s) * 2 + 1] = P1;
        }
    }
    __syncthreads();
    const int rows = s_rows;
    const unsigned c = threadIdx.x & 7u;
    if (blockIdx.x * 8 + c < (unsigned)d.spw)
    {
        const size_t plane = (size_t)d.B * PF_CAP * d.Sp;
        float *dx = d.qlist + ((size_t)b * PF_CAP * d.Sp + sp_y * d.spw + blockIdx.x * 8 + c);
        float *dy = dx + plane, *dz = dy + plane;
        const unsigned sp = (unsigned)d.Sp;
#pragma unroll 2
        for (unsigned r = threadIdx.x >> 3; r < (unsigned)rows; r += 32u)
        {
            const unsigned t = c * PF_CAP + r, o = r * sp; // PF_CAP % 32 == 4: conflict-free
            dx[o] = tile[t];
            dy[o] = tile[PF_CAP * 8 + t];
            dz[o] = tile[2 * PF_CAP * 8 + t];
        }
    }
}

__global__ void __launch_bounds__(128) k_gauss_newton(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.S) return;
    const size_t so = (size_t)b * d.S;
    const float4 sd = d.seed[so + s];
    const float4 P0 = d.pfsum[(so + s) * 2], P1 = d.pfsum[(so + s) * 2 + 1];
    const int n = __float_as_int(P1.w);
    // default record: plane fit rejected -> zero normal / position / view_cos / size (H6-i), Huber mean depth kept
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r1 = make_float4(0.f, 0.f, 0.f, sd.w);
    float4 r2 = make_float4(0.f, sd.z, sd.x, sd.y);
    if (n > 0)
    {
        const float len0 = sqrtf(P0.x * P0.x + P0.y * P0.y + P0.z * P0.z);
        float nx = P0.x / len0, ny = P0.y / len0, nz = P0.z / len0, nb = 0.f; // len0 == 0 -> NaN, propagated (H6-iii)
        const float mxs = P1.x, mys = P1.y, mzs = P1.z;
        const size_t plane = (size_t)d.B * PF_CAP * d.Sp, st = (size_t)d.Sp;
        const float *qx = d.qlist + (size_t)b * PF_CAP * d.Sp + s;
        const float *qy = qx + plane, *qz = qy + plane;
        double hall[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // xx xy xz xw yy yz yw zz zw ww over ALL points
        // Pass skipping: a pass over the points is only needed to find out which of them fall outside the
        // Huber range.  With rmax >= max|r_i| of the last evaluated parameters, qmax = max|q_i| and the step
        // (dn, db) just taken, |r_i(new)| <= rmax + qmax*|dn| + |db|.  If that bound (plus a rounding
        // slack far above the float error of evaluating r) stays below the range, every point is
        // provably in range, so H_R = H_all and J = H_all*theta without touching memory.  Results are
        // bit-identical to evaluating the pass.
        float rmax = 0.f, qmax2 = 0.f;
        bool need_pass = true;
        for (int gn = 0; gn < 5; gn++)
        {
            double ho[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // same, over the points outside the Huber range
            double jo[4] = {0, 0, 0, 0};
            if (need_pass)
            {
                float rm = 0.f;
                bool rnan = false;
                auto point = [&](float ax, float ay, float az)
                {
                    const float r = ax * nx + ay * ny + az * nz + nb; // (:133)
                    const bool inr = r < F_0p4_HI && r > -F_0p4_HI;  // (:134)
                    rm = fmaxf(rm, fabsf(r));
                    rnan |= !(r == r);
                    if (gn == 0 || !inr)
                    {
                        const double t0 = (double)(2 * ax * ax), t1 = (double)(2 * ax * ay), t2 = (double)(2 * ax * az), t3 = (double)(2 * ax);
                        const double t4 = (double)(2 * ay * ay), t5 = (double)(2 * ay * az), t6 = (double)(2 * ay);
                        const double t7 = (double)(2 * az * az), t8 = (double)(2 * az);
                        if (gn == 0)
                        {
                            hall[0] += t0, hall[1] += t1, hall[2] += t2, hall[3] += t3, hall[4] += t4;
                            hall[5] += t5, hall[6] += t6, hall[7] += t7, hall[8] += t8, hall[9] += 2;
                            qmax2 = fmaxf(qmax2, ax * ax + ay * ay + az * az);
                        }
                        if (!inr)
                        {
                            ho[0] += t0, ho[1] += t1, ho[2] += t2, ho[3] += t3, ho[4] += t4;
                            ho[5] += t5, ho[6] += t6, ho[7] += t7, ho[8] += t8, ho[9] += 2;
                            if (r >= F_0p4_HI)
                            { // (double)r >= 0.4 (:157-163)
                                jo[0] += HUBER_RANGE * (double)ax, jo[1] += HUBER_RANGE * (double)ay;
                                jo[2] += HUBER_RANGE * (double)az, jo[3] += HUBER_RANGE;
                            }
                            else if (r <= -F_0p4_HI)
                            { // (double)r <= -0.4 (:164-170)
                                jo[0] += -1 * HUBER_RANGE * (double)ax, jo[1] += -1 * HUBER_RANGE * (double)ay;
                                jo[2] += -1 * HUBER_RANGE * (double)az, jo[3] += -1 * HUBER_RANGE;
                            }
                        }
                    }
                };
                int k = 0;
                for (; k + 4 <= n; k += 4)
                { // four points in flight: 12 coalesced loads issued before the first is consumed
                    const float a0 = qx[k * st], a1 = qx[(k + 1) * st], a2 = qx[(k + 2) * st], a3 = qx[(k + 3) * st];
                    const float b0 = qy[k * st], b1 = qy[(k + 1) * st], b2 = qy[(k + 2) * st], b3 = qy[(k + 3) * st];
                    const float c0 = qz[k * st], c1 = qz[(k + 1) * st], c2 = qz[(k + 2) * st], c3 = qz[(k + 3) * st];
                    point(a0, b0, c0);
                    point(a1, b1, c1);
                    point(a2, b2, c2);
                    point(a3, b3, c3);
                }
                for (; k < n; k++) point(qx[k * st], qy[k * st], qz[k * st]);
                rmax = rnan ? __int_as_float(0x7f800000) : rm; // a NaN residual forces every later pass
            }
            double hh[10], jj[4];
#pragma unroll
            for (int i = 0; i < 10; i++) hh[i] = hall[i] - ho[i];
            const double tx = (double)nx, ty = (double)ny, tz = (double)nz, tb = (double)nb;
            jj[0] = ((hh[0] * tx + hh[1] * ty) + hh[2] * tz) + hh[3] * tb + jo[0];
            jj[1] = ((hh[1] * tx + hh[4] * ty) + hh[5] * tz) + hh[6] * tb + jo[1];
            jj[2] = ((hh[2] * tx + hh[5] * ty) + hh[7] * tz) + hh[8] * tb + jo[2];
            jj[3] = ((hh[3] * tx + hh[6] * ty) + hh[8] * tz) + hh[9] * tb + jo[3];
            hh[0] += 5, hh[4] += 5, hh[7] += 5, hh[9] += 5; // LM damping (:172-175)
            double u[4];
            solve4_spd(hh, jj, u);
            const float ox = nx, oy = ny, oz = nz, ob = nb;
            nx = (float)((double)nx - u[0]);
            ny = (float)((double)ny - u[1]);
            nz = (float)((double)nz - u[2]);
            nb = (float)((double)nb - u[3]);
            // can the next pass be skipped?
            const float dx = nx - ox, dy = ny - oy, dz = nz - oz;
            const float bound = rmax + sqrtf(qmax2) * sqrtf(dx * dx + dy * dy + dz * dz) * 1.0001f + fabsf(nb - ob) + 1e-3f;
            need_pass = !(bound < 0.39f); // NaN-safe: any NaN keeps evaluating
            rmax = bound;
        }
        nb = nb - (nx * mxs + ny * mys + nz * mzs);
        const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nl;
        ny /= nl;
        nz /= nl;
        nb /= nl;
        // centre of the superpixel projected onto the fitted plane (:884-895)
        const float axf = (sd.x - d.cx) / d.fx * sd.w;
        const float ayf = (sd.y - d.cy) / d.fy * sd.w;
        double ax = (double)axf, ay = (double)ayf, az = (double)sd.w;
        const float kk = (float)(-1 * (ax * (double)nx + ay * (double)ny + az * (double)nz) - (double)nb);
        ax += (double)(kk * nx);
        ay += (double)(kk * ny);
        az += (double)(kk * nz);
        const float mean_depth = (float)az;
        float view_cos = (float)(-1.0 * ((double)nx * ax + (double)ny * ay + (double)nz * az) / sqrt(ax * ax + ay * ay + az * az));
        if (view_cos < 0)
        {
            view_cos = -view_cos;
            nx = -nx;
            ny = -ny;
            nz = -nz;
        }
        r0 = make_float4(nx, ny, nz, view_cos);
        r1 = make_float4((float)ax, (float)ay, (float)az, mean_depth);
        r2.x = sqrtf(P0.w);
    }
    float4 *pl = d.plane + (so + s) * 3;
    pl[0] = r0;
    pl[1] = r1;
    pl[2] = r2;
}

// K4b', EXPERIMENTAL (variant bit 3, off by default): k_gauss_newton with every pass over the point
// list streamed through thread-private shared-memory columns by double-buffered 4-byte cp.async instead of
// 4-point register batches whose load latency is exposed once per batch.  Same point order, same arithmetic.
#define GS_CH 16
__global__ void __launch_bounds__(128) k_gauss_newton_staged(const __grid_constant__ DsmDev d)
{
    __shared__ float stage[2 * 3 * GS_CH * 128]; // [buffer][plane][row][thread]: 48 KB
    float *mybuf = stage + threadIdx.x;
    const int b = d.frame0 + blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.S) return;
    const size_t so = (size_t)b * d.S;
    const float4 sd = d.seed[so + s];
    const float4 P0 = d.pfsum[(so + s) * 2], P1 = d.pfsum[(so + s) * 2 + 1];
    const int n = __float_as_int(P1.w);
    // default record: plane fit rejected -> zero normal / position / view_cos / size (H6-i), Huber mean depth kept
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r1 = make_float4(0.f, 0.f, 0.f, sd.w);
    float4 r2 = make_float4(0.f, sd.z, sd.x, sd.y);
    if (n > 0)
    {
        const float len0 = sqrtf(P0.x * P0.x + P0.y * P0.y + P0.z * P0.z);
        float nx = P0.x / len0, ny = P0.y / len0, nz = P0.z / len0, nb = 0.f; // len0 == 0 -> NaN, propagated (H6-iii)
        const float mxs = P1.x, mys = P1.y, mzs = P1.z;
        const size_t plane = (size_t)d.B * PF_CAP * d.Sp, st = (size_t)d.Sp;
        const float *qx = d.qlist + (size_t)b * PF_CAP * d.Sp + s;
        const float *qy = qx + plane, *qz = qy + plane;
        double hall[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // xx xy xz xw yy yz yw zz zw ww over ALL points
        // Pass skipping: a pass over the points is only needed to find out which of them fall outside the
        // Huber range.  With rmax >= max|r_i| of the last evaluated parameters, qmax = max|q_i| and the step
        // (dn, db) just taken, |r_i(new)| <= rmax + qmax*|dn| + |db|.  If that bound (plus a rounding
        // slack far above the float error of evaluating r) stays below the range, every point is
        // provably in range, so H_R = H_all and J = H_all*theta without touching memory.  Results are
        // bit-identical to evaluating the pass.
        float rmax = 0.f, qmax2 = 0.f;
        bool need_pass = true;
        for (int gn = 0; gn < 5; gn++)
        {
            double ho[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // same, over the points outside the Huber range
            double jo[4] = {0, 0, 0, 0};
            if (need_pass)
            {
                float rm = 0.f;
                bool rnan = false;
                auto point = [&](float ax, float ay, float az)
                {
                    const float r = ax * nx + ay * ny + az * nz + nb; // (:133)
                    const bool inr = r < F_0p4_HI && r > -F_0p4_HI;  // (:134)
                    rm = fmaxf(rm, fabsf(r));
                    rnan |= !(r == r);
                    if (gn == 0 || !inr)
                    {
                        const double t0 = (double)(2 * ax * ax), t1 = (double)(2 * ax * ay), t2 = (double)(2 * ax * az), t3 = (double)(2 * ax);
                        const double t4 = (double)(2 * ay * ay), t5 = (double)(2 * ay * az), t6 = (double)(2 * ay);
                        const double t7 = (double)(2 * az * az), t8 = (double)(2 * az);
                        if (gn == 0)
                        {
                            hall[0] += t0, hall[1] += t1, hall[2] += t2, hall[3] += t3, hall[4] += t4;
                            hall[5] += t5, hall[6] += t6, hall[7] += t7, hall[8] += t8, hall[9] += 2;
                            qmax2 = fmaxf(qmax2, ax * ax + ay * ay + az * az);
                        }
                        if (!inr)
                        {
                            ho[0] += t0, ho[1] += t1, ho[2] += t2, ho[3] += t3, ho[4] += t4;
                            ho[5] += t5, ho[6] += t6, ho[7] += t7, ho[8] += t8, ho[9] += 2;
                            if (r >= F_0p4_HI)
                            { // (double)r >= 0.4 (:157-163)
                                jo[0] += HUBER_RANGE * (double)ax, jo[1] += HUBER_RANGE * (double)ay;
                                jo[2] += HUBER_RANGE * (double)az, jo[3] += HUBER_RANGE;
                            }
                            else if (r <= -F_0p4_HI)
                            { // (double)r <= -0.4 (:164-170)
                                jo[0] += -1 * HUBER_RANGE * (double)ax, jo[1] += -1 * HUBER_RANGE * (double)ay;
                                jo[2] += -1 * HUBER_RANGE * (double)az, jo[3] += -1 * HUBER_RANGE;
                            }
                        }
                    }
                };
                // the list streams through this thread's private shared-memory columns in chunks of GS_CH points,
                // double-buffered with cp.async: chunk c+1 is in flight while chunk c is consumed (same point order)
                const int nch = (n + GS_CH - 1) / GS_CH;
                auto issue = [&](int c)
                {
                    if (c < nch)
                    {
                        const int k0 = c * GS_CH, k1 = (k0 + GS_CH) < n ? (k0 + GS_CH) : n;
                        float *dst = mybuf + (c & 1) * (3 * GS_CH * 128);
                        for (int k = k0; k < k1; k++, dst += 128)
                        {
                            __pipeline_memcpy_async(dst, qx + k * st, 4);
                            __pipeline_memcpy_async(dst + GS_CH * 128, qy + k * st, 4);
                            __pipeline_memcpy_async(dst + 2 * GS_CH * 128, qz + k * st, 4);
                        }
                    }
                    __pipeline_commit(); // (possibly empty) group, so that "all but the newest group" below is chunk c
                };
                issue(0);
                for (int c = 0; c < nch; c++)
                {
                    issue(c + 1);
                    __pipeline_wait_prior(1);
                    const int cnt = (n - c * GS_CH) < GS_CH ? (n - c * GS_CH) : GS_CH;
                    const float *src = mybuf + (c & 1) * (3 * GS_CH * 128);
                    for (int r = 0; r < cnt; r++) point(src[r * 128], src[(GS_CH + r) * 128], src[(2 * GS_CH + r) * 128]);
                }
                __pipeline_wait_prior(0);
                rmax = rnan ? __int_as_float(0x7f800000) : rm; // a NaN residual forces every later pass
            }
            double hh[10], jj[4];
#pragma unroll
            for (int i = 0; i < 10; i++) hh[i] = hall[i] - ho[i];
            const double tx = (double)nx, ty = (double)ny, tz = (double)nz, tb = (double)nb;
            jj[0] = ((hh[0] * tx + hh[1] * ty) + hh[2] * tz) + hh[3] * tb + jo[0];
            jj[1] = ((hh[1] * tx + hh[4] * ty) + hh[5] * tz) + hh[6] * tb + jo[1];
            jj[2] = ((hh[2] * tx + hh[5] * ty) + hh[7] * tz) + hh[8] * tb + jo[2];
            jj[3] = ((hh[3] * tx + hh[6] * ty) + hh[8] * tz) + hh[9] * tb + jo[3];
            hh[0] += 5, hh[4] += 5, hh[7] += 5, hh[9] += 5; // LM damping (:172-175)
            double u[4];
            solve4_spd(hh, jj, u);
            const float ox = nx, oy = ny, oz = nz, ob = nb;
            nx = (float)((double)nx - u[0]);
            ny = (float)((double)ny - u[1]);
            nz = (float)((double)nz - u[2]);
            nb = (float)((double)nb - u[3]);
            // can the next pass be skipped?
            const float dx = nx - ox, dy = ny - oy, dz = nz - oz;
            const float bound = rmax + sqrtf(qmax2) * sqrtf(dx * dx + dy * dy + dz * dz) * 1.0001f + fabsf(nb - ob) + 1e-3f;
            need_pass = !(bound < 0.39f); // NaN-safe: any NaN keeps evaluating
            rmax = bound;
        }
        nb = nb - (nx * mxs + ny * mys + nz * mzs);
        const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nl;
        ny /= nl;
        nz /= nl;
        nb /= nl;
        // centre of the superpixel projected onto the fitted plane (:884-895)
        const float axf = (sd.x - d.cx) / d.fx * sd.w;
        const float ayf = (sd.y - d.cy) / d.fy * sd.w;
        double ax = (double)axf, ay = (double)ayf, az = (double)sd.w;
        const float kk = (float)(-1 * (ax * (double)nx + ay * (double)ny + az * (double)nz) - (double)nb);
        ax += (double)(kk * nx);
        ay += (double)(kk * ny);
        az += (double)(kk * nz);
        const float mean_depth = (float)az;
        float view_cos = (float)(-1.0 * ((double)nx * ax + (double)ny * ay + (double)nz * az) / sqrt(ax * ax + ay * ay + az * az));
        if (view_cos < 0)
        {
            view_cos = -view_cos;
            nx = -nx;
            ny = -ny;
            nz = -nz;
        }
        r0 = make_float4(nx, ny, nz, view_cos);
        r1 = make_float4((float)ax, (float)ay, (float)az, mean_depth);
        r2.x = sqrtf(P0.w);
    }
    float4 *pl = d.plane + (so + s) * 3;
    pl[0] = r0;
    pl[1] = r1;
    pl[2] = r2;
}

// k_gauss_newton for SMALL batches (single-frame stream): 8 lanes per seed.  Every lane takes every 8th
// point, the fp64 sums are combined with three width-8 shuffle steps, all lanes of the group then hold
// the same normal equations and solve them redundantly.  Same algebra, pass skipping and thresholds as
// k_gauss_newton; only the summation order of the fp64 accumulators differs (~1e-16 relative).
__device__ __forceinline__ double group8_sum(double v, unsigned gmask)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, 8);
    return v;
}

__global__ void __launch_bounds__(128) k_gauss_newton_small(const __grid_constant__ DsmDev d)
{
    const int b = d.frame0 + blockIdx.y;
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const unsigned gmask = 0xffu << (lane & 24);
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (s >= d.S) return; // whole 8-lane groups leave together
    const size_t so = (size_t)b * d.S;
    const float4 sd = d.seed[so + s];
    const float4 P0 = d.pfsum[(so + s) * 2], P1 = d.pfsum[(so + s) * 2 + 1];
    const int n = __float_as_int(P1.w);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r1 = make_float4(0.f, 0.f, 0.f, sd.w);
    float4 r2 = make_float4(0.f, sd.z, sd.x, sd.y);
    if (n > 0)
    {
        const float len0 = sqrtf(P0.x * P0.x + P0.y * P0.y + P0.z * P0.z);
        float nx = P0.x / len0, ny = P0.y / len0, nz = P0.z / len0, nb = 0.f; // len0 == 0 -> NaN, propagated (H6-iii)
        const float mxs = P1.x, mys = P1.y, mzs = P1.z;
        const size_t plane = (size_t)d.B * PF_CAP * d.Sp, st = (size_t)d.Sp;
        const float *qx = d.qlist + (size_t)b * PF_CAP * d.Sp + s;
        const float *qy = qx + plane, *qz = qy + plane;
        double hall[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        float rmax = 0.f, qmax2 = 0.f;
        bool need_pass = true;
        for (int gn = 0; gn < 5; gn++)
        {
            double ho[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            double jo[4] = {0, 0, 0, 0};
            if (need_pass)
            {
                float rm = 0.f;
                int flags = 0; // bit0: some residual NaN, bit1: some residual outside the Huber range
                for (int k = gl; k < n; k += 8)
                {
                    const float ax = qx[k * st], ay = qy[k * st], az = qz[k * st];
                    const float r = ax * nx + ay * ny + az * nz + nb; // (:133)
                    const bool inr = r < F_0p4_HI && r > -F_0p4_HI;  // (:134)
                    rm = fmaxf(rm, fabsf(r));
                    flags |= (r == r ? 0 : 1) | (inr ? 0 : 2);
                    if (gn == 0 || !inr)
                    {
                        const double t0 = (double)(2 * ax * ax), t1 = (double)(2 * ax * ay), t2 = (double)(2 * ax * az), t3 = (double)(2 * ax);
                        const double t4 = (double)(2 * ay * ay), t5 = (double)(2 * ay * az), t6 = (double)(2 * ay);
                        const double t7 = (double)(2 * az * az), t8 = (double)(2 * az);
                        if (gn == 0)
                        {
                            hall[0] += t0, hall[1] += t1, hall[2] += t2, hall[3] += t3, hall[4] += t4;
                            hall[5] += t5, hall[6] += t6, hall[7] += t7, hall[8] += t8, hall[9] += 2;
                            qmax2 = fmaxf(qmax2, ax * ax + ay * ay + az * az);
                        }
                        if (!inr)
                        {
                            ho[0] += t0, ho[1] += t1, ho[2] += t2, ho[3] += t3, ho[4] += t4;
                            ho[5] += t5, ho[6] += t6, ho[7] += t7, ho[8] += t8, ho[9] += 2;
                            if (r >= F_0p4_HI)
                            {
                                jo[0] += HUBER_RANGE * (double)ax, jo[1] += HUBER_RANGE * (double)ay;
                                jo[2] += HUBER_RANGE * (double)az, jo[3] += HUBER_RANGE;
                            }
                            else if (r <= -F_0p4_HI)
                            {
                                jo[0] += -1 * HUBER_RANGE * (double)ax, jo[1] += -1 * HUBER_RANGE * (double)ay;
                                jo[2] += -1 * HUBER_RANGE * (double)az, jo[3] += -1 * HUBER_RANGE;
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1)
                {
                    rm = fmaxf(rm, __shfl_xor_sync(gmask, rm, o, 8));
                    flags |= __shfl_xor_sync(gmask, flags, o, 8);
                }
                if (gn == 0)
                {
#pragma unroll
                    for (int i = 0; i < 10; i++) hall[i] = group8_sum(hall[i], gmask);
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) qmax2 = fmaxf(qmax2, __shfl_xor_sync(gmask, qmax2, o, 8));
                }
                if (flags & 2)
                { // uniform within the group after the reduction
#pragma unroll
                    for (int i = 0; i < 10; i++) ho[i] = group8_sum(ho[i], gmask);
#pragma unroll
                    for (int i = 0; i < 4; i++) jo[i] = group8_sum(jo[i], gmask);
                }
                rmax = (flags & 1) ? __int_as_float(0x7f800000) : rm;
            }
            double hh[10], jj[4];
#pragma unroll
            for (int i = 0; i < 10; i++) hh[i] = hall[i] - ho[i];
            const double tx = (double)nx, ty = (double)ny, tz = (double)nz, tb = (double)nb;
            jj[0] = ((hh[0] * tx + hh[1] * ty) + hh[2] * tz) + hh[3] * tb + jo[0];
            jj[1] = ((hh[1] * tx + hh[4] * ty) + hh[5] * tz) + hh[6] * tb + jo[1];
            jj[2] = ((hh[2] * tx + hh[5] * ty) + hh[7] * tz) + hh[8] * tb + jo[2];
            jj[3] = ((hh[3] * tx + hh[6] * ty) + hh[8] * tz) + hh[9] * tb + jo[3];
            hh[0] += 5, hh[4] += 5, hh[7] += 5, hh[9] += 5; // LM damping (:172-175)
            double u[4];
            solve4_spd(hh, jj, u);
            const float ox = nx, oy = ny, oz = nz, ob = nb;
            nx = (float)((double)nx - u[0]);
            ny = (float)((double)ny - u[1]);
            nz = (float)((double)nz - u[2]);
            nb = (float)((double)nb - u[3]);
            const float dx = nx - ox, dy = ny - oy, dz = nz - oz;
            const float bound = rmax + sqrtf(qmax2) * sqrtf(dx * dx + dy * dy + dz * dz) * 1.0001f + fabsf(nb - ob) + 1e-3f;
            need_pass = !(bound < 0.39f);
            rmax = bound;
        }
        nb = nb - (nx * mxs + ny * mys + nz * mzs);
        const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nl;
        ny /= nl;
        nz /= nl;
        nb /= nl;
        const float axf = (sd.x - d.cx) / d.fx * sd.w;
        const float ayf = (sd.y - d.cy) / d.fy * sd.w;
        double ax = (double)axf, ay = (double)ayf, az = (double)sd.w;
        const float kk = (float)(-1 * (ax * (double)nx + ay * (double)ny + az * (double)nz) - (double)nb);
        ax += (double)(kk * nx);
        ay += (double)(kk * ny);
        az += (double)(kk * nz);
        const float mean_depth = (float)az;
        float view_cos = (float)(-1.0 * ((double)nx * ax + (double)ny * ay + (double)nz * az) / sqrt(ax * ax + ay * ay + az * az));
        if (view_cos < 0)
        {
            view_cos = -view_cos;
            nx = -nx;
            ny = -ny;
            nz = -nz;
        }
        r0 = make_float4(nx, ny, nz, view_cos);
        r1 = make_float4((float)ax, (float)ay, (float)az, mean_depth);
        r2.x = sqrtf(P0.w);
    }
    if (gl == 0)
    {
        float4 *pl = d.plane + (so + s) * 3;
        pl[0] = r0;
        pl[1] = r1;
        pl[2] = r2;
    }
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
        while (!cuda::ptx::mbarrier_try_wait_parity(cuda::device::barrier_native_handle(*bar), 0)) {}
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
            float tol = (float)((double)(pc[2] * pc[2]) / (BASELINE * (double)d.camera_f) * DISPARITY_ERROR);
            tol = (double)tol < MIN_TOLERATE_DIFF ? (float)MIN_TOLERATE_DIFF : tol;
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
// K6  surfel_init — initialize_surfels (:315-361).  One CTA per frame; ordered compaction by
// ballot + block scan so new_surfels come out in seed-index order exactly like the reference's
// serial push_back loop.
// -------------------------------------------------------------------------------------------
#define INIT_PER_THREAD 8
__global__ void __launch_bounds__(1024) k_init_surfels(const __grid_constant__ DsmDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_total;
    __shared__ float s_pose[16];
    const int b = d.frame0 + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t so = (size_t)b * d.S;
    if (threadIdx.x < 16) s_pose[threadIdx.x] = d.pose[b * 16 + threadIdx.x];
    const int ref = d.refidx[b];
    dsm_surfel_t *out = d.newsurf + so;
    int running = 0;
    // a thread owns INIT_PER_THREAD CONSECUTIVE seeds, so one block scan per 8192 seeds gives every
    // thread the seed-index-ordered output position of its first emitted surfel
    for (int base = 0; base < d.S; base += 1024 * INIT_PER_THREAD)
    {
        const int s0 = base + threadIdx.x * INIT_PER_THREAD;
        unsigned emit = 0;
#pragma unroll
        for (int j = 0; j < INIT_PER_THREAD; j++)
        {
            const int s = s0 + j;
            if (s < d.S)
            {
                const float4 *pl = d.plane + (so + s) * 3;
                const float4 r0 = pl[0];
                const float md = pl[1].w;
                const bool e = !(md == 0) && !d.fused[so + s] && !((double)r0.w < MAX_ANGLE_COS) &&
                               !(r0.x == 0 && r0.y == 0 && r0.z == 0);
                emit |= (e ? 1u : 0u) << j;
            }
        }
        const int cnt = __popc(emit);
        int wtot;
        const int wex = warp_excl_scan(cnt, lane, wtot);
        if (lane == 31) s_warp[warp] = wtot;
        __syncthreads();
        if (warp == 0)
        {
            int t;
            const int e = warp_excl_scan(s_warp[lane], lane, t);
            s_warp[lane] = e;
            if (lane == 0) s_total = t;
        }
        __syncthreads();
        int pos = running + s_warp[warp] + wex;
#pragma unroll
        for (int j = 0; j < INIT_PER_THREAD; j++)
            if ((emit >> j) & 1u)
            {
                const float4 *pl = d.plane + (so + s0 + j) * 3; // second read hits L1/L2
                const float4 r0 = pl[0], r1 = pl[1], r2 = pl[2];
                float pw[4], nw[3];
                mat4_mul(s_pose, r1.x, r1.y, r1.z, 1.0f, pw);
                mat3_mul(s_pose, r0.x, r0.y, r0.z, nw);
                dsm_surfel_t e;
                e.px = pw[0], e.py = pw[1], e.pz = pw[2];
                e.nx = nw[0], e.ny = nw[1], e.nz = nw[2];
                e.size = (float)((double)r2.x * fabs((double)(r1.w / (d.camera_f * r0.w))));
                e.color = r2.y;
                e.weight = get_weight(r1.w);
                e.update_times = 1;
                e.last_update = ref;
                out[pos++] = e;
            }
        running += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) d.nnew[b] = running;
}

// K6', EXPERIMENTAL (variant bit 4, off by default): initialize_surfels with several CTAs per frame.
// k_init_surfels keeps the reference's seed-index output order with ONE 1024-thread CTA per frame (32 CTAs for a
// 32-frame batch on 148 SMs, and a serial ~S/8192-round loop on a single frame's critical path).  Here CTA j owns
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
__global__ void __launch_bounds__(1024) k_init_surfels_mb(const __grid_constant__ DsmDev d)
{
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
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int i = begin + blockIdx.x * 256 + threadIdx.x;
    const bool live = i < end && pool_pred(d.pool[i], mode, key);
    const int c = __syncthreads_count(live);
    if (threadIdx.x == 0) blkcnt[blockIdx.x] = (begin + blockIdx.x * 256 < end) ? c : 0;
}

__global__ void __launch_bounds__(1024) k_pool_scan(const __grid_constant__ DsmDev d, int b, const int *blkcnt, int *blkofs, int *newofs)
{
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

__global__ void __launch_bounds__(256) k_pool_append(const __grid_constant__ DsmDev d, int b, const int *newofs, dsm_surfel_t *dst)
{
    const int n = d.nnew[b];
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < n) dst[newofs[0] + j] = d.newsurf[(size_t)b * d.S + j];
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
    if (d.variants & DSM_VARIANT_SEED_INIT_WIDE)
    {
        dim3 g((d.S + 255) / 256, nb);
        k_seed_init_wide<<<g, 256, 0, s>>>(d);
        return;
    }
    dim3 grid((d.S + 255) / 256, nb);
    k_seed_init<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_assign(const DsmDev &d, int nb, bool first, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.W + 255) / 256, (d.H + 3) / 4, nb);
    if (d.variants & DSM_VARIANT_ASSIGN_FEWER_CVT)
    {
        if (first)
            k_assign_x<true><<<grid, block, 0, s>>>(d);
        else
            k_assign_x<false><<<grid, block, 0, s>>>(d);
        return;
    }
    if (first)
        k_assign<true><<<grid, block, 0, s>>>(d);
    else
        k_assign<false><<<grid, block, 0, s>>>(d);
}
void dsm_launch_relax(const DsmDev &d, int nb, cudaStream_t s) { k_relax<<<nb, 1024, 0, s>>>(d); }
void dsm_launch_gather_depths(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.spw + 7) / 8, d.sph, nb);
    if (d.variants & DSM_VARIANT_GATHER_TILED)
    {
        k_gather_depths_tiled<<<grid, 256, 0, s>>>(d);
        return;
    }
    k_gather_depths<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_newton(const DsmDev &d, int nb, cudaStream_t s)
{
    if (d.variants & DSM_VARIANT_NEWTON_STAGED)
    {
        dim3 grid((d.S + NS_THREADS - 1) / NS_THREADS, nb);
        k_newton_staged<<<grid, NS_THREADS, 0, s>>>(d);
        return;
    }
    dim3 grid((d.S + 127) / 128, nb);
    k_newton<<<grid, 128, 0, s>>>(d);
}
void dsm_launch_pixel_normals(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.Wp + 255) / 256, (d.H + 3) / 4, nb);
    k_pixel_normals<<<grid, block, 0, s>>>(d);
}
void dsm_launch_gather_points(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.spw + 7) / 8, d.sph, nb);
    if (d.variants & DSM_VARIANT_POINTS_TILED)
    {
        k_gather_points_tiled<<<grid, 256, 0, s>>>(d);
        return;
    }
    k_gather_points<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_gauss_newton(const DsmDev &d, int nb, cudaStream_t s)
{
    if ((long)nb * d.S <= 20000)
    { // single-frame streams: 8 lanes per seed
        dim3 grid((d.S * 8 + 127) / 128, nb);
        k_gauss_newton_small<<<grid, 128, 0, s>>>(d);
        return;
    }
    dim3 grid((d.S + 127) / 128, nb);
    if (d.variants & DSM_VARIANT_GN_STAGED)
    {
        k_gauss_newton_staged<<<grid, 128, 0, s>>>(d);
        return;
    }
    k_gauss_newton<<<grid, 128, 0, s>>>(d);
}
void dsm_launch_fuse(const DsmDev &d, int nb, cudaStream_t s)
{
    if (d.max_pool_per_frame <= 0) return;
    dim3 grid((d.max_pool_per_frame + FUSE_BLOCK - 1) / FUSE_BLOCK, nb);
    k_fuse<<<grid, FUSE_BLOCK, 0, s>>>(d);
}
void dsm_launch_init_surfels(const DsmDev &d, int nb, cudaStream_t s)
{
    if (!(d.variants & DSM_VARIANT_LEGACY) || (d.variants & DSM_VARIANT_INIT_MULTIBLOCK))
    { // several CTAs per frame (verified byte-identical on the B200, round 2); the one-CTA kernel stays with the round-1 schedule
        dim3 grid((d.S + 1023) / 1024, nb);
        k_init_surfels_mb<<<grid, 1024, 0, s>>>(d);
        return;
    }
    k_init_surfels<<<nb, 1024, 0, s>>>(d);
}
void dsm_launch_seeds_export(const DsmDev &d, int frame, dsm_seed_t *out_dev, int raw_md, cudaStream_t s)
{
    k_seeds_export<<<(d.S + 255) / 256, 256, 0, s>>>(d, frame, out_dev, raw_md);
}

void dsm_launch_pool_compact(const DsmDev &d, int frame, int upper, int *blkcnt, int *blkofs, int *newofs, dsm_surfel_t *dst, cudaStream_t s)
{
    const int nblk = (upper + 255) / 256;
    if (nblk > 0) k_pool_count<<<nblk, 256, 0, s>>>(d, frame, blkcnt, 0, 0);
    k_pool_scan<<<1, 1024, 0, s>>>(d, frame, blkcnt, blkofs, newofs);
    if (nblk > 0) k_pool_scatter<<<nblk, 256, 0, s>>>(d, frame, blkofs, dst, 0, 0);
    k_pool_append<<<(d.S + 255) / 256, 256, 0, s>>>(d, frame, newofs, dst);
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
