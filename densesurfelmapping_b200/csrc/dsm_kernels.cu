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
#include "dsm_device.cuh"
#include <climits>

#define HUBER_RANGE 0.4       // fusion_functions.h:13
#define MAX_ANGLE_COS 0.1     // fusion_functions.h:11
#define BASELINE 0.5          // fusion_functions.h:14
#define DISPARITY_ERROR 4.0   // fusion_functions.h:15
#define MIN_TOLERATE_DIFF 0.1 // fusion_functions.h:16

#define FULL 0xffffffffu

// -------------------------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ int chunk_of(int s, int S)
{ // which of the reference's 10 static chunks seed s falls in (:471-475)
    int step = S / DSM_THREAD_NUM;
    if (step == 0) return DSM_THREAD_NUM - 1;
    int c = s / step;
    return c > DSM_THREAD_NUM - 1 ? DSM_THREAD_NUM - 1 : c;
}

__device__ __forceinline__ float warp_sum_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int n = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += n;
    }
    total = __shfl_sync(FULL, incl, 31);
    return incl - v;
}

// -------------------------------------------------------------------------------------------
// K0  seed_init   — initialize_seeds_kernel (:577-629) + the per-frame clears (:963-965)
// one thread per seed
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_seed_init(const __grid_constant__ DsmDev d)
{
    const int b = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < 16)
    {
        d.abortc[b * 16 + threadIdx.x] = INT_MAX;
        if (threadIdx.x == 0)
        {
            d.nlist[b] = 0;
            d.nnew[b] = 0;
        }
    }
    if (s >= d.S) return;
    const int W = d.W, H = d.H, Wp = d.Wp;
    const uint8_t *gray = d.gray + (size_t)b * d.px_stride;
    const float *depth = d.depth + (size_t)b * d.px_stride;
    const int sp_x = s % d.spw, sp_y = s / d.spw;
    int ix = sp_x * DSM_SP + DSM_SP / 2, iy = sp_y * DSM_SP + DSM_SP / 2;
    ix = ix < W - 1 ? ix : W - 1;
    iy = iy < H - 1 ? iy : H - 1;
    float md = depth[iy * Wp + ix];
    if ((double)md < 0.01)
    { // first depth > 0.01 in raster order of the clamped END-EXCLUSIVE window (:602-625)
        int xb = sp_x * DSM_SP + DSM_SP / 2 - DSM_SP, yb = sp_y * DSM_SP + DSM_SP / 2 - DSM_SP;
        int xe = xb + DSM_SP * 2, ye = yb + DSM_SP * 2;
        xb = xb > 0 ? xb : 0;
        yb = yb > 0 ? yb : 0;
        xe = xe < W - 1 ? xe : W - 1;
        ye = ye < H - 1 ? ye : H - 1;
        bool found = false;
        for (int j = yb; j < ye && !found; j++)
            for (int i = xb; i < xe; i++)
            {
                float t = depth[j * Wp + i];
                if ((double)t > 0.01)
                {
                    md = t;
                    found = true;
                    break;
                }
            }
    }
    const size_t o = (size_t)b * d.S + s;
    d.seed[o] = make_float4((float)ix, (float)iy, (float)gray[iy * Wp + ix], md);
    d.inv_md[o] = 1.0 / (double)md; // only consumed when md > 0 (:378)
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
struct SeedC
{
    float x, y, I, md;
    double inv;
};

__device__ __forceinline__ bool calc_cost(const SeedC &sd, float pix_i, float pix_inv, int x, int y,
                                          float &nodepth, float &withdepth)
{
    const float ax = sd.x - (float)x, ay = sd.y - (float)y;
    const float dist = ax * ax + ay * ay;
    float n = dist / 16.0f; // (SP_SIZE/2)^2, exact power of two (:374)
    const float idf = sd.I - pix_i;
    n = (float)((double)n + (double)(idf * idf) / 100.0); // (:376)
    nodepth = n;
    withdepth = n;
    if (sd.md > 0 && pix_inv > 0)
    {
        const float idd = (float)(sd.inv - (double)pix_inv);        // (:380)
        withdepth = (float)((double)n + (double)(idd * idd) * 400.0); // (:381)
        return true;
    }
    return false;
}

template <bool FIRST>
__global__ void __launch_bounds__(256) k_assign(const __grid_constant__ DsmDev d)
{
    const int b = blockIdx.z;
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int lane = threadIdx.x & 31;
    const bool active = (x4 < d.W) && (y < d.H);
    // re-arm the chunk-abort slots for the update_seeds pass that follows this one
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.y == 0 && threadIdx.x < 16) d.abortc[b * 16 + threadIdx.x] = INT_MAX;

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
            if (sv[c])
            {
                const float4 s4 = d.seed[so + sidx[c]];
                sc[c].x = s4.x, sc[c].y = s4.y, sc[c].I = s4.z, sc[c].md = s4.w;
                sc[c].inv = d.inv_md[so + sidx[c]];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int x = x4 + i;
            if (x >= d.W) continue;
            const float my_i = gi[i];
            float my_inv = 0.0f;
            if ((double)zi[i] > 0.01) my_inv = (float)(1.0 / (double)zi[i]); // (:404-405)
            float min_d = 1e6f, min_nd = 1e6f;
            int idx_d = -1, idx_nd = -1;
            bool all_has_depth = true;
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                // x%8 == 4 sees only its own column (rx0==4, i==0 -> only xa)
                const bool xvalid = (c >> 1) ? !(rx0 == 4 && i == 0) : true;
                if (sv[c] && xvalid)
                {
                    float cnd, cd;
                    all_has_depth &= calc_cost(sc[c], my_i, my_inv, x, y, cnd, cd);
                    if (cd < min_d)
                    {
                        min_d = cd;
                        idx_d = sidx[c];
                    }
                    if (cnd < min_nd)
                    {
                        min_nd = cnd;
                        idx_nd = sidx[c];
                    }
                }
            }
            win[i] = all_has_depth ? idx_d : idx_nd;
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
    const int b = blockIdx.x;
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
// K2  slic_update — update_seeds_kernel (:468-562)
//
// Block = 32 consecutive seeds.  Phase A (8 warps x 4 seeds): a warp scans the seed's clamped
// 16x16 window (lane = 2*row + half, 8 pixels per lane = raster order), reduces the exactly
// representable integer sums with REDUX, and ballot/scan-compacts the member depths (> 0.1) in
// RASTER ORDER into shared memory.  Phase B (warp 0, lane = seed): the label-affecting float
// sums — sum_depth (:511) and the Huber-Newton sum_a (:536-549) — are order-sensitive (H2), so
// each lane walks its seed's list sequentially exactly like the reference.  Results go to a
// candidate buffer; k_commit_seeds applies the reference's chunk-abort rule (H3).
// Shared list layout: element k of seed sl at dl[k*32 + ((sl+k)&31)]: conflict-free both for
// phase A (fixed seed, consecutive k) and phase B (fixed k, 32 seeds).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_update_seeds(const __grid_constant__ DsmDev d)
{
    __shared__ float dl[256 * 32];
    __shared__ int s_cnt[32], s_sx[32], s_sy[32], s_si[32], s_nd[32];
    const int b = blockIdx.y;
    const int seed0 = blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t fo = (size_t)b * d.px_stride, so = (size_t)b * d.S;
    const int32_t *labels = d.labels + fo;
    const float *depth = d.depth + fo;
    const uint8_t *gray = d.gray + fo;

    for (int q = 0; q < 4; q++)
    {
        const int sl = warp * 4 + q;
        const int s = seed0 + sl;
        if (s >= d.S || d.tstable[so + s] == DSM_STABLE)
        { // stable seeds are skipped (:478-479)
            if (lane == 0) s_cnt[sl] = -1;
            continue;
        }
        const int sp_x = s % d.spw, sp_y = s / d.spw;
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xb = x0 > 0 ? x0 : 0, yb = y0 > 0 ? y0 : 0;
        const int xe = (x0 + 16) < W - 1 ? (x0 + 16) : W - 1; // end-exclusive: last row/col never visited (:488-489)
        const int ye = (y0 + 16) < H - 1 ? (y0 + 16) : H - 1;
        const int y = y0 + (lane >> 1);
        const int xs = x0 + 8 * (lane & 1);
        unsigned m = 0, mdm = 0;
        float dv[8];
        int sumx = 0, sumi = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) dv[k] = 0.f;
        if (y >= yb && y < ye)
        {
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++)
            {
                const int xq = xs + 4 * h2;
                if (xq < 0 || xq >= Wp) continue;
                const int4 l4 = *reinterpret_cast<const int4 *>(labels + (size_t)y * Wp + xq);
                if (l4.x != s && l4.y != s && l4.z != s && l4.w != s) continue;
                const float4 z4 = *reinterpret_cast<const float4 *>(depth + (size_t)y * Wp + xq);
                const uchar4 g4 = *reinterpret_cast<const uchar4 *>(gray + (size_t)y * Wp + xq);
                const int lk[4] = {l4.x, l4.y, l4.z, l4.w};
                const float zk[4] = {z4.x, z4.y, z4.z, z4.w};
                const int gk[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    const int x = xq + k;
                    if (lk[k] == s && x >= xb && x < xe)
                    {
                        m |= 1u << (h2 * 4 + k);
                        sumx += x;
                        sumi += gk[k];
                        if ((double)zk[k] > 0.1) mdm |= 1u << (h2 * 4 + k);
                        dv[h2 * 4 + k] = zk[k];
                    }
                }
            }
        }
        const int cnt_lane = __popc(m);
        const int cnt = __reduce_add_sync(FULL, cnt_lane);
        const int tsx = __reduce_add_sync(FULL, sumx);
        const int tsy = __reduce_add_sync(FULL, cnt_lane * y);
        const int tsi = __reduce_add_sync(FULL, sumi);
        int ndt;
        int pos = warp_excl_scan(__popc(mdm), lane, ndt);
#pragma unroll
        for (int k = 0; k < 8; k++)
            if ((mdm >> k) & 1u)
            {
                dl[pos * 32 + ((sl + pos) & 31)] = dv[k];
                pos++;
            }
        if (lane == 0)
        {
            s_cnt[sl] = cnt;
            s_sx[sl] = tsx;
            s_sy[sl] = tsy;
            s_si[sl] = tsi;
            s_nd[sl] = ndt;
        }
    }
    __syncthreads();
    if (warp != 0) return;

    // ---- phase B: lane == seed-in-block
    const int sl = lane;
    const int s = seed0 + sl;
    if (s >= d.S) return;
    const int n = s_cnt[sl];
    if (n < 0) return; // stable
    if (n == 0)
    { // the reference `return`s here, abandoning the rest of this thread's chunk (:516-517, H3)
        atomicMin(&d.abortc[b * 16 + chunk_of(s, d.S)], s);
        return;
    }
    const float fn = (float)n; // sums below are < 2^24 so the reference's float accumulation is exact
    const float mi = (float)s_si[sl] / fn;
    const float mx = (float)s_sx[sl] / fn;
    const float my = (float)s_sy[sl] / fn;
    const float4 pre = d.seed[so + s];
    // ::fabs(double): float differences, summed in double, rounded once (:527)
    const float diff = (float)(fabs((double)(pre.z - mi)) + fabs((double)(pre.x - mx)) + fabs((double)(pre.y - my)));
    const int newstable = ((double)diff < 0.2) ? 1 : 0;
    const int nd = s_nd[sl];
    float md = 0.0f;
    if (nd > 0)
    {
        float sum_d = 0.0f;
        for (int k = 0; k < nd; k++) sum_d += dl[k * 32 + ((sl + k) & 31)]; // raster order (:511)
        md = sum_d / (float)nd;
        for (int it = 0; it < 5; it++)
        { // damped Huber-Newton (:534-554)
            float sa = 0.0f, sb = 0.0f;
            for (int k = 0; k < nd; k++)
            {
                const float r = md - dl[k * 32 + ((sl + k) & 31)];
                if ((double)r < HUBER_RANGE && (double)r > -HUBER_RANGE)
                {
                    sa += 2 * r;
                    sb += 2;
                }
                else
                    sa = (float)((double)sa + (r > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
            }
            const float delta = (float)((double)(-sa) / ((double)sb + 10.0));
            md = md + delta;
            if ((double)delta < 0.01 && (double)delta > -0.01) break;
        }
    }
    d.cand[so + s] = make_float4(mx, my, mi, md);
    d.cflag[so + s] = newstable;
}

// -------------------------------------------------------------------------------------------
// K2c commit — applies update_seeds results subject to the chunk-abort rule (H3), refreshes the
// hoisted 1/mean_depth, normalises the stable stamps for the next pass and clears the list.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_commit_seeds(const __grid_constant__ DsmDev d)
{
    const int b = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) d.nlist[b] = 0;
    if (s >= d.S) return;
    const size_t o = (size_t)b * d.S + s;
    const int t = d.tstable[o];
    if (t == DSM_STABLE) return; // untouched by update_seeds
    int nt = -1;
    if (s < d.abortc[b * 16 + chunk_of(s, d.S)])
    {
        const float4 c = d.cand[o];
        d.seed[o] = c;
        d.inv_md[o] = 1.0 / (double)c.w;
        if (d.cflag[o] & 1) nt = DSM_STABLE;
    }
    d.tstable[o] = nt;
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
    const int b = blockIdx.z;
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
            if ((double)mz < 0.1 || (double)rz < 0.1 || (double)dz < 0.1) continue; // (:688)
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
            if ((double)view > -MAX_ANGLE_COS && (double)view < MAX_ANGLE_COS) continue; // (:706)
            nx[i] = cxn, ny[i] = cyn, nz[i] = czn;
        }
    }
    *reinterpret_cast<float4 *>(d.nrm + po) = make_float4(nx[0], nx[1], nx[2], nx[3]);
    *reinterpret_cast<float4 *>(d.nrm + d.nrm_plane + po) = make_float4(ny[0], ny[1], ny[2], ny[3]);
    *reinterpret_cast<float4 *>(d.nrm + 2 * d.nrm_plane + po) = make_float4(nz[0], nz[1], nz[2], nz[3]);
}

// -------------------------------------------------------------------------------------------
// K4  seed_plane_fit — calculate_sp_depth_norms_kernel (:792-914) + get_huber_norm (:104-188)
//
// One warp owns 32 consecutive superpixels and works in two phases:
//  A (warp per seed, 32 seeds in turn): scan the 16x16 window (lane = 2*row+half, 8 px/lane),
//    count valid depths (> 0.05), find max_dist, classify inliers |mean_depth-d| < 0.4, reduce the
//    inlier normal / position sums with shuffles, and scan-compact the inlier points into a
//    global scratch list laid out [k][32 seeds] of float4 so that phase B reads are coalesced.
//    Lane sl keeps seed sl's summary in registers.
//  B (thread per seed): the five damped Gauss-Newton steps.  Each lane streams its own seed's
//    points (one coalesced LDG.128 per point per warp), accumulates the 4x4 normal equations in
//    fp64 registers -- no cross-lane reduction at all -- and solves them by symmetric
//    elimination.  Then the superpixel centre is projected onto the plane exactly as (:884-912).
// This stage does not feed the labels, so sums are order-free within the 1e-4 budget
// (SURVEY.md §7 H2/H5); thresholds and float/double promotions follow the reference.
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
    // forward substitution (L y = j)
    const double y0 = j[0];
    const double y1 = j[1] - l10 * y0;
    const double y2 = j[2] - l20 * y0 - l21 * y1;
    const double y3 = j[3] - l30 * y0 - l31 * y1 - l32 * y2;
    // D and back substitution
    u[3] = y3 / a33;
    u[2] = y2 * i2 - l32 * u[3];
    u[1] = y1 * i1 - l21 * u[2] - l31 * u[3];
    u[0] = y0 * i0 - l10 * u[1] - l20 * u[2] - l30 * u[3];
}

#define PF_CAP 232 // >= 15*15 possible members of a superpixel, multiple of 8

__global__ void __launch_bounds__(32) k_plane_fit(const __grid_constant__ DsmDev d)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x;
    const int seed0 = blockIdx.x * 32;
    const int W = d.W, H = d.H, Wp = d.Wp;
    const size_t fo = (size_t)b * d.px_stride, so = (size_t)b * d.S;
    const int32_t *labels = d.labels + fo;
    const float *depth = d.depth + fo;
    const float *nrm = d.nrm + fo;
    float4 *list = d.pflist + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)(PF_CAP * 32);

    // per-lane summary of seed (seed0 + lane), filled in during phase A
    int my_nvalid = 0, my_ninl = 0;
    float my_maxd = 0.f, my_snx = 0.f, my_sny = 0.f, my_snz = 0.f, my_spx = 0.f, my_spy = 0.f, my_spz = 0.f;

    const int row = lane >> 1, half = lane & 1;
    for (int sl = 0; sl < 32; sl++)
    {
        const int s = seed0 + sl;
        if (s >= d.S) break; // warp-uniform
        const float4 sd = d.seed[so + s];
        const int sp_x = s % d.spw, sp_y = s / d.spw;
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int y = y0 + row;
        const int xs = x0 + 8 * half;
        float dv[8];
        unsigned inl = 0;
        int nvalid = 0;
        float maxd = 0.f, snx = 0.f, sny = 0.f, snz = 0.f, spx = 0.f, spy = 0.f, spz = 0.f;
        float kxv[8];
        float kyv = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) dv[k] = 0.f, kxv[k] = 0.f;
        if (y >= 0 && y < H)
        {
            kyv = d.ky[y];
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++)
            {
                const int xq = xs + 4 * h2;
                if (xq < 0 || xq >= Wp) continue;
                const size_t po = (size_t)y * Wp + xq;
                const int4 l4 = *reinterpret_cast<const int4 *>(labels + po);
                if (l4.x != s && l4.y != s && l4.z != s && l4.w != s) continue;
                const float4 z4 = *reinterpret_cast<const float4 *>(depth + po);
                const float4 k4 = *reinterpret_cast<const float4 *>(d.kx + xq);
                const int lk[4] = {l4.x, l4.y, l4.z, l4.w};
                const float zk[4] = {z4.x, z4.y, z4.z, z4.w};
                const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    const int x = xq + k;
                    if (lk[k] != s || x >= W) continue; // window bounded by the flat index only (:816)
                    const float xd = (float)x - sd.x, yd = (float)y - sd.y;
                    const float dist = xd * xd + yd * yd;
                    if (dist > maxd) maxd = dist;
                    const float mz = zk[k];
                    if (!((double)mz > 0.05)) continue; // (:827)
                    nvalid++;
                    const float r = sd.w - mz;
                    if ((double)r < HUBER_RANGE && (double)r > -HUBER_RANGE)
                    { // inlier (:849-860)
                        inl |= 1u << (h2 * 4 + k);
                        dv[h2 * 4 + k] = mz;
                        kxv[h2 * 4 + k] = kk[k];
                        snx += nrm[po + k];
                        sny += nrm[d.nrm_plane + po + k];
                        snz += nrm[2 * d.nrm_plane + po + k];
                        spx += kk[k] * mz; // back_project in float (:94-96)
                        spy += kyv * mz;
                        spz += mz;
                    }
                }
            }
        }
        maxd = warp_max_f(maxd);
        nvalid = __reduce_add_sync(FULL, nvalid);
        int ninl;
        int pos = warp_excl_scan(__popc(inl), lane, ninl);
        const bool ok = nvalid >= 16 && !((double)((float)ninl / (float)nvalid) < 0.8); // (:841, :862) warp-uniform
        if (ok)
        {
            snx = warp_sum_f(snx), sny = warp_sum_f(sny), snz = warp_sum_f(snz);
            spx = warp_sum_f(spx), spy = warp_sum_f(spy), spz = warp_sum_f(spz);
#pragma unroll
            for (int k = 0; k < 8; k++)
                if ((inl >> k) & 1u)
                {
                    list[(size_t)pos * 32 + sl] = make_float4(kxv[k] * dv[k], kyv * dv[k], dv[k], 0.f);
                    pos++;
                }
        }
        if (lane == sl)
        {
            my_nvalid = ok ? nvalid : 0;
            my_ninl = ninl;
            my_maxd = maxd;
            my_snx = snx, my_sny = sny, my_snz = snz;
            my_spx = spx, my_spy = spy, my_spz = spz;
        }
    }
    __syncwarp();
    __threadfence_block(); // phase B reads the scratch list written by other lanes of this warp

    // ---- phase B: thread per seed
    const int s = seed0 + lane;
    if (s >= d.S) return;
    const float4 sd = d.seed[so + s];
    // default record: plane fit rejected -> zero normal / position / view_cos / size (H6-i), Huber mean depth kept
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r1 = make_float4(0.f, 0.f, 0.f, sd.w);
    float4 r2 = make_float4(0.f, sd.z, sd.x, sd.y);
    if (my_nvalid > 0)
    {
        const int n = my_ninl;
        const float len0 = sqrtf(my_snx * my_snx + my_sny * my_sny + my_snz * my_snz);
        float nx = my_snx / len0, ny = my_sny / len0, nz = my_snz / len0, nb = 0.f; // len0 == 0 -> NaN, propagated (H6-iii)
        const float fn = (float)n;
        const float mxs = my_spx / fn, mys = my_spy / fn, mzs = my_spz / fn;
        const float4 *lp = list + lane;
        for (int gn = 0; gn < 5; gn++)
        {
            double j0 = 0, j1 = 0, j2 = 0, j3 = 0;
            double hxx = 0, hxy = 0, hxz = 0, hx = 0, hyy = 0, hyz = 0, hy = 0, hzz = 0, hz = 0, hc = 0;
#pragma unroll 2
            for (int k = 0; k < n; k++)
            {
                const float4 p = lp[(size_t)k * 32];
                const float qx = p.x - mxs, qy = p.y - mys, qz = p.z - mzs; // centred points (:121-126)
                const float r = qx * nx + qy * ny + qz * nz + nb;
                if ((double)r < HUBER_RANGE && (double)r > -1 * HUBER_RANGE)
                { // float products accumulated in double (:136-155)
                    j0 += (double)(2 * r * qx);
                    j1 += (double)(2 * r * qy);
                    j2 += (double)(2 * r * qz);
                    j3 += (double)(2 * r);
                    hxx += (double)(2 * qx * qx);
                    hxy += (double)(2 * qx * qy);
                    hxz += (double)(2 * qx * qz);
                    hx += (double)(2 * qx);
                    hyy += (double)(2 * qy * qy);
                    hyz += (double)(2 * qy * qz);
                    hy += (double)(2 * qy);
                    hzz += (double)(2 * qz * qz);
                    hz += (double)(2 * qz);
                    hc += 2;
                }
                else if ((double)r >= HUBER_RANGE)
                {
                    j0 += HUBER_RANGE * (double)qx;
                    j1 += HUBER_RANGE * (double)qy;
                    j2 += HUBER_RANGE * (double)qz;
                    j3 += HUBER_RANGE;
                }
                else if ((double)r <= -1 * HUBER_RANGE)
                {
                    j0 += -1 * HUBER_RANGE * (double)qx;
                    j1 += -1 * HUBER_RANGE * (double)qy;
                    j2 += -1 * HUBER_RANGE * (double)qz;
                    j3 += -1 * HUBER_RANGE;
                }
            }
            const double hh[10] = {hxx + 5, hxy, hxz, hx, hyy + 5, hyz, hy, hzz + 5, hz, hc + 5}; // LM damping (:172-175)
            const double jj[4] = {j0, j1, j2, j3};
            double u[4];
            solve4_spd(hh, jj, u);
            nx = (float)((double)nx - u[0]);
            ny = (float)((double)ny - u[1]);
            nz = (float)((double)nz - u[2]);
            nb = (float)((double)nb - u[3]);
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
        r2.x = sqrtf(my_maxd);
    }
    float4 *pl = d.plane + (so + s) * 3;
    pl[0] = r0;
    pl[1] = r1;
    pl[2] = r2;
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
    __shared__ float sm[FUSE_BLOCK * 11];
    __shared__ float s_pose[32];
    const int b = blockIdx.y;
    const int begin = d.poolofs[b], end = d.poolofs[b + 1];
    const int first = begin + blockIdx.x * FUSE_BLOCK;
    if (first >= end) return;
    const int cnt = min(FUSE_BLOCK, end - first);
    float *g = reinterpret_cast<float *>(d.pool + first);
    for (int i = threadIdx.x; i < cnt * 11; i += FUSE_BLOCK) sm[i] = g[i];
    if (threadIdx.x < 16) s_pose[threadIdx.x] = d.pose[b * 16 + threadIdx.x];
    else if (threadIdx.x < 32) s_pose[threadIdx.x] = d.ipose[b * 16 + threadIdx.x - 16];
    __syncthreads();
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
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 11; i += FUSE_BLOCK) g[i] = sm[i];
}

// -------------------------------------------------------------------------------------------
// K6  surfel_init — initialize_surfels (:315-361).  One CTA per frame; ordered compaction by
// ballot + block scan so new_surfels come out in seed-index order exactly like the reference's
// serial push_back loop.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_init_surfels(const __grid_constant__ DsmDev d)
{
    __shared__ int s_warp[32];
    __shared__ int s_running;
    __shared__ float s_pose[16];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t so = (size_t)b * d.S;
    if (threadIdx.x < 16) s_pose[threadIdx.x] = d.pose[b * 16 + threadIdx.x];
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    const int ref = d.refidx[b];
    dsm_surfel_t *out = d.newsurf + so;
    for (int base = 0; base < d.S; base += 1024)
    {
        const int s = base + threadIdx.x;
        bool emit = false;
        float4 r0, r1, r2;
        if (s < d.S)
        {
            const float4 *pl = d.plane + (so + s) * 3;
            r0 = pl[0], r1 = pl[1], r2 = pl[2];
            emit = !(r1.w == 0) && !d.fused[so + s] && !((double)r0.w < MAX_ANGLE_COS) && !(r0.x == 0 && r0.y == 0 && r0.z == 0);
        }
        const unsigned bal = __ballot_sync(FULL, emit);
        const int rank = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int wofs = 0, tot = 0;
        for (int w = 0; w < 32; w++)
        {
            const int c = s_warp[w];
            if (w < warp) wofs += c;
            tot += c;
        }
        const int run = s_running;
        if (emit)
        {
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
            out[run + wofs + rank] = e;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_running = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) d.nnew[b] = s_running;
}

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
void dsm_launch_seed_init(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 255) / 256, nb);
    k_seed_init<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_assign(const DsmDev &d, int nb, bool first, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.W + 255) / 256, (d.H + 3) / 4, nb);
    if (first)
        k_assign<true><<<grid, block, 0, s>>>(d);
    else
        k_assign<false><<<grid, block, 0, s>>>(d);
}
void dsm_launch_relax(const DsmDev &d, int nb, cudaStream_t s) { k_relax<<<nb, 1024, 0, s>>>(d); }
void dsm_launch_update_seeds(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 31) / 32, nb);
    k_update_seeds<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_commit_seeds(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 255) / 256, nb);
    k_commit_seeds<<<grid, 256, 0, s>>>(d);
}
void dsm_launch_pixel_normals(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.Wp + 255) / 256, (d.H + 3) / 4, nb);
    k_pixel_normals<<<grid, block, 0, s>>>(d);
}
void dsm_launch_plane_fit(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 31) / 32, nb);
    k_plane_fit<<<grid, 32, 0, s>>>(d);
}
void dsm_launch_fuse(const DsmDev &d, int nb, cudaStream_t s)
{
    if (d.max_pool_per_frame <= 0) return;
    dim3 grid((d.max_pool_per_frame + FUSE_BLOCK - 1) / FUSE_BLOCK, nb);
    k_fuse<<<grid, FUSE_BLOCK, 0, s>>>(d);
}
void dsm_launch_init_surfels(const DsmDev &d, int nb, cudaStream_t s) { k_init_surfels<<<nb, 1024, 0, s>>>(d); }
void dsm_launch_seeds_export(const DsmDev &d, int frame, dsm_seed_t *out_dev, int raw_md, cudaStream_t s)
{
    k_seeds_export<<<(d.S + 255) / 256, 256, 0, s>>>(d, frame, out_dev, raw_md);
}
