// Tile kernels of the superpixel pipeline (the default schedule): filtered assign pass, fused window-gather +
// Huber-Newton update on TMA-staged seed tiles, fused pixel-normal + plane-fit gather, and the plane-fit solver.
//
// Why they look the way they do (measured on the round-1 kernels, profiles/r1_final_*.csv):
//  * assign was bound by the float<->double conversion unit: the reference's cost expression (:364-387) crosses
//    the float/double boundary seven times per (pixel, candidate).  Only the ARGMIN of the costs reaches the
//    labels, so k_assign2 evaluates the costs in plain fp32 with a rigorous error bound (a filtered predicate,
//    as in exact computational geometry): if the two smallest costs are separated by more than the bound the
//    argmin is certain; otherwise (exact ties, near ties: ~1e-5 of real pixels) the pixel is re-evaluated with
//    the reference's exact mixed-precision expression.  Labels stay bit-identical by construction.
//  * the window gathers were bound by the L1 data pipe (six 16-byte global loads per lane over 8 image rows,
//    4x redundant label tests, bank-conflicting compaction stores) and each was followed by a separate per-seed
//    kernel reading the lists back from L2.  k_update / k_plane_gather stage one 8x4-seed tile (72x40 pixels)
//    with three TMA tensor copies, scan windows with lane = window row (16 consecutive pixels per lane,
//    conflict-free LDS.128), compact into shared memory and run the order-sensitive per-seed math from there.
// All citations ":NNN" refer to /root/reference/surfel_fusion/src/fusion_functions.cpp.
#include "dsm_exact.cuh"

// -------------------------------------------------------------------------------------------
// mbarrier / TMA helpers (raw PTX; one elected thread arms the barrier and issues the tensor copies)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // visible to the async proxy before the first copy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a descriptor or byte-count mistake must trap, not hang the GPU (each try_wait blocks for a hardware time slice).
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; spin++)
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok && spin > (1u << 20)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

// -------------------------------------------------------------------------------------------
// K1  slic_assign (filtered) — update_pixels_kernel (:389-453) + calculate_cost (:364-387)
//
// Geometry, candidate order and the raster-order `stable` semantics are those of k_assign (dsm_kernels.cu).
// What changes is how the winner is found:
//   fast path   costs in fp32 (FMA allowed).  Against the reference's value c = fl24(..fl53(..)) the fp32 value c~
//               differs by at most 2^-20 (c~ + c) + 1e-12: squared distance and intensity term carry <= 3 roundings of
//               2^-24 each on either side; the depth term uses 1/mean_depth as hi + lo (error 2^-48 / mean_depth,
//               absolute floor 400 (2 |D| eta + eta^2) <= 1e-12 for 1/mean_depth <= 1024) and the EXACT per-pixel
//               inverse depth (invd plane, correctly rounded once by the first pass).  The test below uses
//               2^-18 (m1 + m2) + 1e-10, i.e. >= 4x slack.  If the smallest and second smallest fp32 cost are
//               separated by more than that, the reference's float comparison (:427, :432, strict '<') picks the
//               same candidate whatever the visiting order, and it is below the 1e6 start value (m1 < 9e5).
//   exact path  everything else (ties -- common on synthetic constant images --, near ties, all costs >= 9e5):
//               calc_cost, expression by expression as the reference.
// The first pass also writes the inverse-depth plane the later passes read instead of the depth image.
// The last CTA of a frame to finish (ticket counter) runs the raster-order `stable` relaxation of k_relax for
// that frame, so the pass needs no separate one-CTA-per-frame launch.
// -------------------------------------------------------------------------------------------
#define A_EPS 3.814697265625e-06f // 2^-18
#define A_ALPHA 1e-10f
#define A_BIG 1e30f

__device__ __forceinline__ void relax_frame(const DsmDev &d, int b, int tid, int nthreads)
{
    const int n = __ldcg(&d.nlist[b]);
    if (n > 0)
    {
        const size_t fo = (size_t)b * d.px_stride;
        int2 *list = d.list + fo;
        int32_t *labels = d.labels + fo;
        int32_t *t = d.tstable + (size_t)b * d.S;
        for (;;)
        {
            int changed = 0;
            for (int e = tid; e < n; e += nthreads)
            {
                const int2 en = __ldcg(&list[e]);
                if (en.x < 0) continue; // already evaluated
                const int owner = __ldcg(&labels[en.x]);
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
    if (tid == 0)
    {
        d.nlist[b] = 0; // the deferred-pixel list of this pass is consumed
        d.done[b] = 0;
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(256, 3) k_assign2(const __grid_constant__ DsmDev d)
{
    __shared__ int s_last;
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
        float iv[4];
        if (FIRST)
        { // (:404-405) my_inv = (float)(1.0 / (double)depth) for depth > 0.01: 53 >= 2*24+2 bits, so the double rounding is
          // innocuous and the correctly rounded float reciprocal is the same value (-ftz=false: subnormals included)
            const float4 z4 = *reinterpret_cast<const float4 *>(d.depth + po);
            iv[0] = (z4.x > F_0p01_LO) ? __frcp_rn(z4.x) : 0.0f;
            iv[1] = (z4.y > F_0p01_LO) ? __frcp_rn(z4.y) : 0.0f;
            iv[2] = (z4.z > F_0p01_LO) ? __frcp_rn(z4.z) : 0.0f;
            iv[3] = (z4.w > F_0p01_LO) ? __frcp_rn(z4.w) : 0.0f;
            *reinterpret_cast<float4 *>(d.invd + po) = make_float4(iv[0], iv[1], iv[2], iv[3]);
        }
        else
        {
            const float4 i4 = *reinterpret_cast<const float4 *>(d.invd + po);
            iv[0] = i4.x, iv[1] = i4.y, iv[2] = i4.z, iv[3] = i4.w;
            const int4 l4 = *reinterpret_cast<const int4 *>(d.labels + po);
            L[0] = l4.x, L[1] = l4.y, L[2] = l4.z, L[3] = l4.w;
        }
        const float gi[4] = {(float)g4.x, (float)g4.y, (float)g4.z, (float)g4.w};
        const int bx = x4 >> 3, by = y >> 3, rx0 = x4 & 7, ry = y & 7;
        const int xa = (rx0 == 0) ? bx - 1 : bx, xb = xa + 1;
        const int ya = (ry < 4) ? by - 1 : by, yb = ya + 1;
        const bool vxa = xa >= 0 && xa < d.spw, vxb = xb >= 0 && xb < d.spw;
        const bool vya = ya >= 0 && ya < d.sph, vyb = (ry != 4) && yb >= 0 && yb < d.sph;
        // candidate c = 2*ix + iy  -> (xa,ya) (xa,yb) (xb,ya) (xb,yb): dx outer, dy inner (:413-414)
        float4 sd[4];
        float2 hl[4];
        bool sv[4];
        int sidx[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            const int cx = (c >> 1) ? xb : xa, cy = (c & 1) ? yb : ya;
            sv[c] = ((c >> 1) ? vxb : vxa) && ((c & 1) ? vyb : vya);
            sidx[c] = cy * d.spw + cx;
            const int li = sv[c] ? sidx[c] : 0; // invalid candidates read seed 0 and are masked out below
            sd[c] = d.seed[so + li];
            hl[c] = d.seed_hl[so + li];
        }
        const float fy = (float)y;
        unsigned uncertain = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const float fx = (float)(x4 + i);
            const float pi = gi[i], pv = iv[i];
            const bool hp = pv > 0.f;
            float cn[4], cd[4];
            bool vc[4];
            bool all_has_depth = true;
#pragma unroll
            for (int c = 0; c < 4; c++)
            { // x%8 == 4 sees only its own column (rx0==4, i==0 -> only xa), (:418-420)
                vc[c] = sv[c] && ((c >> 1) ? !(rx0 == 4 && i == 0) : true);
                const float ax = sd[c].x - fx, ay = sd[c].y - fy;
                const float n = fmaf(ax, ax, ay * ay) * 0.0625f;
                const float idf = sd[c].z - pi;
                cn[c] = fmaf(idf * idf, 0.01f, n);
                const float t = (hl[c].x - pv) + hl[c].y;
                cd[c] = fmaf(t * t, 400.f, cn[c]);
                all_has_depth &= (sd[c].w > 0.f && hp) || !vc[c];
            }
            float m1 = A_BIG, m2 = A_BIG;
            int i1 = -1;
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                const float cost = all_has_depth ? cd[c] : cn[c];
                if (vc[c])
                {
                    if (cost < m1)
                    {
                        m2 = m1;
                        m1 = cost;
                        i1 = sidx[c];
                    }
                    else if (cost < m2)
                        m2 = cost;
                }
            }
            const bool certain = (m2 - m1 > A_EPS * (m2 + m1) + A_ALPHA) && (m1 < 9e5f);
            win[i] = i1;
            if (!certain) uncertain |= 1u << i;
        }
        if (uncertain)
        { // exact path: the reference's expression, candidate order and strict '<' (first wins)
            SeedC sc[4];
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                sc[c].x = sd[c].x, sc[c].y = sd[c].y, sc[c].I = sd[c].z, sc[c].md = sd[c].w;
                sc[c].inv = 1.0 / (double)sd[c].w; // only consumed when mean_depth > 0 (:378)
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                if (!((uncertain >> i) & 1u)) continue;
                const float fx = (float)(x4 + i);
                const float my_inv = iv[i];
                const double my_inv_d = (double)my_inv;
                float min_d = 1e6f, min_nd = 1e6f;
                int idx_d = -1, idx_nd = -1;
                bool all_has_depth = true;
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    const bool valid = sv[c] && ((c >> 1) ? !(rx0 == 4 && i == 0) : true);
                    float cnd, cdd;
                    const bool has = calc_cost(sc[c], gi[i], my_inv, my_inv_d, fx, fy, cnd, cdd);
                    cdd = valid ? cdd : __int_as_float(0x7f800000);
                    cnd = valid ? cnd : __int_as_float(0x7f800000);
                    all_has_depth &= has || !valid;
                    const bool bd = cdd < min_d, bn = cnd < min_nd;
                    min_d = bd ? cdd : min_d;
                    idx_d = bd ? sidx[c] : idx_d;
                    min_nd = bn ? cnd : min_nd;
                    idx_nd = bn ? sidx[c] : idx_nd;
                }
                win[i] = all_has_depth ? idx_d : idx_nd;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (x4 + i >= d.W) win[i] = -1;
    }

    if (FIRST)
    { // every label is 0 and seed 0 is unstable: everything commits (:400)
        if (active)
        {
            const size_t po = fo + (size_t)y * d.Wp + x4;
            *reinterpret_cast<int4 *>(d.labels + po) = make_int4(win[0] < 0 ? 0 : win[0], win[1] < 0 ? 0 : win[1],
                                                                 win[2] < 0 ? 0 : win[2], win[3] < 0 ? 0 : win[3]);
        }
        return;
    }

    // ---- iterations 2..: commit / defer (see k_assign)
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
    // frame-completion ticket: the CTA that takes the last ticket of frame b sees every other CTA's labels, list
    // entries and time stamps (release: fence before the ticket; acquire: fence after it) and resolves the frame
    const int tid = threadIdx.y * 64 + threadIdx.x;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&d.done[b], 1) == (int)(gridDim.x * gridDim.y) - 1) ? 1 : 0;
    __syncthreads();
    if (s_last)
    {
        __threadfence();
        relax_frame(d, b, tid, 256);
    }
}

// -------------------------------------------------------------------------------------------
// K2  slic_update — update_seeds_kernel (:468-562), gather and Huber-Newton in ONE kernel per seed tile
//
// A CTA owns DSM_TILE_SX x DSM_TILE_SY = 32 seeds.  One thread arms an mbarrier and issues three TMA tensor copies
// (labels, depth, gray boxes of 76x41 / 80x41 elements at pixel (64 bx - 4, 32 by - 4); out-of-image parts are
// zero-filled by the hardware and masked by coordinates below exactly like the reference's clamped loops).
// Gather: a half-warp per seed, lane = window row.  A lane reads its row's 16 pixels with conflict-free
// 16-byte shared loads, tests membership once per pixel, and the 16 lanes combine the exactly representable
// integer sums (count, sum x, sum y, sum intensity: < 2^24, so the reference's float accumulation is exact and
// order-free) with packed xor-butterflies.  Member depths > 0.1 are compacted IN RASTER ORDER (row-major = lane
// order, then column order inside the lane) into a per-seed list in shared memory.
// Newton: one thread per seed walks its list sequentially -- the float sums sum_depth (:511) and sum_a (:536-549)
// feed the next pass's costs and are order-sensitive -- with list stride 229 words (conflict-free).
// -------------------------------------------------------------------------------------------
#define UL_STRIDE 229 // >= 15*15 possible members; odd, so 32 lanes reading the same position of 32 lists hit 32 banks
#define TILE_PLANE_BYTES 12544 // 76 * 41 * 4 = 12464 rounded up to 128: TMA destinations are 128-byte aligned
#define UPD_SMEM_LAB 0
#define UPD_SMEM_DEP TILE_PLANE_BYTES
#define UPD_SMEM_GRY (2 * TILE_PLANE_BYTES)
#define UPD_SMEM_LIST (UPD_SMEM_GRY + DSM_TILE_GW * DSM_TILE_H)           // 28208: multiple of 16
#define UPD_SMEM_SUM (UPD_SMEM_LIST + 32 * UL_STRIDE * 4)                 // int4[32]
#define UPD_SMEM_ND (UPD_SMEM_SUM + 32 * 16)                              // int[32] list lengths, int[32] flags
#define UPD_SMEM_BAR (UPD_SMEM_ND + 64 * 4)
#define UPD_SMEM_BYTES (UPD_SMEM_BAR + 16)
#define TILE_TX_BYTES (2u * DSM_TILE_W * DSM_TILE_H * 4u + (unsigned)DSM_TILE_GW * DSM_TILE_H)

__global__ void __launch_bounds__(256, 3) k_update(const __grid_constant__ DsmDev d, const __grid_constant__ DsmMaps mp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int32_t *t_lab = reinterpret_cast<const int32_t *>(smem + UPD_SMEM_LAB);
    const float *t_dep = reinterpret_cast<const float *>(smem + UPD_SMEM_DEP);
    const uint8_t *t_gry = smem + UPD_SMEM_GRY;
    float *lists = reinterpret_cast<float *>(smem + UPD_SMEM_LIST);
    int4 *s_sum = reinterpret_cast<int4 *>(smem + UPD_SMEM_SUM);
    int *s_nd = reinterpret_cast<int *>(smem + UPD_SMEM_ND);
    int *s_act = s_nd + 32;
    const unsigned bar = smem_u32(smem + UPD_SMEM_BAR);

    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int X0 = blockIdx.x * (DSM_TILE_SX * DSM_SP) - DSM_SP / 2, Y0 = blockIdx.y * (DSM_TILE_SY * DSM_SP) - DSM_SP / 2;
    if (threadIdx.x == 0)
    {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, TILE_TX_BYTES);
        tma_load_3d(smem_u32(smem + UPD_SMEM_LAB), &mp.lab, X0, Y0, b, bar);
        tma_load_3d(smem_u32(smem + UPD_SMEM_DEP), &mp.dep, X0, Y0, b, bar);
        tma_load_3d(smem_u32(smem + UPD_SMEM_GRY), &mp.gry, X0, Y0, b, bar);
    }
    const int W = d.W, H = d.H;
    const size_t so = (size_t)b * d.S;
    const int half = lane >> 4, r = lane & 15;
    // the stable flags are fetched while the tile is in flight; stable seeds are skipped by update_seeds (:478-479)
    int tflag[2];
#pragma unroll
    for (int rd = 0; rd < 2; rd++)
    {
        const int sl = rd * 16 + warp * 2 + half;
        const int sp_x = blockIdx.x * DSM_TILE_SX + (sl & 7), sp_y = blockIdx.y * DSM_TILE_SY + (sl >> 3);
        const bool live = sp_x < d.spw && sp_y < d.sph;
        tflag[rd] = live ? d.tstable[so + sp_y * d.spw + sp_x] : DSM_STABLE;
    }
    __syncthreads(); // barrier initialised before anybody polls it
    mbar_wait(bar, 0);
#pragma unroll
    for (int rd = 0; rd < 2; rd++)
    {
        const int sl = rd * 16 + warp * 2 + half;
        const int tx = sl & 7, ty = sl >> 3;
        const int sp_x = blockIdx.x * DSM_TILE_SX + tx, sp_y = blockIdx.y * DSM_TILE_SY + ty;
        const int s = sp_y * d.spw + sp_x;
        const bool act = tflag[rd] != DSM_STABLE;
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int xb = x0 > 0 ? x0 : 0, yb = y0 > 0 ? y0 : 0;
        const int xe = (x0 + 16) < W - 1 ? (x0 + 16) : W - 1; // end-exclusive: last row/col never visited (:488-489)
        const int ye = (y0 + 16) < H - 1 ? (y0 + 16) : H - 1;
        const int y = y0 + r;
        const bool rowin = act && y >= yb && y < ye;
        const int kb = xb - x0, ke = xe - x0; // member columns of the window: kb <= k < ke
        const int trow = ty * DSM_SP + r, tcol = tx * DSM_SP;
        int lk[16], gk[16];
        float zk[16];
        {
            const int4 *pl = reinterpret_cast<const int4 *>(t_lab + trow * DSM_TILE_W + tcol);
            const float4 *pz = reinterpret_cast<const float4 *>(t_dep + trow * DSM_TILE_W + tcol);
            const uint2 *pg = reinterpret_cast<const uint2 *>(t_gry + trow * DSM_TILE_GW + tcol);
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int4 a = pl[q];
                const float4 z = pz[q];
                lk[4 * q] = a.x, lk[4 * q + 1] = a.y, lk[4 * q + 2] = a.z, lk[4 * q + 3] = a.w;
                zk[4 * q] = z.x, zk[4 * q + 1] = z.y, zk[4 * q + 2] = z.z, zk[4 * q + 3] = z.w;
            }
            const uint2 g0 = pg[0], g1 = pg[1];
            const unsigned gw[4] = {g0.x, g0.y, g1.x, g1.y};
#pragma unroll
            for (int k = 0; k < 16; k++) gk[k] = (gw[k >> 2] >> (8 * (k & 3))) & 0xffu;
        }
        int cnt = 0, sdx = 0, si = 0;
        unsigned dm = 0; // bit k: member with depth > 0.1
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            const bool mem = rowin && lk[k] == s && k >= kb && k < ke;
            cnt += mem ? 1 : 0;
            sdx += mem ? k : 0;
            si += mem ? gk[k] : 0;
            if (mem && zk[k] > F_0p1_LO) dm |= 1u << k; // (double)depth > 0.1 (:508)
        }
        // packed 16-lane butterflies: count <= 225 (8 bits) | sum intensity <= 57375 (16 bits);
        // sum (x - x0) <= 3375 (12 bits) | sum (y - y0) <= 3375
        int pa = cnt | (si << 8), pb = sdx | ((cnt * r) << 12);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
        {
            pa += __shfl_xor_sync(FULL, pa, o);
            pb += __shfl_xor_sync(FULL, pb, o);
        }
        const int c = __popc(dm);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
        {
            const int nb = __shfl_up_sync(FULL, incl, o, 16);
            if (r >= o) incl += nb;
        }
        const int ndt = __shfl_sync(FULL, incl, 15, 16);
        float *lp = lists + sl * UL_STRIDE + (incl - c);
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((dm >> k) & 1u) *lp++ = zk[k];
        if (r == 0)
        {
            const int n = pa & 0xff;
            s_sum[sl] = make_int4(n, n * x0 + (pb & 0xfff), n * y0 + (pb >> 12), pa >> 8);
            s_nd[sl] = ndt;
            s_act[sl] = act ? 1 : 0;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    // ---- Huber-Newton, one thread per seed of the tile (text of k_newton, lists in shared memory)
    const int sl = lane;
    if (!s_act[sl]) return; // stable (untouched by update_seeds) or outside the seed grid
    const int sp_x = blockIdx.x * DSM_TILE_SX + (sl & 7), sp_y = blockIdx.y * DSM_TILE_SY + (sl >> 3);
    const int s = sp_y * d.spw + sp_x;
    const int4 su = s_sum[sl];
    const int n = su.x;
    if (n == 0)
    { // unreachable for supported shapes (every seed keeps its centre pixel); recorded, never silently ignored
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
    const int nd = s_nd[sl];
    float md = 0.0f;
    if (nd > 0)
    {
        const float *dl = lists + sl * UL_STRIDE;
        float sum_d = 0.0f;
        {
            int k = 0;
            for (; k + 8 <= nd; k += 8)
            {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = dl[k + j];
#pragma unroll
                for (int j = 0; j < 8; j++) sum_d += v[j]; // raster order (:511)
            }
            for (; k < nd; k++) sum_d += dl[k];
        }
        md = sum_d / (float)nd;
        for (int it = 0; it < 5; it++)
        { // damped Huber-Newton (:534-554)
            float sa = 0.0f, sb = 0.0f;
            int k = 0;
            for (; k + 8 <= nd; k += 8)
            {
                float rr[8];
                bool allin = true;
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    rr[j] = md - dl[k + j];
                    allin &= rr[j] < F_0p4_HI && rr[j] > -F_0p4_HI; // (double)r < 0.4 && (double)r > -0.4
                }
                if (allin)
                { // common case: every residual inside the Huber range -> pure float chain
#pragma unroll
                    for (int j = 0; j < 8; j++) sa += 2 * rr[j];
                    sb += 16; // eight exact +2 steps
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        if (rr[j] < F_0p4_HI && rr[j] > -F_0p4_HI)
                        {
                            sa += 2 * rr[j];
                            sb += 2;
                        }
                        else
                            sa = (float)((double)sa + (rr[j] > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
                    }
                }
            }
            for (; k < nd; k++)
            {
                const float rr = md - dl[k];
                if (rr < F_0p4_HI && rr > -F_0p4_HI)
                {
                    sa += 2 * rr;
                    sb += 2;
                }
                else
                    sa = (float)((double)sa + (rr > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
            }
            const float delta = (float)((double)(-sa) / ((double)sb + 10.0));
            md = md + delta;
            if (delta < F_0p01_HI && delta > -F_0p01_HI) break; // |delta| < 0.01 in double (:552)
        }
    }
    d.seed[so + s] = make_float4(mx, my, mi, md);
    d.seed_hl[so + s] = split_inverse(md);
    d.tstable[so + s] = newstable ? DSM_STABLE : -1;
}

// -------------------------------------------------------------------------------------------
// K3+K4a  plane_gather — calculate_spaces_kernel (:644-662), calculate_pixels_norms_kernel (:664-712) and the
// window scan of calculate_sp_depth_norms_kernel (:792-863) fused on the same TMA-staged seed tile.
//
// The reference computes a normal for EVERY pixel and then reads the ones of a superpixel's inliers.  Here the
// window scan (half-warp per seed, lane = window row) first compacts the INLIER pixel positions of each seed into
// shared memory; a second, dense phase (one warp per seed, lane per inlier) computes exactly those pixels'
// normals from the depth tile -- each inlier belongs to one seed, so no normal is computed twice and none is
// ever written to HBM (the 12 B/px `nrm` planes of the round-1 schedule are gone).  The same phase forms the
// back-projected points, their mean, the centred points (get_huber_norm :111-126), and the first Gauss-Newton
// pass's point sums: H = sum 2 q q^T (fp64, :141-155 with every residual inside the Huber range), max |r| and
// max |q|^2.  If max |r| < 0.4 the solver (k_gn_solve) never needs the points; otherwise it reads the centred
// points this kernel leaves in the [k][seed] list (staged through shared memory: full 32-byte sectors).
// Not label-affecting: float sums are lane-partial + tree (order-free within 1e-4).
// -------------------------------------------------------------------------------------------
#define PF_CAP 228
#define PG_SMEM_LAB 0
#define PG_SMEM_DEP TILE_PLANE_BYTES
#define PG_SMEM_POS (2 * TILE_PLANE_BYTES)                       // u16 [32][PF_CAP + 4]
#define PG_POS_STRIDE 232
#define PG_SMEM_STAGE (PG_SMEM_POS + 32 * PG_POS_STRIDE * 2)     // float [3][8][PF_CAP + 1]
#define PG_ST_STRIDE 229
#define PG_SMEM_KX (PG_SMEM_STAGE + 3 * 8 * PG_ST_STRIDE * 4)    // float [80] kx, float [48] ky
#define PG_SMEM_REC (PG_SMEM_KX + 128 * 4)                       // float maxd[32], int nvalid[32], int ninl[32]
#define PG_SMEM_BAR (PG_SMEM_REC + 96 * 4)
#define PG_SMEM_BYTES (PG_SMEM_BAR + 16)
#define PG_TX_BYTES (2u * DSM_TILE_W * DSM_TILE_H * 4u)

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(256, 3) k_plane_gather(const __grid_constant__ DsmDev d, const __grid_constant__ DsmMaps mp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int32_t *t_lab = reinterpret_cast<const int32_t *>(smem + PG_SMEM_LAB);
    const float *t_dep = reinterpret_cast<const float *>(smem + PG_SMEM_DEP);
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(smem + PG_SMEM_POS);
    float *s_stage = reinterpret_cast<float *>(smem + PG_SMEM_STAGE);
    float *s_kx = reinterpret_cast<float *>(smem + PG_SMEM_KX), *s_ky = s_kx + 80;
    float *s_maxd = reinterpret_cast<float *>(smem + PG_SMEM_REC);
    int *s_nvalid = reinterpret_cast<int *>(s_maxd + 32), *s_ninl = s_nvalid + 32;
    __shared__ int s_rows;
    const unsigned bar = smem_u32(smem + PG_SMEM_BAR);

    const int b = d.frame0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int X0 = blockIdx.x * (DSM_TILE_SX * DSM_SP) - DSM_SP / 2, Y0 = blockIdx.y * (DSM_TILE_SY * DSM_SP) - DSM_SP / 2;
    if (threadIdx.x == 0)
    {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, PG_TX_BYTES);
        tma_load_3d(smem_u32(smem + PG_SMEM_LAB), &mp.lab, X0, Y0, b, bar);
        tma_load_3d(smem_u32(smem + PG_SMEM_DEP), &mp.dep, X0, Y0, b, bar);
    }
    const int W = d.W, H = d.H;
    const size_t so = (size_t)b * d.S;
    // back-projection factors of the tile's columns / rows (tables hold Wp+16 / H+16 entries; clamp the halo)
    if (threadIdx.x < 80)
    {
        int x = X0 + (int)threadIdx.x;
        x = x < 0 ? 0 : (x > d.Wp + 15 ? d.Wp + 15 : x);
        s_kx[threadIdx.x] = d.kx[x];
    }
    else if (threadIdx.x < 128)
    {
        int y = Y0 + (int)threadIdx.x - 80;
        y = y < 0 ? 0 : (y > H + 15 ? H + 15 : y);
        s_ky[threadIdx.x - 80] = d.ky[y];
    }
    const int half = lane >> 4, r = lane & 15;
    float4 sdv[2];
#pragma unroll
    for (int rd = 0; rd < 2; rd++)
    {
        const int sl = rd * 16 + warp * 2 + half;
        const int sp_x = blockIdx.x * DSM_TILE_SX + (sl & 7), sp_y = blockIdx.y * DSM_TILE_SY + (sl >> 3);
        const bool live = sp_x < d.spw && sp_y < d.sph;
        sdv[rd] = d.seed[so + (live ? sp_y * d.spw + sp_x : 0)]; // x, y, I, mean_depth (Huber mean after the 3 iterations)
    }
    __syncthreads();
    mbar_wait(bar, 0);
    // ---- phase 1: window scan, inlier positions in raster order
#pragma unroll
    for (int rd = 0; rd < 2; rd++)
    {
        const int sl = rd * 16 + warp * 2 + half;
        const int tx = sl & 7, ty = sl >> 3;
        const int sp_x = blockIdx.x * DSM_TILE_SX + tx, sp_y = blockIdx.y * DSM_TILE_SY + ty;
        const bool live = sp_x < d.spw && sp_y < d.sph;
        const int s = sp_y * d.spw + sp_x;
        const float4 sd = sdv[rd];
        const int x0 = sp_x * DSM_SP - DSM_SP / 2, y0 = sp_y * DSM_SP - DSM_SP / 2;
        const int y = y0 + r;
        const bool rowin = live && y >= 0 && y < H; // window bounded by the flat index only (:816)
        const int kb = x0 < 0 ? -x0 : 0, ke = (W - x0) < 16 ? (W - x0) : 16;
        const int trow = ty * DSM_SP + r, tcol = tx * DSM_SP;
        int lk[16];
        float zk[16];
        {
            const int4 *pl = reinterpret_cast<const int4 *>(t_lab + trow * DSM_TILE_W + tcol);
            const float4 *pz = reinterpret_cast<const float4 *>(t_dep + trow * DSM_TILE_W + tcol);
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int4 a = pl[q];
                const float4 z = pz[q];
                lk[4 * q] = a.x, lk[4 * q + 1] = a.y, lk[4 * q + 2] = a.z, lk[4 * q + 3] = a.w;
                zk[4 * q] = z.x, zk[4 * q + 1] = z.y, zk[4 * q + 2] = z.z, zk[4 * q + 3] = z.w;
            }
        }
        const float yd = (float)y - sd.y;
        const float yd2 = yd * yd;
        float maxd = 0.f;
        int nvalid = 0;
        unsigned inl = 0;
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            const bool mem = rowin && lk[k] == s && k >= kb && k < ke;
            const float xd = (float)(x0 + k) - sd.x;
            const float dist = xd * xd + yd2;
            if (mem && dist > maxd) maxd = dist; // (:821-823)
            const float mz = zk[k];
            const bool valid = mem && mz > F_0p05_LO; // (double)depth > 0.05 (:827)
            nvalid += valid ? 1 : 0;
            const float rr = sd.w - mz;
            if (valid && rr < F_0p4_HI && rr > -F_0p4_HI) inl |= 1u << k; // inlier (:849-860)
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
        {
            maxd = fmaxf(maxd, __shfl_xor_sync(FULL, maxd, o));
            nvalid += __shfl_xor_sync(FULL, nvalid, o);
        }
        const int c = __popc(inl);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
        {
            const int nb = __shfl_up_sync(FULL, incl, o, 16);
            if (r >= o) incl += nb;
        }
        const int ninl = __shfl_sync(FULL, incl, 15, 16);
        uint16_t *pp = s_pos + sl * PG_POS_STRIDE + (incl - c);
        const int pbase = trow * DSM_TILE_W + tcol;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((inl >> k) & 1u) *pp++ = (uint16_t)(pbase + k);
        if (r == 0)
        {
            s_maxd[sl] = maxd;
            s_nvalid[sl] = nvalid;
            s_ninl[sl] = ninl;
        }
    }
    __syncthreads();
    // ---- phase 2: one warp per seed of a tile row, lane per inlier
    for (int ty = 0; ty < DSM_TILE_SY; ty++)
    {
        if (threadIdx.x == 0) s_rows = 0;
        __syncthreads(); // previous row's copy-out has read the stage tile; s_rows reset
        const int sl = ty * 8 + warp;
        const int sp_x = blockIdx.x * DSM_TILE_SX + warp, sp_y = blockIdx.y * DSM_TILE_SY + ty;
        const bool live = sp_x < d.spw && sp_y < d.sph;
        const int s = sp_y * d.spw + sp_x;
        const int nvalid = s_nvalid[sl], ninl = s_ninl[sl];
        const float maxd = s_maxd[sl];
        const bool ok = live && nvalid >= 16 && !((float)ninl / (float)nvalid < F_0p8_HI); // (:841), (double)ratio < 0.8 (:862)
        if (ok) // warp-uniform
        {
            const uint16_t *pos = s_pos + sl * PG_POS_STRIDE;
            float snx = 0.f, sny = 0.f, snz = 0.f, spx = 0.f, spy = 0.f, spz = 0.f;
            for (int j = lane; j < ninl; j += 32)
            {
                const int p = pos[j];
                const int trow = p / DSM_TILE_W, tcol = p - trow * DSM_TILE_W;
                const int x = X0 + tcol, y = Y0 + trow;
                const float mz = t_dep[p];
                const float kxi = s_kx[tcol], ky0 = s_ky[trow];
                spx += kxi * mz; // back_project in float (:94-96)
                spy += ky0 * mz;
                spz += mz;
                // pixel normal of (x, y), as k_pixel_normals / calculate_pixels_norms_kernel (:664-712)
                if (y >= 1 && y <= H - 2 && x >= 1 && x <= W - 2)
                {
                    const float rz = t_dep[p + 1], dz = t_dep[p + DSM_TILE_W];
                    if (!(mz < F_0p1_HI || rz < F_0p1_HI || dz < F_0p1_HI)) // (double)z < 0.1 (:688)
                    {
                        const float kxr = s_kx[tcol + 1], ky1 = s_ky[trow + 1];
                        const float mx = kxi * mz, my = ky0 * mz;
                        const float rx = kxr * rz - mx, ry = ky0 * rz - my, rzz = rz - mz;
                        const float dx = kxi * dz - mx, dy = ky1 * dz - my, dzz = dz - mz;
                        float cxn = ry * dzz - rzz * dy;
                        float cyn = rzz * dx - rx * dzz;
                        float czn = rx * dy - ry * dx;
                        const float len = sqrtf(cxn * cxn + cyn * cyn + czn * czn);
                        cxn /= len;
                        cyn /= len;
                        czn /= len;
                        const float view = (cxn * mx + cyn * my + czn * mz) / sqrtf(mx * mx + my * my + mz * mz);
                        if (!(view > -F_0p1_HI && view < F_0p1_HI)) // |view| < 0.1 in double -> skipped (:706)
                            snx += cxn, sny += cyn, snz += czn;
                    }
                }
            }
            snx = warp_sum_f(snx), sny = warp_sum_f(sny), snz = warp_sum_f(snz);
            spx = warp_sum_f(spx), spy = warp_sum_f(spy), spz = warp_sum_f(spz);
            const float fn = (float)ninl;
            const float mxs = spx / fn, mys = spy / fn, mzs = spz / fn; // (:117-119)
            // initial normal of get_huber_norm = normalised sum of the inlier pixel normals (:864-871); 0/0 -> NaN, propagated (H6-iii)
            const float len0 = sqrtf(snx * snx + sny * sny + snz * snz);
            const float n0x = snx / len0, n0y = sny / len0, n0z = snz / len0;
            double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // xx xy xz xw yy yz yw zz zw
            float rmax = 0.f, qmax2 = 0.f;
            bool rnan = false;
            float *stx = s_stage + warp * PG_ST_STRIDE, *sty = stx + 8 * PG_ST_STRIDE, *stz = sty + 8 * PG_ST_STRIDE;
            for (int j = lane; j < ninl; j += 32)
            {
                const int p = pos[j];
                const int trow = p / DSM_TILE_W, tcol = p - trow * DSM_TILE_W;
                const float mz = t_dep[p];
                const float ax = s_kx[tcol] * mz - mxs, ay = s_ky[trow] * mz - mys, az = mz - mzs; // centred points (:121-126)
                stx[j] = ax, sty[j] = ay, stz[j] = az;
                const float rr = ax * n0x + ay * n0y + az * n0z + 0.f; // first-pass residual (:133), b = 0
                rmax = fmaxf(rmax, fabsf(rr));
                rnan |= !(rr == rr);
                qmax2 = fmaxf(qmax2, ax * ax + ay * ay + az * az);
                h[0] += (double)(2 * ax * ax), h[1] += (double)(2 * ax * ay), h[2] += (double)(2 * ax * az), h[3] += (double)(2 * ax);
                h[4] += (double)(2 * ay * ay), h[5] += (double)(2 * ay * az), h[6] += (double)(2 * ay);
                h[7] += (double)(2 * az * az), h[8] += (double)(2 * az);
            }
#pragma unroll
            for (int i = 0; i < 9; i++) h[i] = warp_sum_d(h[i]);
            rmax = warp_max_f(rmax);
            qmax2 = warp_max_f(qmax2);
            if (__any_sync(FULL, rnan)) rmax = __int_as_float(0x7f800000); // a NaN residual forces the solver to evaluate every pass
            if (lane < 10)
            {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 9; i++)
                    if (lane == i) v = h[i];
                if (lane == 9) v = __hiloint2double(__float_as_int(qmax2), __float_as_int(rmax));
                d.hrec[(so + s) * 10 + lane] = v;
            }
            if (lane == 0)
            {
                d.pfsum[(so + s) * 2] = make_float4(snx, sny, snz, maxd);
                d.pfsum[(so + s) * 2 + 1] = make_float4(mxs, mys, mzs, __int_as_float(ninl));
                atomicMax(&s_rows, ninl);
            }
        }
        else if (live && lane == 0)
        {
            d.pfsum[(so + s) * 2] = make_float4(0.f, 0.f, 0.f, maxd);
            d.pfsum[(so + s) * 2 + 1] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        }
        __syncthreads();
        // copy-out of the row's centred points as full 32-byte sectors of the [k][seed] lists
        const int rows = s_rows;
        const unsigned c = threadIdx.x & 7u;
        if (blockIdx.x * DSM_TILE_SX + c < (unsigned)d.spw && sp_y < d.sph)
        {
            const size_t plane = (size_t)d.B * PF_CAP * d.Sp;
            float *dx = d.qlist + ((size_t)b * PF_CAP * d.Sp + (size_t)sp_y * d.spw + blockIdx.x * DSM_TILE_SX + c);
            float *dy = dx + plane, *dz = dy + plane;
            const unsigned sp = (unsigned)d.Sp;
            for (unsigned k = threadIdx.x >> 3; k < (unsigned)rows; k += 32u)
            {
                const unsigned t = c * PG_ST_STRIDE + k, o = k * sp;
                dx[o] = s_stage[t];
                dy[o] = s_stage[8 * PG_ST_STRIDE + t];
                dz[o] = s_stage[16 * PG_ST_STRIDE + t];
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// K4b  plane_solve — get_huber_norm (:104-188) + the projection (:884-912), one thread per seed.
// Same algebra as k_gauss_newton (dsm_kernels.cu): H over all points once, passes only classify residuals, pass
// skipping by the |r| bound.  The first pass's H, max |r| and max |q|^2 arrive from k_plane_gather; if that max |r|
// is inside the Huber range (the common case) the solver is five register-only 4x4 solves and never touches the
// point list; otherwise it evaluates the passes from the [k][seed] list exactly like k_gauss_newton.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void solve4_spd_t(const double *h, const double *j, double *u)
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

__global__ void __launch_bounds__(128) k_gn_solve(const __grid_constant__ DsmDev d)
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
        double hall[10]; // xx xy xz xw yy yz yw zz zw ww over ALL points
        const double *hr = d.hrec + (so + s) * 10;
#pragma unroll
        for (int i = 0; i < 9; i++) hall[i] = hr[i];
        hall[9] = 2.0 * (double)n;
        const double pk = hr[9];
        float rmax = __int_as_float(__double2loint(pk)), qmax2 = __int_as_float(__double2hiint(pk));
        // first pass already evaluated by k_plane_gather: every residual inside the Huber range?
        bool need_pass = !(rmax < F_0p4_HI);
        // (if not, the gather's H stays valid -- it is the sum over ALL points -- and only the classification is redone)
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
                    if (!inr)
                    {
                        ho[0] += (double)(2 * ax * ax), ho[1] += (double)(2 * ax * ay), ho[2] += (double)(2 * ax * az), ho[3] += (double)(2 * ax);
                        ho[4] += (double)(2 * ay * ay), ho[5] += (double)(2 * ay * az), ho[6] += (double)(2 * ay);
                        ho[7] += (double)(2 * az * az), ho[8] += (double)(2 * az), ho[9] += 2;
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
            solve4_spd_t(hh, jj, u);
            const float ox = nx, oy = ny, oz = nz, ob = nb;
            nx = (float)((double)nx - u[0]);
            ny = (float)((double)ny - u[1]);
            nz = (float)((double)nz - u[2]);
            nb = (float)((double)nb - u[3]);
            // can the next pass be skipped?  |r_i(new)| <= rmax + qmax |dn| + |db| (+ rounding slack)
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

// -------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------
int dsm_tile_setup()
{
    cudaError_t e = cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, UPD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_plane_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES);
    return e == cudaSuccess ? 0 : -1;
}
void dsm_launch_assign2(const DsmDev &d, int nb, bool first, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((d.W + 255) / 256, (d.H + 3) / 4, nb);
    if (first)
        k_assign2<true><<<grid, block, 0, s>>>(d);
    else
        k_assign2<false><<<grid, block, 0, s>>>(d);
}
void dsm_launch_update(const DsmDev &d, const DsmMaps &m, int nb, cudaStream_t s)
{
    dim3 grid((d.spw + DSM_TILE_SX - 1) / DSM_TILE_SX, (d.sph + DSM_TILE_SY - 1) / DSM_TILE_SY, nb);
    k_update<<<grid, 256, UPD_SMEM_BYTES, s>>>(d, m);
}
void dsm_launch_plane_gather(const DsmDev &d, const DsmMaps &m, int nb, cudaStream_t s)
{
    dim3 grid((d.spw + DSM_TILE_SX - 1) / DSM_TILE_SX, (d.sph + DSM_TILE_SY - 1) / DSM_TILE_SY, nb);
    k_plane_gather<<<grid, 256, PG_SMEM_BYTES, s>>>(d, m);
}
void dsm_launch_gn_solve(const DsmDev &d, int nb, cudaStream_t s)
{
    dim3 grid((d.S + 127) / 128, nb);
    k_gn_solve<<<grid, 128, 0, s>>>(d);
}
