/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of the DenseSurfelMapping per-frame hot path.
 *
 * This file is the readable, single-threaded, plain-C specification the CUDA kernels are
 * tested against.  It is NOT part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's OWN source compiled here
 * (oracle/_ref/libdsm_ref_serial.so, see oracle/ref_driver.cpp): tests/test_oracle.py demands
 * byte-identical labels, seeds and surfels on seeded synthetic frames, and small golden
 * vectors minted from the serialised reference are committed under tests/golden/.
 *
 * Every function cites the reference lines it follows
 * (all in /root/reference/surfel_fusion/src/fusion_functions.cpp unless noted).
 *
 * Arithmetic model: the reference is compiled for baseline x86-64
 * (SSE2, no FMA, FLT_EVAL_METHOD 0), so `float op float` rounds to float, any expression
 * touching a double literal (0.01, 0.1, 0.4, 100.0, 400.0, 10.0, 1.0 ...) is evaluated in
 * double and rounded once on assignment to float.  This file is built with
 * -ffp-contract=off and spells every such promotion explicitly.  Unqualified fabs() in the
 * reference TU resolves to ::fabs(double) (probed with the shim headers), reproduced here.
 *
 * Documented definitions of reference UB:
 *   - `Superpixel_seed this_sp;` (:593) is zero-initialised (serial oracle is built with
 *     -ftrivial-auto-var-init=zero), so seeds rejected by the plane fit keep norm == 0.
 *   - W%8 > 4 or H%8 > 4 is rejected (edge pixels would have no candidate seed, :408-451).
 *   - thread bodies run in thread_i order (10-way partitions kept, needed for the
 *     update_seeds early-`return` quirk, :516-517).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SP_SIZE 8            /* fusion_functions.h:10 */
#define ITERATION_NUM 3      /* fusion_functions.h:8  */
#define THREAD_NUM 10        /* fusion_functions.h:9  */
#define MAX_ANGLE_COS 0.1    /* fusion_functions.h:11 */
/* The reference selects these four at compile time by (un)commenting one of two #define blocks (fusion_functions.h:12-21);
 * here they are process-wide variables so that one library serves both sets: "drive" (the default, :13-16) and "RGBD" (:18-21).
 * dsmor_set_constants switches (pyoracle sets them before every call of a wrapper object). */
static double g_huber_range = 0.4, g_baseline = 0.5, g_disparity_error = 4.0, g_min_tolerate_diff = 0.1;
#define HUBER_RANGE g_huber_range
#define BASELINE g_baseline
#define DISPARITY_ERROR g_disparity_error
#define MIN_TOLERATE_DIFF g_min_tolerate_diff

/* elements.h:5-20 (60 bytes) */
typedef struct
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
    float min_eigen_value;
    float max_eigen_value;
} seed_t;

/* elements.h:22-31 (44 bytes) */
typedef struct
{
    float px, py, pz;
    float nx, ny, nz;
    float size;
    float color;
    float weight;
    int update_times;
    int last_update;
} surfel_t;

typedef char seed_size_check[(sizeof(seed_t) == 60) ? 1 : -1];
typedef char surfel_size_check[(sizeof(surfel_t) == 44) ? 1 : -1];

typedef struct
{
    int W, H, spw, sph, S;
    float fx, fy, cx, cy, fuse_far, fuse_near;
    const uint8_t *gray; /* H x W, continuous */
    const float *depth;  /* H x W, continuous, metres, 0 = invalid */
    seed_t *seeds;       /* S */
    int32_t *labels;     /* W*H */
    float *space;        /* 3*W*H: the reference stores float-computed values in doubles (:658-660) */
    float *normals;      /* 3*W*H */
} ctx_t;

/* ---- initialize (:7-28) ---- */
void dsmor_set_constants(double huber_range, double baseline, double disparity_error, double min_tolerate_diff)
{
    g_huber_range = huber_range;
    g_baseline = baseline;
    g_disparity_error = disparity_error;
    g_min_tolerate_diff = min_tolerate_diff;
}

void *dsmor_create(int w, int h, float fx, float fy, float cx, float cy, float far_d, float near_d)
{
    if (w % SP_SIZE > 4 || h % SP_SIZE > 4 || w < 3 * SP_SIZE || h < 3 * SP_SIZE)
        return NULL;
    ctx_t *c = (ctx_t *)calloc(1, sizeof(ctx_t));
    c->W = w;
    c->H = h;
    c->spw = w / SP_SIZE;
    c->sph = h / SP_SIZE;
    c->S = c->spw * c->sph;
    c->fx = fx;
    c->fy = fy;
    c->cx = cx;
    c->cy = cy;
    c->fuse_far = far_d;
    c->fuse_near = near_d;
    c->seeds = (seed_t *)calloc((size_t)c->S, sizeof(seed_t));
    c->labels = (int32_t *)calloc((size_t)w * h, sizeof(int32_t));
    c->space = (float *)calloc((size_t)w * h * 3, sizeof(float));
    c->normals = (float *)calloc((size_t)w * h * 3, sizeof(float));
    return c;
}

void dsmor_destroy(void *p)
{
    ctx_t *c = (ctx_t *)p;
    if (!c) return;
    free(c->seeds);
    free(c->labels);
    free(c->space);
    free(c->normals);
    free(c);
}

static void chunk_range(int n, int thread_i, int *begin, int *end)
{ /* the static 10-way split used by every per-seed / per-surfel phase (:198-202, :471-475, ...) */
    int step = n / THREAD_NUM;
    *begin = step * thread_i;
    *end = *begin + step;
    if (thread_i == THREAD_NUM - 1) *end = n;
}

/* ---- initialize_seeds_kernel (:577-629) ---- */
static void initialize_seeds(ctx_t *c)
{
    const int W = c->W, H = c->H;
    for (int s = 0; s < c->S; s++)
    {
        int sp_x = s % c->spw, sp_y = s / c->spw;
        int ix = sp_x * SP_SIZE + SP_SIZE / 2;
        int iy = sp_y * SP_SIZE + SP_SIZE / 2;
        ix = ix < W - 1 ? ix : W - 1;
        iy = iy < H - 1 ? iy : H - 1;
        seed_t sp;
        memset(&sp, 0, sizeof(sp)); /* H6-i: defined as zero */
        sp.x = (float)ix;
        sp.y = (float)iy;
        sp.mean_intensity = (float)c->gray[iy * W + ix];
        sp.fused = 0;
        sp.stable = 0;
        sp.mean_depth = c->depth[iy * W + ix];
        if ((double)sp.mean_depth < 0.01)
        { /* first valid depth in raster order of the clamped, END-EXCLUSIVE window (:602-625) */
            int xb = sp_x * SP_SIZE + SP_SIZE / 2 - SP_SIZE, yb = sp_y * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            int xe = xb + SP_SIZE * 2, ye = yb + SP_SIZE * 2;
            xb = xb > 0 ? xb : 0;
            yb = yb > 0 ? yb : 0;
            xe = xe < W - 1 ? xe : W - 1;
            ye = ye < H - 1 ? ye : H - 1;
            int found = 0;
            for (int j = yb; j < ye && !found; j++)
                for (int i = xb; i < xe; i++)
                {
                    float d = c->depth[j * W + i];
                    if ((double)d > 0.01)
                    {
                        sp.mean_depth = d;
                        found = 1;
                        break;
                    }
                }
        }
        c->seeds[s] = sp;
    }
}

/* ---- calculate_cost (:364-387) ---- */
static int calculate_cost(const seed_t *sd, float *nodepth_cost, float *depth_cost,
                          float pix_intensity, float pix_inv_depth, int x, int y)
{
    float nd = 0;
    float dist = (sd->x - (float)x) * (sd->x - (float)x) + (sd->y - (float)y) * (sd->y - (float)y);
    nd += dist / (float)((SP_SIZE / 2) * (SP_SIZE / 2));
    float idiff = sd->mean_intensity - pix_intensity;
    nd = (float)((double)nd + (double)(idiff * idiff) / 100.0);
    *nodepth_cost = nd;
    *depth_cost = nd;
    if (sd->mean_depth > 0 && pix_inv_depth > 0)
    {
        float idd = (float)(1.0 / (double)sd->mean_depth - (double)pix_inv_depth);
        *depth_cost = (float)((double)nd + (double)(idd * idd) * 400.0);
        return 1;
    }
    return 0;
}

/* ---- update_pixels_kernel (:389-453), serial => pure raster order ---- */
static void update_pixels(ctx_t *c)
{
    const int W = c->W, H = c->H;
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++)
        {
            if (c->seeds[c->labels[row * W + col]].stable) continue;
            float my_i = (float)c->gray[row * W + col];
            float my_inv = 0.0f;
            float d = c->depth[row * W + col];
            if ((double)d > 0.01) my_inv = (float)(1.0 / (double)d);
            int bx = col / SP_SIZE, by = row / SP_SIZE;
            float min_d = 1e6f, min_nd = 1e6f;
            int idx_d = -1, idx_nd = -1;
            int all_has_depth = 1;
            for (int ci = -1; ci <= 1; ci++)     /* dx OUTER */
                for (int cj = -1; cj <= 1; cj++) /* dy INNER */
                {
                    int sx = bx + ci, sy = by + cj;
                    int dsx = abs(sx * SP_SIZE + SP_SIZE / 2 - col);
                    int dsy = abs(sy * SP_SIZE + SP_SIZE / 2 - row);
                    if (dsx < SP_SIZE && dsy < SP_SIZE && sx >= 0 && sx < c->spw && sy >= 0 && sy < c->sph)
                    {
                        float cd, cnd;
                        all_has_depth &= calculate_cost(&c->seeds[sy * c->spw + sx], &cnd, &cd, my_i, my_inv, col, row);
                        if (cd < min_d)
                        {
                            min_d = cd;
                            idx_d = sy * c->spw + sx;
                        }
                        if (cnd < min_nd)
                        {
                            min_nd = cnd;
                            idx_nd = sy * c->spw + sx;
                        }
                    }
                }
            int w = all_has_depth ? idx_d : idx_nd;
            /* Outside the supported input domain (valid depth < ~0.02 m) every candidate cost can exceed the
             * 1e6 start value and w stays -1: the reference then indexes superpixel_seeds[-1] (undefined).
             * The restatement stays memory-safe and labels such a pixel 0, like the CUDA path. */
            if (w < 0) w = 0;
            c->labels[row * W + col] = w;
            c->seeds[w].stable = 0;
        }
}

/* test instrumentation: how often the chunk-abandoning `return` (:516-517) fired since the last reset */
static int g_abort_events = 0;
int dsmor_debug_abort_events(int reset)
{
    int v = g_abort_events;
    if (reset) g_abort_events = 0;
    return v;
}

/* ---- update_seeds_kernel (:468-562), with the per-chunk early return (:516-517) ---- */
static void update_seeds(ctx_t *c)
{
    const int W = c->W, H = c->H;
    float dvec[4 * SP_SIZE * SP_SIZE];
    for (int t = 0; t < THREAD_NUM; t++)
    {
        int b, e;
        chunk_range(c->S, t, &b, &e);
        for (int s = b; s < e; s++)
        {
            seed_t *sd = &c->seeds[s];
            if (sd->stable) continue;
            int sp_x = s % c->spw, sp_y = s / c->spw;
            int xb = sp_x * SP_SIZE + SP_SIZE / 2 - SP_SIZE, yb = sp_y * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            int xe = xb + SP_SIZE * 2, ye = yb + SP_SIZE * 2;
            xb = xb > 0 ? xb : 0;
            yb = yb > 0 ? yb : 0;
            xe = xe < W - 1 ? xe : W - 1;
            ye = ye < H - 1 ? ye : H - 1;
            float sum_x = 0, sum_y = 0, sum_i = 0, n_i = 0, sum_d = 0, n_d = 0;
            int nvec = 0;
            for (int j = yb; j < ye; j++)
                for (int i = xb; i < xe; i++)
                    if (c->labels[j * W + i] == s)
                    {
                        sum_x += (float)i;
                        sum_y += (float)j;
                        n_i += 1.0f;
                        sum_i += (float)c->gray[j * W + i];
                        float d = c->depth[j * W + i];
                        if ((double)d > 0.1)
                        {
                            dvec[nvec++] = d;
                            sum_d += d; /* ordered float sum: label-affecting (H2) */
                            n_d += 1.0f;
                        }
                    }
            if (n_i == 0)
            { /* reference `return`: abandons the REST OF THIS CHUNK (H3) */
                g_abort_events++;
                break;
            }
            sum_i /= n_i;
            sum_x /= n_i;
            sum_y /= n_i;
            float pre_i = sd->mean_intensity, pre_x = sd->x, pre_y = sd->y;
            sd->mean_intensity = sum_i;
            sd->x = sum_x;
            sd->y = sum_y;
            /* ::fabs(double): the three terms are float differences summed in double (:527) */
            float diff = (float)(fabs((double)(pre_i - sum_i)) + fabs((double)(pre_x - sum_x)) + fabs((double)(pre_y - sum_y)));
            if ((double)diff < 0.2) sd->stable = 1;
            if (n_d > 0)
            {
                float md = sum_d / n_d;
                for (int it = 0; it < 5; it++)
                { /* damped Huber-Newton refinement of the mean depth (:534-554) */
                    float sa = 0, sb = 0;
                    for (int p = 0; p < nvec; p++)
                    {
                        float r = md - dvec[p];
                        if ((double)r < HUBER_RANGE && (double)r > -HUBER_RANGE)
                        {
                            sa += 2 * r;
                            sb += 2;
                        }
                        else
                            sa = (float)((double)sa + (r > 0 ? HUBER_RANGE : -1 * HUBER_RANGE));
                    }
                    float delta = (float)((double)(-sa) / ((double)sb + 10.0));
                    md = md + delta;
                    if ((double)delta < 0.01 && (double)delta > -0.01) break;
                }
                sd->mean_depth = md;
            }
            else
                sd->mean_depth = 0.0f;
        }
    }
}

/* ---- back_project (:91-97): computed in FLOAT (arguments are float refs), then widened ---- */
static void back_project(const ctx_t *c, float u, float v, float d, float *x, float *y, float *z)
{
    *x = (u - c->cx) / c->fx * d;
    *y = (v - c->cy) / c->fy * d;
    *z = d;
}

/* ---- calculate_spaces_kernel (:644-662) ---- */
static void calculate_spaces(ctx_t *c)
{
    for (int row = 0; row < c->H; row++)
        for (int col = 0; col < c->W; col++)
        {
            int i = row * c->W + col;
            back_project(c, (float)col, (float)row, c->depth[i], &c->space[3 * i], &c->space[3 * i + 1], &c->space[3 * i + 2]);
        }
}

/* ---- calculate_pixels_norms_kernel (:664-712): rows 1..H-2, cols 1..W-2 ---- */
static void calculate_pixel_normals(ctx_t *c)
{
    const int W = c->W, H = c->H;
    for (int row = 1; row < H - 1; row++)
        for (int col = 1; col < W - 1; col++)
        {
            int i = row * W + col;
            float mx = c->space[3 * i], my = c->space[3 * i + 1], mz = c->space[3 * i + 2];
            float rx = c->space[3 * i + 3], ry = c->space[3 * i + 4], rz = c->space[3 * i + 5];
            float dx = c->space[3 * (i + W)], dy = c->space[3 * (i + W) + 1], dz = c->space[3 * (i + W) + 2];
            if ((double)mz < 0.1 || (double)rz < 0.1 || (double)dz < 0.1) continue;
            rx = rx - mx;
            ry = ry - my;
            rz = rz - mz;
            dx = dx - mx;
            dy = dy - my;
            dz = dz - mz;
            float nx = ry * dz - rz * dy;
            float ny = rz * dx - rx * dz;
            float nz = rx * dy - ry * dx;
            float len = sqrtf(nx * nx + ny * ny + nz * nz);
            nx /= len;
            ny /= len;
            nz /= len;
            float view = (nx * mx + ny * my + nz * mz) / sqrtf(mx * mx + my * my + mz * mz);
            if ((double)view > -MAX_ANGLE_COS && (double)view < MAX_ANGLE_COS) continue;
            c->normals[3 * i] = nx;
            c->normals[3 * i + 1] = ny;
            c->normals[3 * i + 2] = nz;
        }
}

/* 4x4 double inverse, adjugate / determinant — same formula as the oracle's Eigen stand-in
 * (oracle/shim/Eigen/Eigen); real Eigen differs in the last bits only. */
static void inverse4d(const double *m, double *out)
{
    double inv[16];
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
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    double inv_det = 1.0 / det;
    for (int i = 0; i < 16; i++) out[i] = inv[i] * inv_det;
}

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

/* ---- get_huber_norm (:104-188): 5 damped Gauss-Newton steps on (n, b), Huber loss ---- */
static void get_huber_norm(float *nx, float *ny, float *nz, float *nb, float *pts, int n)
{
    float sx = 0, sy = 0, sz = 0;
    for (int i = 0; i < n; i++)
    {
        sx += pts[3 * i];
        sy += pts[3 * i + 1];
        sz += pts[3 * i + 2];
    }
    sx /= (float)n;
    sy /= (float)n;
    sz /= (float)n;
    *nb = 0;
    for (int i = 0; i < n; i++)
    {
        pts[3 * i] -= sx;
        pts[3 * i + 1] -= sy;
        pts[3 * i + 2] -= sz;
    }
    for (int gn = 0; gn < 5; gn++)
    {
        double Hm[16], J[4], Hinv[16];
        memset(Hm, 0, sizeof(Hm));
        memset(J, 0, sizeof(J));
        /* Hm is column-major like Eigen: Hm[col*4+row]; it is symmetric so layout is moot */
        for (int i = 0; i < n; i++)
        {
            float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
            float r = px * *nx + py * *ny + pz * *nz + *nb;
            if ((double)r < HUBER_RANGE && (double)r > -1 * HUBER_RANGE)
            { /* products are float (int*float*float), accumulated into doubles */
                J[0] += (double)(2 * r * px);
                J[1] += (double)(2 * r * py);
                J[2] += (double)(2 * r * pz);
                J[3] += (double)(2 * r);
                Hm[0] += (double)(2 * px * px);
                Hm[4] += (double)(2 * px * py);
                Hm[8] += (double)(2 * px * pz);
                Hm[12] += (double)(2 * px);
                Hm[1] += (double)(2 * py * px);
                Hm[5] += (double)(2 * py * py);
                Hm[9] += (double)(2 * py * pz);
                Hm[13] += (double)(2 * py);
                Hm[2] += (double)(2 * pz * px);
                Hm[6] += (double)(2 * pz * py);
                Hm[10] += (double)(2 * pz * pz);
                Hm[14] += (double)(2 * pz);
                Hm[3] += (double)(2 * px);
                Hm[7] += (double)(2 * py);
                Hm[11] += (double)(2 * pz);
                Hm[15] += 2;
            }
            else if ((double)r >= HUBER_RANGE)
            {
                J[0] += HUBER_RANGE * (double)px;
                J[1] += HUBER_RANGE * (double)py;
                J[2] += HUBER_RANGE * (double)pz;
                J[3] += HUBER_RANGE;
            }
            else if ((double)r <= -1 * HUBER_RANGE)
            {
                J[0] += -1 * HUBER_RANGE * (double)px;
                J[1] += -1 * HUBER_RANGE * (double)py;
                J[2] += -1 * HUBER_RANGE * (double)pz;
                J[3] += -1 * HUBER_RANGE;
            }
        }
        Hm[0] += 5;
        Hm[5] += 5;
        Hm[10] += 5;
        Hm[15] += 5;
        inverse4d(Hm, Hinv);
        double upd[4];
        for (int i = 0; i < 4; i++) /* column-by-column accumulation, matches the Eigen stand-in */
            upd[i] = ((Hinv[i] * J[0] + Hinv[4 + i] * J[1]) + Hinv[8 + i] * J[2]) + Hinv[12 + i] * J[3];
        *nx = (float)((double)*nx - upd[0]);
        *ny = (float)((double)*ny - upd[1]);
        *nz = (float)((double)*nz - upd[2]);
        *nb = (float)((double)*nb - upd[3]);
    }
    *nb = *nb - (*nx * sx + *ny * sy + *nz * sz);
    float len = sqrtf(*nx * *nx + *ny * *ny + *nz * *nz);
    *nx /= len;
    *ny /= len;
    *nz /= len;
    *nb /= len;
}

/* ---- calculate_sp_depth_norms_kernel (:792-914) ---- */
static void calculate_seed_planes(ctx_t *c)
{
    const int W = c->W, H = c->H;
    const int N = 4 * SP_SIZE * SP_SIZE;
    float pdepth[4 * SP_SIZE * SP_SIZE], pnorm[3 * 4 * SP_SIZE * SP_SIZE], ppos[3 * 4 * SP_SIZE * SP_SIZE], pin[3 * 4 * SP_SIZE * SP_SIZE];
    (void)N;
    for (int s = 0; s < c->S; s++)
    {
        seed_t *sd = &c->seeds[s];
        int sp_x = s % c->spw, sp_y = s / c->spw;
        int xb = sp_x * SP_SIZE + SP_SIZE / 2 - SP_SIZE, yb = sp_y * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
        int nvalid = 0;
        float valid_depth_num = 0, max_dist = 0;
        for (int j = yb; j < yb + SP_SIZE * 2; j++)
            for (int i = xb; i < xb + SP_SIZE * 2; i++)
            {
                int pi = j * W + i; /* bounded by the flat index only (:816); wrapped pixels never carry label s */
                if (pi < 0 || pi >= W * H) continue;
                if (c->labels[pi] != s) continue;
                float xd = (float)i - sd->x, yd = (float)j - sd->y;
                float dist = xd * xd + yd * yd;
                if (dist > max_dist) max_dist = dist;
                /* depth.at(check_j, check_i): for a pixel that carries label s, (j,i) is in range */
                float d = c->depth[pi];
                if ((double)d > 0.05)
                {
                    pdepth[nvalid] = d;
                    pnorm[3 * nvalid] = c->normals[3 * pi];
                    pnorm[3 * nvalid + 1] = c->normals[3 * pi + 1];
                    pnorm[3 * nvalid + 2] = c->normals[3 * pi + 2];
                    ppos[3 * nvalid] = c->space[3 * pi];
                    ppos[3 * nvalid + 1] = c->space[3 * pi + 1];
                    ppos[3 * nvalid + 2] = c->space[3 * pi + 2];
                    nvalid++;
                    valid_depth_num += 1;
                }
            }
        if (valid_depth_num < 16) continue;
        float mean_depth = sd->mean_depth;
        float nx = 0, ny = 0, nz = 0, nb = 0, inlier_num = 0;
        int nin = 0;
        for (int p = 0; p < nvalid; p++)
        {
            float r = mean_depth - pdepth[p];
            if ((double)r < HUBER_RANGE && (double)r > -HUBER_RANGE)
            {
                nx += pnorm[3 * p];
                ny += pnorm[3 * p + 1];
                nz += pnorm[3 * p + 2];
                inlier_num += 1;
                pin[3 * nin] = ppos[3 * p];
                pin[3 * nin + 1] = ppos[3 * p + 1];
                pin[3 * nin + 2] = ppos[3 * p + 2];
                nin++;
            }
        }
        if ((double)(inlier_num / (float)nvalid) < 0.8) continue;
        float len = sqrtf(nx * nx + ny * ny + nz * nz);
        nx = nx / len; /* len == 0 => NaN normal, propagated exactly like the reference (H6-iii) */
        ny = ny / len;
        nz = nz / len;
        get_huber_norm(&nx, &ny, &nz, &nb, pin, nin);
        float ax_f, ay_f, az_f;
        back_project(c, sd->x, sd->y, mean_depth, &ax_f, &ay_f, &az_f);
        double ax = ax_f, ay = ay_f, az = az_f;
        float k = (float)(-1 * (ax * (double)nx + ay * (double)ny + az * (double)nz) - (double)nb);
        ax += (double)(k * nx);
        ay += (double)(k * ny);
        az += (double)(k * nz);
        mean_depth = (float)az;
        float view_cos = (float)(-1.0 * ((double)nx * ax + (double)ny * ay + (double)nz * az) / sqrt(ax * ax + ay * ay + az * az));
        if (view_cos < 0)
        { /* `x *= -1.0` on a float is exact negation */
            view_cos = -view_cos;
            nx = -nx;
            ny = -ny;
            nz = -nz;
        }
        sd->norm_x = nx;
        sd->norm_y = ny;
        sd->norm_z = nz;
        sd->posi_x = (float)ax;
        sd->posi_y = (float)ay;
        sd->posi_z = (float)az;
        sd->mean_depth = mean_depth;
        sd->view_cos = view_cos;
        sd->size = sqrtf(max_dist);
    }
}

/* ---- generate_super_pixels (:960-975) ---- */
static void generate_super_pixels(ctx_t *c)
{
    memset(c->seeds, 0, (size_t)c->S * sizeof(seed_t));
    memset(c->labels, 0, (size_t)c->W * c->H * sizeof(int32_t));
    memset(c->normals, 0, (size_t)c->W * c->H * 3 * sizeof(float));
    initialize_seeds(c);
    for (int it = 0; it < ITERATION_NUM; it++)
    {
        update_pixels(c);
        update_seeds(c);
    }
    calculate_spaces(c);
    calculate_pixel_normals(c);
    calculate_seed_planes(c);
}

/* column-major 4x4 (Eigen::Matrix4f memory order) times (x,y,z,w); column-by-column accumulation */
static void mat4_mul_vec4(const float *m, float x, float y, float z, float w, float *out)
{
    for (int i = 0; i < 4; i++) out[i] = ((m[i] * x + m[4 + i] * y) + m[8 + i] * z) + m[12 + i] * w;
}
static void mat3_mul_vec3(const float *m /*4x4 col-major, top-left block*/, float x, float y, float z, float *out)
{
    for (int i = 0; i < 3; i++) out[i] = (m[i] * x + m[4 + i] * y) + m[8 + i] * z;
}

static float get_weight(float depth) /* :99-102 */
{
    double w = 1.0 / (double)depth / (double)depth;
    return (float)((1.0 < w) ? 1.0 : w); /* std::min(a, b) == (b < a) ? b : a, so a NaN weight propagates */
}

/* ---- fuse_surfels_kernel (:190-313), all ten chunks in order ---- */
static void fuse_surfels(ctx_t *c, int ref_idx, const float *pose, const float *inv_pose, surfel_t *ls, int n)
{
    const int W = c->W, H = c->H;
    for (int i = 0; i < n; i++)
    {
        surfel_t *e = &ls[i];
        if (ref_idx - e->last_update > 5 && e->update_times < 5)
        {
            e->update_times = 0;
            continue;
        }
        if (e->update_times == 0) continue;
        float pc[4];
        mat4_mul_vec4(inv_pose, e->px, e->py, e->pz, 1.0f, pc);
        if (pc[2] < c->fuse_near || pc[2] > c->fuse_far) continue;
        float nc[3];
        mat3_mul_vec3(inv_pose, e->nx, e->ny, e->nz, nc);
        float pu = pc[0] * c->fx / pc[2] + c->cx; /* project (:85-89) */
        float pv = pc[1] * c->fy / pc[2] + c->cy;
        int ui = (int)((double)pu + 0.5);
        int vi = (int)((double)pv + 0.5);
        if (ui < 1 || ui > W - 2 || vi < 1 || vi > H - 2) continue;
        if ((double)pc[2] < (double)c->depth[vi * W + ui] - 1.0)
        {
            e->update_times = 0;
            continue;
        }
        seed_t *sd = &c->seeds[c->labels[vi * W + ui]];
        if (sd->norm_x == 0 && sd->norm_y == 0 && sd->norm_z == 0) continue;
        if ((double)sd->view_cos < MAX_ANGLE_COS) continue;
        float camera_f = (float)((fabs((double)c->fx) + fabs((double)c->fy)) / 2.0);
        float tol = (float)((double)(pc[2] * pc[2]) / (BASELINE * (double)camera_f) * DISPARITY_ERROR);
        tol = (double)tol < MIN_TOLERATE_DIFF ? (float)MIN_TOLERATE_DIFF : tol;
        if (pc[2] < sd->mean_depth - tol) continue;
        if (pc[2] > sd->mean_depth + tol) continue;
        float ndc = nc[0] * sd->norm_x + nc[1] * sd->norm_y + nc[2] * sd->norm_z;
        if ((double)ndc < MAX_ANGLE_COS)
        {
            e->update_times = 0;
            continue;
        }
        float ow = e->weight;
        float nw = get_weight(sd->mean_depth);
        float sw = ow + nw;
        float pw[4];
        mat4_mul_vec4(pose, sd->posi_x, sd->posi_y, sd->posi_z, 1.0f, pw);
        float fpx = (e->px * ow + nw * pw[0]) / sw;
        float fpy = (e->py * ow + nw * pw[1]) / sw;
        float fpz = (e->pz * ow + nw * pw[2]) / sw;
        float fnx = nc[0] * ow + nw * sd->norm_x;
        float fny = nc[1] * ow + nw * sd->norm_y;
        float fnz = nc[2] * ow + nw * sd->norm_z;
        double nl = (double)sqrtf(fnx * fnx + fny * fny + fnz * fnz); /* std::sqrt(float) -> float, stored in a double (:288) */
        fnx = (float)((double)fnx / nl);
        fny = (float)((double)fny / nl);
        fnz = (float)((double)fnz / nl);
        float nwld[3];
        mat3_mul_vec3(pose, fnx, fny, fnz, nwld);
        e->px = fpx;
        e->py = fpy;
        e->pz = fpz;
        e->nx = nwld[0];
        e->ny = nwld[1];
        e->nz = nwld[2];
        e->weight = sw;
        e->color = sd->mean_intensity;
        float new_size = (float)((double)sd->size * fabs((double)(sd->mean_depth / (camera_f * sd->view_cos))));
        if (new_size < e->size) e->size = new_size;
        e->last_update = ref_idx;
        e->update_times += 1;
        sd->fused = 1;
    }
}

/* ---- initialize_surfels (:315-361): serial over seeds, index order ---- */
static int initialize_surfels(ctx_t *c, int ref_idx, const float *pose, surfel_t *out, int cap)
{
    int n = 0;
    for (int s = 0; s < c->S; s++)
    {
        const seed_t *sd = &c->seeds[s];
        if (sd->mean_depth == 0) continue;
        if (sd->fused) continue;
        if ((double)sd->view_cos < MAX_ANGLE_COS) continue;
        if (sd->norm_x == 0 && sd->norm_y == 0 && sd->norm_z == 0) continue;
        float pw[4], nw[3];
        mat4_mul_vec4(pose, sd->posi_x, sd->posi_y, sd->posi_z, 1.0f, pw);
        mat3_mul_vec3(pose, sd->norm_x, sd->norm_y, sd->norm_z, nw);
        surfel_t e;
        e.px = pw[0];
        e.py = pw[1];
        e.pz = pw[2];
        e.nx = nw[0];
        e.ny = nw[1];
        e.nz = nw[2];
        float camera_f = (float)((fabs((double)c->fx) + fabs((double)c->fy)) / 2.0);
        e.size = (float)((double)sd->size * fabs((double)(sd->mean_depth / (camera_f * sd->view_cos))));
        e.color = sd->mean_intensity;
        e.weight = get_weight(sd->mean_depth);
        e.update_times = 1;
        e.last_update = ref_idx;
        if (n < cap) out[n] = e;
        n++;
    }
    return n;
}

/* ---- the caller-side steps the GPU-resident pool mode takes over ----
 * These follow surfel_fusion/src/surfel_map.cpp and are pinned byte for byte against that file compiled in place
 * behind stand-in ROS / PCL headers (oracle/ref_map_driver.cpp -> oracle/_ref/libdsm_refmap.so, tests/test_refmap.py). */

/* SurfelMap::fuse_map post-step (surfel_map.cpp:1077-1109): recycle deleted slots from the highest
 * index for the new surfels, push_back the rest, then swap-with-back the leftover deleted slots.
 * local must have room for n_local + n_new elements.  Returns the new size. */
int dsmor_fuse_map_poststep(void *local_v, int n_local, const void *fresh_v, int n_new)
{
    surfel_t *local = (surfel_t *)local_v;
    const surfel_t *fresh = (const surfel_t *)fresh_v;
    int *deleted = (int *)malloc(sizeof(int) * (size_t)(n_local > 0 ? n_local : 1));
    int nd = 0, size = n_local;
    for (int i = 0; i < n_local; i++)
        if (local[i].update_times == 0) deleted[nd++] = i;
    for (int i = 0; i < n_new; i++)
    {
        if (fresh[i].update_times != 0)
        {
            if (nd > 0)
                local[deleted[--nd]] = fresh[i];
            else
                local[size++] = fresh[i];
        }
    }
    while (nd > 0)
    {
        local[deleted[--nd]] = local[size - 1];
        size--;
    }
    free(deleted);
    return size;
}

/* move_add_surfels, removal half (surfel_map.cpp:1479-1497): live surfels last updated by keyframe
 * `kf` are copied to `out` in pool order and flagged dead.  Returns how many. */
int dsmor_retire(void *local_v, int n_local, int kf, void *out_v)
{
    surfel_t *local = (surfel_t *)local_v, *out = (surfel_t *)out_v;
    int n = 0;
    for (int i = 0; i < n_local; i++)
        if (local[i].update_times > 0 && local[i].last_update == kf)
        {
            out[n++] = local[i];
            local[i].update_times = 0;
        }
    return n;
}

/* warp_active_surfels_cpu_kernel (surfel_map.cpp:750-789): p <- W p (homogeneous), n <- R_W n */
void dsmor_warp_active(void *surfels_v, int n, const float *W)
{
    surfel_t *e = (surfel_t *)surfels_v;
    for (int i = 0; i < n; i++)
    {
        float p[4], nn[3];
        mat4_mul_vec4(W, e[i].px, e[i].py, e[i].pz, 1.0f, p);
        mat3_mul_vec3(W, e[i].nx, e[i].ny, e[i].nz, nn);
        e[i].px = p[0], e[i].py = p[1], e[i].pz = p[2];
        e[i].nx = nn[0], e[i].ny = nn[1], e[i].nz = nn[2];
    }
}

/* ================= exported test entry points (same shape as oracle/ref_driver.cpp) ================= */
void dsmor_superpixels(void *p, const uint8_t *gray, const float *depth)
{
    ctx_t *c = (ctx_t *)p;
    c->gray = gray;
    c->depth = depth;
    generate_super_pixels(c);
}

/* fuse_initialize_map (:30-83) */
int dsmor_fuse(void *p, int ref_idx, const uint8_t *gray, const float *depth, const float *pose_colmajor16,
               void *local, int n_local, void *new_out, int cap_new)
{
    ctx_t *c = (ctx_t *)p;
    c->gray = gray;
    c->depth = depth;
    generate_super_pixels(c);
    float inv_pose[16];
    inverse4f(pose_colmajor16, inv_pose);
    fuse_surfels(c, ref_idx, pose_colmajor16, inv_pose, (surfel_t *)local, n_local);
    return initialize_surfels(c, ref_idx, pose_colmajor16, (surfel_t *)new_out, cap_new);
}

/* test hook: seeds init + `iters` assign passes, the last one optionally without update_seeds;
 * lets the GPU tests localise a mismatch to a single pass */
void dsmor_debug_iters(void *p, const uint8_t *gray, const float *depth, int iters, int last_with_update)
{
    ctx_t *c = (ctx_t *)p;
    c->gray = gray;
    c->depth = depth;
    memset(c->seeds, 0, (size_t)c->S * sizeof(seed_t));
    memset(c->labels, 0, (size_t)c->W * c->H * sizeof(int32_t));
    memset(c->normals, 0, (size_t)c->W * c->H * 3 * sizeof(float));
    initialize_seeds(c);
    for (int it = 0; it < iters; it++)
    {
        update_pixels(c);
        if (it < iters - 1 || last_with_update) update_seeds(c);
    }
}

void dsmor_get_labels(void *p, int32_t *out)
{
    ctx_t *c = (ctx_t *)p;
    memcpy(out, c->labels, (size_t)c->W * c->H * sizeof(int32_t));
}
void dsmor_get_seeds(void *p, void *out)
{
    ctx_t *c = (ctx_t *)p;
    memcpy(out, c->seeds, (size_t)c->S * sizeof(seed_t));
}
void dsmor_get_norm_map(void *p, float *out)
{
    ctx_t *c = (ctx_t *)p;
    memcpy(out, c->normals, (size_t)c->W * c->H * 3 * sizeof(float));
}
int dsmor_num_seeds(void *p) { return ((ctx_t *)p)->S; }
