// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// C-ABI driver around the reference's OWN hot-path translation unit
// (/root/reference/surfel_fusion/src/fusion_functions.cpp, compiled where it lies,
// nothing copied).  Built by oracle/Makefile into oracle/_ref/ in two variants:
//
//   libdsm_ref_mt.so      the file as shipped: 10 std::threads per phase
//                         (fusion_functions.h:9 THREAD_NUM).  Non-deterministic
//                         (data race on Superpixel_seed::stable, fusion_functions.cpp:400
//                         vs :445/:450).  Used ONLY as the timed CPU baseline.
//   libdsm_ref_serial.so  same source, -DDSM_REF_SERIAL: `std::thread` is token-replaced
//                         by an inline-invoking stand-in so the 10 per-phase bodies run in
//                         thread_i order on the calling thread (keeps the 10-way partitions,
//                         gives pure raster order for update_pixels), and the build adds
//                         -ftrivial-auto-var-init=zero so the uninitialised `Superpixel_seed
//                         this_sp` (fusion_functions.cpp:593) is defined as zero.
//                         Deterministic => THE PARITY ORACLE.
//
// Private members are reached with the `#define private public` trick so labels / seeds
// can be dumped for comparison.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <utility>

#ifdef DSM_REF_SERIAL
namespace std
{
// Stand-in for std::thread that runs the callable immediately on the calling thread.
class dsm_inline_thread
{
  public:
    dsm_inline_thread() {}
    template <typename Obj, typename... MArgs, typename... Args>
    dsm_inline_thread(void (Obj::*fn)(MArgs...), Obj *self, Args... args)
    {
        (self->*fn)(args...);
    }
    dsm_inline_thread(dsm_inline_thread &&) {}
    dsm_inline_thread &operator=(dsm_inline_thread &&) { return *this; }
    bool joinable() const { return false; }
    void join() {}
};
} // namespace std
#define thread dsm_inline_thread
#endif

#include <opencv2/opencv.hpp> // shim
#include <Eigen/Eigen>        // shim

// silence the three per-call printf timing lines (fusion_functions.cpp:55,75,82)
static inline int dsm_ref_noprintf(const char *, ...) { return 0; }
#define printf dsm_ref_noprintf

#define private public
#ifdef DSM_REF_RGBD
// The reference's second constant set ("for RGBD", fusion_functions.h:17-21) ships commented out.  This variant compiles the
// same unmodified source with exactly those four values: the header is included first (its #pragma once makes the
// source's own #include a no-op), the four macros are redefined, then the source follows.
#include "fusion_functions.h"
#undef HUBER_RANGE
#undef BASELINE
#undef DISPARITY_ERROR
#undef MIN_TOLERATE_DIFF
#define HUBER_RANGE 0.05
#define BASELINE 0.08
#define DISPARITY_ERROR 1.0
#define MIN_TOLERATE_DIFF 0.05
#endif
#include DSM_REF_SOURCE // "/root/reference/surfel_fusion/src/fusion_functions.cpp"
#undef private
#undef printf
#ifdef DSM_REF_SERIAL
#undef thread
#endif

static_assert(sizeof(Superpixel_seed) == 60, "reference Superpixel_seed layout (elements.h:5-20)");
static_assert(sizeof(SurfelElement) == 44, "reference SurfelElement layout (elements.h:22-31)");

struct RefCtx
{
    FusionFunctions ff;
    int w, h;
    std::vector<SurfelElement> local, fresh;
};

extern "C"
{
    void *dsmref_create(int w, int h, float fx, float fy, float cx, float cy, float far_d, float near_d)
    {
        RefCtx *c = new RefCtx();
        c->w = w;
        c->h = h;
        c->ff.initialize(w, h, fx, fy, cx, cy, far_d, near_d);
        return c;
    }
    void dsmref_destroy(void *p) { delete (RefCtx *)p; }

    // reference-identical call: FusionFunctions::fuse_initialize_map (fusion_functions.cpp:30-83).
    // local[n_local] is updated in place; returns number of new surfels written to new_out (<= cap_new).
    int dsmref_fuse(void *p, int ref_idx, const uint8_t *gray, const float *depth,
                    const float *pose_colmajor16, void *local, int n_local, void *new_out, int cap_new)
    {
        RefCtx *c = (RefCtx *)p;
        cv::Mat image(c->h, c->w, CV_8UC1, (void *)gray, (size_t)c->w);
        cv::Mat dmap(c->h, c->w, CV_32FC1, (void *)depth, (size_t)c->w * 4);
        Eigen::Matrix4f pose;
        for (int i = 0; i < 16; i++) pose.d[i] = pose_colmajor16[i];
        c->local.resize(n_local);
        if (n_local) memcpy(c->local.data(), local, (size_t)n_local * sizeof(SurfelElement));
        c->ff.fuse_initialize_map(ref_idx, image, dmap, pose, c->local, c->fresh);
        if (n_local) memcpy(local, c->local.data(), (size_t)n_local * sizeof(SurfelElement));
        int n_new = (int)c->fresh.size();
        int n_copy = n_new < cap_new ? n_new : cap_new;
        if (n_copy) memcpy(new_out, c->fresh.data(), (size_t)n_copy * sizeof(SurfelElement));
        return n_new;
    }

    // only generate_super_pixels (fusion_functions.cpp:960-975): labels + seeds, no surfels
    void dsmref_superpixels(void *p, const uint8_t *gray, const float *depth)
    {
        RefCtx *c = (RefCtx *)p;
        c->ff.image = cv::Mat(c->h, c->w, CV_8UC1, (void *)gray, (size_t)c->w);
        c->ff.depth = cv::Mat(c->h, c->w, CV_32FC1, (void *)depth, (size_t)c->w * 4);
        c->ff.generate_super_pixels();
    }

    void dsmref_get_labels(void *p, int32_t *out)
    {
        RefCtx *c = (RefCtx *)p;
        memcpy(out, c->ff.superpixel_index.data(), c->ff.superpixel_index.size() * sizeof(int));
    }
    void dsmref_get_seeds(void *p, void *out)
    {
        RefCtx *c = (RefCtx *)p;
        memcpy(out, c->ff.superpixel_seeds.data(), c->ff.superpixel_seeds.size() * sizeof(Superpixel_seed));
    }
    void dsmref_get_norm_map(void *p, float *out)
    {
        RefCtx *c = (RefCtx *)p;
        memcpy(out, c->ff.norm_map.data(), c->ff.norm_map.size() * sizeof(float));
    }
    int dsmref_num_seeds(void *p) { return (int)((RefCtx *)p)->ff.superpixel_seeds.size(); }
    int dsmref_is_serial(void)
    {
#ifdef DSM_REF_SERIAL
        return 1;
#else
        return 0;
#endif
    }
}
