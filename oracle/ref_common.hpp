// TEST INFRASTRUCTURE ONLY.  Shared prologue of the translation units that #include reference sources for the
// SurfelMap oracle (ref_map_driver.cpp, ref_ff_tu.cpp): every std::thread the reference starts runs on the calling
// thread, in join order (deterministic; same idea as the stand-in of oracle/ref_driver.cpp), and the reference's
// progress printf lines are dropped.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <set>
#include <thread>
#include <utility>
#include <vector>

namespace std
{
// Stand-in for std::thread whose body runs on the CALLING thread when it is joined (join order = creation order in
// the reference).  Running at join() rather than at construction matters for SurfelMap::warp_surfels
// (surfel_map.cpp:795-824): it starts the inactive-surfel workers, THEN reads cam_pose / loop_pose to build the warp
// of the active surfels, and the workers overwrite cam_pose -- a data race in the reference that the real scheduler
// practically always resolves in favour of the main thread.  Deferring to join() reproduces exactly that legal
// schedule, deterministically.
class dsm_inline_thread
{
  public:
    dsm_inline_thread() {}
    template <typename Obj, typename... MArgs, typename... Args>
    dsm_inline_thread(void (Obj::*fn)(MArgs...), Obj *self, Args... args) : body_([=]() { (self->*fn)(args...); })
    {
    }
    dsm_inline_thread(dsm_inline_thread &&o) : body_(std::move(o.body_)) { o.body_ = nullptr; }
    dsm_inline_thread &operator=(dsm_inline_thread &&o)
    {
        body_ = std::move(o.body_);
        o.body_ = nullptr;
        return *this;
    }
    bool joinable() const { return (bool)body_; }
    void join()
    {
        if (body_)
        {
            std::function<void()> f = std::move(body_);
            body_ = nullptr;
            f();
        }
    }

  private:
    std::function<void()> body_;
};
} // namespace std
#define thread dsm_inline_thread

#include <Eigen/Eigen>        // stand-in
#include <opencv2/opencv.hpp> // stand-in

static inline int dsm_ref_noprintf(const char *, ...) { return 0; }
#define printf dsm_ref_noprintf
