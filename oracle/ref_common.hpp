// TEST INFRASTRUCTURE ONLY.  Shared prologue of the translation units that #include reference sources for the
// SurfelMap oracle (ref_map_driver.cpp, ref_ff_tu.cpp): every std::thread the reference starts runs inline, in
// creation order, on the calling thread (deterministic; same stand-in idea as oracle/ref_driver.cpp), and the
// reference's progress printf lines are dropped.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <set>
#include <thread>
#include <utility>
#include <vector>

namespace std
{
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

#include <Eigen/Eigen>        // stand-in
#include <opencv2/opencv.hpp> // stand-in

static inline int dsm_ref_noprintf(const char *, ...) { return 0; }
#define printf dsm_ref_noprintf
