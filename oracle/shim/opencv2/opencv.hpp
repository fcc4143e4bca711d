// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for the handful of OpenCV
// names the reference's fusion_functions.cpp touches: cv::Mat (non-owning 2-D view with
// at<T>(row, col)), cv::Vec3b, and no-op imshow/waitKey used only by the dead debug_show().
// Written from scratch for this repo; not OpenCV code.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <iostream>
#include <vector>
#include <string>
#include <algorithm>
#include <memory>

typedef unsigned char uchar;

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5

namespace cv
{
struct Vec3b
{
    uchar v[3];
    Vec3b() { v[0] = v[1] = v[2] = 0; }
    Vec3b(uchar a, uchar b, uchar c) { v[0] = a; v[1] = b; v[2] = c; }
    uchar &operator[](int i) { return v[i]; }
    const uchar &operator[](int i) const { return v[i]; }
};

class Mat
{
  public:
    int rows, cols;
    size_t step; // bytes per row
    uchar *data;
    std::shared_ptr<std::vector<uchar> > owned;

    Mat() : rows(0), cols(0), step(0), data(NULL) {}
    Mat(int r, int c, int type) : rows(r), cols(c)
    {
        size_t elem = (type == CV_8UC1) ? 1 : (type == CV_8UC3) ? 3 : 4;
        step = elem * (size_t)c;
        owned.reset(new std::vector<uchar>(step * (size_t)r));
        data = owned->data();
    }
    // non-owning view over caller memory (what cv_bridge hands the reference)
    Mat(int r, int c, int /*type*/, void *ptr, size_t step_bytes)
        : rows(r), cols(c), step(step_bytes), data((uchar *)ptr) {}

    template <typename T>
    T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T>
    const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
};

inline void imshow(const char *, const Mat &) {}
inline int waitKey(int) { return 0; }
} // namespace cv
