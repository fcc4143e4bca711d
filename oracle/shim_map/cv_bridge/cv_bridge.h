// TEST INFRASTRUCTURE ONLY: cv_bridge::toCvCopy over the cv::Mat stand-in (a deep copy of the message payload).
#pragma once
#include <dsm_ros_shim.hpp>
#include <opencv2/opencv.hpp>
namespace cv_bridge
{
struct CvImage
{
    std_msgs::Header header;
    std::string encoding;
    cv::Mat image;
};
typedef std::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr &src, const std::string &encoding)
{
    CvImagePtr out(new CvImage);
    out->header = src->header;
    out->encoding = encoding;
    const int type = (encoding == sensor_msgs::image_encodings::MONO8) ? CV_8UC1 : CV_32FC1;
    out->image = cv::Mat((int)src->height, (int)src->width, type);
    const size_t row = (size_t)src->width * (type == CV_8UC1 ? 1 : 4);
    for (uint32_t r = 0; r < src->height; r++) std::memcpy(out->image.data + (size_t)r * out->image.step, src->data.data() + (size_t)r * src->step, row);
    return out;
}
} // namespace cv_bridge
