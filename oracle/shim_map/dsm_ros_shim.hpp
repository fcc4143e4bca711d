// TEST INFRASTRUCTURE ONLY (oracle build of the reference's SurfelMap, oracle/ref_map_driver.cpp).
// Minimal stand-ins, written from scratch for this repo, for the ROS / PCL / boost / cv_bridge names that
// surfel_fusion/src/surfel_map.{h,cpp} touch: message PODs with the fields the file reads and writes, a
// NodeHandle whose parameters come from a process-wide table, and Publishers that keep the last message they were
// given so that a test can look at what the node would have put on its topics.  Not ROS / PCL / boost code.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <typeindex>
#include <vector>

namespace boost
{
template <class T>
using shared_ptr = std::shared_ptr<T>;
}

namespace ros
{
struct Time
{
    uint32_t sec, nsec;
    Time() : sec(0), nsec(0) {}
    explicit Time(double t) : sec((uint32_t)t), nsec((uint32_t)((t - (double)(uint32_t)t) * 1e9 + 0.5)) {}
    double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
    uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + nsec; }
    bool operator==(const Time &o) const { return sec == o.sec && nsec == o.nsec; }
    bool operator<(const Time &o) const { return sec < o.sec || (sec == o.sec && nsec < o.nsec); }
    static Time now() { return Time(); }
};

// what the last publish() on a topic carried (type-erased; the driver knows the type per topic)
struct Latch
{
    std::shared_ptr<void> msg;
    int count;
    Latch() : count(0) {}
};
inline std::map<std::string, Latch> &latches()
{
    static std::map<std::string, Latch> m;
    return m;
}

class Publisher
{
  public:
    std::string topic;
    template <class M>
    void publish(const M &m) const
    {
        Latch &l = latches()[topic];
        l.msg = std::make_shared<M>(m);
        l.count++;
    }
    template <class M>
    void publish(const std::shared_ptr<M> &m) const
    {
        Latch &l = latches()[topic];
        l.msg = std::make_shared<M>(*m);
        l.count++;
    }
};

inline std::map<std::string, double> &params()
{
    static std::map<std::string, double> p;
    return p;
}

class NodeHandle
{
  public:
    template <class T>
    bool getParam(const std::string &name, T &v) const
    {
        std::map<std::string, double>::const_iterator it = params().find(name);
        if (it == params().end()) return false;
        v = (T)it->second;
        return true;
    }
    template <class M>
    Publisher advertise(const std::string &topic, int)
    {
        Publisher p;
        p.topic = topic;
        return p;
    }
};
} // namespace ros

namespace std_msgs
{
struct Header
{
    uint32_t seq;
    ros::Time stamp;
    std::string frame_id;
    Header() : seq(0) {}
};
struct String
{
    std::string data;
};
typedef std::shared_ptr<const String> StringConstPtr;
struct ColorRGBA
{
    float r, g, b, a;
    ColorRGBA() : r(0), g(0), b(0), a(0) {}
};
} // namespace std_msgs

namespace geometry_msgs
{
struct Point
{
    double x, y, z;
    Point() : x(0), y(0), z(0) {}
};
struct Point32
{
    float x, y, z;
    Point32() : x(0), y(0), z(0) {}
};
struct Quaternion
{
    double x, y, z, w;
    Quaternion() : x(0), y(0), z(0), w(0) {}
};
struct Vector3
{
    double x, y, z;
    Vector3() : x(0), y(0), z(0) {}
};
struct Pose
{
    Point position;
    Quaternion orientation;
};
struct PoseStamped
{
    std_msgs::Header header;
    Pose pose;
};
struct PointStamped
{
    std_msgs::Header header;
    Point point;
};
typedef std::shared_ptr<const PointStamped> PointStampedConstPtr;
struct PoseWithCovariance
{
    Pose pose;
    double covariance[36];
    PoseWithCovariance()
    {
        for (int i = 0; i < 36; i++) covariance[i] = 0;
    }
};
} // namespace geometry_msgs

namespace sensor_msgs
{
namespace image_encodings
{
const std::string MONO8 = "mono8";
const std::string TYPE_32FC1 = "32FC1";
} // namespace image_encodings
struct Image
{
    std_msgs::Header header;
    uint32_t height, width, step;
    std::string encoding;
    std::vector<uint8_t> data;
    Image() : height(0), width(0), step(0) {}
};
typedef std::shared_ptr<const Image> ImageConstPtr;
struct ChannelFloat32
{
    std::string name;
    std::vector<float> values;
};
struct PointCloud
{
    std_msgs::Header header;
    std::vector<geometry_msgs::Point32> points;
    std::vector<ChannelFloat32> channels;
};
typedef std::shared_ptr<const PointCloud> PointCloudConstPtr;
typedef PointCloudConstPtr PointcloudConstPtr;
} // namespace sensor_msgs

namespace nav_msgs
{
struct Path
{
    std_msgs::Header header;
    std::vector<geometry_msgs::PoseStamped> poses;
};
typedef std::shared_ptr<const Path> PathConstPtr;
struct Odometry
{
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
};
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
} // namespace nav_msgs

namespace visualization_msgs
{
struct Marker
{
    enum
    {
        ARROW = 0,
        CUBE = 1,
        SPHERE = 2,
        CYLINDER = 3,
        LINE_STRIP = 4,
        LINE_LIST = 5,
        CUBE_LIST = 6,
        SPHERE_LIST = 7,
        POINTS = 8
    };
    enum
    {
        ADD = 0,
        MODIFY = 0,
        DELETE = 2,
        DELETEALL = 3
    };
    std_msgs::Header header;
    std::string ns;
    int32_t id, type, action;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    std::vector<geometry_msgs::Point> points;
    std::vector<std_msgs::ColorRGBA> colors;
    Marker() : id(0), type(0), action(0) {}
};
struct MarkerArray
{
    std::vector<Marker> markers;
};
} // namespace visualization_msgs

// ---- PCL: the PointXYZI cloud container the node publishes ----
namespace pcl
{
struct PCLHeader
{
    uint32_t seq;
    uint64_t stamp;
    std::string frame_id;
    PCLHeader() : seq(0), stamp(0) {}
};
struct PointXYZI
{
    float x, y, z, intensity;
    PointXYZI() : x(0), y(0), z(0), intensity(0) {}
};
template <class P>
class PointCloud
{
  public:
    typedef std::shared_ptr<PointCloud<P> > Ptr;
    typedef typename std::vector<P>::iterator iterator;
    typedef typename std::vector<P>::const_iterator const_iterator;
    PCLHeader header;
    std::vector<P> points;
    uint32_t width, height;
    PointCloud() : width(0), height(1) {}
    void push_back(const P &p)
    {
        points.push_back(p);
        width = (uint32_t)points.size();
    }
    size_t size() const { return points.size(); }
    void reserve(size_t n) { points.reserve(n); }
    iterator begin() { return points.begin(); }
    iterator end() { return points.end(); }
    P &front() { return points.front(); }
    P &back() { return points.back(); }
    P &at(size_t i) { return points.at(i); }
    P &operator[](size_t i) { return points[i]; }
    iterator erase(iterator a, iterator b)
    {
        iterator r = points.erase(a, b);
        width = (uint32_t)points.size();
        return r;
    }
    template <class It>
    void insert(iterator pos, It a, It b)
    {
        points.insert(pos, a, b);
        width = (uint32_t)points.size();
    }
    PointCloud &operator+=(const PointCloud &o)
    {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (uint32_t)points.size();
        return *this;
    }
};
namespace io
{
// files are not the subject of this oracle (the product's writers are tested against the published formats);
// the cloud handed to the writer is latched like a published message
template <class P>
int savePCDFile(const std::string &name, const PointCloud<P> &cloud)
{
    ros::Latch &l = ros::latches()["file:" + name];
    l.msg = std::make_shared<PointCloud<P> >(cloud);
    l.count++;
    return 0;
}
template <class P>
int savePLYFile(const std::string &name, const PointCloud<P> &cloud)
{
    return savePCDFile(name, cloud);
}
} // namespace io
} // namespace pcl

namespace pcl_conversions
{
inline void toPCL(const ros::Time &t, uint64_t &stamp) { stamp = t.toNSec() / 1000ull; }
} // namespace pcl_conversions
