// TEST INFRASTRUCTURE ONLY.  The reference's own SurfelMap (surfel_fusion/src/surfel_map.cpp) compiled where it lies,
// against the stand-in headers in oracle/shim_map + oracle/shim, behind a small C interface.  Two builds
// (oracle/Makefile): libdsm_refmap.so keeps the reference's FusionFunctions (the CPU hot path: a full-system oracle);
// libdsm_refmap_b200.so is the SAME translation unit with INTEGRATION.md's three-line patch applied through the
// include path (oracle/shim_map_b200/fusion_functions.h: FusionFunctions := dsm::FusionFunctions), i.e. the
// reference's node logic running on top of the product library -- the drop-in claim, executed.
//
// Nothing of the reference is copied: the sources are #included from $(REF_ROOT) at build time.
#include <map>
#include <string>
#include "ref_common.hpp"

#define private public
#include DSM_REFMAP_SOURCE
#undef private
#undef printf

// CameraPoseVisualization.cpp is not part of this build (markers only): empty bodies for what surfel_map.cpp calls
CameraPoseVisualization::CameraPoseVisualization(float, float, float, float) : m_scale(0), m_line_width(0) {}
void CameraPoseVisualization::setImageBoundaryColor(float, float, float, float) {}
void CameraPoseVisualization::setOpticalCenterConnectorColor(float, float, float, float) {}
void CameraPoseVisualization::setScale(double) {}
void CameraPoseVisualization::setLineWidth(double) {}
void CameraPoseVisualization::add_pose(const Eigen::Vector3d &, const Eigen::Quaterniond &) {}
void CameraPoseVisualization::reset() {}
void CameraPoseVisualization::publish_by(ros::Publisher &, ros::Time &) {}
void CameraPoseVisualization::add_edge(const Eigen::Vector3d &, const Eigen::Vector3d &) {}
void CameraPoseVisualization::add_loopedge(const Eigen::Vector3d &, const Eigen::Vector3d &) {}

struct RefMap
{
    ros::NodeHandle nh;
    SurfelMap *map;
};

static sensor_msgs::ImageConstPtr make_image(double stamp, int W, int H, const void *data, bool is_depth)
{
    std::shared_ptr<sensor_msgs::Image> m(new sensor_msgs::Image);
    m->header.stamp = ros::Time(stamp);
    m->height = (uint32_t)H, m->width = (uint32_t)W;
    m->step = (uint32_t)(W * (is_depth ? 4 : 1));
    m->encoding = is_depth ? "32FC1" : "mono8";
    m->data.assign((const uint8_t *)data, (const uint8_t *)data + (size_t)m->step * H);
    return m;
}

extern "C"
{
    // parameters of surfel_map.cpp:14-29 (kitti_orb.launch)
    RefMap *dsmmap_create(int W, int H, float fx, float fy, float cx, float cy, float far_dist, float near_dist, int drift_free_poses)
    {
        std::map<std::string, double> &p = ros::params();
        p["cam_width"] = W, p["cam_height"] = H, p["cam_fx"] = fx, p["cam_fy"] = fy, p["cam_cx"] = cx, p["cam_cy"] = cy;
        p["fuse_far_distence"] = far_dist, p["fuse_near_distence"] = near_dist, p["drift_free_poses"] = drift_free_poses;
        RefMap *r = new RefMap;
        r->map = new SurfelMap(r->nh);
        return r;
    }
    void dsmmap_destroy(RefMap *r)
    {
        if (!r) return;
        delete r->map;
        delete r;
    }
    // the three callbacks of ros_node.cpp, in the order the node receives a synchronised frame:
    // orb_results_input(loop_stamps, loop_path, this_pose) then image_input / depth_input with the same stamp.
    // pose = {px,py,pz,qx,qy,qz,qw}; path = n_path of those; loops = n_loops pairs of keyframe indices;
    // is_keyframe / reference_index travel in pose.covariance[0] / [1] like the ORB node sends them.
    void dsmmap_frame(RefMap *r, double stamp, const unsigned char *gray, const float *depth, const double *pose7, int is_keyframe,
                      int reference_index, const double *path7, int n_path, const int *loops, int n_loops)
    {
        const int W = r->map->cam_width, H = r->map->cam_height;
        std::shared_ptr<sensor_msgs::PointCloud> ls(new sensor_msgs::PointCloud);
        ls->header.stamp = ros::Time(stamp);
        ls->channels.resize(1);
        for (int i = 0; i < 2 * n_loops; i++) ls->channels[0].values.push_back((float)loops[i]);
        std::shared_ptr<nav_msgs::Path> path(new nav_msgs::Path);
        path->header.stamp = ros::Time(stamp);
        for (int i = 0; i < n_path; i++)
        {
            geometry_msgs::PoseStamped ps;
            const double *q = path7 + 7 * i;
            ps.pose.position.x = q[0], ps.pose.position.y = q[1], ps.pose.position.z = q[2];
            ps.pose.orientation.x = q[3], ps.pose.orientation.y = q[4], ps.pose.orientation.z = q[5], ps.pose.orientation.w = q[6];
            path->poses.push_back(ps);
        }
        std::shared_ptr<nav_msgs::Odometry> od(new nav_msgs::Odometry);
        od->header.stamp = ros::Time(stamp);
        od->pose.pose.position.x = pose7[0], od->pose.pose.position.y = pose7[1], od->pose.pose.position.z = pose7[2];
        od->pose.pose.orientation.x = pose7[3], od->pose.pose.orientation.y = pose7[4], od->pose.pose.orientation.z = pose7[5];
        od->pose.pose.orientation.w = pose7[6];
        od->pose.covariance[0] = is_keyframe ? 1.0 : 0.0;
        od->pose.covariance[1] = (double)reference_index;
        r->map->orb_results_input(ls, path, od);
        if (!gray || !depth) return; // pose feed only (lets a test look at the state between the callbacks)
        r->map->image_input(make_image(stamp, W, H, gray, false));
        r->map->depth_input(make_image(stamp, W, H, depth, true));
    }
    int dsmmap_num_local(RefMap *r) { return (int)r->map->local_surfels.size(); }
    void dsmmap_get_local(RefMap *r, SurfelElement *out) { std::memcpy(out, r->map->local_surfels.data(), r->map->local_surfels.size() * sizeof(SurfelElement)); }
    int dsmmap_num_poses(RefMap *r) { return (int)r->map->poses_database.size(); }
    int dsmmap_num_attached(RefMap *r, int pose) { return (int)r->map->poses_database[(size_t)pose].attached_surfels.size(); }
    void dsmmap_get_attached(RefMap *r, int pose, SurfelElement *out)
    {
        const std::vector<SurfelElement> &v = r->map->poses_database[(size_t)pose].attached_surfels;
        std::memcpy(out, v.data(), v.size() * sizeof(SurfelElement));
    }
    int dsmmap_num_inactive_points(RefMap *r) { return (int)r->map->inactive_pointcloud->size(); }
    void dsmmap_get_inactive_points(RefMap *r, float *xyzi)
    {
        for (size_t i = 0; i < r->map->inactive_pointcloud->size(); i++)
        {
            const PointType &p = r->map->inactive_pointcloud->points[i];
            xyzi[4 * i] = p.x, xyzi[4 * i + 1] = p.y, xyzi[4 * i + 2] = p.z, xyzi[4 * i + 3] = p.intensity;
        }
    }
    // last cloud published on a topic ("active_pointcloud", "inactive_pointcloud", "pointcloud", "raw_pointcloud")
    int dsmmap_published_points(const char *topic, float *xyzi, int cap)
    {
        std::map<std::string, ros::Latch>::iterator it = ros::latches().find(topic);
        if (it == ros::latches().end() || !it->second.msg) return -1;
        const PointCloud *c = (const PointCloud *)it->second.msg.get();
        const int n = (int)c->size();
        for (int i = 0; i < n && i < cap; i++)
        {
            const PointType &p = c->points[(size_t)i];
            xyzi[4 * i] = p.x, xyzi[4 * i + 1] = p.y, xyzi[4 * i + 2] = p.z, xyzi[4 * i + 3] = p.intensity;
        }
        return n;
    }
    void dsmmap_save_mesh(RefMap *r, const char *path) { r->map->save_mesh(path); }
    void dsmmap_save_cloud(RefMap *r, const char *path) { r->map->save_cloud(path); }
    // the individual members, for unit-level pinning of the restatements in dsm_oracle.c / pyoracle.py
    void dsmmap_set_local(RefMap *r, const SurfelElement *s, int n) { r->map->local_surfels.assign(s, s + n); }
    void dsmmap_warp_active(RefMap *r, const float *W_colmajor)
    {
        Eigen::Matrix4f m;
        for (int i = 0; i < 16; i++) m.d[i] = W_colmajor[i];
        r->map->warp_active_surfels_cpu_kernel(0, 1, m);
    }
    void dsmmap_fuse_map(RefMap *r, const unsigned char *gray, const float *depth, const float *pose_colmajor, int reference_index)
    {
        const int W = r->map->cam_width, H = r->map->cam_height;
        cv::Mat image(H, W, CV_8UC1, (void *)gray, (size_t)W), dm(H, W, CV_32FC1, (void *)depth, (size_t)W * 4);
        Eigen::Matrix4f m;
        for (int i = 0; i < 16; i++) m.d[i] = pose_colmajor[i];
        r->map->fuse_map(image, dm, m, reference_index);
    }
    void dsmmap_move_add_surfels(RefMap *r, int reference_index) { r->map->move_add_surfels(reference_index); }
    void dsmmap_publish_clouds(RefMap *r, int reference_index)
    { // the four cloud publishers, including the two the node has commented out of its callback
        ros::Time t;
        r->map->publish_active_pointcloud(t);
        r->map->publish_inactive_pointcloud(t);
        r->map->publish_all_pointcloud(t);
        r->map->pointcloud_publish.topic = "neighbor_pointcloud"; // same publisher as "pointcloud": keep both observable
        r->map->publish_neighbor_pointcloud(t, reference_index);
        r->map->pointcloud_publish.topic = "pointcloud";
    }
    int dsmmap_local_pose_indexs(RefMap *r, int *out, int cap)
    {
        int n = 0;
        for (std::set<int>::iterator it = r->map->local_surfels_indexs.begin(); it != r->map->local_surfels_indexs.end(); ++it, ++n)
            if (n < cap) out[n] = *it;
        return n;
    }
    int dsmmap_mesh_vertices(RefMap *r, const SurfelElement *s, int n, float *out36)
    {
        std::vector<float> v;
        for (int i = 0; i < n; i++)
        {
            SurfelElement e = s[i];
            r->map->push_a_surfel(v, e);
        }
        std::memcpy(out36, v.data(), v.size() * sizeof(float));
        return (int)v.size();
    }
}
