// Wire formats on the output side of the hot path: the two files the reference's
// ~save_map callback writes (ros_node.cpp -> SurfelMap::save_cloud / save_mesh).  Host code only — text
// formatting is not GPU work; the device part is the filtered, ordered export of the resident pool
// (dsm_pool_export_cloud / dsm_pool_export_surfels in dsm_capi.cu).
//
//  dsm_write_pcd       pcl::io::savePCDFile(name, cloud) as called at surfel_map.cpp:1171 for a
//                      pcl::PointCloud<pcl::PointXYZI>: PCD v0.7, fields x y z intensity, one float each.
//                      PCL is a third-party dependency that is not vendored in the reference tree; the layout
//                      below restates its published PCD v0.7 writer (ASCII: 8 significant digits, "nan" for NaN,
//                      single spaces, '\n' line ends; binary: the packed 16-byte records after "DATA binary").
//  dsm_write_ply_mesh  SurfelMap::save_mesh (surfel_map.cpp:1229-1280) with push_a_surfel (:1175-1226): every
//                      surfel becomes a hexagon of 6 vertices (grey colour) and 4 triangles, ASCII PLY, numbers
//                      in the default ostream format (%g), each vertex value followed by one space.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/dsm.h"

namespace
{
struct File
{
    FILE *f;
    explicit File(const char *path) : f(std::fopen(path, "wb")) {}
    ~File()
    {
        if (f) std::fclose(f);
    }
    bool close()
    {
        const bool ok = f && std::fflush(f) == 0 && !std::ferror(f);
        if (f) std::fclose(f);
        f = nullptr;
        return ok;
    }
};

// one PCD ASCII value: ostream << float at precision 8, NaN spelled "nan" whatever its sign
inline int put_pcd_value(char *dst, float v)
{
    if (std::isnan(v))
    {
        std::memcpy(dst, "nan", 3);
        return 3;
    }
    return std::snprintf(dst, 32, "%.8g", (double)v);
}

// push_a_surfel (surfel_map.cpp:1175-1226): 6 vertices x (x, y, z, c, c, c); float arithmetic in the
// reference's operation order (this file is built with -ffp-contract=off)
void hexagon(const dsm_surfel_t &e, float *v)
{
    const float c = (float)(int)e.color; // "int surfel_color = this_surfel.color" (:1177), pushed back as float
    const float px = e.px, py = e.py, pz = e.pz;
    float xd[3] = {-1 * e.ny, e.nx, 0.f}; // :1186-1188
    {                                     // x_dir.normalize() (:1189): Eigen leaves a zero vector untouched
        const float z = xd[0] * xd[0] + xd[1] * xd[1] + xd[2] * xd[2];
        if (z > 0.f)
        {
            const float n = std::sqrt(z);
            xd[0] /= n, xd[1] /= n, xd[2] /= n;
        }
    }
    const float yd[3] = {e.ny * xd[2] - e.nz * xd[1], e.nz * xd[0] - e.nx * xd[2], e.nx * xd[1] - e.ny * xd[0]}; // :1191
    const float radius = e.size;
    const float h_r = (float)((double)radius * 0.5);     // :1193
    const float t_r = (float)((double)radius * 0.86603); // :1194
    const float p[3] = {px, py, pz};
    const float sx[6] = {-h_r, h_r, -radius, radius, -h_r, h_r}; // :1196-1201
    const float sy[6] = {-t_r, -t_r, 0.f, 0.f, t_r, t_r};
    for (int k = 0; k < 6; k++)
    {
        for (int a = 0; a < 3; a++)
        {
            float q = p[a] + xd[a] * sx[k]; // p -/+ x_dir * r  (x*(-r) == -(x*r) exactly)
            if (k != 2 && k != 3) q = q + yd[a] * sy[k];
            v[k * 6 + a] = q;
        }
        v[k * 6 + 3] = v[k * 6 + 4] = v[k * 6 + 5] = c;
    }
}
} // namespace

extern "C" int dsm_write_pcd(const char *path, const dsm_point_t *pts, size_t n, int binary)
{
    if (!path || (n > 0 && !pts)) return DSM_E_INVALID;
    File out(path);
    if (!out.f) return DSM_E_IO;
    std::fprintf(out.f,
                 "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                 "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA %s\n",
                 n, n, binary ? "binary" : "ascii");
    if (binary)
    {
        if (n > 0 && std::fwrite(pts, sizeof(dsm_point_t), n, out.f) != n) return DSM_E_IO;
    }
    else
    {
        char line[160];
        for (size_t i = 0; i < n; i++)
        {
            const float v[4] = {pts[i].x, pts[i].y, pts[i].z, pts[i].intensity};
            int len = 0;
            for (int k = 0; k < 4; k++)
            {
                len += put_pcd_value(line + len, v[k]);
                line[len++] = k < 3 ? ' ' : '\n';
            }
            if (std::fwrite(line, 1, (size_t)len, out.f) != (size_t)len) return DSM_E_IO;
        }
    }
    return out.close() ? DSM_OK : DSM_E_IO;
}

extern "C" int dsm_mesh_vertices(const dsm_surfel_t *surfels, size_t n, float *vertices36)
{
    if (n > 0 && (!surfels || !vertices36)) return DSM_E_INVALID;
    for (size_t i = 0; i < n; i++) hexagon(surfels[i], vertices36 + i * 36);
    return DSM_OK;
}

extern "C" int dsm_write_ply_mesh(const char *path, const dsm_surfel_t *surfels, size_t n)
{
    if (!path || (n > 0 && !surfels)) return DSM_E_INVALID;
    File out(path);
    if (!out.f) return DSM_E_IO; // the reference returns silently when the stream cannot be opened (:1232-1233)
    std::fprintf(out.f,
                 "ply\nformat ascii 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %zu\n"
                 "property list uchar int vertex_index\nend_header\n",
                 n * 6, n * 4);
    float v[36];
    char line[6 * 24 + 8];
    for (size_t i = 0; i < n; i++)
    {
        hexagon(surfels[i], v);
        for (int k = 0; k < 6; k++)
        {
            int len = 0;
            for (int j = 0; j < 6; j++)
            {
                len += std::snprintf(line + len, 24, "%g", (double)v[k * 6 + j]);
                line[len++] = ' ';
            }
            line[len++] = '\n';
            if (std::fwrite(line, 1, (size_t)len, out.f) != (size_t)len) return DSM_E_IO;
        }
    }
    for (size_t i = 0; i < n; i++)
    { // :1266-1278
        const size_t p1 = i * 6, p2 = p1 + 1, p3 = p1 + 2, p4 = p1 + 3, p5 = p1 + 4, p6 = p1 + 5;
        std::fprintf(out.f, "3 %zu %zu %zu\n3 %zu %zu %zu\n3 %zu %zu %zu\n3 %zu %zu %zu\n", p1, p2, p3, p2, p4, p3, p3, p4, p5,
                     p5, p4, p6);
    }
    return out.close() ? DSM_OK : DSM_E_IO;
}
