// oracle/ref_kd_driver.cc -- TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's own kd-tree (nanoflann, vendored header-only under /root/reference/include/nanoflann/, included
// from there -- nothing is copied) with the call sequence of flannTranslatePointsToClosestInputCell
// (/root/reference/src/CDMInterpolator.cc:991-1062: point cloud on the unit sphere, KDTreeSingleIndexAdaptor<L2_Simple_Adaptor,
// ..., 3> with max leaf 12, radiusSearch with sorted results, first match wins).  CDMInterpolator.cc itself cannot be compiled
// here (Boost, libxml2, ...), so the ~40 lines around the nanoflann calls are restated; the tree, its pruning and the result
// set are the reference's.  Used to pin orc_coordkd (brute force) in tests/test_oracle_golden.py.
#include <nanoflann/nanoflann.hpp>

#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace {
template <typename T>
struct PointCloud { // the adaptor interface nanoflann asks for (CDMInterpolator.cc:952-989)
    struct Point {
        T x, y, z;
    };
    std::vector<Point> pts;
    inline size_t kdtree_get_point_count() const { return pts.size(); }
    inline T kdtree_distance(const T* p1, const size_t idx_p2, size_t) const
    {
        const T d0 = p1[0] - pts[idx_p2].x;
        const T d1 = p1[1] - pts[idx_p2].y;
        const T d2 = p1[2] - pts[idx_p2].z;
        return d0 * d0 + d1 * d1 + d2 * d2;
    }
    inline T kdtree_get_pt(const size_t idx, int dim) const { return dim == 0 ? pts[idx].x : (dim == 1 ? pts[idx].y : pts[idx].z); }
    template <class BBOX>
    bool kdtree_get_bbox(BBOX&) const { return false; }
};
} // namespace

extern "C" int ref_coordkd(double* px, double* py, size_t n, const double* lon, const double* lat, size_t nx, size_t ny, double max_dist_m)
{
    using namespace nanoflann;
    double maxDist = max_dist_m / 6371000.;
    PointCloud<double> cloud;
    cloud.pts.resize(nx * ny);
    for (size_t ix = 0; ix < nx; ix++) {
        for (size_t iy = 0; iy < ny; iy++) {
            const size_t pos = ix + iy * nx;
            if (!(std::isnan(lat[pos]) || std::isnan(lon[pos]))) {
                cloud.pts[pos].x = std::cos(lat[pos]) * std::cos(lon[pos]);
                cloud.pts[pos].y = std::cos(lat[pos]) * std::sin(lon[pos]);
                cloud.pts[pos].z = std::sin(lat[pos]);
            } else {
                cloud.pts[pos].x = cloud.pts[pos].y = cloud.pts[pos].z = NAN;
            }
        }
    }
    typedef KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<double, PointCloud<double> >, PointCloud<double>, 3> my_kd_tree_t;
    my_kd_tree_t index(3, cloud, KDTreeSingleIndexAdaptorParams(12));
    index.buildIndex();
    const double search_radius = maxDist * maxDist;
    nanoflann::SearchParams params;
    params.sorted = true;
    for (size_t i = 0; i < n; i++) {
        const double query_pt[3] = {std::cos(py[i]) * std::cos(px[i]), std::cos(py[i]) * std::sin(px[i]), std::sin(py[i])};
        std::vector<std::pair<size_t, double> > ret_matches;
        const size_t nMatches = index.radiusSearch(&query_pt[0], search_radius, ret_matches, params);
        if (nMatches > 0) {
            const size_t pos = ret_matches.at(0).first;
            px[i] = (double)(pos % nx);
            py[i] = (double)(pos / nx);
        } else {
            px[i] = -1000;
            py[i] = -1000;
        }
    }
    return 1;
}
