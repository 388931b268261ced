// shim_pcl.h -- a minimal stand-in for the PCL types that /root/reference/include/RegisterPhotoICP.h
// names (pcl::getTime, point clouds and the GICP object of its alignGICP helper), so the header
// compiles here.   *** TEST INFRASTRUCTURE ONLY. ***   Nothing here is on the spherical dense
// registration path except pcl::getTime (a wall clock).
#pragma once
#include <cstdlib>
#include <iostream>
#include <vector>
#include <sys/time.h>
#include "shim_eigen.h"

namespace boost {
template <typename T>
class shared_ptr {   // just enough for PointCloud<T>::Ptr
    T* p;
    int* n;
public:
    shared_ptr() : p(0), n(0) {}
    explicit shared_ptr(T* q) : p(q), n(new int(1)) {}
    shared_ptr(const shared_ptr& o) : p(o.p), n(o.n) { if (n) ++*n; }
    shared_ptr& operator=(const shared_ptr& o) { if (o.n) ++*o.n; drop(); p = o.p; n = o.n; return *this; }
    ~shared_ptr() { drop(); }
    void drop() { if (n && --*n == 0) { delete p; delete n; } p = 0; n = 0; }
    void reset(T* q = 0) { drop(); if (q) { p = q; n = new int(1); } }
    T& operator*() const { return *p; }
    T* operator->() const { return p; }
    T* get() const { return p; }
    operator bool() const { return p != 0; }
};
}  // namespace boost

namespace pcl {
inline double getTime() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

struct PointXYZ { float x, y, z; PointXYZ() : x(0), y(0), z(0) {} };
struct PointXYZRGBA { float x, y, z; unsigned char r, g, b, a; unsigned rgba; PointXYZRGBA() : x(0), y(0), z(0), r(0), g(0), b(0), a(0), rgba(0) {} };

template <typename PointT>
struct PointCloud {
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT> points;
    unsigned width, height;
    bool is_dense;
    PointCloud() : width(0), height(0), is_dense(true) {}
    size_t size() const { return points.size(); }
    void resize(size_t n) { points.resize(n); }
    void clear() { points.clear(); }
    void push_back(const PointT& p) { points.push_back(p); }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    PointT& at(size_t i) { return points[i]; }
};

template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
    out.points.clear(); index.clear();
    for (size_t i = 0; i < in.points.size(); ++i)
        if (in.points[i].x == in.points[i].x && in.points[i].y == in.points[i].y && in.points[i].z == in.points[i].z) {
            out.points.push_back(in.points[i]);
            index.push_back((int)i);
        }
    out.width = (unsigned)out.points.size(); out.height = 1; out.is_dense = true;
}

template <typename PS, typename PT>
struct GeneralizedIterativeClosestPoint {   // alignGICP helper only (RPI.h:4800-4900); never called here
    static void na() { std::cerr << "refshim: pcl GICP is not implemented\n"; abort(); }
    void setMaxCorrespondenceDistance(double) {}
    void setMaximumIterations(int) {}
    void setTransformationEpsilon(double) {}
    void setRotationEpsilon(double) {}
    void setEuclideanFitnessEpsilon(double) {}
    void setRANSACOutlierRejectionThreshold(double) {}
    template <typename P> void setInputSource(const P&) {}
    template <typename P> void setInputTarget(const P&) {}
    template <typename C> void align(C&) { na(); }
    template <typename C, typename M> void align(C&, const M&) { na(); }
    bool hasConverged() const { return false; }
    double getFitnessScore() const { return 0; }
    Eigen::Matrix4f getFinalTransformation() const { return Eigen::Matrix4f::Identity(); }
};
}  // namespace pcl
