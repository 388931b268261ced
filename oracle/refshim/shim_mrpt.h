// shim_mrpt.h -- a minimal stand-in for the MRPT (1.x) types that
// /root/reference/include/RegisterPhotoICP.h and Miscellaneous.h name, so the reference header
// compiles here.   *** TEST INFRASTRUCTURE ONLY. ***   Written from scratch (C++98).
//
// The one MRPT call on the spherical path is (RPI.h:4697)
//     mrpt::poses::CPose3D::exp(CArrayNumeric<double,6>(v), /*pseudo_exponential=*/true)
//         .getHomogeneousMatrixVal()
// restated (MRPT 1.x: translation copied verbatim, R = Rodrigues(v[3:6]) in double, Taylor
// fall-backs for tiny angles) by r360_pseudo_exp_AB / r360_rodrigues_small of
// rgbd360_b200/csrc/gn_math.h -- here with glibc sin/cos, as a g++ build of MRPT would call.
// The full exponential (pseudo_exponential=false) adds the V matrix on the translation; it is only
// reached from the pinhole alignFrames path.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <string>
#include <vector>
#include "shim_eigen.h"
#include "shim_pcl.h"

#ifndef DEG2RAD
#define DEG2RAD(x) ((x) * 3.14159265358979323846 / 180.0)
#define RAD2DEG(x) ((x) * 180.0 / 3.14159265358979323846)
#endif

namespace mrpt {

inline std::string format(const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}

namespace utils { class CStream {}; }

namespace math {
template <typename T, int N>
struct CArrayNumeric : public Eigen::Matrix<T, N, 1> {
    CArrayNumeric() {}
    template <int R, int C>
    CArrayNumeric(const Eigen::Matrix<T, R, C>& m) : Eigen::Matrix<T, N, 1>(m) {}
};
template <int N>
struct CArrayDouble : public CArrayNumeric<double, N> {
    CArrayDouble() {}
    template <int R, int C>
    CArrayDouble(const Eigen::Matrix<double, R, C>& m) : CArrayNumeric<double, N>(m) {}
};
struct CMatrixDouble44 : public Eigen::Matrix<double, 4, 4> {
    CMatrixDouble44() {}
    template <typename U, int R, int C>
    CMatrixDouble44(const Eigen::Matrix<U, R, C>& m) : Eigen::Matrix<double, 4, 4>(m.template cast<double>()) {}
};
}  // namespace math

namespace poses {
class CPose3D {
    math::CMatrixDouble44 H;
public:
    CPose3D() { static_cast<Eigen::Matrix<double, 4, 4>&>(H) = Eigen::Matrix<double, 4, 4>::Identity(); }
    explicit CPose3D(const math::CMatrixDouble44& m) : H(m) {}
    static CPose3D exp(const math::CArrayNumeric<double, 6>& v, bool pseudo_exponential = false) {
        double vv[6], T[16], A, B;
        for (int i = 0; i < 6; ++i) vv[i] = v(i);
        const double th2 = vv[3] * vv[3] + vv[4] * vv[4] + vv[5] * vv[5];
        if (r360_rodrigues_small(th2, &A, &B)) {
            const double th = sqrt(th2), inv_th = 1.0 / th;
            A = sin(th) * inv_th;
            B = (1 - cos(th)) * (inv_th * inv_th);
        }
        r360_pseudo_exp_AB(vv, A, B, T);
        if (!pseudo_exponential) r360_exp_translation(vv, A, B, th2, T);     // t = V u (gn_math.h)
        CPose3D p;
        for (int i = 0; i < 16; ++i) p.H.data()[i] = T[i];
        return p;
    }
    math::CMatrixDouble44 getHomogeneousMatrixVal() const { return H; }
    math::CArrayDouble<3> ln_rotation() const {
        math::CArrayDouble<3> r;
        const double tr = H(0, 0) + H(1, 1) + H(2, 2);
        double c = 0.5 * (tr - 1);
        if (c > 1) c = 1;
        if (c < -1) c = -1;
        const double th = acos(c);
        const double k = th < 1e-9 ? 0.5 : th / (2 * sin(th));
        r(0) = k * (H(2, 1) - H(1, 2));
        r(1) = k * (H(0, 2) - H(2, 0));
        r(2) = k * (H(1, 0) - H(0, 1));
        return r;
    }
};
}  // namespace poses

namespace pbmap {
struct Plane {
    Eigen::Vector3f v3normal, v3center;
    float areaHull;
    pcl::PointCloud<pcl::PointXYZRGBA>::Ptr polygonContourPtr;
    Plane() : areaHull(0), polygonContourPtr(new pcl::PointCloud<pcl::PointXYZRGBA>()) {}
};
struct PbMap {
    std::vector<Plane> vPlanes;
};
}  // namespace pbmap

}  // namespace mrpt

namespace Eigen {
template <typename T>
struct Quaternion {   // Miscellaneous.h:127-136 only (diffRotation)
    T w, x, y, z;
    Quaternion() : w(1), x(0), y(0), z(0) {}
    explicit Quaternion(const Matrix<T, 3, 3>& m) {
        const T tr = m(0, 0) + m(1, 1) + m(2, 2);
        if (tr > 0) {
            T s = std::sqrt(tr + 1) * 2; w = s / 4; x = (m(2, 1) - m(1, 2)) / s; y = (m(0, 2) - m(2, 0)) / s; z = (m(1, 0) - m(0, 1)) / s;
        } else if (m(0, 0) > m(1, 1) && m(0, 0) > m(2, 2)) {
            T s = std::sqrt(1 + m(0, 0) - m(1, 1) - m(2, 2)) * 2; w = (m(2, 1) - m(1, 2)) / s; x = s / 4; y = (m(0, 1) + m(1, 0)) / s; z = (m(0, 2) + m(2, 0)) / s;
        } else if (m(1, 1) > m(2, 2)) {
            T s = std::sqrt(1 + m(1, 1) - m(0, 0) - m(2, 2)) * 2; w = (m(0, 2) - m(2, 0)) / s; x = (m(0, 1) + m(1, 0)) / s; y = s / 4; z = (m(1, 2) + m(2, 1)) / s;
        } else {
            T s = std::sqrt(1 + m(2, 2) - m(0, 0) - m(1, 1)) * 2; w = (m(1, 0) - m(0, 1)) / s; x = (m(0, 2) + m(2, 0)) / s; y = (m(1, 2) + m(2, 1)) / s; z = s / 4;
        }
    }
    T angularDistance(const Quaternion& o) const {
        T d = std::fabs(w * o.w + x * o.x + y * o.y + z * o.z);
        if (d > 1) d = 1;
        return 2 * std::acos(d);
    }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;
}  // namespace Eigen
