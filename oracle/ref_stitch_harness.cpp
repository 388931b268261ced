// ref_stitch_harness.cpp -- the reference's OWN stitching code compiled from where it lies, into
// oracle/_ref/librpi_ref_stitch.so (and the pinned-trig twin), to pin the ingest row.
//
// *** TEST INFRASTRUCTURE ONLY. ***
//
// What is compiled:
//   * /root/reference/include/Calib360.h, WHOLE and unmodified (#include): the camera matrix of the constructor
//     (Calib360.h:75-77) and loadExtrinsicCalibration (Calib360.h:122-131), which reads the reference's own
//     Calibration/Extrinsics/Rt_0N.txt and inverts them.
//   * Frame360::stitchSphericalImage and Frame360::stitchImage (Frame360.h:386-405, 1099-1148), VERBATIM: the
//     Makefile cuts the two member functions out of /root/reference/include/Frame360.h at build time (awk, from each
//     function's signature line to its closing brace) into oracle/_ref/frame360_stitch_members.inc -- a build product
//     in the git-ignored output directory, never committed -- and that text is #included inside the scaffold class
//     below.  Frame360.h as a whole cannot be compiled here: the rest of the class is PCL plane segmentation, MRPT
//     PbMap, CLAMS undistortion and boost serialization; the scaffold supplies exactly the
//     members the two functions touch: frameRGBD_[8] (getRGBImage / getDepthImage), sphereRGB, sphereDepth, calib.
// Third-party stand-ins: oracle/refshim (Eigen: block / product / inverse / loadFromTextFile; OpenCV: Mat, at<>, zeros;
// PCL: getTime, pcl_isfinite; mrpt::format).  Built as gnu++98 with a leaked `using namespace std`, like the
// registration harness: Frame360.h itself relies on both (unqualified `cout`, Frame360.h:208; Miscellaneous.h:120-124
// is not C++11) -- which is what makes `pow(float, 2)` / `sqrt` of Frame360.h:1141 the FLOAT overloads.
#include <stdint.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <numeric>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
using namespace std;

#ifdef REF_PINNED_MATH
// second build: the float sin / cos calls of stitchImage routed to the pinned sequences the GPU executes
#include "../rgbd360_b200/csrc/sphere_math.h"
inline float ref_sin(float x) { return r360_sinf(x); }
inline float ref_cos(float x) { return r360_cosf(x); }
inline double ref_sin(double x) { return std::sin(x); }
inline double ref_cos(double x) { return std::cos(x); }
#define sin ref_sin
#define cos ref_cos
#endif

#include "shim_mrpt.h"              // mrpt::format, Eigen and PCL stand-ins
#include <opencv2/opencv.hpp>
#ifndef pcl_isfinite
#define pcl_isfinite(x) std::isfinite(x)      // pcl/pcl_macros.h on Linux
#endif
#ifndef PROJECT_SOURCE_PATH
#define PROJECT_SOURCE_PATH "/root/reference"
#endif
#include "Miscellaneous.h"          // PI (Miscellaneous.h:44), as Frame360.h:47 includes it
#include "Calib360.h"

struct RefSensorFrame {             // the two accessors of CloudRGBD_Ext the stitch functions call
    cv::Mat rgb, depth;
    cv::Mat& getRGBImage() { return rgb; }
    cv::Mat& getDepthImage() { return depth; }
};

class Frame360 {
public:
    cv::Mat sphereRGB, sphereDepth;
    RefSensorFrame frameRGBD_[8];
    Calib360* calib;
    explicit Frame360(Calib360* c) : calib(c) {}
#include "_ref/frame360_stitch_members.inc"
};

extern "C" {

// sensor_rgb: 8 x h x w x 3 u8, sensor_depth: 8 x h x w u16.  extrinsics_dir: directory holding Rt_01.txt .. Rt_08.txt
// (NULL: the reference's own Calibration/Extrinsics), read by Calib360::loadExtrinsicCalibration; or Rt_inv_in != NULL:
// 8 column-major 4x4 matrices put straight into Calib360::Rt_inv.  Outputs: sphere images (rows x cols of
// stitchSphericalImage), the Rt_inv actually used (8 x 16, column-major) and the sphere size.
int refstitch_run(const uint8_t* sensor_rgb, const uint16_t* sensor_depth, int h, int w, const char* extrinsics_dir,
                  const float* Rt_inv_in, uint8_t* sphere_rgb, uint16_t* sphere_depth, float* Rt_inv_out, int* rows, int* cols) {
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    Calib360 calib;                                             // QVGA camera matrix
    if (Rt_inv_in) {
        for (int s = 0; s < 8; ++s)
            for (int k = 0; k < 16; ++k) calib.Rt_inv[s].data()[k] = Rt_inv_in[16 * s + k];
    } else {
        calib.loadExtrinsicCalibration(extrinsics_dir ? std::string(extrinsics_dir) : std::string(""));
    }
    Frame360 f(&calib);
    for (int s = 0; s < 8; ++s) {
        f.frameRGBD_[s].rgb = cv::Mat(h, w, CV_8UC3, (void*)(sensor_rgb + (size_t)s * h * w * 3)).clone();
        f.frameRGBD_[s].depth = cv::Mat(h, w, CV_16UC1, (void*)(sensor_depth + (size_t)s * h * w)).clone();
    }
    omp_set_dynamic(0);                                         // `#pragma omp parallel num_threads(8)`: one thread per sensor
    f.stitchSphericalImage();
    std::cout.rdbuf(old);
    if (rows) *rows = f.sphereRGB.rows;
    if (cols) *cols = f.sphereRGB.cols;
    if (sphere_rgb)
        for (int r = 0; r < f.sphereRGB.rows; ++r) memcpy(sphere_rgb + (size_t)r * f.sphereRGB.cols * 3, f.sphereRGB.ptr<uint8_t>(r), (size_t)f.sphereRGB.cols * 3);
    if (sphere_depth)
        for (int r = 0; r < f.sphereDepth.rows; ++r) memcpy(sphere_depth + (size_t)r * f.sphereDepth.cols, f.sphereDepth.ptr<uint16_t>(r), (size_t)f.sphereDepth.cols * 2);
    if (Rt_inv_out)
        for (int s = 0; s < 8; ++s)
            for (int k = 0; k < 16; ++k) Rt_inv_out[16 * s + k] = calib.Rt_inv[s].data()[k];
    return 0;
}

// Calib360's camera matrix (fx, fy, cx, cy) as its constructor sets it.
void refstitch_camera(float out[4]) {
    Calib360 calib;
    out[0] = calib.cameraMatrix(0, 0); out[1] = calib.cameraMatrix(1, 1); out[2] = calib.cameraMatrix(0, 2); out[3] = calib.cameraMatrix(1, 2);
}

}  // extern "C"
