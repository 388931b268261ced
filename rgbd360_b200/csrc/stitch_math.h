// stitch_math.h -- per-pixel arithmetic of Frame360::stitchImage (the ingest step right before the
// registration path; /root/reference/include/Frame360.h:1099-1148), shared host/device.
//
// A sphere pixel (row_phi, col_theta) of the (8*size_h) x (8*size_h*0.5*60/180) equirectangular
// image belongs to sensor 7 - col_theta / size_h.  Its unit ray is rotated into that sensor's
// frame by the inverse extrinsics, projected through the pinhole model (Calib360.h:75-77) and the
// NEAREST-BELOW sensor pixel (float -> int truncation of at<>(v, u)) supplies the colour; depth is
// the sensor's z-depth scaled to Euclidean range, truncated to u16.
// Same compile rules as sphere_math.h (no implicit contraction), so host and device agree bit for bit.
#pragma once
#include "sphere_math.h"

struct R360StitchGeom {
    int rows, cols;             // sphere image: cols = 8 * size_h, rows = (int)(cols * 0.5 * 60.0 / 180)
    int size_h, size_w;         // sensor image rows (240), cols (320)
    float offset_phi;           // rows / 2 - 0.5            Frame360.h:1104 (integer division, then - 0.5)
    float offset_theta;         // -size_h * 15 / 2 + 0.5    Frame360.h:1105 (integer division)
    float angle_pixel;          // 2 PI / cols               Frame360.h:1106
    float fx, fy, cx, cy;       // Calib360.h:75-77
};

R360_HD R360StitchGeom r360_stitch_geom(int size_h, int size_w, float fx, float fy, float cx, float cy) {
    R360StitchGeom g;
    g.size_h = size_h; g.size_w = size_w;
    g.cols = size_h * 8;                                   // Frame360.h:391
    g.rows = (int)(g.cols * 0.5 * 60.0 / 180);             // Frame360.h:392
    g.offset_phi = (float)(g.rows / 2 - 0.5);
    g.offset_theta = (float)(-size_h * 15 / 2 + 0.5);
    g.angle_pixel = (float)(2 * R360_PI_D / g.cols);
    g.fx = fx; g.fy = fy; g.cx = cx; g.cy = cy;
    return g;
}

// Sensor pixel hit by sphere pixel (row, col); sphi/cphi and sth/cth are sin/cos of
// phi = (offset_phi - row) * angle_pixel and theta = (col + offset_theta) * angle_pixel.
// Rt_inv: column-major 4x4 of the sensor.  Returns 0 when the ray misses the sensor image.
// *range_scale = sqrt(1 + ((u-cx)/fx)^2 + ((v-cy)/fy)^2) in FLOAT (Frame360.h:1141): the reference is C++98 code with
// a leaked `using namespace std` (unqualified `cout` in Frame360.h:208; Miscellaneous.h:120-124 does not compile as
// C++11), where `pow(float, 2)` is std::pow(float, int) = x * x in float and `sqrt` the float overload -- what the
// reference's own lines produce when compiled here (tests/test_ingest.py pins this against them).
R360_HD int r360_stitch_pixel(const R360StitchGeom& g, const float* Rt_inv, float sphi, float cphi, float sth,
                              float cth, int* ui, int* vi, float* range_scale) {
    const float v0 = sphi, v1 = cphi * sth, v2 = cphi * cth;
    // Eigen 3x3 * 3x1 + 3x1, coefficient sums left to right
    const float p0 = ((Rt_inv[0] * v0 + Rt_inv[4] * v1) + Rt_inv[8] * v2) + Rt_inv[12];
    const float p1 = ((Rt_inv[1] * v0 + Rt_inv[5] * v1) + Rt_inv[9] * v2) + Rt_inv[13];
    const float p2 = ((Rt_inv[2] * v0 + Rt_inv[6] * v1) + Rt_inv[10] * v2) + Rt_inv[14];
    const float u = g.fx * p0 / p2 + g.cx;
    const float v = g.fy * p1 / p2 + g.cy;
    if (!(u >= 0 && u < g.size_w && v >= 0 && v < g.size_h)) return 0;
    *ui = (int)u;
    *vi = (int)v;
    const float a = (u - g.cx) / g.fx, b = (v - g.cy) / g.fy;
    *range_scale = sqrtf((1 + a * a) + b * b);
    return 1;
}

// depth (u16 mm, z) -> Euclidean range (u16 mm): float product truncated (Frame360.h:1141);
// values beyond the u16 range saturate (the reference's conversion is undefined there).
R360_HD unsigned short r360_stitch_range(unsigned short d, float range_scale) {
    const float x = (float)d * range_scale;
    return (unsigned short)(x < 65535.0f ? x : 65535.0f);
}
