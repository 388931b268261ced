// synth.h -- deterministic synthetic sphere frames, shared host/device.
//
// A camera inside a convex axis-aligned box room (|x|<=1.5, |y|<=2.5, |z|<=3.0 m, x = up,
// the reference's sphere convention RPI.h:4580-4582).  Convex => no occlusion between any
// two viewpoints, all ranges in [1.0, 5.9] m lie inside (minDepth, maxDepth) = (0.3, 6.0).
// Pixel (r,c) looks along (sin phi, -cos phi sin theta, -cos phi cos theta) with
// phi = (H/2 - 1/2 - r) res, theta = c res, res = 2 PI / W  (RPI.h:4555-4582).
// Output dtypes are the reference's: RGB u8 x3 (R=G=B), depth u16 millimetres (Euclidean range).
// Same compile rules as sphere_math.h, so host and device render identical bytes.
#pragma once
#include "sphere_math.h"

#define R360_SYNTH_ODOMETRY 0
#define R360_SYNTH_KEYFRAMES 1

R360_HD uint64_t r360_splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
R360_HD double r360_u01(uint64_t* s) { return (double)(r360_splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

// Rodrigues rotation (double, pinned sin/cos), row-major R.
R360_HD void r360_rodrigues(const double w[3], double R[9]) {
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    double A, B;
    if (th2 < 1e-12) { A = 1.0; B = 0.5; }
    else { double s, c; r360_sincos(th, &s, &c); A = s / th; B = (1.0 - c) / th2; }
    double x = w[0], y = w[1], z = w[2];
    R[0] = 1.0 - B * (y * y + z * z); R[1] = B * x * y - A * z;         R[2] = B * x * z + A * y;
    R[3] = B * x * y + A * z;         R[4] = 1.0 - B * (x * x + z * z); R[5] = B * y * z - A * x;
    R[6] = B * x * z - A * y;         R[7] = B * y * z + A * x;         R[8] = 1.0 - B * (x * x + y * y);
}

// Camera-to-world pose of synthetic frame `id`: X_world = R X_cam + t.
R360_HD void r360_synth_pose(int kind, int id, double R[9], double t[3]) {
    double w[3];
    if (kind == R360_SYNTH_ODOMETRY) {
        double k = (double)id, s, c;
        r360_sincos(0.07 * k, &s, &c); t[0] = 0.3 * s;
        r360_sincos(0.10 * k, &s, &c); t[1] = 0.5 * s;
        r360_sincos(0.09 * k, &s, &c); t[2] = 0.5 * c;
        r360_sincos(0.05 * k, &s, &c); w[0] = 0.5 * s;
        r360_sincos(0.04 * k, &s, &c); w[1] = 0.05 * c;
        r360_sincos(0.03 * k, &s, &c); w[2] = 0.05 * s;
    } else {
        uint64_t st = 0x360ull + (uint64_t)id;
        t[0] = (2.0 * r360_u01(&st) - 1.0) * 0.5;
        t[1] = (2.0 * r360_u01(&st) - 1.0) * 1.0;
        t[2] = (2.0 * r360_u01(&st) - 1.0) * 1.25;
        w[0] = (2.0 * r360_u01(&st) - 1.0) * 3.14159265358979323846;   // yaw about x (up)
        w[1] = (2.0 * r360_u01(&st) - 1.0) * 0.05;
        w[2] = (2.0 * r360_u01(&st) - 1.0) * 0.05;
    }
    r360_rodrigues(w, R);
}

// Ground truth T_trg<-src (column-major 4x4): p_trg = Rt^T (Rs p_src + ts - tt).
R360_HD void r360_synth_relpose(int kind, int src_id, int trg_id, double T[16]) {
    double Rs[9], ts[3], Rt[9], tt[3];
    r360_synth_pose(kind, src_id, Rs, ts);
    r360_synth_pose(kind, trg_id, Rt, tt);
    double d[3] = { ts[0] - tt[0], ts[1] - tt[1], ts[2] - tt[2] };
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += Rt[3 * k + i] * Rs[3 * k + j];
            T[i + 4 * j] = a;
        }
        T[i + 12] = Rt[0 + i] * d[0] + Rt[3 + i] * d[1] + Rt[6 + i] * d[2];
        T[3 + 4 * i] = 0.0;
    }
    T[15] = 1.0;
}

// Procedural wall texture in [0.05, 0.95].
R360_HD float r360_synth_texture(float u, float v) {
    float a = r360_sinf(7.0f * u) * r360_sinf(5.0f * v);
    float b = r360_sinf(23.0f * u + 1.0f) * r360_sinf(19.0f * v + 2.0f);
    float c = r360_sinf(61.0f * u) * r360_sinf(53.0f * v);
    return 0.5f + 0.2f * a + 0.15f * b + 0.1f * c;
}

// One pixel: ray (sphi, -cphi*sth, -cphi*cth) in the camera frame of pose (R row-major
// float, t float) -> gray level and range in millimetres.
R360_HD void r360_synth_pixel(const float* R, const float* t, float sphi, float cphi, float sth,
                              float cth, uint8_t* gray, uint16_t* depth_mm) {
    const float dc0 = sphi, dc1 = -cphi * sth, dc2 = -cphi * cth;
    const float d[3] = { R[0] * dc0 + R[1] * dc1 + R[2] * dc2,
                         R[3] * dc0 + R[4] * dc1 + R[5] * dc2,
                         R[6] * dc0 + R[7] * dc1 + R[8] * dc2 };
    const float bound[3] = { 1.5f, 2.5f, 3.0f };
    float best = 1e30f;
    int wall = 0;
    for (int i = 0; i < 3; ++i) {
        if (d[i] > 1e-9f) {
            float s = (bound[i] - t[i]) / d[i];
            if (s < best) { best = s; wall = 2 * i; }
        } else if (d[i] < -1e-9f) {
            float s = (-bound[i] - t[i]) / d[i];
            if (s < best) { best = s; wall = 2 * i + 1; }
        }
    }
    const float h[3] = { t[0] + best * d[0], t[1] + best * d[1], t[2] + best * d[2] };
    const int ax = wall >> 1;
    float u = h[(ax + 1) % 3] + 1.7f * (float)wall;
    float v = h[(ax + 2) % 3] + 0.9f * (float)wall;
    float I = r360_synth_texture(u, v);
    float g = roundf(255.0f * I);
    g = fminf(fmaxf(g, 0.0f), 255.0f);
    *gray = (uint8_t)(int)g;
    float mm = roundf(best * 1000.0f);
    mm = fminf(fmaxf(mm, 0.0f), 65535.0f);
    *depth_mm = (uint16_t)(int)mm;
}
