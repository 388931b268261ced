// gn_math.h -- the 6-DoF Gauss-Newton step algebra of alignFrames360, shared host/device.
//
// Restates, dependency-free, the third-party pieces the reference calls at
// include/RegisterPhotoICP.h:4682-4697:
//   * (H + lambda*diag(H)).rank()        MRPT Eigen plugin -> ColPivHouseholderQR::rank()
//   * -H.inverse() * g                   Eigen fixed 6x6 inverse (partial-pivot LU), float
//   * CPose3D::exp(v, pseudo_exponential=true)   MRPT 1.x: t = v[0:3] verbatim,
//                                         R = Rodrigues(v[3:6]) in double
//   * Matrix4f product                    exp(v).cast<float>() * pose_estim
// Same compile rules as sphere_math.h (no implicit contraction on either side).
#pragma once
#include "sphere_math.h"

// Column-major 4x4 float product C = A*B, summed k = 0..3 left to right
// (Eigen's packet product adds A.col(k)*B(k,j) in that order).
R360_HD void r360_mat4_mul(const float* A, const float* B, float* C) {
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float acc = A[i + 0] * B[0 + 4 * j];
            acc = acc + A[i + 4] * B[1 + 4 * j];
            acc = acc + A[i + 8] * B[2 + 4 * j];
            acc = acc + A[i + 12] * B[3 + 4 * j];
            C[i + 4 * j] = acc;
        }
}

// General 4x4 float inverse (column-major), Gauss-Jordan with partial pivoting on the augmented matrix: the
// restatement of Eigen's Matrix4f::inverse() used for the sensor extrinsics (poseCamRobot.inverse(), RPI.h:4922, 5104;
// Calib360.h:129).  Real Eigen uses a cofactor formula for fixed 4x4: third-party arithmetic, restated ONCE here for the
// reference stand-in, the oracle and the product alike.
R360_HD void r360_inverse4(const float* A, float* inv) {
    const int n = 4;
    float a[4 * 8];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i * 2 * n + j] = A[i + 4 * j]; a[i * 2 * n + n + j] = (i == j) ? 1.f : 0.f; }
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i) if (fabsf(a[i * 2 * n + k]) > fabsf(a[p * 2 * n + k])) p = i;
        if (p != k) for (int j = 0; j < 2 * n; ++j) { const float t = a[k * 2 * n + j]; a[k * 2 * n + j] = a[p * 2 * n + j]; a[p * 2 * n + j] = t; }
        const float d = 1.f / a[k * 2 * n + k];
        for (int j = 0; j < 2 * n; ++j) a[k * 2 * n + j] *= d;
        for (int i = 0; i < n; ++i)
            if (i != k) { const float f = a[i * 2 * n + k]; for (int j = 0; j < 2 * n; ++j) a[i * 2 * n + j] -= f * a[k * 2 * n + j]; }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) inv[i + 4 * j] = a[i * 2 * n + n + j];
}

// MRPT CPose3D::exp(..., pseudo_exponential = true): translation copied, rotation =
// Rodrigues formula (rodrigues_so3_exp).  A = sin(theta)/theta and B = (1-cos(theta))/theta^2
// come from r360_rodrigues_AB (Taylor fall-backs for tiny angles, as MRPT).  Output
// column-major 4x4 double.
R360_HD void r360_pseudo_exp_AB(const double v[6], double A, double B, double T[16]) {
    const double wx = v[3], wy = v[4], wz = v[5];
    const double wx2 = wx * wx, wy2 = wy * wy, wz2 = wz * wz;
    double R00 = 1.0 - B * (wy2 + wz2);
    double R11 = 1.0 - B * (wx2 + wz2);
    double R22 = 1.0 - B * (wx2 + wy2);
    double a = A * wz, b = B * (wx * wy);
    double R01 = b - a, R10 = b + a;
    a = A * wy; b = B * (wx * wz);
    double R02 = b + a, R20 = b - a;
    a = A * wx; b = B * (wy * wz);
    double R12 = b - a, R21 = b + a;
    T[0] = R00; T[1] = R10; T[2] = R20; T[3] = 0.0;
    T[4] = R01; T[5] = R11; T[6] = R21; T[7] = 0.0;
    T[8] = R02; T[9] = R12; T[10] = R22; T[11] = 0.0;
    T[12] = v[0]; T[13] = v[1]; T[14] = v[2]; T[15] = 1.0;
}
// Returns 1 when sin/cos of theta are needed (A, B then come from the caller's sin/cos).
R360_HD int r360_rodrigues_small(double theta_sq, double* A, double* B) {
    const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
    if (theta_sq < 1e-8) {
        *A = 1.0 - one_6th * theta_sq;
        *B = 0.5;
        return 0;
    }
    if (theta_sq < 1e-6) {
        *B = 0.5 - 0.25 * one_6th * theta_sq;
        *A = 1.0 - theta_sq * one_6th * (1.0 - one_20th * theta_sq);
        return 0;
    }
    return 1;
}
R360_HD void r360_pseudo_exp(const double v[6], double T[16]) {
    const double theta_sq = v[3] * v[3] + v[4] * v[4] + v[5] * v[5];
    double A, B;
    if (r360_rodrigues_small(theta_sq, &A, &B)) {
        const double theta = sqrt(theta_sq);
        double s, c;
        r360_sincos(theta, &s, &c);
        const double inv_theta = 1.0 / theta;
        A = s * inv_theta;
        B = (1 - c) * (inv_theta * inv_theta);
    }
    r360_pseudo_exp_AB(v, A, B, T);
}

// MRPT CPose3D::exp(v) with pseudo_exponential = false (what the pinhole alignFrames calls, RPI.h:4375):
// the rotation of the pseudo-exponential and the translation t = V u, V = I + B [w]x + C [w]x^2,
// C = (1 - A) / theta^2 (1/6 for tiny angles).  Overwrites the translation column of a pseudo-exp T.
R360_HD void r360_exp_translation(const double v[6], double A, double B, double theta_sq, double T[16]) {
    const double C = theta_sq < 1e-8 ? 1.0 / 6.0 : (1 - A) / theta_sq;
    const double w[3] = { v[3], v[4], v[5] }, u[3] = { v[0], v[1], v[2] };
    const double wu[3] = { w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0] };
    const double wwu[3] = { w[1] * wu[2] - w[2] * wu[1], w[2] * wu[0] - w[0] * wu[2], w[0] * wu[1] - w[1] * wu[0] };
    for (int i = 0; i < 3; ++i) T[12 + i] = u[i] + B * wu[i] + C * wwu[i];
}
R360_HD void r360_se3_exp(const double v[6], double T[16]) {
    const double theta_sq = v[3] * v[3] + v[4] * v[4] + v[5] * v[5];
    double A, B;
    if (r360_rodrigues_small(theta_sq, &A, &B)) {
        const double theta = sqrt(theta_sq);
        double s, c;
        r360_sincos(theta, &s, &c);
        const double inv_theta = 1.0 / theta;
        A = s * inv_theta;
        B = (1 - c) * (inv_theta * inv_theta);
    }
    r360_pseudo_exp_AB(v, A, B, T);
    r360_exp_translation(v, A, B, theta_sq, T);
}

// Partial-pivot LU inverse of a 6x6 float matrix (column-major, symmetric input so the
// layout is immaterial).  Returns 0 when a pivot is exactly zero (inverse undefined).
R360_HD int r360_inverse6(const float* Hin, float* inv) {
    float a[6][12];
    for (int i = 0; i < 36; ++i) inv[i] = NAN;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            a[i][j] = Hin[i + 6 * j];
            a[i][6 + j] = (i == j) ? 1.0f : 0.0f;
        }
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        float big = fabsf(a[k][k]);
        for (int i = k + 1; i < 6; ++i) {
            float v = fabsf(a[i][k]);
            if (v > big) { big = v; piv = i; }
        }
        if (!(big > 0.0f)) return 0;
        if (piv != k)
            for (int j = 0; j < 12; ++j) { float t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
        float d = 1.0f / a[k][k];
        for (int j = 0; j < 12; ++j) a[k][j] = a[k][j] * d;
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            float f = a[i][k];
            for (int j = 0; j < 12; ++j) a[i][j] = a[i][j] - f * a[k][j];
        }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) inv[i + 6 * j] = a[i][6 + j];
    return 1;
}

// update = -(H^-1) * g     (RPI.h:4693), row sums k = 0..5 left to right.
R360_HD void r360_solve_update(const float* inv, const float* g, float* upd) {
    for (int i = 0; i < 6; ++i) {
        float acc = (-inv[i + 0]) * g[0];
        for (int k = 1; k < 6; ++k) acc = acc + (-inv[i + 6 * k]) * g[k];
        upd[i] = acc;
    }
}

// rank() of a 6x6 float matrix by column-pivoted Householder QR with Eigen's default
// threshold: |R_ii| > eps * 6 * max|R_ii|   (ColPivHouseholderQR::rank()).
R360_HD int r360_rank6(const float* Min) {
    float a[6][6];
    float colnorm2[6];
    for (int j = 0; j < 6; ++j) {
        float s = 0.0f;
        for (int i = 0; i < 6; ++i) { a[i][j] = Min[i + 6 * j]; s = s + a[i][j] * a[i][j]; }
        colnorm2[j] = s;
    }
    float diag[6];
    float maxpivot = 0.0f;
    int nonzero = 6;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        float big = colnorm2[k];
        for (int j = k + 1; j < 6; ++j)
            if (colnorm2[j] > big) { big = colnorm2[j]; piv = j; }
        // recompute the winning column norm exactly on the remaining rows
        big = 0.0f;
        for (int i = k; i < 6; ++i) big = big + a[i][piv] * a[i][piv];
        if (big == 0.0f) { nonzero = k; for (int j = k; j < 6; ++j) diag[j] = 0.0f; break; }
        if (piv != k) {
            for (int i = 0; i < 6; ++i) { float t = a[i][k]; a[i][k] = a[i][piv]; a[i][piv] = t; }
            float t = colnorm2[k]; colnorm2[k] = colnorm2[piv]; colnorm2[piv] = t;
        }
        // Householder reflector for a[k..5][k]
        float norm = sqrtf(big);
        float alpha = a[k][k];
        float beta = (alpha >= 0.0f) ? -norm : norm;
        float tail2 = 0.0f;
        for (int i = k + 1; i < 6; ++i) tail2 = tail2 + a[i][k] * a[i][k];
        if (tail2 == 0.0f) {
            beta = alpha;                                 // already upper-triangular here
        } else {
            float v0 = alpha - beta;
            float vnorm2 = v0 * v0 + tail2;
            for (int j = k + 1; j < 6; ++j) {
                float dot = v0 * a[k][j];
                for (int i = k + 1; i < 6; ++i) dot = dot + a[i][k] * a[i][j];
                float f = 2.0f * dot / vnorm2;
                a[k][j] = a[k][j] - f * v0;
                for (int i = k + 1; i < 6; ++i) a[i][j] = a[i][j] - f * a[i][k];
            }
        }
        diag[k] = beta;
        if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
        for (int j = k + 1; j < 6; ++j) colnorm2[j] = colnorm2[j] - a[k][j] * a[k][j];
    }
    const float thr = maxpivot * (1.1920928955078125e-07f * 6.0f);
    int rank = 0;
    for (int i = 0; i < nonzero; ++i)
        if (fabsf(diag[i]) > thr) ++rank;
    return rank;
}
