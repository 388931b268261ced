// r360_kernels.cu -- hand-written sm_100a kernels of the spherical dense registration path.
//
//   K1  k_level0 / k_down / k_texel   pyramids of equirectangular sphere images
//                                      (RPI.h:292-354, 365-398, 429-516, 4537-4549)
//   K3  k_pass<METHOD>                 back-projection, SE(3) warp, spherical re-projection,
//                                      nearest-neighbour gather, Huber-weighted photometric +
//                                      depth residuals, 6-DoF Jacobians and the fused
//                                      J^T W J / J^T W r / sum r^2 reduction
//                                      (errorPhotoICP_sphere RPI.h:2545-2739 +
//                                       calcHessGrad_sphere RPI.h:2745-3228, one pass)
//   K4  k_level_begin / k_gn_step / k_finalize   per-pair Gauss-Newton state machine of
//                                      alignFrames360 (RPI.h:4519-4784), no host sync
//       k_warp_dump                    parity hook: index maps + validity masks
//       k_synth                        synthetic sphere frames
//
// HBM-bound gather/reduction: no tensor cores.  Compiled with --fmad=false (sphere_math.h).
#include "r360_device.cuh"
#include "synth.h"
#include "r360_kernels.h"

// =========================================================================== K1: pyramids
// Level 0: RGB8 -> gray (cvtColor fixed point, RPI.h:485) * 1/255 (RPI.h:486);
//          depth u16 mm -> metres (* 0.001f, RPI.h:317).  Output float2 {depth, gray}.
__global__ void __launch_bounds__(256)
k_level0(const uint8_t* __restrict__ rgb, const uint16_t* __restrict__ depth_mm,
         const float* __restrict__ depth_m, float2* const* __restrict__ dst, int n_px) {
    const int f = blockIdx.y;
    const uint8_t* c = rgb + (size_t)f * n_px * 3;
    float2* out = dst[f];
    const float gs = (float)(1. / 255), ds = (float)0.001;
    // 4 pixels per thread: 12 B of RGB as 3 x u32, 8 B of depth as uint2
    const int n4 = n_px >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
        const uint32_t* c4 = reinterpret_cast<const uint32_t*>(c) + 3 * (size_t)q;
        uint32_t w0 = __ldg(c4), w1 = __ldg(c4 + 1), w2 = __ldg(c4 + 2);
        uint8_t b[12];
        b[0] = w0; b[1] = w0 >> 8; b[2] = w0 >> 16; b[3] = w0 >> 24;
        b[4] = w1; b[5] = w1 >> 8; b[6] = w1 >> 16; b[7] = w1 >> 24;
        b[8] = w2; b[9] = w2 >> 8; b[10] = w2 >> 16; b[11] = w2 >> 24;
        float d[4];
        if (depth_mm) {
            uint2 dd = __ldg(reinterpret_cast<const uint2*>(depth_mm + (size_t)f * n_px) + q);
            d[0] = (float)(dd.x & 0xffffu) * ds; d[1] = (float)(dd.x >> 16) * ds;
            d[2] = (float)(dd.y & 0xffffu) * ds; d[3] = (float)(dd.y >> 16) * ds;
        } else {
            float4 dd = __ldg(reinterpret_cast<const float4*>(depth_m + (size_t)f * n_px) + q);
            d[0] = dd.x; d[1] = dd.y; d[2] = dd.z; d[3] = dd.w;
        }
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int v = (b[3 * k] * 9798 + b[3 * k + 1] * 19235 + b[3 * k + 2] * 3735 + 16384) >> 15;
            g[k] = (float)v * gs;
        }
        float4* o = reinterpret_cast<float4*>(out) + 2 * (size_t)q;
        o[0] = make_float4(d[0], g[0], d[1], g[1]);
        o[1] = make_float4(d[2], g[2], d[3], g[3]);
    }
    // tail (n_px not a multiple of 4)
    for (int i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
        int v = (c[3 * (size_t)i] * 9798 + c[3 * (size_t)i + 1] * 19235 + c[3 * (size_t)i + 2] * 3735 + 16384) >> 15;
        float d = depth_mm ? (float)depth_mm[(size_t)f * n_px + i] * ds : depth_m[(size_t)f * n_px + i];
        out[i] = make_float2(d, (float)v * gs);
    }
}

__device__ __forceinline__ int r360_reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// Level l-1 -> l: gray = cv::pyrDown (5x5 Gaussian, REFLECT_101, RPI.h:303), depth = mean of the
// 2x2 parents inside (minDepth, maxDepth) else 0 (RPI.h:322-350).  Same op order as the oracle.
__global__ void __launch_bounds__(256)
k_down(float2* const* __restrict__ pyr, long long off_src, long long off_dst, int rows, int cols,
       float min_d, float max_d) {
    const int f = blockIdx.y;
    const float2* __restrict__ s = pyr[f] + off_src;
    float2* __restrict__ o = pyr[f] + off_dst;
    const int h = rows >> 1, w = cols >> 1, n = h * w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        const int c0 = r360_reflect101(2 * x - 2, cols), c1 = r360_reflect101(2 * x - 1, cols), c2 = 2 * x,
                  c3 = r360_reflect101(2 * x + 1, cols), c4 = r360_reflect101(2 * x + 2, cols);
        float hrow[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float2* row = s + (size_t)r360_reflect101(2 * y - 2 + k, rows) * cols;
            hrow[k] = __ldg(&row[c2]).y * 6 + (__ldg(&row[c1]).y + __ldg(&row[c3]).y) * 4 + __ldg(&row[c0]).y +
                      __ldg(&row[c4]).y;
        }
        const float a = (hrow[0] + hrow[4]) + (hrow[2] + hrow[2]);
        const float b = ((hrow[1] + hrow[3]) + hrow[2]) * 4.0f;
        const float gray = (a + b) * (1.f / 256);
        float av = 0.f;
        unsigned cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float z = __ldg(&s[(size_t)(2 * y + (k >> 1)) * cols + 2 * x + (k & 1)]).x;
            if (z > min_d && z < max_d) { av += z; ++cnt; }
        }
        const float depth = cnt > 0 ? av / cnt : 0.f;
        o[i] = make_float2(depth, gray);
    }
}

// calcGradientXY (RPI.h:365-398): harmonic mean of the one-sided differences on strictly
// monotone triples, 0 elsewhere and on the border.
__device__ __forceinline__ float r360_hgrad(float v, float nxt, float prv) {
    if ((v > nxt && v < prv) || (v < nxt && v > prv)) return 2.f / (1 / (nxt - v) + 1 / (v - prv));
    return 0.f;
}

// Target texels of one level: {gray, depth, Ix, Iy, Dx, Dy} with the sensor-joint columns of the
// four gradient planes zeroed (RPI.h:4537-4549).
__global__ void __launch_bounds__(256)
k_texel(float2* const* __restrict__ pyr, float* const* __restrict__ trg, long long off, int rows,
        int cols, int n_sensors) {
    const int f = blockIdx.y;
    const float2* __restrict__ s = pyr[f] + off;
    float2* __restrict__ o = reinterpret_cast<float2*>(trg[f] + off * R360_TEXEL_FLOATS);
    const int n = rows * cols;
    const int ws = n_sensors > 1 ? cols / n_sensors : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        const float2 v = __ldg(&s[i]);
        float ix = 0.f, iy = 0.f, dx = 0.f, dy = 0.f;
        if (r > 0 && r < rows - 1 && c > 0 && c < cols - 1) {
            const float2 e = __ldg(&s[i + 1]), wv = __ldg(&s[i - 1]), d = __ldg(&s[i + cols]), u = __ldg(&s[i - cols]);
            ix = r360_hgrad(v.y, e.y, wv.y);
            iy = r360_hgrad(v.y, d.y, u.y);
            dx = r360_hgrad(v.x, e.x, wv.x);
            dy = r360_hgrad(v.x, d.x, u.x);
        }
        if (ws > 0) {
            // columns k*ws-1 and k*ws, k = 1..n_sensors-1
            const int k0 = c / ws, rem = c - k0 * ws;
            const bool masked = (rem == 0 && k0 >= 1 && k0 <= n_sensors - 1) ||
                                (rem == ws - 1 && k0 + 1 <= n_sensors - 1);
            if (masked) { ix = iy = dx = dy = 0.f; }
        }
        o[3 * (size_t)i + 0] = make_float2(v.y, v.x);
        o[3 * (size_t)i + 1] = make_float2(ix, iy);
        o[3 * (size_t)i + 2] = make_float2(dx, dy);
    }
}

// =========================================================================== K3: fused pixel pass
// One work item = `px_per_item` consecutive source pixels of one active pair.  A persistent grid
// strides over the (active pair, item) space; per item the CTA keeps 28 float partial sums and
// 3 counters per thread, reduces them with warp shuffles + one shared-memory stage and issues
// one double atomicAdd per sum into the pair's accumulator.
template <int METHOD>
__global__ void __launch_bounds__(R360_PASS_THREADS, 2)
k_pass(R360PassArgs a) {
    __shared__ float s_red[R360_PASS_THREADS / 32][R360_ACC_DOUBLES];
    __shared__ int s_cnt[R360_PASS_THREADS / 32][R360_ACC_INTS];
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const float inv_std_photo = a.inv_std_photo;
    const int n_items = (*a.n_active) * a.items_per_pair;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ap = item / a.items_per_pair;
        const int sub = item - ap * a.items_per_pair;
        const int pair = a.active_list[ap];
        const R360Pair* __restrict__ ps = a.pairs + pair;
        float T[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) T[k] = __ldg(&ps->pose_eval[k]);
        const float2* __restrict__ src = a.src_base[pair] + lv.px_off;
        const float2* __restrict__ trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);

        R360Acc A;
        r360_acc_zero(A);
        int n_vis = 0, n_photo = 0, n_depth = 0;

        const int i_end = min(lv.n, (sub + 1) * a.px_per_item);
        // Batches of R360_PASS_U pixels per thread, staged so that the U source loads, then the U
        // texel gathers, are in flight together (the two dependent HBM round trips per pixel are
        // what bounds this kernel); the source loads of the next batch are prefetched.
        int i0 = sub * a.px_per_item + threadIdx.x;
        float2 sd[R360_PASS_U];
#pragma unroll
        for (int u = 0; u < R360_PASS_U; ++u) {
            const int i = i0 + u * R360_PASS_THREADS;
            sd[u] = i < i_end ? __ldg(&src[i]) : make_float2(0.f, 0.f);
        }
        for (; i0 < i_end; i0 += R360_PASS_THREADS * R360_PASS_U) {
            float2 sdn[R360_PASS_U];
#pragma unroll
            for (int u = 0; u < R360_PASS_U; ++u) {
                const int i = i0 + (u + R360_PASS_U) * R360_PASS_THREADS;
                sdn[u] = i < i_end ? __ldg(&src[i]) : make_float2(0.f, 0.f);
            }
            R360Warp w[R360_PASS_U];
            bool ok[R360_PASS_U];
            const float2* tx[R360_PASS_U];
#pragma unroll
            for (int u = 0; u < R360_PASS_U; ++u) {
                const int i = i0 + u * R360_PASS_THREADS;
                const float d = sd[u].x;
                // LUT INVALID_POINT (RPI.h:4575,4585); out-of-range lanes of the tail carry d = 0
                ok[u] = i < i_end && (P.min_depth < d && d < P.max_depth);
                const int ii = ok[u] ? i : 0;
                const int r = (int)(((unsigned long long)ii * lv.div_magic) >> 40);
                const int c = ii - r * lv.cols;
                float X[3];
                r360_backproject(d, __ldg(&lv.sin_p[r]), __ldg(&lv.cos_p[r]), __ldg(&lv.sin_t[c]),
                                 __ldg(&lv.cos_t[c]), X);
                const bool inb = r360_warp_point(T, X, lv.res_inv, lv.half_rows, lv.rows, lv.cols, w[u]);
                ok[u] = ok[u] && inb;
                tx[u] = trg + 3 * (ok[u] ? ((size_t)w[u].r * lv.cols + w[u].c) : 0);
            }
            float2 t0[R360_PASS_U], t1[R360_PASS_U], t2[R360_PASS_U];
#pragma unroll
            for (int u = 0; u < R360_PASS_U; ++u) {
                t0[u] = __ldg(tx[u]); t1[u] = __ldg(tx[u] + 1); t2[u] = __ldg(tx[u] + 2);
            }
#pragma unroll
            for (int u = 0; u < R360_PASS_U; ++u) {
                if (ok[u]) {
                    ++n_vis;
                    R360Row ph, dp;
                    const int v = r360_rows<METHOD>(w[u], lv.res_inv, sd[u].y, t0[u].x, t0[u].y, t1[u].x, t1[u].y,
                                                    t2[u].x, t2[u].y, P, inv_std_photo, ph, dp);
                    if (v & 1) { r360_accumulate(A, ph); ++n_photo; }
                    if (v & 2) { r360_accumulate(A, dp); ++n_depth; }
                }
            }
#pragma unroll
            for (int u = 0; u < R360_PASS_U; ++u) sd[u] = sdn[u];
        }

        // ---- block reduction
        float acc[R360_ACC_DOUBLES];
        r360_acc_unpack(A, acc);
#pragma unroll
        for (int k = 0; k < R360_ACC_DOUBLES; ++k) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[wid][k] = v;
        }
        n_vis = __reduce_add_sync(0xffffffffu, n_vis);
        n_photo = __reduce_add_sync(0xffffffffu, n_photo);
        n_depth = __reduce_add_sync(0xffffffffu, n_depth);
        if (lane == 0) { s_cnt[wid][0] = n_vis; s_cnt[wid][1] = n_photo; s_cnt[wid][2] = n_depth; }
        __syncthreads();
        if (threadIdx.x < R360_ACC_DOUBLES) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < R360_PASS_THREADS / 32; ++k) s += (double)s_red[k][threadIdx.x];
            atomicAdd(&a.acc[(size_t)pair * R360_ACC_DOUBLES + threadIdx.x], s);
        } else if (threadIdx.x >= 32 && threadIdx.x < 35) {
            int s = 0;
#pragma unroll
            for (int k = 0; k < R360_PASS_THREADS / 32; ++k) s += s_cnt[k][threadIdx.x - 32];
            atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], s);
        }
        __syncthreads();
    }
}

// Parity hook: per source pixel the rounded target index and the validPixelsPhoto/Depth masks.
__global__ void __launch_bounds__(256)
k_warp_dump(R360PassArgs a, int pair, int method, int32_t* __restrict__ r_idx, int32_t* __restrict__ c_idx,
            uint8_t* __restrict__ vphoto, uint8_t* __restrict__ vdepth) {
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const R360Pair* ps = a.pairs + pair;
    float T[16];
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float2* src = a.src_base[pair] + lv.px_off;
    const float2* trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lv.n; i += gridDim.x * blockDim.x) {
        int rr = INT_MIN, cc = INT_MIN, v = 0;
        const float2 sd = src[i];
        if (P.min_depth < sd.x && sd.x < P.max_depth) {
            const int r = (int)(((unsigned long long)i * lv.div_magic) >> 40);
            const int c = i - r * lv.cols;
            float X[3];
            r360_backproject(sd.x, lv.sin_p[r], lv.cos_p[r], lv.sin_t[c], lv.cos_t[c], X);
            R360Warp w;
            const bool inb = r360_warp_point(T, X, lv.res_inv, lv.half_rows, lv.rows, lv.cols, w);
            rr = w.r; cc = w.c;
            if (inb) {
                const float2* tx = trg + 3 * ((size_t)w.r * lv.cols + w.c);
                const float2 t0 = tx[0], t1 = tx[1], t2 = tx[2];
                R360Row ph, dp;
                if (method == R360_PHOTO_CONSISTENCY)
                    v = r360_rows<R360_PHOTO_CONSISTENCY>(w, lv.res_inv, sd.y, t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, P, a.inv_std_photo, ph, dp);
                else if (method == R360_DEPTH_CONSISTENCY)
                    v = r360_rows<R360_DEPTH_CONSISTENCY>(w, lv.res_inv, sd.y, t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, P, a.inv_std_photo, ph, dp);
                else
                    v = r360_rows<R360_PHOTO_DEPTH>(w, lv.res_inv, sd.y, t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, P, a.inv_std_photo, ph, dp);
            }
        }
        if (r_idx) r_idx[i] = rr;
        if (c_idx) c_idx[i] = cc;
        if (vphoto) vphoto[i] = (uint8_t)(v & 1);
        if (vdepth) vdepth[i] = (uint8_t)((v >> 1) & 1);
    }
}

// =========================================================================== K4: Gauss-Newton state machine
__device__ void r360_zero_acc(double* acc, int* cnt, int pair) {
    for (int k = 0; k < R360_ACC_DOUBLES; ++k) acc[(size_t)pair * R360_ACC_DOUBLES + k] = 0.0;
    for (int k = 0; k < R360_ACC_INTS; ++k) cnt[(size_t)pair * R360_ACC_INTS + k] = 0;
}

// Compacts the active flags of pairs [0, n) into active_list (ascending pair order), one warp.
__device__ void r360_compact_active(R360Pair* pairs, int n, int* active_list, int* n_active) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        const bool on = p < n && pairs[p].active;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) active_list[base + __popc(m & ((1u << lane) - 1))] = p;
        base += __popc(m);
    }
    if (lane == 0) *n_active = base;
}

// Start of a pyramid level (RPI.h:4589-4605): evaluate the current estimate first.
__global__ void k_level_begin(R360GnArgs g, int level) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pairs; p += gridDim.x * blockDim.x) {
        R360Pair* ps = g.pairs + p;
        r360_zero_acc(g.acc, g.cnt, p);
        if (ps->status != R360_PAIR_OK) { ps->active = 0; continue; }
        for (int k = 0; k < 16; ++k) ps->pose_eval[k] = ps->pose_estim[k];
        for (int k = 0; k < 6; ++k) ps->upd[k] = 1.f;
        ps->lambda = 1.0;
        ps->it = 0;
        ps->phase = 0;
        ps->ev = 0;
        ps->active = 1;
    }
}

__global__ void k_compact(R360GnArgs g) {
    if (threadIdx.x < 32 && blockIdx.x == 0) r360_compact_active(g.pairs, g.n_pairs, g.active_list, g.n_active);
}

// One step of the per-pair state machine after a pixel pass (RPI.h:4599-4722).
//   phase 0: the pass evaluated pose_estim at the start of the level.
//   phase 1: the pass evaluated the candidate pose_tmp = exp(update) * pose_estim.
// The reference runs errorPhotoICP_sphere(pose_tmp) and, if accepted, calcHessGrad_sphere at the
// same pose in the next loop body; the fused pass already produced both.
__global__ void k_gn_step(R360GnArgs g, int level) {
    const r360_params P = g.params;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pairs; p += gridDim.x * blockDim.x) {
        R360Pair* ps = g.pairs + p;
        if (!ps->active) continue;
        const double* acc = g.acc + (size_t)p * R360_ACC_DOUBLES;
        const int* cnt = g.cnt + (size_t)p * R360_ACC_INTS;
        const double e2 = acc[27];
        const int n_valid = cnt[1] + cnt[2];
        const int n_vis = cnt[0];
        const double err = sqrt(e2 / (double)n_valid);      // RPI.h:2738
        ps->passes[level] += 1;
        double diff_error;
        int accepted;
        if (ps->phase == 0) {
            diff_error = err;                               // RPI.h:4605
            accepted = 1;
        } else {
            diff_error = ps->error - err;                   // RPI.h:4711
            accepted = diff_error > P.tol_residual;         // RPI.h:4715
        }
        r360_iter_record* rec = nullptr;
        if (g.trace && ps->ev < P.max_iters + 2)
            rec = g.trace + ((size_t)p * P.n_levels + level) * (P.max_iters + 2) + ps->ev;
        ps->ev += 1;
        if (accepted) {
            if (ps->phase == 1) {
                ps->lambda /= 5.0;                          // RPI.h:4718
                for (int k = 0; k < 16; ++k) ps->pose_estim[k] = ps->pose_eval[k];
                ps->it += 1;
            }
            ps->error = err; ps->err2 = e2; ps->n_valid = n_valid;
            for (int k = 0; k < 21; ++k) ps->Hc[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) ps->gc[k] = (float)acc[21 + k];
            ps->nvis_c = n_vis;
        }
        if (rec) {
            rec->err2 = e2; rec->n_valid = n_valid; rec->n_visible = n_vis; rec->level = level;
            rec->it = ps->it; rec->accepted = accepted; rec->used = 3;
            for (int k = 0; k < 16; ++k) rec->pose[k] = ps->pose_eval[k];
            for (int k = 0; k < 21; ++k) rec->hessian[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) rec->gradient[k] = (float)acc[21 + k];
            rec->pad = 0.f;
        }
        ps->phase = 1;
        r360_zero_acc(g.acc, g.cnt, p);

        // while (it < maxIters && update_pose.norm() > tol_update && diff_error > tol_residual)
        const float* u = ps->upd;
        const float na = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
        const float nb = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
        const float unorm = sqrtf(na + nb);
        const bool go = ps->it < P.max_iters && (double)unorm > P.tol_update && diff_error > P.tol_residual;
        if (!go) {
            ps->iters[level] = ps->it;                      // RPI.h:4772
            ps->active = 0;
            continue;
        }
        // loop body: calcHessGrad_sphere(pose_estim) == (Hc, gc)            RPI.h:4623
        for (int k = 0; k < 21; ++k) ps->Hl[k] = ps->Hc[k];
        for (int k = 0; k < 6; ++k) ps->gl[k] = ps->gc[k];
        ps->nvis_l = ps->nvis_c;
        ps->lvl_l = level;
        float H[36], Hlam[36];
        {
            int q = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b, ++q) H[a + 6 * b] = H[b + 6 * a] = ps->Hc[q];
        }
        const float lam = (float)ps->lambda;
        for (int k = 0; k < 36; ++k) Hlam[k] = H[k];
        for (int a = 0; a < 6; ++a) Hlam[a + 6 * a] = H[a + 6 * a] + lam * H[a + 6 * a];
        if (r360_rank6(Hlam) != 6) {                        // RPI.h:4682-4690
            ps->status = R360_PAIR_ILL_POSED;
            ps->active = 0;
            continue;
        }
        float inv[36], upd[6];
        r360_inverse6(H, inv);
        r360_solve_update(inv, ps->gc, upd);                // RPI.h:4693
        double ud[6], Td[16];
        for (int k = 0; k < 6; ++k) { ps->upd[k] = upd[k]; ud[k] = (double)upd[k]; }
        r360_pseudo_exp(ud, Td);                            // RPI.h:4695-4697
        float Tf[16], Tn[16];
        for (int k = 0; k < 16; ++k) Tf[k] = (float)Td[k];
        r360_mat4_mul(Tf, ps->pose_estim, Tn);
        for (int k = 0; k < 16; ++k) ps->pose_eval[k] = Tn[k];
    }
}

__global__ void k_pairs_init(R360GnArgs g, const int32_t* __restrict__ src_idx,
                             const int32_t* __restrict__ trg_idx, const float* __restrict__ init_pose) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pairs; p += gridDim.x * blockDim.x) {
        R360Pair* ps = g.pairs + p;
        for (int k = 0; k < 16; ++k) {
            float v = init_pose ? init_pose[(size_t)p * 16 + k] : ((k % 5 == 0) ? 1.f : 0.f);
            ps->pose_estim[k] = v;
            ps->pose_eval[k] = v;
        }
        for (int k = 0; k < 21; ++k) { ps->Hc[k] = 0.f; ps->Hl[k] = 0.f; }
        for (int k = 0; k < 6; ++k) { ps->gc[k] = 0.f; ps->gl[k] = 0.f; ps->upd[k] = 1.f; }
        ps->error = 0.0; ps->err2 = 0.0; ps->lambda = 1.0;
        ps->n_valid = 0; ps->nvis_c = 0; ps->nvis_l = 0; ps->lvl_l = -1;
        ps->it = 0; ps->phase = 0; ps->active = 0; ps->status = R360_PAIR_OK; ps->ev = 0;
        ps->src = src_idx[p]; ps->trg = trg_idx[p];
        for (int k = 0; k < R360_MAX_LEVELS; ++k) { ps->iters[k] = 0; ps->passes[k] = 0; }
    }
}

__global__ void k_finalize(R360GnArgs g, r360_result* __restrict__ out, int rows, int cols) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pairs; p += gridDim.x * blockDim.x) {
        const R360Pair* ps = g.pairs + p;
        r360_result* r = out + p;
        for (int k = 0; k < 16; ++k) r->pose[k] = ps->pose_estim[k];        // RPI.h:4783 / 4687
        int q = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b, ++q) r->hessian[a + 6 * b] = r->hessian[b + 6 * a] = ps->Hl[q];
        for (int k = 0; k < 6; ++k) r->gradient[k] = ps->gl[k];
        r->n_visible = ps->nvis_l;
        r->sso = ps->lvl_l >= 0 ? (float)ps->nvis_l / (float)((rows >> ps->lvl_l) * (cols >> ps->lvl_l)) : 0.f;
        r->final_error = ps->error;
        r->final_err2 = ps->err2;
        r->final_n_valid = ps->n_valid;
        r->status = ps->status;
        for (int k = 0; k < R360_MAX_LEVELS; ++k) { r->iters[k] = ps->iters[k]; r->passes[k] = ps->passes[k]; }
        r->pair_id = p;
        r->reserved = 0;
    }
}

// =========================================================================== synthetic frames
__global__ void __launch_bounds__(256)
k_synth(int kind, int first_id, int rows, int cols, const float* __restrict__ cams /* n x 12 */,
        uint8_t* __restrict__ rgb, uint16_t* __restrict__ depth_mm) {
    const int f = blockIdx.y;
    const float* R = cams + 12 * f;
    const float* t = R + 9;
    const float res = (float)(2 * R360_PI_D / cols);
    const float half_rows = (float)(0.5 * rows - 0.5);
    const int n = rows * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        float sp, cp, st, ct;
        r360_sincosf((half_rows - r) * res, &sp, &cp);
        r360_sincosf(c * res, &st, &ct);
        uint8_t g; uint16_t d;
        r360_synth_pixel(R, t, sp, cp, st, ct, &g, &d);
        const size_t o = (size_t)f * n + i;
        rgb[3 * o] = g; rgb[3 * o + 1] = g; rgb[3 * o + 2] = g;
        depth_mm[o] = d;
    }
}

// =========================================================================== launch wrappers
static inline int r360_blocks(long long n, int threads, int cap) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

void r360_launch_level0(cudaStream_t st, const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m,
                        float2* const* dst, int n_frames, int n_px, int sm_count) {
    dim3 grid(r360_blocks(n_px / 4 + 1, 256, sm_count * 8), n_frames);
    k_level0<<<grid, 256, 0, st>>>(rgb, depth_mm, depth_m, dst, n_px);
}
void r360_launch_down(cudaStream_t st, float2* const* pyr, long long off_src, long long off_dst, int rows,
                      int cols, float min_d, float max_d, int n_frames, int sm_count) {
    dim3 grid(r360_blocks((long long)(rows / 2) * (cols / 2), 256, sm_count * 8), n_frames);
    k_down<<<grid, 256, 0, st>>>(pyr, off_src, off_dst, rows, cols, min_d, max_d);
}
void r360_launch_texel(cudaStream_t st, float2* const* pyr, float* const* trg, long long off, int rows, int cols,
                       int n_sensors, int n_frames, int sm_count) {
    dim3 grid(r360_blocks((long long)rows * cols, 256, sm_count * 8), n_frames);
    k_texel<<<grid, 256, 0, st>>>(pyr, trg, off, rows, cols, n_sensors);
}
void r360_launch_pass(cudaStream_t st, const R360PassArgs& a, int grid) {
    switch (a.params.method) {
        case R360_PHOTO_CONSISTENCY: k_pass<R360_PHOTO_CONSISTENCY><<<grid, R360_PASS_THREADS, 0, st>>>(a); break;
        case R360_DEPTH_CONSISTENCY: k_pass<R360_DEPTH_CONSISTENCY><<<grid, R360_PASS_THREADS, 0, st>>>(a); break;
        default: k_pass<R360_PHOTO_DEPTH><<<grid, R360_PASS_THREADS, 0, st>>>(a); break;
    }
}
void r360_launch_warp_dump(cudaStream_t st, const R360PassArgs& a, int pair, int32_t* r_idx, int32_t* c_idx,
                           uint8_t* vp, uint8_t* vd, int sm_count) {
    k_warp_dump<<<r360_blocks(a.lv.n, 256, sm_count * 8), 256, 0, st>>>(a, pair, a.params.method, r_idx, c_idx, vp, vd);
}
void r360_launch_pairs_init(cudaStream_t st, const R360GnArgs& g, const int32_t* src_idx, const int32_t* trg_idx,
                            const float* init_pose) {
    k_pairs_init<<<r360_blocks(g.n_pairs, 128, 1024), 128, 0, st>>>(g, src_idx, trg_idx, init_pose);
}
void r360_launch_level_begin(cudaStream_t st, const R360GnArgs& g, int level) {
    k_level_begin<<<r360_blocks(g.n_pairs, 64, 1024), 64, 0, st>>>(g, level);
    k_compact<<<1, 32, 0, st>>>(g);
}
void r360_launch_gn_step(cudaStream_t st, const R360GnArgs& g, int level) {
    k_gn_step<<<r360_blocks(g.n_pairs, 32, 1024), 32, 0, st>>>(g, level);
    k_compact<<<1, 32, 0, st>>>(g);
}
void r360_launch_finalize(cudaStream_t st, const R360GnArgs& g, r360_result* out, int rows, int cols) {
    k_finalize<<<r360_blocks(g.n_pairs, 128, 1024), 128, 0, st>>>(g, out, rows, cols);
}
void r360_launch_synth(cudaStream_t st, int kind, int first_id, int rows, int cols, const float* cams, int n_frames,
                       uint8_t* rgb, uint16_t* depth_mm, int sm_count) {
    dim3 grid(r360_blocks((long long)rows * cols, 256, sm_count * 8), n_frames);
    k_synth<<<grid, 256, 0, st>>>(kind, first_id, rows, cols, cams, rgb, depth_mm);
}
