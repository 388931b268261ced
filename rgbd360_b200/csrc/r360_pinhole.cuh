// r360_pinhole.cuh -- the pinhole registration of RegisterPhotoICP, included by
// r360_kernels.cu (it shares the state-machine helpers of that translation unit):
//   k_pin_eval<METHOD>   errorPhotoICP (RPI.h:560-775) + calcHessGrad (RPI.h:776-1104) at one pose, fused:
//                        the reference evaluates the error of a candidate and, when the step is accepted,
//                        the Hessian at the same pose in the next loop body.
//   k_gn_step_pin        the Levenberg-Marquardt state machine of alignFrames (RPI.h:4254-4512): accept on
//                        diff_error > 0, one damped retry otherwise, full SE(3) exponential.
// Same pyramids and texel layout as the spherical path (built without the sensor-joint mask).  The
// index maps are bit-exact by construction: the projection is evaluated with the reference's own
// operation sequence -- including its double-precision 1 / z narrowed to float (RPI.h:659) -- and the
// file is compiled with --fmad=false.  k_pin_eval: one pixel per thread (the plain statement, R360_PIN_PACKED=0);
// k_pin_eval2: two pixels per thread in packed fp32x2 (the default).  Both gather directly.
#pragma once

#define R360_PIN_THREADS 256
#ifndef R360_PIN_CAP
#define R360_PIN_CAP 16                       // blocks per pair so that the grid is <= 16 CTAs per SM (8: 7.33 ms, 16: 6.37 ms, 32: 6.78 ms per 256-pair batch)
#endif

template <int METHOD>
__global__ void __launch_bounds__(R360_PIN_THREADS)
k_pin_eval(R360PassArgs a, R360PinLevel pl) {
    __shared__ float s_red[R360_PIN_THREADS / 32][R360_ACC_DOUBLES + 1];
    __shared__ int s_cnt[R360_PIN_THREADS / 32][R360_ACC_INTS];
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    float T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float2* __restrict__ src = a.src_base[pair] + lv.px_off;
    const float2* __restrict__ trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);

    float H[21], g[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) H[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = 0.f;
    float sumP = 0.f, sumD = 0.f;
    int n_vis = 0, n_photo = 0, n_depth = 0;
    auto accumulate = [&](const float J[6], float r) {
        int q = 0;
#pragma unroll
        for (int x = 0; x < 6; ++x) {
#pragma unroll
            for (int y = x; y < 6; ++y, ++q) H[q] = fmaf(J[x], J[y], H[q]);
            g[x] = fmaf(J[x], r, g[x]);
        }
    };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lv.n; i += gridDim.x * blockDim.x) {
        const int r = (int)(((unsigned long long)i * lv.div_magic) >> 40), c = i - r * lv.cols;
        const float2 s = __ldg(&src[i]);                                   // {depth, gray}
        const float z = s.x;
        if (!(P.min_depth < z && z < P.max_depth)) continue;               // LUT x = INVALID_POINT, RPI.h:4293-4299
        const float X = (c - pl.ox) * z * pl.inv_fx;                       // RPI.h:4295
        const float Y = (r - pl.oy) * z * pl.inv_fy;
        const float px = ((T[0] * X + T[4] * Y) + T[8] * z) + T[12];
        const float py = ((T[1] * X + T[5] * Y) + T[9] * z) + T[13];
        const float pz = ((T[2] * X + T[6] * Y) + T[10] * z) + T[14];
        const float iz = (float)(1.0 / (double)pz);                        // RPI.h:659
        const float tc = (px * pl.fx) * iz + pl.ox;                        // RPI.h:662
        const float tr = (py * pl.fy) * iz + pl.oy;
        const int ri = r360_round_to_int_dev(tr), ci = r360_round_to_int_dev(tc);
        if (!((unsigned)ri < (unsigned)lv.rows && (unsigned)ci < (unsigned)lv.cols)) continue;   // RPI.h:667-668
        const float2* tx = trg + 3u * (unsigned)(ri * lv.cols + ci);
        const float2 t0 = __ldg(tx), t1 = __ldg(tx + 1), t2 = __ldg(tx + 2);   // {gray, depth}, {Ix, Iy}, {Dx, Dy}
        const bool fin = fabsf(t0.y) < INFINITY;
        // weighted residuals (shared by the error and the Hessian rows)
        float rp = 0.f, wp = 0.f, rd = 0.f, wd = 0.f;
        if (METHOD != R360_DEPTH_CONSISTENCY) {
            const float e = t0.x - s.y;
            wp = a.inv_std_photo;
            if (!(fabsf(e) < P.std_photo)) { const float u = r360_rcp_fast(fabsf(e)); wp = r360_sqrt_fast(u * (2.f * a.inv_std_photo - u)); }
            rp = wp * e;
        }
        if (METHOD != R360_PHOTO_CONSISTENCY && fin) {
            const float f = t0.y - pz;
            const float sd = P.std_depth * pz;                             // RPI.h:694: transformed SOURCE depth
            wd = r360_rcp_fast(sd);
            if (!(fabsf(f) < sd)) { const float u = r360_rcp_fast(fabsf(f)); wd = r360_sqrt_fast(u * (2.f * wd - u)); }
            rd = wd * f;
        }
        // ---- errorPhotoICP: no saliency test (RPI.h:669-703)
        if (METHOD != R360_DEPTH_CONSISTENCY) { sumP += rp * rp; ++n_photo; }
        if (METHOD != R360_PHOTO_CONSISTENCY && fin) { sumD += rd * rd; ++n_depth; }
        // ---- calcHessGrad (RPI.h:967-1085): both saliency `continue`s drop the whole pixel
        ++n_vis;
        if (METHOD != R360_DEPTH_CONSISTENCY &&
            (fabsf(t1.x) < P.thres_sal_int) & (fabsf(t1.y) < P.thres_sal_int)) continue;
        if (METHOD != R360_PHOTO_CONSISTENCY &&
            (fabsf(t2.x) < P.thres_sal_depth) & (fabsf(t2.y) < P.thres_sal_depth)) continue;
        const float iz2 = iz * iz;
        float J0[6], J1[6];                                                // jacobianWarpRt, RPI.h:970-984
        J0[0] = pl.fx * iz; J1[0] = 0.f;
        J0[1] = 0.f; J1[1] = pl.fy * iz;
        J0[2] = -pl.fx * px * iz2; J1[2] = -pl.fy * py * iz2;
        J0[3] = -pl.fx * py * px * iz2; J1[3] = -pl.fy * (1 + py * py * iz2);
        J0[4] = pl.fx * (1 + px * px * iz2); J1[4] = pl.fy * px * py * iz2;
        J0[5] = -pl.fx * py * iz; J1[5] = pl.fy * px * iz;
        float J[6];
        if (METHOD != R360_DEPTH_CONSISTENCY) {
            const float ga = wp * t1.x, gb = wp * t1.y;
#pragma unroll
            for (int q = 0; q < 6; ++q) J[q] = ga * J0[q] + gb * J1[q];
            accumulate(J, rp);
        }
        if (METHOD != R360_PHOTO_CONSISTENCY && fin) {
            const float Rz[6] = { 0.f, 0.f, 1.f, py, -px, 0.f };           // jacobianRt_z, RPI.h:1053
#pragma unroll
            for (int q = 0; q < 6; ++q) J[q] = wd * ((t2.x * J0[q] + t2.y * J1[q]) - Rz[q]);
            accumulate(J, rd);
        }
    }

    // ---- block reduction (as k_occ_eval): 27 normal-equation sums + PhotoResidual + DepthResidual, 3 counters
    float acc[R360_ACC_DOUBLES + 1];
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[k] = H[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[21 + k] = g[k];
    acc[27] = sumP;
    acc[28] = sumD;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < R360_ACC_DOUBLES + 1; ++k) {
        float v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[wid][k] = v;
    }
    n_vis = __reduce_add_sync(0xffffffffu, n_vis);
    n_photo = __reduce_add_sync(0xffffffffu, n_photo);
    n_depth = __reduce_add_sync(0xffffffffu, n_depth);
    if (lane == 0) { s_cnt[wid][0] = n_vis; s_cnt[wid][1] = n_photo; s_cnt[wid][2] = n_depth; s_cnt[wid][3] = 0; }
    __syncthreads();
    if (threadIdx.x < R360_ACC_DOUBLES + 1) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
        r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
        atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
    }
}

// The same evaluation, two horizontally adjacent pixels per thread in packed fp32x2 (the formulation of k_pass): the
// projection is the reference's operation sequence on both pixels at once -- every product and sum of R X + t and of the
// projection rounded separately (f2add_sep), the reciprocal of z still formed in DOUBLE and narrowed (RPI.h:659), the
// rounding to the nearest texel by the 1.5 * 2^23 trick with the scalar function on exact ties and out-of-range values -- so
// the index maps and the counters are those of k_pin_eval bit for bit; residuals, weights and Jacobian rows use fused
// multiply-adds like the spherical kernel (sums within float rounding of the scalar kernel's).  Both pixels' six texel
// gathers are issued together; 28 FFMA2 per residual row pair.
template <int METHOD>
__global__ void __launch_bounds__(R360_PIN_THREADS, 2)
k_pin_eval2(R360PassArgs a, R360PinLevel pl) {
    __shared__ float s_red[R360_PIN_THREADS / 32][R360_ACC_DOUBLES + 1];
    __shared__ int s_cnt[R360_PIN_THREADS / 32][R360_ACC_INTS];
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    float T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float4* __restrict__ src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    const float2* __restrict__ trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    const float one = a.one;

    R360Acc2 A;
    r360_acc_zero(A);
    float2 sumP = make_float2(0.f, 0.f), sumD = sumP;
    int n_vis = 0, n_photo = 0, n_depth = 0;
    const int n2 = lv.n >> 1;                                               // cols is even at every level: pixel pairs never straddle rows
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += gridDim.x * blockDim.x) {
        const int i = 2 * q;
        const int r = (int)(((unsigned long long)i * lv.div_magic) >> 40), c = i - r * lv.cols;
        const float4 s = __ldg(&src4[q]);                                   // {d0, g0, d1, g1}
        const bool v0 = (P.min_depth < s.x) & (s.x < P.max_depth), v1 = (P.min_depth < s.z) & (s.z < P.max_depth);   // RPI.h:4293-4299
        if (!(v0 | v1)) continue;
        const float2 z = make_float2(v0 ? s.x : 1.f, v1 ? s.z : 1.f);       // invalid lanes stay finite
        const float2 cf = make_float2((float)c - pl.ox, (float)(c + 1) - pl.ox);
        const float rf = (float)r - pl.oy;
        const float2 X = f2mul(f2mul(cf, z), R360_F2(pl.inv_fx));           // RPI.h:4295
        const float2 Y = f2mul(f2mul(R360_F2(rf), z), R360_F2(pl.inv_fy));
        const float2 px = f2add(f2add_sep(f2add_sep(f2mul(X, R360_F2(T[0])), one, f2mul(Y, R360_F2(T[4]))), one, f2mul(z, R360_F2(T[8]))), R360_F2(T[12]));
        const float2 py = f2add(f2add_sep(f2add_sep(f2mul(X, R360_F2(T[1])), one, f2mul(Y, R360_F2(T[5]))), one, f2mul(z, R360_F2(T[9]))), R360_F2(T[13]));
        const float2 pz = f2add(f2add_sep(f2add_sep(f2mul(X, R360_F2(T[2])), one, f2mul(Y, R360_F2(T[6]))), one, f2mul(z, R360_F2(T[10]))), R360_F2(T[14]));
        const float2 iz = make_float2((float)(1.0 / (double)pz.x), (float)(1.0 / (double)pz.y));     // RPI.h:659
        const float2 tc = f2add_sep(f2mul(f2mul(px, R360_F2(pl.fx)), iz), one, R360_F2(pl.ox));      // RPI.h:662
        const float2 tr = f2add_sep(f2mul(f2mul(py, R360_F2(pl.fy)), iz), one, R360_F2(pl.oy));
        // round half away from zero == round to nearest except on exact ties; nearest via 1.5 * 2^23, exact for |v| < 2^22
        const float M = 12582912.0f;
        const float2 mr = f2add_sep(tr, one, R360_F2(M)), mc = f2add_sep(tc, one, R360_F2(M));
        const float2 dr = f2add(tr, f2add(R360_F2(M), f2neg(mr))), dc = f2add(tc, f2add(R360_F2(M), f2neg(mc)));
        int ri0 = __float_as_int(mr.x) - 0x4B400000, ri1 = __float_as_int(mr.y) - 0x4B400000;
        int ci0 = __float_as_int(mc.x) - 0x4B400000, ci1 = __float_as_int(mc.y) - 0x4B400000;
        const bool e0 = !((fmaxf(fabsf(dr.x), fabsf(dc.x)) < 0.5f) & (fmaxf(fabsf(tr.x), fabsf(tc.x)) < 4194304.f));
        const bool e1 = !((fmaxf(fabsf(dr.y), fabsf(dc.y)) < 0.5f) & (fmaxf(fabsf(tr.y), fabsf(tc.y)) < 4194304.f));
        if (e0) { ri0 = r360_round_to_int_dev(tr.x); ci0 = r360_round_to_int_dev(tc.x); }            // ties, huge values, NaN: the scalar rule
        if (e1) { ri1 = r360_round_to_int_dev(tr.y); ci1 = r360_round_to_int_dev(tc.y); }
        const bool ok0 = v0 & ((unsigned)ri0 < (unsigned)lv.rows) & ((unsigned)ci0 < (unsigned)lv.cols);   // RPI.h:667-668
        const bool ok1 = v1 & ((unsigned)ri1 < (unsigned)lv.rows) & ((unsigned)ci1 < (unsigned)lv.cols);
        if (!(ok0 | ok1)) continue;
        const float2* ta = trg + 3u * (ok0 ? (unsigned)(ri0 * lv.cols + ci0) : 0u);
        const float2* tb = trg + 3u * (ok1 ? (unsigned)(ri1 * lv.cols + ci1) : 0u);
        const float2 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2);     // {gray, depth}, {Ix, Iy}, {Dx, Dy}
        const float2 b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2);
        const bool fin0 = ok0 & (fabsf(a0.y) < INFINITY), fin1 = ok1 & (fabsf(b0.y) < INFINITY);
        // weighted residuals (shared by the error and the Hessian rows)
        float2 rp = make_float2(0.f, 0.f), wp = rp, rd = rp, wd = rp;
        if (METHOD != R360_DEPTH_CONSISTENCY) {
            const float2 e = make_float2(a0.x - s.y, b0.x - s.w);
            float w0 = a.inv_std_photo, w1 = a.inv_std_photo;
            if (!(fabsf(e.x) < P.std_photo)) { const float u = r360_rcp_fast(fabsf(e.x)); w0 = r360_sqrt_fast(u * (2.f * a.inv_std_photo - u)); }
            if (!(fabsf(e.y) < P.std_photo)) { const float u = r360_rcp_fast(fabsf(e.y)); w1 = r360_sqrt_fast(u * (2.f * a.inv_std_photo - u)); }
            wp = make_float2(ok0 ? w0 : 0.f, ok1 ? w1 : 0.f);
            rp = f2mul(wp, e);
            sumP = f2fma(rp, rp, sumP);                                     // errorPhotoICP: no saliency test (RPI.h:669-703)
            r360_count(n_photo, ok0); r360_count(n_photo, ok1);
        }
        if (METHOD != R360_PHOTO_CONSISTENCY) {
            const float2 D = make_float2(fin0 ? a0.y : 1.f, fin1 ? b0.y : 1.f);
            const float2 f = f2add(D, f2neg(pz));
            const float2 sd = f2mul(R360_F2(P.std_depth), pz);              // RPI.h:694: transformed SOURCE depth
            float w0 = r360_rcp_fast(sd.x), w1 = r360_rcp_fast(sd.y);
            if (!(fabsf(f.x) < sd.x)) { const float u = r360_rcp_fast(fabsf(f.x)); w0 = r360_sqrt_fast(u * (2.f * w0 - u)); }
            if (!(fabsf(f.y) < sd.y)) { const float u = r360_rcp_fast(fabsf(f.y)); w1 = r360_sqrt_fast(u * (2.f * w1 - u)); }
            wd = make_float2(fin0 ? w0 : 0.f, fin1 ? w1 : 0.f);
            rd = f2mul(wd, f);
            sumD = f2fma(rd, rd, sumD);
            r360_count(n_depth, fin0); r360_count(n_depth, fin1);
        }
        r360_count(n_vis, ok0); r360_count(n_vis, ok1);
        // ---- calcHessGrad (RPI.h:967-1085): both saliency `continue`s drop the whole pixel
        bool h0 = ok0, h1 = ok1;
        if (METHOD != R360_DEPTH_CONSISTENCY) {
            h0 = h0 & !((fabsf(a1.x) < P.thres_sal_int) & (fabsf(a1.y) < P.thres_sal_int));
            h1 = h1 & !((fabsf(b1.x) < P.thres_sal_int) & (fabsf(b1.y) < P.thres_sal_int));
        }
        if (METHOD != R360_PHOTO_CONSISTENCY) {
            h0 = h0 & !((fabsf(a2.x) < P.thres_sal_depth) & (fabsf(a2.y) < P.thres_sal_depth));
            h1 = h1 & !((fabsf(b2.x) < P.thres_sal_depth) & (fabsf(b2.y) < P.thres_sal_depth));
        }
        if (!(h0 | h1)) continue;
        const float2 iz2 = f2mul(iz, iz);
        const float2 fxz = f2mul(R360_F2(pl.fx), iz), fyz = f2mul(R360_F2(pl.fy), iz);
        const float2 fxz2 = f2mul(R360_F2(pl.fx), iz2), fyz2 = f2mul(R360_F2(pl.fy), iz2);
        // jacobianWarpRt, RPI.h:970-984 (rows J0 = d col, J1 = d row)
        const float2 J02 = f2neg(f2mul(fxz2, px)), J12 = f2neg(f2mul(fyz2, py));
        const float2 J03 = f2mul(J02, py), J13 = f2neg(f2fma(f2mul(fyz2, py), py, R360_F2(pl.fy)));
        const float2 J04 = f2fma(f2mul(fxz2, px), px, R360_F2(pl.fx)), J14 = f2mul(f2mul(fyz2, px), py);
        const float2 J05 = f2neg(f2mul(fxz, py)), J15 = f2mul(fyz, px);
        float2 J[6];
        if (METHOD != R360_DEPTH_CONSISTENCY) {
            const float2 w = make_float2(h0 ? wp.x : 0.f, h1 ? wp.y : 0.f);     // selects, not products: a dropped pixel adds exact zeros whatever its weight
            const float2 ga = f2mul(w, make_float2(a1.x, b1.x)), gb = f2mul(w, make_float2(a1.y, b1.y));
            J[0] = f2mul(ga, fxz); J[1] = f2mul(gb, fyz);
            J[2] = f2fma(ga, J02, f2mul(gb, J12)); J[3] = f2fma(ga, J03, f2mul(gb, J13));
            J[4] = f2fma(ga, J04, f2mul(gb, J14)); J[5] = f2fma(ga, J05, f2mul(gb, J15));
            r360_accumulate(A, J, make_float2(h0 ? rp.x : 0.f, h1 ? rp.y : 0.f));
        }
        if (METHOD != R360_PHOTO_CONSISTENCY) {
            const float2 w = make_float2(h0 ? wd.x : 0.f, h1 ? wd.y : 0.f);     // 0 where the depth is not finite
            const float2 ga = f2mul(w, make_float2(a2.x, b2.x)), gb = f2mul(w, make_float2(a2.y, b2.y));
            // w ((Dx J0 + Dy J1) - jacobianRt_z),  jacobianRt_z = [0, 0, 1, py, -px, 0]   (RPI.h:1053)
            J[0] = f2mul(ga, fxz); J[1] = f2mul(gb, fyz);
            J[2] = f2add(f2fma(ga, J02, f2mul(gb, J12)), f2neg(w));
            J[3] = f2fma(f2neg(w), py, f2fma(ga, J03, f2mul(gb, J13)));
            J[4] = f2fma(w, px, f2fma(ga, J04, f2mul(gb, J14)));
            J[5] = f2fma(ga, J05, f2mul(gb, J15));
            r360_accumulate(A, J, make_float2(h0 ? rd.x : 0.f, h1 ? rd.y : 0.f));
        }
    }

    // ---- block reduction (as k_pin_eval): 27 normal-equation sums + PhotoResidual + DepthResidual, 3 counters
    float acc[R360_ACC_DOUBLES + 1];
    r360_acc_unpack(A, acc);                                                // fills 0..27 (27 = A.e2, unused here)
    acc[27] = sumP.x + sumP.y;
    acc[28] = sumD.x + sumD.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < R360_ACC_DOUBLES + 1; ++k) {
        float v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[wid][k] = v;
    }
    n_vis = __reduce_add_sync(0xffffffffu, n_vis);
    n_photo = __reduce_add_sync(0xffffffffu, n_photo);
    n_depth = __reduce_add_sync(0xffffffffu, n_depth);
    if (lane == 0) { s_cnt[wid][0] = n_vis; s_cnt[wid][1] = n_photo; s_cnt[wid][2] = n_depth; s_cnt[wid][3] = 0; }
    __syncthreads();
    if (threadIdx.x < R360_ACC_DOUBLES + 1) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
        r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
        atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
    }
}

// exp(update) * pose_estim with the FULL exponential (CPose3D::exp(v), RPI.h:4375)
__device__ void r360_pin_candidate(R360Pair* ps, const float upd[6]) {
    double ud[6], Td[16];
    for (int k = 0; k < 6; ++k) { ps->upd[k] = upd[k]; ud[k] = (double)upd[k]; }
    r360_se3_exp(ud, Td);
    float Tf[16], Tn[16];
    for (int k = 0; k < 16; ++k) Tf[k] = (float)Td[k];
    r360_mat4_mul(Tf, ps->pose_estim, Tn);
    for (int k = 0; k < 16; ++k) ps->pose_eval[k] = Tn[k];
}

// One step of the per-pair state machine of alignFrames after an evaluation.
//   phase 0: pose_estim at the start of the level;  phase 1: the Gauss-Newton candidate;
//   phase 2: the damped candidate of the LM retry (LM_maxIters = 1, RPI.h:4306).
__global__ void k_gn_step_pin(R360GnArgs g, int level) {
    const r360_params P = g.params;
    const int per_level = 2 * P.max_iters + 2;
    if (*g.n_active == 0) return;                                // every pair has left the level: the rest of its schedule is empty
    for (int p = threadIdx.x == 0 ? (int)blockIdx.x : g.n_pairs; p < g.n_pairs; p += gridDim.x) {   // one pair per warp, as k_gn_step
        R360Pair* ps = g.pairs + p;
        if (!ps->active) continue;
        double acc[R360_ACC_DOUBLES + 1];
        for (int k = 0; k < R360_ACC_DOUBLES + 1; ++k) acc[k] = r360_fx_get(g.acc + (size_t)p * R360_ACC_STRIDE, k);
        const int* cnt = g.cnt + (size_t)p * R360_ACC_INTS;
        const double n_d = (double)cnt[2];
        // avResidual is a float member (RPI.h:183): the value alignFrames compares has float precision
        const double err = (double)(float)(sqrt(acc[27] / n_d) + sqrt(acc[28] / n_d));    // RPI.h:768-771
        ps->passes[level] += 1;
        double diff_error;
        int accepted = 0;
        bool retry = false;
        if (ps->phase == 0) {
            diff_error = err;                                   // RPI.h:4318
            accepted = 1;
        } else {
            diff_error = ps->error - err;                       // RPI.h:4384 / 4411
            accepted = diff_error > 0;                          // RPI.h:4390 / 4415
            if (accepted) {
                if (ps->phase == 1) ps->lambda /= 10.0;         // RPI.h:4392 (the retry keeps lambda)
                for (int k = 0; k < 16; ++k) ps->pose_estim[k] = ps->pose_eval[k];
                ps->it += 1;
            } else if (ps->phase == 1 && diff_error < 0) {
                retry = true;                                   // RPI.h:4399-4402
            }
        }
        if (accepted) {
            ps->error = err; ps->err2 = acc[27] + acc[28]; ps->n_valid = cnt[2];
            for (int k = 0; k < 21; ++k) ps->Hc[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) ps->gc[k] = (float)acc[21 + k];
            ps->nvis_c = cnt[0];
        }
        if (g.trace && ps->ev < per_level) {
            r360_iter_record* rec = g.trace + ((size_t)p * P.n_levels + level) * per_level + ps->ev;
            rec->err2 = acc[27]; rec->err2_depth = acc[28]; rec->n_valid = cnt[1]; rec->n_valid_depth = cnt[2];
            rec->n_visible = cnt[0]; rec->level = level; rec->it = ps->it; rec->accepted = accepted; rec->used = 3;
            for (int k = 0; k < 16; ++k) rec->pose[k] = ps->pose_eval[k];
            for (int k = 0; k < 21; ++k) rec->hessian[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) rec->gradient[k] = (float)acc[21 + k];
            rec->pad = 0.f; rec->reserved = 0;
        }
        ps->ev += 1;
        r360_zero_acc(g.acc, g.cnt, p);

        if (retry) {
            // lambda *= step; update = -(H + lambda diag H)^-1 g with the H of the last calcHessGrad (RPI.h:4401-4404)
            ps->lambda *= 10.0;
            const float lam = (float)ps->lambda;
            float Hd[36], inv[36], upd[6];
            int q = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b, ++q) Hd[a + 6 * b] = Hd[b + 6 * a] = ps->Hl[q];
            for (int a = 0; a < 6; ++a) Hd[a + 6 * a] = Hd[a + 6 * a] + lam * Hd[a + 6 * a];
            r360_inverse6(Hd, inv);
            r360_solve_update(inv, ps->gl, upd);
            r360_pin_candidate(ps, upd);
            ps->phase = 2;
            continue;
        }
        // while (it < maxIters && update_pose.norm() > tol_update && diff_error > tol_residual)   RPI.h:4324
        const float* u = ps->upd;
        const float na = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
        const float nb = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
        const float unorm = sqrtf(na + nb);
        const bool go = ps->it < P.max_iters && (double)unorm > P.tol_update && diff_error > P.tol_residual;
        if (!go) {
            ps->iters[level] = ps->it;
            ps->active = 0;
            continue;
        }
        // loop body: calcHessGrad(pose_estim) == (Hc, gc)                     RPI.h:4340
        for (int k = 0; k < 21; ++k) ps->Hl[k] = ps->Hc[k];
        for (int k = 0; k < 6; ++k) ps->gl[k] = ps->gc[k];
        ps->nvis_l = ps->nvis_c;
        ps->lvl_l = level;
        float Hm[36], Hlam[36];
        {
            int q = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b, ++q) Hm[a + 6 * b] = Hm[b + 6 * a] = ps->Hc[q];
        }
        const float lam = (float)ps->lambda;
        for (int k = 0; k < 36; ++k) Hlam[k] = Hm[k];
        for (int a = 0; a < 6; ++a) Hlam[a + 6 * a] = Hm[a + 6 * a] + lam * Hm[a + 6 * a];
        if (r360_rank6(Hlam) != 6) {                            // RPI.h:4360-4368
            ps->status = R360_PAIR_ILL_POSED;
            ps->active = 0;
            continue;
        }
        float inv[36], upd[6];
        r360_inverse6(Hm, inv);
        r360_solve_update(inv, ps->gc, upd);                    // RPI.h:4371
        r360_pin_candidate(ps, upd);
        ps->phase = 1;
    }
    r360_compact_when_last(g);
}

// =========================================================================== the 8-sensor rig
// calcPhotoICPError_robot (RPI.h:4905-5092) + calcHessianGradient_robot (RPI.h:5100-5407) of ONE sensor of ONE pair at
// the pair's pose_eval, PHOTO_CONSISTENCY (the only method whose Hessian is defined upstream: the depth row uses a matrix
// that is never assigned, RPI.h:5372-5374).  blockIdx.y = 8 * active pair + sensor; all 8 sensors add into the pair's
// accumulators, which IS the driver's sum over the rig (RegisterRGBD360.h:403-440).  The two functions warp differently
// (one float matrix vs three matrix-vector products with double intrinsics) and so may hit different texels: both chains
// are evaluated, each with the reference's own operation sequence.
__device__ __forceinline__ int r360_round_d_to_int(double v) {
    const double r = round(v);
    return (r >= -2147483648.0 && r <= 2147483647.0) ? (int)r : INT_MIN;
}
__device__ __forceinline__ void r360_mat4_vec4(const float* M, const float* v, float* o) {    // Eigen 4x4 * 4x1, k ascending
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = ((M[i] * v[0] + M[i + 4] * v[1]) + M[i + 8] * v[2]) + M[i + 12] * v[3];
}
__global__ void __launch_bounds__(R360_PIN_THREADS)
k_rig_eval(R360PassArgs a, R360RigArgs rig) {
    __shared__ float s_red[R360_PIN_THREADS / 32][R360_ACC_DOUBLES + 1];
    __shared__ int s_cnt[R360_PIN_THREADS / 32][R360_ACC_INTS];
    __shared__ float s_tmp[16], s_rel[16];
    const int ap = blockIdx.y >> 3, sensor = blockIdx.y & 7;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    const float* Rt = rig.Rt[sensor];
    const float* Rti = rig.Rt_inv[sensor];
    // relPoseCam = Rt^-1 * pose * Rt (RPI.h:4924), float products summed k ascending (r360_mat4_mul's order)
    if (threadIdx.x < 16) {
        const int i = threadIdx.x & 3, j = threadIdx.x >> 2;
        const float* B = ps->pose_eval;
        float acc = Rti[i] * B[4 * j];
        acc = acc + Rti[i + 4] * B[1 + 4 * j];
        acc = acc + Rti[i + 8] * B[2 + 4 * j];
        acc = acc + Rti[i + 12] * B[3 + 4 * j];
        s_tmp[threadIdx.x] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int i = threadIdx.x & 3, j = threadIdx.x >> 2;
        float acc = s_tmp[i] * Rt[4 * j];
        acc = acc + s_tmp[i + 4] * Rt[1 + 4 * j];
        acc = acc + s_tmp[i + 8] * Rt[2 + 4 * j];
        acc = acc + s_tmp[i + 12] * Rt[3 + 4 * j];
        s_rel[threadIdx.x] = acc;
    }
    __syncthreads();
    float rel[16], T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { rel[k] = s_rel[k]; T[k] = ps->pose_eval[k]; }
    const float2* __restrict__ src = rig.src8[pair * 8 + sensor] + lv.px_off;
    const float2* __restrict__ trg = reinterpret_cast<const float2*>(rig.trg8[pair * 8 + sensor] + lv.px_off * R360_TEXEL_FLOATS);
    const double stdDevPhoto_inv = 1. / (double)P.std_photo;                // RPI.h:4927 / 5122

    float H[21], g[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) H[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = 0.f;
    float sumE = 0.f;
    int n_vis = 0, n_err = 0, n_photo = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lv.n; i += gridDim.x * blockDim.x) {
        const int r = (int)(((unsigned long long)i * lv.div_magic) >> 40), c = i - r * lv.cols;
        const float2 s = __ldg(&src[i]);                                   // {depth, gray}
        const float z = s.x;
        if (!(P.min_depth < z && z < P.max_depth)) continue;               // RPI.h:5018 / 5297
        // ---- calcPhotoICPError_robot: float back-projection, ONE matrix, double 1/z and pixel coordinates, no saliency test
        {
            float p[4] = { (c - rig.ox) * z * rig.inv_fx, (r - rig.oy) * z * rig.inv_fy, z, 1.f }, tp[4];
            r360_mat4_vec4(rel, p, tp);
            const double inv_z = 1.0 / (double)tp[2];
            const double tc = (double)(tp[0] * rig.fx) * inv_z + (double)rig.ox;
            const double tr = (double)(tp[1] * rig.fy) * inv_z + (double)rig.oy;
            const int ri = r360_round_d_to_int(tr), ci = r360_round_d_to_int(tc);
            if ((unsigned)ri < (unsigned)lv.rows && (unsigned)ci < (unsigned)lv.cols) {
                const float photoDiff = __ldg(trg + 3u * (unsigned)(ri * lv.cols + ci)).x - s.y;
                const double weight_photo = (double)r360_huber(photoDiff, P.std_photo) * stdDevPhoto_inv;
                const float werr = (float)(weight_photo * (double)photoDiff);
                sumE += werr * werr;
                ++n_err;
            }
        }
        // ---- calcHessianGradient_robot: double intrinsics, three matrix-vector products, photo saliency `continue`
        float p[4] = { (float)(((double)c - rig.dox) * (double)z * rig.dinv_fx), (float)(((double)r - rig.doy) * (double)z * rig.dinv_fy), z, 1.f };
        float p1[4], p2[4], tp[4];
        r360_mat4_vec4(Rt, p, p1);
        r360_mat4_vec4(T, p1, p2);
        r360_mat4_vec4(Rti, p2, tp);
        const double inv_z = 1.0 / (double)tp[2];
        const double tc = ((double)tp[0] * rig.dfx) * inv_z + rig.dox;
        const double tr = ((double)tp[1] * rig.dfy) * inv_z + rig.doy;
        const int ri = r360_round_d_to_int(tr), ci = r360_round_d_to_int(tc);
        if (!((unsigned)ri < (unsigned)lv.rows && (unsigned)ci < (unsigned)lv.cols)) continue;
        ++n_vis;
        const float2* tx = trg + 3u * (unsigned)(ri * lv.cols + ci);
        const float2 t0 = __ldg(tx), t1 = __ldg(tx + 1);                   // {gray, depth}, {Ix, Iy}
        if ((fabsf(t1.x) < P.thres_sal_int) & (fabsf(t1.y) < P.thres_sal_int)) continue;      // RPI.h:5354-5355
        const float x = p2[0], y = p2[1], zz = p2[2];
        const float P00 = (float)(rig.dfx * inv_z), P11 = (float)(rig.dfy * inv_z);
        const float P02 = (float)(-rig.dfx * (double)tp[0] * inv_z * inv_z), P12 = (float)(-rig.dfy * (double)tp[1] * inv_z * inv_z);
        float W0[6], W1[6];                                                // jacobianWarpRt = jacobianProj23 * (Rt^-1(3x3) * [I | -skew(p2)])
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            float Tq[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float R0 = Rti[k], R1 = Rti[k + 4], R2 = Rti[k + 8];
                Tq[k] = q == 0 ? (R0 * 1.f + R1 * 0.f) + R2 * 0.f
                      : q == 1 ? (R0 * 0.f + R1 * 1.f) + R2 * 0.f
                      : q == 2 ? (R0 * 0.f + R1 * 0.f) + R2 * 1.f
                      : q == 3 ? (R0 * 0.f + R1 * (-zz)) + R2 * y
                      : q == 4 ? (R0 * zz + R1 * 0.f) + R2 * (-x)
                               : (R0 * (-y) + R1 * x) + R2 * 0.f;
            }
            W0[q] = (P00 * Tq[0] + 0.f * Tq[1]) + P02 * Tq[2];
            W1[q] = (0.f * Tq[0] + P11 * Tq[1]) + P12 * Tq[2];
        }
        const float photoDiff = t0.x - s.y;
        const double weight_photo = (double)r360_huber(photoDiff, P.std_photo) * stdDevPhoto_inv;
        const double weightedErrorPhoto = weight_photo * (double)photoDiff;
        const float wf = (float)weight_photo;
        const float a0 = wf * t1.x, a1 = wf * t1.y;
        const float rf = (float)weightedErrorPhoto;
        float J[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) J[q] = a0 * W0[q] + a1 * W1[q];
        ++n_photo;
        int q = 0;
#pragma unroll
        for (int xx = 0; xx < 6; ++xx) {
#pragma unroll
            for (int yy = xx; yy < 6; ++yy, ++q) H[q] = fmaf(J[xx], J[yy], H[q]);
            g[xx] = fmaf(J[xx], rf, g[xx]);
        }
    }
    // ---- block reduction: 27 normal-equation sums + the error sum, 3 counters
    float acc[R360_ACC_DOUBLES + 1];
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[k] = H[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[21 + k] = g[k];
    acc[27] = sumE;
    acc[28] = 0.f;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < R360_ACC_DOUBLES; ++k) {
        float v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[wid][k] = v;
    }
    n_vis = __reduce_add_sync(0xffffffffu, n_vis);
    n_err = __reduce_add_sync(0xffffffffu, n_err);
    n_photo = __reduce_add_sync(0xffffffffu, n_photo);
    if (lane == 0) { s_cnt[wid][0] = n_vis; s_cnt[wid][1] = n_err; s_cnt[wid][2] = n_photo; s_cnt[wid][3] = 0; }
    __syncthreads();
    if (threadIdx.x < R360_ACC_DOUBLES) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
        r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < R360_PIN_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
        atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
    }
}

// -(H + lambda diag H)^-1 g, then exp(update) * pose_estim -> `cand` (RegisterRGBD360.h:453-455, 478-480)
__device__ void r360_rig_candidate(R360Pair* ps, float* cand) {
    const float lam = (float)ps->lambda;
    float Hd[36], inv[36], upd[6];
    int q = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++q) Hd[a + 6 * b] = Hd[b + 6 * a] = ps->Hl[q];
    for (int a = 0; a < 6; ++a) Hd[a + 6 * a] = Hd[a + 6 * a] + lam * Hd[a + 6 * a];
    r360_inverse6(Hd, inv);
    r360_solve_update(inv, ps->gl, upd);
    double ud[6], Td[16];
    for (int k = 0; k < 6; ++k) { ps->upd[k] = upd[k]; ud[k] = (double)upd[k]; }
    r360_se3_exp(ud, Td);
    float Tf[16];
    for (int k = 0; k < 16; ++k) Tf[k] = (float)Td[k];
    r360_mat4_mul(Tf, ps->pose_estim, cand);
}

// The state machine of RegisterRGBD360::RegisterDensePhotoICP after an evaluation (phases as k_gn_step_pin).  The
// candidate pose_estim_temp is kept in Hc (16 of its 21 floats: the rig path does not use Hc otherwise); with
// g.rig_faithful the pass that follows evaluates pose_estim again, as upstream does (RegisterRGBD360.h:462, 488): the sums
// are bit-reproducible, so diff_error is exactly 0, the candidate is never taken and every level runs one loop body.
__global__ void k_gn_step_rig(R360GnArgs g, int level) {
    const r360_params P = g.params;
    if (*g.n_active == 0) return;                                // every pair has left the level: the rest of its schedule is empty
    for (int p = threadIdx.x == 0 ? (int)blockIdx.x : g.n_pairs; p < g.n_pairs; p += gridDim.x) {   // one pair per warp, as k_gn_step
        R360Pair* ps = g.pairs + p;
        if (!ps->active) continue;
        double acc[R360_ACC_DOUBLES + 1];
        for (int k = 0; k < R360_ACC_DOUBLES + 1; ++k) acc[k] = r360_fx_get(g.acc + (size_t)p * R360_ACC_STRIDE, k);
        const int* cnt = g.cnt + (size_t)p * R360_ACC_INTS;
        const double err = acc[27];                             // the SUM of squared weighted residuals over the 8 sensors
        ps->passes[level] += 1;
        double diff_error;
        bool retry = false;
        const bool evaluated_candidate = !g.rig_faithful;       // else: pose_estim again
        auto take_h = [&]() {
            for (int k = 0; k < 21; ++k) ps->Hl[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) ps->gl[k] = (float)acc[21 + k];
            ps->nvis_l = cnt[0];
        };
        if (ps->phase == 0) {
            diff_error = err;                                   // RegisterRGBD360.h:412
            ps->error = err; ps->err2 = err; ps->n_valid = cnt[1];
            take_h();                                           // calcHessianGradient_robot(pose_estim) of the first loop body
        } else {
            diff_error = ps->error - err;                       // :464 / :489
            if (diff_error > 0) {
                if (ps->phase == 1) ps->lambda /= 10.0;         // :467 (the retry keeps lambda)
                for (int k = 0; k < 16; ++k) ps->pose_estim[k] = ps->Hc[k];     // pose_estim = pose_estim_temp
                ps->error = err; ps->err2 = err; ps->n_valid = cnt[1];
                ps->it += 1;
                if (evaluated_candidate) take_h();              // the pass ran at the new pose_estim: its H serves the next loop body
            } else if (ps->phase == 1 && diff_error < 0) {
                retry = true;                                   // :474
            }
        }
        r360_zero_acc(g.acc, g.cnt, p);
        if (retry) {
            ps->lambda *= 10.0;                                 // :476
            r360_rig_candidate(ps, ps->Hc);
            for (int k = 0; k < 16; ++k) ps->pose_eval[k] = evaluated_candidate ? ps->Hc[k] : ps->pose_estim[k];
            ps->phase = 2;
            continue;
        }
        const float* u = ps->upd;
        const float na = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
        const float nb = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
        const float unorm = sqrtf(na + nb);
        const bool go = ps->it < P.max_iters && (double)unorm > P.tol_update && diff_error > P.tol_residual;   // :421
        if (!go) {
            ps->iters[level] = ps->it;
            ps->active = 0;
            continue;
        }
        // loop body: Hessian / Gradient summed over the sensors at pose_estim == (Hl, gl)        :424-440
        ps->lvl_l = level;
        float Hm[36];
        {
            int q = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b, ++q) Hm[a + 6 * b] = Hm[b + 6 * a] = ps->Hl[q];
        }
        const float lam = (float)ps->lambda;
        for (int a = 0; a < 6; ++a) Hm[a + 6 * a] = Hm[a + 6 * a] + lam * Hm[a + 6 * a];
        if (r360_rank6(Hm) != 6) {                              // :443-450: returns false with rigidTransf = pose_estim
            ps->status = R360_PAIR_ILL_POSED;
            ps->active = 0;
            continue;
        }
        r360_rig_candidate(ps, ps->Hc);
        for (int k = 0; k < 16; ++k) ps->pose_eval[k] = evaluated_candidate ? ps->Hc[k] : ps->pose_estim[k];
        ps->phase = 1;
    }
    r360_compact_when_last(g);
}

void r360_launch_rig_eval(cudaStream_t st, const R360PassArgs& a, const R360RigArgs& rig, int n_pairs, int sm_count) {
    long long blocks = ((long long)a.lv.n + R360_PIN_THREADS - 1) / R360_PIN_THREADS;
    long long cap = (long long)R360_PIN_CAP * sm_count / (8LL * (n_pairs > 0 ? n_pairs : 1));
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const dim3 grid((unsigned)blocks, (unsigned)(8 * n_pairs));
    k_rig_eval<<<grid, R360_PIN_THREADS, 0, st>>>(a, rig);
}
void r360_launch_gn_step_rig(cudaStream_t st, const R360GnArgs& g, int level) {
    k_gn_step_rig<<<r360_blocks(g.n_pairs, 1, 1024), 32, 0, st>>>(g, level);
}

void r360_launch_pin_eval(cudaStream_t st, const R360PassArgs& a, const R360PinLevel& pl, int n_pairs, int sm_count, bool packed) {
    long long blocks = ((long long)a.lv.n + R360_PIN_THREADS - 1) / R360_PIN_THREADS;
    long long cap = (long long)R360_PIN_CAP * sm_count / (n_pairs > 0 ? n_pairs : 1);
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const dim3 grid((unsigned)blocks, (unsigned)n_pairs);
    if (packed && (a.lv.cols & 1) == 0) {
        switch (a.params.method) {
            case R360_PHOTO_CONSISTENCY: k_pin_eval2<R360_PHOTO_CONSISTENCY><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
            case R360_DEPTH_CONSISTENCY: k_pin_eval2<R360_DEPTH_CONSISTENCY><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
            default: k_pin_eval2<R360_PHOTO_DEPTH><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
        }
        return;
    }
    switch (a.params.method) {
        case R360_PHOTO_CONSISTENCY: k_pin_eval<R360_PHOTO_CONSISTENCY><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
        case R360_DEPTH_CONSISTENCY: k_pin_eval<R360_DEPTH_CONSISTENCY><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
        default: k_pin_eval<R360_PHOTO_DEPTH><<<grid, R360_PIN_THREADS, 0, st>>>(a, pl); break;
    }
}
void r360_launch_gn_step_pin(cudaStream_t st, const R360GnArgs& g, int level) {
    k_gn_step_pin<<<r360_blocks(g.n_pairs, 1, 1024), 32, 0, st>>>(g, level);
}
