// r360_occ.cu -- the occlusion variants of the spherical registration:
//   errorPhotoICP_sphereOcc1 / calcHessGrad_sphereOcc1   (RPI.h:3232-3369, 3373-3716)   occlusion = 1
//   errorPhotoICP_sphereOcc2 / calcHessGrad_sphereOcc2   (RPI.h:3720-3858, 3861-4249)   occlusion = 2
//
// Upstream these loops run `#pragma omp parallel for` over a z-buffer that every iteration reads
// and writes: the result depends on the thread schedule.  What is implemented here -- and pinned by
// the compiled reference at one OpenMP thread, tests/test_occlusion.py -- is the semantics of ONE
// thread (source pixels in index order), evaluated deterministically and in parallel:
//
//   * a z-test passer is a source pixel whose 1/|p| is >= that of every EARLIER candidate of the
//     same target texel (the buffer always holds the running maximum);
//   * "stored per target texel, last writer wins" = the candidate with the largest source index
//     (Occ2 H/g), resp. the last passer = the latest occurrence of the maximal 1/|p| (Occ1 error).
//
// Two kernels per evaluated pose instead of the fused k_pass:
//   k_occ_scatter  builds, per target texel, the linked list of its candidate source pixels
//                  (head[texel] / next[pixel], atomicExch; list ORDER is arbitrary, list CONTENT is not)
//                  and stores 1/|p| per candidate;
//   k_occ_eval     every candidate walks its texel's list (1-3 entries) to decide passer / last /
//                  winner, then accumulates the error sums (PhotoResidual, DepthResidual, counters)
//                  and the J^T W J / J^T W r rows of ITS variant.  Same packed index path, Jacobian
//                  rows and block reduction as k_pass; the reference evaluates the error and, when a
//                  step is accepted, the Hessian at the same pose, so one evaluation yields both.
#include <algorithm>
#include <cstdlib>
#include "r360_device.cuh"
#include "r360_kernels.h"

#define R360_OCC_THREADS 256
#define R360_OCC_GATE 0.3f                     // thresDepthOutliers, RPI.h:4525

namespace {

struct OccPixelPair {
    R360SrcPair sp;
    R360Geo2 g;
    int ii[2];          // target texel index r' * cols + c' (valid when cand)
    bool inb[2];        // LUT point valid and (r', c') inside the image (RPI.h:3292)
    float2 ta[3], tb[3];
};

}  // namespace

// {depth, gray} x 2 of source pixel pair i (zeros past the end of the level); the kernels load it one
// iteration ahead of its use.
__device__ __forceinline__ float4 r360_occ_src(const float4* __restrict__ src4, int i, int n) {
    return i < n ? __ldg(&src4[i >> 1]) : make_float4(0.f, 0.f, 0.f, 0.f);
}
// Source pixel pair -> warped geometry, target texel, its six floats.  All 32 lanes must call.
// TEX: 2 = all six floats of the texel (k_occ_eval), 1 = {gray, depth} only (the Occ2 gate of k_occ_scatter),
// 0 = none (k_occ_scatter of Occ1 needs the texel index alone).
template <int OCC, int TEX = 2>
__device__ __forceinline__ void r360_occ_pixel_pair(const R360PassArgs& a, const R360Level& lv, const r360_params& P,
                                                    const float* T, const float* Tg, float4 s,
                                                    const float2* __restrict__ trg, int i, OccPixelPair& o) {
    const bool in0 = i < lv.n, in1 = i + 1 < lv.n;
    const int r = in0 ? (int)(((unsigned long long)i * lv.div_magic) >> 40) : 0;
    const int c = in0 ? i - r * lv.cols : 0;
    r360_load_src_pair(lv, P, s, r, c, in0, in1, o.sp);
    int rr[2], cc[2];
    unsigned n_fb = 0;
    r360_index_pair(T, Tg, lv, o.sp, a.one, o.g, rr, cc, n_fb);
    o.inb[0] = o.sp.v0 && (unsigned)rr[0] < (unsigned)lv.rows && (unsigned)cc[0] < (unsigned)lv.cols;
    o.inb[1] = o.sp.v1 && (unsigned)rr[1] < (unsigned)lv.rows && (unsigned)cc[1] < (unsigned)lv.cols;
    o.ii[0] = o.inb[0] ? rr[0] * lv.cols + cc[0] : 0;
    o.ii[1] = o.inb[1] ? rr[1] * lv.cols + cc[1] : 0;
    const float2* tx0 = trg + 3u * (unsigned)o.ii[0];
    const float2* tx1 = trg + 3u * (unsigned)o.ii[1];
    const float2 z2 = make_float2(0.f, 0.f);
    o.ta[0] = TEX >= 1 ? __ldg(tx0) : z2; o.ta[1] = TEX >= 2 ? __ldg(tx0 + 1) : z2; o.ta[2] = TEX >= 2 ? __ldg(tx0 + 2) : z2;
    o.tb[0] = TEX >= 1 ? __ldg(tx1) : z2; o.tb[1] = TEX >= 2 ? __ldg(tx1 + 1) : z2; o.tb[2] = TEX >= 2 ? __ldg(tx1 + 2) : z2;
}

// Candidate test: Occ1 every in-bounds pixel; Occ2 those within the 0.3 m gate of the target depth
// (RPI.h:3788-3791; a NaN difference is NOT an outlier upstream: fabs(NaN) > t is false).
template <int OCC>
__device__ __forceinline__ bool r360_occ_candidate(bool inb, float depth2, float dist) {
    if (OCC == 2) return inb && !(fabsf(depth2 - dist) > R360_OCC_GATE);
    return inb;
}

// Empties the candidate lists of the ACTIVE pairs (a memset over the whole batch would also pay for the pairs that
// have already converged: every level enqueues max_iters + 1 evaluations whatever the number of active pairs).
__global__ void __launch_bounds__(R360_OCC_THREADS)
k_occ_reset(R360PassArgs a, int* __restrict__ head) {
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    int2* h = reinterpret_cast<int2*>(head + (size_t)ap * a.lv.n);          // lv.n is even at every level
    const int n2 = a.lv.n >> 1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += gridDim.x * blockDim.x) h[q] = make_int2(-1, -1);
}

template <int OCC>
__global__ void __launch_bounds__(R360_OCC_THREADS)
k_occ_scatter(R360PassArgs a, int* __restrict__ head, int* __restrict__ next, float* __restrict__ dinv) {
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    float T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float4* src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    const float2* trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    int* hd = head + (size_t)ap * lv.n;
    int* nx = next + (size_t)ap * lv.n;
    float* dv = dinv + (size_t)ap * lv.n;
    const int stride = 2 * gridDim.x * blockDim.x;
    float4 s_nxt = r360_occ_src(src4, 2 * (blockIdx.x * blockDim.x + threadIdx.x), lv.n);
    for (int base = 0; base < lv.n; base += stride) {           // warp-uniform trip count (warp votes inside)
        const int i = base + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
        const float4 s_cur = s_nxt;
        s_nxt = r360_occ_src(src4, i + stride, lv.n);           // in flight across this iteration's atomics
        OccPixelPair o;
        r360_occ_pixel_pair<OCC, OCC == 2 ? 1 : 0>(a, lv, P, T, ps->pose_eval, s_cur, trg, i, o);
        // both exchanges are issued before either result is stored: two round trips to L2 in flight instead of one after
        // the other (same thread, so two candidates of one texel still enter its list in pixel order)
        const bool c0 = r360_occ_candidate<OCC>(o.inb[0], o.ta[0].y, o.g.dist.x);
        const bool c1 = r360_occ_candidate<OCC>(o.inb[1], o.tb[0].y, o.g.dist.y);
        int p0 = -1, p1 = -1;
        if (c0) p0 = atomicExch(&hd[o.ii[0]], i);
        if (c1) p1 = atomicExch(&hd[o.ii[1]], i + 1);
        if (c0) { dv[i] = o.g.dinv.x; nx[i] = p0; }
        if (c1) { dv[i + 1] = o.g.dinv.y; nx[i + 1] = p1; }
    }
}

// Walks the candidate list of target texel `ii` on behalf of candidate `i`:
//   (di = the candidate's own 1/|p|: k_occ_scatter stored the result of the same packed sequence in dv[i])
//   passer : no earlier candidate (smaller source index) has a larger 1/|p|        (RPI.h:3299-3301)
//   last   : no later candidate exists                                              (last writer of a per-texel slot)
//   winner : passer and no later candidate passes (none has 1/|p| >= ours)          (last passer)
__device__ __forceinline__ void r360_occ_walk(const int* __restrict__ hd, const int* __restrict__ nx,
                                              const float* __restrict__ dv, int ii, int i, float di, int nx_i,
                                              bool& passer, bool& last, bool& winner) {
    passer = true; last = true; winner = true;
    int j = hd[ii];
    while (j >= 0) {
        if (j == i) { j = nx_i; continue; }                      // nx_i == nx[i], loaded by the caller
        const float dj = dv[j];
        if (j < i) { if (dj > di) passer = false; }
        else { last = false; if (dj >= di) winner = false; }
        j = nx[j];
    }
    winner = winner && passer;
}

// Straightforward form (one pixel pair per thread and iteration, direct gathers), kept as the A/B baseline and
// fallback of the pipelined kernel below: R360_OCC_PIPE=0 in the environment selects it.
// Same walk with the list head already loaded (j0 = hd[texel of i]).
__device__ __forceinline__ void r360_occ_walk_from(int j0, const int* __restrict__ nx, const float* __restrict__ dv, int i,
                                                   float di, int nx_i, bool& passer, bool& last, bool& winner) {
    passer = true; last = true; winner = true;
    int j = j0;
    while (j >= 0) {
        if (j == i) { j = nx_i; continue; }
        const float dj = dv[j];
        if (j < i) { if (dj > di) passer = false; }
        else { last = false; if (dj >= di) winner = false; }
        j = nx[j];
    }
    winner = winner && passer;
}

// Pipelined form of the evaluation, the structure of k_pass (r360_kernels.cu): per thread and iteration one
// pixel pair in two stages through shared memory.
//   stage A (pair k+1): source pair (loaded two iterations ahead), packed pinned index, then the six 8-byte texel
//            gathers, the two list heads hd[texel] and the pair's own links nx[i], nx[i+1] are issued as cp.async
//            into the thread's slot; the warped geometry is parked next to them;
//   stage B (pair k):   cp.async.wait_group, the candidate / passer / last / winner decisions (the list walk only
//            touches global memory again for texels with more than one candidate), error sums, rows, accumulation.
// Slot per thread and stage: 48 B texels | 48 B geometry | 16 B {hd0, hd1, nx_i, nx_i+1}.
#define R360_OCC_STAGES 2
#define R360_OCC_SLOT_BYTES 112
#define R360_OCC_DYN_SMEM (R360_OCC_STAGES * R360_OCC_THREADS * R360_OCC_SLOT_BYTES)
template <int METHOD, int OCC>
__global__ void __launch_bounds__(R360_OCC_THREADS, 2)
k_occ_eval(R360PassArgs a, const int* __restrict__ head, const int* __restrict__ next, const float* __restrict__ dinv) {
    extern __shared__ float4 s_pipe[];
    __shared__ float s_red[R360_OCC_THREADS / 32][R360_ACC_DOUBLES + 1];
    __shared__ int s_cnt[R360_OCC_THREADS / 32][R360_ACC_INTS];
    __shared__ __align__(16) float s_T[16];
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    if (threadIdx.x < 16) s_T[threadIdx.x] = ps->pose_eval[threadIdx.x];
    __syncthreads();
    const float4* __restrict__ src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    const float2* __restrict__ trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    const int* __restrict__ hd = head + (size_t)ap * lv.n;
    const int* __restrict__ nx = next + (size_t)ap * lv.n;
    const float* __restrict__ dv = dinv + (size_t)ap * lv.n;

    constexpr unsigned STAGE_BYTES = R360_OCC_THREADS * R360_OCC_SLOT_BYTES;
    constexpr unsigned GEO_OFF = R360_OCC_THREADS * 48, LIST_OFF = R360_OCC_THREADS * 96;
    const unsigned slot_tex = r360_smem_addr(reinterpret_cast<char*>(s_pipe) + 48 * threadIdx.x);
    const unsigned slot_list = r360_smem_addr(reinterpret_cast<char*>(s_pipe) + LIST_OFF + 16 * threadIdx.x);

    R360Acc2 A;
    r360_acc_zero(A);
    float sumP = 0.f, sumD = 0.f;
    int n_vis = 0, n_photo = 0, n_depth = 0;
    const int stride = 2 * gridDim.x * blockDim.x;
    const int n_it = (lv.n + stride - 1) / stride;               // uniform over the grid (warp votes inside)
    int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);         // pixel pair of the next stage A
    int i_b = i;                                                 // pixel pair of the next stage B
    float4 s_cur = r360_occ_src(src4, i, lv.n);
    float4 s_nxt = r360_occ_src(src4, i + stride, lv.n);

    auto stage_a = [&](unsigned st) {
        const bool in0 = i < lv.n, in1 = i + 1 < lv.n;
        const int r = in0 ? (int)(((unsigned long long)i * lv.div_magic) >> 40) : 0;
        const int c = in0 ? i - r * lv.cols : 0;
        R360SrcPair sp;
        r360_load_src_pair(lv, P, s_cur, r, c, in0, in1, sp);
        float T[16];
        {
            const float4 c0 = reinterpret_cast<const float4*>(s_T)[0], c1 = reinterpret_cast<const float4*>(s_T)[1],
                         c2 = reinterpret_cast<const float4*>(s_T)[2], c3 = reinterpret_cast<const float4*>(s_T)[3];
            T[0] = c0.x; T[1] = c0.y; T[2] = c0.z; T[4] = c1.x; T[5] = c1.y; T[6] = c1.z;
            T[8] = c2.x; T[9] = c2.y; T[10] = c2.z; T[12] = c3.x; T[13] = c3.y; T[14] = c3.z;
        }
        R360Geo2 g;
        int rr[2], cc[2];
        unsigned n_fb = 0;
        r360_index_pair(T, s_T, lv, sp, a.one, g, rr, cc, n_fb);
        const bool ok0 = sp.v0 & ((unsigned)rr[0] < (unsigned)lv.rows) & ((unsigned)cc[0] < (unsigned)lv.cols);   // RPI.h:3292
        const bool ok1 = sp.v1 & ((unsigned)rr[1] < (unsigned)lv.rows) & ((unsigned)cc[1] < (unsigned)lv.cols);
        const unsigned ii0 = ok0 ? (unsigned)(rr[0] * lv.cols + cc[0]) : 0u;
        const unsigned ii1 = ok1 ? (unsigned)(rr[1] * lv.cols + cc[1]) : 0u;
        const float2* tx0 = trg + 3u * ii0;
        const float2* tx1 = trg + 3u * ii1;
        const unsigned dst = slot_tex + st;
        r360_cp_async8(dst + 0, tx0); r360_cp_async8(dst + 8, tx0 + 1); r360_cp_async8(dst + 16, tx0 + 2);
        r360_cp_async8(dst + 24, tx1); r360_cp_async8(dst + 32, tx1 + 1); r360_cp_async8(dst + 40, tx1 + 2);
        const unsigned dl = slot_list + st;
        r360_cp_async4(dl + 0, hd + ii0);
        r360_cp_async4(dl + 4, hd + ii1);
        r360_cp_async8(dl + 8, nx + min(i, lv.n - 2));           // i and lv.n are even; tail lanes load a link they never use
        r360_cp_async_commit();
        r360_sts128(dst + GEO_OFF, make_float4(g.px.x, g.px.y, g.py.x, g.py.y));
        r360_sts128(dst + GEO_OFF + 16, make_float4(g.pz.x, g.pz.y, g.dinv.x, g.dinv.y));
        // |p| > 0: its sign carries the in-bounds flag of the pixel
        r360_sts128(dst + GEO_OFF + 32, make_float4(sp.Is.x, sp.Is.y, ok0 ? g.dist.x : -g.dist.x, ok1 ? g.dist.y : -g.dist.y));
        i += stride;
    };

    constexpr unsigned RING_END = R360_OCC_STAGES * STAGE_BYTES;
    stage_a(0u);
    unsigned st_b = 0u, st_a = STAGE_BYTES;
    for (int k = 0; k < n_it; ++k) {
        if (k + 1 < n_it) {
            s_cur = s_nxt;
            s_nxt = r360_occ_src(src4, i + stride, lv.n);
            stage_a(st_a);
        } else {
            r360_cp_async_commit();
        }
        st_a += STAGE_BYTES;
        if (st_a == RING_END) st_a = 0u;
        r360_cp_async_wait<1>();
        // ---- stage B: pixel pair i_b
        const unsigned rd = slot_tex + st_b, rl = slot_list + st_b;
        st_b += STAGE_BYTES;
        if (st_b == RING_END) st_b = 0u;
        const float4 q0 = r360_lds128(rd), q1 = r360_lds128(rd + 16), q2 = r360_lds128(rd + 32);
        const float4 g0 = r360_lds128(rd + GEO_OFF), g1 = r360_lds128(rd + GEO_OFF + 16), g2 = r360_lds128(rd + GEO_OFF + 32);
        const float4 lf = r360_lds128(rl);
        const int hd_v[2] = { __float_as_int(lf.x), __float_as_int(lf.y) };
        const int nx_own[2] = { __float_as_int(lf.z), __float_as_int(lf.w) };
        const float2 ta[3] = { make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y) };
        const float2 tb[3] = { make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w) };
        R360Geo2 g;
        g.px = make_float2(g0.x, g0.y); g.py = make_float2(g0.z, g0.w);
        g.pz = make_float2(g1.x, g1.y); g.dinv = make_float2(g1.z, g1.w);
        g.dist = make_float2(fabsf(g2.z), fabsf(g2.w));
        g.rho2 = f2fma(g.py, g.py, f2mul(g.pz, g.pz));
        const bool inb[2] = { g2.z > 0.f, g2.w > 0.f };
        bool okH[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float2* t = q ? tb : ta;
            const float dist = q ? g.dist.y : g.dist.x;
            const float Is = q ? g2.y : g2.x;
            const bool cand = r360_occ_candidate<OCC>(inb[q], t[0].y, dist);
            bool passer = false, last = false, winner = false;
            if (cand) r360_occ_walk_from(hd_v[q], nx, dv, i_b + q, q ? g.dinv.y : g.dinv.x, nx_own[q], passer, last, winner);
            // ---- error function of the variant (see k_occ_eval_simple for the line citations)
            const bool cnt_on = passer;
            const bool sum_on = OCC == 1 ? winner : passer;
            const bool photo_sal = !((fabsf(t[1].x) < P.thres_sal_int) & (fabsf(t[1].y) < P.thres_sal_int));
            const bool depth_ok = (fabsf(t[0].y) < INFINITY) &&
                                  !((fabsf(t[2].x) < P.thres_sal_depth) & (fabsf(t[2].y) < P.thres_sal_depth));
            if (OCC == 2 && cnt_on) ++n_depth;
            bool reach = true;
            if (METHOD != R360_DEPTH_CONSISTENCY) {
                reach = photo_sal;
                if (photo_sal) {
                    if (OCC == 1 && cnt_on) ++n_photo;
                    if (sum_on) { const float r = r360_wres_photo(t[0].x, Is, P, a.inv_std_photo); sumP += r * r; }
                }
            }
            if (METHOD != R360_PHOTO_CONSISTENCY) {
                if (reach && depth_ok) {
                    if (OCC == 1 && cnt_on) ++n_depth;
                    if (sum_on) { const float r = r360_wres_depth(t[0].y, dist, P); sumD += r * r; }
                }
            }
            okH[q] = OCC == 1 ? inb[q] : last;
            n_vis += okH[q] ? 1 : 0;
        }
        r360_rows_pair<METHOD, 1>(g, lv.res_inv, make_float2(g2.x, g2.y), ta, tb, okH[0], okH[1], P, a.inv_std_photo, A);
        i_b += stride;
    }

    // ---- block reduction: 27 normal-equation sums + PhotoResidual + DepthResidual, 3 counters
    float acc[R360_ACC_DOUBLES + 1];
    r360_acc_unpack(A, acc);
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
        for (int k = 0; k < R360_OCC_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
        r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < R360_OCC_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
        atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
    }
}

template <int METHOD, int OCC>
__global__ void __launch_bounds__(R360_OCC_THREADS)
k_occ_eval_simple(R360PassArgs a, const int* __restrict__ head, const int* __restrict__ next, const float* __restrict__ dinv) {
    __shared__ float s_red[R360_OCC_THREADS / 32][R360_ACC_DOUBLES + 1];
    __shared__ int s_cnt[R360_OCC_THREADS / 32][R360_ACC_INTS];
    const int ap = blockIdx.y;
    if (ap >= *a.n_active) return;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const int pair = a.active_list[ap];
    const R360Pair* ps = a.pairs + pair;
    float T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float4* src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    const float2* trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    const int* hd = head + (size_t)ap * lv.n;
    const int* nx = next + (size_t)ap * lv.n;
    const float* dv = dinv + (size_t)ap * lv.n;

    R360Acc2 A;
    r360_acc_zero(A);
    float sumP = 0.f, sumD = 0.f;
    int n_vis = 0, n_photo = 0, n_depth = 0;
    const int stride = 2 * gridDim.x * blockDim.x;
    float4 s_nxt = r360_occ_src(src4, 2 * (blockIdx.x * blockDim.x + threadIdx.x), lv.n);
    for (int base = 0; base < lv.n; base += stride) {
        const int i = base + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
        const float4 s_cur = s_nxt;
        s_nxt = r360_occ_src(src4, i + stride, lv.n);
        // the pixels' own list links, needed by the walk when they head their texel's list (the common
        // single-candidate case): loaded up front instead of after the head (i and lv.n are even)
        const int2 nx_own = i < lv.n ? __ldg(reinterpret_cast<const int2*>(nx + i)) : make_int2(-1, -1);
        OccPixelPair o;
        r360_occ_pixel_pair<OCC>(a, lv, P, T, ps->pose_eval, s_cur, trg, i, o);
        bool okH[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float2* t = q ? o.tb : o.ta;
            const float dist = q ? o.g.dist.y : o.g.dist.x;
            const float Is = q ? o.sp.Is.y : o.sp.Is.x;
            const bool cand = r360_occ_candidate<OCC>(o.inb[q], t[0].y, dist);
            bool passer = false, last = false, winner = false;
            if (cand) r360_occ_walk(hd, nx, dv, o.ii[q], i + q, q ? o.g.dinv.y : o.g.dinv.x, q ? nx_own.y : nx_own.x, passer, last, winner);
            // ---- error function of the variant
            const bool cnt_on = passer;                                   // counters count every passer
            const bool sum_on = OCC == 1 ? winner : passer;               // Occ1 sums one value per texel (RPI.h:3318)
            const bool photo_sal = !((fabsf(t[1].x) < P.thres_sal_int) & (fabsf(t[1].y) < P.thres_sal_int));
            const bool depth_ok = (fabsf(t[0].y) < INFINITY) &&
                                  !((fabsf(t[2].x) < P.thres_sal_depth) & (fabsf(t[2].y) < P.thres_sal_depth));
            if (OCC == 2 && cnt_on) ++n_depth;                            // nValidDepthPts, RPI.h:3798
            bool reach = true;                                            // the photo `continue` precedes the depth term
            if (METHOD != R360_DEPTH_CONSISTENCY) {
                reach = photo_sal;
                if (photo_sal) {
                    if (OCC == 1 && cnt_on) ++n_photo;                    // nValidPhotoPts, RPI.h:3319
                    if (sum_on) { const float r = r360_wres_photo(t[0].x, Is, P, a.inv_std_photo); sumP += r * r; }
                }
            }
            if (METHOD != R360_PHOTO_CONSISTENCY) {
                if (reach && depth_ok) {
                    if (OCC == 1 && cnt_on) ++n_depth;                    // RPI.h:3338
                    if (sum_on) { const float r = r360_wres_depth(t[0].y, dist, P); sumD += r * r; }
                }
            }
            // ---- rows of calcHessGrad_sphereOccN: Occ1 every in-bounds pixel (its z-buffer is indexed by
            //      the SOURCE pixel, RPI.h:3473-3475, and never rejects); Occ2 the last candidate of a texel
            okH[q] = OCC == 1 ? o.inb[q] : last;
            n_vis += okH[q] ? 1 : 0;                                      // numVisiblePixels (Occ2: distinct texels)
        }
        r360_rows_pair<METHOD, 1>(o.g, lv.res_inv, o.sp.Is, o.ta, o.tb, okH[0], okH[1], P, a.inv_std_photo, A);
    }

    // ---- block reduction: 27 normal-equation sums + PhotoResidual + DepthResidual, 3 counters
    float acc[R360_ACC_DOUBLES + 1];
    r360_acc_unpack(A, acc);
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
        for (int k = 0; k < R360_OCC_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
        r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < R360_OCC_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
        atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
    }
}

// =========================================================================== launch wrappers
// The pipeline slots of k_occ_eval need more than the 48 KB default of dynamic shared memory (per device: called by r360_create).
template <int METHOD, int OCC>
static cudaError_t r360_occ_attr() {
    return cudaFuncSetAttribute(k_occ_eval<METHOD, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, R360_OCC_DYN_SMEM);
}
cudaError_t r360_occ_init() {
    cudaError_t e = r360_occ_attr<R360_PHOTO_CONSISTENCY, 1>();
    if (e == cudaSuccess) e = r360_occ_attr<R360_DEPTH_CONSISTENCY, 1>();
    if (e == cudaSuccess) e = r360_occ_attr<R360_PHOTO_DEPTH, 1>();
    if (e == cudaSuccess) e = r360_occ_attr<R360_PHOTO_CONSISTENCY, 2>();
    if (e == cudaSuccess) e = r360_occ_attr<R360_DEPTH_CONSISTENCY, 2>();
    if (e == cudaSuccess) e = r360_occ_attr<R360_PHOTO_DEPTH, 2>();
    return e;
}
static dim3 occ_grid(const R360PassArgs& a, int n_pairs, int sm_count) {
    long long blocks = ((long long)(a.lv.n + 1) / 2 + R360_OCC_THREADS - 1) / R360_OCC_THREADS;
    const long long cap = std::max(1LL, 8LL * sm_count / std::max(n_pairs, 1));     // ~8 CTAs per SM over all pairs
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)n_pairs);
}

// One evaluation (error + Hessian / gradient of the occlusion variant) of the active pairs at their
// pose_eval: list heads reset, scatter, evaluate.  scratch: head | next | dinv, each n_pairs * lv.n.
void r360_launch_occ_pass(cudaStream_t st, const R360PassArgs& a, int n_pairs, int* head, int* next, float* dinv,
                          int sm_count) {
    const dim3 grid = occ_grid(a, n_pairs, sm_count);
    k_occ_reset<<<grid, R360_OCC_THREADS, 0, st>>>(a, head);
    const int occ = a.params.occlusion;
    if (occ == 1) k_occ_scatter<1><<<grid, R360_OCC_THREADS, 0, st>>>(a, head, next, dinv);
    else k_occ_scatter<2><<<grid, R360_OCC_THREADS, 0, st>>>(a, head, next, dinv);
    static const bool pipe = [] { const char* e = getenv("R360_OCC_PIPE"); return !e || atoi(e) != 0; }();
#define R360_OCC_EVAL(M)                                                                                 \
    do {                                                                                                 \
        if (!pipe) {                                                                                     \
            if (occ == 1) k_occ_eval_simple<M, 1><<<grid, R360_OCC_THREADS, 0, st>>>(a, head, next, dinv); \
            else k_occ_eval_simple<M, 2><<<grid, R360_OCC_THREADS, 0, st>>>(a, head, next, dinv);        \
        } else if (occ == 1) {                                                                           \
            k_occ_eval<M, 1><<<grid, R360_OCC_THREADS, R360_OCC_DYN_SMEM, st>>>(a, head, next, dinv);    \
        } else {                                                                                         \
            k_occ_eval<M, 2><<<grid, R360_OCC_THREADS, R360_OCC_DYN_SMEM, st>>>(a, head, next, dinv);    \
        }                                                                                                \
    } while (0)
    switch (a.params.method) {
        case R360_PHOTO_CONSISTENCY: R360_OCC_EVAL(R360_PHOTO_CONSISTENCY); break;
        case R360_DEPTH_CONSISTENCY: R360_OCC_EVAL(R360_DEPTH_CONSISTENCY); break;
        default: R360_OCC_EVAL(R360_PHOTO_DEPTH); break;
    }
#undef R360_OCC_EVAL
}
