// r360_kernels.cu -- hand-written sm_100a kernels of the spherical dense registration path.
//
//   K1  k_pyr_head                    level 0 + level-0 target texels + level 1 in one tiled pass over the
//                                      raw input; k_level0 / k_down / k_texel for the remaining levels
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
//       k_stitch                       Frame360 ingest: 8 sensor images -> sphere RGB8 / depth u16 (Frame360.h:1099-1148)
//       r360_pinhole.cuh (included at the end): k_pin_eval / k_gn_step_pin, the pinhole alignFrames;
//       r360_occ.cu: k_occ_scatter / k_occ_eval, the occlusion variants of the spherical path
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
    // 4 pixels per thread: 12 B of RGB as 3 x u32, 8 B of depth as uint2.  The vector loads of frame f start at
    // f * n_px * {3, 2, 4} bytes: aligned for every frame only when n_px is a multiple of 4 (else: the scalar loop)
    const int n4 = (n_px & 3) ? 0 : n_px >> 2;
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
// One thread = one output column x of a strip of R360_DOWN_R output rows: it walks down the source
// rows with a rolling window of five horizontally filtered values in registers, so every source
// row costs three 16-byte loads ({depth, gray} of columns 2x-2 .. 2x+3; lanes are consecutive x,
// so a warp reads one contiguous 512-byte span per load) and every output needs two new rows.
#define R360_DOWN_R 8
struct R360HRow { float h, dl, dr; };    // filtered gray, depth of source columns 2x and 2x+1
__device__ __forceinline__ R360HRow r360_down_hrow(const float2* __restrict__ s, int sr, int rows, int cols, int x) {
    const float2* row = s + (size_t)r360_reflect101(sr, rows) * cols;
    float c0, c1, c2, c3, c4;
    R360HRow o;
    if (x > 0 && 2 * x + 2 < cols) {
        const float4 A = __ldg(reinterpret_cast<const float4*>(row + 2 * x - 2));
        const float4 B = __ldg(reinterpret_cast<const float4*>(row + 2 * x));
        const float4 C = __ldg(reinterpret_cast<const float4*>(row + 2 * x + 2));
        c0 = A.y; c1 = A.w; c2 = B.y; c3 = B.w; c4 = C.y;
        o.dl = B.x; o.dr = B.z;
    } else {                                                         // image border columns: REFLECT_101
        const float4 B = __ldg(reinterpret_cast<const float4*>(row + 2 * x));
        c0 = __ldg(&row[r360_reflect101(2 * x - 2, cols)]).y;
        c1 = __ldg(&row[r360_reflect101(2 * x - 1, cols)]).y;
        c2 = B.y; c3 = B.w;
        c4 = __ldg(&row[r360_reflect101(2 * x + 2, cols)]).y;
        o.dl = B.x; o.dr = B.z;
    }
    o.h = c2 * 6 + (c1 + c3) * 4 + c0 + c4;
    return o;
}
__global__ void __launch_bounds__(256)
k_down(float2* const* __restrict__ pyr, long long off_src, long long off_dst, int rows, int cols,
       float min_d, float max_d) {
    const int f = blockIdx.y;
    const float2* __restrict__ s = pyr[f] + off_src;
    float2* __restrict__ o = pyr[f] + off_dst;
    const int h = rows >> 1, w = cols >> 1;
    const int n_strips = (h + R360_DOWN_R - 1) / R360_DOWN_R;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < w * n_strips; t += gridDim.x * blockDim.x) {
        const int strip = t / w, x = t - strip * w;
        const int y0 = strip * R360_DOWN_R, y1 = min(h, y0 + R360_DOWN_R);
        float hm2 = r360_down_hrow(s, 2 * y0 - 2, rows, cols, x).h;
        float hm1 = r360_down_hrow(s, 2 * y0 - 1, rows, cols, x).h;
        R360HRow r0 = r360_down_hrow(s, 2 * y0, rows, cols, x);
        for (int y = y0; y < y1; ++y) {
            const R360HRow r1 = r360_down_hrow(s, 2 * y + 1, rows, cols, x);
            const R360HRow r2 = r360_down_hrow(s, 2 * y + 2, rows, cols, x);
            const float a = (hm2 + r2.h) + (r0.h + r0.h);
            const float b = ((hm1 + r1.h) + r0.h) * 4.0f;
            const float gray = (a + b) * (1.f / 256);
            float av = 0.f;
            unsigned cnt = 0;
            if (r0.dl > min_d && r0.dl < max_d) { av += r0.dl; ++cnt; }
            if (r0.dr > min_d && r0.dr < max_d) { av += r0.dr; ++cnt; }
            if (r1.dl > min_d && r1.dl < max_d) { av += r1.dl; ++cnt; }
            if (r1.dr > min_d && r1.dr < max_d) { av += r1.dr; ++cnt; }
            const float depth = cnt > 0 ? av / cnt : 0.f;
            o[(size_t)y * w + x] = make_float2(depth, gray);
            hm2 = r0.h; hm1 = r1.h; r0 = r2;
        }
    }
}

// calcGradientXY (RPI.h:365-398): harmonic mean of the one-sided differences on strictly
// monotone triples, 0 elsewhere and on the border:  g = 2 / (1/(nxt - v) + 1/(v - prv)).
// Evaluated for two pixels at once with the packed IEEE reciprocal sequences (the differences of a
// strictly monotone triple of finite floats are normal numbers of equal sign, so the sequences are
// exact; 2 / s == 2 * RN(1 / s) because scaling by 2 commutes with rounding).  The four gradients of
// a pixel pair share ONE check for a non-finite result (Inf / NaN neighbours, CV_32F depth only, or
// operands outside the normal range), which recomputes them with the scalar IEEE operators.
__device__ __forceinline__ float r360_hgrad(float v, float nxt, float prv) {
    if ((v > nxt && v < prv) || (v < nxt && v > prv)) return 2.f / (1 / (nxt - v) + 1 / (v - prv));
    return 0.f;
}
__device__ __forceinline__ float2 r360_hgrad2(float2 v, float2 nxt, float2 prv) {
    const float2 a = f2add(nxt, f2neg(v)), b = f2add(v, f2neg(prv));
    const float2 r = f2rcp_rn(f2add(f2rcp_rn(a), f2rcp_rn(b)));
    const float2 g = f2add(r, r);
    // strictly monotone  <=>  both differences non-zero with equal sign (float subtraction keeps signs
    // exactly)  <=>  (a 2^100)(b 2^100) > 0: the scaling is exact and keeps the product of two non-zero
    // floats (denormals included) away from underflow; an overflow keeps the sign, 0 * Inf and NaN compare false.
    const float K = 1.2676506002282294e30f;                                  // 2^100
    const float2 p = f2mul(f2mul(a, R360_F2(K)), f2mul(b, R360_F2(K)));
    return make_float2(p.x > 0.f ? g.x : 0.f, p.y > 0.f ? g.y : 0.f);
}
struct R360Grad4 { float2 ix, iy, dx, dy; };
// Rare path, out of line (registers in, registers out): one gradient pair again, scalar IEEE operators.
static __device__ __noinline__ float2 r360_hgrad2_scalar(float v0, float n0, float p0, float v1, float n1, float p1) {
    return make_float2(r360_hgrad(v0, n0, p0), r360_hgrad(v1, n1, p1));
}
// Texels of the pixel pair (r, c), (r, c + 1), c even: v = {d0, g0, d1, g1} of the pair, u / d the pairs
// above / below, wv / e the {depth, gray} west of pixel 0 / east of pixel 1 (any finite value on the image
// border: border gradients are forced to 0).  Writes {gray, depth, Ix, Iy, Dx, Dy} x 2 as three float4.
struct R360MaskGeom { int ws, n_sensors; unsigned magic; };     // ws = cols / n_sensors (0: no mask), magic = ceil(2^32 / ws) (0 when ws == 1)
__device__ __forceinline__ void r360_texel_pair(float4 v, float4 u, float4 d, float2 wv, float2 e, int r, int c, int rows,
                                                int cols, const R360MaskGeom& mg, float4 out[3]) {
    R360Grad4 G;
    G.ix = G.iy = G.dx = G.dy = make_float2(0.f, 0.f);
    if (r > 0 && r < rows - 1) {
        G.ix = r360_hgrad2(make_float2(v.y, v.w), make_float2(v.w, e.y), make_float2(wv.y, v.y));
        G.dx = r360_hgrad2(make_float2(v.x, v.z), make_float2(v.z, e.x), make_float2(wv.x, v.x));
        G.iy = r360_hgrad2(make_float2(v.y, v.w), make_float2(d.y, d.w), make_float2(u.y, u.w));
        G.dy = r360_hgrad2(make_float2(v.x, v.z), make_float2(d.x, d.z), make_float2(u.x, u.z));
        const float2 chk = f2add(f2add(G.ix, G.iy), f2add(G.dx, G.dy));                      // Inf / NaN if any term is
        if (!(fabsf(chk.x) < INFINITY) | !(fabsf(chk.y) < INFINITY)) {
            G.ix = r360_hgrad2_scalar(v.y, v.w, wv.y, v.w, e.y, v.y);
            G.dx = r360_hgrad2_scalar(v.x, v.z, wv.x, v.z, e.x, v.x);
            G.iy = r360_hgrad2_scalar(v.y, d.y, u.y, v.w, d.w, u.w);
            G.dy = r360_hgrad2_scalar(v.x, d.x, u.x, v.z, d.z, u.z);
        }
        if (c == 0) { G.ix.x = 0.f; G.iy.x = 0.f; G.dx.x = 0.f; G.dy.x = 0.f; }              // border columns
        if (c + 2 == cols) { G.ix.y = 0.f; G.iy.y = 0.f; G.dx.y = 0.f; G.dy.y = 0.f; }
    }
    if (mg.ws > 0) {
        // sensor-joint columns k*ws-1 and k*ws, k = 1..n_sensors-1 (RPI.h:4537-4549); c / ws by multiplication
        // (exact for c * ws < 2^32)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int cc = c + p, k0 = mg.magic ? (int)__umulhi((unsigned)cc, mg.magic) : cc, rem = cc - k0 * mg.ws;
            const bool masked = (rem == 0 && k0 >= 1 && k0 <= mg.n_sensors - 1) ||
                                (rem == mg.ws - 1 && k0 + 1 <= mg.n_sensors - 1);
            if (masked) {
                if (p == 0) { G.ix.x = 0.f; G.iy.x = 0.f; G.dx.x = 0.f; G.dy.x = 0.f; }
                else { G.ix.y = 0.f; G.iy.y = 0.f; G.dx.y = 0.f; G.dy.y = 0.f; }
            }
        }
    }
    out[0] = make_float4(v.y, v.x, G.ix.x, G.iy.x);
    out[1] = make_float4(G.dx.x, G.dy.x, v.w, v.z);
    out[2] = make_float4(G.ix.y, G.iy.y, G.dx.y, G.dy.y);
}
// ---- the same gradients for a thread that walks DOWN a column of pixel pairs (k_pyr_head, r360_texel_pair_colz below): of the 12 reciprocals per
// pixel of calcGradientXY, three are shared.  Lanes are packed {depth, gray} of ONE pixel -- the order the plane is stored
// in, so differences are formed on the loaded register pairs as they are.
//   x: the three horizontal differences of a pair (v0 - w, v1 - v0, e - v1) serve both pixels: pixel 0's forward
//      difference IS pixel 1's backward difference (same operands, same bits) -> 3 reciprocal pairs instead of 4;
//   y: the backward difference of row r is the forward difference of row r - 1 -> its reciprocal (and its scaled copy for
//      the monotonicity test) is carried in registers from row to row -> 2 reciprocal pairs per row instead of 4.
// Every value is produced by the same IEEE operation on the same operands as r360_hgrad2, so the planes keep their bits.
struct R360ColState { float2 ra0, ra1, ka0, ka1; };     // 1 / (v[r] - v[r-1]) and (v[r] - v[r-1]) 2^100 of pixel 0 / 1, {depth, gray}
__device__ __forceinline__ float2 r360_lo(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 r360_hi(float4 v) { return make_float2(v.z, v.w); }
static inline R360MaskGeom r360_mask_geom(int cols, int n_sensors) {
    R360MaskGeom mg;
    mg.n_sensors = n_sensors;
    mg.ws = n_sensors > 1 ? cols / n_sensors : 0;
    mg.magic = mg.ws > 1 ? (unsigned)(((1ULL << 32) + mg.ws - 1) / mg.ws) : 0u;     // 0: ws == 1, c / ws = c
    return mg;
}

// Target texels of one level: {gray, depth, Ix, Iy, Dx, Dy} with the sensor-joint columns of the
// four gradient planes zeroed (RPI.h:4537-4549).  Two horizontally adjacent pixels per thread
// (cols is even): 5 loads and three 16-byte stores per pixel pair.
__global__ void __launch_bounds__(256)
k_texel(float2* const* __restrict__ pyr, float* const* __restrict__ trg, long long off, int rows,
        int cols, R360MaskGeom mg) {
    const int f = blockIdx.y;
    const float2* __restrict__ s = pyr[f] + off;
    float4* __restrict__ o = reinterpret_cast<float4*>(trg[f] + off * R360_TEXEL_FLOATS);
    const int n2 = (rows * cols) >> 1, half = cols >> 1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += gridDim.x * blockDim.x) {
        const int r = q / half, c = 2 * (q - r * half), i = r * cols + c;
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + i));              // {d0, g0, d1, g1}
        float4 u = v, d = v;
        float2 wv = make_float2(0.f, 0.f), e = wv;
        if (r > 0 && r < rows - 1) {
            u = __ldg(reinterpret_cast<const float4*>(s + i - cols));
            d = __ldg(reinterpret_cast<const float4*>(s + i + cols));
            if (c > 0) wv = __ldg(&s[i - 1]);
            if (c + 2 < cols) e = __ldg(&s[i + 2]);
        }
        float4 t[3];
        r360_texel_pair(v, u, d, wv, e, r, c, rows, cols, mg, t);
        o[3 * (size_t)q + 0] = t[0];
        o[3 * (size_t)q + 1] = t[1];
        o[3 * (size_t)q + 2] = t[2];
    }
}

// Fused head of the pyramid build: level 0 (RGB8 -> gray, depth -> metres), the level-0 target
// texels (gradients + joint mask) and level 1 (pyrDown + valid-mean) of a frame in ONE pass over the
// raw input, tile by tile through shared memory.  The separate kernels move 5 + 8 (k_level0) + 8 + 2
// (k_down) + 8 + 24 (k_texel) = 55 B per level-0 pixel of a target frame; this one moves 5 + 24 + 2
// (+ 8 only when the frame is also a source) -- the level-0 {depth, gray} plane of a target-only frame
// never exists in HBM.  Per-pixel arithmetic is that of k_level0 / k_down / k_texel, operation for
// operation (the planes stay bit-identical to the oracle's).
//   tile: R360_F0_TW x R360_F0_TH level-0 pixels + a 2-pixel REFLECT_101 halo (5-tap filter; the
//   gradients need 1), staged as {depth, gray}; x is padded to groups of 4 pixels so that the raw
//   loads are the 12-byte RGB / 8-byte depth vectors of k_level0.
#define R360_F0_TW 64
#ifndef R360_F0_TH
#define R360_F0_TH 64
#endif
#ifndef R360_F0_MINB
#define R360_F0_MINB 4                         // resident CTAs per SM the register allocation aims at
#endif
#define R360_F0_SW (R360_F0_TW + 8)            // smem columns: global x = tx0 - 4 + sx
#define R360_F0_SH (R360_F0_TH + 4)            // smem rows:    global y = ty0 - 2 + sy
#define R360_F0_ROWB (R360_F0_SW * 8)          // bytes per smem row
__device__ __forceinline__ int r360_reflect101_clamped(int i, int n) {
    i = r360_reflect101(i, n);
    return min(max(i, 0), n - 1);               // partial tiles reach far outside; those values are never used
}
__device__ __forceinline__ float r360_gray_u8(unsigned r, unsigned g, unsigned b) {
    const int v = (int)(r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15;
    return (float)v * (float)(1. / 255);
}
__device__ __forceinline__ float2 r360_lds64(unsigned smem) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem) : "memory");
    return v;
}
// g if p > 0 and the pixel is not masked, else 0: ONE compare that takes the mask as its predicate input and one select
// (nvcc makes two selects of `(p > 0) & !z ? g : 0`).  NaN compares false, as `p > 0.f` does.
__device__ __forceinline__ float r360_sel_pos(float p, float g, bool z) {
    float o;
    asm("{\n\t.reg .pred q, t;\n\tsetp.ne.s32 q, %3, 0;\n\tsetp.gt.and.f32 t, %1, 0f00000000, !q;\n\tselp.f32 %0, %2, 0f00000000, t;\n\t}"
        : "=f"(o) : "f"(p), "f"(g), "r"((int)z));
    return o;
}
// r360_hsel with the border / sensor-joint mask folded into the select.
__device__ __forceinline__ float2 r360_hsel_z(float2 rsum, float2 ka, float2 kb, bool z) {
    const float2 r = f2rcp_rn(rsum);
    const float2 g = f2add(r, r);
    const float2 p = f2mul(ka, kb);
    return make_float2(r360_sel_pos(p.x, g.x, z), r360_sel_pos(p.y, g.y, z));
}
// r360_texel_pair_col with the masks of the pair's two pixels (image border, sensor joints) decided by the caller --
// the column part is the same for every row of a strip.  SCALED = false: the differences themselves stand in for their
// 2^100 multiples in the monotonicity test -- exact when no product of two non-zero differences can underflow, which
// holds for 8-bit gray (steps of 1/255) and 16-bit millimetre depth (steps of 0.001), not for CV_32F depth.
template <bool SCALED>
__device__ __forceinline__ void r360_col_state_init_t(R360ColState& st, float4 u, float4 v) {
    const float K = 1.2676506002282294e30f;                                  // 2^100
    const float2 a0 = f2add(r360_lo(v), f2neg(r360_lo(u))), a1 = f2add(r360_hi(v), f2neg(r360_hi(u)));
    st.ra0 = f2rcp_rn(a0); st.ra1 = f2rcp_rn(a1);
    st.ka0 = SCALED ? f2mul(a0, R360_F2(K)) : a0; st.ka1 = SCALED ? f2mul(a1, R360_F2(K)) : a1;
}
template <bool SCALED>
__device__ __forceinline__ void r360_texel_pair_colz(R360ColState& st, float4 v, float4 u, float4 d, float2 wv, float2 e, bool z0,
                                                     bool z1, float4 out[3]) {
    const float K = 1.2676506002282294e30f;
    const float2 ay0 = f2add(r360_lo(d), f2neg(r360_lo(v))), ay1 = f2add(r360_hi(d), f2neg(r360_hi(v)));
    const float2 ray0 = f2rcp_rn(ay0), ray1 = f2rcp_rn(ay1);
    const float2 kay0 = SCALED ? f2mul(ay0, R360_F2(K)) : ay0, kay1 = SCALED ? f2mul(ay1, R360_F2(K)) : ay1;
    const float2 gy0 = r360_hsel_z(f2add(ray0, st.ra0), kay0, st.ka0, z0);   // {Dy, Iy} of pixel 0
    const float2 gy1 = r360_hsel_z(f2add(ray1, st.ra1), kay1, st.ka1, z1);
    st.ra0 = ray0; st.ra1 = ray1; st.ka0 = kay0; st.ka1 = kay1;
    const float2 d0 = f2add(r360_lo(v), f2neg(wv)), d1 = f2add(r360_hi(v), f2neg(r360_lo(v))), d2 = f2add(e, f2neg(r360_hi(v)));
    const float2 r0 = f2rcp_rn(d0), r1 = f2rcp_rn(d1), r2 = f2rcp_rn(d2);
    const float2 k0 = SCALED ? f2mul(d0, R360_F2(K)) : d0, k1 = SCALED ? f2mul(d1, R360_F2(K)) : d1, k2 = SCALED ? f2mul(d2, R360_F2(K)) : d2;
    const float2 gx0 = r360_hsel_z(f2add(r1, r0), k1, k0, z0);               // {Dx, Ix} of pixel 0
    const float2 gx1 = r360_hsel_z(f2add(r2, r1), k2, k1, z1);
    // the selects write the texel registers directly; the finite check adds them up in the pairing they are stored in
    // (any order will do: it only has to come out Inf / NaN when an unmasked term is)
    float2 i0 = make_float2(gx0.y, gy0.y), q0 = make_float2(gx0.x, gy0.x);   // {Ix, Iy}, {Dx, Dy} of pixel 0
    float2 i1 = make_float2(gx1.y, gy1.y), q1 = make_float2(gx1.x, gy1.x);
    const float2 chk = f2add(f2add(i0, q0), f2add(i1, q1));
    if (!(fabsf(chk.x) < INFINITY) | !(fabsf(chk.y) < INFINITY)) {           // rare: the scalar IEEE operators, out of line
        const float2 ix = r360_hgrad2_scalar(v.y, v.w, wv.y, v.w, e.y, v.y), dx = r360_hgrad2_scalar(v.x, v.z, wv.x, v.z, e.x, v.x);
        const float2 iy = r360_hgrad2_scalar(v.y, d.y, u.y, v.w, d.w, u.w), dy = r360_hgrad2_scalar(v.x, d.x, u.x, v.z, d.z, u.z);
        i0 = z0 ? make_float2(0.f, 0.f) : make_float2(ix.x, iy.x); q0 = z0 ? make_float2(0.f, 0.f) : make_float2(dx.x, dy.x);
        i1 = z1 ? make_float2(0.f, 0.f) : make_float2(ix.y, iy.y); q1 = z1 ? make_float2(0.f, 0.f) : make_float2(dx.y, dy.y);
    }
    out[0] = make_float4(v.y, v.x, i0.x, i0.y);
    out[1] = make_float4(q0.x, q0.y, v.w, v.z);
    out[2] = make_float4(i1.x, i1.y, q1.x, q1.y);
}
// Sensor-joint columns k*ws-1 and k*ws, k = 1..n_sensors-1 (RPI.h:4537-4549); c / ws by multiplication (exact for c * ws < 2^32).
__device__ __forceinline__ bool r360_joint_column(int cc, const R360MaskGeom& mg) {
    if (mg.ws <= 0) return false;
    const int k0 = mg.magic ? (int)__umulhi((unsigned)cc, mg.magic) : cc, rem = cc - k0 * mg.ws;
    return (rem == 0 && k0 >= 1 && k0 <= mg.n_sensors - 1) || (rem == mg.ws - 1 && k0 + 1 <= mg.n_sensors - 1);
}
// Horizontal 5-tap of one staged row at level-1 column ox (global columns 2x-2 .. 2x+2 = smem columns 2ox+2 .. 2ox+6) and
// the two depth parents of that column: k_down's r360_down_hrow, from shared memory.
__device__ __forceinline__ R360HRow r360_head_hrow(unsigned row_addr) {     // address of smem column 2ox+2 of the row
    const float4 A = r360_lds128(row_addr), B = r360_lds128(row_addr + 16);
    const float2 C = r360_lds64(row_addr + 32);
    R360HRow o;
    o.h = B.y * 6 + (A.w + B.w) * 4 + A.y + C.y;
    o.dl = B.x; o.dr = B.z;
    return o;
}
// IN: 0 = RGB8 + 16-bit depth, 1 = RGB8 + CV_32F depth, 2 = the {depth, gray} plane of a level >= 1 (plane_in[f] + off_in):
// the same tile pass then writes that level's target texels and the next level (k_texel + k_down in one read of the plane).
// off_l1 / off_tex: element offsets added to l1_dst[f] / texel_dst[f]; a null table or entry = that output is not wanted.
#define R360_IN_PLANE 2
template <int IN>
__global__ void __launch_bounds__(256, R360_F0_MINB)
k_pyr_head(const uint8_t* __restrict__ rgb, const uint16_t* __restrict__ depth_mm, const float* __restrict__ depth_m,
           const float2* const* __restrict__ plane_in, long long off_in,
           float2* const* __restrict__ l0_dst, float2* const* __restrict__ l1_dst, long long off_l1,
           float* const* __restrict__ texel_dst, long long off_tex,
           int rows, int cols, int tiles_x, float min_d, float max_d, R360MaskGeom mg) {
    constexpr bool F32DEPTH = IN == 1;
    constexpr bool SCALED = IN != 0;                                             // see r360_texel_pair_colz
    __shared__ __align__(16) float2 s_dg[R360_F0_SH][R360_F0_SW];
    const int f = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * R360_F0_TH, tx0 = (blockIdx.x % tiles_x) * R360_F0_TW;
    const size_t n_px = (size_t)rows * cols;
    const uint8_t* __restrict__ c8 = IN == R360_IN_PLANE ? nullptr : rgb + (size_t)f * n_px * 3;
    const float2* __restrict__ pin = IN == R360_IN_PLANE ? plane_in[f] + off_in : nullptr;
    const float ds = (float)0.001;
    const unsigned sbase = r360_smem_addr(&s_dg[0][0]);

    // ---- phase 1: raw input -> {depth, gray} of the tile + halo.  A thread owns ONE group of 4 columns and walks down the
    //      rows (row stride 14: 18 groups x 14 rows = 252 threads), so the column part of its addresses and the border
    //      decision are made once; pixel offsets are 32-bit.  All raw loads of the thread are issued first (registers),
    //      then converted: the HBM latency is paid once per tile, not once per group.
    constexpr int GROUPS = R360_F0_SW / 4, RPP = 256 / GROUPS, NIT = (R360_F0_SH + RPP - 1) / RPP;
    {
        const int tr = threadIdx.x / GROUPS, gq = threadIdx.x - tr * GROUPS;
        const int gx0 = tx0 - 4 + 4 * gq;
        if (tr < RPP) {
            const unsigned sdst = sbase + (unsigned)(tr * R360_F0_SW + 4 * gq) * 8u;
            if (IN == R360_IN_PLANE && gx0 >= 0 && gx0 + 3 < cols) {               // four staged pixels = two 16-byte loads
                float4 a[NIT], b[NIT];
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    const int sy = tr + RPP * k;
                    if (sy < R360_F0_SH) {
                        const unsigned off = (unsigned)r360_reflect101_clamped(ty0 - 2 + sy, rows) * (unsigned)cols + (unsigned)gx0;   // even
                        const float4* q = reinterpret_cast<const float4*>(pin + off);
                        a[k] = __ldg(q); b[k] = __ldg(q + 1);
                    }
                }
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    if (tr + RPP * k < R360_F0_SH) {
                        r360_sts128(sdst + (unsigned)(RPP * k) * R360_F0_ROWB, a[k]);
                        r360_sts128(sdst + (unsigned)(RPP * k) * R360_F0_ROWB + 16, b[k]);
                    }
                }
            } else if (IN != R360_IN_PLANE && gx0 >= 0 && gx0 + 3 < cols) {
                uint32_t w0[NIT], w1[NIT], w2[NIT];
                uint32_t dw[NIT][F32DEPTH ? 4 : 2];
                const uint32_t* __restrict__ c32 = reinterpret_cast<const uint32_t*>(c8);
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    const int sy = tr + RPP * k;
                    if (sy < R360_F0_SH) {
                        const unsigned off = (unsigned)r360_reflect101_clamped(ty0 - 2 + sy, rows) * (unsigned)cols + (unsigned)gx0;   // multiple of 4
                        const uint32_t* c4 = c32 + 3u * (off >> 2);
                        w0[k] = __ldg(c4); w1[k] = __ldg(c4 + 1); w2[k] = __ldg(c4 + 2);
                        if (F32DEPTH) {
                            const uint4 t = __ldg(reinterpret_cast<const uint4*>(depth_m + (size_t)f * n_px) + (off >> 2));
                            dw[k][0] = t.x; dw[k][1] = t.y; dw[k][F32DEPTH ? 2 : 0] = t.z; dw[k][F32DEPTH ? 3 : 1] = t.w;
                        } else {
                            const uint2 t = __ldg(reinterpret_cast<const uint2*>(depth_mm + (size_t)f * n_px) + (off >> 2));
                            dw[k][0] = t.x; dw[k][1] = t.y;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    if (tr + RPP * k < R360_F0_SH) {
                        float d[4], g[4];
                        g[0] = r360_gray_u8(w0[k] & 0xffu, (w0[k] >> 8) & 0xffu, (w0[k] >> 16) & 0xffu);
                        g[1] = r360_gray_u8(w0[k] >> 24, w1[k] & 0xffu, (w1[k] >> 8) & 0xffu);
                        g[2] = r360_gray_u8((w1[k] >> 16) & 0xffu, w1[k] >> 24, w2[k] & 0xffu);
                        g[3] = r360_gray_u8((w2[k] >> 8) & 0xffu, (w2[k] >> 16) & 0xffu, w2[k] >> 24);
                        if (F32DEPTH) {
                            d[0] = __uint_as_float(dw[k][0]); d[1] = __uint_as_float(dw[k][1]);
                            d[2] = __uint_as_float(dw[k][F32DEPTH ? 2 : 0]); d[3] = __uint_as_float(dw[k][F32DEPTH ? 3 : 1]);
                        } else {
                            d[0] = (float)(dw[k][0] & 0xffffu) * ds; d[1] = (float)(dw[k][0] >> 16) * ds;
                            d[2] = (float)(dw[k][1] & 0xffffu) * ds; d[3] = (float)(dw[k][1] >> 16) * ds;
                        }
                        r360_sts128(sdst + (unsigned)(RPP * k) * R360_F0_ROWB, make_float4(d[0], g[0], d[1], g[1]));
                        r360_sts128(sdst + (unsigned)(RPP * k) * R360_F0_ROWB + 16, make_float4(d[2], g[2], d[3], g[3]));
                    }
                }
            } else {                                                           // image border columns: REFLECT_101, pixel by pixel
                for (int sy = tr; sy < R360_F0_SH; sy += RPP) {
                    const int gy = r360_reflect101_clamped(ty0 - 2 + sy, rows);
                    float d[4], g[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int gx = r360_reflect101_clamped(gx0 + j, cols);
                        const size_t i = (size_t)gy * cols + gx;
                        if (IN == R360_IN_PLANE) {
                            const float2 t = __ldg(pin + i);
                            d[j] = t.x; g[j] = t.y;
                        } else {
                            g[j] = r360_gray_u8(c8[3 * i], c8[3 * i + 1], c8[3 * i + 2]);
                            d[j] = F32DEPTH ? depth_m[(size_t)f * n_px + i] : (float)depth_mm[(size_t)f * n_px + i] * ds;
                        }
                    }
                    const unsigned a = sbase + (unsigned)(sy * R360_F0_SW + 4 * gq) * 8u;
                    r360_sts128(a, make_float4(d[0], g[0], d[1], g[1]));
                    r360_sts128(a + 16, make_float4(d[2], g[2], d[3], g[3]));
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: level-0 outputs.  A thread walks down a strip of rows of ONE pixel-pair column (lane = column pair, so a
    //      warp still writes one contiguous row segment per step) and carries the vertical differences' reciprocals from
    //      row to row (r360_texel_pair_colz).  Everything that depends only on the column -- the joint / border masks, the
    //      shared-memory address -- is set up once; the strip is unrolled, so the row offsets are immediates.
    {
        float4* __restrict__ l0 = IN == R360_IN_PLANE ? nullptr : reinterpret_cast<float4*>(l0_dst[f]);
        float* const tex_f = texel_dst ? texel_dst[f] : nullptr;
        float4* __restrict__ tex = tex_f ? reinterpret_cast<float4*>(tex_f + off_tex) : nullptr;
        constexpr int STRIP = R360_F0_TH / (256 / (R360_F0_TW / 2));             // rows per thread
        const int lp = threadIdx.x % (R360_F0_TW / 2), ly0 = (threadIdx.x / (R360_F0_TW / 2)) * STRIP;
        const int c = tx0 + 2 * lp, r0 = ty0 + ly0;
        if (c < cols && r0 < rows) {                                             // cols is even: the pair is inside or outside
            const unsigned sa = sbase + (unsigned)((ly0 + 1) * R360_F0_SW + 2 * lp + 4) * 8u;   // the row above the strip
            const int n_row = min(STRIP, rows - r0);
            const unsigned half = (unsigned)cols >> 1;
            unsigned po = ((unsigned)r0 * (unsigned)cols + (unsigned)c) >> 1;    // pixel-pair index
            float4 v = r360_lds128(sa + R360_F0_ROWB);
            if (tex) {
                float4 u = r360_lds128(sa);
                R360ColState st;
                r360_col_state_init_t<SCALED>(st, u, v);
                const bool zc0 = (c == 0) | r360_joint_column(c, mg), zc1 = (c + 2 == cols) | r360_joint_column(c + 1, mg);
                auto strip = [&](auto with_l0) {                                 // a frame with both roles also gets its plane
#pragma unroll
                    for (int k = 0; k < STRIP; ++k) {
                        if (k >= n_row) break;
                        const float4 d = r360_lds128(sa + (k + 2) * R360_F0_ROWB);
                        const float2 wv = r360_lds64(sa + (k + 1) * R360_F0_ROWB - 8), e = r360_lds64(sa + (k + 1) * R360_F0_ROWB + 16);
                        if (decltype(with_l0)::value) l0[po] = v;
                        const bool rb = (unsigned)(r0 + k - 1) >= (unsigned)(rows - 2);      // first / last image row
                        float4 t[3];
                        r360_texel_pair_colz<SCALED>(st, v, u, d, wv, e, zc0 | rb, zc1 | rb, t);
                        float4* o = tex + 3 * (size_t)po;
                        o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
                        u = v; v = d; po += half;
                    }
                };
                if (l0) strip(std::true_type{}); else strip(std::false_type{});
            } else if (l0) {
#pragma unroll
                for (int k = 0; k < STRIP; ++k) {
                    if (k >= n_row) break;
                    l0[po] = v;
                    v = r360_lds128(sa + (k + 2) * R360_F0_ROWB);
                    po += half;
                }
            }
        }
    }

    // ---- phase 3: level 1.  A thread owns one output column of a strip of 4 output rows and rolls the horizontally
    //      filtered rows through registers (k_down's scheme, from shared memory): 2 new rows per output row.
    {
        constexpr int OSTRIP = (R360_F0_TH / 2) / (256 / (R360_F0_TW / 2));      // output rows per thread
        const int ox = threadIdx.x % (R360_F0_TW / 2), oy0 = (threadIdx.x / (R360_F0_TW / 2)) * OSTRIP;
        const int h1 = rows >> 1, wd1 = cols >> 1;
        const int x = (tx0 >> 1) + ox, y0 = (ty0 >> 1) + oy0;
        if (l1_dst && x < wd1 && y0 < h1) {
            float2* __restrict__ l1 = l1_dst[f] + off_l1 + (size_t)y0 * wd1 + x;
            const int n_out = min(OSTRIP, h1 - y0);
            const unsigned ra = sbase + (unsigned)(2 * oy0 * R360_F0_SW + 2 * ox + 2) * 8u;    // smem row 2 oy0, column 2 ox + 2
            float hm2 = r360_head_hrow(ra).h, hm1 = r360_head_hrow(ra + R360_F0_ROWB).h;
            R360HRow q0 = r360_head_hrow(ra + 2 * R360_F0_ROWB);
#pragma unroll
            for (int j = 0; j < OSTRIP; ++j) {
                if (j >= n_out) break;
                const R360HRow q1 = r360_head_hrow(ra + (2 * j + 3) * R360_F0_ROWB);
                const R360HRow q2 = r360_head_hrow(ra + (2 * j + 4) * R360_F0_ROWB);
                const float a = (hm2 + q2.h) + (q0.h + q0.h);
                const float b = ((hm1 + q1.h) + q0.h) * 4.0f;
                const float gray = (a + b) * (1.f / 256);
                float av = 0.f;
                unsigned cnt = 0;
                if (q0.dl > min_d && q0.dl < max_d) { av += q0.dl; ++cnt; }
                if (q0.dr > min_d && q0.dr < max_d) { av += q0.dr; ++cnt; }
                if (q1.dl > min_d && q1.dl < max_d) { av += q1.dl; ++cnt; }
                if (q1.dr > min_d && q1.dr < max_d) { av += q1.dr; ++cnt; }
                const float depth = cnt > 0 ? av / cnt : 0.f;
                l1[(size_t)j * wd1] = make_float2(depth, gray);
                hm2 = q0.h; hm1 = q1.h; q0 = q2;
            }
        }
    }
}

// =========================================================================== K3: fused pixel pass
// Work decomposition: the (active pair, pixel block) space is cut into `items` of px_per_item
// pixels.  STATIC part (the first 1 - dyn_permille / 1000 of the items): every CTA of the persistent
// grid takes one CONTIGUOUS run of items, so it touches one or two pairs, keeps its 28 packed partial
// sums in registers across the whole run and flushes them (warp shuffles + shared memory + 28
// order-independent fixed-point atomic sums, r360_fx_add) once per pair it touched.  DYNAMIC part (the
// remaining items): CTAs that are done fetch single items from a global counter and flush after each.
// Equal pixel counts do not take equal time -- the row skips are content dependent and the SMs do not
// stream equally fast (ncu: SMs busy 88 % of a static launch on average, 80 % the fastest) -- so the
// dynamic tail lets the fast CTAs take what the slow ones would still be working on.  The sums stay
// reproducible bit for bit: a static run and a dynamic item each cover a FIXED set of pixels with a fixed
// pixel-to-thread mapping whoever executes them, and the cross-CTA accumulation is order-independent.
//
// Per thread and iteration one pixel pair, software-pipelined through shared memory in two stages:
//   stage A (pair k+1): LDG.128 of {depth, gray} x 2 (loaded two iterations ahead), back-projection,
//            packed pinned index sequence, then the six 8-byte texel gathers are issued as
//            cp.async (LDGSTS) straight into the thread's shared-memory slot and the warped
//            geometry is parked next to them -- no register is held across the HBM round trip;
//   stage B (pair k):   cp.async.wait_group, 6 x LDS.128, residuals / Jacobians / 56 FFMA2 of
//            normal-equation accumulation.
// Every thread reads only what it wrote itself, so the pipeline needs no block barrier.
#define R360_SLOT_BYTES 48                                   // texels of 2 pixels = geometry of 2 pixels = 48 B
#define R360_PASS_DYN_SMEM (R360_PASS_STAGES * 2 * R360_PASS_THREADS * R360_SLOT_BYTES)

// WITH_H = false: the error-only pass (errorPhotoICP_sphere alone) of the pairs whose candidate pose the
// Gauss-Newton model predicts will end the level: same stage A, same residuals, no Jacobians and no 56
// accumulator registers -- one more resident CTA per SM.
template <int METHOD, bool WITH_H>
#ifdef R360_PASS_MAXNREG
__global__ void __maxnreg__(R360_PASS_MAXNREG)               // explicit register cap (variant builds)
#else
__global__ void __launch_bounds__(R360_PASS_THREADS, WITH_H ? R360_PASS_CTAS : R360_ERR_CTAS)
#endif
k_pass(R360PassArgs a) {
    extern __shared__ float4 s_pipe[];                       // [stage][texel | geometry][thread][3]
    __shared__ float s_red[R360_PASS_THREADS / 32][R360_ACC_DOUBLES];
    __shared__ int s_cnt[R360_PASS_THREADS / 32][R360_ACC_INTS];
    __shared__ __align__(16) float s_T[16];
    __shared__ int s_next;
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const float inv_std_photo = a.inv_std_photo;
    const int ipp = a.items_per_pair, ppi = a.px_per_item;
    int item, item_end;                                              // item_end < 0: the dynamic phase
    {
        const int n_items = (*a.n_active) * ipp;
        const int n_static = n_items - (int)(((long long)n_items * a.dyn_permille) / 1000);
        const int per = n_static / (int)gridDim.x, rem = n_static - per * (int)gridDim.x;
        const int bid = (int)blockIdx.x;
        item = bid * per + min(bid, rem);
        item_end = item + per + (bid < rem ? 1 : 0);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int STRIDE = 2 * R360_PASS_THREADS;
    // this thread's slots: stage s at + s * STAGE_BYTES; texels at +0, geometry at + GEO_OFF
    constexpr unsigned STAGE_BYTES = 2 * R360_PASS_THREADS * R360_SLOT_BYTES, GEO_OFF = R360_PASS_THREADS * R360_SLOT_BYTES;
    const unsigned slot0 = r360_smem_addr(reinterpret_cast<char*>(s_pipe) + R360_SLOT_BYTES * threadIdx.x);
    const unsigned ts = r360_smem_addr(s_T);                         // held in a register: no generic pointer, no window arithmetic in the loop

    int ap = item / ipp;
    int sub = item - ap * ipp;
    for (;;) {
        __syncthreads();                                             // s_T / s_red / s_next of the previous segment
        if (item >= item_end) {                                      // static run done (uniform over the CTA): single items from the counter
            item_end = -1;
            const int n_items = (*a.n_active) * ipp;
            const int n_static = n_items - (int)(((long long)n_items * a.dyn_permille) / 1000);
            if (n_static == n_items) break;
            if (threadIdx.x == 0) s_next = n_static + atomicAdd(a.work_counter, 1);
            __syncthreads();
            item = s_next;
            if (item >= n_items) break;
            ap = item / ipp;
            sub = item - ap * ipp;
        }
        const int n_seg = item_end < 0 ? 1 : min(ipp - sub, item_end - item);   // items of pair `ap` in this segment
        const int pair = a.active_list[ap];
        const R360Pair* __restrict__ ps = a.pairs + pair;
        if (threadIdx.x < 16) s_T[threadIdx.x] = ps->pose_eval[threadIdx.x];
        __syncthreads();
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
        const float2* __restrict__ trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);

        R360Acc2 A;
        r360_acc_zero(A);
        int n_vis = 0, n_photo = 0, n_depth = 0;
        unsigned n_fb = 0;

        const int p_begin = sub * ppi;
        const int p_end = min(lv.n, (sub + n_seg) * ppi);
        const int n_it = (p_end - p_begin + STRIDE - 1) / STRIDE;   // uniform over the CTA
        int i = p_begin + 2 * (int)threadIdx.x;                      // pixel pair of the next stage A
        int r = (int)(((unsigned long long)i * lv.div_magic) >> 40);
        int c = i - r * lv.cols;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 s_cur = i < p_end ? __ldg(&src4[i >> 1]) : zero4;
        float4 s_nxt = i + STRIDE < p_end ? __ldg(&src4[(i + STRIDE) >> 1]) : zero4;
#ifdef R360_PREFETCH_TAB
        float2 tp_n;                                                 // trig-table entries of the next stage A,
        float4 tt_n;                                                 // loaded one stage ahead (variant build)
        r360_load_tabs(lv, r, c, tp_n, tt_n);
#endif

        // stage A: index of pixel pair (i, i+1), async gathers + geometry into slot `st`
        auto stage_a = [&](unsigned st) {                            // st: byte offset of the stage
            R360SrcPair sp;
#ifdef R360_PREFETCH_TAB
            r360_load_src_pair(lv, P, s_cur, tp_n, tt_n, i < p_end, i + 1 < p_end, sp);
#else
            r360_load_src_pair(lv, P, s_cur, r, c, i < p_end, i + 1 < p_end, sp);
#endif
            R360Geo2 g;
            int rr[2], cc[2];
#ifdef R360_T_IN_REGS
            r360_index_pair(T, ts, lv, sp, a.one, g, rr, cc, n_fb);
#else
            float T[16];                                             // pose: 3 x LDS.128 per iteration, no long-lived registers
            {
                const float4 c0 = r360_lds128(ts), c1 = r360_lds128(ts + 16), c2 = r360_lds128(ts + 32), c3 = r360_lds128(ts + 48);
                T[0] = c0.x; T[1] = c0.y; T[2] = c0.z; T[4] = c1.x; T[5] = c1.y; T[6] = c1.z;
                T[8] = c2.x; T[9] = c2.y; T[10] = c2.z; T[12] = c3.x; T[13] = c3.y; T[14] = c3.z;
            }
            r360_index_pair(T, ts, lv, sp, a.one, g, rr, cc, n_fb);
#endif
            // RPI.h:2683 / 2989 (no c' >= 0 test upstream; c' is never negative, the unsigned compare
            // only guards memory)
            const bool ok0 = sp.v0 & ((unsigned)rr[0] < (unsigned)lv.rows) & ((unsigned)cc[0] < (unsigned)lv.cols);
            const bool ok1 = sp.v1 & ((unsigned)rr[1] < (unsigned)lv.rows) & ((unsigned)cc[1] < (unsigned)lv.cols);
            const float2* tx0 = trg + 3u * (ok0 ? (unsigned)(rr[0] * lv.cols + cc[0]) : 0u);
            const float2* tx1 = trg + 3u * (ok1 ? (unsigned)(rr[1] * lv.cols + cc[1]) : 0u);
            const unsigned dst = slot0 + st;
            r360_cp_async8(dst + 0, tx0); r360_cp_async8(dst + 8, tx0 + 1); r360_cp_async8(dst + 16, tx0 + 2);
            r360_cp_async8(dst + 24, tx1); r360_cp_async8(dst + 32, tx1 + 1); r360_cp_async8(dst + 40, tx1 + 2);
            r360_cp_async_commit();
            r360_sts128(dst + GEO_OFF, make_float4(g.px.x, g.px.y, g.py.x, g.py.y));
            r360_sts128(dst + GEO_OFF + 16, make_float4(g.pz.x, g.pz.y, g.dinv.x, g.dinv.y));
            // |p| > 0: its sign carries the in-bounds flag of the pixel
            r360_sts128(dst + GEO_OFF + 32, make_float4(sp.Is.x, sp.Is.y, ok0 ? g.dist.x : -g.dist.x, ok1 ? g.dist.y : -g.dist.y));
            r360_count(n_vis, ok0); r360_count(n_vis, ok1);
            // next pixel pair of this thread
            i += STRIDE;
            r += lv.stride_r;
            c += lv.stride_c;
            if (c >= lv.cols) { c -= lv.cols; ++r; }
#ifdef R360_PREFETCH_TAB
            r360_load_tabs(lv, r, c, tp_n, tt_n);
#endif
        };

        // prologue: pixel pairs 0 .. STAGES-2 in flight before the first stage B
        constexpr int S = R360_PASS_STAGES;
        constexpr unsigned RING_END = S * STAGE_BYTES;
        stage_a(0u);
#pragma unroll
        for (int p = 1; p < S - 1; ++p) {
            if (p < n_it) {
                s_cur = s_nxt;
                s_nxt = i + STRIDE < p_end ? __ldg(&src4[(i + STRIDE) >> 1]) : zero4;
                stage_a(p * STAGE_BYTES);
            } else {
                r360_cp_async_commit();
            }
        }
        unsigned st_b = 0u;                                          // stage of pixel pair k
        unsigned st_a = (S - 1) * STAGE_BYTES;                       // stage the next stage A fills
#ifdef R360_PASS_UNROLL
        constexpr int kUnroll = R360_PASS_UNROLL;
#pragma unroll kUnroll
#endif
        for (int k = 0; k < n_it; ++k) {
#ifdef R360_PASS_SYNC_EVERY
            // keeps the CTA's warps in step (they share texel rows in L1 and stream one region): n_it is uniform over the CTA
            if ((k & (R360_PASS_SYNC_EVERY - 1)) == R360_PASS_SYNC_EVERY - 1) __syncthreads();
#endif
            if (k + S - 1 < n_it) {
                s_cur = s_nxt;
                s_nxt = i + STRIDE < p_end ? __ldg(&src4[(i + STRIDE) >> 1]) : zero4;
                stage_a(st_a);
            } else {
                r360_cp_async_commit();                              // keeps "all but the S-1 newest groups" == group k
            }
            st_a += STAGE_BYTES;
            if (st_a == RING_END) st_a = 0u;
            r360_cp_async_wait<S - 1>();
            // stage B: pixel pair k
            const unsigned rd = slot0 + st_b;
            st_b += STAGE_BYTES;
            if (st_b == RING_END) st_b = 0u;
            const float4 q0 = r360_lds128(rd), q1 = r360_lds128(rd + 16), q2 = r360_lds128(rd + 32);
            const float4 g0 = r360_lds128(rd + GEO_OFF), g1 = r360_lds128(rd + GEO_OFF + 16), g2 = r360_lds128(rd + GEO_OFF + 32);
            const float2 ta[3] = { make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y) };
            const float2 tb[3] = { make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w) };
            R360Geo2 g;
            g.px = make_float2(g0.x, g0.y); g.py = make_float2(g0.z, g0.w);
            g.pz = make_float2(g1.x, g1.y); g.dinv = make_float2(g1.z, g1.w);
            g.dist = make_float2(fabsf(g2.z), fabsf(g2.w));
            g.rho2 = f2fma(g.py, g.py, f2mul(g.pz, g.pz));
            const bool ok0 = g2.z > 0.f, ok1 = g2.w > 0.f;
            if (WITH_H)
                r360_rows_pair<METHOD>(g, lv.res_inv, make_float2(g2.x, g2.y), ta, tb, ok0, ok1, P, inv_std_photo, A,
                                       &n_photo, &n_depth);
            else
                r360_err_pair<METHOD>(g, make_float2(g2.x, g2.y), ta, tb, ok0, ok1, P, inv_std_photo, A.e2, n_photo, n_depth);
        }

        // ---- flush: warp shuffles, one shared-memory stage, 28 fixed-point atomic sums per CTA and pair (order-independent)
        float acc[R360_ACC_DOUBLES];
        r360_acc_unpack(A, acc);
#pragma unroll
        for (int k = WITH_H ? 0 : R360_ACC_DOUBLES - 1; k < R360_ACC_DOUBLES; ++k) {      // error-only: just sum r^2
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[wid][k] = v;
        }
        n_vis = __reduce_add_sync(0xffffffffu, n_vis);
        n_photo = __reduce_add_sync(0xffffffffu, n_photo);
        n_depth = __reduce_add_sync(0xffffffffu, n_depth);
        n_fb = __reduce_add_sync(0xffffffffu, n_fb);
        if (lane == 0) { s_cnt[wid][0] = n_vis; s_cnt[wid][1] = n_photo; s_cnt[wid][2] = n_depth; s_cnt[wid][3] = (int)n_fb; }
        __syncthreads();
        if (threadIdx.x < R360_ACC_DOUBLES && (WITH_H || threadIdx.x == R360_ACC_DOUBLES - 1)) {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < R360_PASS_THREADS / 32; ++k) sum += (double)s_red[k][threadIdx.x];
            r360_fx_add(a.acc + (size_t)pair * R360_ACC_STRIDE, threadIdx.x, sum);
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + R360_ACC_INTS) {
            int sum = 0;
#pragma unroll
            for (int k = 0; k < R360_PASS_THREADS / 32; ++k) sum += s_cnt[k][threadIdx.x - 32];
            atomicAdd(&a.cnt[(size_t)pair * R360_ACC_INTS + threadIdx.x - 32], sum);
        }
        item += n_seg;
        ++ap;
        sub = 0;
    }
}

// Parity hook: per source pixel the rounded target index and the validPixelsPhoto/Depth masks,
// through the same index function as k_pass.  One warp-uniform loop; 2 pixels per thread.
__global__ void __launch_bounds__(256)
k_warp_dump(R360PassArgs a, int pair, int method, int32_t* __restrict__ r_idx, int32_t* __restrict__ c_idx,
            uint8_t* __restrict__ vphoto, uint8_t* __restrict__ vdepth) {
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const R360Pair* ps = a.pairs + pair;
    float T[16];
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float4* src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    const float2* trg = reinterpret_cast<const float2*>(a.trg_base[pair] + lv.px_off * R360_TEXEL_FLOATS);
    const int stride = 2 * gridDim.x * blockDim.x;
    for (int base = 0; base < lv.n; base += stride) {
        const int i = base + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
        const bool in0 = i < lv.n, in1 = i + 1 < lv.n;
        const int r = in0 ? (int)(((unsigned long long)i * lv.div_magic) >> 40) : 0;
        const int c = in0 ? i - r * lv.cols : 0;
        const float4 s = in0 ? src4[i >> 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        R360SrcPair sp;
        r360_load_src_pair(lv, P, s, r, c, in0, in1, sp);
        R360Geo2 g;
        int rr[2], cc[2];
        unsigned n_fb = 0;
        r360_index_pair(T, ps->pose_eval, lv, sp, a.one, g, rr, cc, n_fb);
        const bool ok0 = sp.v0 && (unsigned)rr[0] < (unsigned)lv.rows && (unsigned)cc[0] < (unsigned)lv.cols;
        const bool ok1 = sp.v1 && (unsigned)rr[1] < (unsigned)lv.rows && (unsigned)cc[1] < (unsigned)lv.cols;
        const float2* tx0 = trg + 3u * (ok0 ? (unsigned)(rr[0] * lv.cols + cc[0]) : 0u);
        const float2* tx1 = trg + 3u * (ok1 ? (unsigned)(rr[1] * lv.cols + cc[1]) : 0u);
        const float2 ta[3] = { tx0[0], tx0[1], tx0[2] }, tb[3] = { tx1[0], tx1[1], tx1[2] };
        R360Acc2 A;
        r360_acc_zero(A);
        unsigned v;
        if (method == R360_PHOTO_CONSISTENCY)
            v = r360_rows_pair<R360_PHOTO_CONSISTENCY>(g, lv.res_inv, sp.Is, ta, tb, ok0, ok1, P, a.inv_std_photo, A);
        else if (method == R360_DEPTH_CONSISTENCY)
            v = r360_rows_pair<R360_DEPTH_CONSISTENCY>(g, lv.res_inv, sp.Is, ta, tb, ok0, ok1, P, a.inv_std_photo, A);
        else
            v = r360_rows_pair<R360_PHOTO_DEPTH>(g, lv.res_inv, sp.Is, ta, tb, ok0, ok1, P, a.inv_std_photo, A);
        if (in0) {
            if (r_idx) r_idx[i] = sp.v0 ? rr[0] : INT_MIN;
            if (c_idx) c_idx[i] = sp.v0 ? cc[0] : INT_MIN;
            if (vphoto) vphoto[i] = (uint8_t)(v & 1u);
            if (vdepth) vdepth[i] = (uint8_t)((v >> 2) & 1u);
        }
        if (in1) {
            if (r_idx) r_idx[i + 1] = sp.v1 ? rr[1] : INT_MIN;
            if (c_idx) c_idx[i + 1] = sp.v1 ? cc[1] : INT_MIN;
            if (vphoto) vphoto[i + 1] = (uint8_t)((v >> 1) & 1u);
            if (vdepth) vdepth[i + 1] = (uint8_t)((v >> 3) & 1u);
        }
    }
}

// Cross-check of the packed index path against the scalar pinned sequence over every valid source
// pixel of one pair: out[0] = valid pixels, out[1] = pixels the packed path sent to the scalar path,
// out[2] = pixels it kept whose (r', c') differ from the scalar result (must be 0).
__global__ void __launch_bounds__(256)
k_index_stats(R360PassArgs a, int pair, unsigned long long* __restrict__ out) {
    const R360Level lv = a.lv;
    const r360_params P = a.params;
    const R360Pair* ps = a.pairs + pair;
    float T[16];
    for (int k = 0; k < 16; ++k) T[k] = ps->pose_eval[k];
    const float4* src4 = reinterpret_cast<const float4*>(a.src_base[pair] + lv.px_off);
    unsigned long long n_valid = 0, n_fb = 0, n_bad = 0;
    const int stride = 2 * gridDim.x * blockDim.x;
    for (int base = 0; base < lv.n; base += stride) {
        const int i = base + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
        const bool in0 = i < lv.n, in1 = i + 1 < lv.n;
        const int r = in0 ? (int)(((unsigned long long)i * lv.div_magic) >> 40) : 0;
        const int c = in0 ? i - r * lv.cols : 0;
        const float4 s = in0 ? src4[i >> 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        R360SrcPair sp;
        r360_load_src_pair(lv, P, s, r, c, in0, in1, sp);
        R360Geo2 g;
        int rf[2], cf[2];
        bool bad[2];
        r360_index_pair_packed(T, sp.X0, sp.X1, sp.X2, lv.res_inv, lv.half_rows, a.one, g, rf, cf, bad);
        const unsigned need = (bad[0] ? 1u : 0u) | (bad[1] ? 2u : 0u);
        for (int q = 0; q < 2; ++q) {
            if (!(q ? sp.v1 : sp.v0)) continue;
            const float X[3] = { q ? sp.X0.y : sp.X0.x, q ? sp.X1.y : sp.X1.x, q ? sp.X2.y : sp.X2.x };
            int re, ce;
            float vre, vce;
            r360_index_exact_inl(T, X, lv.res_inv, lv.half_rows, re, ce, vre, vce);
            ++n_valid;
            if (need & (1u << q)) ++n_fb;
            else if (re != rf[q] || ce != cf[q]) ++n_bad;
        }
    }
    atomicAdd(&out[0], n_valid);
    atomicAdd(&out[1], n_fb);
    atomicAdd(&out[2], n_bad);
}

// =========================================================================== K4: Gauss-Newton state machine
__device__ void r360_zero_acc(R360Fx* acc, int* cnt, int pair) {
    for (int k = 0; k < R360_ACC_STRIDE; ++k) { acc[(size_t)pair * R360_ACC_STRIDE + k].lo = 0ull; acc[(size_t)pair * R360_ACC_STRIDE + k].hi = 0ll; }
    for (int k = 0; k < R360_ACC_INTS; ++k) cnt[(size_t)pair * R360_ACC_INTS + k] = 0;
}

// Compacts the active flags of pairs [0, n) into the two pass lists (ascending pair order), one warp:
// pairs whose next pass is the fused one, and pairs whose next pass is error-only.
__device__ void r360_compact_active(R360Pair* pairs, int n, int* active_list, int* n_active, int* list_err, int* n_err) {
    const int lane = threadIdx.x & 31;
    int base = 0, base_e = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        const bool act = p < n && *(volatile int*)&pairs[p].active;
        const bool full = act && *(volatile int*)&pairs[p].want_h;
        const bool err = act && !full;
        const unsigned m = __ballot_sync(0xffffffffu, full), me = __ballot_sync(0xffffffffu, err);
        if (full) active_list[base + __popc(m & ((1u << lane) - 1))] = p;
        if (err) list_err[base_e + __popc(me & ((1u << lane) - 1))] = p;
        base += __popc(m);
        base_e += __popc(me);
    }
    if (lane == 0) { *n_active = base; *n_err = base_e; }
}

// The last block of a state-machine kernel to finish rebuilds the compact active list (what a
// separate one-warp launch did before: 135 launches per 64-pair batch fewer).
__device__ void r360_compact_when_last(const R360GnArgs& g) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(g.ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (threadIdx.x < 32) r360_compact_active(g.pairs, g.n_pairs, g.active_list, g.n_active, g.active_list_err, g.n_active_err);
        if (threadIdx.x == 0) {
            *g.ticket = 0;
            g.work_counters[0] = 0;                                  // dynamic-item counters of the next k_pass launches
            g.work_counters[1] = 0;
        }
    }
}

// Start of a pyramid level (RPI.h:4589-4605): evaluate the current estimate first.
__global__ void k_level_begin(R360GnArgs g, int level) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.n_pairs; p += gridDim.x * blockDim.x) {
        R360Pair* ps = g.pairs + p;
        r360_zero_acc(g.acc, g.cnt, p);
        if (ps->status != R360_PAIR_OK) { ps->active = 0; continue; }
        for (int k = 0; k < 16; ++k) ps->pose_eval[k] = ps->pose_estim[k];
        for (int k = 0; k < 6; ++k) ps->upd[k] = 1.f;
        ps->lambda = g.lambda0 > 0.0 ? g.lambda0 : (g.params.projection == R360_PINHOLE ? 0.01 : 1.0);     // RPI.h:4589 / 4304
        ps->it = 0;
        ps->phase = 0;
        ps->ev = 0;
        ps->want_h = 1;
        ps->extra = 0;
        ps->active = 1;
    }
    r360_compact_when_last(g);
}

// One step of the per-pair state machine after a pixel pass (RPI.h:4599-4722).
//   phase 0: the pass evaluated pose_estim at the start of the level.
//   phase 1: the pass evaluated the candidate pose_tmp = exp(update) * pose_estim.
// The reference runs errorPhotoICP_sphere(pose_tmp) and, if accepted, calcHessGrad_sphere at the
// same pose in the next loop body; the fused pass already produced both.
// One pair per WARP (lane 0 works): the pivoting solves and the accept / reject branches of different pairs diverge,
// and 32 of them in one warp run one after the other -- the step is pure latency between two pixel passes.
__global__ void k_gn_step(R360GnArgs g, int level) {
    const r360_params P = g.params;
    if (*g.n_active == 0 && *g.n_active_err == 0) return;       // every pair has left the level: the remaining launches of its schedule are empty
    for (int p = threadIdx.x == 0 ? (int)blockIdx.x : g.n_pairs; p < g.n_pairs; p += gridDim.x) {
        R360Pair* ps = g.pairs + p;
        if (!ps->active) continue;
        double acc[R360_ACC_DOUBLES + 1];
        for (int k = 0; k < R360_ACC_DOUBLES + 1; ++k) acc[k] = r360_fx_get(g.acc + (size_t)p * R360_ACC_STRIDE, k);
        const int* cnt = g.cnt + (size_t)p * R360_ACC_INTS;
        const int n_vis = cnt[0];
        double e2, err;
        int n_valid;
        if (P.occlusion == 0) {
            e2 = acc[27];
            n_valid = cnt[1] + cnt[2];
            err = sqrt(e2 / (double)n_valid);               // RPI.h:2738
        } else {
            // errorPhotoICP_sphereOcc1: sqrt(Photo / nValidPhotoPts) + sqrt(Depth / nValidDepthPts)  RPI.h:3360-3367
            // errorPhotoICP_sphereOcc2: both over nValidDepthPts                                      RPI.h:3849-3856
            // (a missing term gives 0/0 = NaN upstream as well: the loop then never runs)
            const double n_p = (double)(P.occlusion == 1 ? cnt[1] : cnt[2]), n_d = (double)cnt[2];
            err = sqrt(acc[27] / n_p) + sqrt(acc[28] / n_d);
            e2 = acc[27] + acc[28];
            n_valid = P.occlusion == 1 ? cnt[1] + cnt[2] : cnt[2];
        }
        ps->passes[level] += 1;
        if (ps->phase == 3) {
            // The accepted candidate had been evaluated by an error-only pass (its step was predicted to end the
            // level); this pass is calcHessGrad_sphere at the same pose (RPI.h:4623), the loop goes on.
            for (int k = 0; k < 21; ++k) ps->Hc[k] = (float)acc[k];
            for (int k = 0; k < 6; ++k) ps->gc[k] = (float)acc[21 + k];
            ps->nvis_c = n_vis;
            if (g.trace && ps->ev >= 1 && ps->ev - 1 < P.max_iters + 2) {
                r360_iter_record* pr = g.trace + ((size_t)p * P.n_levels + level) * (P.max_iters + 2) + ps->ev - 1;
                for (int k = 0; k < 21; ++k) pr->hessian[k] = (float)acc[k];
                for (int k = 0; k < 6; ++k) pr->gradient[k] = (float)acc[21 + k];
                pr->n_visible = n_vis; pr->used = 3;
            }
        }
        const bool resume = ps->phase == 3;
        const bool had_h = ps->want_h != 0;                 // what the pass that just ran computed
        double diff_error = 0.0;
        int accepted = 0;
        if (!resume) {
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
            if (had_h) {
                for (int k = 0; k < 21; ++k) ps->Hc[k] = (float)acc[k];
                for (int k = 0; k < 6; ++k) ps->gc[k] = (float)acc[21 + k];
            }
            ps->nvis_c = n_vis;
        }
        if (rec) {
            rec->err2 = e2; rec->n_valid = n_valid; rec->n_visible = n_vis; rec->level = level;
            rec->err2_depth = 0.0; rec->n_valid_depth = 0; rec->reserved = 0;
            if (P.occlusion != 0) {
                rec->err2 = acc[27]; rec->err2_depth = acc[28];
                rec->n_valid = P.occlusion == 1 ? cnt[1] : cnt[2]; rec->n_valid_depth = cnt[2];
            }
            rec->it = ps->it; rec->accepted = accepted; rec->used = had_h ? 3 : 1;   // bit 1: normal equations recorded
            for (int k = 0; k < 16; ++k) rec->pose[k] = ps->pose_eval[k];
            for (int k = 0; k < 21; ++k) rec->hessian[k] = had_h ? (float)acc[k] : 0.f;
            for (int k = 0; k < 6; ++k) rec->gradient[k] = had_h ? (float)acc[21 + k] : 0.f;
            rec->pad = 0.f;
        }
        }   // !resume
        ps->phase = 1;
        r360_zero_acc(g.acc, g.cnt, p);

        if (!resume) {
            // while (it < maxIters && update_pose.norm() > tol_update && diff_error > tol_residual)
            const float* u = ps->upd;
            const float na = u[0] * u[0] + (u[1] * u[1] + u[2] * u[2]);
            const float nb = u[3] * u[3] + (u[4] * u[4] + u[5] * u[5]);
            const float unorm = sqrtf(na + nb);
            const bool go = ps->it < P.max_iters && (double)unorm > P.tol_update && diff_error > P.tol_residual;
            if (!go) {
                ps->iters[level] = ps->it;                  // RPI.h:4772
                ps->active = 0;
                continue;
            }
            if (!had_h) {
                // mispredicted: the candidate was accepted and the loop goes on, but its pass was error-only --
                // run the fused pass at the same pose before the loop body (one extra pass, same results)
                ps->phase = 3;
                ps->want_h = 1;
                ps->extra += 1;
                continue;
            }
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
        // Which pass does this candidate need?  The reference evaluates errorPhotoICP_sphere(candidate) and only
        // if the step is accepted AND the loop goes on calcHessGrad_sphere at the same pose.  The Gauss-Newton
        // model predicts the new sum of squares, e2 + g^T update (= e2 - g^T H^-1 g); when the predicted RMS
        // decrease is below tol_residual the step will most likely end the level, so its pass is error-only
        // (fewer instructions, no accumulator registers).  A misprediction costs one extra fused pass (phase 3)
        // and changes no result: every decision is still taken on the exactly evaluated error.
        ps->want_h = 1;
        if (g.speculate && P.occlusion == 0 && ps->extra < R360_SPEC_EXTRA) {    // the host enqueues max_iters + 1 + R360_SPEC_EXTRA passes
            double gu = 0.0;
            for (int k = 0; k < 6; ++k) gu += (double)ps->gc[k] * (double)upd[k];
            const double e2p = ps->err2 + gu;
            const double ep = sqrt((e2p > 0.0 ? e2p : 0.0) / (double)ps->n_valid);
            if (ps->error - ep < P.tol_residual * (double)g.spec_margin) ps->want_h = 0;
        }
    }
    r360_compact_when_last(g);
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
        ps->it = 0; ps->phase = 0; ps->active = 0; ps->status = R360_PAIR_OK; ps->ev = 0; ps->want_h = 1; ps->extra = 0;
        ps->src = src_idx[p]; ps->trg = trg_idx[p];
        for (int k = 0; k < R360_MAX_LEVELS; ++k) { ps->iters[k] = 0; ps->passes[k] = 0; }
    }
}

__global__ void k_finalize(R360GnArgs g, r360_result* __restrict__ out, int rows, int cols, int pair_id0) {
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
        r->pair_id = pair_id0 + p;
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

// =========================================================================== K0: Frame360 ingest
// Frame360::stitchSphericalImage (Frame360.h:386-405, 1099-1148): one thread per sphere pixel looks
// its ray up in the sensor that owns its column band (stitch_math.h) and copies the nearest-below
// sensor pixel; depth is rescaled from z to Euclidean range.  Writes exactly the RGB8 / depth-u16
// sphere images the pyramid kernels (K1) read, so ingest + pyramids stay on the device.
__global__ void __launch_bounds__(256)
k_stitch(R360StitchArgs a, const uint8_t* __restrict__ sensor_rgb, const uint16_t* __restrict__ sensor_depth,
         uint8_t* __restrict__ rgb, uint16_t* __restrict__ depth_mm) {
    const R360StitchGeom g = a.g;
    const int f = blockIdx.y;
    const int n = g.rows * g.cols;
    const size_t spx = (size_t)g.size_h * g.size_w;
    const uint8_t* __restrict__ srgb = sensor_rgb + (size_t)f * 8 * spx * 3;
    const uint16_t* __restrict__ sdep = sensor_depth + (size_t)f * 8 * spx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / g.cols, c = i - r * g.cols;
        const int s = 7 - c / g.size_h;
        float sp, cp, st, ct;
        r360_sincosf((g.offset_phi - r) * g.angle_pixel, &sp, &cp);
        r360_sincosf((c + g.offset_theta) * g.angle_pixel, &st, &ct);
        int ui, vi;
        float sc;
        uint8_t p0 = 0, p1 = 0, p2 = 0;
        uint16_t d = 0;
        if (r360_stitch_pixel(g, a.Rt_inv[s], sp, cp, st, ct, &ui, &vi, &sc)) {
            const size_t j = (size_t)s * spx + (size_t)vi * g.size_w + ui;
            p0 = srgb[3 * j]; p1 = srgb[3 * j + 1]; p2 = srgb[3 * j + 2];
            d = r360_stitch_range(sdep[j], sc);
        }
        const size_t o = (size_t)f * n + i;
        rgb[3 * o] = p0; rgb[3 * o + 1] = p1; rgb[3 * o + 2] = p2;
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
    const long long n_thr = (long long)(cols / 2) * ((rows / 2 + R360_DOWN_R - 1) / R360_DOWN_R);
    dim3 grid(r360_blocks(n_thr, 256, sm_count * 8), n_frames);
    k_down<<<grid, 256, 0, st>>>(pyr, off_src, off_dst, rows, cols, min_d, max_d);
}
void r360_launch_texel(cudaStream_t st, float2* const* pyr, float* const* trg, long long off, int rows, int cols,
                       int n_sensors, int n_frames, int sm_count) {
    dim3 grid(r360_blocks((long long)rows * cols / 2, 256, sm_count * 8), n_frames);
    k_texel<<<grid, 256, 0, st>>>(pyr, trg, off, rows, cols, r360_mask_geom(cols, n_sensors));
}
void r360_launch_pyr_head(cudaStream_t st, const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m,
                          float2* const* l0_dst, float2* const* l1_dst, float* const* texel_dst, int rows, int cols,
                          float min_d, float max_d, int n_sensors, int n_frames) {
    const int tiles_x = (cols + R360_F0_TW - 1) / R360_F0_TW, tiles_y = (rows + R360_F0_TH - 1) / R360_F0_TH;
    dim3 grid(tiles_x * tiles_y, n_frames);
    if (depth_mm)
        k_pyr_head<0><<<grid, 256, 0, st>>>(rgb, depth_mm, depth_m, nullptr, 0, l0_dst, l1_dst, 0, texel_dst, 0, rows, cols,
                                                           tiles_x, min_d, max_d, r360_mask_geom(cols, n_sensors));
    else
        k_pyr_head<1><<<grid, 256, 0, st>>>(rgb, depth_mm, depth_m, nullptr, 0, l0_dst, l1_dst, 0, texel_dst, 0, rows, cols,
                                                           tiles_x, min_d, max_d, r360_mask_geom(cols, n_sensors));
}
// Level l >= 1 of every frame of a chunk in one read of its plane: the level's target texels (frames whose entry of
// `tex` is not null; tex == nullptr: none) and level l + 1 (off_next >= 0).
void r360_launch_pyr_mid(cudaStream_t st, float2* const* pyr, float* const* tex, long long off, long long off_next, int rows, int cols,
                         float min_d, float max_d, int n_sensors, int n_frames) {
    const int tiles_x = (cols + R360_F0_TW - 1) / R360_F0_TW, tiles_y = (rows + R360_F0_TH - 1) / R360_F0_TH;
    dim3 grid(tiles_x * tiles_y, n_frames);
    k_pyr_head<R360_IN_PLANE><<<grid, 256, 0, st>>>(nullptr, nullptr, nullptr, pyr, off, nullptr, off_next >= 0 ? pyr : nullptr,
                                                                   off_next >= 0 ? off_next : 0, tex, off * R360_TEXEL_FLOATS, rows, cols, tiles_x,
                                                                   min_d, max_d, r360_mask_geom(cols, n_sensors));
}
// The pass kernel's pipeline slots need more than the 48 KB default of dynamic shared memory.
template <int METHOD, bool WITH_H>
static cudaError_t r360_pass_attr() {
    return cudaFuncSetAttribute(k_pass<METHOD, WITH_H>, cudaFuncAttributeMaxDynamicSharedMemorySize, R360_PASS_DYN_SMEM);
}
cudaError_t r360_pass_init() {
    cudaError_t e = r360_pass_attr<R360_PHOTO_CONSISTENCY, true>();
    if (e == cudaSuccess) e = r360_pass_attr<R360_DEPTH_CONSISTENCY, true>();
    if (e == cudaSuccess) e = r360_pass_attr<R360_PHOTO_DEPTH, true>();
    if (e == cudaSuccess) e = r360_pass_attr<R360_PHOTO_CONSISTENCY, false>();
    if (e == cudaSuccess) e = r360_pass_attr<R360_DEPTH_CONSISTENCY, false>();
    if (e == cudaSuccess) e = r360_pass_attr<R360_PHOTO_DEPTH, false>();
    return e;
}
template <bool WITH_H>
static void r360_launch_pass_t(cudaStream_t st, const R360PassArgs& a, int grid) {
    switch (a.params.method) {
        case R360_PHOTO_CONSISTENCY: k_pass<R360_PHOTO_CONSISTENCY, WITH_H><<<grid, R360_PASS_THREADS, R360_PASS_DYN_SMEM, st>>>(a); break;
        case R360_DEPTH_CONSISTENCY: k_pass<R360_DEPTH_CONSISTENCY, WITH_H><<<grid, R360_PASS_THREADS, R360_PASS_DYN_SMEM, st>>>(a); break;
        default: k_pass<R360_PHOTO_DEPTH, WITH_H><<<grid, R360_PASS_THREADS, R360_PASS_DYN_SMEM, st>>>(a); break;
    }
}
void r360_launch_pass(cudaStream_t st, const R360PassArgs& a, int grid, bool with_h) {
    if (with_h) r360_launch_pass_t<true>(st, a, grid);
    else r360_launch_pass_t<false>(st, a, grid);
}
void r360_launch_warp_dump(cudaStream_t st, const R360PassArgs& a, int pair, int32_t* r_idx, int32_t* c_idx,
                           uint8_t* vp, uint8_t* vd, int sm_count) {
    k_warp_dump<<<r360_blocks((a.lv.n + 1) / 2, 256, sm_count * 8), 256, 0, st>>>(a, pair, a.params.method, r_idx, c_idx, vp, vd);
}
void r360_launch_index_stats(cudaStream_t st, const R360PassArgs& a, int pair, unsigned long long* out, int sm_count) {
    k_index_stats<<<r360_blocks((a.lv.n + 1) / 2, 256, sm_count * 8), 256, 0, st>>>(a, pair, out);
}
void r360_launch_pairs_init(cudaStream_t st, const R360GnArgs& g, const int32_t* src_idx, const int32_t* trg_idx,
                            const float* init_pose) {
    k_pairs_init<<<r360_blocks(g.n_pairs, 128, 1024), 128, 0, st>>>(g, src_idx, trg_idx, init_pose);
}
void r360_launch_level_begin(cudaStream_t st, const R360GnArgs& g, int level) {
    k_level_begin<<<r360_blocks(g.n_pairs, 64, 1024), 64, 0, st>>>(g, level);
}
void r360_launch_gn_step(cudaStream_t st, const R360GnArgs& g, int level) {
    k_gn_step<<<r360_blocks(g.n_pairs, 1, 1024), 32, 0, st>>>(g, level);
}
void r360_launch_finalize(cudaStream_t st, const R360GnArgs& g, r360_result* out, int rows, int cols, int pair_id0) {
    k_finalize<<<r360_blocks(g.n_pairs, 128, 1024), 128, 0, st>>>(g, out, rows, cols, pair_id0);
}
void r360_launch_stitch(cudaStream_t st, const R360StitchArgs& a, const uint8_t* sensor_rgb, const uint16_t* sensor_depth,
                        uint8_t* rgb, uint16_t* depth_mm, int n_frames, int sm_count) {
    dim3 grid(r360_blocks((long long)a.g.rows * a.g.cols, 256, sm_count * 8), n_frames);
    k_stitch<<<grid, 256, 0, st>>>(a, sensor_rgb, sensor_depth, rgb, depth_mm);
}
void r360_launch_synth(cudaStream_t st, int kind, int first_id, int rows, int cols, const float* cams, int n_frames,
                       uint8_t* rgb, uint16_t* depth_mm, int sm_count) {
    dim3 grid(r360_blocks((long long)rows * cols, 256, sm_count * 8), n_frames);
    k_synth<<<grid, 256, 0, st>>>(kind, first_id, rows, cols, cams, rgb, depth_mm);
}

// =========================================================================== pinhole registration
#include "r360_pinhole.cuh"
