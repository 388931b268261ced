// Dev check 2: stage-by-stage comparison of the packed index path with the scalar pinned one.
#include <cstdio>
#include <cstdint>
#include "../../rgbd360_b200/csrc/r360_device.cuh"
__device__ uint32_t rng(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
__device__ float rnd01(uint64_t& s) { return (rng(s) >> 7) * (1.0f / 16777216.0f); }
__global__ void k(unsigned long long* out, int iters) {
    uint64_t s = 0x1234567ull + 7919ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad[8] = {0,0,0,0,0,0,0,0};
    for (int it = 0; it < iters; ++it) {
        float x = (rnd01(s) - 0.5f) * 2.f, y = (rnd01(s) - 0.5f) * 8.f, z = (rnd01(s) - 0.5f) * 8.f;
        float x2 = (rnd01(s) - 0.5f) * 2.f;
        // asin: packed (as in r360_index_pair_packed) vs r360_asinf
        {
            float2 sx = make_float2(x, x2);
            const float ax0 = fabsf(sx.x), ax1 = fabsf(sx.y);
            const bool big0 = ax0 > 0.5f, big1 = ax1 > 0.5f;
            const float2 za = f2mul(sx, sx);
            const float2 zb = f2mul(f2add(R360_F2(1.0f), make_float2(-ax0, -ax1)), R360_F2(0.5f));
            const float2 sq = f2sqrt_rn(zb);
            const float2 zz = make_float2(big0 ? zb.x : za.x, big1 ? zb.y : za.y);
            const float2 u = make_float2(big0 ? sq.x : sx.x, big1 ? sq.y : sx.y);
            float2 p = f2fma(zz, R360_F2(3.8206567683e-02f), R360_F2(2.6494211752e-02f));
            p = f2fma(zz, p, R360_F2(4.5010712250e-02f));
            p = f2fma(zz, p, R360_F2(7.4988090911e-02f));
            p = f2fma(zz, p, R360_F2(1.6666672766e-01f));
            const float2 t = f2fma(f2mul(u, zz), p, u);
            const float2 tb = f2add(f2fma(R360_F2(-2.0f), t, R360_F2(1.57079637050628662109375f)), R360_F2(-4.37113900018624283e-8f));
            const float2 phi = make_float2(big0 ? (sx.x < 0.0f ? -tb.x : tb.x) : t.x, big1 ? (sx.y < 0.0f ? -tb.y : tb.y) : t.y);
            float e0 = r360_asinf(x), e1 = r360_asinf(x2);
            if (phi.x != e0) { ++bad[0]; if (big0) ++bad[1]; }
            if (phi.y != e1) { ++bad[0]; if (big1) ++bad[1]; }
        }
        {
            float2 py = make_float2(y, z), pz = make_float2(z, y * 0.3f);
            const float ay0 = fabsf(py.x), ay1 = fabsf(py.y), az0 = fabsf(pz.x), az1 = fabsf(pz.y);
            const float2 mx = make_float2(fmaxf(az0, ay0), fmaxf(az1, ay1));
            const float2 mn = make_float2(fminf(az0, ay0), fminf(az1, ay1));
            const float2 q = f2div_rn(mn, mx);
            const float2 t2 = f2mul(q, q);
            float2 pa = f2fma(t2, R360_F2(-2.4470558835e-03f), R360_F2(1.3750389111e-02f));
            pa = f2fma(t2, pa, R360_F2(-3.6270357867e-02f));
            pa = f2fma(t2, pa, R360_F2(6.2843779659e-02f));
            pa = f2fma(t2, pa, R360_F2(-8.6731798886e-02f));
            pa = f2fma(t2, pa, R360_F2(1.1037996988e-01f));
            pa = f2fma(t2, pa, R360_F2(-1.4279111346e-01f));
            pa = f2fma(t2, pa, R360_F2(1.9999766029e-01f));
            pa = f2fma(t2, pa, R360_F2(-3.3333331951e-01f));
            float2 at = f2fma(f2mul(q, t2), pa, q);
            float b0 = at.x;
            const float2 a1 = f2add(f2add(R360_F2(1.57079637050628662109375f), f2neg(at)), R360_F2(-4.37113900018624283e-8f));
            at = make_float2(ay0 > az0 ? a1.x : at.x, ay1 > az1 ? a1.y : at.y);
            const float2 a2 = f2add(f2add(R360_F2(3.1415927410125732421875f), f2neg(at)), R360_F2(-8.74227800037248566e-8f));
            at = make_float2(pz.x < 0.0f ? a2.x : at.x, pz.y < 0.0f ? a2.y : at.y);
            float r0 = copysignf(at.x, py.x), r1 = copysignf(at.y, py.y);
            float e0 = r360_atan2f(py.x, pz.x), e1 = r360_atan2f(py.y, pz.y);
            if (r0 != e0) { ++bad[2]; if (ay0 > az0) ++bad[3]; if (pz.x < 0) ++bad[4]; }
            if (r1 != e1) ++bad[2];
            // base atan only (first octant)
            float qq = fminf(ay0, az0) / fmaxf(ay0, az0);
            float eb = r360_atan2f(qq, 1.0f);
            if (eb != b0) ++bad[5];
        }
        ++bad[7];
    }
    for (int q = 0; q < 8; ++q) atomicAdd(&out[q], bad[q]);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    k<<<592, 256>>>(d, 1000);
    unsigned long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("n=%llu bad: asin %llu (big %llu) | atan2 %llu (swap %llu, neg %llu) base %llu (%s)\n", h[7], h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
