// Dev check: packed IEEE sequences vs the compiler's scalar operators, on random inputs.
#include <cstdio>
#include <cstdint>
#include "../../rgbd360_b200/csrc/r360_device.cuh"
__device__ uint32_t rng(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
__device__ float rnd01(uint64_t& s) { return (rng(s) >> 7) * (1.0f / 16777216.0f); }
__global__ void k(unsigned long long* out, int iters) {
    uint64_t s = 0x1234567ull + 7919ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad[8] = {0,0,0,0,0,0,0,0};
    for (int it = 0; it < iters; ++it) {
        float a = ldexpf(rnd01(s) + 0.5f, (int)(rng(s) % 40) - 20), b = ldexpf(rnd01(s) + 0.5f, (int)(rng(s) % 40) - 20);
        float2 A = make_float2(a, b), B = make_float2(b, a);
        float2 sq = f2sqrt_rn(A);
        if (sq.x != sqrtf(a) || sq.y != sqrtf(b)) ++bad[0];
        float2 rc = f2rcp_rn(A);
        if (rc.x != 1.f / a || rc.y != 1.f / b) ++bad[1];
        float mn = fminf(a, b), mx = fmaxf(a, b);
        float2 q = f2div_rn(make_float2(mn, mn * 0.37f), make_float2(mx, mx));
        if (q.x != mn / mx || q.y != (mn * 0.37f) / mx) ++bad[2];
        // asin / atan2 through the full index functions
        float X[3] = { (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f };
        float Y[3] = { (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 0.01f, (rnd01(s) - 0.5f) * 0.01f };
        float T[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.01f, -0.02f, 0.03f, 1 };
        const float res = (float)(2 * R360_PI_D / 2048), res_inv = 1 / res, half_rows = 511.5f;
        R360Geo2 g; int r[2], c[2];
        unsigned need = r360_index_pair_packed(T, make_float2(X[0], Y[0]), make_float2(X[1], Y[1]), make_float2(X[2], Y[2]), res_inv, half_rows, g, r, c);
        int re, ce; float vr, vc;
        r360_index_exact_inl(T, X, res_inv, half_rows, re, ce, vr, vc);
        if (!(need & 1)) { if (re != r[0]) ++bad[3]; if (ce != c[0]) ++bad[4]; }
        r360_index_exact_inl(T, Y, res_inv, half_rows, re, ce, vr, vc);
        if (!(need & 2)) { if (re != r[1]) ++bad[5]; if (ce != c[1]) ++bad[6]; }
        ++bad[7];
    }
    for (int q = 0; q < 8; ++q) atomicAdd(&out[q], bad[q]);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    k<<<592, 256>>>(d, 2000);
    unsigned long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("n=%llu bad: sqrt %llu rcp %llu div %llu | r0 %llu c0 %llu r1(polar) %llu c1(polar) %llu  (%s)\n", h[7], h[0], h[1], h[2], h[3], h[4], h[5], h[6], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
