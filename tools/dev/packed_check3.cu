// Dev check 3: full packed index path vs scalar pinned, stage by stage (values exposed via copies of the code).
#include <cstdio>
#include <cstdint>
#include "../../rgbd360_b200/csrc/r360_device.cuh"
__device__ uint32_t rng(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
__device__ float rnd01(uint64_t& s) { return (rng(s) >> 7) * (1.0f / 16777216.0f); }
__global__ void k(unsigned long long* out, int iters, const float* Tg, float one) {
    uint64_t s = 0x1234567ull + 7919ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad[12] = {0};
    float T[16];
    for (int q = 0; q < 16; ++q) T[q] = Tg[q];
    const float res = (float)(2 * R360_PI_D / 2048), res_inv = 1 / res, half_rows = 511.5f;
    for (int it = 0; it < iters; ++it) {
        float X[3] = { (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f };
        float Y[3] = { (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f, (rnd01(s) - 0.5f) * 8.f };
        R360Geo2 g; int r[2], c[2];
        bool bad2[2]; r360_index_pair_packed(T, make_float2(X[0], Y[0]), make_float2(X[1], Y[1]), make_float2(X[2], Y[2]), res_inv, half_rows, one, g, r, c, bad2); unsigned need = (bad2[0] ? 1u : 0u) | (bad2[1] ? 2u : 0u);
        // scalar pinned, step by step (copy of r360_index_exact_inl)
        const float px = ((T[0] * X[0] + T[4] * X[1]) + T[8] * X[2]) + T[12];
        const float py = ((T[1] * X[0] + T[5] * X[1]) + T[9] * X[2]) + T[13];
        const float pz = ((T[2] * X[0] + T[6] * X[1]) + T[10] * X[2]) + T[14];
        const float dist = sqrtf(px * px + (py * py + pz * pz));
        const float dinv = 1.f / dist;
        const float phi = r360_asinf(px * dinv);
        const float theta = (float)((double)r360_atan2f(py, pz) + R360_PI_D);
        const float vr = half_rows - phi * res_inv;
        const float vc = theta * res_inv;
        const int re = r360_round_to_int_dev(vr), ce = r360_round_to_int_dev(vc);
        if (px != g.px.x || py != g.py.x || pz != g.pz.x) ++bad[0];
        if (dist != g.dist.x) ++bad[1];
        if (dinv != g.dinv.x) ++bad[2];
        if (!(need & 1)) { if (re != r[0]) ++bad[3]; if (ce != c[0]) ++bad[4]; }
        else ++bad[5];
        ++bad[7];
    }
    for (int q = 0; q < 8; ++q) atomicAdd(&out[q], bad[q]);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    float Th[16] = { 0.9987f, 0.0499f, -0.0120f, 0, -0.0497f, 0.9986f, 0.0170f, 0, 0.0128f, -0.0164f, 0.9998f, 0, 0.01f, -0.02f, 0.03f, 1 };
    float* Td; cudaMalloc(&Td, 64); cudaMemcpy(Td, Th, 64, cudaMemcpyHostToDevice);
    k<<<592, 256>>>(d, 1000, Td, 1.0f);
    unsigned long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("n=%llu bad: p %llu dist %llu dinv %llu | r %llu c %llu (flagged %llu) (%s)\n", h[7], h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
