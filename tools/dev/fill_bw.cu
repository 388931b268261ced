// Write-bandwidth probe (development tool): how fast can a B200 WRITE HBM, and does the store flavour matter?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fill_bw fill_bw.cu && ./fill_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_st128(float4* p, size_t n) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_st128_cs(float4* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "f"(1.f) : "memory");
}
__global__ void k_st256_ef(float4* p, size_t n) {                // 32-byte stores, L2 evict-first
    const size_t n2 = n / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p + 2 * i), "r"(0x3f800000) : "memory");
}
__global__ void k_st256(float4* p, size_t n) {                    // 32-byte stores (sm_100)
    const size_t n2 = n / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p + 2 * i), "r"(0x3f800000) : "memory");
}
// every CTA fills a 16 KB shared buffer once and bulk-stores it (TMA, shared -> global) over and over
__global__ void k_tma(char* p, size_t bytes) {
    extern __shared__ __align__(128) char sm[];
    const int CH = 16384;
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
        for (size_t off = (size_t)blockIdx.x * CH; off + CH <= bytes; off += (size_t)gridDim.x * CH) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void k_copy128(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_read128(const float4* __restrict__ a, float* out, size_t n) {
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const float4 v = a[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 123.456f) *out = s;
}
template <typename F> float best_ms(F f, int reps = 8) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    const size_t bytes = 4ull << 30, n = bytes / 16;
    char *a, *b; float* o;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, 4));
    CK(cudaMemset(a, 1, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    for (int mult : {4, 8, 16}) {
        const int g = sms * mult;
        printf("grid %d x 256\n", g);
        printf("  st.v4           %7.0f GB/s\n", bytes / (best_ms([&] { k_st128<<<g, 256>>>((float4*)b, n); }) * 1e6));
        printf("  st.cs.v4        %7.0f GB/s\n", bytes / (best_ms([&] { k_st128_cs<<<g, 256>>>((float4*)b, n); }) * 1e6));
        printf("  st.v8.evict_1st  %7.0f GB/s\n", bytes / (best_ms([&] { k_st256_ef<<<g, 256>>>((float4*)b, n); }) * 1e6));
        printf("  st.v8 (256 bit) %7.0f GB/s\n", bytes / (best_ms([&] { k_st256<<<g, 256>>>((float4*)b, n); }) * 1e6));
        printf("  TMA bulk store  %7.0f GB/s\n", bytes / (best_ms([&] { k_tma<<<g, 128, 16384>>>(b, bytes); }) * 1e6));
        printf("  copy (r+w)      %7.0f GB/s\n", 2.0 * bytes / (best_ms([&] { k_copy128<<<g, 256>>>((const float4*)a, (float4*)b, n); }) * 1e6));
        printf("  read            %7.0f GB/s\n", bytes / (best_ms([&] { k_read128<<<g, 256>>>((const float4*)a, o, n); }) * 1e6));
    }
    printf("cudaMemset        %7.0f GB/s\n", bytes / (best_ms([&] { cudaMemsetAsync(b, 0, bytes); }) * 1e6));
    printf("cudaMemcpy D2D    %7.0f GB/s\n", 2.0 * bytes / (best_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }) * 1e6));
    CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    return 0;
}
