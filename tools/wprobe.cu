// Write-only HBM bandwidth probe (B200): 256-bit stores, constant vs incompressible data, one or two output planes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/wprobe tools/wprobe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void store8(float* dst, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// each CTA (256 threads) writes `span` floats (multiple of 2048) to each of `planes` planes; tile index = blockIdx.x
template <int RANDOM>
__global__ void __launch_bounds__(256) wkernel(float* __restrict__ p0, float* __restrict__ p1, long long span, int planes) {
    const long long base = (long long)blockIdx.x * span;
    for (long long o = threadIdx.x * 8; o < span; o += 2048) {
        float v[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned h = (unsigned)(base + o + i) * 2654435761u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v[i] = RANDOM ? __uint_as_float((h & 0x007fffffu) | 0x3f800000u) : 1.25f;
            w[i] = RANDOM ? __uint_as_float(((h * 3266489917u) & 0x007fffffu) | 0x40000000u) : 2.5f;
        }
        store8(p0 + base + o, v);
        if (planes > 1) store8(p1 + base + o, w);
    }
}
int main() {
    const long long n = 1ll << 32;            // floats per plane (16 GiB)
    float *a, *b;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int planes = 1; planes <= 2; ++planes)
        for (int rnd = 0; rnd < 2; ++rnd)
            for (long long span : {16384ll, 65536ll, 1048576ll}) {
                const long long per = n / planes;             // same total bytes either way
                const unsigned grid = (unsigned)(per / span);
                float best = 1e9f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    if (rnd) wkernel<1><<<grid, 256>>>(a, b, span, planes); else wkernel<0><<<grid, 256>>>(a, b, span, planes);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                printf("planes %d  data %-8s span %8lld floats/CTA : %8.1f GB/s\n", planes, rnd ? "random" : "constant", span,
                       4.0 * n / best / 1e6);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
