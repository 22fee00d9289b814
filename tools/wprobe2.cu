// Write stream + a small read share (B200): does a 4 % read share from HBM cost the plane-store kernels their last 15 %?
// Each CTA (256 threads, 4 per SM) first reads `rd` bytes (0 .. 16 KB; from a 2 GB buffer = HBM, or from a 32 MB
// buffer = L2), waits for them, then writes 64 KB to each of two planes with 256-bit stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/wprobe2 tools/wprobe2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void store8(float* dst, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__global__ void __launch_bounds__(256, 4) wkernel(float* __restrict__ p0, float* __restrict__ p1, const float2* __restrict__ src,
                                                  long long src_mask, int rd_per_thread, int spin) {
    __shared__ float2 sm[4096];           // 32 KB, like the expansion: 4 CTAs / SM
    const long long span = 16384;
    const long long base = (long long)blockIdx.x * span;
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (u < rd_per_thread) v[u] = src[(((long long)blockIdx.x * 2048 + u * 256 + threadIdx.x)) & src_mask];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (u < rd_per_thread) sm[u * 256 + threadIdx.x] = v[u];
    __syncthreads();
    float s = rd_per_thread ? sm[(threadIdx.x * 7) & 255].x : 0.0f;
    for (long long o = threadIdx.x * 8; o < span; o += 2048) {
        float a[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned h = (unsigned)(base + o + i) * 2654435761u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            a[i] = __uint_as_float((h & 0x007fffffu) | 0x3f800000u) + s;
            w[i] = __uint_as_float(((h * 3266489917u) & 0x007fffffu) | 0x40000000u) + s;
        }
        for (int k = 0; k < spin; ++k)     // dependent FMAs: stand-in for the interpolation work between two stores
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0000001f, w[i] * 1e-9f);
        store8(p0 + base + o, a);
        store8(p1 + base + o, w);
    }
}
// Same work, but a CTA walks `tiles` tiles (tile = blockIdx.x + j * gridDim.x) and requests the reads of tile j + 1
// BEFORE it issues the stores of tile j: the reads are then ahead of the stores in the SM's memory queues.
__global__ void __launch_bounds__(256, 4) wkernel_pipe(float* __restrict__ p0, float* __restrict__ p1, const float2* __restrict__ src,
                                                       long long src_mask, int tiles) {
    __shared__ float2 sm[2][2048];
    const long long span = 16384;
    float2 v[2];
    long long tile = blockIdx.x;
#pragma unroll
    for (int u = 0; u < 2; ++u) v[u] = src[((tile * 2048 + u * 256 + threadIdx.x)) & src_mask];
    for (int j = 0; j < tiles; ++j, tile += gridDim.x) {
        float2* buf = sm[j & 1];
#pragma unroll
        for (int u = 0; u < 2; ++u) buf[u * 256 + threadIdx.x] = v[u];
        __syncthreads();
        if (j + 1 < tiles) {
#pragma unroll
            for (int u = 0; u < 2; ++u) v[u] = src[(((tile + gridDim.x) * 2048 + u * 256 + threadIdx.x)) & src_mask];
        }
        const float s = buf[(threadIdx.x * 7) & 255].x;
        const long long base = tile * span;
        for (long long o = threadIdx.x * 8; o < span; o += 2048) {
            float a[8], w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned h = (unsigned)(base + o + i) * 2654435761u;
                h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                a[i] = __uint_as_float((h & 0x007fffffu) | 0x3f800000u) + s;
                w[i] = __uint_as_float(((h * 3266489917u) & 0x007fffffu) | 0x40000000u) + s;
            }
            store8(p0 + base + o, a);
            store8(p1 + base + o, w);
        }
    }
}
// Reads whose result nothing waits for until the very end of the CTA (mode 0), or only ONE warp reads and the others
// wait at the barrier (mode 1: 128-bit loads, 4 KB by 32 threads x 8), or ld.global.nc.L1::no_allocate (mode 2).
__global__ void __launch_bounds__(256, 4) wkernel_var(float* __restrict__ p0, float* __restrict__ p1, const float4* __restrict__ src,
                                                      long long src_mask, int mode) {
    __shared__ float4 sm[2048];
    const long long span = 16384;
    const long long base = (long long)blockIdx.x * span;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.0f;
    if (mode == 0) {
        v = src[((long long)blockIdx.x * 256 + threadIdx.x) & src_mask];        // 4 KB per CTA, consumed at the end
    } else if (mode >= 3) {
        // same total read bytes, but only every 2^(mode-1)-th CTA reads (16 KB for mode 3, 64 KB for mode 5), nothing waits
        const int every = 1 << (mode - 1);
        if ((blockIdx.x & (every - 1)) == 0) {
            for (int u = 0; u < every; ++u) {
                const float4 q = src[((long long)blockIdx.x * 256 + u * 256 + threadIdx.x) & src_mask];
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
        }
    } else if (mode == 1) {
        if (threadIdx.x < 32) {
            float4 q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = src[((long long)blockIdx.x * 256 + u * 32 + threadIdx.x) & src_mask];
#pragma unroll
            for (int u = 0; u < 8; ++u) sm[u * 32 + threadIdx.x] = q[u];
        }
        __syncthreads();
        s = sm[threadIdx.x].x;
    } else {
        const float4* p = src + (((long long)blockIdx.x * 256 + threadIdx.x) & src_mask);
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        sm[threadIdx.x] = v;
        __syncthreads();
        s = sm[(threadIdx.x * 7) & 255].x;
    }
    for (long long o = threadIdx.x * 8; o < span; o += 2048) {
        float a[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned h = (unsigned)(base + o + i) * 2654435761u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            a[i] = __uint_as_float((h & 0x007fffffu) | 0x3f800000u) + s;
            w[i] = __uint_as_float(((h * 3266489917u) & 0x007fffffu) | 0x40000000u) + s;
        }
        store8(p0 + base + o, a);
        store8(p1 + base + o, w);
    }
    if ((mode == 0 || mode >= 3) && v.x + v.y + v.z + v.w == 12345.678f) p0[base] = 0.0f;
}
int main() {
    const long long n = 1ll << 31;            // floats per plane (8 GiB)
    float *a, *b; float2* src;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&src, (1ll << 28) * 8);
    cudaMemset(src, 0, (1ll << 28) * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)(n / 16384);
    struct { const char* name; long long mask; int rd; int spin; } cases[] = {
        {"no reads", 0, 0, 0}, {"2 KB/CTA from HBM", (1ll << 28) - 1, 1, 0}, {"4 KB/CTA from HBM", (1ll << 28) - 1, 2, 0},
        {"16 KB/CTA from HBM", (1ll << 28) - 1, 8, 0}, {"4 KB/CTA from L2", (1ll << 22) - 1, 2, 0},
        {"16 KB/CTA from L2", (1ll << 22) - 1, 8, 0}, {"no reads, 8 FMA rounds", 0, 0, 8},
        {"4 KB HBM, 8 FMA rounds", (1ll << 28) - 1, 2, 8}, {"no reads, 16 FMA rounds", 0, 0, 16}};
    for (auto& c : cases) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            wkernel<<<grid, 256>>>(a, b, src, c.mask, c.rd, c.spin);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("%-24s : stores %8.1f GB/s  (%.3f ms)\n", c.name, 8.0 * n / best / 1e6, best);
    }
    for (int mode = 0; mode < 1; ++mode)
        for (long long mask : {(1ll << 27) - 1, (1ll << 21) - 1, (1ll << 14) - 1, (1ll << 10) - 1}) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                wkernel_var<<<grid, 256>>>(a, b, reinterpret_cast<const float4*>(src), mask, mode);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("mode %d (%s), 4 KB from %s : stores %8.1f GB/s\n", mode,
                   mode == 0 ? "nothing waits for the reads" : mode == 1 ? "one warp reads" : mode == 2 ? "ld.nc no_allocate" : mode == 3 ? "every 4th CTA reads 16 KB" : mode == 4 ? "every 8th CTA reads 32 KB" : "every 16th CTA reads 64 KB",
                   mask > (1ll << 24) ? "HBM" : mask > (1ll << 20) ? "32 MB" : mask > (1ll << 12) ? "256 KB" : "16 KB", 8.0 * n / best / 1e6);
        }
    for (int tiles : {2, 4, 8, 16})
        for (long long mask : {(1ll << 28) - 1, (1ll << 22) - 1}) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                wkernel_pipe<<<grid / tiles, 256>>>(a, b, src, mask, tiles);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("pipelined, %2d tiles/CTA, 4 KB from %s : stores %8.1f GB/s\n", tiles, mask > (1ll << 24) ? "HBM" : "L2 ", 8.0 * n / best / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
