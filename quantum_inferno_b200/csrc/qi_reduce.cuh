// qi_reduce.cuh -- block-level reductions (fp64 accumulators; SURVEY 7.3-g: power spans >55 bits).
#pragma once
#include "qi_platform.cuh"

namespace qi {

QI_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
QI_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
QI_DEV double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double t = __shfl_down_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// Sum over the CTA; result valid in thread 0.  scratch: >= 32 doubles of shared memory.
QI_DEV double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}
QI_DEV double block_max(double v, double* scratch) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : -1.0e300;
    if (w == 0) v = warp_max(v);
    return v;
}

// atomic max on non-negative doubles via their (order-preserving) bit pattern
QI_DEV void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

}  // namespace qi
