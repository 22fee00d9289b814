// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny CUDA-execution-model emulator so the kernel sources under
// quantum_inferno_b200/csrc/*.cu can be compiled with plain g++ (-DQI_EMUL)
// and their LOGIC debugged on a box without a GPU.  Every CUDA thread of a
// block is a ucontext fiber; __syncthreads()/__shfl_*_sync() are cooperative
// yield points; blocks are spread over host threads.  It is slow, it is not a
// product path, and nothing in the quantum_inferno_b200 package can load it:
// only tests/emul/ builds and dlopens the resulting libqi_emul.so.
#pragma once
#ifndef QI_EMUL
#error "cuda_emul.h is only for -DQI_EMUL builds"
#endif

#include <ucontext.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
// one block runs on one host thread at a time -> thread_local statics behave like static __shared__
#define __shared__ static thread_local

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace qi_emul {

struct WarpState {
    uint64_t slot[32];
    int arrived = 0, gen = 0, arrived2 = 0, gen2 = 0, live = 0;
};

struct BlockState;
struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    int linear = 0;
    bool done = false;
    BlockState* blk = nullptr;
};

struct BlockState {
    uint3 bid;
    dim3 bdim, gdim;
    unsigned char* smem = nullptr;
    int nthreads = 0, live = 0;
    int bar_arrived = 0, bar_gen = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    ucontext_t sched;
    int current = 0;
    const std::function<void()>* body = nullptr;
};

extern thread_local Fiber* cur;

void yield();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

inline void syncthreads() {
    BlockState* b = cur->blk;
    b->bar_arrived++;
    if (b->bar_arrived >= b->live) {
        b->bar_arrived = 0;
        b->bar_gen++;
    } else {
        int g = b->bar_gen;
        while (b->bar_gen == g) yield();
    }
}

inline uint64_t warp_exchange(uint64_t v, int src_lane) {
    BlockState* b = cur->blk;
    int lane = cur->linear & 31;
    WarpState& w = b->warps[cur->linear >> 5];
    w.slot[lane] = v;
    w.arrived++;
    if (w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    else { int g = w.gen; while (w.gen == g) yield(); }
    uint64_t r = w.slot[src_lane & 31];
    w.arrived2++;
    if (w.arrived2 >= w.live) { w.arrived2 = 0; w.gen2++; }
    else { int g = w.gen2; while (w.gen2 == g) yield(); }
    return r;
}

template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

}  // namespace qi_emul

#define threadIdx (qi_emul::cur->tid)
#define blockIdx (qi_emul::cur->blk->bid)
#define blockDim (qi_emul::cur->blk->bdim)
#define gridDim (qi_emul::cur->blk->gdim)
#define warpSize 32

static inline void __syncthreads() { qi_emul::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { qi_emul::warp_exchange(0, 0); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return qi_emul::from_bits<T>(qi_emul::warp_exchange(qi_emul::to_bits(v), src));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    int lane = qi_emul::cur->linear & 31;
    return qi_emul::from_bits<T>(qi_emul::warp_exchange(qi_emul::to_bits(v), lane ^ m));
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
    int lane = qi_emul::cur->linear & 31;
    int src = lane + d > 31 ? lane : lane + d;
    return qi_emul::from_bits<T>(qi_emul::warp_exchange(qi_emul::to_bits(v), src));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d, int = 32) {
    int lane = qi_emul::cur->linear & 31;
    int src = lane - d < 0 ? lane : lane - d;
    return qi_emul::from_bits<T>(qi_emul::warp_exchange(qi_emul::to_bits(v), src));
}
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

// atomics (blocks run concurrently on host threads -> use real atomics)
static inline double atomicAdd(double* p, double v) {
    uint64_t* up = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(up, __ATOMIC_RELAXED), nw;
    double od;
    do { std::memcpy(&od, &old, 8); double nd = od + v; std::memcpy(&nw, &nd, 8); }
    while (!__atomic_compare_exchange_n(up, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    return od;
}
static inline float atomicAdd(float* p, float v) {
    uint32_t* up = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(up, __ATOMIC_RELAXED), nw;
    float od;
    do { std::memcpy(&od, &old, 4); float nd = od + v; std::memcpy(&nw, &nd, 4); }
    while (!__atomic_compare_exchange_n(up, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    return od;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}

// math intrinsics used by the kernels
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
static inline double sinpi(double x) { return std::sin(M_PI * x); }
static inline double cospi(double x) { return std::cos(M_PI * x); }
static inline float exp2f_(float x) { return std::exp2(x); }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::exp; using std::log2; using std::hypot; using std::rint; using std::sqrt; using std::fabs; using std::fma; using std::floor;
using std::min; using std::max;

#define QI_EMUL_DYN_SMEM (qi_emul::cur->blk->smem)
