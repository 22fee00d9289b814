// qi_pick.cu -- the step AFTER the time-frequency path: reduce a record or a [bands, time]
// plane that lives in HBM to a displayable / pickable summary without leaving the device.
//
// Replaces quantum_inferno/utilities/sampling.py:13-50 (subsample) and :87-120 (subsample_2d): block reductions
// "average" / "median" / "max" / "min" / "nth" of `factor` consecutive samples along the time axis, and the
// scan half of quantum_inferno/utilities/picker.py:34-53,108-149 (scale_signal_by_extraction_type,
// find_peaks_by_extraction_type, find_peaks_with_bits): extrema of a record and the plateau-aware local maxima
// of scipy.signal.find_peaks (scipy/signal/_peak_finding_utils.pyx::_local_maxima_1d) with the height test.
// The O(#peaks) distance selection stays on the host (quantum_inferno_b200/utilities/picker.py).
//
// All kernels are HBM-streaming reads: 4 or 8 B per input sample, 1/factor of that written.
#include <limits>
#include <string.h>
#include "qi_platform.cuh"
#include "qi_host.h"
#include "qi_reduce.cuh"

namespace qi {

constexpr int SUB_TILE = 4096;       // samples a CTA stages in shared memory (staged path)
constexpr int SUB_SMALL_MAX = 128;   // largest factor of the staged path; longer groups get one warp each
QI_HD int sub_slot(int i) { return i + (i >> 5); }   // one pad word per 32: strided group walks stay conflict-free

template <typename T> QI_DEV T quiet_nan() { return std::numeric_limits<T>::quiet_NaN(); }

// order-preserving unsigned keys of IEEE values (NaNs are handled separately by the callers)
QI_DEV unsigned key_of(float v) {
    unsigned u;
    memcpy(&u, &v, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
QI_DEV unsigned long long key_of(double v) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
QI_DEV float value_of(unsigned k) {
    const unsigned u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    float v;
    memcpy(&v, &u, 4);
    return v;
}
QI_DEV double value_of(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
    double v;
    memcpy(&v, &u, 8);
    return v;
}
template <typename T> struct key_traits;
template <> struct key_traits<float> { typedef unsigned type; static constexpr int bits = 32; };
template <> struct key_traits<double> { typedef unsigned long long type; static constexpr int bits = 64; };

// np.median of an even count: mean of the two middle values in the array's own dtype (numpy/lib/_function_base_impl.py::_median)
template <typename T> QI_DEV T mid_of(T lo, T hi) { return (lo + hi) / (T)2; }

// coalesced copy of ne consecutive samples into the padded tile (128-bit loads when the source allows)
template <typename T>
QI_DEV void sub_stage(T* tile, const T* __restrict__ src, int ne) {
    constexpr int V = 16 / (int)sizeof(T);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nv = ne / V;
        for (int q = threadIdx.x; q < nv; q += blockDim.x) {
            T v[V];
            if (sizeof(T) == 4) {
                const float4 w = *reinterpret_cast<const float4*>(src + (i64)q * V);
                v[0] = (T)w.x; v[1] = (T)w.y; v[V - 2] = (T)w.z; v[V - 1] = (T)w.w;
            } else {
                const double2 w = *reinterpret_cast<const double2*>(src + (i64)q * V);
                v[0] = (T)w.x; v[V - 1] = (T)w.y;
            }
#pragma unroll
            for (int e = 0; e < V; ++e) tile[sub_slot(q * V + e)] = v[e];
        }
        for (int i = nv * V + threadIdx.x; i < ne; i += blockDim.x) tile[sub_slot(i)] = src[i];
    } else {
        for (int i = threadIdx.x; i < ne; i += blockDim.x) tile[sub_slot(i)] = src[i];
    }
}

// ---------------------------------------------------------------- staged path: factor <= SUB_SMALL_MAX
// A CTA stages floor(SUB_TILE / factor) whole groups (coalesced, 128-bit when the source allows) and L = lanes-per-group
// threads (a power of two <= 32) reduce one group.  grid: (ceil(n_out / groups_per_tile), M).
template <typename T, int METHOD>
__global__ void __launch_bounds__(256)
subsample_small_kernel(const T* __restrict__ in, i64 stride, i64 n_out, int factor, int lanes, T* __restrict__ out,
                       i64 out_stride) {
    __shared__ __align__(16) T tile[SUB_TILE + SUB_TILE / 32 + 4];
    const i64 m = blockIdx.y;
    const int gpt = SUB_TILE / factor;
    const i64 g0 = (i64)blockIdx.x * gpt;
    const i64 left = n_out - g0;
    const int ng = left < gpt ? (int)left : gpt;
    const T* src = in + m * stride + g0 * factor;
    const int ne = ng * factor;
    sub_stage<T>(tile, src, ne);
    __syncthreads();
    const int gstep = blockDim.x / lanes;
    const int sub = threadIdx.x & (lanes - 1);
    const int k_hi = factor >> 1, k_lo = (factor & 1) ? k_hi : k_hi - 1;
    for (int gb = 0; gb < ng; gb += gstep) {                       // uniform trip count: the shuffles need every lane
        const int g = gb + threadIdx.x / lanes;
        const bool live = g < ng;
        const int base = live ? g * factor : 0;
        T res;
        if (METHOD == QI_SUB_AVERAGE) {
            double s = 0.0;
            if (live) for (int j = sub; j < factor; j += lanes) s += (double)tile[sub_slot(base + j)];
            for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            res = (T)(s / (double)factor);
        } else if (METHOD == QI_SUB_MAX || METHOD == QI_SUB_MIN) {
            T r = tile[sub_slot(base + (sub < factor ? sub : 0))];
            bool bad = false;
            if (live) for (int j = sub; j < factor; j += lanes) {
                const T v = tile[sub_slot(base + j)];
                bad |= (v != v);
                r = (METHOD == QI_SUB_MAX) ? (v > r ? v : r) : (v < r ? v : r);
            }
            for (int o = lanes >> 1; o > 0; o >>= 1) {
                const T v = __shfl_xor_sync(0xffffffffu, r, o);
                const int b = __shfl_xor_sync(0xffffffffu, (int)bad, o);
                bad |= (b != 0);
                r = (METHOD == QI_SUB_MAX) ? (v > r ? v : r) : (v < r ? v : r);
            }
            res = bad ? quiet_nan<T>() : r;
        } else {                                                   // median: the k-th order statistics by rank counting
            T lo = (T)0, hi = (T)0;
            int f_lo = 0, f_hi = 0, bad = 0;
            if (live) for (int i = sub; i < factor; i += lanes) {
                const T v = tile[sub_slot(base + i)];
                bad |= (v != v);
                int less = 0, eq = 0;
                for (int j = 0; j < factor; ++j) {
                    const T u = tile[sub_slot(base + j)];
                    less += (u < v);
                    eq += (u == v);
                }
                if (less <= k_lo && k_lo < less + eq) { lo = v; f_lo = 1; }
                if (less <= k_hi && k_hi < less + eq) { hi = v; f_hi = 1; }
            }
            for (int o = lanes >> 1; o > 0; o >>= 1) {
                const T olo = __shfl_xor_sync(0xffffffffu, lo, o), ohi = __shfl_xor_sync(0xffffffffu, hi, o);
                const int oflo = __shfl_xor_sync(0xffffffffu, f_lo, o), ofhi = __shfl_xor_sync(0xffffffffu, f_hi, o);
                bad |= __shfl_xor_sync(0xffffffffu, bad, o);
                if (!f_lo && oflo) { lo = olo; f_lo = 1; }
                if (!f_hi && ofhi) { hi = ohi; f_hi = 1; }
            }
            res = bad ? quiet_nan<T>() : ((factor & 1) ? hi : mid_of(lo, hi));
        }
        if (live && sub == 0) out[m * out_stride + g0 + g] = res;
    }
}

// ---------------------------------------------------------------- median of short groups: factor <= 32
// One thread per group: the group is read from the staged tile into P = 2^m >= factor registers (padded with +inf,
// which sorts behind every sample and leaves the middle ranks where they are) and sorted by a fully unrolled bitonic
// network of min / max pairs -- ~15 instructions per sample at P = 32 against ~130 for the rank counting above.
template <typename T, int P>
__global__ void __launch_bounds__(256)
subsample_median_sort_kernel(const T* __restrict__ in, i64 stride, i64 n_out, int factor, T* __restrict__ out,
                             i64 out_stride) {
    __shared__ __align__(16) T tile[SUB_TILE + SUB_TILE / 32 + 4];
    const i64 m = blockIdx.y;
    const int gpt = SUB_TILE / factor;
    const i64 g0 = (i64)blockIdx.x * gpt;
    const i64 left = n_out - g0;
    const int ng = left < gpt ? (int)left : gpt;
    sub_stage<T>(tile, in + m * stride + g0 * factor, ng * factor);
    __syncthreads();
    const int k_hi = factor >> 1, k_lo = (factor & 1) ? k_hi : k_hi - 1;
    const T inf = std::numeric_limits<T>::infinity();
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
        T a[P];
        bool bad = false;
        const int base = g * factor;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            a[i] = inf;
            if (i < factor) {
                const T v = tile[sub_slot(base + i)];
                bad |= (v != v);
                a[i] = v;
            }
        }
#pragma unroll
        for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int l = i ^ j;
                    if (l > i) {
                        const T x = a[i], y = a[l];
                        const T lo = x < y ? x : y, hi = x < y ? y : x;
                        if ((i & k) == 0) { a[i] = lo; a[l] = hi; } else { a[i] = hi; a[l] = lo; }
                    }
                }
            }
        }
        T lo = a[0], hi = a[0];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            if (i == k_lo) lo = a[i];
            if (i == k_hi) hi = a[i];
        }
        out[m * out_stride + g0 + g] = bad ? quiet_nan<T>() : ((factor & 1) ? hi : mid_of(lo, hi));
    }
}

// ---------------------------------------------------------------- register path: factor = V * 2^k (or a divisor of V)
// V = samples per 128-bit load.  A thread reduces its own vector, 2^k <= 32 neighbouring lanes finish the group with
// xor-shuffles: no shared memory, four independent 128-bit loads in flight per thread.  "average" / "max" / "min" only.
template <typename T, int METHOD> QI_DEV T sub_combine(T r, T v) {
    if (METHOD == QI_SUB_MAX) return (v > r || v != v) ? v : r;      // NaN-propagating: once r is NaN it stays
    return (v < r || v != v) ? v : r;
}
#ifndef QI_EMUL
// float: the hardware's NaN-propagating max / min (numpy's semantics) in one instruction
template <> QI_DEV float sub_combine<float, QI_SUB_MAX>(float r, float v) {
    float o;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(o) : "f"(r), "f"(v));
    return o;
}
template <> QI_DEV float sub_combine<float, QI_SUB_MIN>(float r, float v) {
    float o;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(o) : "f"(r), "f"(v));
    return o;
}
#endif
template <typename T, int METHOD>
__global__ void __launch_bounds__(256, 5)
subsample_vec_kernel(const T* __restrict__ in, i64 stride, i64 n_vec, int factor, T* __restrict__ out, i64 n_out) {
    constexpr int V = 16 / (int)sizeof(T), U = 4;
    typedef double Acc;
    const i64 m = blockIdx.y;
    const T* row = in + m * stride;
    T* orow = out + m * n_out;
    const int lanes = factor >= V ? factor / V : 1;                   // lanes that share a group
    const int per = factor >= V ? V : factor;                         // samples of one group inside a vector (fp32, factor 2: 2)
    const Acc inv = (Acc)1 / (Acc)factor;                             // factor is a power of two here: s * inv == s / factor exactly
    for (i64 base = (i64)blockIdx.x * (256 * U); base < n_vec; base += (i64)gridDim.x * (256 * U)) {
        T v[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 idx = base + u * 256 + threadIdx.x;
            if (idx < n_vec) {
                if (sizeof(T) == 4) {
                    const float4 w = *reinterpret_cast<const float4*>(row + idx * V);
                    v[u][0] = (T)w.x; v[u][1] = (T)w.y; v[u][V - 2] = (T)w.z; v[u][V - 1] = (T)w.w;
                } else {
                    const double2 w = *reinterpret_cast<const double2*>(row + idx * V);
                    v[u][0] = (T)w.x; v[u][V - 1] = (T)w.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) v[u][e] = (T)0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 idx = base + u * 256 + threadIdx.x;
            const bool live = idx < n_vec;
            if (per == V) {                                           // whole vector belongs to one group
                if (METHOD == QI_SUB_AVERAGE) {
                    // the samples of one 128-bit load are added in the record's own dtype (fp32: three additions, as
                    // numpy's float32 mean would), every level above that in fp64
                    const Acc s0 = (Acc)((v[u][0] + v[u][1]) + (v[u][V - 2] + v[u][V - 1]));
                    Acc s = (sizeof(T) == 4) ? s0 : (Acc)v[u][0] + (Acc)v[u][V - 1];
                    for (int o = lanes >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (live && (threadIdx.x & (lanes - 1)) == 0) orow[idx / lanes] = (T)(s * inv);
                } else {
                    T r = v[u][0];
#pragma unroll
                    for (int e = 1; e < V; ++e) r = sub_combine<T, METHOD>(r, v[u][e]);
                    for (int o = lanes >> 1; o > 0; o >>= 1) r = sub_combine<T, METHOD>(r, __shfl_xor_sync(0xffffffffu, r, o));
                    if (live && (threadIdx.x & (lanes - 1)) == 0) orow[idx / lanes] = r;
                }
            } else if (live) {                                        // factor 2 with four samples per vector
                T r0, r1;
                if (METHOD == QI_SUB_AVERAGE) {
                    r0 = (T)(((Acc)v[u][0] + (Acc)v[u][1]) / (Acc)2);
                    r1 = (T)(((Acc)v[u][V - 2] + (Acc)v[u][V - 1]) / (Acc)2);
                } else {
                    r0 = sub_combine<T, METHOD>(v[u][0], v[u][1]);
                    r1 = sub_combine<T, METHOD>(v[u][V - 2], v[u][V - 1]);
                }
                orow[2 * idx] = r0;
                orow[2 * idx + 1] = r1;
            }
        }
    }
}

// ---------------------------------------------------------------- long groups: one warp per group
// k-th smallest (0-based) of src[0..count) by a most-significant-byte-first radix select; hist: 256 ints of this warp.
template <typename T>
QI_DEV T warp_radix_select(const T* __restrict__ src, i64 count, i64 k, int* hist, int lane) {
    typedef typename key_traits<T>::type K;
    K prefix = 0, mask = 0;
    for (int shift = key_traits<T>::bits - 8; shift >= 0; shift -= 8) {
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (i64 i = lane; i < count; i += 32) {
            const K key = key_of(src[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 0xFF)], 1);
        }
        __syncwarp();
        int c[8], s = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { c[e] = hist[8 * lane + e]; s += c[e]; }
        int incl = s;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_sync(0xffffffffu, incl, lane >= o ? lane - o : lane);
            if (lane >= o) incl += t;
        }
        i64 below = incl - s;                                      // members of this pass in lower bins than mine
        int bin = -1;
        i64 under = 0;
        if (below <= k && k < incl) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (bin < 0 && k < below + c[e]) { bin = 8 * lane + e; under = below; }
                below += c[e];
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, bin, o);
            const i64 ou = __shfl_xor_sync(0xffffffffu, under, o);
            bin = ob > bin ? ob : bin;
            under += ou;
        }
        prefix |= (K)bin << shift;
        mask |= (K)0xFF << shift;
        k -= under;
        __syncwarp();
    }
    return value_of(prefix);
}

// grid: (blocks, M); warps stride over the groups of a row.
template <typename T, int METHOD>
__global__ void __launch_bounds__(256)
subsample_large_kernel(const T* __restrict__ in, i64 stride, i64 n_out, i64 factor, T* __restrict__ out, i64 out_stride) {
    __shared__ int hist[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 m = blockIdx.y;
    for (i64 g = (i64)blockIdx.x * 8 + warp; g < n_out; g += (i64)gridDim.x * 8) {
        const T* src = in + m * stride + g * factor;
        T res;
        if (METHOD == QI_SUB_AVERAGE) {
            double s = 0.0;
            for (i64 i = lane; i < factor; i += 32) s += (double)src[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            res = (T)(s / (double)factor);
        } else {
            T r = src[0];
            int bad = 0;
            for (i64 i = lane; i < factor; i += 32) {
                const T v = src[i];
                bad |= (v != v);
                if (METHOD == QI_SUB_MAX) r = v > r ? v : r;
                if (METHOD == QI_SUB_MIN) r = v < r ? v : r;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const T v = __shfl_xor_sync(0xffffffffu, r, o);
                bad |= __shfl_xor_sync(0xffffffffu, bad, o);
                if (METHOD == QI_SUB_MAX) r = v > r ? v : r;
                if (METHOD == QI_SUB_MIN) r = v < r ? v : r;
            }
            if (METHOD == QI_SUB_MEDIAN && !bad) {
                const T hi = warp_radix_select<T>(src, factor, factor >> 1, hist[warp], lane);
                r = (factor & 1) ? hi : mid_of(warp_radix_select<T>(src, factor, (factor >> 1) - 1, hist[warp], lane), hi);
            }
            res = bad ? quiet_nan<T>() : r;
        }
        if (lane == 0) out[m * out_stride + g] = res;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
subsample_nth_kernel(const T* __restrict__ in, i64 stride, i64 n_out, i64 factor, T* __restrict__ out, i64 out_stride) {
    const i64 m = blockIdx.y;
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_out) out[m * out_stride + j] = in[m * stride + j * factor];
}

template <typename T, int METHOD>
static void subsample_launch(const T* in, i64 M, i64 stride, i64 factor, T* out, i64 n_out, cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(T);
    const bool pow2 = (factor & (factor - 1)) == 0 && factor <= 32 * V && (factor >= V || factor == 2);
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && stride % V == 0;
    if (METHOD != QI_SUB_MEDIAN && pow2 && aligned && (n_out * factor) % V == 0) {
        const i64 n_vec = n_out * factor / V;
        i64 blocks = (n_vec + 1023) / 1024;
        const i64 cap = (148 * 8 * 8 + M - 1) / M;
        if (blocks > cap) blocks = cap;
        dim3 grid((unsigned)blocks, (unsigned)M);
        QI_LAUNCH((subsample_vec_kernel<T, METHOD>), grid, dim3(256), 0, st, in, stride, n_vec, (int)factor, out, n_out);
    } else if (METHOD == QI_SUB_MEDIAN && factor <= 32) {
        const int f = (int)factor;
        const int gpt = SUB_TILE / f;
        dim3 grid((unsigned)((n_out + gpt - 1) / gpt), (unsigned)M);
        const dim3 block(gpt >= 256 ? 256 : 128);
        if (f <= 4) QI_LAUNCH((subsample_median_sort_kernel<T, 4>), grid, block, 0, st, in, stride, n_out, f, out, n_out);
        else if (f <= 8) QI_LAUNCH((subsample_median_sort_kernel<T, 8>), grid, block, 0, st, in, stride, n_out, f, out, n_out);
        else if (f <= 16) QI_LAUNCH((subsample_median_sort_kernel<T, 16>), grid, block, 0, st, in, stride, n_out, f, out, n_out);
        else QI_LAUNCH((subsample_median_sort_kernel<T, 32>), grid, block, 0, st, in, stride, n_out, f, out, n_out);
    } else if (factor <= SUB_SMALL_MAX) {
        const int f = (int)factor;
        int lanes = 1;
        while (lanes < 32 && lanes * 8 <= f) lanes <<= 1;          // >= 8 samples per lane
        const int gpt = SUB_TILE / f;
        dim3 grid((unsigned)((n_out + gpt - 1) / gpt), (unsigned)M);
        QI_LAUNCH((subsample_small_kernel<T, METHOD>), grid, dim3(256), 0, st, in, stride, n_out, f, lanes, out, n_out);
    } else {
        i64 blocks = (n_out + 7) / 8;
        const i64 cap = (148 * 8 * 4 + M - 1) / M;                  // a few waves of resident CTAs over all rows
        if (blocks > cap) blocks = cap;
        dim3 grid((unsigned)blocks, (unsigned)M);
        QI_LAUNCH((subsample_large_kernel<T, METHOD>), grid, dim3(256), 0, st, in, stride, n_out, factor, out, n_out);
    }
}

template <typename T>
static int subsample_impl(const void* in_, i64 M, i64 n_in, i64 stride, i64 factor, int method, void* out_, i64 n_out,
                          cudaStream_t st) {
    const T* in = static_cast<const T*>(in_);
    T* out = static_cast<T*>(out_);
    const i64 expect = method == QI_SUB_NTH ? (n_in + factor - 1) / factor : n_in / factor;
    if (n_out != expect) return QI_ERR_ARG;
    if (n_out == 0) return QI_OK;
    prof_set_category(QI_CAT_INFO);
    switch (method) {
        case QI_SUB_NTH: {
            dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)M);
            QI_LAUNCH((subsample_nth_kernel<T>), grid, dim3(256), 0, st, in, stride, n_out, factor, out, n_out);
            break;
        }
        case QI_SUB_AVERAGE: subsample_launch<T, QI_SUB_AVERAGE>(in, M, stride, factor, out, n_out, st); break;
        case QI_SUB_MEDIAN: subsample_launch<T, QI_SUB_MEDIAN>(in, M, stride, factor, out, n_out, st); break;
        case QI_SUB_MAX: subsample_launch<T, QI_SUB_MAX>(in, M, stride, factor, out, n_out, st); break;
        case QI_SUB_MIN: subsample_launch<T, QI_SUB_MIN>(in, M, stride, factor, out, n_out, st); break;
        default: return QI_ERR_ARG;
    }
    return check_cuda("qi_subsample");
}

// ---------------------------------------------------------------- extrema of a record (NaN-ignoring) + NaN count
// acc[0] = key(max), acc[1] = key(min), acc[2] = key(max |x|), acc[3] = number of NaNs; decoded in place afterwards.
__global__ void extrema_init_kernel(unsigned long long* acc, i64 M) {
    const i64 m = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) { acc[4 * m] = 0ull; acc[4 * m + 1] = ~0ull; acc[4 * m + 2] = 0ull; acc[4 * m + 3] = 0ull; }
}
__global__ void extrema_decode_kernel(unsigned long long* acc, i64 M) {
    const i64 m = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) {
        double* o = reinterpret_cast<double*>(acc + 4 * m);
        const unsigned long long k0 = acc[4 * m], k1 = acc[4 * m + 1], k2 = acc[4 * m + 2], c = acc[4 * m + 3];
        o[0] = value_of(k0);                                        // all-NaN rows decode to NaN (numpy: nanmax -> nan)
        o[1] = value_of(k1);
        o[2] = value_of(k2);
        o[3] = (double)c;
    }
}
template <typename T>
__global__ void __launch_bounds__(256)
extrema_kernel(const T* __restrict__ x, i64 n, i64 stride, unsigned long long* __restrict__ acc) {
    const i64 m = blockIdx.y;
    const T* row = x + m * stride;
    unsigned long long kmax = 0ull, kmin = ~0ull, kabs = 0ull, nans = 0ull;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = (double)row[i];
        if (v != v) { ++nans; continue; }
        const unsigned long long k = key_of(v), ka = key_of(v < 0.0 ? -v : v);
        kmax = k > kmax ? k : kmax;
        kmin = k < kmin ? k : kmin;
        kabs = ka > kabs ? ka : kabs;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmax, o), b = __shfl_xor_sync(0xffffffffu, kmin, o),
                                 c = __shfl_xor_sync(0xffffffffu, kabs, o), d = __shfl_xor_sync(0xffffffffu, nans, o);
        kmax = a > kmax ? a : kmax;
        kmin = b < kmin ? b : kmin;
        kabs = c > kabs ? c : kabs;
        nans += d;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&acc[4 * m], kmax);
        atomicMin(&acc[4 * m + 1], kmin);
        atomicMax(&acc[4 * m + 2], kabs);
        if (nans) atomicAdd(&acc[4 * m + 3], nans);
    }
}

// ---------------------------------------------------------------- local maxima (scipy _local_maxima_1d + height)
// Sample i starts a candidate when x[i-1] < x[i]; the plateau x[i..j) is a peak when x[j] < x[i]; its position is the
// plateau midpoint (i + j - 1) / 2.  Peaks are appended unordered (the host sorts the short list).
template <typename T>
__global__ void __launch_bounds__(256)
local_maxima_kernel(const T* __restrict__ x, i64 n, double height, int use_height, i64* __restrict__ peaks,
                    double* __restrict__ values, i64 capacity, unsigned long long* __restrict__ count) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= n - 1) return;
    const T v = x[i];
    if (!(x[i - 1] < v)) return;
    i64 j = i + 1;
    while (j < n && x[j] == v) ++j;
    if (j >= n || !(x[j] < v)) return;
    if (use_height && !((double)v >= height)) return;
    const unsigned long long slot = atomicAdd(count, 1ull);
    if ((i64)slot < capacity) {
        peaks[slot] = (i + j - 1) / 2;
        if (values) values[slot] = (double)v;
    }
}

// out = x / divisor in the record's own dtype (numpy: array / scalar of the same dtype)
template <typename T>
__global__ void __launch_bounds__(256)
divide_kernel(const T* __restrict__ x, i64 n, T divisor, T* __restrict__ out) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] / divisor;
}

}  // namespace qi

extern "C" {

int qi_subsample(const void* in, int64_t M, int64_t n_in, int64_t stride, int64_t factor, int method, int dtype,
                 void* out, int64_t n_out, void* stream) {
    if (!in || !out || M <= 0 || M > 65535 || n_in <= 0 || stride < n_in || factor < 1 || factor > (int64_t)0x7fffffff)
        return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32) return qi::subsample_impl<float>(in, M, n_in, stride, factor, method, out, n_out, st);
    if (dtype == QI_F64) return qi::subsample_impl<double>(in, M, n_in, stride, factor, method, out, n_out, st);
    return QI_ERR_ARG;
}

int qi_extrema(const void* in, int64_t M, int64_t n, int64_t stride, int dtype, double* out, void* stream) {
    if (!in || !out || M <= 0 || M > 65535 || n <= 0 || stride < n) return QI_ERR_ARG;
    if (dtype != QI_F32 && dtype != QI_F64) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(out);
    qi::prof_set_category(QI_CAT_INFO);
    QI_LAUNCH((qi::extrema_init_kernel), dim3((unsigned)((M + 255) / 256)), dim3(256), 0, st, acc, (qi::i64)M);
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (148 * 8 + M - 1) / M;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)M);
    if (dtype == QI_F32)
        QI_LAUNCH((qi::extrema_kernel<float>), grid, dim3(256), 0, st, static_cast<const float*>(in), (qi::i64)n, (qi::i64)stride, acc);
    else
        QI_LAUNCH((qi::extrema_kernel<double>), grid, dim3(256), 0, st, static_cast<const double*>(in), (qi::i64)n, (qi::i64)stride, acc);
    QI_LAUNCH((qi::extrema_decode_kernel), dim3((unsigned)((M + 255) / 256)), dim3(256), 0, st, acc, (qi::i64)M);
    return qi::check_cuda("qi_extrema");
}

int qi_select_peaks_by_distance(const int64_t* peaks, const int64_t* order, int64_t n, int64_t distance, uint8_t* keep) {
    if (n < 0 || distance < 1 || (n > 0 && (!peaks || !order || !keep))) return QI_ERR_ARG;
    for (int64_t i = 0; i < n; ++i) keep[i] = 1;
    for (int64_t i = n - 1; i >= 0; --i) {
        const int64_t j = order[i];
        if (j < 0 || j >= n) return QI_ERR_ARG;
        if (!keep[j]) continue;
        for (int64_t k = j - 1; k >= 0 && peaks[j] - peaks[k] < distance; --k) keep[k] = 0;
        for (int64_t k = j + 1; k < n && peaks[k] - peaks[j] < distance; ++k) keep[k] = 0;
    }
    return QI_OK;
}

int qi_divide(const void* in, int64_t n, int dtype, double divisor, void* out, void* stream) {
    if (!in || !out || n <= 0) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((n + 255) / 256));
    qi::prof_set_category(QI_CAT_INFO);
    if (dtype == QI_F32)
        QI_LAUNCH((qi::divide_kernel<float>), grid, dim3(256), 0, st, static_cast<const float*>(in), (qi::i64)n, (float)divisor, static_cast<float*>(out));
    else if (dtype == QI_F64)
        QI_LAUNCH((qi::divide_kernel<double>), grid, dim3(256), 0, st, static_cast<const double*>(in), (qi::i64)n, divisor, static_cast<double*>(out));
    else return QI_ERR_ARG;
    return qi::check_cuda("qi_divide");
}

int qi_local_maxima(const void* in, int64_t n, int dtype, double height, int use_height, int64_t* peaks,
                    double* values, int64_t capacity, int64_t* count, void* stream) {
    if (!in || !count || n < 0 || capacity < 0 || (capacity > 0 && !peaks)) return QI_ERR_ARG;
    if (dtype != QI_F32 && dtype != QI_F64) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(count, 0, sizeof(int64_t), st);
    if (n >= 3) {
        qi::prof_set_category(QI_CAT_INFO);
        dim3 grid((unsigned)((n - 2 + 255) / 256));
        unsigned long long* cnt = reinterpret_cast<unsigned long long*>(count);
        if (dtype == QI_F32)
            QI_LAUNCH((qi::local_maxima_kernel<float>), grid, dim3(256), 0, st, static_cast<const float*>(in), (qi::i64)n, height, use_height, reinterpret_cast<qi::i64*>(peaks), values, (qi::i64)capacity, cnt);
        else
            QI_LAUNCH((qi::local_maxima_kernel<double>), grid, dim3(256), 0, st, static_cast<const double*>(in), (qi::i64)n, height, use_height, reinterpret_cast<qi::i64*>(peaks), values, (qi::i64)capacity, cnt);
    }
    return qi::check_cuda("qi_local_maxima");
}

}  // extern "C"
