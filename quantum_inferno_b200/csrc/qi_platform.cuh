// qi_platform.cuh -- build-mode switch and small device helpers.
//
// Product build : nvcc -gencode arch=compute_100a,code=sm_100a  (the only build the package loads)
// Debug build   : g++ -DQI_EMUL (tests/emul only; CPU emulation of the kernel logic, never shipped)
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef QI_EMUL
#include "cuda_emul.h"
#define QI_LAUNCH(kern, grid, block, smem, stream, ...) \
    qi_emul::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
namespace qi { inline void prof_set_category(int) {} }
#define QI_DYN_SMEM(name) unsigned char* name = QI_EMUL_DYN_SMEM
#define QI_HD inline
#define QI_DEV inline
#else
#include <cuda_runtime.h>
namespace qi {
// launch accounting + optional per-category CUDA-event timing (qi_profile_* in include/qi_b200.h)
void prof_set_category(int cat);
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);
}
#define QI_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
    do {                                                                                   \
        auto qi_kfn_ = kern;                                                               \
        qi::prof_begin(stream);                                                            \
        qi_kfn_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                       \
        qi::prof_end(stream);                                                              \
    } while (0)
#define QI_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define QI_HD __host__ __device__ __forceinline__
#define QI_DEV __device__ __forceinline__
#endif

#include "../../include/qi_b200.h"

namespace qi {

typedef long long i64;

// ---------------------------------------------------------------- complex
template <typename T> struct __align__(8) cplx_base { T re, im; };
template <typename T> struct cplx;
template <> struct __align__(8) cplx<float> { float re, im; };
template <> struct __align__(16) cplx<double> { double re, im; };

template <typename T> QI_HD cplx<T> mk(T a, T b) { cplx<T> r; r.re = a; r.im = b; return r; }
template <typename T> QI_HD cplx<T> operator+(cplx<T> a, cplx<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> QI_HD cplx<T> operator-(cplx<T> a, cplx<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <typename T> QI_HD cplx<T> operator*(cplx<T> a, cplx<T> b) {
    return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> QI_HD cplx<T> operator*(cplx<T> a, T s) { return mk<T>(a.re * s, a.im * s); }
template <typename T> QI_HD cplx<T> conj(cplx<T> a) { return mk<T>(a.re, -a.im); }
// a * conj(b)
template <typename T> QI_HD cplx<T> mul_conj(cplx<T> a, cplx<T> b) {
    return mk<T>(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
// multiply by -i (forward quarter turn) or +i
template <typename T> QI_HD cplx<T> mul_mi(cplx<T> a) { return mk<T>(a.im, -a.re); }
template <typename T> QI_HD cplx<T> mul_pi(cplx<T> a) { return mk<T>(-a.im, a.re); }
template <typename T> QI_HD T norm2(cplx<T> a) { return a.re * a.re + a.im * a.im; }

// ---------------------------------------------------------------- packed pair of floats
// One 64-bit register pair on the device: the operand of the f32x2 instructions of sm_100 (fma / add / sub / mul on both
// halves).  The CPU emulation build spells the same operations out.
#ifdef QI_EMUL
struct f32x2 { float x, y; };
QI_DEV f32x2 f2_make(float x, float y) { f32x2 r; r.x = x; r.y = y; return r; }
QI_DEV float f2_lo(f32x2 v) { return v.x; }
QI_DEV float f2_hi(f32x2 v) { return v.y; }
QI_DEV f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { return f2_make(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
QI_DEV f32x2 f2_add(f32x2 a, f32x2 b) { return f2_make(a.x + b.x, a.y + b.y); }
#else
typedef unsigned long long f32x2;
QI_DEV f32x2 f2_make(float x, float y) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
QI_DEV float f2_lo(f32x2 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return x; }
QI_DEV float f2_hi(f32x2 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return y; }
QI_DEV f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
QI_DEV f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#endif

// ---------------------------------------------------------------- exact twiddles
// exp(sign * 2*pi*i * m / 2^lb), 0 <= m < 2^lb, evaluated with an exact quadrant
// reduction so the argument handed to sincospi is exactly representable.
QI_DEV void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
QI_DEV void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }

template <typename T> QI_DEV cplx<T> unit_root_small(unsigned long long m, int lb) {
    // requires lb <= 2 or (m mod 2^(lb-2)) exactly representable in T
    if (lb == 0) return mk<T>((T)1, (T)0);
    if (lb == 1) return mk<T>(m ? (T)-1 : (T)1, (T)0);
    unsigned q = (unsigned)(m >> (lb - 2)) & 3u;
    unsigned long long r = m & ((1ull << (lb - 2)) - 1ull);
    // angle / pi = 2*r / 2^lb = r * 2^-(lb-1)
    T frac = (T)r;
    // scale by power of two exactly
    frac = frac * (T)(1.0 / (double)(1ull << (lb - 1)));
    T s, c;
    sincospi_t(frac, &s, &c);
    // rotate by q quarter turns: (c + i s) * i^q
    T cr, ci;
    if (q == 0) { cr = c; ci = s; }
    else if (q == 1) { cr = -s; ci = c; }
    else if (q == 2) { cr = -c; ci = -s; }
    else { cr = s; ci = -c; }
    return mk<T>(cr, ci);
}

// exp(+2*pi*i*m/2^lb)
template <typename T> QI_DEV cplx<T> unit_root(unsigned long long m, int lb);
template <> QI_DEV cplx<double> unit_root<double>(unsigned long long m, int lb) {
    return unit_root_small<double>(m, lb);   // r < 2^51 always exact for lb <= 53
}
template <> QI_DEV cplx<float> unit_root<float>(unsigned long long m, int lb) {
    if (lb <= 26) return unit_root_small<float>(m, lb);
    // split m = hi * 2^13 + lo  ->  w^m = (w^(2^13))^hi * w^lo
    unsigned long long lo = m & 8191ull, hi = m >> 13;
    cplx<float> a = unit_root_small<float>(hi, lb - 13);
    cplx<float> b = unit_root_small<float>(lo, lb);   // lo < 2^13: exact
    return a * b;
}

QI_HD unsigned brev_bits(unsigned v, int bits) {
#if defined(__CUDA_ARCH__) || defined(QI_EMUL)
    return bits == 0 ? 0u : (__brev(v) >> (32 - bits));
#else
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

template <typename T> struct real_traits;
template <> struct real_traits<float> { static constexpr int dtype = QI_F32; };
template <> struct real_traits<double> { static constexpr int dtype = QI_F64; };

}  // namespace qi
