// qi_interp.cuh -- the Kaiser-windowed-sinc interpolator of the band-limited routes (Gabor CWT: qi_cwt_fast.cuh,
// Stockwell: qi_stx.cu), float32 and float64.
//
// A band whose spectrum is negligible outside K = 2^m bins around its centre is known exactly from the K-point inverse
// transform of those bins: every D = L / K-th sample of its L-point (circular) output, demodulated by the centre bin.  A
// TAPS-per-phase interpolator brings it to the full rate, puts the carrier back (exact integer phase, stepped inside a
// thread) and fuses the output slice, |.|^2 and the fp64 band sums.
//     float32: RHO = 2 (oversampling of the decimated samples), 16 taps, beta 11.2 (stop band -110 dB)
//     float64: RHO = 4, 24 taps, beta 24.5 (-231 dB)
#pragma once
#include "qi_tfr.cuh"

namespace qi {

constexpr int CWTF_MAX_LOGD = 8;
constexpr int CWTF_TILE = 2048;
constexpr int CWTF_SPAN = 4;                 // tiles per CTA of the interpolator
constexpr int CWTF_OS_LOGF = 11;             // overlap-save block length 2048
constexpr int CWTF_OS_MAX_HALF = 384;        // longest kernel half-support taken by the overlap-save route

template <typename T> struct CwtFastCfg;
template <> struct CwtFastCfg<float> {
    static constexpr int TAPS = 16, RHO = 2, PER = 8;
    static constexpr double U_CUT = 5.6, BETA = 11.2;
};
template <> struct CwtFastCfg<double> {
    static constexpr int TAPS = 24, RHO = 4, PER = 4;
    static constexpr double U_CUT = 7.7, BETA = 24.5;
};

template <typename T> struct CwtFastAcc { typedef double type; };
template <> struct CwtFastAcc<float> { typedef float type; };

// modified Bessel function I0 by its power series (converged to 1e-19 of the sum for x <= 23 after 64 terms)
QI_HD double cwtf_bessel_i0(double x) {
    double s = 1.0, term = 1.0;
    const double hh = 0.25 * x * x;
    for (int k = 1; k <= 64; ++k) { term *= hh / (double)(k * k); s += term; }
    return s;
}

// coef[logD][j * D + p] = h(p - (j - (TAPS/2 - 1)) D),  h(t) = sinc(t / D) kaiser(t / (TAPS/2 D); beta)
template <typename T>
__global__ void cwtf_coef_kernel(T* __restrict__ coef_all, unsigned need_mask, double inv_i0_beta) {
    constexpr int TAPS = CwtFastCfg<T>::TAPS;
    const int logD = blockIdx.y;
    if (!((need_mask >> logD) & 1u)) return;
    T* coef = coef_all + (size_t)logD * TAPS * (1u << CWTF_MAX_LOGD);
    const int D = 1 << logD;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= TAPS * D) return;
    const int j = idx >> logD, p = idx & (D - 1);
    const double t = (double)(p - (j - (TAPS / 2 - 1)) * D);
    const double x = t / (0.5 * TAPS * D);
    const double arg = 1.0 - x * x;
    const double w = cwtf_bessel_i0(CwtFastCfg<T>::BETA * sqrt(arg > 0.0 ? arg : 0.0)) * inv_i0_beta;
    const double y = t / (double)D;
    const double sinc = t == 0.0 ? 1.0 : sinpi(y) / (M_PI * y);
    coef[idx] = (T)(sinc * w);
}

// grid: (ceil(N / (SPAN * TILE)), bands of the group, channels);  dec: [band_in_group][chan][K]
// float64 keeps the real and the imaginary parts of the decimated samples in separate shared arrays and runs the taps over
// one part at a time: 27 + 24 doubles live instead of 54 + 24, i.e. ~120 instead of 198 registers and two CTAs per SM
// instead of one (ncu before: occupancy 12 %, issue-active 45 %, no memory stall -- the FP64 pipe waiting for issue slots).
template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 2 : 1)
cwtf_interp_kernel(const cplx<T>* __restrict__ dec, const int* __restrict__ ids, const long long* __restrict__ kc_of_band,
                   int kc_stride, CwtGeom geo, int logD, const T* __restrict__ coef, cplx<T>* __restrict__ out_c, T* __restrict__ out_p,
                   double* __restrict__ band_sum) {
    constexpr int TAPS = CwtFastCfg<T>::TAPS, PER = CwtFastCfg<T>::PER, NSUB = 8 / PER, J0 = TAPS / 2 - 1;
    constexpr bool SPLIT = sizeof(T) == 8;
    // one pad slot per 8 decimated samples: at small D the lanes of a warp start their windows PER samples apart
    constexpr int SEGN = (CWTF_SPAN * CWTF_TILE / 4 + TAPS) * 9 / 8 + 2;
    __shared__ cplx<T> seg[SEGN];
    T* seg_re = reinterpret_cast<T*>(seg);                  // SPLIT: [SEGN] real parts, then [SEGN] imaginary parts
    T* seg_im = seg_re + SEGN;
    __shared__ double scratch[32];
    const int D = 1 << logD, logK = geo.logL - logD;
    const i64 K = 1ll << logK, N = geo.n_points;
    const i64 span0 = (i64)blockIdx.x * (CWTF_SPAN * CWTF_TILE);
    const i64 left = (N - span0 + CWTF_TILE - 1) / CWTF_TILE;
    const int ntile = left < CWTF_SPAN ? (int)left : CWTF_SPAN;
    const int bi = blockIdx.y, chan = blockIdx.z, band = ids[bi];
    // carrier bin of the band (0 without a table: the decimated samples are the output's own baseband)
    const unsigned long long kc = kc_of_band ? (unsigned long long)kc_of_band[(size_t)band * kc_stride] : 0ull;
    const cplx<T>* src = dec + (((i64)bi * geo.n_channels + chan) << logK);
    const i64 m_base = (span0 >> logD) - J0;
    const int nseg = ((ntile * CWTF_TILE) >> logD) + TAPS;
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) {
        const cplx<T> v = src[(m_base + i) & (K - 1)];
        if (SPLIT) { seg_re[i + (i >> 3)] = v.re; seg_im[i + (i >> 3)] = v.im; }
        else seg[i + (i >> 3)] = v;
    }
    __syncthreads();
    const i64 row = ((i64)chan * geo.n_bands + band) * N;
    const int p = threadIdx.x & (D - 1);
    const int MT = CWTF_TILE >> logD;                       // decimated samples per tile
    T cf[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) cf[j] = coef[(j << logD) + p];
    // carrier exp(2 pi i k_c n / L) at n = m D + p: exact at the first sample of a run, stepped by exp(2 pi i k_c D / L)
    const unsigned long long Lmask = (1ull << geo.logL) - 1ull;
    const cplx<T> step = unit_root<T>((kc << logD) & Lmask, geo.logL);
    // float32: partial sums of a thread's <= 64 non-negative outputs in the arithmetic type, fp64 across the CTA
    typename CwtFastAcc<T>::type acc = 0;
    for (int tl = 0; tl < ntile; ++tl) {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            const int m0 = tl * MT + sub * (MT / NSUB) + (threadIdx.x >> logD) * PER;
            T re[PER], im[PER];
            if (SPLIT) {
                // m0 is a multiple of PER = 4: the window starts at slot m0 + (m0 >> 3) and crosses a padding slot where
                // (m0 & 7) + j reaches a multiple of 8
                const int s0 = m0 + (m0 >> 3), r0 = m0 & 7;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const T* sp = (part ? seg_im : seg_re) + s0;
                    T win[PER + TAPS - 1];
#pragma unroll
                    for (int j = 0; j < PER + TAPS - 1; ++j) win[j] = sp[j + ((r0 + j) >> 3)];
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        T a = (T)0;
#pragma unroll
                        for (int j = 0; j < TAPS; ++j) a += cf[j] * win[i + j];
                        if (part) im[i] = a; else re[i] = a;
                    }
                }
            } else {
                // float32: one packed fma.rn.f32x2 per tap on the (re, im) pair (the kernel is issue-bound: ncu 76 %
                // issue-active, 87 thread instructions per output cell with scalar taps)
                const f32x2* seg2 = reinterpret_cast<const f32x2*>(seg);
                f32x2 win[PER + TAPS - 1];
#pragma unroll
                for (int j = 0; j < PER + TAPS - 1; ++j) win[j] = seg2[m0 + j + ((m0 + j) >> 3)];
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    f32x2 a = f2_make(0.0f, 0.0f);
#pragma unroll
                    for (int j = 0; j < TAPS; ++j) a = f2_fma(f2_make((float)cf[j], (float)cf[j]), win[i + j], a);
                    re[i] = (T)f2_lo(a); im[i] = (T)f2_hi(a);
                }
            }
            const i64 n0 = span0 + ((i64)m0 << logD) + p;
            cplx<T> car = mk<T>((T)1, (T)0);
            if (out_c) car = unit_root<T>((kc * (unsigned long long)n0) & Lmask, geo.logL);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const i64 n = n0 + ((i64)i << logD);
                if (n < N) {
                    const T pw = re[i] * re[i] + im[i] * im[i];
                    if (out_c) out_c[row + n] = mk<T>(re[i], im[i]) * car;
                    if (out_p) out_p[row + n] = pw;
                    acc += pw;
                }
                if (out_c) car = car * step;
            }
        }
    }
    if (band_sum) {
        const double tot = block_sum((double)acc, scratch);
        if (threadIdx.x == 0) atomicAdd(&band_sum[(i64)chan * geo.n_bands + band], tot);
    }
}

}  // namespace qi
