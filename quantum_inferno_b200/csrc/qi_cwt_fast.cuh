// qi_cwt_fast.cuh -- band-limited routes of the exact (FFT) Gabor CWT, float32 and float64.
//
// The plain route of qi_cwt.cu spends three full-length HBM passes (length L = 2N) on every band.  A Gaussian atom
// whose support has decayed inside the record does not need them:
//
//  (D) DECIMATED route, long atoms (s large).  The band response exp(-0.5 s^2 (theta - omega)^2) is below CUT of its
//      peak beyond |k - k_c| > kmax = U / (s 2 pi / L) bins.  The K = 2^m >= RHO (2 kmax + 4) bins around k_c, moved to
//      baseband, are inverse-transformed at length K (same pass kernels, K << L): that is the band output at every
//      D = L / K-th sample, demodulated by exp(-2 pi i k_c n / L), EXACTLY (the product is band-limited on the circular
//      L-grid).  A TAPS-per-phase Kaiser-windowed-sinc interpolator brings it to the full rate, the carrier is put back
//      (exact integer phase, stepped inside a thread), and the store fuses the 'same' slice, |.|^2 and the fp64 band sum.
//          float32: RHO = 2, 16 taps, beta 11.2  (stop band -110 dB);   float64: RHO = 4, 24 taps, beta 24.5 (-231 dB)
//      HBM traffic per cell: the output + 2 RHO-ish / D reads instead of ~6 x 2 x sizeof(complex).
//
//  (S) OVERLAP-SAVE route, short atoms (the top bands, which are too wide to decimate).  Their kernels are a few
//      hundred samples long (|t| <= U s): the record is cut into blocks of F = 2048 samples held in shared memory, one
//      forward transform per block, and per band one product with the band's response on the F-grid (the same closed
//      form, periodised at F -- exact because the kernel has died out well inside the block) and one inverse transform;
//      the V = F - 2 half valid samples go straight to the planes.  Nothing but the record and the output touches HBM.
//
// Record-long truncated atoms (table bands) and the circular-correlation mode keep the plain route.
// Replaces quantum_inferno/styx_cwt.py:195-196 (scipy fftconvolve on the tiled record) for those bands.
#pragma once

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

// First inverse pass of the K-point transform of a band's baseband bins.  batch = band_in_group * C + chan
template <typename T> struct SrcCwtDec {
    const cplx<T>* spec; const DevBand* bands; const int* ids; int n_channels, logL, logK, half_shift;
    QI_DEV cplx<T> load(i64 batch, i64 e) const {
        const i64 chan = batch % n_channels;
        const DevBand& b = bands[ids[batch / n_channels]];
        const i64 K = 1ll << logK, L = 1ll << logL;
        const i64 q = (i64)brev_bits((unsigned)e, logK);
        const i64 ks = (q < (K >> 1)) ? q : q - K;
        if (ks > b.dec_kmax || -ks > b.dec_kmax) return mk<T>((T)0, (T)0);
        const i64 k = (b.kc + ks) & (L - 1);
        const cplx<T> X = spec[(chan << logL) + (i64)brev_bits((unsigned)k, logL)];
        return X * gabor_response<T>(b, k, logL, half_shift);
    }
};

// grid: (ceil(N / (SPAN * TILE)), bands of the group, channels);  dec: [band_in_group][chan][K]
template <typename T>
__global__ void __launch_bounds__(256)
cwtf_interp_kernel(const cplx<T>* __restrict__ dec, const int* __restrict__ ids, const DevBand* __restrict__ bands,
                   CwtGeom geo, int logD, const T* __restrict__ coef, cplx<T>* __restrict__ out_c, T* __restrict__ out_p,
                   double* __restrict__ band_sum) {
    constexpr int TAPS = CwtFastCfg<T>::TAPS, PER = CwtFastCfg<T>::PER, NSUB = 8 / PER, J0 = TAPS / 2 - 1;
    // one pad slot per 8 decimated samples: at small D the lanes of a warp start their windows PER samples apart
    __shared__ cplx<T> seg[(CWTF_SPAN * CWTF_TILE / 4 + TAPS) * 9 / 8 + 2];
    __shared__ double scratch[32];
    const int D = 1 << logD, logK = geo.logL - logD;
    const i64 K = 1ll << logK, N = geo.n_points;
    const i64 span0 = (i64)blockIdx.x * (CWTF_SPAN * CWTF_TILE);
    const i64 left = (N - span0 + CWTF_TILE - 1) / CWTF_TILE;
    const int ntile = left < CWTF_SPAN ? (int)left : CWTF_SPAN;
    const int bi = blockIdx.y, chan = blockIdx.z, band = ids[bi];
    const DevBand b = bands[band];
    const cplx<T>* src = dec + (((i64)bi * geo.n_channels + chan) << logK);
    const i64 m_base = (span0 >> logD) - J0;
    const int nseg = ((ntile * CWTF_TILE) >> logD) + TAPS;
    for (int i = threadIdx.x; i < nseg; i += blockDim.x) seg[i + (i >> 3)] = src[(m_base + i) & (K - 1)];
    __syncthreads();
    const i64 row = ((i64)chan * geo.n_bands + band) * N;
    const int p = threadIdx.x & (D - 1);
    const int MT = CWTF_TILE >> logD;                       // decimated samples per tile
    T cf[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) cf[j] = coef[(j << logD) + p];
    // carrier exp(2 pi i k_c n / L) at n = m D + p: exact at the first sample of a run, stepped by exp(2 pi i k_c D / L)
    const unsigned long long Lmask = (1ull << geo.logL) - 1ull;
    const cplx<T> step = unit_root<T>(((unsigned long long)b.kc << logD) & Lmask, geo.logL);
    double acc = 0.0;
    for (int tl = 0; tl < ntile; ++tl) {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            const int m0 = tl * MT + sub * (MT / NSUB) + (threadIdx.x >> logD) * PER;
            cplx<T> win[PER + TAPS - 1];
#pragma unroll
            for (int j = 0; j < PER + TAPS - 1; ++j) win[j] = seg[m0 + j + ((m0 + j) >> 3)];
            const i64 n0 = span0 + ((i64)m0 << logD) + p;
            cplx<T> car = mk<T>((T)1, (T)0);
            if (out_c) car = unit_root<T>(((unsigned long long)b.kc * (unsigned long long)n0) & Lmask, geo.logL);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                T re = (T)0, im = (T)0;
#pragma unroll
                for (int j = 0; j < TAPS; ++j) { re += cf[j] * win[i + j].re; im += cf[j] * win[i + j].im; }
                const i64 n = n0 + ((i64)i << logD);
                if (n < N) {
                    const T pw = re * re + im * im;
                    if (out_c) out_c[row + n] = mk<T>(re, im) * car;
                    if (out_p) out_p[row + n] = pw;
                    acc += (double)pw;
                }
                if (out_c) car = car * step;
            }
        }
    }
    if (band_sum) {
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) atomicAdd(&band_sum[(i64)chan * geo.n_bands + band], acc);
    }
}

// ---------------------------------------------------------------- (S) overlap-save route
// response tables of the overlap-save bands on the F-grid, bit-reversed order: tabF[i][e]
template <typename T>
__global__ void cwtf_os_table_kernel(const DevBand* __restrict__ bandsF, int logF, int half_shift, cplx<T>* __restrict__ tabF) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (1 << logF)) return;
    const i64 k = (i64)brev_bits((unsigned)e, logF);
    tabF[((size_t)blockIdx.y << logF) + e] = gabor_response<T>(bandsF[blockIdx.y], k, logF, half_shift);
}

// grid: (ceil(N / V), channels); one block of F samples, all n_os bands.  dyn smem: (2 pad8(F) + F) complex + 256 B
template <typename T>
__global__ void __launch_bounds__(512)
cwtf_os_kernel(const T* __restrict__ sig, i64 stride, CwtGeom geo, const int* __restrict__ ids, int n_os, int logF,
               int half, const cplx<T>* __restrict__ tabF, cplx<T>* __restrict__ out_c, T* __restrict__ out_p,
               double* __restrict__ band_sum) {
    QI_DYN_SMEM(smem_raw);
    const int F = 1 << logF, V = F - 2 * half;
    const int FP = pad8(F);                                   // tiles in the padded single-column layout of tile_fft<.., true>
    cplx<T>* tile_x = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* tile_y = tile_x + FP;
    cplx<T>* tw = tile_y + FP;
    double* scratch = reinterpret_cast<double*>(tw + F);
    const i64 chan = blockIdx.y, N = geo.n_points;
    const i64 n0 = (i64)blockIdx.x * V;
    const T* xs = sig + chan * stride;
    fill_stage_twiddles<T>(tw, logF);
    for (int p = threadIdx.x; p < F; p += blockDim.x) {
        const i64 k = n0 - half + p;
        tile_x[pad8(p)] = mk<T>((k >= 0 && k < N) ? xs[k] : (T)0, (T)0);
    }
    __syncthreads();
    tile_fft<T, FFT_FWD, true>(tile_x, tw, logF, 1, 1);
    for (int i = 0; i < n_os; ++i) {
        const int band = ids[i];
        const cplx<T>* H = tabF + ((size_t)i << logF);
        for (int r = threadIdx.x; r < F; r += blockDim.x) tile_y[pad8(r)] = tile_x[pad8(r)] * H[r];
        __syncthreads();
        tile_fft<T, FFT_INV, true>(tile_y, tw, logF, 1, 1);
        const i64 row = (chan * geo.n_bands + band) * N;
        double acc = 0.0;
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            const i64 n = n0 + v;
            if (n < N) {
                const cplx<T> y = tile_y[pad8(half + v)];
                const T pw = norm2(y);
                if (out_c) out_c[row + n] = y;
                if (out_p) out_p[row + n] = pw;
                acc += (double)pw;
            }
        }
        if (band_sum) {
            acc = block_sum(acc, scratch);
            if (threadIdx.x == 0) atomicAdd(&band_sum[chan * geo.n_bands + band], acc);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host side: which band takes which route
enum { CWT_ROUTE_PLAIN = 0, CWT_ROUTE_DEC = 1, CWT_ROUTE_OS = 2 };

struct CwtFastPlan {
    std::vector<int> route, logK;          // per band
    std::vector<int> dec_ids;              // decimated bands ordered by logK
    std::vector<int> os_ids;
    int os_half;
    int n_plain;
};

template <typename T>
static void cwt_fast_plan(const QiAtomBand* hb, int B, i64 N, int logL, int conv_mode, bool enable, CwtFastPlan& fp) {
    typedef CwtFastCfg<T> Cfg;
    fp.route.assign(B, CWT_ROUTE_PLAIN);
    fp.logK.assign(B, logL);
    fp.dec_ids.clear(); fp.os_ids.clear();
    fp.os_half = 0;
    fp.n_plain = B;
    if (!enable || conv_mode != QI_CONV_LINEAR_SAME || logL < 14 || N < 4 * CWTF_TILE) return;
    const double L = (double)(1ll << logL);
    for (int b = 0; b < B; ++b) {
        if (!hb[b].analytic || hb[b].p_im != 0.0 || !(hb[b].p_re > 0.0)) continue;
        const double s = 1.0 / sqrt(2.0 * hb[b].p_re);
        if (s < 1.0) continue;                                   // sub-sample atoms: more aliases than the closed form keeps
        const int half = (int)ceil(Cfg::U_CUT * s) + 2;
        if (half <= CWTF_OS_MAX_HALF) {
            fp.route[b] = CWT_ROUTE_OS;
            fp.os_ids.push_back(b);
            if (half > fp.os_half) fp.os_half = half;
            continue;
        }
        const double kmax = ceil(Cfg::U_CUT * L / (2.0 * M_PI * s)) + 1.0;
        const double need = Cfg::RHO * (2.0 * kmax + 4.0);
        int lk = logL - CWTF_MAX_LOGD;
        if (lk < 6) lk = 6;
        while (lk < logL && (double)(1ll << lk) < need) ++lk;
        if (lk <= logL - 2) { fp.route[b] = CWT_ROUTE_DEC; fp.logK[b] = lk; }
    }
    for (int lk = 0; lk <= logL; ++lk)
        for (int b = 0; b < B; ++b) if (fp.route[b] == CWT_ROUTE_DEC && fp.logK[b] == lk) fp.dec_ids.push_back(b);
    fp.os_half = (fp.os_half + 15) & ~15;
    fp.n_plain = B - (int)fp.dec_ids.size() - (int)fp.os_ids.size();
}

}  // namespace qi
