// qi_mr_autocorr.cuh -- band powers of the level-0 bands of the multirate CWT BEFORE their rows are computed.
//
// The information plane -log2(P / S + eps) needs the total power S of a record when a power value is stored.  The bands
// of the levels >= 1 know their power from their decimated samples (mr_total_kernel); the level-0 bands are convolved at
// the full rate, and their power used to be known only after the convolution had written their rows -- which cost a
// second pass over those rows (read P, write the information: 8.6 GB of the 78 GB a headline step moved).
// With y_b = kappa_b * x (the taps the convolution kernel uses, |d| <= h_b) the power of a band is a quadratic form of
// the record's autocorrelation:
//     sum_{all n} |y_b[n]|^2 = R_b[0] r[0] + 2 sum_{tau >= 1} Re R_b[tau] r[tau],
//     r[tau] = sum_n x[n] x[n + tau],   R_b[tau] = sum_d kappa_b[d] conj(kappa_b[d - tau]),   0 <= tau <= 2 h_b,
// minus the 2 h_b outputs that fall outside the record (n < 0, n >= N), which are evaluated directly.
//
// r[tau] is GEMM-shaped: with A[i][m] = x[16 m + i] (the record itself, read as a 16-row matrix),
//     C_t[i][j] = sum_m A[i][m] A[j][m + t]        ->        r[16 t + j - i] += C_t[i][j]      (t = 0: j >= i only),
// i.e. (16 x M) . (M x 16) products with M = N / 16 as the contraction length: it runs on the tensor cores
// (mma.sync m16n8k8, TF32 operands, fp32 accumulators per 2048-sample tile, fp64 across tiles).  Rounding x to TF32
// perturbs r by uncorrelated errors of relative size 2^-11 per sample: 1e-7 of r[0] over a 2^24-sample record, and a
// bias of 2^-22 / 3 = 8e-8 on r[0] -- two orders below the 1e-5 of the estimates of the other bands.
// The same sums on the FP32 pipe cost 153 multiply-adds per sample (0.47 ms per headline step at 60 % of the FFMA2
// peak); on the tensor pipe the kernel is bound by its shared-memory fragment loads.
#pragma once
#include "qi_platform.cuh"
#include "qi_reduce.cuh"
#include "qi_mr_expand.cuh"

namespace qi {

constexpr int AC_TS = 2048;                 // samples of the A side per CTA (26 KB of static shared memory)
constexpr int AC_TB = 11;                   // lag blocks of 16 per pass
constexpr int AC_LAGS = 16 * AC_TB;         // lags [AC_LAGS * pass - 15, AC_LAGS * (pass + 1) + 15) touched by a pass
constexpr int AC_PITCH = 24;                // shared-memory pitch of a 16-sample column: fragment loads are conflict free
constexpr int AC_THREADS = 128;

QI_DEV unsigned ac_tf32(float v) {
#ifdef QI_EMUL
    unsigned u;
    memcpy(&u, &v, 4);
    return u;
#else
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
#endif
}

// r[c][tau] += sums over this CTA's tiles.  grid (CTAs per channel, C, passes); a CTA walks the tiles x, x + gridDim.x, ...
// of its channel.  The mma accumulators stay in registers for AC_FLUSH tiles (256 products per element), are then summed
// over the four warps in shared memory and folded, lag by lag, into one fp64 register per thread; one global atomic per
// lag and CTA at the very end (one per lag and TILE serialised 13.6 M fp64 atomics on 1400 addresses: 4 ms).
// r: [C][r_stride] fp64, zeroed beforehand.
constexpr int AC_FLUSH = 8;
__global__ void __launch_bounds__(AC_THREADS)
mr_autocorr_kernel(const float* __restrict__ x, i64 stride, i64 n_points, double* __restrict__ r, int r_stride, int n_lags) {
    __shared__ unsigned sa[(AC_TS / 16) * AC_PITCH];                              // A side: x[n0 .. n0 + TS)
    __shared__ unsigned sb[(AC_TS / 16 + AC_TB + 1) * AC_PITCH];                 // B side: x[n0 + 176 pass .. + TS + 192)
    __shared__ float stage[2 * AC_TB * 4 * 32];                                   // one warp's accumulators: [n-tile][e][lane]
    const i64 chan = blockIdx.y;
    const int pass = blockIdx.z;
    const float* xs = x + chan * stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    constexpr int KSTEPS = AC_TS / 16 / 8;                                        // k-steps of 8 columns per tile
    constexpr int NT = 2 * AC_TB;                                                 // n-tiles of 8 lags-within-block
    const i64 n_tiles = (n_points + AC_TS - 1) / AC_TS;
    double lag_sum[2] = {0.0, 0.0};                                               // thread t owns the slots t and t + 128 (rel + 16)
#ifndef QI_EMUL
    float d[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.0f; }
#else
    float lag_emul[AC_LAGS + 32];
    for (int i = 0; i < AC_LAGS + 32; ++i) lag_emul[i] = 0.0f;
#endif
    int pending = 0;
    for (i64 tile = blockIdx.x; tile < n_tiles || pending; tile += gridDim.x) {
        const bool live = tile < n_tiles;
        if (live) {
            const i64 n0 = tile * AC_TS;
            __syncthreads();                                                      // the previous tile's fragments are read
            for (int i = threadIdx.x; i < AC_TS; i += AC_THREADS) {
                const i64 n = n0 + i;
                sa[(i >> 4) * AC_PITCH + (i & 15)] = ac_tf32(n < n_points ? xs[n] : 0.0f);
            }
            for (int i = threadIdx.x; i < AC_TS + 16 * (AC_TB + 1); i += AC_THREADS) {
                const i64 n = n0 + (i64)AC_LAGS * pass + i;
                sb[(i >> 4) * AC_PITCH + (i & 15)] = ac_tf32(n < n_points ? xs[n] : 0.0f);
            }
            __syncthreads();
#ifndef QI_EMUL
            for (int ks = warp; ks < KSTEPS; ks += AC_THREADS / 32) {
                const int m0 = ks * 8;
                const unsigned a0 = sa[(m0 + tig) * AC_PITCH + g], a1 = sa[(m0 + tig) * AC_PITCH + g + 8];
                const unsigned a2 = sa[(m0 + tig + 4) * AC_PITCH + g], a3 = sa[(m0 + tig + 4) * AC_PITCH + g + 8];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int t = nt >> 1, j0 = 8 * (nt & 1);
                    const unsigned b0 = sb[(m0 + tig + t) * AC_PITCH + j0 + g], b1 = sb[(m0 + tig + 4 + t) * AC_PITCH + j0 + g];
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(d[nt][0]), "+f"(d[nt][1]), "+f"(d[nt][2]), "+f"(d[nt][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                }
            }
#else
            if (threadIdx.x == 0) {                                               // CPU emulation of the same sums
                for (int m = 0; m < AC_TS / 16; ++m)
                    for (int t = 0; t < AC_TB; ++t)
                        for (int i = 0; i < 16; ++i)
                            for (int j = 0; j < 16; ++j) {
                                float va, vb;
                                memcpy(&va, &sa[m * AC_PITCH + i], 4);
                                memcpy(&vb, &sb[(m + t) * AC_PITCH + j], 4);
                                lag_emul[16 * t + j - i + 16] += va * vb;
                            }
            }
#endif
            ++pending;
        }
        if (pending == AC_FLUSH || (!live && pending)) {
            // C_t[i][j] of the four warps -> stage (summed warp after warp), then every thread gathers its two lags:
            // rel = 16 t + j - i lives in n-tile 2 t + (j >> 3), element e = 2 (i >> 3) + (j & 1), lane 4 (i & 7) + ((j & 7) >> 1)
#ifndef QI_EMUL
            for (int w = 0; w < AC_THREADS / 32; ++w) {
                __syncthreads();
                if (warp == w) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float* slot = &stage[(nt * 4 + e) * 32 + lane];
                            *slot = (w ? *slot : 0.0f) + d[nt][e];
                            d[nt][e] = 0.0f;
                        }
                }
            }
            __syncthreads();
            for (int half = 0; half < 2; ++half) {
                const int rel = (int)threadIdx.x + 128 * half - 16;
                if (rel >= AC_LAGS + 16 || (pass == 0 && rel < 0)) continue;      // pass 0, t = 0: pairs counted once, as (n, n + tau)
                float sum = 0.0f;
                for (int t = 0; t < AC_TB; ++t)
                    for (int i = 0; i < 16; ++i) {
                        const int j = rel - 16 * t + i;
                        if (j < 0 || j > 15) continue;
                        sum += stage[((2 * t + (j >> 3)) * 4 + 2 * (i >> 3) + (j & 1)) * 32 + 4 * (i & 7) + ((j & 7) >> 1)];
                    }
                lag_sum[half] += (double)sum;
            }
#else
            if (threadIdx.x == 0) {
                for (int i = 0; i < AC_LAGS + 32; ++i) {
                    const int tau = AC_LAGS * pass + i - 16;
                    if (!(pass == 0 && i < 16) && tau >= 0 && tau < n_lags && lag_emul[i] != 0.0f)
                        atomicAdd(&r[chan * r_stride + tau], (double)lag_emul[i]);
                    lag_emul[i] = 0.0f;
                }
            }
#endif
            pending = 0;
        }
        if (!live) break;
    }
#ifndef QI_EMUL
    for (int half = 0; half < 2; ++half) {
        const int rel = (int)threadIdx.x + 128 * half - 16;
        const int tau = AC_LAGS * pass + rel;
        if (rel < AC_LAGS + 16 && tau >= 0 && tau < n_lags && lag_sum[half] != 0.0) atomicAdd(&r[chan * r_stride + tau], lag_sum[half]);
    }
#else
    (void)lag_sum; (void)warp; (void)g; (void)tig; (void)KSTEPS; (void)NT; (void)lane;
#endif
}

// Predicted power of one level-0 band of one record: grid (level-0 bands, C), 256 threads.
// kappa[d] = amp exp(-t^2 / 2 s^2) exp(i omega t), t = d - 1/2, |d| <= h, |t| <= (N - 1) / 2, rounded to float32 exactly as
// mr_table_kernel stores it (the convolution kernel's effective taps).  dyn smem: (2 h + 1) double2 + (2 h + 1) doubles.
__global__ void __launch_bounds__(256)
mr_level0_predict_kernel(const MrDevBand* __restrict__ bands, int band_first, int n_bands, i64 n_points, int half_w_cap,
                         const float* __restrict__ x, i64 stride, const double* __restrict__ r, int r_stride,
                         double* __restrict__ predicted) {
    QI_DYN_SMEM(smem_raw);
    __shared__ double scratch[32];
    const int b = band_first + blockIdx.x;
    const i64 chan = blockIdx.y;
    const MrDevBand band = bands[b];
    int h = (int)ceil(5.2 * band.scale) + 1;
    if (h > half_w_cap) h = half_w_cap;
    const int nk = 2 * h + 1;
    double2* kap = reinterpret_cast<double2*>(smem_raw);                          // kap[d + h]
    const double tmax = 0.5 * (double)(n_points - 1);
    for (int p = threadIdx.x; p < nk; p += blockDim.x) {
        const double t = (double)(p - h) - 0.5;
        double re = 0.0, im = 0.0;
        if (fabs(t) <= tmax) {
            const double u = t / band.scale;
            const double env = band.amp * exp(-0.5 * u * u);
            double s, c;
            sincos(band.omega * t, &s, &c);
            re = (double)(float)(env * c); im = (double)(float)(env * s);
        }
        kap[p] = make_double2(re, im);
    }
    __syncthreads();
    const double* rc = r + chan * r_stride;
    double acc = 0.0;
    // all outputs of the full linear convolution, through the autocorrelation
    for (int tau = threadIdx.x; tau < nk; tau += blockDim.x) {
        double rr = 0.0;                                                          // Re R[tau]
        for (int p = tau; p < nk; ++p) rr += kap[p].x * kap[p - tau].x + kap[p].y * kap[p - tau].y;
        acc += (tau ? 2.0 : 1.0) * rr * rc[tau];
    }
    // minus the outputs outside the record: y[n] = sum_d kappa[d] x[n - d] for n in [-h, 0) and [N, N + h)
    const float* xs = x + chan * stride;
    for (int e = threadIdx.x; e < 2 * h; e += blockDim.x) {
        const i64 n = e < h ? (i64)(e - h) : n_points + (e - h);
        double yr = 0.0, yi = 0.0;
        for (int p = 0; p < nk; ++p) {
            const i64 k = n - (p - h);
            if (k >= 0 && k < n_points) { const double v = (double)xs[k]; yr += kap[p].x * v; yi += kap[p].y * v; }
        }
        acc -= yr * yr + yi * yi;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) predicted[chan * n_bands + b] = acc;
}

}  // namespace qi
