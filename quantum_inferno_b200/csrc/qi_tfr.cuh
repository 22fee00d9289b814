// qi_tfr.cuh -- pieces shared by the CWT and Stockwell drivers: output geometry and the fused
// "slice + |.|^2 + fp64 band-sum" sink of the last inverse-FFT pass.
#pragma once
#include "qi_fft.cuh"
#include "qi_reduce.cuh"

namespace qi {

struct CwtGeom {
    i64 n_points;      // N
    i64 n_channels;    // C
    int n_bands;       // B
    int logL;
    int half_shift;    // 1 if the 'same' slice leaves a half-sample offset (N even)
    int conv_mode;
    i64 d_min, d_max;  // kernel lag range placed circularly (linear mode)
    i64 centre_idx;    // (N-1)//2
    double fs;
};

// Sink of the last inverse pass: slice, optional rotation, complex / power planes, fp64 band sums.
template <typename T> struct DstCwtOut {
    cplx<T>* out_c; T* out_p; double* band_sum; int band0; CwtGeom geo; double acc;
    QI_DEV void store(i64 batch, i64 n, cplx<T> v) {
        if (n >= geo.n_points) return;
        const i64 chan = batch % geo.n_channels;
        const i64 band = band0 + batch / geo.n_channels;
        i64 no = n;
        if (geo.conv_mode == QI_CONV_CIRC_CORR) no = (n + (geo.n_points >> 1)) & (geo.n_points - 1);
        const i64 o = (chan * geo.n_bands + band) * geo.n_points + no;
        if (out_c) out_c[o] = v;
        const T p = norm2(v);
        if (out_p) out_p[o] = p;
        acc += (double)p;
    }
    QI_DEV void finish(i64 batch, unsigned char* scratch) {
        if (!band_sum) return;
        const double s = block_sum(acc, reinterpret_cast<double*>(scratch));
        if (threadIdx.x == 0) {
            const i64 chan = batch % geo.n_channels;
            const i64 band = band0 + batch / geo.n_channels;
            atomicAdd(&band_sum[chan * geo.n_bands + band], s);
        }
    }
};

inline int ceil_log2_i64(i64 v) { int l = 0; while ((1ll << l) < v) ++l; return l; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }


}  // namespace qi
