// qi_stx.cu -- Stockwell transform on the shared FFT core.
//
// One forward FFT per record (n = 2^m, no padding: the reference's product is circular), then per band
// the first inverse pass gathers the spectrum shifted by the band's bin index, multiplies by the Gaussian
// window exp(-0.5*sigma^2*w_k^2) synthesised on the fly, and the last inverse pass fuses |.|^2 / band sums.
// Reference: quantum_inferno/styx_stx.py:213-234 and :166-190.
#include "qi_fft.cuh"
#include "qi_host.h"
#include "qi_tfr.cuh"

#include <vector>

namespace qi {

struct DevStxBand { double q; long long shift; long long kmax; };   // q = sigma*2*pi/n; |k| > kmax: the window underflows to 0

template <typename T> struct SrcStxSpec {
    const cplx<T>* spec; const DevStxBand* bands; int band0; CwtGeom geo;
    QI_DEV cplx<T> load(i64 batch, i64 e) const {
        const i64 chan = batch % geo.n_channels;
        const DevStxBand b = bands[band0 + (int)(batch / geo.n_channels)];
        const i64 n = 1ll << geo.logL;
        const i64 k = (i64)brev_bits((unsigned)e, geo.logL);
        const i64 ks = (k < (n >> 1)) ? k : k - n;                 // signed bin (fftfreq ordering)
        // beyond kmax exp() returns exactly 0 in T: same product, without the gather and the exponential
        if (ks > b.kmax || -ks > b.kmax) return mk<T>((T)0, (T)0);
        const i64 ksrc = (k + b.shift) & (n - 1);
        const cplx<T> X = spec[(chan << geo.logL) + (i64)brev_bits((unsigned)ksrc, geo.logL)];
        const T u = (T)b.q * (T)ks;
        const T w = exp((T)-0.5 * u * u) * (T)(1.0 / (double)n);
        return X * w;
    }
};

template <typename T>
__global__ void stx_windows_kernel(const DevStxBand* bands, i64 n, cplx<T>* out) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const DevStxBand b = bands[blockIdx.y];
    const i64 ks = (k < (n >> 1)) ? k : k - n;
    const T u = (T)b.q * (T)ks;
    out[(i64)blockIdx.y * n + k] = mk<T>(exp((T)-0.5 * u * u), (T)0);
}

struct StxLayout { int logL; i64 L; size_t off_bands, off_spec, off_work, total; int group; };

template <typename T> static StxLayout stx_layout(i64 C, i64 N, int B, int group) {
    StxLayout lo;
    lo.logL = ceil_log2_i64(N);
    lo.L = 1ll << lo.logL;
    if (group < 1) group = 1;
    if (group > B) group = B;
    while ((i64)group * C > 65535 && group > 1) --group;
    lo.group = group;
    size_t o = 0;
    lo.off_bands = o; o = align_up(o + sizeof(DevStxBand) * (size_t)B, 256);
    lo.off_spec = o; o = align_up(o + sizeof(cplx<T>) * (size_t)C * lo.L, 256);
    lo.off_work = o; o = align_up(o + sizeof(cplx<T>) * (size_t)group * C * lo.L, 256);
    lo.total = o;
    return lo;
}

// u_zero: |u| from which exp(-0.5 u^2) is exactly 0 in the arithmetic type (below the smallest subnormal, with margin:
// float32 exp(-104) = 6.8e-46 < 2^-150; float64 exp(-746) < 2^-1075)
static void upload_stx_bands(const QiStxBand* hb, int B, i64 N, double u_zero, DevStxBand* d_bands, cudaStream_t st) {
    std::vector<DevStxBand> db(B);
    for (int b = 0; b < B; ++b) {
        db[b].q = hb[b].sigma * 2.0 * M_PI / (double)N;
        db[b].shift = ((hb[b].shift % N) + N) % N;
        const double aq = fabs(db[b].q);
        const double km = aq > 0.0 ? ceil(u_zero / aq) + 2.0 : (double)N;
        db[b].kmax = km < (double)N ? (long long)km : (long long)N;
    }
    stage_to_device(d_bands, db.data(), sizeof(DevStxBand) * (size_t)B, st);
}

template <typename T>
static int stx_fft_impl(const void* sig, i64 C, i64 N, i64 stride, const QiStxBand* hb, int B, void* out_c,
                        void* out_p, double* band_sum, void* ws, size_t ws_bytes, int group, cudaStream_t st) {
    const StxLayout lo = stx_layout<T>(C, N, B, group);
    if (ws_bytes < lo.total) return QI_ERR_WORKSPACE;
    if (C > 65535) return QI_ERR_UNSUPPORTED;
    unsigned char* base = static_cast<unsigned char*>(ws);
    DevStxBand* d_bands = reinterpret_cast<DevStxBand*>(base + lo.off_bands);
    cplx<T>* spec = reinterpret_cast<cplx<T>*>(base + lo.off_spec);
    cplx<T>* work = reinterpret_cast<cplx<T>*>(base + lo.off_work);
    upload_stx_bands(hb, B, N, sizeof(T) == 4 ? 14.5 : 38.7, d_bands, st);

    CwtGeom geo;
    geo.n_points = N; geo.n_channels = C; geo.n_bands = B; geo.logL = lo.logL;
    geo.conv_mode = QI_CONV_LINEAR_SAME; geo.fs = 0; geo.centre_idx = 0; geo.half_shift = 0; geo.d_min = 0; geo.d_max = 0;

    const FftPlan plan = make_plan(lo.logL, (int)sizeof(cplx<T>));
    const int np = plan.npass;
    const T one = (T)1;
    prof_set_category(QI_CAT_FFT_FWD);
    for (int p = 0; p < np; ++p) {
        DstComplex<T> d{spec, lo.L, one};
        if (p == 0) {
            SrcRealPad<T> s{static_cast<const T*>(sig), stride, N};
            launch_pass<T, FFT_FWD>(plan, p, C, s, d, 0, st);
        } else {
            SrcComplex<T> s{spec, lo.L};
            launch_pass<T, FFT_FWD>(plan, p, C, s, d, 0, st);
        }
    }
    if (band_sum) cudaMemsetAsync(band_sum, 0, sizeof(double) * (size_t)C * B, st);
    for (int band0 = 0; band0 < B; band0 += lo.group) {
        const int g = (B - band0 < lo.group) ? (B - band0) : lo.group;
        const i64 nb = (i64)g * C;
        for (int p = np - 1; p >= 0; --p) {
            const bool first = (p == np - 1), last = (p == 0);
            SrcStxSpec<T> s1{spec, d_bands, band0, geo};
            SrcComplex<T> s2{work, lo.L};
            DstComplex<T> d1{work, lo.L, one};
            DstCwtOut<T> d2{static_cast<cplx<T>*>(out_c), static_cast<T*>(out_p), band_sum, band0, geo, 0.0};
            prof_set_category(last ? QI_CAT_INV_LAST : (first ? QI_CAT_INV_FIRST : QI_CAT_INV_MID));
            if (first && last) launch_pass<T, FFT_INV>(plan, p, nb, s1, d2, 256, st);
            else if (first) launch_pass<T, FFT_INV>(plan, p, nb, s1, d1, 0, st);
            else if (last) launch_pass<T, FFT_INV>(plan, p, nb, s2, d2, 256, st);
            else launch_pass<T, FFT_INV>(plan, p, nb, s2, d1, 0, st);
        }
    }
    prof_set_category(QI_CAT_OTHER);
    return check_cuda("qi_stx_fft");
}

template <typename T>
static int stx_windows_impl(const QiStxBand* hb, int B, i64 N, void* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < sizeof(DevStxBand) * (size_t)B) return QI_ERR_WORKSPACE;
    DevStxBand* d_bands = static_cast<DevStxBand*>(ws);
    upload_stx_bands(hb, B, N, 1e300, d_bands, st);
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
    QI_LAUNCH((stx_windows_kernel<T>), grid, dim3(256), 0, st, d_bands, N, static_cast<cplx<T>*>(out));
    return check_cuda("qi_stx_windows");
}

}  // namespace qi

extern "C" {

size_t qi_stx_workspace_bytes(int64_t C, int64_t N, int B, int group, int dtype) {
    if (C <= 0 || N <= 0 || B <= 0) return 0;
    return dtype == QI_F32 ? qi::stx_layout<float>(C, N, B, group).total : qi::stx_layout<double>(C, N, B, group).total;
}

int qi_stx_fft(const void* sig, int64_t C, int64_t N, int64_t stride, const QiStxBand* bands, int B, int dtype,
               void* out_tfr, void* out_power, double* band_sum, void* ws, size_t ws_bytes, int group, void* stream) {
    if (!sig || !bands || !ws || C <= 0 || N <= 0 || B <= 0 || stride < N) return QI_ERR_ARG;
    if (N & (N - 1)) return QI_ERR_ARG;
    if (N > (1ll << 30)) return QI_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32) return qi::stx_fft_impl<float>(sig, C, N, stride, bands, B, out_tfr, out_power, band_sum, ws, ws_bytes, group, st);
    if (dtype == QI_F64) return qi::stx_fft_impl<double>(sig, C, N, stride, bands, B, out_tfr, out_power, band_sum, ws, ws_bytes, group, st);
    return QI_ERR_ARG;
}

int qi_stx_windows(const QiStxBand* bands, int B, int64_t N, int dtype, void* out, void* ws, size_t ws_bytes, void* stream) {
    if (!bands || !out || !ws || B <= 0 || N <= 0) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32) return qi::stx_windows_impl<float>(bands, B, N, out, ws, ws_bytes, st);
    if (dtype == QI_F64) return qi::stx_windows_impl<double>(bands, B, N, out, ws, ws_bytes, st);
    return QI_ERR_ARG;
}

}  // extern "C"
