// qi_info.cu -- power / information / entropy reductions of a time-frequency power array.
//
// Replaces quantum_inferno/tfr_info.py:65-94 (scale_log2_64, scale_power_bits, power_dynamics_scaled_bits),
// :97-135 (Shannon marginals, EPS32) and :203-260 (ShannonStft global / per-time / per-frequency, EPS64).
// Every reduction accumulates in fp64 whatever the plane dtype is.  All kernels are
// HBM-streaming: one coalesced read of the power plane, float4/double2-friendly contiguous rows.
#include "qi_platform.cuh"
#include "qi_host.h"
#include "qi_reduce.cuh"

namespace qi {

// ---------------------------------------------------------------- row sums + max  (grid: (splits, F, M))
template <typename T>
__global__ void __launch_bounds__(256)
row_reduce_kernel(const T* __restrict__ p, i64 F, i64 Tn, double* __restrict__ row_sum, double* __restrict__ tot,
                  double* __restrict__ mx) {
    __shared__ double scratch[32];
    const i64 m = blockIdx.z, f = blockIdx.y;
    const T* row = p + (m * F + f) * Tn;
    const i64 chunk = (Tn + gridDim.x - 1) / gridDim.x;
    const i64 t0 = (i64)blockIdx.x * chunk;
    const i64 t1 = t0 + chunk < Tn ? t0 + chunk : Tn;
    double s = 0.0, mv = 0.0;
    for (i64 t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const double v = (double)row[t];
        s += v;
        mv = v > mv ? v : mv;
    }
    s = block_sum(s, scratch);
    mv = block_max(mv, scratch);
    if (threadIdx.x == 0) {
        if (row_sum) atomicAdd(&row_sum[m * F + f], s);
        if (tot) atomicAdd(&tot[m], s);
        if (mx) atomic_max_nonneg(&mx[m], mv);
    }
}

// ---------------------------------------------------------------- column sums (grid: (ceil(T/256), M))
template <typename T>
__global__ void __launch_bounds__(256)
col_sum_kernel(const T* __restrict__ p, i64 F, i64 Tn, double* __restrict__ col_sum) {
    const i64 m = blockIdx.y;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tn) return;
    const T* base = p + m * F * Tn + t;
    double s = 0.0;
    for (i64 f = 0; f < F; ++f) s += (double)base[f * Tn];
    col_sum[m * Tn + t] = s;
}

// ---------------------------------------------------------------- Shannon planes
// mode 0: pdf = P / norm[m]                (tfr_info.py:236)
// mode 1: pdf = (1/norm[m,t] + eps) * P    (tfr_info.py:247, per time)
// mode 2: pdf = (1/norm[m,f] + eps) * P    (tfr_info.py:259, per frequency)
// mode 3: pdf = P                          (tfr_info.py:97-103 marginals; eps = EPS32)
struct ShannonArgs {
    i64 F, Tn;
    int mode;
    double eps, log2_d, inv_ref_bits;
};

template <typename T>
__global__ void __launch_bounds__(256)
shannon_kernel(const T* __restrict__ p, const double* __restrict__ norm, ShannonArgs a, T* __restrict__ o_pdf,
               T* __restrict__ o_info, T* __restrict__ o_bits, T* __restrict__ o_isnr, T* __restrict__ o_esnr, double* __restrict__ ent_sum) {
    __shared__ double scratch[32];
    const i64 m = blockIdx.z, f = blockIdx.y;
    const i64 row = (m * a.F + f) * a.Tn;
    const T eps = (T)a.eps;
    T rnorm = (T)1;
    if (a.mode == 0) rnorm = (T)norm[m];
    else if (a.mode == 2) rnorm = (T)((T)1 / (T)norm[m * a.F + f] + eps);
    double acc = 0.0;
    const i64 chunk = (a.Tn + gridDim.x - 1) / gridDim.x;
    const i64 t0 = (i64)blockIdx.x * chunk;
    const i64 t1 = t0 + chunk < a.Tn ? t0 + chunk : a.Tn;
    for (i64 t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const T pw = p[row + t];
        T pdf;
        if (a.mode == 0) pdf = pw / rnorm;
        else if (a.mode == 1) pdf = ((T)1 / (T)norm[m * a.Tn + t] + eps) * pw;
        else if (a.mode == 2) pdf = rnorm * pw;
        else pdf = pw;
        const T info = -log2(pdf + eps);
        const T bits = pdf * info;
        if (o_pdf) o_pdf[row + t] = pdf;
        if (o_info) o_info[row + t] = info;
        if (o_bits) o_bits[row + t] = bits;
        if (o_isnr) o_isnr[row + t] = (T)a.log2_d - info;
        if (o_esnr) o_esnr[row + t] = bits * (T)a.inv_ref_bits;
        acc += (double)bits;
    }
    if (ent_sum) {
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) atomicAdd(&ent_sum[m * a.F + f], acc);
    }
}

// Streaming special case of the above for the fused north-star pass: float32 plane, mode 0, only the information
// plane and the per-band entropy sums.  128-bit loads/stores, MUFU log2, one fp64 add per four cells.
QI_DEV float fast_log2f(float v) {
#ifdef QI_EMUL
    return std::log2(v);
#else
    return __log2f(v);
#endif
}

__global__ void __launch_bounds__(256)
info_plane_f32_kernel(const float4* __restrict__ p, const double* __restrict__ norm, i64 F, i64 T4, float eps,
                      float4* __restrict__ o_info, double* __restrict__ ent_sum) {
    __shared__ double scratch[32];
    const i64 m = blockIdx.z, f = blockIdx.y;
    const i64 row = (m * F + f) * T4;
    const float inv = (float)(1.0 / norm[m]);
    const i64 chunk = (T4 + gridDim.x - 1) / gridDim.x;
    const i64 t0 = (i64)blockIdx.x * chunk;
    const i64 t1 = t0 + chunk < T4 ? t0 + chunk : T4;
    double acc = 0.0;
    for (i64 t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const float4 v = p[row + t];
        const float d0 = v.x * inv, d1 = v.y * inv, d2 = v.z * inv, d3 = v.w * inv;
        const float i0 = -fast_log2f(d0 + eps), i1 = -fast_log2f(d1 + eps);
        const float i2 = -fast_log2f(d2 + eps), i3 = -fast_log2f(d3 + eps);
        o_info[row + t] = make_float4(i0, i1, i2, i3);
        acc += (double)((d0 * i0 + d1 * i1) + (d2 * i2 + d3 * i3));
    }
    if (ent_sum) {
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) atomicAdd(&ent_sum[m * F + f], acc);
    }
}

// The float64 counterpart (the information pass of the exact float64 CWT path): 128-bit loads / stores, two in flight,
// one reciprocal per row instead of a division per cell (P / S and P * (1 / S) differ by one ulp).
__global__ void __launch_bounds__(256)
info_plane_f64_kernel(const double2* __restrict__ p, const double* __restrict__ norm, i64 F, i64 T2, double eps,
                      double2* __restrict__ o_info, double* __restrict__ ent_sum) {
    __shared__ double scratch[32];
    const i64 m = blockIdx.z, f = blockIdx.y;
    const i64 row = (m * F + f) * T2;
    const double inv = 1.0 / norm[m];
    const i64 chunk = (T2 + gridDim.x - 1) / gridDim.x;
    const i64 t0 = (i64)blockIdx.x * chunk;
    const i64 t1 = t0 + chunk < T2 ? t0 + chunk : T2;
    double acc = 0.0;
    auto one = [&](const double2 v, i64 t) {
        const double d0 = v.x * inv, d1 = v.y * inv;
        const double i0 = -log2(d0 + eps), i1 = -log2(d1 + eps);
        o_info[row + t] = make_double2(i0, i1);
        acc += d0 * i0 + d1 * i1;
    };
    i64 t = t0 + threadIdx.x;
    for (; t + (i64)blockDim.x < t1; t += 2 * (i64)blockDim.x) {
        const double2 v0 = p[row + t], v1 = p[row + t + blockDim.x];
        one(v0, t); one(v1, t + blockDim.x);
    }
    for (; t < t1; t += blockDim.x) one(p[row + t], t);
    if (ent_sum) {
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) atomicAdd(&ent_sum[m * F + f], acc);
    }
}

// bits = log2(P + eps) - log2(max[m] + eps)   (tfr_info.py:73-79)
template <typename T>
__global__ void __launch_bounds__(256)
power_bits_kernel(const T* __restrict__ p, i64 per_mat, const double* __restrict__ mx, double eps, T* __restrict__ out) {
    const i64 m = blockIdx.y;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_mat) return;
    const T e = (T)eps;
    const T top = log2((T)mx[m] + e);
    out[m * per_mat + i] = log2(p[m * per_mat + i] + e) - top;
}

// marginal of the time-domain record: sig_n = x/sqrt(sum x^2), marginal = sig_n^2  (tfr_info.py:147-151)
template <typename T>
__global__ void __launch_bounds__(256)
sumsq_kernel(const T* __restrict__ x, i64 n, i64 stride, double* __restrict__ out) {
    __shared__ double scratch[32];
    const i64 m = blockIdx.y;
    double s = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = (double)x[m * stride + i];
        s += v * v;
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) atomicAdd(&out[m], s);
}
template <typename T>
__global__ void __launch_bounds__(256)
tdr_marginal_kernel(const T* __restrict__ x, i64 n, i64 stride, const double* __restrict__ sumsq,
                    T* __restrict__ sig_n, T* __restrict__ marg) {
    const i64 m = blockIdx.y;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = x[m * stride + i] / (T)sqrt(sumsq[m]);
    if (sig_n) sig_n[m * n + i] = v;
    marg[m * n + i] = v * v;
}

template <typename T>
__global__ void __launch_bounds__(256)
abs_log2_kernel(const T* __restrict__ in, i64 n, int is_complex, int square, double eps, T* __restrict__ out) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T mag2, mag;
    if (is_complex == 2) {              // signed real: log2(x + eps), no modulus (tfr_info.py:65-70)
        const T v = in[i];
        mag2 = v * v;
        mag = v;
    } else if (is_complex) {
        const cplx<T> v = reinterpret_cast<const cplx<T>*>(in)[i];
        mag2 = v.re * v.re + v.im * v.im;
        mag = square == 1 ? (T)0 : (T)hypot(v.re, v.im);
    } else {
        const T v = in[i];
        mag2 = v * v;
        mag = v < (T)0 ? -v : v;
    }
    out[i] = square == 1 ? mag2 + (T)eps : (square == 2 ? mag + (T)eps : log2(mag + (T)eps));
}

static unsigned splits_for(i64 Tn, i64 rows) {
    // many short CTAs (>= 64 waves over 148 SMs x 8 resident CTAs, >= 32K elements each) so the tail wave is noise
    i64 s = (148 * 8 * 64 + rows - 1) / rows;
    const i64 cap = (Tn + 32767) / 32768;
    if (s > cap) s = cap;
    if (s < 1) s = 1;
    return (unsigned)s;
}

template <typename T>
static int power_reduce_impl(const void* p, i64 M, i64 F, i64 Tn, double* row_sum, double* col_sum, double* tot,
                             double* mx, cudaStream_t st) {
    if (F > 65535 || M > 65535) return QI_ERR_UNSUPPORTED;
    prof_set_category(QI_CAT_INFO);
    if (row_sum) cudaMemsetAsync(row_sum, 0, sizeof(double) * (size_t)M * F, st);
    if (tot) cudaMemsetAsync(tot, 0, sizeof(double) * (size_t)M, st);
    if (mx) cudaMemsetAsync(mx, 0, sizeof(double) * (size_t)M, st);
    if (row_sum || tot || mx) {
        dim3 grid(splits_for(Tn, M * F), (unsigned)F, (unsigned)M);
        QI_LAUNCH((row_reduce_kernel<T>), grid, dim3(256), 0, st, static_cast<const T*>(p), F, Tn, row_sum, tot, mx);
    }
    if (col_sum) {
        dim3 grid((unsigned)((Tn + 255) / 256), (unsigned)M);
        QI_LAUNCH((col_sum_kernel<T>), grid, dim3(256), 0, st, static_cast<const T*>(p), F, Tn, col_sum);
    }
    return check_cuda("qi_power_reduce");
}

template <typename T>
static int shannon_impl(const void* p, i64 M, i64 F, i64 Tn, int mode, const double* norm, double eps, double deg_free,
                        void* o_pdf, void* o_info, void* o_bits, void* o_isnr, void* o_esnr, double* ent_sum,
                        cudaStream_t st) {
    if (F > 65535 || M > 65535) return QI_ERR_UNSUPPORTED;
    prof_set_category(QI_CAT_INFO);
    ShannonArgs a;
    a.F = F; a.Tn = Tn; a.mode = mode; a.eps = eps;
    a.log2_d = log2(deg_free);
    a.inv_ref_bits = 1.0 / (log2(deg_free) / deg_free);
    if (ent_sum) cudaMemsetAsync(ent_sum, 0, sizeof(double) * (size_t)M * F, st);
    dim3 grid(splits_for(Tn, M * F), (unsigned)F, (unsigned)M);
    if (sizeof(T) == 4 && mode == 0 && o_info && !o_pdf && !o_bits && !o_isnr && !o_esnr && (Tn % 4) == 0 &&
        ((uintptr_t)p % 16) == 0 && ((uintptr_t)o_info % 16) == 0) {
        QI_LAUNCH(info_plane_f32_kernel, grid, dim3(256), 0, st, static_cast<const float4*>(p), norm, F, Tn / 4,
                  (float)eps, static_cast<float4*>(o_info), ent_sum);
        return check_cuda("qi_shannon");
    }
    if (sizeof(T) == 8 && mode == 0 && o_info && !o_pdf && !o_bits && !o_isnr && !o_esnr && (Tn % 2) == 0 &&
        ((uintptr_t)p % 16) == 0 && ((uintptr_t)o_info % 16) == 0) {
        QI_LAUNCH(info_plane_f64_kernel, grid, dim3(256), 0, st, static_cast<const double2*>(p), norm, F, Tn / 2, eps,
                  static_cast<double2*>(o_info), ent_sum);
        return check_cuda("qi_shannon");
    }
    QI_LAUNCH((shannon_kernel<T>), grid, dim3(256), 0, st, static_cast<const T*>(p), norm, a, static_cast<T*>(o_pdf), static_cast<T*>(o_info),
              static_cast<T*>(o_bits), static_cast<T*>(o_isnr), static_cast<T*>(o_esnr), ent_sum);
    return check_cuda("qi_shannon");
}

}  // namespace qi

extern "C" {

int qi_power_reduce(const void* power, int64_t M, int64_t F, int64_t Tn, int dtype, double* row_sum, double* col_sum,
                    double* total, double* max_value, void* stream) {
    if (!power || M <= 0 || F <= 0 || Tn <= 0) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32) return qi::power_reduce_impl<float>(power, M, F, Tn, row_sum, col_sum, total, max_value, st);
    if (dtype == QI_F64) return qi::power_reduce_impl<double>(power, M, F, Tn, row_sum, col_sum, total, max_value, st);
    return QI_ERR_ARG;
}

int qi_shannon(const void* power, int64_t M, int64_t F, int64_t Tn, int dtype, int mode, const double* norm, double eps,
               double deg_free, void* out_pdf, void* out_info, void* out_bits, void* out_isnr, void* out_esnr,
               double* entropy_sum, void* stream) {
    if (!power || M <= 0 || F <= 0 || Tn <= 0 || mode < 0 || mode > 3 || (mode != 3 && !norm)) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32)
        return qi::shannon_impl<float>(power, M, F, Tn, mode, norm, eps, deg_free, out_pdf, out_info, out_bits, out_isnr, out_esnr, entropy_sum, st);
    if (dtype == QI_F64)
        return qi::shannon_impl<double>(power, M, F, Tn, mode, norm, eps, deg_free, out_pdf, out_info, out_bits, out_isnr, out_esnr, entropy_sum, st);
    return QI_ERR_ARG;
}

int qi_power_bits(const void* power, int64_t M, int64_t per_mat, int dtype, const double* max_value, double eps, void* out,
                  void* stream) {
    if (!power || !max_value || !out || M <= 0 || per_mat <= 0 || M > 65535) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((per_mat + 255) / 256), (unsigned)M);
    if (dtype == QI_F32)
        QI_LAUNCH((qi::power_bits_kernel<float>), grid, dim3(256), 0, st, static_cast<const float*>(power), (qi::i64)per_mat, max_value, eps, static_cast<float*>(out));
    else if (dtype == QI_F64)
        QI_LAUNCH((qi::power_bits_kernel<double>), grid, dim3(256), 0, st, static_cast<const double*>(power), (qi::i64)per_mat, max_value, eps, static_cast<double*>(out));
    else return QI_ERR_ARG;
    return qi::check_cuda("qi_power_bits");
}

int qi_abs_log2(const void* in, int64_t n, int dtype, int is_complex, int square, double eps, void* out, void* stream) {
    if (!in || !out || n <= 0) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((n + 255) / 256));
    if (dtype == QI_F32)
        QI_LAUNCH((qi::abs_log2_kernel<float>), grid, dim3(256), 0, st, static_cast<const float*>(in), (qi::i64)n, is_complex, square, eps, static_cast<float*>(out));
    else if (dtype == QI_F64)
        QI_LAUNCH((qi::abs_log2_kernel<double>), grid, dim3(256), 0, st, static_cast<const double*>(in), (qi::i64)n, is_complex, square, eps, static_cast<double*>(out));
    else return QI_ERR_ARG;
    return qi::check_cuda("qi_abs_log2");
}

int qi_tdr_marginal(const void* sig, int64_t M, int64_t n, int64_t stride, int dtype, double* sumsq, void* out_sig,
                    void* out_marginal, void* stream) {
    if (!sig || !sumsq || !out_marginal || M <= 0 || n <= 0 || M > 65535) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(sumsq, 0, sizeof(double) * (size_t)M, st);
    unsigned nb = (unsigned)((n + 255) / 256);
    dim3 g1(nb > 1024 ? 1024 : nb, (unsigned)M), g2(nb, (unsigned)M);
    if (dtype == QI_F32) {
        QI_LAUNCH((qi::sumsq_kernel<float>), g1, dim3(256), 0, st, static_cast<const float*>(sig), (qi::i64)n, (qi::i64)stride, sumsq);
        QI_LAUNCH((qi::tdr_marginal_kernel<float>), g2, dim3(256), 0, st, static_cast<const float*>(sig), (qi::i64)n, (qi::i64)stride, (const double*)sumsq, static_cast<float*>(out_sig), static_cast<float*>(out_marginal));
    } else if (dtype == QI_F64) {
        QI_LAUNCH((qi::sumsq_kernel<double>), g1, dim3(256), 0, st, static_cast<const double*>(sig), (qi::i64)n, (qi::i64)stride, sumsq);
        QI_LAUNCH((qi::tdr_marginal_kernel<double>), g2, dim3(256), 0, st, static_cast<const double*>(sig), (qi::i64)n, (qi::i64)stride, (const double*)sumsq, static_cast<double*>(out_sig), static_cast<double*>(out_marginal));
    } else return QI_ERR_ARG;
    return qi::check_cuda("qi_tdr_marginal");
}

}  // extern "C"
