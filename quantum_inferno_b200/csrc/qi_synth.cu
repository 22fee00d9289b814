// qi_synth.cu -- device-side synthetic inputs: the deterministic generators of
// quantum_inferno/synth/benchmark_signals.py, so that bench-size records are made where they are used instead of
// crossing PCIe.
//
// Replaces the array expressions of quantum_chirp (benchmark_signals.py:92-101):
//     time = k - t_center;  u = time / chirp_scale
//     chirp_phase = omega * time + (0.5 * gamma) * u**2
//     wf = exp(-0.5 * u**2 + 1j * chirp_phase)      (gauss)      |      exp(1j * chirp_phase)
// and of well_tempered_tone (benchmark_signals.py:323-335):  cos((2 pi f_c) * k)  (t_center = 0, gamma = 0, no envelope,
// real part only).  Every product is formed in float64 in the reference's order, so the arguments handed to
// exp / sincos are bit-identical to numpy's; the results are rounded to the output dtype.
#include <math.h>
#include "qi_platform.cuh"
#include "qi_host.h"

namespace qi {

template <typename T>
__global__ void __launch_bounds__(256)
synth_chirp_kernel(i64 n, double t_center, double omega, double half_gamma, double chirp_scale, int gauss,
                   T* __restrict__ out_re, T* __restrict__ out_im, i64 stride, i64 k0) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const i64 m = blockIdx.y;
    const double time = (double)(k + k0) - t_center;
    const double u = time / chirp_scale;
    const double u2 = u * u;
    const double phase = omega * time + half_gamma * u2;
    double s, c;
    sincos(phase, &s, &c);
    const double amp = gauss ? exp(-0.5 * u2) : 1.0;
    out_re[m * stride + k] = (T)(amp * c);
    if (out_im) out_im[m * stride + k] = (T)(amp * s);
}

}  // namespace qi

extern "C" {

int qi_synth_chirp(int64_t M, int64_t n, int64_t stride, int64_t k0, double t_center, double omega, double half_gamma,
                   double chirp_scale, int gauss, int dtype, void* out_re, void* out_im, void* stream) {
    if (!out_re || M <= 0 || M > 65535 || n <= 0 || stride < n || !(chirp_scale > 0.0)) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)M);
    qi::prof_set_category(QI_CAT_OTHER);
    if (dtype == QI_F32)
        QI_LAUNCH((qi::synth_chirp_kernel<float>), grid, dim3(256), 0, st, (qi::i64)n, t_center, omega, half_gamma, chirp_scale, gauss,
                  static_cast<float*>(out_re), static_cast<float*>(out_im), (qi::i64)stride, (qi::i64)k0);
    else if (dtype == QI_F64)
        QI_LAUNCH((qi::synth_chirp_kernel<double>), grid, dim3(256), 0, st, (qi::i64)n, t_center, omega, half_gamma, chirp_scale, gauss,
                  static_cast<double*>(out_re), static_cast<double*>(out_im), (qi::i64)stride, (qi::i64)k0);
    else return QI_ERR_ARG;
    return qi::check_cuda("qi_synth_chirp");
}

}  // extern "C"
