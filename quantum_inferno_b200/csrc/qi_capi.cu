// qi_capi.cu -- extern "C" boundary of libqi_b200.so (see include/qi_b200.h).
#include "qi_fft.cuh"
#include "qi_host.h"

namespace qi {

thread_local char g_last_cuda_error[256] = {0};

int check_cuda(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", where, cudaGetErrorString(e));
        return QI_ERR_CUDA;
    }
    return QI_OK;
}

template <typename T>
static int fft_c2c_impl(const void* in, void* out, i64 batch, int log2n, int inverse, cudaStream_t st) {
    const FftPlan plan = make_plan(log2n, (int)sizeof(cplx<T>));
    const i64 n = 1ll << log2n;
    const cplx<T>* src0 = static_cast<const cplx<T>*>(in);
    cplx<T>* dstp = static_cast<cplx<T>*>(out);
    if (!inverse) {
        for (int p = 0; p < plan.npass; ++p) {
            SrcComplex<T> s{p == 0 ? src0 : dstp, n};
            DstComplex<T> d{dstp, n, (T)1};
            launch_pass<T, FFT_FWD>(plan, p, batch, s, d, 0, st);
        }
    } else {
        for (int p = plan.npass - 1; p >= 0; --p) {
            SrcComplex<T> s{p == plan.npass - 1 ? src0 : dstp, n};
            DstComplex<T> d{dstp, n, p == 0 ? (T)(1.0 / (double)n) : (T)1};
            launch_pass<T, FFT_INV>(plan, p, batch, s, d, 0, st);
        }
    }
    return check_cuda("qi_fft_c2c");
}

}  // namespace qi

extern "C" {

int qi_abi_version(void) { return QI_ABI_VERSION; }

const char* qi_error_string(int code) {
    switch (code) {
        case QI_OK: return "ok";
        case QI_ERR_ARG: return "invalid argument";
        case QI_ERR_WORKSPACE: return "workspace too small";
        case QI_ERR_CUDA: return "CUDA error";
        case QI_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

const char* qi_last_cuda_error(void) { return qi::g_last_cuda_error; }

int qi_fft_c2c(const void* in, void* out, int64_t batch, int log2n, int inverse, int dtype, void* stream) {
    if (!in || !out || batch <= 0 || log2n < 0 || log2n > 30) return QI_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == QI_F32) return qi::fft_c2c_impl<float>(in, out, batch, log2n, inverse, st);
    if (dtype == QI_F64) return qi::fft_c2c_impl<double>(in, out, batch, log2n, inverse, st);
    return QI_ERR_ARG;
}

}  // extern "C"
