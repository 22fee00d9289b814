/* qi_b200.h -- C ABI of libqi_b200.so: the B200 (sm_100a) kernels behind the quantum-inferno
 * time-frequency hot path.
 *
 * The reference (ISLA-UH/quantum-inferno v1.1.3) has no FFI layer: its boundary is the
 * module-level Python API.  Each entry point below names the reference function whose
 * arithmetic it replaces (file:line relative to the reference checkout); the Python modules
 * in quantum_inferno_b200/ keep the reference's names and signatures and call these through
 * ctypes.  All pointers except those marked HOST are device pointers (HBM); no entry point
 * allocates; all work is enqueued on `stream` (a cudaStream_t passed as void*); every call
 * returns 0 on success or a negative QI_ERR_* code and never throws.
 */
#ifndef QI_B200_H
#define QI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QI_ABI_VERSION 1

/* dtype of the arithmetic and of every real/complex buffer in a call */
#define QI_F32 0
#define QI_F64 1

#define QI_OK 0
#define QI_ERR_ARG (-1)       /* bad size / pointer / dtype                     -> ValueError   */
#define QI_ERR_WORKSPACE (-2) /* workspace too small                            -> ValueError   */
#define QI_ERR_CUDA (-3)      /* a CUDA runtime error was raised                -> RuntimeError */
#define QI_ERR_UNSUPPORTED (-4)

int qi_abi_version(void);
const char* qi_error_string(int code);
/* text of the last CUDA error seen by this thread (empty string if none) */
const char* qi_last_cuda_error(void);

/* ---- plain batched FFT (building block; also used by tfr_info.ShannonFFT, tfr_info.py:177) ----
 * in : complex [batch, 2^log2n] natural order
 * out: forward -> spectrum in BIT-REVERSED order (position p holds bin bitrev(p));
 *      inverse -> takes bit-reversed order, returns natural order scaled by 1/n.
 * in == out is allowed. */
int qi_fft_c2c(const void* in, void* out, int64_t batch, int log2n, int inverse, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QI_B200_H */
