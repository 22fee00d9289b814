// qi_iir.cu -- the step BEFORE the time-frequency path: zero-phase IIR filtering of records in HBM.
//
// Replaces scipy.signal.filtfilt(b, a, x) as called by quantum_inferno/styx_fft.py:60-149 (butter_bandpass /
// butter_highpass / butter_lowpass: Tukey taper + Butterworth + filtfilt) and quantum_inferno/synth/
// synthetic_signals.py:180-192 (antialias_half_nyquist), and scipy.signal.sosfiltfilt(sos, x) as called by
// quantum_inferno/utilities/picker.py:56-76 (apply_bandpass).  Same algorithm as scipy's (scipy/signal/
// _signaltools.py::filtfilt / sosfiltfilt, method="pad", padtype="odd"): odd extension by padlen samples, forward
// recursion started from zi * ext[0], recursion over the reversed result started from zi * y[-1], middle part kept.
//
// The recursion itself (scipy _lfilter.c: direct form II transposed; _sosfilt.pyx: cascade of biquads) is a linear
// recurrence  s[p] = A s[p-1] + B x[p]  with a K-dimensional state, evaluated here as a three-kernel blocked scan:
//   1. iir_tile_kernel<false>: a CTA stages a 4096-sample tile (coalesced), every thread runs its 16-sample chunk from
//      the zero state, a Kogge-Stone scan over the 256 chunk end states with the constant matrices A^(16 * 2^j) gives the
//      tile's zero-state end vector;
//   2. iir_scan_kernel: one CTA per record chains the tile vectors (thread-serial runs + the same scan with
//      A^(4096 * R * 2^j)) into the true state at the start of every tile;
//   3. iir_tile_kernel<true>: as 1. with the tile's true initial state injected into chunk 0, then every chunk is
//      re-run from its true initial state and the outputs are stored (coalesced).
// All arithmetic is fp64 whatever the record dtype.  HBM traffic: 48 B per extended sample for the two directions.
#include <math.h>
#include <string.h>
#include <vector>
#include "qi_platform.cuh"
#include "qi_host.h"

namespace qi {

// samples per thread chunk; measured on B200 (order-4 band-pass, 16 x 2^22 fp64): 8 -> 4.75 ms, 16 -> 3.57 ms,
// 32 -> 3.96 ms (make EXTRA=-DQI_IIR_CHUNK=n to repeat)
#ifndef QI_IIR_CHUNK
#define QI_IIR_CHUNK 16
#endif
constexpr int IIR_L = QI_IIR_CHUNK;           // samples per chunk (one thread)
constexpr int IIR_T = 256;                    // chunks per tile (threads per CTA)
constexpr int IIR_TILE = IIR_L * IIR_T;
constexpr int IIR_PITCH = IIR_L + 1;          // tile row pitch in doubles (conflict-free chunk walks)
constexpr int IIR_LEVELS = 8;                 // log2(IIR_T); the record scan also uses 256 threads
constexpr int IIR_NMATS = 2 * IIR_LEVELS + 1; // A^(L 2^j) | A^TILE | (A^(TILE R))^(2^j)

struct IirCoef {
    int form, kp;
    double b[QI_IIR_MAX_STATE + 1], a[QI_IIR_MAX_STATE + 1];
    double sos[QI_IIR_MAX_STATE / 2][6];
    double zi[QI_IIR_MAX_STATE];
};

struct IirIo {
    const void* x;       // record [M, n], f32 or f64
    i64 x_stride, n, n_ext;
    int x_f32, padlen;
    double alpha;        // Tukey taper applied to the record first (< 0: none)
    double* y1;          // [M, n_ext] forward result
    void* out;           // [M, n]
    int out_f32;
};

// one sample of the recursion; z: KP state values (coefficients and sections beyond the filter's own are neutral)
template <int KP, typename S> QI_HD S iir_step(const IirCoef& c, S x, S* z) {
    if (c.form == QI_IIR_BA) {                               // scipy _lfilter.c (direct form II transposed)
        const S y = (S)c.b[0] * x + z[0];
#pragma unroll
        for (int i = 0; i < KP - 1; ++i) z[i] = z[i + 1] + x * (S)c.b[i + 1] - y * (S)c.a[i + 1];
        z[KP - 1] = x * (S)c.b[KP] - y * (S)c.a[KP];
        return y;
    }
#pragma unroll
    for (int s = 0; s < KP / 2; ++s) {                       // scipy _sosfilt.pyx
        const S y = (S)c.sos[s][0] * x + z[2 * s];
        z[2 * s] = (S)c.sos[s][1] * x - (S)c.sos[s][4] * y + z[2 * s + 1];
        z[2 * s + 1] = (S)c.sos[s][2] * x - (S)c.sos[s][5] * y;
        x = y;
    }
    return x;
}

// scipy.signal.windows.tukey(M, alpha, sym=True)[j]
QI_DEV double tukey_weight(i64 j, i64 M, double alpha) {
    if (alpha <= 0.0 || M <= 1) return 1.0;
    if (alpha >= 1.0) return 0.5 - 0.5 * cospi(2.0 * (double)j / (double)(M - 1));          // hann
    const i64 width = (i64)floor(alpha * (double)(M - 1) / 2.0);
    if (j <= width) return 0.5 * (1.0 + cospi(-1.0 + 2.0 * (double)j / alpha / (double)(M - 1)));
    if (j >= M - width - 1) return 0.5 * (1.0 + cospi(-2.0 / alpha + 1.0 + 2.0 * (double)j / alpha / (double)(M - 1)));
    return 1.0;
}

QI_DEV double iir_record(const IirIo& io, i64 m, i64 j) {
    const double v = io.x_f32 ? (double)static_cast<const float*>(io.x)[m * io.x_stride + j]
                              : static_cast<const double*>(io.x)[m * io.x_stride + j];
    return io.alpha >= 0.0 ? v * tukey_weight(j, io.n, io.alpha) : v;
}

// sample p of the sequence a direction filters: the odd extension of the (tapered) record, or the reversed forward result
QI_DEV double iir_input(const IirIo& io, i64 m, i64 p, int backward) {
    if (backward) return io.y1[m * io.n_ext + (io.n_ext - 1 - p)];
    const i64 pl = io.padlen;
    if (p < pl) return 2.0 * iir_record(io, m, 0) - iir_record(io, m, pl - p);
    if (p < pl + io.n) return iir_record(io, m, p - pl);
    return 2.0 * iir_record(io, m, io.n - 1) - iir_record(io, m, io.n - 2 - (p - pl - io.n));
}

// z += P * left over a 256-entry state table; the scan both kernels share.  st: [256][KP + 1] shared doubles.
template <int KP>
QI_DEV void iir_block_scan(double* z, double* st, const double* mats) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int r = 0; r < KP; ++r) st[tid * (KP + 1) + r] = z[r];
    __syncthreads();
    for (int lev = 0; lev < IIR_LEVELS; ++lev) {
        const int d = 1 << lev;
        double add[KP];
#pragma unroll
        for (int r = 0; r < KP; ++r) add[r] = 0.0;
        if (tid >= d) {
            const double* left = st + (tid - d) * (KP + 1);
            const double* P = mats + lev * KP * KP;
#pragma unroll
            for (int cidx = 0; cidx < KP; ++cidx) {
                const double lv = left[cidx];
#pragma unroll
                for (int r = 0; r < KP; ++r) add[r] += P[r * KP + cidx] * lv;
            }
        }
        __syncthreads();
        if (tid >= d) {
#pragma unroll
            for (int r = 0; r < KP; ++r) { z[r] += add[r]; st[tid * (KP + 1) + r] = z[r]; }
        }
        __syncthreads();
    }
}

// grid: (ntiles, M); dynamic shared memory: (IIR_T * IIR_PITCH + IIR_T * (KP + 1) + IIR_LEVELS * KP * KP) doubles.
// The chunk's samples stay in the shared tile (not in registers) and the level matrices are staged in shared memory,
// so that three CTAs share an SM.
template <int KP, bool FINAL>
__global__ void __launch_bounds__(IIR_T, (KP <= 8 ? 3 : 2))
iir_tile_kernel(IirCoef c, IirIo io, int backward, const double* __restrict__ mats, double* __restrict__ agg,
                const double* __restrict__ tile_in, i64 ntiles) {
    QI_DYN_SMEM(raw);
    double* tile = reinterpret_cast<double*>(raw);
    double* st = tile + IIR_T * IIR_PITCH;
    double* smats = st + IIR_T * (KP + 1);
    const int tid = threadIdx.x;
    const i64 m = blockIdx.y, t = blockIdx.x, p0 = t * IIR_TILE;
    for (int i = tid; i < IIR_LEVELS * KP * KP; i += IIR_T) smats[i] = mats[i];
    for (int i = tid; i < IIR_TILE; i += IIR_T) {
        const i64 p = p0 + i;
        tile[(i / IIR_L) * IIR_PITCH + (i % IIR_L)] = p < io.n_ext ? iir_input(io, m, p, backward) : 0.0;
    }
    __syncthreads();
    double* xs = tile + tid * IIR_PITCH;
    double z[KP], s0[KP];
#pragma unroll
    for (int r = 0; r < KP; ++r) {
        s0[r] = (FINAL && tid == 0) ? tile_in[(m * ntiles + t) * KP + r] : 0.0;
        z[r] = s0[r];
    }
#pragma unroll
    for (int i = 0; i < IIR_L; ++i) iir_step<KP, double>(c, xs[i], z);
    iir_block_scan<KP>(z, st, smats);
    if (!FINAL) {
        if (tid == IIR_T - 1) {
#pragma unroll
            for (int r = 0; r < KP; ++r) agg[(m * ntiles + t) * KP + r] = z[r];
        }
        return;
    }
    if (tid > 0) {
#pragma unroll
        for (int r = 0; r < KP; ++r) s0[r] = st[(tid - 1) * (KP + 1) + r];
    }
#pragma unroll
    for (int i = 0; i < IIR_L; ++i) xs[i] = iir_step<KP, double>(c, xs[i], s0);
    __syncthreads();
    for (int i = tid; i < IIR_TILE; i += IIR_T) {
        const i64 p = p0 + i;
        if (p >= io.n_ext) break;
        const double y = tile[(i / IIR_L) * IIR_PITCH + (i % IIR_L)];
        if (!backward) {
            io.y1[m * io.n_ext + p] = y;
        } else {
            const i64 j = io.n_ext - 1 - p - io.padlen;
            if (j >= 0 && j < io.n) {
                if (io.out_f32) static_cast<float*>(io.out)[m * io.n + j] = (float)y;
                else static_cast<double*>(io.out)[m * io.n + j] = y;
            }
        }
    }
}

// grid: (M); 256 threads; thread t chains tiles [t R, (t+1) R).  mats: [0] = A^TILE, [1 + j] = (A^(TILE R))^(2^j)
template <int KP>
__global__ void __launch_bounds__(IIR_T)
iir_scan_kernel(IirCoef c, IirIo io, int backward, const double* __restrict__ mats, const double* __restrict__ agg,
                double* __restrict__ tile_in, i64 ntiles, i64 R) {
    __shared__ double st[IIR_T * (KP + 1)];
    const int tid = threadIdx.x;
    const i64 m = blockIdx.x;
    const double* P = mats;
    const double first = iir_input(io, m, 0, backward);
    double s[KP], s_in[KP];
#pragma unroll
    for (int r = 0; r < KP; ++r) s[r] = tid == 0 ? c.zi[r] * first : 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: end state of this thread's run from zero (thread 0: from zi * first); pass 1: the same walk from the
        // true incoming state, recording the state in front of every tile
        for (i64 k = 0; k < R; ++k) {
            const i64 t = (i64)tid * R + k;
            if (t >= ntiles) break;
            if (pass == 1) {
#pragma unroll
                for (int r = 0; r < KP; ++r) tile_in[(m * ntiles + t) * KP + r] = s[r];
            }
            double nx[KP];
#pragma unroll
            for (int r = 0; r < KP; ++r) nx[r] = agg[(m * ntiles + t) * KP + r];
#pragma unroll
            for (int cidx = 0; cidx < KP; ++cidx) {
#pragma unroll
                for (int r = 0; r < KP; ++r) nx[r] += P[r * KP + cidx] * s[cidx];
            }
#pragma unroll
            for (int r = 0; r < KP; ++r) s[r] = nx[r];
        }
        if (pass == 0) {
            iir_block_scan<KP>(s, st, mats + KP * KP);
#pragma unroll
            for (int r = 0; r < KP; ++r) s_in[r] = tid == 0 ? c.zi[r] * first : st[(tid - 1) * (KP + 1) + r];
#pragma unroll
            for (int r = 0; r < KP; ++r) s[r] = s_in[r];
        }
    }
}

// ---------------------------------------------------------------- host: powers of the transition matrix
// Column j of A^m is the state after m zero-input steps from the unit state e_j.  The powers the scans need are read
// off one long-double run of the recursion itself (backward stable; repeated squaring of these non-normal matrices
// is not: its error squares at every level).  Once every entry is below 1e-60 all later powers are taken as zero.
typedef long double ld;
template <int KP>
static void iir_powers(const IirCoef& c, const std::vector<unsigned long long>& wanted, double* out /*[wanted][KP][KP]*/) {
    std::vector<ld> col((size_t)KP * KP, 0.0L);               // col[j*KP + r] = (A^m)[r][j]
    for (int j = 0; j < KP; ++j) col[(size_t)j * KP + j] = 1.0L;
    unsigned long long m = 0;
    bool dead = false;
    for (size_t w = 0; w < wanted.size(); ++w) {              // wanted is ascending
        while (m < wanted[w] && !dead) {
            ld big = 0.0L;
            for (int j = 0; j < KP; ++j) {
                iir_step<KP, ld>(c, 0.0L, &col[(size_t)j * KP]);
                for (int r = 0; r < KP; ++r) { const ld v = fabsl(col[(size_t)j * KP + r]); big = v > big ? v : big; }
            }
            ++m;
            dead = big < 1e-60L;
        }
        for (int r = 0; r < KP; ++r)
            for (int j = 0; j < KP; ++j) out[(w * KP + r) * KP + j] = dead ? 0.0 : (double)col[(size_t)j * KP + r];
    }
}

template <int KP>
static int filtfilt_impl(const IirCoef& c, IirIo io, i64 M, void* ws, size_t ws_bytes, cudaStream_t st) {
    const i64 ntiles = (io.n_ext + IIR_TILE - 1) / IIR_TILE;
    const i64 R = (ntiles + IIR_T - 1) / IIR_T;
    const size_t mats_bytes = sizeof(double) * IIR_NMATS * KP * KP;
    const size_t vec_bytes = sizeof(double) * (size_t)M * (size_t)ntiles * KP;
    const size_t need = mats_bytes + 2 * vec_bytes + sizeof(double) * (size_t)M * (size_t)io.n_ext;
    if (ws_bytes < need) return QI_ERR_WORKSPACE;
    std::vector<unsigned long long> wanted;                      // A^(L 2^lev) lev = 0..8 (8: A^TILE), then A^(TILE R 2^lev)
    for (int lev = 0; lev <= IIR_LEVELS; ++lev) wanted.push_back((unsigned long long)IIR_L << lev);
    for (int lev = 0; lev < IIR_LEVELS; ++lev) wanted.push_back(((unsigned long long)IIR_TILE * (unsigned long long)R) << lev);
    std::vector<double> host((size_t)IIR_NMATS * KP * KP);
    iir_powers<KP>(c, wanted, host.data());
    unsigned char* base = static_cast<unsigned char*>(ws);
    double* d_mats = reinterpret_cast<double*>(base);
    double* d_agg = reinterpret_cast<double*>(base + mats_bytes);
    double* d_in = reinterpret_cast<double*>(base + mats_bytes + vec_bytes);
    io.y1 = reinterpret_cast<double*>(base + mats_bytes + 2 * vec_bytes);
    stage_to_device(d_mats, host.data(), mats_bytes, st);
    const size_t smem = sizeof(double) * (IIR_T * IIR_PITCH + IIR_T * (KP + 1) + IIR_LEVELS * KP * KP);
#ifndef QI_EMUL
    cudaFuncSetAttribute(iir_tile_kernel<KP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(iir_tile_kernel<KP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    prof_set_category(QI_CAT_OTHER);
    dim3 grid((unsigned)ntiles, (unsigned)M);
    for (int backward = 0; backward < 2; ++backward) {
        QI_LAUNCH((iir_tile_kernel<KP, false>), grid, dim3(IIR_T), smem, st, c, io, backward, (const double*)d_mats, d_agg,
                  (const double*)d_in, ntiles);
        QI_LAUNCH((iir_scan_kernel<KP>), dim3((unsigned)M), dim3(IIR_T), 0, st, c, io, backward,
                  (const double*)(d_mats + IIR_LEVELS * KP * KP), (const double*)d_agg, d_in, ntiles, R);
        QI_LAUNCH((iir_tile_kernel<KP, true>), grid, dim3(IIR_T), smem, st, c, io, backward, (const double*)d_mats, d_agg,
                  (const double*)d_in, ntiles);
    }
    return check_cuda("qi_filtfilt");
}

static int iir_kp(int n_state) { return n_state <= 4 ? 4 : (n_state <= 8 ? 8 : 16); }

}  // namespace qi

extern "C" {

size_t qi_filtfilt_workspace_bytes(int64_t M, int64_t n, int padlen, int n_state) {
    if (M <= 0 || n <= 0 || padlen < 0 || n_state < 1 || n_state > QI_IIR_MAX_STATE) return 0;
    const int kp = qi::iir_kp(n_state);
    const int64_t n_ext = n + 2 * (int64_t)padlen;
    const int64_t ntiles = (n_ext + qi::IIR_TILE - 1) / qi::IIR_TILE;
    return sizeof(double) * ((size_t)qi::IIR_NMATS * kp * kp + 2 * (size_t)M * ntiles * kp + (size_t)M * n_ext) + 256;
}

int qi_filtfilt(const void* sig, int64_t M, int64_t n, int64_t stride, const QiIirFilter* f, int padlen, double tukey_alpha,
                int dtype, void* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!sig || !out || !f || !workspace || M <= 0 || M > 65535 || n <= 0 || stride < n || padlen < 0) return QI_ERR_ARG;
    if (n <= padlen) return QI_ERR_ARG;                       // scipy: "The length of the input vector x must be greater than padlen"
    if (dtype != QI_F32 && dtype != QI_F64) return QI_ERR_ARG;
    qi::IirCoef c;
    memset(&c, 0, sizeof(c));
    int n_state;
    if (f->form == QI_IIR_BA) {
        if (f->n_coef < 2 || f->n_coef > QI_IIR_MAX_STATE + 1) return QI_ERR_ARG;
        n_state = f->n_coef - 1;
        for (int i = 0; i < f->n_coef; ++i) { c.b[i] = f->b[i]; c.a[i] = f->a[i]; }
        if (c.a[0] != 1.0) return QI_ERR_ARG;                 // the caller normalises by a[0], as scipy does
    } else if (f->form == QI_IIR_SOS) {
        if (f->n_coef < 1 || f->n_coef > QI_IIR_MAX_STATE / 2) return QI_ERR_ARG;
        n_state = 2 * f->n_coef;
        for (int s = 0; s < QI_IIR_MAX_STATE / 2; ++s) {
            if (s < f->n_coef) {
                for (int k = 0; k < 6; ++k) c.sos[s][k] = f->sos[s][k];
                if (c.sos[s][3] != 1.0) return QI_ERR_ARG;
            } else {
                c.sos[s][0] = 1.0;                            // neutral section: y = x
                c.sos[s][3] = 1.0;
            }
        }
    } else return QI_ERR_ARG;
    c.form = f->form;
    c.kp = qi::iir_kp(n_state);
    for (int i = 0; i < n_state; ++i) c.zi[i] = f->zi[i];
    qi::IirIo io;
    memset(&io, 0, sizeof(io));
    io.x = sig; io.x_stride = stride; io.n = n; io.padlen = padlen; io.n_ext = n + 2 * (int64_t)padlen;
    io.x_f32 = dtype == QI_F32; io.out_f32 = dtype == QI_F32; io.alpha = tukey_alpha; io.out = out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c.kp == 4) return qi::filtfilt_impl<4>(c, io, M, workspace, workspace_bytes, st);
    if (c.kp == 8) return qi::filtfilt_impl<8>(c, io, M, workspace, workspace_bytes, st);
    return qi::filtfilt_impl<16>(c, io, M, workspace, workspace_bytes, st);
}

}  // extern "C"
