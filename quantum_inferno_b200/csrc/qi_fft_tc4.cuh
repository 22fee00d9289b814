// qi_fft_tc4.cuh -- compile-time specialisation of the shared-memory FFT tile core for the overlap-save kernel of
// the multirate path: float, 2^LOGR rows x 4 columns, no padding, XOR-swizzled so that EVERY access pattern the
// kernel uses is bank-conflict free:
//
//     slot(r, c) = 4 * (r ^ s) + (c ^ s),   s = (r >> 2) & 3          (8-byte complex slots, 16 slots = 32 banks)
//
//   * butterfly stages with sub-stride >= 4: a half-warp touches 4 consecutive rows x 4 columns -> 16 distinct slots
//   * the final radix-4 stage (rows 4g+i, lanes over g and c): the row swizzle spreads the four g's
//   * row-wise staging (lanes along r, fixed c): the column swizzle spreads rows 4a+b over a
// (the generic [R][TC+1] layout of qi_fft.cuh is 2-way conflicted for TC = 4 in the butterfly stages; ncu showed
// ~2 extra wavefronts per shared-memory instruction in mr_level_kernel).
#pragma once
#include "qi_fft.cuh"

namespace qi {

QI_DEV int sw4(int r, int c) {
    const int s = (r >> 2) & 3;
    return ((r ^ s) << 2) | (c ^ s);
}

template <int DIR, int STEP, int LOGR, int LOGB>
QI_DEV void tc4_stage(cplx<float>* __restrict__ tile, const cplx<float>* __restrict__ tw) {
    constexpr int Q = 1 << STEP;
    constexpr int LOGH = LOGB - STEP;
    constexpr int H = 1 << LOGH;
    constexpr int NTASK = (1 << (LOGR - STEP)) * 4;
    constexpr int TWSHIFT = LOGR - LOGB;
    for (int task = threadIdx.x; task < NTASK; task += blockDim.x) {
        const int c = task & 3;
        const int u = task >> 2;
        const int j = u & (H - 1);
        const int base = ((u >> LOGH) << LOGB) + j;
        // slot of element i of this butterfly (row base + i*H), with the swizzle folded per stage geometry:
        //   H >= 16 : i*H does not touch the 4 low row bits -> one swizzle, compile-time offsets
        //   H == 1  : rows 4g + i share s = g & 3
        //   else    : generic
        int slot[Q];
        if constexpr (H >= 16) {
            const int s0 = sw4(base, c);
#pragma unroll
            for (int i = 0; i < Q; ++i) slot[i] = s0 + i * (H << 2);
        } else if constexpr (H == 1 && Q == 4) {
            const int s = (base >> 2) & 3;
            const int b4 = ((base >> 2) << 4) | (c ^ s);
#pragma unroll
            for (int i = 0; i < Q; ++i) slot[i] = b4 + ((i ^ s) << 2);
        } else {
#pragma unroll
            for (int i = 0; i < Q; ++i) slot[i] = sw4(base + i * H, c);
        }
        cplx<float> a[Q];
        if (DIR == FFT_FWD) {
#pragma unroll
            for (int i = 0; i < Q; ++i) a[i] = tile[slot[i]];
            if (STEP == 3) dif8<float, DIR>(a); else if (STEP == 2) dif4<float, DIR>(a); else dif2<float, DIR>(a);
#pragma unroll
            for (int s = 0; s < Q; ++s) {
                const int f = STEP == 3 ? brev3(s) : (STEP == 2 ? brev2(s) : s);
                cplx<float> v = a[s];
                if (f != 0) v = v * tw[(j * f) << TWSHIFT];
                tile[slot[s]] = v;
            }
        } else {
#pragma unroll
            for (int s = 0; s < Q; ++s) {
                const int f = STEP == 3 ? brev3(s) : (STEP == 2 ? brev2(s) : s);
                cplx<float> v = tile[slot[s]];
                if (f != 0) v = mul_conj(v, tw[(j * f) << TWSHIFT]);
                a[s] = v;
            }
            if (STEP == 3) dit8<float, DIR>(a); else if (STEP == 2) dit4<float, DIR>(a); else dit2<float, DIR>(a);
#pragma unroll
            for (int i = 0; i < Q; ++i) tile[slot[i]] = a[i];
        }
    }
}

template <int LOGR, int LOGB>
QI_DEV void tc4_fwd_from(cplx<float>* tile, const cplx<float>* tw) {
    if constexpr (LOGB >= 3 && LOGB - 3 >= LOGR % 3) {
        tc4_stage<FFT_FWD, 3, LOGR, LOGB>(tile, tw);
        __syncthreads();
        tc4_fwd_from<LOGR, LOGB - 3>(tile, tw);
    } else if constexpr (LOGB == 2) {
        tc4_stage<FFT_FWD, 2, LOGR, 2>(tile, tw);
        __syncthreads();
    } else if constexpr (LOGB == 1) {
        tc4_stage<FFT_FWD, 1, LOGR, 1>(tile, tw);
        __syncthreads();
    }
}

template <int LOGR, int LOGB>
QI_DEV void tc4_inv_from(cplx<float>* tile, const cplx<float>* tw) {
    if constexpr (LOGB < LOGR) {
        tc4_stage<FFT_INV, 3, LOGR, LOGB + 3>(tile, tw);
        __syncthreads();
        tc4_inv_from<LOGR, LOGB + 3>(tile, tw);
    }
}

// All threads of the CTA must call these; they end with a __syncthreads().
template <int LOGR> QI_DEV void tc4_fft_fwd(cplx<float>* tile, const cplx<float>* tw) { tc4_fwd_from<LOGR, LOGR>(tile, tw); }

template <int LOGR> QI_DEV void tc4_fft_inv(cplx<float>* tile, const cplx<float>* tw) {
    constexpr int REM = LOGR % 3;
    if constexpr (REM == 2) { tc4_stage<FFT_INV, 2, LOGR, 2>(tile, tw); __syncthreads(); }
    else if constexpr (REM == 1) { tc4_stage<FFT_INV, 1, LOGR, 1>(tile, tw); __syncthreads(); }
    tc4_inv_from<LOGR, REM>(tile, tw);
}

}  // namespace qi
