// qi_mr_level2k.cuh -- "A" of the multirate CWT (see qi_mr.cu): overlap-save convolution of one level signal with all
// bands of that level, 2048-point blocks, written for the lowest instruction count per output sample.
//
//   * A CTA of 256 threads convolves PAIRS of consecutive blocks: the two blocks are the two complex columns of a
//     float4 row, so every shared-memory access is one 128-bit instruction and every index / twiddle computation is
//     shared by two transforms.  The tile is [2048 rows] x 16 B with one padding row per 8 rows
//     (phys(r) = r + (r >> 3)): all four stage geometries and the staging pattern are bank-conflict free and every
//     address is "one base register + immediate".
//   * radix 8-8-8-4, one radix-8 butterfly (two columns) per thread and stage.  Forward = DIF (natural -> bit-reversed),
//     inverse = DIT (bit-reversed -> natural), as everywhere in this library (qi_fft.cuh).
//   * forward: the first stage reads the real level signal straight from HBM (coalesced), the last (radix-4) stage
//     leaves the block spectra X in REGISTERS (rows 4g .. 4g+3 of thread g and g+256).
//   * per band: the first inverse stage multiplies those registers with the band's kernel table (two 128-bit loads per
//     four rows, L1/L2 resident) and runs its radix-4 butterfly before anything touches shared memory; the last
//     inverse stage hands its natural-order outputs from registers to HBM (level 0: |.|^2 -> power plane; deeper
//     levels: the decimated complex output w_b).  Two tile buffers alternate between bands -> 3 barriers per band.
//     band_sum (optional) receives sum |y|^2 over the record's own samples: the exact band power at level 0, the raw
//     sum that mr_total_kernel turns into the band-power estimate at deeper levels.
//   * twiddles come from a per-stage table laid out [slot pair][j] (built once per call by mr_twiddle2k_kernel with
//     the exact sincospi roots, copied to shared memory per CTA): 4 conflict-free 128-bit loads per butterfly.
#pragma once
#include "qi_fft.cuh"
#include "qi_mr_expand.cuh"
#include "qi_reduce.cuh"

namespace qi {

constexpr int L2K_LOGF = 11;
constexpr int L2K_F = 1 << L2K_LOGF;
constexpr int L2K_TILE = L2K_F + L2K_F / 8;          // padded rows per tile buffer
// twiddle rows: 2048-point stages B=2048 (j<256), B=256 (j<32), B=32 (j<4); 1024-point stages of the envelope bands
// B=1024 (j<128), B=128 (j<16), B=16 (j<2); then 16 rows holding the 64 demodulation phases exp(-2 pi i m / 64)
constexpr int L2K_TW0 = 0, L2K_TW1 = 256, L2K_TW2 = 288;
constexpr int L2K_TWE0 = 292, L2K_TWE1 = 420, L2K_TWE2 = 436;
constexpr int L2K_TWJ = 438;
constexpr int L2K_DEMOD = 4 * L2K_TWJ;               // float4 index of the demodulation table (16 float4 = 64 phases / 2 ... see kernel)
constexpr int L2K_TW_TOTAL = 4 * L2K_TWJ + 32;       // float4 elements in the whole table
constexpr int L2K_MAXB = 64;                         // bands per level handled by this kernel
constexpr int L2K_THREADS = 256;
constexpr size_t L2K_SMEM = (size_t)(2 * L2K_TILE + L2K_TW_TOTAL) * 16 + (size_t)(L2K_THREADS / 32) * L2K_MAXB * 4;

// tw[m * L2K_TWJ + row] = ( w_B^(j * brev3(2m)), w_B^(j * brev3(2m+1)) ),  w_B = exp(-2 pi i / B)
__global__ void mr_twiddle2k_kernel(float4* __restrict__ tw) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L2K_TW_TOTAL) return;
    if (idx >= L2K_DEMOD) {                                    // two consecutive phases per float4
        const int m = 2 * (idx - L2K_DEMOD);
        const cplx<float> w0 = conj(unit_root<float>((unsigned long long)m, 6));
        const cplx<float> w1 = conj(unit_root<float>((unsigned long long)(m + 1), 6));
        tw[idx] = make_float4(w0.re, w0.im, w1.re, w1.im);
        return;
    }
    const int m = idx / L2K_TWJ, row = idx % L2K_TWJ;
    int logB, j;
    if (row < L2K_TW1) { logB = 11; j = row; }
    else if (row < L2K_TW2) { logB = 8; j = row - L2K_TW1; }
    else if (row < L2K_TWE0) { logB = 5; j = row - L2K_TW2; }
    else if (row < L2K_TWE1) { logB = 10; j = row - L2K_TWE0; }
    else if (row < L2K_TWE2) { logB = 7; j = row - L2K_TWE1; }
    else { logB = 4; j = row - L2K_TWE2; }
    const cplx<float> w0 = conj(unit_root<float>((unsigned long long)(j * brev3(2 * m)), logB));
    const cplx<float> w1 = conj(unit_root<float>((unsigned long long)(j * brev3(2 * m + 1)), logB));
    tw[idx] = make_float4(w0.re, w0.im, w1.re, w1.im);
}

struct MrLevelGeom {
    int level, logF, TC;
    int band_first, band_count, n_bands;
    int wk;                 // two-sided kernel support reserved per block (even); valid outputs per block V = F - wk
    i64 n_points;           // N
    i64 n_level;            // N >> level
    i64 q_first;            // first output index (level rate): -MR_HALO for level >= 1, 0 for level 0
    i64 n_out;              // outputs per channel at this level
    i64 n_blocks;
    i64 x_stride, x_len;    // level signal: per-channel stride, stored length
    int x_halo;             // halo of the stored level signal (0 for level 0)
    int env;                // every band of this level is stored as a demodulated envelope at level + 1
    int no_sums;            // edge source bands (indices >= n_bands): no row of their own in the sum arrays
};

// radix-8 stage on the two columns of a float4 tile; rows base + i*H live at tile[p0 + i*STRIDE]
template <int DIR, int STRIDE>
QI_DEV void l2k_stage8(float4* __restrict__ tile, int p0, const float4* __restrict__ tw, int twrow) {
    cplx<float> a[8], b[8];
    float4 w[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) w[m] = tw[m * L2K_TWJ + twrow];
    if (DIR == FFT_FWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = tile[p0 + i * STRIDE];
            a[i] = mk<float>(v.x, v.y); b[i] = mk<float>(v.z, v.w);
        }
        dif8<float, DIR>(a); dif8<float, DIR>(b);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            cplx<float> va = a[s], vb = b[s];
            if (s) {
                const float4 ww = w[s >> 1];
                const cplx<float> t = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                va = va * t; vb = vb * t;
            }
            tile[p0 + s * STRIDE] = make_float4(va.re, va.im, vb.re, vb.im);
        }
    } else {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float4 v = tile[p0 + s * STRIDE];
            cplx<float> va = mk<float>(v.x, v.y), vb = mk<float>(v.z, v.w);
            if (s) {
                const float4 ww = w[s >> 1];
                const cplx<float> t = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                va = mul_conj(va, t); vb = mul_conj(vb, t);
            }
            a[s] = va; b[s] = vb;
        }
        dit8<float, DIR>(a); dit8<float, DIR>(b);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[p0 + i * STRIDE] = make_float4(a[i].re, a[i].im, b[i].re, b[i].im);
    }
}

// bx: index of this CTA among the CTAs of its level (blockIdx.x, or the offset inside a merged multi-level launch)
// ENV = false compiles the envelope path out (the level-0 launch: fewer registers)
template <bool ENV>
QI_DEV void l2k_body(const float* __restrict__ x, const MrLevelGeom& g, const MrDevBand* __restrict__ bands,
                     const cplx<float>* __restrict__ tables, const float4* __restrict__ tw_g,
                     cplx<float>* __restrict__ wbuf, float* __restrict__ out_power, cplx<float>* __restrict__ out_complex,
                     double* __restrict__ band_sum, int pairs_per_cta, int bx) {
    QI_DYN_SMEM(smem_raw);
    float4* tile0 = reinterpret_cast<float4*>(smem_raw);
    float4* tile1 = tile0 + L2K_TILE;
    float4* tw = tile1 + L2K_TILE;
    float* wsum = reinterpret_cast<float*>(tw + L2K_TW_TOTAL);         // [warps][L2K_MAXB]
    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const i64 chan = blockIdx.y;
    const int V = L2K_F - g.wk;
    const int half = g.wk / 2;
    const float* xs = x + chan * g.x_stride;

    for (int i = t; i < L2K_TW_TOTAL; i += L2K_THREADS) tw[i] = tw_g[i];
    for (int i = t; i < (L2K_THREADS / 32) * L2K_MAXB; i += L2K_THREADS) wsum[i] = 0.0f;

    // per-thread tile addresses (padded rows), fixed for the whole kernel
    const int pA = t + (t >> 3);                                           // stage B=2048: rows t + 256 i   -> + 288 i
    const int rB = ((t >> 5) << 8) + (t & 31);                             // stage B=256 : rows rB + 32 i   -> + 36 i
    const int pB = rB + (rB >> 3);
    const int pC = (t >> 2) * 36 + (t & 3);                                // stage B=32  : rows 32G + j + 4i -> + 4i + (i>>1)
    const int pD0 = 4 * t + (t >> 1);                                      // radix 4     : rows 4g + i, g = t
    const int pD1 = 4 * (t + 256) + ((t + 256) >> 1);                      //               g = t + 256

    // twiddles of the stage B=2048 depend on the thread only: kept in registers for the forward stage and every band
    // (the kernel is bound by the shared-memory / L1 data pipe: 73 % of its wavefront rate, ncu r3c)
    __syncthreads();
    float4 wA[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) wA[m] = tw[m * L2K_TWJ + L2K_TW0 + t];

    for (int pp = 0; pp < pairs_per_cta; ++pp) {
        const i64 blk0 = 2 * ((i64)bx * pairs_per_cta + pp);
        if (blk0 >= g.n_blocks) break;                                     // uniform over the CTA
        __syncthreads();                                                   // previous pair's last stage has left the tiles
        // ---- forward, stage B=2048 straight from HBM (real input, two blocks)
        {
            cplx<float> a[8], b[8];
            const i64 k0 = g.q_first + blk0 * V - half + g.x_halo + t;
            const bool has_b = blk0 + 1 < g.n_blocks;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const i64 ka = k0 + 256 * i, kb = ka + V;
                a[i] = mk<float>((ka >= 0 && ka < g.x_len) ? xs[ka] : 0.0f, 0.0f);
                b[i] = mk<float>((has_b && kb >= 0 && kb < g.x_len) ? xs[kb] : 0.0f, 0.0f);
            }
            dif8<float, FFT_FWD>(a); dif8<float, FFT_FWD>(b);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                cplx<float> va = a[s], vb = b[s];
                if (s) {
                    const float4 ww = wA[s >> 1];
                    const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                    va = va * tt; vb = vb * tt;
                }
                tile0[pA + s * 288] = make_float4(va.re, va.im, vb.re, vb.im);
            }
        }
        __syncthreads();
        l2k_stage8<FFT_FWD, 36>(tile0, pB, tw, L2K_TW1 + (t & 31));
        __syncthreads();
        {   // stage B=32: rows 32G + j + 4i -> pC + 4i + (i >> 1): not a constant stride, spelled out
            cplx<float> a[8], b[8];
            float4 w[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) w[m] = tw[m * L2K_TWJ + L2K_TW2 + (t & 3)];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = tile0[pC + 4 * i + (i >> 1)];
                a[i] = mk<float>(v.x, v.y); b[i] = mk<float>(v.z, v.w);
            }
            dif8<float, FFT_FWD>(a); dif8<float, FFT_FWD>(b);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                cplx<float> va = a[s], vb = b[s];
                if (s) {
                    const float4 ww = w[s >> 1];
                    const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                    va = va * tt; vb = vb * tt;
                }
                tile0[pC + 4 * s + (s >> 1)] = make_float4(va.re, va.im, vb.re, vb.im);
            }
        }
        __syncthreads();
        // ---- forward radix-4 stage: spectra stay in registers (rows 4g .. 4g+3, g = t and t + 256)
        cplx<float> X0[2][4], X1[2][4];
#pragma unroll
        for (int task = 0; task < 2; ++task) {
            const int p = task ? pD1 : pD0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = tile0[p + i];
                X0[task][i] = mk<float>(v.x, v.y); X1[task][i] = mk<float>(v.z, v.w);
            }
            dif4<float, FFT_FWD>(X0[task]); dif4<float, FFT_FWD>(X1[task]);
        }

        // ---- envelope bands: two bands per pass, half-size inverse transforms (see mr_plan)
        //   fold   Z[k'] = Y[k'] + Y[k' + 1024]  <->  rows (2q, 2q+1) of the bit-reversed 2048-row spectrum add up to
        //          row q of the bit-reversed 1024-row one: y(2m) = IFFT_1024(Z)(m)
        //   the first (radix-2) stage of that transform joins rows (2g, 2g+1): still inside the thread that holds rows
        //   4g .. 4g+3 of Y; then radix 8-8-8 with 128 butterflies per band, i.e. 256 = both bands of the pass
        for (int pi = 0; ENV && g.env && pi < g.band_count; pi += 2) {
            const int nbp = g.band_count - pi < 2 ? g.band_count - pi : 2;
            float4* tile = ((pi >> 1) & 1) ? tile0 : tile1;
#pragma unroll
            for (int task = 0; task < 2; ++task) {
                const int gq = t + 256 * task;
                const int p = 2 * gq + (gq >> 2);                          // padded row of 2 gq in a 1024-row tile
                for (int bb = 0; bb < nbp; ++bb) {
                    const float4* K = reinterpret_cast<const float4*>(tables + bands[g.band_first + pi + bb].table_off);
                    const float4 k01 = K[2 * gq], k23 = K[2 * gq + 1];
                    const cplx<float> kk[4] = {mk<float>(k01.x, k01.y), mk<float>(k01.z, k01.w), mk<float>(k23.x, k23.y),
                                               mk<float>(k23.z, k23.w)};
                    const cplx<float> za0 = X0[task][0] * kk[0] + X0[task][1] * kk[1];
                    const cplx<float> za1 = X0[task][2] * kk[2] + X0[task][3] * kk[3];
                    const cplx<float> zc0 = X1[task][0] * kk[0] + X1[task][1] * kk[1];
                    const cplx<float> zc1 = X1[task][2] * kk[2] + X1[task][3] * kk[3];
                    const cplx<float> a0 = za0 + za1, a1 = za0 - za1, c0 = zc0 + zc1, c1 = zc0 - zc1;
                    tile[bb * (L2K_TILE / 2) + p] = make_float4(a0.re, a0.im, c0.re, c0.im);
                    tile[bb * (L2K_TILE / 2) + p + 1] = make_float4(a1.re, a1.im, c1.re, c1.im);
                }
            }
            __syncthreads();
            const int bb = t >> 7, u = t & 127;                            // band of the pass / butterfly of that band
            const bool live = bb < nbp;
            float4* tb = tile + bb * (L2K_TILE / 2);
            if (live) {   // stage B=16: rows 16G + j + 2i -> 18G + j + 2i + (i >> 2)
                const int p0 = (u >> 1) * 18 + (u & 1);
                cplx<float> a[8], c[8];
                float4 w[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) w[m] = tw[m * L2K_TWJ + L2K_TWE2 + (u & 1)];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float4 v = tb[p0 + 2 * s + (s >> 2)];
                    cplx<float> va = mk<float>(v.x, v.y), vb = mk<float>(v.z, v.w);
                    if (s) {
                        const float4 ww = w[s >> 1];
                        const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                        va = mul_conj(va, tt); vb = mul_conj(vb, tt);
                    }
                    a[s] = va; c[s] = vb;
                }
                dit8<float, FFT_INV>(a); dit8<float, FFT_INV>(c);
#pragma unroll
                for (int i = 0; i < 8; ++i) tb[p0 + 2 * i + (i >> 2)] = make_float4(a[i].re, a[i].im, c[i].re, c[i].im);
            }
            __syncthreads();
            if (live) {   // stage B=128: rows 128G + j + 16i -> 144G + j + (j >> 3) + 18i
                const int j = u & 15;
                l2k_stage8<FFT_INV, 18>(tb, (u >> 4) * 144 + j + (j >> 3), tw, L2K_TWE1 + j);
            }
            __syncthreads();
            if (live) {   // stage B=1024: rows j + 128i -> j + (j >> 3) + 144i; outputs m = u + 128 i leave from registers
                cplx<float> a[8], c[8];
                float4 w[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) w[m] = tw[m * L2K_TWJ + L2K_TWE0 + u];
                const int p0 = u + (u >> 3);
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float4 v = tb[p0 + 144 * s];
                    cplx<float> va = mk<float>(v.x, v.y), vb = mk<float>(v.z, v.w);
                    if (s) {
                        const float4 ww = w[s >> 1];
                        const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                        va = mul_conj(va, tt); vb = mul_conj(vb, tt);
                    }
                    a[s] = va; c[s] = vb;
                }
                dit8<float, FFT_INV>(a); dit8<float, FFT_INV>(c);
                const int b = g.band_first + pi + bb;
                const MrDevBand band = bands[b];
                const cplx<float>* dm = reinterpret_cast<const cplx<float>*>(tw + L2K_DEMOD);   // exp(-2 pi i m / 64)
                const int Vh = V >> 1, hh = half >> 1;
                const i64 n_w = (g.n_level >> 1) + 2 * MR_HALO;              // stored samples of w at level + 1
                float acc = 0.0f;
#pragma unroll
                for (int col = 0; col < 2; ++col) {
                    const i64 blk = blk0 + col;
                    if (blk >= g.n_blocks) break;
                    const cplx<float>* y = col ? c : a;
                    cplx<float>* wdst = wbuf + band.w_off + chan * band.w_stride;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int mm = u + 128 * i - hh;
                        const i64 wi = blk * Vh + mm;                       // index into w (halo included): q' = wi - HALO
                        if (mm >= 0 && mm < Vh && wi < n_w) {
                            const i64 q = wi - MR_HALO;
                            const cplx<float> e = y[i] * dm[(int)((band.demod * q) & 63)];
                            wdst[wi] = e;
                            if (q >= 0 && q < (g.n_level >> 1)) acc += norm2(e);
                        }
                    }
                }
                if (band_sum) {
                    acc = warp_sum(acc);
                    if (lane == 0) wsum[warp * L2K_MAXB + pi + bb] += acc;
                }
            }
        }

        // ---- bands
        for (int bi = 0; !(ENV && g.env) && bi < g.band_count; ++bi) {
            const int b = g.band_first + bi;
            const MrDevBand band = bands[b];
            const float4* K = reinterpret_cast<const float4*>(tables + band.table_off);
            float4* tile = (bi & 1) ? tile0 : tile1;
            // inverse radix-4 stage on X * K, from registers
#pragma unroll
            for (int task = 0; task < 2; ++task) {
                const int gq = t + 256 * task;
                const float4 k01 = K[2 * gq], k23 = K[2 * gq + 1];
                const cplx<float> kk[4] = {mk<float>(k01.x, k01.y), mk<float>(k01.z, k01.w), mk<float>(k23.x, k23.y),
                                           mk<float>(k23.z, k23.w)};
                cplx<float> a[4], c[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) { a[s] = X0[task][s] * kk[s]; c[s] = X1[task][s] * kk[s]; }
                dit4<float, FFT_INV>(a); dit4<float, FFT_INV>(c);
                const int p = task ? pD1 : pD0;
#pragma unroll
                for (int i = 0; i < 4; ++i) tile[p + i] = make_float4(a[i].re, a[i].im, c[i].re, c[i].im);
            }
            __syncthreads();
            {   // stage B=32
                cplx<float> a[8], c[8];
                float4 w[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) w[m] = tw[m * L2K_TWJ + L2K_TW2 + (t & 3)];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float4 v = tile[pC + 4 * s + (s >> 1)];
                    cplx<float> va = mk<float>(v.x, v.y), vb = mk<float>(v.z, v.w);
                    if (s) {
                        const float4 ww = w[s >> 1];
                        const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                        va = mul_conj(va, tt); vb = mul_conj(vb, tt);
                    }
                    a[s] = va; c[s] = vb;
                }
                dit8<float, FFT_INV>(a); dit8<float, FFT_INV>(c);
#pragma unroll
                for (int i = 0; i < 8; ++i) tile[pC + 4 * i + (i >> 1)] = make_float4(a[i].re, a[i].im, c[i].re, c[i].im);
            }
            __syncthreads();
            l2k_stage8<FFT_INV, 36>(tile, pB, tw, L2K_TW1 + (t & 31));
            __syncthreads();
            // ---- last stage B=2048: outputs n = t + 256 i leave from registers
            cplx<float> a[8], c[8];
            {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float4 v = tile[pA + s * 288];
                    cplx<float> va = mk<float>(v.x, v.y), vb = mk<float>(v.z, v.w);
                    if (s) {
                        const float4 ww = wA[s >> 1];
                        const cplx<float> tt = (s & 1) ? mk<float>(ww.z, ww.w) : mk<float>(ww.x, ww.y);
                        va = mul_conj(va, tt); vb = mul_conj(vb, tt);
                    }
                    a[s] = va; c[s] = vb;
                }
                dit8<float, FFT_INV>(a); dit8<float, FFT_INV>(c);
            }
            float acc = 0.0f;
#pragma unroll
            for (int col = 0; col < 2; ++col) {
                const i64 blk = blk0 + col;
                if (blk >= g.n_blocks) break;
                const i64 o0 = blk * V;                  // output ordinal of this block's first valid sample
                const cplx<float>* y = col ? c : a;
                if (g.level == 0) {
                    const i64 cell0 = (chan * g.n_bands + b) * g.n_points + o0 - half;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int pv = t + 256 * i - half;
                        if (pv >= 0 && pv < V && o0 + pv < g.n_out) {
                            const float pw = norm2(y[i]);
                            if (out_power) out_power[cell0 + t + 256 * i] = pw;
                            if (out_complex) out_complex[cell0 + t + 256 * i] = y[i];
                            acc += pw;
                        }
                    }
                } else {
                    cplx<float>* wdst = wbuf + band.w_off + chan * band.w_stride + o0 - half;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int pv = t + 256 * i - half;
                        if (pv >= 0 && pv < V && o0 + pv < g.n_out) {
                            wdst[t + 256 * i] = y[i];
                            // sum over the record proper (q in [0, N >> level), not the halo) for the power estimate
                            const i64 q = g.q_first + o0 + pv;
                            if (q >= 0 && q < g.n_level) acc += norm2(y[i]);
                        }
                    }
                }
            }
            if (band_sum) {
                acc = warp_sum(acc);
                if (lane == 0) wsum[warp * L2K_MAXB + bi] += acc;
            }
        }
    }
    if (band_sum) {
        __syncthreads();
        if (t < g.band_count) {
            double s = 0.0;
            for (int w = 0; w < L2K_THREADS / 32; ++w) s += (double)wsum[w * L2K_MAXB + t];
            atomicAdd(&band_sum[chan * g.n_bands + g.band_first + t], s);
        }
    }
}

template <bool ENV>
__global__ void __launch_bounds__(L2K_THREADS, 2)
mr_level2k_kernel(const float* __restrict__ x, MrLevelGeom g, const MrDevBand* __restrict__ bands,
                  const cplx<float>* __restrict__ tables, const float4* __restrict__ tw_g,
                  cplx<float>* __restrict__ wbuf, float* __restrict__ out_power, cplx<float>* __restrict__ out_complex,
                  double* __restrict__ band_sum, int pairs_per_cta) {
    l2k_body<ENV>(x, g, bands, tables, tw_g, wbuf, out_power, out_complex, band_sum, pairs_per_cta, (int)blockIdx.x);
}

// The deep levels are a few CTAs each and independent of one another: one launch runs up to L2K_MULTI of them side by
// side instead of a chain of latency-bound launches.
constexpr int L2K_MULTI = 8;
struct MrMultiLevel {
    int n;
    int cta_end[L2K_MULTI];          // exclusive prefix sums of CTAs per level (grid.x = cta_end[n-1])
    int ppc[L2K_MULTI];
    MrLevelGeom g[L2K_MULTI];
    const float* x[L2K_MULTI];
    double* sum[L2K_MULTI];
};

__global__ void __launch_bounds__(L2K_THREADS, 2)
mr_level2k_multi_kernel(MrMultiLevel m, const MrDevBand* __restrict__ bands, const cplx<float>* __restrict__ tables,
                        const float4* __restrict__ tw_g, cplx<float>* __restrict__ wbuf) {
    int li = 0;
    while (li + 1 < m.n && (int)blockIdx.x >= m.cta_end[li]) ++li;
    const int bx = (int)blockIdx.x - (li ? m.cta_end[li - 1] : 0);
    const MrLevelGeom g = m.g[li];
    l2k_body<true>(m.x[li], g, bands, tables, tw_g, wbuf, nullptr, nullptr, m.sum[li], m.ppc[li], bx);
}

}  // namespace qi
