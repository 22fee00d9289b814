// qi_mr_level2kp.cuh -- the 2048-point level convolution of qi_mr_level2k.cuh (bands that keep their carrier: level 0
// of the headline path, every level of the complex-TFR mode) with PACKED arithmetic.
//
// Same algorithm, tile geometry, padding and barriers as l2k_body (read that header first).  What changes is the
// arithmetic type: the two blocks a CTA convolves side by side are no longer two complex numbers (a.re, a.im, b.re, b.im)
// of a tile row but the two halves of packed pairs,
//     row = ( {a.re, b.re}, {a.im, b.im} )          (one 128-bit shared-memory access, two 64-bit register pairs)
// so that every complex addition of the butterflies is two add/sub.f32x2 for BOTH blocks, every twiddle product five
// packed instructions (mul, mul, sub, mul, fma) instead of eight scalar ones, and the quarter turns stay free (they only
// choose between add and sub).  The sqrt(1/2) rotations of the radix-8 butterfly are folded into the additions that
// consume them (fma with {h, h} / {-h, -h}).  A radix-8 butterfly on two blocks is 52 packed + 35 twiddle instructions
// where the scalar version needs about 170.
//
// The packed operands need the twiddles and the band tables as {t.re, t.re}, {t.im, t.im}: the stage twiddles get their
// own shared-memory table (7 slots x 292 rows x 16 B = 32 KB; mr_twiddle2kp_kernel), the band tables a second,
// duplicated copy in the workspace (mr_table_kernel writes both).
#pragma once
#include "qi_mr_level2k.cuh"

namespace qi {

constexpr int L2KP_ROWS = 292;                        // twiddle rows: stage B=2048 (j<256), B=256 (j<32), B=32 (j<4)
constexpr int L2KP_TW_TOTAL = 7 * L2KP_ROWS;          // float4 entries, [slot s-1][row]
constexpr size_t L2KP_SMEM = (size_t)(2 * L2K_TILE + L2KP_TW_TOTAL) * 16 + (size_t)(L2K_THREADS / 32) * L2K_MAXB * 4;

struct __align__(16) c2 { f32x2 re, im; };            // two complex numbers: ({a.re, b.re}, {a.im, b.im})

// twp[(s - 1) * L2KP_ROWS + row] = ({t.re, t.re}, {t.im, t.im}),  t = w_B^(j * brev3(s)),  w_B = exp(-2 pi i / B)
__global__ void mr_twiddle2kp_kernel(float4* __restrict__ twp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L2KP_TW_TOTAL) return;
    const int s = idx / L2KP_ROWS + 1, row = idx % L2KP_ROWS;
    int logB, j;
    if (row < L2K_TW1) { logB = 11; j = row; }
    else if (row < L2K_TW2) { logB = 8; j = row - L2K_TW1; }
    else { logB = 5; j = row - L2K_TW2; }
    const cplx<float> w = conj(unit_root<float>((unsigned long long)(j * brev3(s)), logB));
    twp[idx] = make_float4(w.re, w.re, w.im, w.im);
}

QI_DEV c2 c2_add(c2 a, c2 b) { c2 r; r.re = f2_add(a.re, b.re); r.im = f2_add(a.im, b.im); return r; }
QI_DEV c2 c2_sub(c2 a, c2 b) { c2 r; r.re = f2_sub(a.re, b.re); r.im = f2_sub(a.im, b.im); return r; }
// u + q(d), u - q(d) with q the quarter turn of the transform direction: * (-i) forward, * (+i) inverse
template <int DIR> QI_DEV c2 c2_addq(c2 u, c2 d) {
    c2 r;
    if (DIR == FFT_FWD) { r.re = f2_add(u.re, d.im); r.im = f2_sub(u.im, d.re); }
    else { r.re = f2_sub(u.re, d.im); r.im = f2_add(u.im, d.re); }
    return r;
}
template <int DIR> QI_DEV c2 c2_subq(c2 u, c2 d) {
    c2 r;
    if (DIR == FFT_FWD) { r.re = f2_sub(u.re, d.im); r.im = f2_add(u.im, d.re); }
    else { r.re = f2_add(u.re, d.im); r.im = f2_sub(u.im, d.re); }
    return r;
}
// v * t (forward) or v * conj(t) (inverse); t holds duplicated halves
template <int DIR> QI_DEV c2 c2_tw(c2 v, c2 t) {
    c2 r;
    if (DIR == FFT_FWD) {
        r.re = f2_sub(f2_mul(v.re, t.re), f2_mul(v.im, t.im));
        r.im = f2_fma(v.re, t.im, f2_mul(v.im, t.re));
    } else {
        r.re = f2_fma(v.im, t.im, f2_mul(v.re, t.re));
        r.im = f2_sub(f2_mul(v.im, t.re), f2_mul(v.re, t.im));
    }
    return r;
}

// forward DIF butterflies (natural slots in, bit-reversed slots out) -- the packed mirror of dif4 / dif8<., FFT_FWD>
QI_DEV void p_dif4(c2* a) {
    const c2 b0 = c2_add(a[0], a[2]), b2 = c2_sub(a[0], a[2]);
    const c2 b1 = c2_add(a[1], a[3]), d13 = c2_sub(a[1], a[3]);
    a[0] = c2_add(b0, b1); a[1] = c2_sub(b0, b1);
    a[2] = c2_addq<FFT_FWD>(b2, d13); a[3] = c2_subq<FFT_FWD>(b2, d13);
}
QI_DEV void p_dif8(c2* a, f32x2 H, f32x2 NH) {
    const c2 b0 = c2_add(a[0], a[4]), b4 = c2_sub(a[0], a[4]);
    const c2 b1 = c2_add(a[1], a[5]), d15 = c2_sub(a[1], a[5]);
    const c2 b2 = c2_add(a[2], a[6]), d26 = c2_sub(a[2], a[6]);
    const c2 b3 = c2_add(a[3], a[7]), d37 = c2_sub(a[3], a[7]);
    const c2 c0 = c2_add(b0, b2), cc2 = c2_sub(b0, b2), c1 = c2_add(b1, b3), d13 = c2_sub(b1, b3);
    a[0] = c2_add(c0, c1); a[1] = c2_sub(c0, c1);
    a[2] = c2_addq<FFT_FWD>(cc2, d13); a[3] = c2_subq<FFT_FWD>(cc2, d13);
    // b5 = w8 (a1 - a5) = h (p1, q1),  b7 = w8^3 (a3 - a7) = h (p3, -q3)
    const f32x2 p1 = f2_add(d15.re, d15.im), q1 = f2_sub(d15.im, d15.re);
    const f32x2 p3 = f2_sub(d37.im, d37.re), q3 = f2_add(d37.re, d37.im);
    c2 e, f;                                   // c5 = b5 + b7 = h e,  b5 - b7 = h f
    e.re = f2_add(p1, p3); e.im = f2_sub(q1, q3);
    f.re = f2_sub(p1, p3); f.im = f2_add(q1, q3);
    const c2 c4 = c2_addq<FFT_FWD>(b4, d26), c6 = c2_subq<FFT_FWD>(b4, d26);
    a[4].re = f2_fma(e.re, H, c4.re);  a[4].im = f2_fma(e.im, H, c4.im);
    a[5].re = f2_fma(e.re, NH, c4.re); a[5].im = f2_fma(e.im, NH, c4.im);
    a[6].re = f2_fma(f.im, H, c6.re);  a[6].im = f2_fma(f.re, NH, c6.im);      // c6 + (-i) h f
    a[7].re = f2_fma(f.im, NH, c6.re); a[7].im = f2_fma(f.re, H, c6.im);
}
// inverse DIT butterflies (bit-reversed slots in, natural slots out) -- the packed mirror of dit4 / dit8<., FFT_INV>
QI_DEV void p_dit4(c2* a) {
    const c2 b0 = c2_add(a[0], a[1]), b1 = c2_sub(a[0], a[1]);
    const c2 b2 = c2_add(a[2], a[3]), d23 = c2_sub(a[2], a[3]);
    a[0] = c2_add(b0, b2); a[2] = c2_sub(b0, b2);
    a[1] = c2_addq<FFT_INV>(b1, d23); a[3] = c2_subq<FFT_INV>(b1, d23);
}
QI_DEV void p_dit8(c2* a, f32x2 H, f32x2 NH) {
    const c2 c0 = c2_add(a[0], a[1]), c1 = c2_sub(a[0], a[1]);
    const c2 cc2 = c2_add(a[2], a[3]), d23 = c2_sub(a[2], a[3]);
    const c2 c4 = c2_add(a[4], a[5]), c5 = c2_sub(a[4], a[5]);
    const c2 c6 = c2_add(a[6], a[7]), d67 = c2_sub(a[6], a[7]);
    const c2 b0 = c2_add(c0, cc2), b2 = c2_sub(c0, cc2);
    const c2 b1 = c2_addq<FFT_INV>(c1, d23), b3 = c2_subq<FFT_INV>(c1, d23);
    const c2 b4 = c2_add(c4, c6), d46 = c2_sub(c4, c6);
    const c2 s = c2_addq<FFT_INV>(c5, d67), d = c2_subq<FFT_INV>(c5, d67);
    a[0] = c2_add(b0, b4); a[4] = c2_sub(b0, b4);
    a[2] = c2_addq<FFT_INV>(b2, d46); a[6] = c2_subq<FFT_INV>(b2, d46);
    // b5 = conj(w8) s = h (s.re - s.im, s.re + s.im),  b7 = conj(w8^3) d = h (-(d.re + d.im), d.re - d.im)
    const f32x2 p = f2_sub(s.re, s.im), q = f2_add(s.re, s.im);
    const f32x2 pp = f2_add(d.re, d.im), qq = f2_sub(d.re, d.im);
    a[1].re = f2_fma(p, H, b1.re);   a[1].im = f2_fma(q, H, b1.im);
    a[5].re = f2_fma(p, NH, b1.re);  a[5].im = f2_fma(q, NH, b1.im);
    a[3].re = f2_fma(pp, NH, b3.re); a[3].im = f2_fma(qq, H, b3.im);
    a[7].re = f2_fma(pp, H, b3.re);  a[7].im = f2_fma(qq, NH, b3.im);
}

// radix-8 stage; rows base + i*H live at tile[p0 + row_off(i)], twiddles of the stage at twp[. * L2KP_ROWS + twrow]
template <int DIR, typename RowOff>
QI_DEV void p_stage8(c2* __restrict__ tile, int p0, RowOff row_off, const c2* __restrict__ twp, int twrow, f32x2 H, f32x2 NH) {
    c2 a[8];
    if (DIR == FFT_FWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = tile[p0 + row_off(i)];
        p_dif8(a, H, NH);
#pragma unroll
        for (int s = 0; s < 8; ++s)
            tile[p0 + row_off(s)] = s ? c2_tw<FFT_FWD>(a[s], twp[(s - 1) * L2KP_ROWS + twrow]) : a[s];
    } else {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const c2 v = tile[p0 + row_off(s)];
            a[s] = s ? c2_tw<FFT_INV>(v, twp[(s - 1) * L2KP_ROWS + twrow]) : v;
        }
        p_dit8(a, H, NH);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[p0 + row_off(i)] = a[i];
    }
}

// tables_dup: the band tables as ({k.re, k.re}, {k.im, k.im}) per row, same offsets as `tables`
QI_DEV void l2kp_body(const float* __restrict__ x, const MrLevelGeom& g, const MrDevBand* __restrict__ bands,
                      const float4* __restrict__ tables_dup, const float4* __restrict__ twp_g,
                      cplx<float>* __restrict__ wbuf, float* __restrict__ out_power, cplx<float>* __restrict__ out_complex,
                      double* __restrict__ band_sum, int pairs_per_cta, int bx) {
    QI_DYN_SMEM(smem_raw);
    c2* tile0 = reinterpret_cast<c2*>(smem_raw);
    c2* tile1 = tile0 + L2K_TILE;
    c2* twp = tile1 + L2K_TILE;
    float* wsum = reinterpret_cast<float*>(twp + L2KP_TW_TOTAL);          // [warps][L2K_MAXB]
    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const i64 chan = blockIdx.y;
    const int V = L2K_F - g.wk;
    const int half = g.wk / 2;
    const float* xs = x + chan * g.x_stride;
    const f32x2 H = f2_make(0.70710678118654752440f, 0.70710678118654752440f);
    const f32x2 NH = f2_make(-0.70710678118654752440f, -0.70710678118654752440f);

    {
        float4* dst = reinterpret_cast<float4*>(twp);
        for (int i = t; i < L2KP_TW_TOTAL; i += L2K_THREADS) dst[i] = twp_g[i];
    }
    for (int i = t; i < (L2K_THREADS / 32) * L2K_MAXB; i += L2K_THREADS) wsum[i] = 0.0f;

    // per-thread tile addresses (padded rows), as in l2k_body
    const int pA = t + (t >> 3);
    const int rB = ((t >> 5) << 8) + (t & 31);
    const int pB = rB + (rB >> 3);
    const int pC = (t >> 2) * 36 + (t & 3);
    const int pD0 = 4 * t + (t >> 1);
    const int pD1 = 4 * (t + 256) + ((t + 256) >> 1);
    auto offB = [](int i) { return i * 36; };
    auto offC = [](int i) { return 4 * i + (i >> 1); };

    for (int pp = 0; pp < pairs_per_cta; ++pp) {
        const i64 blk0 = 2 * ((i64)bx * pairs_per_cta + pp);
        if (blk0 >= g.n_blocks) break;                                     // uniform over the CTA
        __syncthreads();
        // ---- forward, stage B=2048 straight from HBM (real input, two blocks)
        {
            c2 a[8];
            const i64 k0 = g.q_first + blk0 * V - half + g.x_halo + t;
            const bool has_b = blk0 + 1 < g.n_blocks;
            const f32x2 zero = f2_make(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const i64 ka = k0 + 256 * i, kb = ka + V;
                const float va = (ka >= 0 && ka < g.x_len) ? xs[ka] : 0.0f;
                const float vb = (has_b && kb >= 0 && kb < g.x_len) ? xs[kb] : 0.0f;
                a[i].re = f2_make(va, vb); a[i].im = zero;
            }
            p_dif8(a, H, NH);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tile0[pA + s * 288] = s ? c2_tw<FFT_FWD>(a[s], twp[(s - 1) * L2KP_ROWS + L2K_TW0 + t]) : a[s];
        }
        __syncthreads();
        p_stage8<FFT_FWD>(tile0, pB, offB, twp, L2K_TW1 + (t & 31), H, NH);
        __syncthreads();
        p_stage8<FFT_FWD>(tile0, pC, offC, twp, L2K_TW2 + (t & 3), H, NH);
        __syncthreads();
        // ---- forward radix-4 stage: spectra stay in registers (rows 4g .. 4g+3, g = t and t + 256)
        c2 X[2][4];
#pragma unroll
        for (int task = 0; task < 2; ++task) {
            const int p = task ? pD1 : pD0;
#pragma unroll
            for (int i = 0; i < 4; ++i) X[task][i] = tile0[p + i];
            p_dif4(X[task]);
        }

        // valid output range of the two blocks (pv = position among the V valid samples of a block)
        int lim[2], s0[2], s1[2];
        i64 o0[2];
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            const i64 blk = blk0 + col;
            o0[col] = blk * V;
            i64 rem = blk < g.n_blocks ? g.n_out - o0[col] : 0;
            lim[col] = rem < V ? (rem < 0 ? 0 : (int)rem) : V;
            // samples that count for the raw power sum: q = q_first + o0 + pv in [0, n_level)
            const i64 qa = -(g.q_first + o0[col]), qb = g.n_level - g.q_first - o0[col];
            s0[col] = qa < 0 ? 0 : (qa > V ? V : (int)qa);
            s1[col] = qb < 0 ? 0 : (qb > lim[col] ? lim[col] : (int)qb);
        }

        for (int bi = 0; bi < g.band_count; ++bi) {
            const int b = g.band_first + bi;
            const MrDevBand band = bands[b];
            const c2* K = reinterpret_cast<const c2*>(tables_dup + band.table_off);
            c2* tile = (bi & 1) ? tile0 : tile1;
            // inverse radix-4 stage on X * K, from registers
#pragma unroll
            for (int task = 0; task < 2; ++task) {
                const int gq = t + 256 * task;
                c2 a[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) a[s] = c2_tw<FFT_FWD>(X[task][s], K[4 * gq + s]);
                p_dit4(a);
                const int p = task ? pD1 : pD0;
#pragma unroll
                for (int i = 0; i < 4; ++i) tile[p + i] = a[i];
            }
            __syncthreads();
            p_stage8<FFT_INV>(tile, pC, offC, twp, L2K_TW2 + (t & 3), H, NH);
            __syncthreads();
            p_stage8<FFT_INV>(tile, pB, offB, twp, L2K_TW1 + (t & 31), H, NH);
            __syncthreads();
            // ---- last stage B=2048: outputs n = t + 256 i leave from registers
            c2 y[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const c2 v = tile[pA + s * 288];
                y[s] = s ? c2_tw<FFT_INV>(v, twp[(s - 1) * L2KP_ROWS + L2K_TW0 + t]) : v;
            }
            p_dit8(y, H, NH);
            float acc = 0.0f;
            if (g.level == 0) {
                float* pa = out_power ? out_power + (chan * g.n_bands + b) * g.n_points - half + t : nullptr;
                cplx<float>* ca = out_complex ? out_complex + (chan * g.n_bands + b) * g.n_points - half + t : nullptr;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pv = t + 256 * i - half;
                    const f32x2 pw2 = f2_fma(y[i].re, y[i].re, f2_mul(y[i].im, y[i].im));
                    const float pw[2] = {f2_lo(pw2), f2_hi(pw2)};
#pragma unroll
                    for (int col = 0; col < 2; ++col) {
                        if ((unsigned)pv < (unsigned)lim[col]) {
                            if (pa) pa[o0[col] + 256 * i] = pw[col];
                            if (ca) ca[o0[col] + 256 * i] = col ? mk<float>(f2_hi(y[i].re), f2_hi(y[i].im))
                                                                : mk<float>(f2_lo(y[i].re), f2_lo(y[i].im));
                            acc += pw[col];
                        }
                    }
                }
            } else {
                cplx<float>* wd = wbuf + band.w_off + chan * band.w_stride - half + t;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pv = t + 256 * i - half;
                    const f32x2 pw2 = f2_fma(y[i].re, y[i].re, f2_mul(y[i].im, y[i].im));
                    const float pw[2] = {f2_lo(pw2), f2_hi(pw2)};
#pragma unroll
                    for (int col = 0; col < 2; ++col) {
                        if ((unsigned)pv < (unsigned)lim[col]) {
                            wd[o0[col] + 256 * i] = col ? mk<float>(f2_hi(y[i].re), f2_hi(y[i].im))
                                                        : mk<float>(f2_lo(y[i].re), f2_lo(y[i].im));
                            if (pv >= s0[col] && pv < s1[col]) acc += pw[col];
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

__global__ void __launch_bounds__(L2K_THREADS, 2)
mr_level2kp_kernel(const float* __restrict__ x, MrLevelGeom g, const MrDevBand* __restrict__ bands,
                   const float4* __restrict__ tables_dup, const float4* __restrict__ twp_g,
                   cplx<float>* __restrict__ wbuf, float* __restrict__ out_power, cplx<float>* __restrict__ out_complex,
                   double* __restrict__ band_sum, int pairs_per_cta) {
    l2kp_body(x, g, bands, tables_dup, twp_g, wbuf, out_power, out_complex, band_sum, pairs_per_cta, (int)blockIdx.x);
}

}  // namespace qi
