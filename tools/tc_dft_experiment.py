"""
Decision experiment for the north star's "tensor cores only if a DFT-as-GEMM radix stage beats the SIMT FFT"
(SURVEY 7.1 step 9): the 2048-point block transform of the level-0 convolution (mr_level2k_kernel) written as two
dense DFT stages, 2048 = 64 x 32,

    stage 1   [Re; Im] (128 x cols)  <-  [[C, S], [-S, C]] (128 x 128) @ [Re; Im]      radix 64, 512 flop / point
    twiddle   elementwise complex multiply
    stage 2   radix 32 the same way (64 x 64 real matrix), 256 flop / point

run on the tensor cores by the vendor's GEMM (cuBLAS through torch.matmul: the best case a hand-written tcgen05 kernel
could hope for on these shapes) in the three arithmetic modes that come into question:
    bf16        one product, 8-bit mantissa           -> fails the 1e-4 power tolerance, listed for the rate only
    bf16 x3     hi*hi + hi*lo + lo*hi split, ~16 bits -> meets the tolerance
    tf32        10-bit mantissa                       -> fails the tolerance
and compared with what the SIMT kernels of this repository sustain on the same blocks.  Prints one JSON object.
"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from quantum_inferno_b200 import _runtime  # noqa: E402

dev = torch.device("cuda", 0)
rt = _runtime.get_runtime()
N_BLOCKS = int(os.environ.get("QI_TC_BLOCKS", "71112"))       # 8 channels x 8889 blocks of the headline step
F = 2048
R1, R2 = 64, 32


def dft_real_matrix(r, dtype):
    k = torch.arange(r, dtype=torch.float64)
    ang = -2 * math.pi * torch.outer(k, k) / r
    c, s = torch.cos(ang), torch.sin(ang)
    return torch.cat([torch.cat([c, -s], 1), torch.cat([s, c], 1)], 0).to(device=dev, dtype=dtype)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    points = N_BLOCKS * F
    out = {"blocks": N_BLOCKS, "points": points}
    x = torch.randn(N_BLOCKS, F, dtype=torch.complex64, device=dev)
    ref = torch.fft.fft(x.to(torch.complex128), dim=1)
    # layout of stage 1: rows = (re, im) x the radix-64 index, columns = everything else
    xs = x.view(N_BLOCKS, R1, R2).permute(1, 0, 2).reshape(R1, -1)             # [64, blocks * 32]
    b1 = torch.cat([xs.real, xs.imag], 0).contiguous()                          # [128, cols]
    cols = b1.shape[1]
    tw = torch.exp(-2j * math.pi * torch.outer(torch.arange(R1, dtype=torch.float64), torch.arange(R2, dtype=torch.float64)) / F)
    tw = tw.to(device=dev, dtype=torch.complex64)

    def split(t):
        hi = t.to(torch.bfloat16)
        lo = (t - hi.float()).to(torch.bfloat16)
        return hi, lo

    def run(mode):
        def gemm(a_f32, b_f32):
            if mode == "fp32":
                torch.backends.cuda.matmul.allow_tf32 = False
                return a_f32 @ b_f32
            if mode == "tf32":
                torch.backends.cuda.matmul.allow_tf32 = True
                return a_f32 @ b_f32
            ah, al = split(a_f32)
            bh, bl = split(b_f32)
            mm = lambda p, q: torch.mm(p, q, out_dtype=torch.float32)          # bf16 operands, fp32 accumulator and result
            if mode == "bf16":
                return mm(ah, bh)
            return mm(ah, bh) + mm(ah, bl) + mm(al, bh)
        m1, m2 = dft_real_matrix(R1, torch.float32), dft_real_matrix(R2, torch.float32)
        y1 = gemm(m1, b1)                                                      # [128, cols]
        z = torch.complex(y1[:R1], y1[R1:]).view(R1, N_BLOCKS, R2) * tw[:, None, :]          # twiddle
        zs = z.permute(2, 1, 0).reshape(R2, -1)                               # radix-32 index to the rows
        b2 = torch.cat([zs.real, zs.imag], 0).contiguous()
        y2 = gemm(m2, b2)
        res = torch.complex(y2[:R2], y2[R2:]).view(R2, N_BLOCKS, R1)          # [k2, block, k1] -> X[k1 + 64 k2]
        return res.permute(1, 0, 2).reshape(N_BLOCKS, F)

    def gemms_only(mode):
        """Only the matrix products (the part a fused kernel cannot remove), operands prepared beforehand."""
        m1, m2 = dft_real_matrix(R1, torch.float32), dft_real_matrix(R2, torch.float32)
        b2 = torch.randn(2 * R2, N_BLOCKS * R1, device=dev)
        if mode in ("fp32", "tf32"):
            torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
            return lambda: (m1 @ b1, m2 @ b2)
        (m1h, m1l), (m2h, m2l), (b1h, b1l), (b2h, b2l) = split(m1), split(m2), split(b1), split(b2)
        mm = lambda p, q: torch.mm(p, q, out_dtype=torch.float32)
        if mode == "bf16":
            return lambda: (mm(m1h, b1h), mm(m2h, b2h))
        return lambda: (mm(m1h, b1h), mm(m1h, b1l), mm(m1l, b1h), mm(m2h, b2h), mm(m2h, b2l), mm(m2l, b2h))

    flop_single = 2 * (2 * R1) ** 2 * cols + 2 * (2 * R2) ** 2 * (N_BLOCKS * R1)
    for mode in ("bf16", "bf16x3", "tf32", "fp32"):
        got = run(mode)
        err = float((got.to(torch.complex128) - ref).norm() / ref.norm())
        ms_all = timed(lambda: run(mode), 3)
        ms_mm = timed(gemms_only(mode), 5)
        nprod = 3 if mode == "bf16x3" else 1
        out[mode] = {"rel_l2_error": err, "ms_whole_unfused": ms_all, "ms_matrix_products_only": ms_mm,
                     "points_per_s_products_only": points / ms_mm * 1e3,
                     "tflops_products_only": nprod * flop_single / ms_mm / 1e9}
        del got
    # SIMT: the generic multi-pass FFT kernel of this library on the same blocks (one 2048-point pass)
    y = torch.empty_like(x)
    code = _runtime.DTYPE_CODE["float32"]
    half = N_BLOCKS // 2                                                   # the launch grid holds 65535 batches

    def simt():
        assert rt.lib.qi_fft_c2c(rt.ptr(x), rt.ptr(y), half, 11, 0, code, rt.stream()) == 0
        assert rt.lib.qi_fft_c2c(rt.ptr(x[half:]), rt.ptr(y[half:]), N_BLOCKS - half, 11, 0, code, rt.stream()) == 0
    ms = timed(simt)
    out["simt_generic_pass"] = {"ms": ms, "points_per_s": points / ms * 1e3}
    ms = timed(lambda: torch.fft.fft(x, dim=1))
    out["cufft_c2c"] = {"ms": ms, "points_per_s": points / ms * 1e3}
    out["note"] = ("mr_level2k_kernel<level 0> (the kernel the question is about) sustains 9 transforms x 8 ch x 8889 blocks "
                   "x 2048 points in the time bench.py reports as category inv_mid, spectrum product, |.|^2 and plane "
                   "stores included; see profiles/r02_tc_dft_decision.md")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
