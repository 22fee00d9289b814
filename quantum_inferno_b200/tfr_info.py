"""
Power, information and entropy of a time-frequency representation on the B200 -- drop-in for
``quantum_inferno.tfr_info`` (same names, attributes and return order; reference tfr_info.py:13-260).

Every array-sized reduction / elementwise plane runs in csrc/qi_info.cu with fp64 accumulation; numpy in ->
numpy out, CUDA tensor in -> tensors out.  ``dtype=`` (keyword-only, default float64) selects the plane
precision.  A leading batch axis ([channels, bands, time]) is accepted as an extension: every matrix is
normalised independently, exactly as if the reference were called once per channel.
"""
from typing import Tuple

import numpy as np

from . import _driver
from . import scales_dyadic as scales
from ._runtime import dtype_name, finish, get_runtime


# ----------------------------------------------------------------------------- scalar helpers (host)
def log2_ceil(x: float, epsilon: float = scales.EPSILON64) -> float:
    """ceil(log2(|x| + eps)) (reference tfr_info.py:13-22)."""
    return np.ceil(np.log2(np.abs(x) + epsilon))


def log2_round(x: float, epsilon: float = scales.EPSILON64) -> float:
    """round(log2(|x| + eps)) (reference tfr_info.py:25-34)."""
    return float(np.round(np.log2(np.abs(x) + epsilon)))


def log2_floor(x: float, epsilon: float = scales.EPSILON64) -> float:
    """floor(log2(|x| + eps)) (reference tfr_info.py:37-46)."""
    return np.floor(np.log2(np.abs(x) + epsilon))


def _host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def mat_max_idx(a) -> Tuple[np.ndarray]:
    """Index tuple of the maximum (reference tfr_info.py:49-54)."""
    a = _host(a)
    return np.unravel_index(a.argmax(), a.shape)


def mat_min_idx(a) -> Tuple[np.ndarray]:
    """Index tuple of the minimum (reference tfr_info.py:57-62)."""
    a = _host(a)
    return np.unravel_index(a.argmin(), a.shape)


# ----------------------------------------------------------------------------- device helpers
class _Ctx:
    """Per-call context: runtime, dtype, and whether results go back to numpy."""

    def __init__(self, array_in, dtype):
        self.rt = get_runtime(array_in)
        self.dt = dtype_name(dtype)
        self.want_numpy = not self.rt.is_device_array(array_in)

    def matrix(self, power):
        """[F,T] or [M,F,T] -> device [M,F,T]; returns (buffer, had_batch)."""
        p = self.rt.asarray(power, self.dt)
        if p.ndim == 2:
            return self.rt.reshape(p, (1,) + tuple(p.shape)), False
        if p.ndim != 3:
            raise TypeError(f"Cannot handle an array of shape {tuple(p.shape)}.")
        return p, True

    def out(self, buf, shape=None):
        if buf is None:
            return None
        if shape is not None:
            buf = self.rt.reshape(buf, shape)
        return finish(self.rt, buf, self.want_numpy)


def scale_log2_64(in_array, *, dtype=None):
    """log2(x + EPS64) (reference tfr_info.py:65-70)."""
    c = _Ctx(in_array, dtype)
    x = c.rt.asarray(in_array, c.dt)
    return c.out(_driver.abs_log2(x, c.dt, False, rt=c.rt, signed=True))


def scale_power_bits(power, *, dtype=None):
    """log2(P + EPS64) minus its maximum (reference tfr_info.py:73-79)."""
    c = _Ctx(power, dtype)
    p = c.rt.asarray(power, c.dt)
    shape = tuple(int(s) for s in p.shape)
    flat = c.rt.reshape(p, (1, 1, int(np.prod(shape))))
    mx = _driver.power_reduce(flat, c.dt, maximum=True, rt=c.rt)["max"]
    return c.out(_driver.power_bits(flat, c.dt, mx, rt=c.rt), shape)


def power_dynamics_scaled_bits(tfr_power, *, dtype=None):
    """(bits re max, per-time dynamic range, per-frequency dynamic range) (reference tfr_info.py:82-94)."""
    c = _Ctx(tfr_power, dtype)
    rt, dt = c.rt, c.dt
    p, batched = c.matrix(tfr_power)
    M, F, T = (int(s) for s in p.shape)
    red = _driver.power_reduce(p, dt, rows=True, cols=True, maximum=True, rt=rt)
    outs = [(_driver.power_bits(p, dt, red["max"], rt=rt), (M, F, T) if batched else (F, T))]
    for sums, n in ((red["col_sum"], T), (red["row_sum"], F)):
        v3 = rt.reshape(rt.asarray(sums, dt), (M, 1, n))                 # fp64 sums -> plane dtype
        mx = _driver.power_reduce(v3, dt, maximum=True, rt=rt)["max"]
        outs.append((_driver.power_bits(v3, dt, mx, rt=rt), (M, n) if batched else (n,)))
    return tuple(c.out(b, s) for b, s in outs)


def get_info_and_entropy_32(marginal, *, dtype=None):
    """(info, entropy, ref_entropy) of a 1-D marginal with EPS32 (reference tfr_info.py:97-103)."""
    s = Shannon(marginal, dtype=dtype)
    return s.info, s.entropy, s.ref_entropy


class Shannon:
    """Shannon information of a 1-D marginal (reference tfr_info.py:106-135).
    Attributes: marginal, info, entropy, ref_entropy, isnr, esnr."""

    def __init__(self, marginal, *, dtype=None):
        c = _Ctx(marginal, dtype)
        m = c.rt.asarray(marginal, c.dt)
        n = int(m.shape[-1])
        self.marginal = marginal
        self._fill(c, c.rt.reshape(m, (1, 1, n)), 3, None, want_pdf=False)

    def _fill(self, c, p3, mode, norm, want_pdf):
        n = int(p3.shape[-1])
        planes = ("info", "bits", "isnr", "esnr") + (("pdf",) if want_pdf else ())
        res = _driver.shannon(p3, c.dt, mode, norm, n, eps=_driver.EPS32, planes=planes, rt=c.rt)
        if want_pdf:
            self.marginal = c.out(res["pdf"], (n,))
        self.info = c.out(res["info"], (n,))
        self.entropy = c.out(res["bits"], (n,))
        self.ref_entropy = np.log2(n) / n
        self.isnr = c.out(res["isnr"], (n,))
        self.esnr = c.out(res["esnr"], (n,))


class ShannonTDR(Shannon):
    """Time-domain Shannon information: marginal = (x/||x||)^2 (reference tfr_info.py:138-160)."""

    def __init__(self, sig_in_real, *, dtype=None):
        c = _Ctx(sig_in_real, dtype)
        x = c.rt.asarray(sig_in_real, c.dt)
        n = int(x.shape[-1])
        sn, mg, _ = _driver.tdr_marginal(c.rt.reshape(x, (1, n)), c.dt, rt=c.rt)
        self.sig = c.out(sn, (n,))
        self.marginal = c.out(mg, (n,))
        self._fill(c, c.rt.reshape(mg, (1, 1, n)), 3, None, want_pdf=False)

    def print_total_ref_entropy(self):
        print("Ref entropy, time:", self.ref_entropy)

    def print_total_entropy(self):
        print("Total Entropy, time:", _host(self.entropy).sum())

    def print_total_marginal(self):
        print("Sum of time marginal:", _host(self.marginal).sum())


class ShannonFFT(Shannon):
    """Spectral Shannon information: marginal = |rfft(x)|^2 / sum (reference tfr_info.py:163-190).  Any record length
    (lengths that are not a power of two go through Bluestein's identity, see _driver._rfft_bluestein)."""

    def __init__(self, sig_in_real, *, dtype=None):
        c = _Ctx(sig_in_real, dtype)
        rt, dt = c.rt, c.dt
        x = rt.asarray(sig_in_real, dt)
        n = int(x.shape[-1])
        k = n // 2 + 1
        z = _driver.rfft(rt.reshape(x, (1, n)), dt, rt=rt)
        p3 = rt.reshape(_driver.abs_log2(z, dt, True, eps=0.0, square=True, rt=rt), (1, 1, k))     # |z|^2
        tot = _driver.power_reduce(p3, dt, total=True, rt=rt)["total"]
        self.sig = c.out(z, (k,))
        self.angle_rads = np.unwrap(np.angle(_host(self.sig)))        # 1-D sequential phase unwrap: host
        self.frequency = np.arange(k) / k / 2.0
        self._fill(c, p3, 0, tot, want_pdf=True)

    def print_total_ref_entropy(self):
        print("Ref entropy, frequency:", self.ref_entropy)

    def print_total_entropy(self):
        print("Total Entropy, frequency:", _host(self.entropy).sum())

    def print_total_marginal(self):
        print("Sum of frequency marginal:", _host(self.marginal).sum())


def shannon_tdr_fft(sig_in_real, *, dtype=None) -> Tuple[ShannonTDR, ShannonFFT]:
    """(ShannonTDR, ShannonFFT) of a record (reference tfr_info.py:193-200)."""
    return ShannonTDR(sig_in_real, dtype=dtype), ShannonFFT(sig_in_real, dtype=dtype)


class ShannonStft:
    """Information planes of a normalised time-frequency pdf (reference tfr_info.py:203-228).
    Attributes: info, shannon_bits, ref_bits, isnr, esnr."""

    def __init__(self, tfr_pow_pdf, deg_free: int, *, dtype=None):
        c = _Ctx(tfr_pow_pdf, dtype)
        p, batched = c.matrix(tfr_pow_pdf)
        self._planes(c, p, batched, 3, None, deg_free)

    def _planes(self, c, p, batched, mode, norm, deg_free):
        res = _driver.shannon(p, c.dt, mode, norm, deg_free, eps=_driver.EPS64, rt=c.rt)
        shape = tuple(int(s) for s in p.shape) if batched else tuple(int(s) for s in p.shape[1:])
        self.info = c.out(res["info"], shape)
        self.shannon_bits = c.out(res["bits"], shape)
        self.ref_bits = np.log2(deg_free) / deg_free
        self.isnr = c.out(res["isnr"], shape)
        self.esnr = c.out(res["esnr"], shape)


def shannon_stft_from_tfr_power(tfr_power, *, dtype=None) -> ShannonStft:
    """Globally normalised pdf = P / sum(P), D = bands*times (reference tfr_info.py:231-236)."""
    c = _Ctx(tfr_power, dtype)
    p, batched = c.matrix(tfr_power)
    tot = _driver.power_reduce(p, c.dt, total=True, rt=c.rt)["total"]
    obj = ShannonStft.__new__(ShannonStft)
    obj._planes(c, p, batched, 0, tot, int(p.shape[1]) * int(p.shape[2]))
    return obj


class ShannonStftPerTime(ShannonStft):
    """pdf normalised per time step, D = bands (reference tfr_info.py:239-248)."""

    def __init__(self, tfr_power, *, dtype=None):
        c = _Ctx(tfr_power, dtype)
        p, batched = c.matrix(tfr_power)
        cs = _driver.power_reduce(p, c.dt, cols=True, rt=c.rt)["col_sum"]
        self._planes(c, p, batched, 1, cs, int(p.shape[1]))


class ShannonStftPerFreq(ShannonStft):
    """pdf normalised per frequency band, D = times (reference tfr_info.py:251-260)."""

    def __init__(self, tfr_power, *, dtype=None):
        c = _Ctx(tfr_power, dtype)
        p, batched = c.matrix(tfr_power)
        rs = _driver.power_reduce(p, c.dt, rows=True, rt=c.rt)["row_sum"]
        self._planes(c, p, batched, 2, rs, int(p.shape[2]))
