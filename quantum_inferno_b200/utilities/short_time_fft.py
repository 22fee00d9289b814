"""
Tukey-window STFT / spectrogram / inverse STFT on the B200 -- drop-in for
``quantum_inferno.utilities.short_time_fft`` (reference utilities/short_time_fft.py:19-175), the STFT wrapper the
reference's examples call.

The reference builds a ``scipy.signal.ShortTimeFFT`` and calls its ``stft_detrend`` / ``spectrogram`` / ``istft``.
``get_stft_object_tukey`` here returns a SUBCLASS of that SciPy class: SciPy keeps doing the host-side bookkeeping
(window scaling, slice ranges ``p_min`` / ``p_max``, canonical dual window, axes), while the three transforms are
overridden to run on the GPU through the C ABI (csrc/qi_stft.cu: frame gather with the virtual zero extension, per
slice mean removal, window, FFT, phase rotation; inverse FFT pairs, dual window, overlap-add).  Nothing in the
overridden methods calls SciPy's transform code: without a CUDA device they raise like the rest of the package.

Supported (what the reference's wrappers use): real input, ``fft_mode='onesided'``, ``mfft = 2^m``, ``detr`` None or
'constant', ``k_offset = 0``, ``y = None``; anything else raises ``NotImplementedError``.  Extra keyword: ``dtype``.
"""
from typing import Tuple, Union

import numpy as np
from scipy import signal

from .. import _driver
from .._runtime import dtype_name, finish, get_runtime
from .calculations import round_value
from .rescaling import is_power_of_two

# Create dictionaries for the types to avoid having to use Literal when running the functions
scaling_type = ["magnitude", "psd", None]
padding_type = ["zeros", "edge", "even", "odd"]


class ShortTimeFFT(signal.ShortTimeFFT):
    """``scipy.signal.ShortTimeFFT`` whose transforms run on the B200 (see the module docstring)."""

    compute_dtype = "float64"

    # ---- helpers
    def _check_supported(self):
        if self.fft_mode != "onesided":
            raise NotImplementedError("the B200 STFT kernels implement fft_mode='onesided' only")
        if not is_power_of_two(int(self.mfft)):
            raise NotImplementedError(f"the B200 STFT kernels need mfft = 2^m, got {self.mfft}")
        if np.iscomplexobj(self.win):
            raise NotImplementedError("complex windows are not supported")

    def _roll(self):
        return 0 if self.phase_shift is None else int((self.phase_shift + self.m_num_mid) % self.m_num)

    def _padded(self, rt, x2, n, p0, p1, padding):
        """The record extended like np.pad does inside scipy's _x_slices (scipy/signal/_short_time_fft.py), for the
        padding modes the kernel's virtual zero extension does not cover; slice f then starts at sample f*hop."""
        k0 = p0 * self.hop - self.m_num_mid
        k1 = k0 + (p1 - p0) * self.hop + self.m_num
        left, right = -min(k0, 0), max(k1 - n, 0)
        if max(left, right) > n - 1:
            raise NotImplementedError("padding longer than the record is not supported")
        on_gpu = getattr(rt, "name", "") == "cuda"
        flip = (lambda a: a.flip(-1)) if on_gpu else (lambda a: a[..., ::-1])
        rep = (lambda a, r: a.repeat(1, r)) if on_gpu else (lambda a, r: np.repeat(a, r, axis=-1))
        cat = (lambda parts: rt.torch.cat(parts, dim=-1)) if on_gpu else (lambda parts: np.concatenate(parts, axis=-1))
        first, last = x2[:, :1], x2[:, n - 1:]
        if padding == "edge":
            lo, hi = rep(first, left), rep(last, right)
        else:
            lo, hi = flip(x2[:, 1:left + 1]), flip(x2[:, n - 1 - right:n - 1])
            if padding == "odd":
                lo, hi = 2 * first - lo, 2 * last - hi
        return rt.asarray(cat([lo, x2[:, max(k0, 0):min(k1, n)], hi]), self.compute_dtype)

    # ---- forward
    def stft(self, x, p0=None, p1=None, *, k_offset=0, padding="zeros", axis=-1):
        return self.stft_detrend(x, None, p0, p1, k_offset=k_offset, padding=padding, axis=axis)

    def stft_detrend(self, x, detr, p0=None, p1=None, *, k_offset=0, padding="zeros", axis=-1):
        self._check_supported()
        if not (detr is None or detr == "constant"):
            raise NotImplementedError("detr must be None or 'constant'")
        if k_offset != 0:
            raise NotImplementedError("k_offset != 0 is not supported")
        if padding not in padding_type:
            raise ValueError(f"Parameter {padding=} not in {tuple(padding_type)}!")
        rt = get_runtime(x)
        want_numpy = not rt.is_device_array(x)
        if np.iscomplexobj(x) if want_numpy else x.is_complex():
            raise ValueError(f"Complex-valued `x` not allowed for {self.fft_mode=}'! "
                             "Set property `fft_mode` to 'twosided' or 'centered'.")
        dt = self.compute_dtype
        xd = rt.asarray(x, dt)
        nd = len(xd.shape)
        if nd > 1 and axis not in (-1, nd - 1):
            raise NotImplementedError("the transform axis must be the last axis")
        n = int(xd.shape[-1])
        if not (n >= (m2p := self.m_num - self.m_num_mid)):
            raise ValueError(f"len(x)={n} must be >= ceil(m_num/2) = {m2p}!")
        p0, p1 = self.p_range(n, p0, p1)
        lead = tuple(int(s) for s in xd.shape[:-1])
        x2 = rt.reshape(xd, (int(np.prod(lead)) if lead else 1, n))
        if padding == "zeros":
            sig, pad_left = x2, self.m_num_mid - p0 * self.hop
        else:
            sig, pad_left = self._padded(rt, x2, n, p0, p1, padding), 0
        z = _driver.stft(sig, np.asarray(self.win, dtype=np.float64), self.m_num, self.hop, self.mfft, p1 - p0, pad_left,
                         1.0, dt, detrend=detr == "constant", rt=rt, roll=self._roll())
        z = rt.reshape(z, lead + (self.f_pts, p1 - p0))
        return finish(rt, z, want_numpy)

    def spectrogram(self, x, y=None, detr=None, *, p0=None, p1=None, k_offset=0, padding="zeros", axis=-1):
        if y is not None:
            raise NotImplementedError("cross-spectrograms (y is not None) are not supported")
        rt = get_runtime(x)
        want_numpy = not rt.is_device_array(x)
        sx = self.stft_detrend(rt.asarray(x, self.compute_dtype), detr, p0, p1, k_offset=k_offset, padding=padding,
                               axis=axis)
        return finish(rt, _driver.abs_log2(sx, self.compute_dtype, True, eps=0.0, square=True, rt=rt), want_numpy)

    # ---- inverse
    def istft(self, S, k0=0, k1=None, *, f_axis=-2, t_axis=-1):
        self._check_supported()
        rt = get_runtime(S)
        want_numpy = not rt.is_device_array(S)
        nd = len(S.shape)
        if (f_axis % nd, t_axis % nd) != (nd - 2, nd - 1):
            raise NotImplementedError("f_axis / t_axis must be the last two axes")
        if S.shape[-2] != self.f_pts:
            raise ValueError(f"S.shape[f_axis]={S.shape[-2]} must be equal to self.f_pts={self.f_pts} (S.shape={tuple(S.shape)})!")
        n_min = self.m_num - self.m_num_mid
        if not (S.shape[-1] >= (q_num := self.p_num(n_min))):
            raise ValueError(f"S.shape[t_axis]={S.shape[-1]} needs to have at least {q_num} slices (S.shape={tuple(S.shape)})!")
        q_max = int(S.shape[-1]) + self.p_min
        k_max = (q_max - 1) * self.hop + self.m_num - self.m_num_mid
        k1 = k_max if k1 is None else int(k1)
        if not (self.k_min <= k0 < k1 <= k_max):
            raise ValueError(f"(self.k_min={self.k_min}) <= (k0={k0}) < (k1={k1}) <= (k_max={k_max}) is false!")
        if not (num_pts := k1 - k0) >= n_min:
            raise ValueError(f"(k1={k1}) - (k0={k0}) = {num_pts} has to be at least the half the window length {n_min}!")
        q0 = (k0 // self.hop + self.p_min if k0 >= 0 else k0 // self.hop)
        q1 = min(self.p_max(k1), q_max)
        dt = self.compute_dtype
        cdt = "complex64" if dt == "float32" else "complex128"
        sd = rt.asarray(S, cdt)
        lead = tuple(int(s) for s in sd.shape[:-2])
        s3 = rt.reshape(sd, (int(np.prod(lead)) if lead else 1, self.f_pts, int(sd.shape[-1])))
        out = _driver.istft(s3, np.asarray(self.dual_win, dtype=np.float64), self.m_num, self.hop, self.mfft, self._roll(),
                            first_start=self.p_min * self.hop - self.m_num_mid, frame_lo=q0 - self.p_min,
                            frame_hi=q1 - self.p_min, k0=k0, n_out=k1 - k0, dt=dt, rt=rt)
        return finish(rt, rt.reshape(out, lead + (k1 - k0,)), want_numpy)


# return the Short-Time Fourier Transform (STFT) object with default parameters
def get_stft_object_tukey(sample_rate_hz: float, tukey_alpha: float, segment_length: int, overlap_length: int,
                          scaling: str = "magnitude", *, dtype=None) -> ShortTimeFFT:
    """ShortTimeFFT object with a (symmetric) Tukey window, mfft = ceil_power_of_two(segment_length), one-sided
    (reference utilities/short_time_fft.py:19-60, same warnings and fallbacks)."""
    if segment_length < overlap_length:
        print(f"overlap length {overlap_length} must be smaller than segment length {segment_length}"
              " using half of the segment length as the overlap length")
        overlap_length = segment_length // 2
    if tukey_alpha < 0 or tukey_alpha > 1:
        print(f"Warning: Tukey alpha {tukey_alpha} must be between 0 and 1, using 0.25 as the default value")
        tukey_alpha = 0.25
    if scaling not in scaling_type:
        print(f"Warning: scaling {scaling} must be one of {scaling_type}, using 'magnitude' as the default value")
        scaling = "magnitude"
    tukey_window = signal.windows.tukey(segment_length, alpha=tukey_alpha)
    fft_points = round_value(segment_length, "ceil_power_of_two")
    hop_length = segment_length - overlap_length
    stft_obj = ShortTimeFFT(win=tukey_window, hop=hop_length, fs=sample_rate_hz, mfft=fft_points, fft_mode="onesided",
                            scale_to=scaling)
    stft_obj.compute_dtype = dtype_name(dtype)
    return stft_obj


def stft_tukey(timeseries, sample_rate_hz: Union[float, int], tukey_alpha: float, segment_length: int,
               overlap_length: int, scaling: str = "magnitude", padding: str = "zeros", *, dtype=None
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """|STFT| of the per-slice mean-removed signal (reference utilities/short_time_fft.py:64-102).

    :return: frequency bins, time bins, magnitude [..., f, t]
    """
    if padding not in padding_type:
        print(f"Warning: padding {padding} must be one of {padding_type}, using 'zeros' as the default value")
        padding = "zeros"
    stft_obj = get_stft_object_tukey(sample_rate_hz, tukey_alpha, segment_length, overlap_length, scaling, dtype=dtype)
    rt = get_runtime(timeseries)
    want_numpy = not rt.is_device_array(timeseries)
    z = stft_obj.stft_detrend(rt.asarray(timeseries, stft_obj.compute_dtype), "constant", padding=padding)
    mag = finish(rt, _driver.abs_log2(z, stft_obj.compute_dtype, True, eps=0.0, square=2, rt=rt), want_numpy)
    time_bins = np.arange(start=0, stop=stft_obj.delta_t * np.shape(mag)[-1], step=stft_obj.delta_t)
    return stft_obj.f, time_bins, mag


def istft_tukey(stft_to_invert, sample_rate_hz: Union[float, int], tukey_alpha: float, segment_length: int,
                overlap_length: int, scaling: str = "magnitude", *, dtype=None) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse STFT up to the last window that is half filled by the signal (reference
    utilities/short_time_fft.py:106-134).

    :return: timestamps, reconstructed signal
    """
    stft_obj = get_stft_object_tukey(sample_rate_hz, tukey_alpha, segment_length, overlap_length, scaling, dtype=dtype)
    last_window_index = int((np.shape(stft_to_invert)[-1] - 1) * stft_obj.hop)
    timestamps = np.arange(start=0, stop=last_window_index / sample_rate_hz, step=1 / sample_rate_hz)
    return timestamps, stft_obj.istft(stft_to_invert, k1=last_window_index)


def spectrogram_tukey(timeseries, sample_rate_hz: Union[float, int], tukey_alpha: float, segment_length: int,
                      overlap_length: int, scaling: str = "magnitude", padding: str = "zeros", *, dtype=None
                      ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """|STFT|^2 without detrending (reference utilities/short_time_fft.py:138-175).

    :return: frequency bins, time bins, spectrogram [..., f, t]
    """
    if padding not in padding_type:
        print(f"Warning: padding {padding} must be one of {padding_type}, using 'zeros' as the default value")
        padding = "zeros"
    stft_obj = get_stft_object_tukey(sample_rate_hz, tukey_alpha, segment_length, overlap_length, scaling, dtype=dtype)
    spectrogram = stft_obj.spectrogram(x=timeseries, padding=padding)
    time_bins = np.arange(start=0, stop=stft_obj.delta_t * np.shape(spectrogram)[-1], step=stft_obj.delta_t)
    return stft_obj.f, time_bins, spectrogram
