"""
qi_oracle -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A plain numpy (float64) restatement of the time-frequency hot path of ISLA-UH/quantum-inferno
v1.1.3, written band-by-band so it also runs at sizes where the reference's own B x 2N tiles do
not fit in host memory.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it, and only as the checker or as the timed CPU baseline.
The quantum_inferno_b200 package never imports it.

Parity pin: every function here is checked against outputs of the *reference itself*
(imported from /root/reference in the build container by oracle/make_golden.py, which wrote
tests/golden/*.npz) in tests/test_oracle_golden.py.  The reference's own test-suite holds no
vectors for this path except the commented-out band-table KAT of
quantum_inferno/tests/test_scales_dyadic.py:8-21, which is pinned too.

Citations are file:line relative to the reference checkout.  The only third-party arithmetic
the reference uses on this path is scipy (>=1.15, pyproject.toml:20): signal.fftconvolve,
signal.stft / welch, fft.fft/ifft/rfft; their published algorithms are restated here with
numpy.fft only.
"""
import numpy as np

EPS64 = float(np.finfo(np.float64).eps)      # scales_dyadic.py:16
EPS32 = float(np.finfo(np.float32).eps)      # scales_dyadic.py:17
G2 = 2.0                                     # scales_dyadic.py:53
G3 = 10.0 ** 0.3                             # scales_dyadic.py:54
M_OVER_N = 0.75 * np.pi                      # scales_dyadic.py:21
ORDER_MIN = 0.75                             # scales_dyadic.py:97
T0S = 1e-42                                  # scales_dyadic.py:57
VALID_ORDERS = [0.75, 1, 1.5, 3, 6, 12, 24, 48]   # scales_dyadic.py:102


# ----------------------------------------------------------------------------- scales_dyadic
def order_checked(order):
    """scales_dyadic.py:105-122 (print side effect dropped)."""
    order = np.abs(order)
    return ORDER_MIN if order < ORDER_MIN else order


def cycles_from_order(order):
    """scales_dyadic.py:125-141: M = 0.75*pi*N."""
    return M_OVER_N * order_checked(order)


def scale_from_frequency_hz(order, f_hz, fs):
    """scales_dyadic.py:167-180 -> (scale_atom, omega)."""
    omega = 2.0 * np.pi * f_hz / fs
    return cycles_from_order(order) / omega, omega


def log_frequency_hz_from_fft_points(fs, n_points, order, ref_hz=1.0, base=G3):
    """scales_dyadic.py:355-393: ascending band centres ref*base^(-j/N)."""
    log2_len = int(np.ceil(np.log2(n_points)))
    mult = cycles_from_order(order)
    n_over_log2g = order_checked(order) / np.log2(base)
    log2_mult = np.log2(mult)
    log2_ref = np.log2(fs / ref_hz)
    j_lo = int(np.ceil(n_over_log2g * (np.log2(2.5) - log2_ref)))
    j_hi = int(np.floor(n_over_log2g * (log2_len - log2_mult - log2_ref)))
    j = np.arange(j_lo, j_hi + 1)
    return np.flip(ref_hz * base ** (-j / order))


def band_intervals_periods(order_in, base_in, ref_in, low_in, high_in):
    """scales_dyadic.py:241-352 (warnings dropped)."""
    ref, low, high, base, order = np.absolute([ref_in, low_in, high_in, base_in, order_in])
    if not (base == G3 or base == G2) and base < 1.0:
        base = G2
    if order not in VALID_ORDERS and order < 0.75:
        order = 1
    edge = base ** (1.0 / (2.0 * order))
    if low < T0S:
        low = T0S / edge
    if high < low:
        low = high / base
    if high == low:
        high *= edge
        low /= edge
    n_max = np.round(order * np.log(high / ref) / np.log(base))
    n_min = np.floor(order * np.log(low / ref) / np.log(base))
    c_min = ref * np.power(base, n_min / order)
    if (c_min < low) or (c_min / edge < low - EPS64):
        n_min += 1
    if n_max < n_min:
        n_max = np.floor(np.log10(high) / np.log10(base))
        n_min = n_max - order
    band = np.arange(n_min, n_max + 1)
    centre = ref * np.power(base * np.ones(band.shape), band / order)
    start = centre / edge
    end = centre * edge
    return order, base, band, ref, (start + end) / 2.0, centre, start, end


def band_frequency_low_high(order_in, base_in, ref_in, f_low, f_high, fs):
    """scales_dyadic.py:183-238."""
    s_ref = 1 / ref_in
    s_nyq = 2 / fs
    s_low = 1 / f_high
    if s_low < s_nyq:
        s_low = s_nyq
    s_high = 1 / f_low
    order, base, band, ref, _, centre, start, end = band_intervals_periods(order_in, base_in, s_ref, s_low, s_high)
    f_end = 1 / start
    f_start = 1 / end
    return order, base, -band, 1 / ref, (f_end + f_start) / 2.0, 1 / centre, f_start, f_end


# ----------------------------------------------------------------------------- styx_cwt
def wavelet_amplitude(scale):
    """styx_cwt.py:29-40 (kept un-simplified as the source asks)."""
    a_norm = (np.pi * scale ** 2) ** (-1 / 4)
    a_spect = (4 * np.pi * scale ** 2) ** (-1 / 4) * a_norm
    return a_norm, a_spect


def gabor_atom_centered(order, n_points, f_hz, fs, dictionary_type="norm"):
    """One band of styx_cwt.py:113-144 / :68-110: atom centred at (n_points-1)/2 samples.
    Returns (atom[n_points] complex128, scale, omega, amp)."""
    t = np.arange(n_points) / fs
    x = fs * (t - t[-1] / 2.0)                       # styx_cwt.py:65,132,135
    scale, omega = scale_from_frequency_hz(order, f_hz, fs)
    gabor = np.exp(-0.5 * (x / scale) ** 2) * np.exp(1j * omega * x)    # styx_cwt.py:107
    a_norm, a_spect = wavelet_amplitude(scale)
    if dictionary_type == "spect":
        amp = a_spect
    elif dictionary_type == "unit":
        amp = 1.0
    else:
        amp = a_norm
    return amp * gabor, scale, omega, amp


def gabor_atom_centered_arbiter(order, n_points, f_hz, fs, dictionary_type="norm"):
    """ARBITER, not a restatement: the same atom as gabor_atom_centered with the time axis and the carrier phase carried
    in 80-bit long double (the axis fs * (k / fs - t[-1] / 2) is exactly k - (n_points - 1) / 2 in exact arithmetic).
    The reference's float64 axis and its float64 product omega * x carry a phase error of about eps64 * omega * N / 2
    (1.4e-10 rad for the top band of a 2^20-sample record, 2.3e-9 at 2^24), which is above the 1e-10 fp64 tolerance;
    tests on long records bound the distance to gabor_atom_centered by that figure and the distance to this arbiter by
    the tolerance itself (the same device as the long-double arbiter of the IIR filters)."""
    ld = np.longdouble
    x = np.arange(n_points, dtype=ld) - ld(n_points - 1) / ld(2)
    scale, omega = scale_from_frequency_hz(order, f_hz, fs)
    two_pi = ld(2) * np.arctan2(ld(0), ld(-1))
    phase = ld(omega) * x
    phase -= two_pi * np.rint(phase / two_pi)
    gabor = np.exp(-0.5 * (x / ld(scale)) ** 2) * (np.cos(phase) + 1j * np.sin(phase))
    a_norm, a_spect = wavelet_amplitude(scale)
    amp = a_spect if dictionary_type == "spect" else (1.0 if dictionary_type == "unit" else a_norm)
    return (amp * gabor).astype(np.complex128), scale, omega, amp


def reference_axis_phase_noise(order, n_points, f_hz, fs):
    """Bound on the reference's own carrier-phase rounding for one band (see gabor_atom_centered_arbiter):
    the float64 axis (three roundings at magnitude N/2 samples) times omega, plus the rounding of omega * x."""
    _, omega = scale_from_frequency_hz(order, f_hz, fs)
    return 4.0 * np.finfo(np.float64).eps * float(omega) * n_points / 2.0


def cwt_band(sig_fft_2n, order, n_points, f_hz, fs, dictionary_type="norm", arbiter=False):
    """One row of styx_cwt.py:195-196: fftconvolve(sig, conj(fliplr(atom)), 'same'), i.e.
    ifft(fft(x,2N)*fft(h,2N))[(N-1)//2 : (N-1)//2+N] (scipy _freq_domain_conv + _centered)."""
    make = gabor_atom_centered_arbiter if arbiter else gabor_atom_centered
    atom, _, _, _ = make(order, n_points, f_hz, fs, dictionary_type)
    h = np.conj(atom[::-1])
    full = np.fft.ifft(sig_fft_2n * np.fft.fft(h, 2 * n_points))
    s = (n_points - 1) // 2
    return full[s:s + n_points]


def cwt_complex_any_scale_pow2(order, sig, fs, cwt_type="fft", dictionary_type="norm"):
    """styx_cwt.py:147-198 -> (frequency_hz[B], time_s[N], cwt[B,N] complex128)."""
    sig = np.asarray(sig, dtype=np.float64)
    n = len(sig)
    freqs = log_frequency_hz_from_fft_points(fs, n, order)
    xf = np.fft.fft(sig, 2 * n)
    out = np.empty((len(freqs), n), dtype=np.complex128)
    for b, f in enumerate(freqs):
        out[b] = cwt_band(xf, order, n, f, fs, dictionary_type)
    return freqs, np.arange(n) / fs, out


# ----------------------------------------------------------------------------- styx_stx
def stx_complex_any_scale_pow2(order, sig, fs):
    """styx_stx.py:195-236."""
    sig = np.asarray(sig, dtype=np.float64)
    n = len(sig)
    freqs = log_frequency_hz_from_fft_points(fs, n, order)
    xf = np.fft.fft(sig)
    xf2 = np.concatenate([xf, xf])
    f_fft = np.fft.fftfreq(n, 1 / fs)
    w_fft = 2 * np.pi * f_fft / fs
    w_stx = 2 * np.pi * freqs / fs
    sigma = cycles_from_order(order) / w_stx
    out = np.empty((len(freqs), n), dtype=np.complex128)
    for b, f in enumerate(freqs):
        idx = int(np.abs(f_fft - f).argmin())                      # styx_stx.py:233 (first minimum)
        win = np.exp(-0.5 * (sigma[b] ** 2.0) * (w_fft ** 2.0))    # styx_stx.py:223-225
        out[b] = np.fft.ifft(xf2[idx:idx + n] * win)
    return freqs, np.arange(n) / fs, out


def stx_shift_indices(order, n, fs):
    """Bit-exact integer part of styx_stx.py:231-233."""
    freqs = log_frequency_hz_from_fft_points(fs, n, order)
    f_fft = np.fft.fftfreq(n, 1 / fs)
    return np.array([int(np.abs(f_fft - f).argmin()) for f in freqs], dtype=np.int64)


def tfr_stx_fft(sig, dt, order=8.0, n_fft_in=None, frequency_min=None, frequency_max=None, frequency_step=None,
                factor_q=0.0, power_p=0.0, power_r=1.0, is_geometric=False, is_inferno=False,
                base=G3, ref=1.0):
    """styx_stx.py:52-192, on its working domain n_fft_in == len(sig) == 2^m."""
    sig = np.asarray(sig, dtype=np.float64)
    n = sig.shape[-1]
    if n_fft_in is None:
        raise TypeError("'<' not supported between instances of 'NoneType' and 'int'")   # styx_stx.py:30
    if n_fft_in < n:
        raise ValueError(f"n_fft cannot be smaller than signal size. Got {n_fft_in} < {n}.")
    if n_fft_in != n:
        raise TypeError("unsupported padding path (styx_stx.py:44)")
    fs = 1 / dt
    cycles = 12.0 / 5.0 * order
    xf = np.fft.fft(sig)
    xf2 = np.concatenate([xf, xf], axis=-1)
    f_fft = np.fft.fftfreq(n, dt)
    w_fft = 2 * np.pi * f_fft / fs
    f_min_nth = cycles / (n / fs)
    if frequency_min is None:
        frequency_min = f_min_nth
    if frequency_max is None:
        frequency_max = fs / 2.0
    i0 = np.abs(f_fft - frequency_min).argmin()
    i1 = np.abs(f_fft - frequency_max).argmin()
    f_start, f_stop = f_fft[i0], f_fft[i1]
    if frequency_step is None:
        frequency_step = (frequency_max - frequency_min) * 2.0 / len(f_fft)
    f_stx = np.arange(f_start, f_stop, frequency_step)
    if is_geometric is True:
        if is_inferno is True:
            f_stx = band_frequency_low_high(order, base, ref, f_start, f_stop, fs)[5]
        else:
            n_bands = int(np.log2(f_stop / f_start) * order)
            f_stx = np.logspace(np.log2(f_start), np.log2(f_stop), num=n_bands, base=base)
    nb = len(f_stx)
    f_snap = np.empty(nb)
    windows = np.empty((nb, n), dtype=np.complex128)
    tfr = np.empty((nb, n), dtype=np.complex128)
    psd = np.empty((nb, n))
    for b, f in enumerate(f_stx):
        idx = np.abs(f_fft - f).argmin()
        f_snap[b] = f_fft[idx]
        w_sx = 2 * np.pi * f_snap[b] / fs
        if w_sx == 0.0:
            raise TypeError("object of type 'int' has no len()")     # styx_stx.py:173
        sigma = cycles / w_sx * ((1 + factor_q * (w_sx ** power_p)) * (w_sx ** (1 - power_r)))
        windows[b] = np.exp(-0.5 * (sigma ** 2.0) * (w_fft ** 2.0))
        tfr[b] = np.fft.ifft(xf2[idx:idx + n] * windows[b])
        psd[b] = np.abs(tfr[b]) ** 2 + EPS64
    return tfr, psd, f_stx, f_snap, windows


# ----------------------------------------------------------------------------- styx_fft (scipy.signal.stft / welch)
def window_periodic(kind, param, n):
    """scipy.signal.get_window((kind, param), n) with fftbins=True: the symmetric n+1 window minus
    its last sample.  kind in {'tukey','gaussian'} (styx_fft.py:178,218,257)."""
    m = n + 1
    k = np.arange(m)
    if kind == "tukey":
        alpha = param
        if alpha <= 0:
            w = np.ones(m)
        elif alpha >= 1.0:
            w = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (m - 1))      # hann
        else:
            width = int(np.floor(alpha * (m - 1) / 2.0))
            k1 = k[0:width + 1]
            k3 = k[m - width - 1:]
            w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * k1 / alpha / (m - 1))))
            w2 = np.ones(m - 2 * width - 2)
            w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * k3 / alpha / (m - 1))))
            w = np.concatenate((w1, w2, w3))
    elif kind == "gaussian":
        x = k - (m - 1.0) / 2.0
        w = np.exp(-0.5 * (x / param) ** 2)
    else:
        raise ValueError(kind)
    return w[:-1]


def _frames(x, nperseg, noverlap):
    step = nperseg - noverlap
    nfr = (x.shape[-1] - noverlap) // step
    idx = np.arange(nperseg)[None, :] + step * np.arange(nfr)[:, None]
    return x[..., idx]                                             # [..., frame, sample]


def stft_scipy(x, fs, window, nperseg, noverlap, nfft):
    """scipy.signal.stft(x, fs, window, nperseg, noverlap, nfft, detrend='constant',
    return_onesided=True, boundary='zeros', padded=True, axis=-1, scaling='spectrum')
    -- _spectral_helper(mode='stft'), as called at styx_fft.py:175-187 / :215-227."""
    x = np.asarray(x, dtype=np.float64)
    win = window_periodic(window[0], window[1], nperseg)
    half = nperseg // 2
    pad = [(0, 0)] * (x.ndim - 1)
    x = np.pad(x, pad + [(half, half)])                            # boundary='zeros'
    step = nperseg - noverlap
    nadd = (-(x.shape[-1] - nperseg) % step) % nperseg             # padded=True
    x = np.pad(x, pad + [(0, nadd)])
    fr = _frames(x, nperseg, noverlap)
    fr = fr - fr.mean(axis=-1, keepdims=True)                      # detrend='constant'
    z = np.fft.rfft(fr * win, n=nfft, axis=-1) * (1.0 / win.sum())
    t = np.arange(nperseg / 2, x.shape[-1] - nperseg / 2 + 1, step) / float(fs) - (nperseg / 2) / fs
    f = np.fft.rfftfreq(nfft, 1 / fs)
    return f, t, np.moveaxis(z, -1, -2)                            # [..., freq, time]


def _pow2_defaults(segment, overlap, nfft):
    if nfft is None:
        nfft = int(2 ** np.ceil(np.log2(segment)))
    if overlap is None:
        overlap = int(segment / 2)
    return overlap, nfft


def stft_complex_pow2(sig, fs, segment_points, overlap_points=None, nfft_points=None, alpha=0.25):
    """styx_fft.py:152-187."""
    overlap_points, nfft_points = _pow2_defaults(segment_points, overlap_points, nfft_points)
    return stft_scipy(sig, fs, ("tukey", alpha), segment_points, overlap_points, nfft_points)


def gtx_complex_pow2(sig, fs, segment_points, gaussian_sigma=None, overlap_points=None, nfft_points=None):
    """styx_fft.py:190-227."""
    overlap_points, nfft_points = _pow2_defaults(segment_points, overlap_points, nfft_points)
    if gaussian_sigma is None:
        gaussian_sigma = int(segment_points / 4)
    return stft_scipy(sig, fs, ("gaussian", gaussian_sigma), segment_points, overlap_points, nfft_points)


def welch_power_pow2(sig, fs, segment_points, nfft_points=None, overlap_points=None, alpha=0.25):
    """styx_fft.py:230-266: scipy.signal.welch(scaling='spectrum', average='mean', detrend='constant')
    = _spectral_helper(mode='psd', boundary=None, padded=False) then mean over segments."""
    overlap_points, nfft_points = _pow2_defaults(segment_points, overlap_points, nfft_points)
    x = np.asarray(sig, dtype=np.float64)
    win = window_periodic("tukey", alpha, segment_points)
    fr = _frames(x, segment_points, overlap_points)
    fr = fr - fr.mean(axis=-1, keepdims=True)
    z = np.fft.rfft(fr * win, n=nfft_points, axis=-1)
    p = (np.conjugate(z) * z).real * (1.0 / win.sum() ** 2)
    if nfft_points % 2:
        p[..., 1:] *= 2
    else:
        p[..., 1:-1] *= 2
    return np.fft.rfftfreq(nfft_points, 1 / fs), p.mean(axis=-2)


def get_num_points_ceil_log2(fs, duration_s):
    """utilities/calculations.py:187-205 with rounding_type='ceil', output_unit='log2'."""
    return int(np.ceil(np.log2(fs * duration_s)))


def stft_from_sig(sig, fs, order, center_frequency_hz=None, octaves_below_center=4):
    """styx_fft.py:14-57 -> (stft, stft_bits, time_s, frequency_hz)."""
    if center_frequency_hz is None:
        center_frequency_hz = fs * 0.075
    f_ave = center_frequency_hz / octaves_below_center
    nd = 2 ** get_num_points_ceil_log2(fs, cycles_from_order(order) / f_ave)
    if len(sig) < nd:
        raise ValueError(f"Signal length: {len(sig)} is less than time_fft_nd: {nd}")
    f, t, z = stft_complex_pow2(sig, fs, nd, alpha=1.0)
    z = z * (2 * np.sqrt(np.pi) / nd)
    return z, np.log2(np.abs(z) + EPS64), t, f


# ----------------------------------------------------------------------------- cwt_atoms
def chirp_mqg_from_n(order, index_shift=0, base=G2):
    """cwt_atoms.py:122-144 -> (cycles M, Q, gamma)."""
    if order < 0.7:
        order = 3.0
    edge = base ** (1.0 / 2.0 / order)
    q = 1.0 / (edge - 1.0 / edge)
    gamma = np.sqrt(np.log(2)) * (1 - np.log(2) * (index_shift / np.pi) ** 2) ** (-0.5)
    return 2 * q * gamma, q, gamma


def chirp_atom_centered(order, n_points, f_hz, fs, index_shift=0, base=G2, dictionary_type="norm"):
    """cwt_atoms.py:303-340 + :16-50 for one band."""
    t = np.arange(n_points) / fs
    x = fs * (t - t[-1] / 2.0)
    m, _, gamma = chirp_mqg_from_n(order, index_shift, base)
    scale = m * fs / f_hz / (2.0 * np.pi)                                      # cwt_atoms.py:158
    p = (1 - 1j * index_shift * gamma / np.pi) / (2 * scale ** 2)               # cwt_atoms.py:211
    a_norm = 1 / np.pi ** 0.25 * 1 / np.sqrt(scale)                             # cwt_atoms.py:224
    a_spect = np.sqrt(np.abs(p) / np.pi)                                        # cwt_atoms.py:225
    atom = np.exp(-p * x ** 2) * np.exp(1j * m * x / scale)
    return (a_norm if dictionary_type == "norm" else a_spect) * atom


def cwt_chirp_complex(order, sig, f_low, fs, f_high=1.0e42, cwt_type="fft", index_shift=0, ref=1.0, base=G2,
                      dictionary_type="norm"):
    """cwt_atoms.py:343-444 -> (cwt, cwt_bits, time_s, frequency_hz)."""
    sig = np.asarray(sig, dtype=np.float64)
    n = len(sig)
    if f_high > fs / 2.0:
        f_high = fs / 2.0
    order_nth, base_out, _, _, _, f_desc, _, _ = band_frequency_low_high(order, base, ref, f_low, f_high, fs)
    nb = len(f_desc)
    out = np.empty((nb, n), dtype=np.complex128)
    if cwt_type == "fft":
        xf = np.fft.fft(sig)
        for b in range(nb):
            a = chirp_atom_centered(order_nth, n, f_desc[b], fs, index_shift, base_out, dictionary_type)
            raw = np.fft.ifft(xf * np.conj(np.fft.fft(a)))
            out[b] = np.append(raw[n // 2:], raw[0:n // 2])                   # cwt_atoms.py:421
    elif cwt_type == "conv":
        xf2 = np.fft.fft(sig, 2 * n)
        for b in range(nb):
            a = chirp_atom_centered(order_nth, n, f_desc[b], fs, index_shift, base_out, dictionary_type)
            # scipy.signal.convolve(sig, conj(a)[::-1], 'same') == linear convolution, centred slice
            full = np.fft.ifft(xf2 * np.fft.fft(np.conj(a)[::-1], 2 * n))
            s = (n - 1) // 2
            out[b] = full[s:s + n]
    else:
        raise ValueError(f"Incorrect cwt_type: {cwt_type} specified in cwt_chirp_complex")
    out = np.flipud(out)
    return out, np.log2(np.abs(out) + EPS64), np.arange(n) / fs, np.flip(f_desc)


def cwt_chirp_from_sig(sig, fs, order=3, cwt_type="fft", index_shift=0, ref=1.0, base=G2, dictionary_type="norm"):
    """cwt_atoms.py:447-486."""
    m, _, _ = chirp_mqg_from_n(order, index_shift, base)
    f_min = 1 / ((len(sig) / fs) / m)                                          # cwt_atoms.py:253-255
    return cwt_chirp_complex(order, sig, f_min, fs, fs / 2.0, cwt_type, index_shift, ref, base, dictionary_type)


# ----------------------------------------------------------------------------- tfr_info
class ShannonStft:
    """tfr_info.py:203-228."""

    def __init__(self, pdf, deg_free):
        self.info = -np.log2(pdf + EPS64)
        self.shannon_bits = pdf * self.info
        self.ref_bits = np.log2(deg_free) / deg_free
        self.isnr = np.log2(deg_free) - self.info
        self.esnr = self.shannon_bits / self.ref_bits


def shannon_stft_from_tfr_power(p):
    """tfr_info.py:231-236."""
    return ShannonStft(p / np.sum(p), p.shape[0] * p.shape[1])


def shannon_stft_per_time(p):
    """tfr_info.py:239-248: eps is added to the reciprocal of the column sums."""
    return ShannonStft((1 / np.sum(p, axis=0) + EPS64)[None, :] * p, p.shape[0])


def shannon_stft_per_freq(p):
    """tfr_info.py:251-260."""
    return ShannonStft((1 / np.sum(p, axis=1) + EPS64)[:, None] * p, p.shape[1])


def scale_power_bits(p):
    """tfr_info.py:65-79."""
    b = np.log2(p + EPS64)
    return b - np.max(b)


def power_dynamics_scaled_bits(p):
    """tfr_info.py:82-94."""
    return scale_power_bits(p), scale_power_bits(np.sum(p, axis=0)), scale_power_bits(np.sum(p, axis=1))


class Shannon1D:
    """tfr_info.py:97-135 (EPS32 inside the log)."""

    def __init__(self, marginal):
        self.marginal = marginal
        self.info = -np.log2(marginal + EPS32)
        self.entropy = marginal * self.info
        self.ref_entropy = np.log2(len(marginal)) / len(marginal)
        self.isnr = np.log2(len(self.info)) - self.info
        self.esnr = self.entropy / self.ref_entropy


def shannon_tdr(sig):
    """tfr_info.py:138-151."""
    s = sig / np.sqrt(np.sum(sig ** 2))
    out = Shannon1D(s ** 2)
    out.sig = s
    return out


def shannon_fft(sig):
    """tfr_info.py:163-181."""
    z = np.fft.rfft(sig)
    p = np.abs(z) ** 2
    out = Shannon1D(p / np.sum(p))
    out.sig = z
    out.angle_rads = np.unwrap(np.angle(z))
    out.frequency = np.arange(len(z)) / len(z) / 2.0
    return out


# ----------------------------------------------------------------------------- fused north-star quantity
def cwt_power_entropy(order, sig, fs, dictionary_type="norm"):
    """The north-star composite: styx_cwt.py:147-198 -> |.|^2 -> tfr_info.py:231-236 and :251-260,
    evaluated band-by-band with a two-pass total so it runs at any record length.
    Returns dict(freq, power[B,N], info[B,N], band_sum[B], total, entropy_bits (sum of shannon_bits),
    band_entropy_bits[B] (per-frequency ShannonStftPerFreq row sums))."""
    freqs, _, c = cwt_complex_any_scale_pow2(order, sig, fs, dictionary_type=dictionary_type)
    p = np.abs(c) ** 2
    g = shannon_stft_from_tfr_power(p)
    pf = shannon_stft_per_freq(p)
    return dict(freq=freqs, power=p, info=g.info, band_sum=p.sum(axis=1), total=float(p.sum()),
                entropy_bits=float(g.shannon_bits.sum()), band_entropy_bits=pf.shannon_bits.sum(axis=1))


# ---------------------------------------------------------------- utilities/short_time_fft.py
# The reference delegates to scipy.signal.ShortTimeFFT (scipy >= 1.15, pyproject.toml:20); its published algorithm
# (scipy/signal/_short_time_fft.py: _pre_padding, _post_padding, _x_slices, stft_detrend, _fft_func, _ifft_func,
# istft, _calc_dual_canonical_window) is restated here with numpy only.
def tukey_symmetric(m, alpha):
    """scipy.signal.windows.tukey(m, alpha, sym=True) as called at utilities/short_time_fft.py:51."""
    if m == 1 or alpha <= 0:
        return np.ones(m)
    if alpha >= 1.0:
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(m) / (m - 1))
    n = np.arange(m)
    width = int(np.floor(alpha * (m - 1) / 2.0))
    n1, n2, n3 = n[0:width + 1], n[width + 1:m - width - 1], n[m - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (m - 1))))
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (m - 1))))
    return np.concatenate((w1, np.ones(n2.shape), w3))


class StftTukeyPlan:
    """Bookkeeping of get_stft_object_tukey (utilities/short_time_fft.py:19-60) + ShortTimeFFT's slice ranges."""

    def __init__(self, fs, alpha, segment_length, overlap_length, scaling="magnitude"):
        if segment_length < overlap_length:
            overlap_length = segment_length // 2
        if alpha < 0 or alpha > 1:
            alpha = 0.25
        if scaling not in ("magnitude", "psd", None):
            scaling = "magnitude"
        self.fs, self.m, self.hop = float(fs), int(segment_length), int(segment_length - overlap_length)
        self.mfft = 2 ** int(np.ceil(np.log2(segment_length)))        # round_value(..., "ceil_power_of_two")
        self.mid = self.m // 2
        win = tukey_symmetric(self.m, alpha)
        if scaling == "magnitude":                                    # ShortTimeFFT.scale_to: the window itself is scaled
            win = win / abs(win.sum())
        elif scaling == "psd":
            win = win / np.sqrt((win ** 2).sum() / (1.0 / self.fs))
        self.win = win
        w2 = win ** 2
        dd = w2.copy()                                                # canonical dual window
        for k in range(self.hop, self.m, self.hop):
            dd[k:] += w2[:-k]
            dd[:-k] += w2[k:]
        self.dual = win / dd
        self.roll = self.mid % self.m                                 # phase_shift = 0
        self.f = np.fft.rfftfreq(self.mfft, 1.0 / self.fs)
        self.delta_t = self.hop / self.fs
        # _pre_padding: move the window left until it no longer overlaps t >= 0
        n0 = -self.mid
        for p_, n_ in enumerate(range(n0, n0 - self.m - 1, -self.hop)):
            n_next = n_ - self.hop
            if n_next + self.m <= 0 or np.all(w2[n_next:] == 0):
                self.k_min, self.p_min = n_, -p_
                break

    def p_max(self, n):
        w2 = self.win ** 2
        q1 = n // self.hop
        k1 = q1 * self.hop - self.mid
        for q_, k_ in enumerate(range(k1, n + self.m, self.hop), start=q1):
            n_next = k_ + self.hop
            if n_next >= n or np.all(w2[:n - n_next] == 0):
                return q_ + 1
        raise RuntimeError("unreachable")

    def slices(self, x, padding):
        n = len(x)
        p0, p1 = self.p_min, self.p_max(n)
        k0 = p0 * self.hop - self.mid
        k1 = k0 + (p1 - p0) * self.hop + self.m
        kw = {"zeros": dict(mode="constant"), "edge": dict(mode="edge"), "even": dict(mode="reflect", reflect_type="even"),
              "odd": dict(mode="reflect", reflect_type="odd")}[padding]
        x1 = np.pad(x[max(k0, 0):min(k1, n)], (-min(k0, 0), max(k1 - n, 0)), **kw)
        return [x1[k:k + self.m] for k in range(0, (p1 - p0) * self.hop, self.hop)]

    def stft(self, x, detrend_constant=False, padding="zeros"):
        cols = []
        for s in self.slices(np.asarray(x, dtype=np.float64), padding):
            if detrend_constant:
                s = s - s.mean()
            z = np.zeros(self.mfft)
            z[:self.m] = s * self.win
            cols.append(np.fft.rfft(np.roll(z, -self.roll)))
        return np.array(cols).T

    def istft(self, spec, k1):
        q_max = spec.shape[-1] + self.p_min
        q0, q1 = self.p_min, min(self.p_max(k1), q_max)
        x = np.zeros(k1 + self.m)
        for q in range(q0, q1):
            xs = np.roll(np.fft.irfft(spec[:, q - self.p_min], n=self.mfft), self.roll)[:self.m] * self.dual
            i0 = q * self.hop - self.mid
            j0 = max(0, -i0)
            x[i0 + j0:i0 + self.m] += xs[j0:]
        return x[:k1]


def stft_tukey(x, fs, alpha, segment_length, overlap_length, scaling="magnitude", padding="zeros"):
    """utilities/short_time_fft.py:64-102 -> (frequency bins, time bins, |stft_detrend(x, 'constant')|)."""
    if padding not in ("zeros", "edge", "even", "odd"):
        padding = "zeros"
    plan = StftTukeyPlan(fs, alpha, segment_length, overlap_length, scaling)
    mag = np.abs(plan.stft(x, True, padding))
    return plan.f, np.arange(0, plan.delta_t * mag.shape[1], plan.delta_t), mag


def spectrogram_tukey(x, fs, alpha, segment_length, overlap_length, scaling="magnitude", padding="zeros"):
    """utilities/short_time_fft.py:138-175."""
    if padding not in ("zeros", "edge", "even", "odd"):
        padding = "zeros"
    plan = StftTukeyPlan(fs, alpha, segment_length, overlap_length, scaling)
    s = plan.stft(x, False, padding)
    sp = s.real ** 2 + s.imag ** 2
    return plan.f, np.arange(0, plan.delta_t * sp.shape[1], plan.delta_t), sp


def istft_tukey(spec, fs, alpha, segment_length, overlap_length, scaling="magnitude"):
    """utilities/short_time_fft.py:106-134."""
    plan = StftTukeyPlan(fs, alpha, segment_length, overlap_length, scaling)
    last = int((spec.shape[1] - 1) * plan.hop)
    return np.arange(0, last / fs, 1 / fs), plan.istft(spec, last)


# ============================================================================ after the path
# utilities/sampling.py:13-50,87-120 and utilities/picker.py:34-53,108-149 restated with numpy; scipy.signal.find_peaks
# (scipy >= 1.15, scipy/signal/_peak_finding.py::find_peaks, _peak_finding_utils.pyx::_local_maxima_1d and
# ::_select_by_peak_distance) restated for the two arguments the reference passes (height, distance).
SUBSAMPLE_METHODS = ["average", "median", "max", "min", "nth"]     # utilities/sampling.py:10


def subsample_2d(array, factor, method="nth"):
    """utilities/sampling.py:87-120: group reductions along the second axis, remainder dropped (kept for "nth")."""
    array = np.asarray(array)
    if factor < 2:
        return array
    if method not in SUBSAMPLE_METHODS:
        method = "nth"
    if method == "nth":
        return array[:, ::factor]
    n_out = array.shape[1] // factor
    groups = array[:, :n_out * factor].reshape(array.shape[0], n_out, factor)
    if method == "average":
        return groups.mean(axis=2)
    if method == "median":
        srt = np.sort(groups, axis=2)                               # NaNs sort last
        mid = srt[:, :, factor // 2] if factor % 2 else (srt[:, :, factor // 2 - 1] + srt[:, :, factor // 2]) / 2
        return np.where(np.isnan(srt[:, :, -1]), np.nan, mid).astype(array.dtype)
    return groups.max(axis=2) if method == "max" else groups.min(axis=2)


def subsample(timeseries, sample_rate_hz, factor, method="nth"):
    """utilities/sampling.py:13-50."""
    if factor < 2:
        return timeseries, sample_rate_hz
    return subsample_2d(np.asarray(timeseries)[None, :], factor, method)[0], sample_rate_hz / factor


def local_maxima_1d(x):
    """scipy _local_maxima_1d: midpoints of the flat tops that rise strictly on the left and fall strictly on the
    right; the first and last sample never count.  Restated on runs of equal samples."""
    x = np.asarray(x)
    n = x.shape[0]
    if n < 3:
        return np.zeros(0, dtype=np.intp)
    start = np.concatenate(([0], np.flatnonzero(x[1:] != x[:-1]) + 1))      # first index of every run
    end = np.concatenate((start[1:], [n])) - 1                              # last index of every run
    v = x[start]
    inner = np.arange(1, len(start) - 1)
    is_peak = (v[inner - 1] < v[inner]) & (v[inner + 1] < v[inner])
    runs = inner[is_peak]
    return ((start[runs] + end[runs]) // 2).astype(np.intp)


def select_by_peak_distance(peaks, priority, distance):
    """scipy _select_by_peak_distance: peaks in order of falling priority remove their neighbours nearer than
    ceil(distance) samples."""
    n = len(peaks)
    d = int(np.ceil(distance))
    keep = np.ones(n, dtype=bool)
    order = np.argsort(priority)
    for i in range(n - 1, -1, -1):
        j = order[i]
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < d:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < d:
            keep[k] = False
            k += 1
    return keep


def find_peaks(x, height=None, distance=None):
    """scipy.signal.find_peaks(x, height=, distance=)[0] (conditions applied in scipy's order: height, distance)."""
    x = np.asarray(x)
    peaks = local_maxima_1d(x)
    if height is not None:
        peaks = peaks[x[peaks] >= height]
    if distance is not None:
        peaks = peaks[select_by_peak_distance(peaks, x[peaks], distance)]
    return peaks


def to_log2_with_epsilon(x):
    """utilities/rescaling.py:13-20."""
    return np.log2(np.abs(x) + EPS64)


def scale_signal_by_extraction_type(sig, extraction_type="sigmax"):
    """utilities/picker.py:34-53."""
    sig = np.asarray(sig)
    if extraction_type == "sigmin":
        return sig / np.nanmin(sig)
    if extraction_type == "sigabs":
        return sig / np.nanmax(np.abs(sig))
    if extraction_type == "log2":
        return to_log2_with_epsilon(sig)
    if extraction_type == "log2max":
        return to_log2_with_epsilon(sig) / np.nanmax(to_log2_with_epsilon(sig))
    return sig / np.nanmax(sig)                                           # "sigmax" and the invalid-type default


def find_peaks_by_extraction_type(timeseries, extraction_type="sigmax", height=0.7):
    """utilities/picker.py:108-120."""
    return find_peaks(scale_signal_by_extraction_type(timeseries, extraction_type), height=height)


def find_peaks_with_bits(timeseries, sample_rate_hz, scaling_type="amplitude", threshold_bits=1,
                         time_distance_seconds=0.1):
    """utilities/picker.py:123-149."""
    bits = to_log2_with_epsilon(timeseries)
    if scaling_type == "log2":
        height = np.max(bits) - threshold_bits
    else:
        height = np.max(timeseries) - 2 ** threshold_bits
    return find_peaks(bits, height=height, distance=int(time_distance_seconds * sample_rate_hz))


# ============================================================================ before the path
# styx_fft.py:60-149 (butter_bandpass / butter_highpass / butter_lowpass), synth/synthetic_signals.py:180-192
# (antialias_half_nyquist) and utilities/picker.py:56-76 (apply_bandpass) call scipy.signal.filtfilt / sosfiltfilt
# (scipy >= 1.15).  Their published algorithm is restated here sample by sample: scipy/signal/_lfilter.c.in (direct
# form II transposed), _sosfilt.pyx (biquad cascade), _signaltools.py::lfilter_zi / sosfilt_zi / filtfilt /
# sosfiltfilt (method="pad", padtype="odd"), windows/_windows.py::tukey.  `dtype=np.longdouble` runs the SAME
# recursion in 80-bit arithmetic: the arbiter when float64 evaluations of these ill-conditioned recursions disagree.
# The filter taps themselves come from scipy.signal.butter (design formulas, not record arithmetic).
def tukey(m, alpha):
    """scipy.signal.windows.tukey(m, alpha, sym=True)."""
    if m == 1 or alpha <= 0:
        return np.ones(m)
    k = np.arange(m)
    if alpha >= 1.0:
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (m - 1))
    width = int(np.floor(alpha * (m - 1) / 2.0))
    n1, n3 = k[:width + 1], k[m - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (m - 1))))
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (m - 1))))
    return np.concatenate((w1, np.ones(m - 2 * width - 2), w3))


def lfilter_zi(b, a):
    """scipy.signal.lfilter_zi: the direct-form-II-transposed state of the step response's steady state."""
    b, a = np.atleast_1d(np.asarray(b, dtype=np.float64)), np.atleast_1d(np.asarray(a, dtype=np.float64))
    b, a = b / a[0], a / a[0]
    n = max(len(a), len(b))
    a, b = np.r_[a, np.zeros(n - len(a))], np.r_[b, np.zeros(n - len(b))]
    companion_t = np.zeros((n - 1, n - 1))
    companion_t[:, 0] = -a[1:]
    companion_t[np.arange(n - 2), np.arange(1, n - 1)] = 1.0
    return np.linalg.solve(np.eye(n - 1) - companion_t, b[1:] - a[1:] * b[0])


def lfilter(b, a, x, zi, dtype=np.float64):
    """scipy.signal.lfilter(b, a, x, zi=zi)[0] for a[0] == 1 (scipy/signal/_lfilter.c.in)."""
    n = max(len(a), len(b))
    b = np.r_[np.asarray(b, dtype=dtype), np.zeros(n - len(b), dtype=dtype)]
    a = np.r_[np.asarray(a, dtype=dtype), np.zeros(n - len(a), dtype=dtype)]
    z = np.asarray(zi, dtype=dtype).copy()
    y = np.empty(len(x), dtype=dtype)
    for i, xn in enumerate(np.asarray(x, dtype=dtype)):
        yn = z[0] + b[0] * xn
        for k in range(n - 2):
            z[k] = z[k + 1] + xn * b[k + 1] - yn * a[k + 1]
        z[n - 2] = xn * b[n - 1] - yn * a[n - 1]
        y[i] = yn
    return y


def sosfilt_zi(sos):
    """scipy.signal.sosfilt_zi."""
    sos = np.asarray(sos, dtype=np.float64)
    zi = np.empty((sos.shape[0], 2))
    scale = 1.0
    for s in range(sos.shape[0]):
        b, a = sos[s, :3], sos[s, 3:]
        zi[s] = scale * lfilter_zi(b, a)
        scale *= b.sum() / a.sum()
    return zi


def sosfilt(sos, x, zi, dtype=np.float64):
    """scipy.signal.sosfilt(sos, x, zi=zi)[0] (scipy/signal/_sosfilt.pyx)."""
    sos = np.asarray(sos, dtype=dtype)
    z = np.asarray(zi, dtype=dtype).copy()
    y = np.empty(len(x), dtype=dtype)
    for i, x_cur in enumerate(np.asarray(x, dtype=dtype)):
        for s in range(sos.shape[0]):
            x_new = sos[s, 0] * x_cur + z[s, 0]
            z[s, 0] = sos[s, 1] * x_cur - sos[s, 4] * x_new + z[s, 1]
            z[s, 1] = sos[s, 2] * x_cur - sos[s, 5] * x_new
            x_cur = x_new
        y[i] = x_cur
    return y


def _odd_ext(x, n):
    return np.concatenate((2 * x[0] - x[n:0:-1], x, 2 * x[-1] - x[-2:-n - 2:-1]))


def filtfilt(b, a, x, dtype=np.float64):
    """scipy.signal.filtfilt(b, a, x): odd extension by 3 * ntaps, forward and backward lfilter from the steady state."""
    b, a = np.atleast_1d(np.asarray(b, dtype=np.float64)), np.atleast_1d(np.asarray(a, dtype=np.float64))
    b, a = b / a[0], a / a[0]
    edge = 3 * max(len(a), len(b))
    if len(x) <= edge:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {edge}.")
    ext = _odd_ext(np.asarray(x, dtype=dtype), edge)
    zi = lfilter_zi(b, a).astype(dtype)
    y = lfilter(b, a, ext, zi * ext[0], dtype)
    y = lfilter(b, a, y[::-1], zi * y[-1], dtype)
    return y[::-1][edge:-edge]


def sosfiltfilt(sos, x, dtype=np.float64):
    """scipy.signal.sosfiltfilt(sos, x)."""
    sos = np.asarray(sos, dtype=np.float64)
    ntaps = 2 * sos.shape[0] + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    edge = 3 * int(ntaps)
    if len(x) <= edge:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {edge}.")
    ext = _odd_ext(np.asarray(x, dtype=dtype), edge)
    zi = sosfilt_zi(sos).astype(dtype)
    y = sosfilt(sos, ext, zi * ext[0], dtype)
    y = sosfilt(sos, y[::-1], zi * y[-1], dtype)
    return y[::-1][edge:-edge]


def butter_filtfilt(b, a, sig, tukey_alpha, dtype=np.float64):
    """styx_fft.py:87-90: filtfilt(b, a, sig * tukey(len(sig), alpha))."""
    return filtfilt(b, a, np.asarray(sig, dtype=np.float64) * tukey(len(sig), tukey_alpha), dtype)


# ============================================================================ synthetic inputs
def well_tempered_tone(fs=800.0, f_center=60.0, duration_s=10.24, fft_s=0.64, use_fft_frequency=True):
    """synth/benchmark_signals.py:268-355 (deterministic branch): cos(2 pi f_c n) with f_c snapped to an FFT bin."""
    n = 2 ** int(np.log2(duration_s * fs))
    n_fft = 2 ** int(np.log2(fft_s * fs))
    f_pos = np.fft.rfftfreq(n_fft, d=1 / fs)
    f_fft = f_pos[np.argmin(np.abs(f_pos - f_center))]
    f_c = (f_fft if use_fft_frequency else f_center) / fs
    k = np.arange(n)
    return np.cos(2.0 * np.pi * f_c * k), k / fs, n_fft, fs, f_fft, fs / n_fft


def decimate(x, q, dtype=np.float64):
    """scipy.signal.decimate(x, q, zero_phase=True): sosfiltfilt with cheby1(8, 0.05, 0.8 / q), every q-th sample
    (the Chebyshev sections are scipy's design; the record arithmetic is restated)."""
    from scipy.signal import cheby1
    return sosfiltfilt(cheby1(8, 0.05, 0.8 / q, output="sos"), x, dtype)[::q]


def quantum_chirp(omega, order=12.0, gamma=0.0, gauss=True, oversample_scale=2):
    """synth/benchmark_signals.py:57-109."""
    if omega >= 0.8 * np.pi:
        omega = np.pi * 2 ** (-1 / order)
    chirp_scale = (3.0 / 4.0 * np.pi * order / omega) * np.sqrt(1 + gamma ** 2)
    support = 2 ** int(np.ceil(np.log2(2.0 * np.pi * chirp_scale)))
    time0 = np.arange(oversample_scale * support)
    time = time0 - time0[-1] / 2
    phase = omega * time + 0.5 * gamma * (time / chirp_scale) ** 2
    wf = np.exp(-0.5 * (time / chirp_scale) ** 2 + 1j * phase) if gauss else np.exp(1j * phase)
    return decimate(np.real(wf), oversample_scale) + 1j * decimate(np.imag(wf), oversample_scale), support
