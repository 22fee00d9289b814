"""
Host-side band tables and constants of the constant-Q framework (float64 numpy; stays on the CPU).

Drop-in for ``quantum_inferno.scales_dyadic`` on the time-frequency hot path: same public names,
signatures and -- because band indexing must be bit-exact -- the same floating-point expression
order as the reference (quantum_inferno/scales_dyadic.py:105-393).  Nothing here touches the GPU.
"""
import sys
from typing import List, Tuple, Union

import numpy as np

EPSILON64 = np.finfo(np.float64).eps
EPSILON32 = np.finfo(np.float32).eps
EPSILON16 = np.finfo(np.float16).eps

M_OVER_N = 0.75 * np.pi          # cycles per band order, reference scales_dyadic.py:21


def get_epsilon() -> float:
    """Machine epsilon matched to the interpreter word size (reference scales_dyadic.py:28-37)."""
    if sys.maxsize > 2 ** 32:
        return EPSILON64
    return EPSILON32 if sys.maxsize > 2 ** 16 else EPSILON16


class Slice:
    """Named constants (reference scales_dyadic.py:40-81)."""
    ORD1, ORD3, ORD6, ORD12, ORD24, ORD48 = 1.0, 3.0, 6.0, 12.0, 24.0, 48.0
    G2 = 2.0
    G3 = 10.0 ** 0.3
    T_PLANCK = 5.4e-44
    T0S = 1e-42
    T1S = 1.0
    T100S = 100.0
    T1000S = 1000.0
    T1M = 60.0
    T1H = T1M * 60.0
    T1D = T1H * 24.0
    TU = 2.0 ** 58
    F1HZ = 1.0
    F1KHZ = 1_000.0
    F0HZ = 1.0e42
    FU = 2.0 ** -58
    FS1HZ, FS10HZ, FS30HZ, FS80HZ, FS200HZ = 1.0, 10.0, 30.0, 80.0, 200.0
    FS400HZ, FS800HZ, FS8KHZ, FS16KHZ, FS48KHZ = 400.0, 800.0, 8_000.0, 16_000.0, 48_000.0


DEFAULT_SCALE_BASE = Slice.G3
DEFAULT_SCALE_ORDER = Slice.ORD3
DEFAULT_REF_FREQUENCY_HZ = Slice.F1HZ
DEFAULT_SCALE_ORDER_MIN: float = 0.75
_POW2_LIMIT = {EPSILON64: 63, EPSILON32: 31}.get(get_epsilon(), 15)
DEFAULT_FFT_POW2_POINTS_MAX: int = 2 ** _POW2_LIMIT
DEFAULT_FFT_POW2_POINTS_MIN: int = 2 ** 8
DEFAULT_MESH_POW2_PIXELS: int = 2 ** 19
DEFAULT_TIME_DISPLAY_S: float = 60.0
VALID_SCALE_ORDERS: List[float] = [0.75, 1, 1.5, 3, 6, 12, 24, 48]


def scale_order_check(scale_order: float = DEFAULT_SCALE_ORDER, show_warning: bool = True) -> float:
    """|order| clamped from below at 0.75, with the reference's printed warning (scales_dyadic.py:105-122)."""
    order = np.abs(scale_order)
    if order >= DEFAULT_SCALE_ORDER_MIN:
        return order
    if show_warning:
        print(f"** Warning from scales_dyadic.scale_order_check:\n"
              f"N < {DEFAULT_SCALE_ORDER_MIN} specified, overriding using N = {DEFAULT_SCALE_ORDER_MIN}")
    return DEFAULT_SCALE_ORDER_MIN


def scale_multiplier(scale_order: float = DEFAULT_SCALE_ORDER) -> float:
    """M = 0.75*pi*N (scales_dyadic.py:125-130)."""
    return M_OVER_N * scale_order_check(scale_order)


def cycles_from_order(scale_order: float) -> float:
    """Number of cycles M of the order-N atom (scales_dyadic.py:133-141)."""
    return scale_multiplier(scale_order)


def order_from_cycles(cycles_per_scale: float) -> float:
    """Inverse of cycles_from_order with a one-cycle floor (scales_dyadic.py:144-155)."""
    cycles = 1.0 if np.abs(cycles_per_scale) < 1 else cycles_per_scale
    return scale_order_check(cycles / M_OVER_N)


def base_multiplier(scale_order: float = DEFAULT_SCALE_ORDER, scale_base: float = DEFAULT_SCALE_BASE) -> float:
    """N / log2(G) (scales_dyadic.py:158-164)."""
    return scale_order_check(scale_order) / np.log2(scale_base)


def scale_from_frequency_hz(scale_order: float, scale_frequency_center_hz: Union[np.ndarray, float],
                            frequency_sample_rate_hz: float
                            ) -> Tuple[Union[np.ndarray, float], Union[np.ndarray, float]]:
    """(atom scale s = M/omega, omega = 2*pi*f/fs) (scales_dyadic.py:167-180)."""
    omega = 2.0 * np.pi * scale_frequency_center_hz / frequency_sample_rate_hz
    return cycles_from_order(scale_order) / omega, omega


def band_intervals_periods(scale_order_input: float, scale_base_input: float, scale_ref_input: float,
                           scale_low_input: float, scale_high_input: float, show_warnings: bool = True
                           ) -> Tuple[float, float, np.ndarray, float, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Nth-octave band numbers, centres and edges in the period domain (scales_dyadic.py:241-352).
    Returns (order, base, band_number, ref, centre_algebraic, centre_geometric, start, end)."""
    ref, low, high, base, order = np.absolute(
        [scale_ref_input, scale_low_input, scale_high_input, scale_base_input, scale_order_input])

    if base != Slice.G3 and base != Slice.G2:
        if base < 1.0:
            if show_warnings:
                print("\nWARNING: Base must be greater than unity. Overriding to G = 2")
            base = Slice.G2
        elif show_warnings:
            print("\nWARNING: Base is not ISO3 or ANSI S1.11 compliant")
            print(f"Continuing With Non-standard base = {base}...")
    if order not in VALID_SCALE_ORDERS:
        if order < 0.75:
            if show_warnings:
                print("Order must be greater than 0.75. Overriding to Order 1")
            order = 1
        elif show_warnings:
            print(f"\nWARNING: Recommend Orders {VALID_SCALE_ORDERS}")
            print(f"Continuing With Non-standard Order = {order}...")

    half_band = base ** (1.0 / (2.0 * order))
    rel_width = half_band - 1.0 / half_band

    if low < Slice.T0S:
        low = Slice.T0S / half_band
    if high < low:
        if show_warnings:
            print("\nWARNING: Upper scale must be larger than the lowest scale")
            print("Overriding to min = max/G\n")
        low = high / base
    if high == low:
        if show_warnings:
            print("\nWARNING: Upper scale = lowest scale, returning closest band edges")
        high *= half_band
        low /= half_band

    top = np.round(order * np.log(high / ref) / np.log(base))
    bottom = np.floor(order * np.log(low / ref) / np.log(base))
    lowest_centre = ref * np.power(base, bottom / order)
    if (lowest_centre < low) or (lowest_centre / half_band < low - get_epsilon()):
        bottom += 1          # keep the first band above the Nyquist period
    if top < bottom:
        if show_warnings:
            print("\nSPECMOD: Insufficient bandwidth for Nth band specification")
            print(f"Minimum scaled bandwidth (scale_high - scale_low)/scale_center = {rel_width}")
            print("Correct scale High/Low input parameters")
            print("Apply one order")
        top = np.floor(np.log10(high) / np.log10(base))
        bottom = top - order

    band_number = np.arange(bottom, top + 1)
    centre = ref * np.power(base * np.ones(band_number.shape), band_number / order)
    start, end = centre / half_band, centre * half_band
    return order, base, band_number, ref, (start + end) / 2.0, centre, start, end


def band_frequency_low_high(frequency_order_input: float, frequency_base_input: float, frequency_ref_input: float,
                            frequency_low_input: float, frequency_high_input: float,
                            frequency_sample_rate_input: float
                            ) -> Tuple[float, float, np.ndarray, float, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Frequency-domain view of band_intervals_periods (scales_dyadic.py:183-238); centres descend."""
    period_low = 1 / frequency_high_input
    nyquist_period = 2 / frequency_sample_rate_input
    if period_low < nyquist_period:
        period_low = nyquist_period
    order, base, band, ref, _, centre, start, end = band_intervals_periods(
        frequency_order_input, frequency_base_input, 1 / frequency_ref_input, period_low, 1 / frequency_low_input)
    f_hi, f_lo = 1 / start, 1 / end
    return order, base, -band, 1 / ref, (f_hi + f_lo) / 2.0, 1 / centre, f_lo, f_hi


def log_frequency_hz_from_fft_points(frequency_sample_hz: float, fft_points: int,
                                     scale_order: float = DEFAULT_SCALE_BASE,
                                     scale_ref_hz: float = DEFAULT_REF_FREQUENCY_HZ,
                                     scale_base: float = DEFAULT_SCALE_BASE) -> np.ndarray:
    """Ascending band centres ref*G^(-j/N) between 0.8 Nyquist and the longest atom that fits
    2^ceil(log2(fft_points)) samples (scales_dyadic.py:355-393).  The odd default of ``scale_order``
    (= the base G3) is the reference's and is kept."""
    log2_window = int(np.ceil(np.log2(fft_points)))
    cycles = scale_multiplier(scale_order)
    bands_per_log2 = base_multiplier(scale_order, scale_base)
    log2_cycles = np.log2(cycles)
    log2_rate = np.log2(frequency_sample_hz / scale_ref_hz)
    j_first = int(np.ceil(bands_per_log2 * (np.log2(2.5) - log2_rate)))
    j_last = int(np.floor(bands_per_log2 * (log2_window - log2_cycles - log2_rate)))
    j = np.arange(j_first, j_last + 1)
    return np.flip(scale_ref_hz * scale_base ** (-j / scale_order))
