"""
Integer helpers of ``quantum_inferno.utilities.calculations`` used by ``styx_fft.stft_from_sig``
(reference utilities/calculations.py:160-205).
"""
import numpy as np

ROUNDING_TYPES = ["floor", "ceil", "round", "floor_power_of_two", "ceil_power_of_two"]
OUTPUT_TYPES = ["points", "log2", "pow2"]


def round_value(value: float, rounding_type: str = "round") -> int:
    """Round by name; 'round' is round-half-to-even like numpy (reference utilities/calculations.py:160-184)."""
    if rounding_type not in ROUNDING_TYPES:
        raise ValueError(f"Invalid rounding type {rounding_type}, must be one of {ROUNDING_TYPES}")
    if rounding_type == "floor":
        return int(np.floor(value))
    if rounding_type == "ceil":
        return int(np.ceil(value))
    if rounding_type == "round":
        return int(np.round(value))
    if rounding_type == "ceil_power_of_two":
        return 2 ** int(np.ceil(np.log2(value)))
    return 2 ** int(np.floor(np.log2(value)))


def get_num_points(sample_rate_hz: float, duration_s: float, rounding_type: str, output_unit: str) -> int:
    """Points in ``duration_s`` as points / log2(points) / 2^points (reference utilities/calculations.py:187-205)."""
    if output_unit not in OUTPUT_TYPES:
        raise ValueError(f"Invalid output unit {output_unit}, must be one of {OUTPUT_TYPES}")
    span = sample_rate_hz * duration_s
    if output_unit == "points":
        return round_value(span, rounding_type)
    if output_unit == "log2":
        return round_value(np.log2(span), rounding_type)
    return round_value(2 ** span, rounding_type)
