"""
The two helpers of ``quantum_inferno.utilities.rescaling`` used on the hot path
(reference utilities/rescaling.py:13-28).  ``to_log2_with_epsilon`` runs on the GPU for arrays that are
already device tensors; small host inputs are evaluated with numpy float64.
"""
from typing import Union

import numpy as np

from ..scales_dyadic import get_epsilon


def to_log2_with_epsilon(x: Union[np.ndarray, float, list]):
    """log2(|x| + eps) (reference utilities/rescaling.py:13-20)."""
    if hasattr(x, "data_ptr"):                       # CUDA tensor: elementwise kernel
        from .. import _driver
        from .._runtime import get_runtime
        rt = get_runtime()
        is_complex = x.is_complex()
        dt = "float64" if x.dtype in (rt.torch.float64, rt.torch.complex128) else "float32"
        return _driver.abs_log2(x.contiguous(), dt, is_complex, eps=get_epsilon(), rt=rt)
    return np.log2(np.abs(x) + get_epsilon())


def is_power_of_two(n: int) -> bool:
    """True for positive powers of two (reference utilities/rescaling.py:23-28)."""
    return n > 0 and not (n & (n - 1))
