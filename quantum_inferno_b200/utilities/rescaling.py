"""
The two helpers of ``quantum_inferno.utilities.rescaling`` used on the hot path
(reference utilities/rescaling.py:13-28).  ``to_log2_with_epsilon`` runs on the GPU for every array input.
"""
from typing import Union

import numpy as np

from ..scales_dyadic import get_epsilon


def to_log2_with_epsilon(x: Union[np.ndarray, float, list]):
    """log2(|x| + eps) (reference utilities/rescaling.py:13-20).  Arrays (numpy or CUDA tensors, real or complex) go
    through the elementwise kernel ``qi_abs_log2``; numpy in -> numpy out.  Python scalars and lists -- the scalar
    helper use of the reference -- are evaluated in place."""
    is_tensor = hasattr(x, "data_ptr")
    if is_tensor or (isinstance(x, np.ndarray) and x.ndim >= 1 and x.size > 0):
        from .. import _driver
        from .._runtime import finish, get_runtime
        rt = get_runtime(x)
        if is_tensor:
            is_complex = x.is_complex()
            dt = "float64" if x.dtype in (rt.torch.float64, rt.torch.complex128) else "float32"
            buf = x.contiguous()
        else:
            is_complex = np.iscomplexobj(x)
            dt = "float32" if x.dtype in (np.float32, np.complex64) else "float64"
            buf = rt.asarray(x, {"float32": "complex64", "float64": "complex128"}[dt] if is_complex else dt)
        return finish(rt, _driver.abs_log2(buf, dt, is_complex, eps=get_epsilon(), rt=rt), not is_tensor)
    return np.log2(np.abs(x) + get_epsilon())


def is_power_of_two(n: int) -> bool:
    """True for positive powers of two (reference utilities/rescaling.py:23-28)."""
    return n > 0 and not (n & (n - 1))
