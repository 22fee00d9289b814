"""
Row / column broadcast products of ``quantum_inferno.utilities.matrix`` (reference utilities/matrix.py:89-134).
The reference materialises the tiled operand with np.tile; the product is the same with broadcasting.
"""
import numpy as np


def d0tile_x_d0d1(d0, d0d1: np.ndarray) -> np.ndarray:
    """Multiply every row f of ``d0d1`` by ``d0[f]`` (reference utilities/matrix.py:89-110)."""
    d0d1 = np.asarray(d0d1)
    if d0d1.ndim == 1:
        tiled = np.tile(d0, d0d1.shape[0])
        if tiled.shape == d0d1.shape:
            return tiled * d0d1
    elif d0d1.ndim == 2 and np.shape(d0) == (d0d1.shape[0],):
        return np.asarray(d0)[:, None] * d0d1
    raise TypeError(f"Cannot handle an array of shape {np.shape(d0)}.")


def d1tile_x_d0d1(d1, d0d1: np.ndarray) -> np.ndarray:
    """Multiply every column t of ``d0d1`` by ``d1[t]`` (reference utilities/matrix.py:113-134)."""
    d0d1 = np.asarray(d0d1)
    if d0d1.ndim == 1:
        tiled = np.tile(d1, d0d1.shape[0])
        if tiled.shape == d0d1.shape:
            return tiled * d0d1
    elif d0d1.ndim == 2 and np.shape(d1) == (d0d1.shape[1],):
        return np.asarray(d1)[None, :] * d0d1
    raise TypeError(f"Cannot handle an array of shape {np.shape(d1)}.")
