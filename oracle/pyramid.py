"""Restatement of the multiscale step the reference applies to the destriped volume.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``/root/reference/code/aind_smartspim_destripe/zarr_destriper.py:365-407`` calls
``xarray_multiscale.multiscale(array, windowed_mean, scale_factors, preserve_dtype=True)``
(xarray-multiscale==2.1.0, pinned in ``environment/Dockerfile:29``, not vendored) with scale
``(1, 1, 2, 2, 2)`` and keeps levels ``[:n_lvls]``; ``compute_multiscale`` (:741-749) builds every
level from the previously written one.  Published semantics: the array is cropped to a multiple
of the window, ``windowed_mean`` is ``reshape -> mean`` in float64, ``preserve_dtype`` casts the
result back (truncation for unsigned integers).
"""

from __future__ import annotations

from typing import List, Sequence

import numpy as np


def windowed_mean(array: np.ndarray, window: Sequence[int]) -> np.ndarray:
    array = np.asarray(array)
    crop = tuple(slice(0, (s // w) * w) for s, w in zip(array.shape, window))
    a = array[crop]
    new_shape = []
    for s, w in zip(a.shape, window):
        new_shape += [s // w, w]
    r = a.reshape(new_shape).mean(axis=tuple(range(1, 2 * a.ndim, 2)))
    return r.astype(array.dtype)


def compute_pyramid(data: np.ndarray, n_lvls: int, scale_axis: Sequence[int]) -> List[np.ndarray]:
    """Levels 0 .. n_lvls-1; level k is the windowed mean of level k-1."""
    levels = [np.asarray(data)]
    for _ in range(1, n_lvls):
        levels.append(windowed_mean(levels[-1], scale_axis))
    return levels
