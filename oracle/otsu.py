"""Restatement of ``skimage.filters.threshold_otsu`` (scikit-image==0.24.0) on top of
``np.histogram`` (numpy==1.26.4), as called at
``/root/reference/code/aind_smartspim_destripe/filtering.py:190-192`` on ``ch**2``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Neither library is vendored in
``/root/reference`` (pins: ``environment/Dockerfile:16,18``); the published algorithm
is restated:

1. all pixels equal the first pixel -> return that value;
2. 256 equal-width bins over ``[min, max]`` (last bin right-closed); a value's bin is
   fixed by comparisons against the *float32* edge array (np.histogram's fast path
   corrects its arithmetic guess by +-1 against ``bin_edges``), the edges being
   ``np.linspace(min, max, 257)``.  numpy 1.26.4 evaluates that linspace in float64
   from the float32 ``step = (max - min) / 256`` and rounds to float32 (numpy >= 2
   evaluates it in float32); the pinned 1.26.4 behaviour is restated with explicit
   arithmetic so that the result does not depend on the numpy installed here;
3. counts -> float32; cumulative sums and class means in float32 (sequential
   ``cumsum``), between-class variance, first arg-max, return that bin's centre.
"""

from __future__ import annotations

import numpy as np

NBINS = 256


def histogram_edges_f32(vmin: np.float32, vmax: np.float32, nbins: int = NBINS) -> np.ndarray:
    """``np.histogram_bin_edges`` for a float32 array under numpy 1.26.4 semantics."""
    first = np.float32(vmin)
    last = np.float32(vmax)
    if first == last:
        first = np.float32(first - np.float32(0.5))
        last = np.float32(last + np.float32(0.5))
    delta = np.float32(last - first)  # float32 - float32
    step = np.float32(delta / np.float32(nbins))  # float32 / python int -> float32
    y = np.arange(0, nbins + 1, dtype=np.float64)
    if step == 0:
        y = y / nbins
        y = y * np.float64(delta)
    else:
        y = y * np.float64(step)
    y = y + np.float64(first)
    y[-1] = np.float64(last)
    return y.astype(np.float32)


def histogram_f32(values: np.ndarray, nbins: int = NBINS):
    """(counts int64, edges float32) with np.histogram's uniform-bin semantics."""
    a = np.asarray(values).reshape(-1)
    assert a.dtype == np.float32, "oracle histogram restates the float32 path only"
    first = a.min()
    last = a.max()
    edges = histogram_edges_f32(first, last, nbins)
    first_edge = edges[0]
    last_edge = edges[-1]
    norm_denom = np.float32(last_edge - first_edge)
    keep = (a >= first_edge) & (a <= last_edge)
    t = a[keep]
    f_idx = ((t - first_edge) / norm_denom) * nbins  # float32 arithmetic, like numpy
    idx = f_idx.astype(np.intp)
    idx[idx == nbins] -= 1
    dec = t < edges[idx]
    idx[dec] -= 1
    inc = (t >= edges[idx + 1]) & (idx != nbins - 1)
    idx[inc] += 1
    counts = np.bincount(idx, minlength=nbins).astype(np.int64)
    return counts, edges


def otsu_from_histogram(counts: np.ndarray, edges: np.ndarray):
    """skimage threshold_otsu tail: float32 counts, float32 cumulative sums."""
    counts = counts.astype(np.float32)
    centers = (edges[:-1] + edges[1:]) / 2.0  # float32 stays float32
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * centers) / weight1
        mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
        variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = int(np.argmax(variance12))
    return centers[idx], idx


def threshold_otsu(image: np.ndarray, nbins: int = NBINS):
    """``skimage.filters.threshold_otsu(image)`` for a float image."""
    image = np.asarray(image)
    first_pixel = image.reshape(-1)[0]
    if np.all(image == first_pixel):
        return first_pixel
    if image.dtype == np.float32:
        counts, edges = histogram_f32(image, nbins)
    else:
        # float64 (TIFF path: uint16 input -> float64 log): numpy computes the edges in
        # float64 in every version, so np.histogram itself is the specification.
        counts, edges = np.histogram(image.reshape(-1), bins=nbins)
    thr, _ = otsu_from_histogram(counts, edges)
    return thr
