"""Restatement of the PyWavelets 1.6.0 routines the reference calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference call sites: ``/root/reference/code/aind_smartspim_destripe/filtering.py:176``
(``pywt.wavedec2(input_image_log, wavelet=wavelet, level=level)``) and ``:221``
(``pywt.waverec2(coeff_filtered, wavelet)``).  PyWavelets==1.6.0 is pinned in
``/root/reference/environment/Dockerfile:20`` and is NOT vendored, so the published
algorithm is restated here:

* mode ``symmetric`` (half-sample symmetric extension, repeated for short signals),
* 1-D analysis  ``c[o] = sum_j filt[j] * x_ext[2*o + 1 - j]``, ``o < (N + F - 1) // 2``,
  accumulated in the input's float type in tap order j = 0..F-1,
* 2-D step transforms axis -2 first, then axis -1; detail tuple is
  ``(cH, cV, cD) = ('da', 'ad', 'dd')`` (first letter = axis -2),
* synthesis transforms axis -1 first, then axis -2, output length ``2 n - F + 2``,
* ``waverec2`` drops the last row/column of the running approximation when it is one
  sample longer than the next level's details.

Pinned by: PyWavelets' documented known answers (``tests/test_oracle_dwt.py``) and
perfect reconstruction.
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

_DEC_LO = {
    "db1": [0.7071067811865476, 0.7071067811865476],
    "haar": [0.7071067811865476, 0.7071067811865476],
    "db2": [
        -0.12940952255092145,
        0.22414386804185735,
        0.836516303737469,
        0.48296291314469025,
    ],
    "db3": [
        0.035226291882100656,
        -0.08544127388224149,
        -0.13501102001039084,
        0.4598775021193313,
        0.8068915093133388,
        0.3326705529509569,
    ],
}


def filter_bank(wavelet: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(dec_lo, dec_hi, rec_lo, rec_hi) as float64, pywt orthogonal-wavelet convention."""
    if wavelet not in _DEC_LO:
        raise ValueError(f"Unknown wavelet name '{wavelet}' (oracle knows {sorted(_DEC_LO)})")
    dec_lo = np.asarray(_DEC_LO[wavelet], dtype=np.float64)
    F = dec_lo.size
    rec_lo = dec_lo[::-1].copy()
    # dec_hi[k] = (-1)^(k+1) * dec_lo[F-1-k]
    dec_hi = np.array([(-1.0) ** (k + 1) * dec_lo[F - 1 - k] for k in range(F)])
    rec_hi = dec_hi[::-1].copy()
    return dec_lo, dec_hi, rec_lo, rec_hi


def dwt_coeff_len(n: int, F: int) -> int:
    """pywt.dwt_coeff_len for every mode except periodization."""
    return (n + F - 1) // 2


def dwt_max_level(n: int, F: int) -> int:
    """pywt.dwt_max_level: floor(log2(n / (F - 1))), clipped at 0."""
    if n < F - 1 or F < 2:
        return 0
    return max(int(math.floor(math.log2(n / (F - 1.0)))), 0)


def dwtn_max_level(shape: Sequence[int], wavelet: str) -> int:
    F = len(_DEC_LO[wavelet])
    return min(dwt_max_level(int(n), F) for n in shape)


def _float_type(x: np.ndarray) -> np.dtype:
    # pywt computes in float32 for float32 input (and float16 -> float32), else float64
    if x.dtype == np.float32 or x.dtype == np.float16:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def dwt_axis(x: np.ndarray, wavelet: str, axis: int) -> Tuple[np.ndarray, np.ndarray]:
    """Single-level 1-D analysis along ``axis`` (mode symmetric)."""
    dec_lo, dec_hi, _, _ = filter_bank(wavelet)
    dt = _float_type(x)
    x = np.moveaxis(np.asarray(x, dtype=dt), axis, -1)
    N = x.shape[-1]
    F = dec_lo.size
    n_out = dwt_coeff_len(N, F)
    pad = [(0, 0)] * (x.ndim - 1) + [(F - 1, F - 1)]
    ext = np.pad(x, pad, mode="symmetric")  # x_ext[i] lives at ext[..., i + F - 1]
    lo = dec_lo.astype(dt)
    hi = dec_hi.astype(dt)
    cA = np.zeros(x.shape[:-1] + (n_out,), dtype=dt)
    cD = np.zeros_like(cA)
    for j in range(F):
        start = 1 - j + (F - 1)
        sl = ext[..., start : start + 2 * n_out : 2]
        cA += sl * lo[j]
        cD += sl * hi[j]
    return np.moveaxis(cA, -1, axis), np.moveaxis(cD, -1, axis)


def idwt_axis(cA: np.ndarray, cD: np.ndarray, wavelet: str, axis: int) -> np.ndarray:
    """Single-level 1-D synthesis along ``axis``; output length 2 n - F + 2."""
    _, _, rec_lo, rec_hi = filter_bank(wavelet)
    dt = np.result_type(_float_type(np.asarray(cA)), _float_type(np.asarray(cD)))
    a = np.moveaxis(np.asarray(cA, dtype=dt), axis, -1)
    d = np.moveaxis(np.asarray(cD, dtype=dt), axis, -1)
    if a.shape != d.shape:
        raise ValueError("Coefficients arrays must have the same size.")
    n = a.shape[-1]
    F = rec_lo.size
    half = F // 2
    m_count = n - half + 1  # number of (even, odd) output pairs
    out = np.zeros(a.shape[:-1] + (2 * m_count,), dtype=dt)
    lo = rec_lo.astype(dt)
    hi = rec_hi.astype(dt)
    even = np.zeros(a.shape[:-1] + (m_count,), dtype=dt)
    odd = np.zeros_like(even)
    for j in range(half):
        sa = a[..., half - 1 - j : half - 1 - j + m_count]
        sd = d[..., half - 1 - j : half - 1 - j + m_count]
        even += lo[2 * j] * sa
        even += hi[2 * j] * sd
        odd += lo[2 * j + 1] * sa
        odd += hi[2 * j + 1] * sd
    out[..., 0::2] = even
    out[..., 1::2] = odd
    return np.moveaxis(out, -1, axis)


def dwt2(x: np.ndarray, wavelet: str):
    """pywt.dwt2 with axes=(-2, -1): returns cA, (cH, cV, cD)."""
    a, d = dwt_axis(x, wavelet, -2)  # first letter: axis -2
    aa, ad = dwt_axis(a, wavelet, -1)
    da, dd = dwt_axis(d, wavelet, -1)
    return aa, (da, ad, dd)


def idwt2(cA, details, wavelet: str) -> np.ndarray:
    cH, cV, cD = details
    lo = idwt_axis(cA, cV, wavelet, -1)  # 'aa','ad' -> 'a' along axis -2
    hi = idwt_axis(cH, cD, wavelet, -1)  # 'da','dd' -> 'd' along axis -2
    return idwt_axis(lo, hi, wavelet, -2)


def wavedec2(x: np.ndarray, wavelet: str = "db3", level=None) -> List:
    """pywt.wavedec2(x, wavelet, mode='symmetric', level=level, axes=(-2, -1))."""
    x = np.asarray(x)
    if x.ndim < 2:
        raise ValueError("Expected input data to have at least 2 dimensions.")
    if level is None:
        level = dwtn_max_level(x.shape[-2:], wavelet)
    elif level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    coeffs = []
    a = x
    for _ in range(int(level)):
        a, det = dwt2(a, wavelet)
        coeffs.append(det)
    coeffs.append(a)
    coeffs.reverse()
    return coeffs


def waverec2(coeffs: Sequence, wavelet: str = "db3") -> np.ndarray:
    """pywt.waverec2(coeffs, wavelet, mode='symmetric', axes=(-2, -1))."""
    if len(coeffs) < 1:
        raise ValueError("Coefficient list too short (minimum 1 array required).")
    a = np.asarray(coeffs[0])
    if len(coeffs) == 1:
        return a
    for det in coeffs[1:]:
        cH = np.asarray(det[0])
        # drop the extra trailing row / column left by an odd-sized level
        if a.shape[-2] == cH.shape[-2] + 1:
            a = a[..., :-1, :]
        if a.shape[-1] == cH.shape[-1] + 1:
            a = a[..., :-1]
        a = idwt2(a, det, wavelet)
    return a
