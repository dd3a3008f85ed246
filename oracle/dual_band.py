"""CPU restatement of pystripe's classic dual-band filter (``pystripe.core.filter_streaks`` /
``filter_subband``; the "Dual-band" picture of the reference README, ``/root/reference/README.md:7-8``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference snapshot does NOT contain this
mode: only its helpers survive (``sigmoid`` ``filtering.py:13-22``,
``foreground_fraction`` ``:25-51``).  It is restated from the published pystripe algorithm with the
reference's helpers; nothing under ``/root/reference`` pins its results -> **parity unpinned**.

    threshold = given, or skimage.filters.threshold_otsu(img) (integer image: one bin per value)
    background = clip(img, None, threshold);  foreground = clip(img, threshold, None)
    subband(x, sigma): log(1 + x) -> wavedec2(db3, level) -> per level
                       cH <- irfft(rfft(cH) * gaussian_filter(shape, s)),  s = cH.shape[0] * sigma / img.shape[0]
                       -> waverec2 -> exp(y) - 1
    f = foreground_fraction(img, threshold, crossover)      (no smoothing: the reference helper has none)
    out = foreground_filtered * f + background_filtered * (1 - f)
    dark > 0: out -= dark;  flat: out /= flat;  clip to [0, 65535];  astype(uint16)
"""
from __future__ import annotations

import numpy as np

from . import dwt as _dwt
from . import plane_filter as _pf


def threshold_otsu_integer(img: np.ndarray) -> float:
    """``skimage.filters.threshold_otsu`` (0.24) on an integer image: ``histogram`` takes the
    bincount path (one bin per integer value between min and max), counts as float32."""
    img = np.asarray(img)
    first = img.reshape(-1)[0]
    if np.all(img == first):
        return float(first)
    lo, hi = int(img.min()), int(img.max())
    counts = np.bincount(img.reshape(-1).astype(np.int64) - lo, minlength=hi - lo + 1).astype(np.float32)
    centers = np.arange(lo, hi + 1)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * centers) / weight1
        mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
        variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return float(centers[int(np.argmax(variance12))])


def filter_subband(img: np.ndarray, sigma: float, level, wavelet: str = "db3") -> np.ndarray:
    """pystripe ``filter_subband``: pure notch on every horizontal detail band (float64 flow)."""
    log_image = np.log(1.0 + img)
    pyramid = _dwt.wavedec2(log_image, wavelet, None if not level else level)  # pystripe: level 0 = maximum
    width_fraction = sigma / img.shape[0]
    rebuilt = [pyramid[0]]
    for ch, cv, cd in pyramid[1:]:
        s = ch.shape[0] * width_fraction
        rebuilt.append((_pf._row_notch(ch, s), cv, cd))
    out = _dwt.waverec2(rebuilt, wavelet)
    return np.exp(out[: img.shape[0], : img.shape[1]]) - 1.0


def filter_streaks(img, sigma, level=0, wavelet="db3", crossover=10, threshold=-1, flat=None, dark=0):
    """pystripe ``filter_streaks(img, sigma=[foreground, background], ...)`` -> uint16 plane."""
    if threshold == -1:
        try:
            threshold = threshold_otsu_integer(img)
        except ValueError:
            threshold = 1
    img = np.array(img, dtype=float)
    sigma_fg, sigma_bg = float(sigma[0]), float(sigma[1])
    if sigma_fg > 0:
        if sigma_bg > 0:
            if sigma_fg == sigma_bg:
                out = filter_subband(img, sigma_fg, level, wavelet)
            else:
                background = np.clip(img, None, threshold)
                foreground = np.clip(img, threshold, None)
                bgf = filter_subband(background, sigma_bg, level, wavelet)
                fgf = filter_subband(foreground, sigma_fg, level, wavelet)
                f = _pf.foreground_fraction(img, threshold, crossover)
                out = fgf * f + bgf * (1 - f)
        else:
            foreground = np.clip(img, threshold, None)
            fgf = filter_subband(foreground, sigma_fg, level, wavelet)
            f = _pf.foreground_fraction(img, threshold, crossover)
            out = fgf * f + img * (1 - f)
    else:
        if sigma_bg > 0:
            background = np.clip(img, None, threshold)
            bgf = filter_subband(background, sigma_bg, level, wavelet)
            f = _pf.foreground_fraction(img, threshold, crossover)
            out = img * f + bgf * (1 - f)
        else:
            out = img
    if dark > 0:
        out = out - dark
    if flat is not None:
        out = out / flat
    out = np.clip(out, 0, 2**16 - 1)
    return out.astype(np.uint16)
