"""Dtype-faithful CPU restatement of the reference plane filter.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/code/aind_smartspim_destripe/filtering.py`` function by
function (line ranges cited per function).  ``pywt`` and ``skimage`` are replaced by
``oracle.dwt`` / ``oracle.otsu``; ``scipy.fftpack`` and ``np.median`` are the real ones.
If the real ``pywt`` and ``skimage`` ever become importable, ``USE_REAL_THIRD_PARTY``
switches the two third-party calls to them so the restatement can be cross-checked.
"""

from __future__ import annotations

import importlib.util
from typing import List, Optional, Tuple

import numpy as np
from scipy import fftpack

from . import dwt as _dwt
from . import otsu as _otsu

USE_REAL_THIRD_PARTY = (
    importlib.util.find_spec("pywt") is not None
    and importlib.util.find_spec("skimage") is not None
)

if USE_REAL_THIRD_PARTY:  # pragma: no cover - not available in the build container
    import pywt as _pywt
    from skimage import filters as _skfilters

    def _wavedec2(x, wavelet, level):
        return _pywt.wavedec2(x, wavelet=wavelet, level=level)

    def _waverec2(coeffs, wavelet):
        return _pywt.waverec2(coeffs, wavelet)

    def _threshold_otsu(x):
        return _skfilters.threshold_otsu(x)

else:

    def _wavedec2(x, wavelet, level):
        return _dwt.wavedec2(x, wavelet=wavelet, level=level)

    def _waverec2(coeffs, wavelet):
        return _dwt.waverec2(coeffs, wavelet)

    def _threshold_otsu(x):
        return _otsu.threshold_otsu(x)


def sigmoid(data):
    """filtering.py:13-22"""
    return 1 / (1 + np.exp(-data))


def foreground_fraction(img, center, crossover):
    """filtering.py:25-51"""
    z = (img - center) / crossover
    return sigmoid(z)


def get_foreground_background_mean(img, threshold_mask=0.3) -> Tuple:
    """filtering.py:54-88 (float16 sigmoid mask; means of the two pixel classes)."""
    with np.errstate(over="ignore"):
        cell_for = foreground_fraction(img.astype(np.float16), 400, 20)
    cell_for[cell_for > threshold_mask] = 1
    cell_for[cell_for <= threshold_mask] = 0
    foreground = img[cell_for == 1]
    background = img[cell_for == 0]
    foreground_mean = foreground.mean() if foreground.size else 0.0
    background_mean = background.mean() if background.size else 0.0
    return foreground_mean, background_mean, cell_for


def notch(n, sigma):
    """filtering.py:91-115"""
    if n <= 0:
        raise ValueError("n must be positive")
    else:
        n = int(n)
    if sigma <= 0:
        raise ValueError("sigma must be positive")
    x = np.arange(n)
    return 1 - np.exp(-(x**2) / (2 * sigma**2))


def gaussian_filter(shape, sigma):
    """filtering.py:118-136"""
    g = notch(n=shape[-1], sigma=sigma)
    return np.broadcast_to(g, shape).copy()


def _row_notch(band_rows: np.ndarray, s: float) -> np.ndarray:
    """filtering.py:206-215: packed rfft along the last axis, gaussian_filter mask, irfft."""
    spectrum = fftpack.rfft(band_rows, axis=-1)
    damp = gaussian_filter(shape=spectrum.shape, sigma=s)
    return fftpack.irfft(spectrum * damp)


def _filter_detail_level(ch, width_fraction, max_threshold, record=None):
    """One iteration of the level loop, filtering.py:186-219 (only cH is modified)."""
    energy = ch**2  # :187
    magnitude = np.sqrt(energy)  # :188
    otsu_raw = _threshold_otsu(energy)  # :190-192
    threshold = min(max_threshold, np.sqrt(otsu_raw))  # :193
    is_fg = magnitude > threshold  # :195
    fg_part = ch * is_fg  # :196
    bg_part = ch * (1 - is_fg)  # :197  (1 - bool) is int64 -> float64 from here on
    row_median = np.median(bg_part, axis=-1)  # :199-202
    inpainted = bg_part + np.broadcast_to(row_median[..., np.newaxis], ch.shape) * is_fg  # :204
    n_rows = inpainted.shape[1] if inpainted.ndim == 3 else inpainted.shape[0]  # :208-211
    s = n_rows * width_fraction  # :213
    bg_filtered = _row_notch(inpainted, s)  # :206, :214-215
    result = fg_part + bg_filtered * (1 - is_fg)  # :217
    if record is not None:
        record.append(
            dict(ch=ch, otsu=otsu_raw, threshold=threshold, mask=is_fg, median=row_median, s=s,
                 background_filtered=bg_filtered, ch_filtered=result)
        )
    return result


def log_space_fft_filtering(
    input_image,
    wavelet="db3",
    level=0,
    sigma=64,
    max_threshold=4,
    _trace: Optional[dict] = None,
):
    """filtering.py:139-224.  ``_trace`` (oracle-only) collects per-level intermediates."""
    log_image = np.log(1.0 + input_image)  # :175
    pyramid = _wavedec2(log_image, wavelet, level)  # :176
    spatial = input_image.shape[1:] if input_image.ndim == 3 else input_image.shape  # :180-183
    width_fraction = sigma / min(spatial)
    levels_rec = [] if _trace is not None else None
    rebuilt = [pyramid[0]]
    for ch, cv, cd in pyramid[1:]:  # :186
        rebuilt.append((_filter_detail_level(ch, width_fraction, max_threshold, levels_rec), cv, cd))
    log_filtered = _waverec2(rebuilt, wavelet)  # :221
    if _trace is not None:
        _trace.update(log=log_image, approx=pyramid[0], levels=levels_rec, log_filtered=log_filtered)
    return np.exp(log_filtered) + 1.0  # :222  (plus one, as in the reference)


def normalize_image(images: List[np.ndarray]) -> np.ndarray:
    """filtering.py:227-250"""
    images = np.array(images)
    min_val = np.min(images)
    max_val = np.max(images)
    return 1 + np.divide(images - min_val, max_val - min_val).astype(np.float16)


def invert_image(image) -> np.ndarray:
    """filtering.py:253-270"""
    image = np.array(image)
    return image.max() - image


def get_hemisphere_flatfield(input_tile_path, tile_config, flatfields, zarr=True):
    """filtering.py:273-335"""
    if zarr:
        xy = str(input_tile_path).split("_")
    else:
        xy = str(input_tile_path).split("/")[-2].split("_")
    x_folder, y_folder = xy[0], xy[1]
    if tile_config.get(x_folder) is None:
        raise KeyError(f"Please, check the tile config while trying to reach: {x_folder}")
    brain_side = tile_config[x_folder].get(y_folder)
    if brain_side is None:
        raise KeyError(f"Please, check the tile config while trying to reach: {y_folder}")
    return flatfields[brain_side]


def flatfield_correction(image_tiles, flatfield, darkfield, baseline=None) -> np.ndarray:
    """filtering.py:338-414 (dark subtract with floor at 0, divide by flat, clip, TRUNCATE)."""
    tiles = np.array(image_tiles)  # :369 (copy)
    flat = flatfield if tiles.ndim == flatfield.ndim else np.expand_dims(flatfield, axis=0)  # :371-372
    dark = darkfield if tiles.ndim == darkfield.ndim else np.expand_dims(darkfield, axis=0)  # :374-375
    dark = dark[: tiles.shape[-2], : tiles.shape[-1]]  # :377
    for name, field in (("darkfield. ", dark), ("flatfield.", flat)):  # :379-391
        if field.shape != tiles.shape:
            raise ValueError(
                f"Please, check the shape of the {name}Image: {tiles.shape} - {name.strip('. ').capitalize()}: {field.shape}"
            )
    base = np.zeros((tiles.shape[0],)) if baseline is None else baseline  # :393-394
    base = base[tuple([slice(None)] + [np.newaxis] * (tiles.ndim - 1))]  # :396
    below = tiles <= dark  # :399-406: pixels at or under the dark level become 0, the rest lose it
    tiles[below] = 0
    tiles[~below] = tiles[~below] - dark[~below]
    corrected = tiles / flat - base  # :409
    return np.clip(corrected, 0, 65535).astype("uint16")  # :412 (truncation)


def filter_stripes(
    image,
    input_tile_path,
    no_cells_config,
    cells_config,
    shadow_correction=None,
    microscope_high_int=2700,
):
    """filtering.py:417-491"""
    fore_mean, back_mean, _ = get_foreground_background_mean(image)
    if fore_mean > back_mean and fore_mean > microscope_high_int:
        filtered_image = log_space_fft_filtering(input_image=image, **cells_config)
    else:
        filtered_image = log_space_fft_filtering(input_image=image, **no_cells_config)

    if shadow_correction is not None:
        retrospective = shadow_correction.get("retrospective")
        flatfield = shadow_correction.get("flatfield")
        darkfield = shadow_correction.get("darkfield")
        tile_config = shadow_correction.get("tile_config")
        if not retrospective:
            flatfield = get_hemisphere_flatfield(
                input_tile_path=input_tile_path,
                tile_config=tile_config,
                flatfields=flatfield,
            )
        filtered_image = flatfield_correction(
            image_tiles=filtered_image,
            flatfield=flatfield,
            darkfield=darkfield,
            baseline=None,
        )
    return filtered_image
