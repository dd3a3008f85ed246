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


def log_space_fft_filtering(
    input_image,
    wavelet="db3",
    level=0,
    sigma=64,
    max_threshold=4,
    _trace: Optional[dict] = None,
):
    """filtering.py:139-224.  ``_trace`` (oracle-only) collects per-level intermediates."""
    input_image_log = np.log(1.0 + input_image)
    coeffs = _wavedec2(input_image_log, wavelet, level)
    approx = coeffs[0]
    detail = coeffs[1:]

    width_fraction = sigma / min(input_image.shape)
    if len(input_image.shape) == 3:
        width_fraction = sigma / min(input_image.shape[1:])

    if _trace is not None:
        _trace["log"] = input_image_log
        _trace["approx"] = approx
        _trace["levels"] = []

    coeff_filtered = [approx]
    for i, (ch, cv, cd) in enumerate(detail):
        ch_sq = ch**2
        ch_power = np.sqrt(ch_sq)

        otsu_raw = _threshold_otsu(ch_sq)
        otsu_threshold_sqrt = np.sqrt(otsu_raw)
        threshold = min(max_threshold, otsu_threshold_sqrt)

        mask = ch_power > threshold
        foreground = ch * mask
        background = ch * (1 - mask)

        background_means = np.broadcast_to(
            np.median(background, axis=-1)[..., np.newaxis], ch.shape
        )
        background_inpainted = background + background_means * mask

        fft = fftpack.rfft(background_inpainted, axis=-1)
        s_shape = fft.shape[0]
        if len(fft.shape) == 3:
            s_shape = fft.shape[1]
        s = s_shape * width_fraction
        g = gaussian_filter(shape=fft.shape, sigma=s)
        background_filtered = fftpack.irfft(fft * g)

        ch_filtered = foreground + background_filtered * (1 - mask)
        coeff_filtered.append((ch_filtered, cv, cd))

        if _trace is not None:
            _trace["levels"].append(
                dict(
                    ch=ch,
                    otsu=otsu_raw,
                    threshold=threshold,
                    mask=mask,
                    median=np.median(background, axis=-1),
                    s=s,
                    background_filtered=background_filtered,
                    ch_filtered=ch_filtered,
                )
            )

    img_log_filtered = _waverec2(coeff_filtered, wavelet)
    img_filtered = np.exp(img_log_filtered) + 1.0
    if _trace is not None:
        _trace["log_filtered"] = img_log_filtered
    return img_filtered


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
    image_tiles = np.array(image_tiles)
    if image_tiles.ndim != flatfield.ndim:
        flatfield = np.expand_dims(flatfield, axis=0)
    if image_tiles.ndim != darkfield.ndim:
        darkfield = np.expand_dims(darkfield, axis=0)
    darkfield = darkfield[: image_tiles.shape[-2], : image_tiles.shape[-1]]
    if darkfield.shape != image_tiles.shape:
        raise ValueError(
            "Please, check the shape of the darkfield. "
            f"Image: {image_tiles.shape} - Darkfield: {darkfield.shape}"
        )
    if flatfield.shape != image_tiles.shape:
        raise ValueError(
            "Please, check the shape of the flatfield."
            f"Image: {image_tiles.shape} - Flatfield: {flatfield.shape}"
        )
    if baseline is None:
        baseline = np.zeros((image_tiles.shape[0],))
    baseline_indxs = tuple([slice(None)] + ([np.newaxis] * (image_tiles.ndim - 1)))
    negative_darkfield = np.where(image_tiles <= darkfield)
    positive_darkfield = np.where(image_tiles > darkfield)
    image_tiles[negative_darkfield] = 0
    image_tiles[positive_darkfield] = (
        image_tiles[positive_darkfield] - darkfield[positive_darkfield]
    )
    corrected_tiles = image_tiles / flatfield - baseline[baseline_indxs]
    return np.clip(corrected_tiles, 0, 65535).astype("uint16")


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
