"""Shared helpers for the GPU parity tests (oracle = checker, never the thing under test)."""
import numpy as np
from scipy import fftpack

from oracle import dwt as odwt
from oracle import otsu as ootsu
from oracle import plane_filter as OF

# north-star tolerances (BASELINE.json): uint16 within +-1 count on >= 99.99 % of pixels,
# fp32 intermediates within 1e-4 relative error
U16_FRACTION = 0.9999
REL_TOL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def u16_agreement(out, ref):
    d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
    return float((d <= 1).mean()), int(d.max()), float((d == 0).mean())


def oracle_levels(plane_f32, level):
    """cA_l, cH_l for l = 1..L exactly as pywt.wavedec2 would produce them (float32 path)."""
    lg = np.log(1.0 + plane_f32)
    out = []
    a = lg
    L = odwt.dwtn_max_level(plane_f32.shape, "db3") if level is None else level
    for _ in range(L):
        a, (ch, cv, cd) = odwt.dwt2(a, "db3")
        out.append((a, ch))
    return out


def oracle_level_filter(ch, thr, s):
    """filtering.py:195-217 for one band given the threshold; returns (dH, mask, median)."""
    ch_power = np.sqrt(ch**2)
    mask = ch_power > thr
    background = ch * (1 - mask)
    med = np.median(background, axis=-1)
    inp = background + med[:, None] * mask
    g = OF.notch(ch.shape[-1], s)
    bgf = fftpack.irfft(fftpack.rfft(inp, axis=-1) * g)
    ch_f = ch * mask + bgf * (1 - mask)
    return ch_f - ch, mask, med


def oracle_threshold(ch, max_threshold):
    q = ch**2
    raw = ootsu.threshold_otsu(q)
    sq = np.sqrt(raw)
    return raw, min(max_threshold, sq)
