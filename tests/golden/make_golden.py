"""Regenerates tests/golden/*.npz from the CPU oracle (the reference itself cannot be imported
here: pywt / skimage are not installable).  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from aind_smartspim_destripe_b200 import synthetic as S  # noqa: E402
from oracle import plane_filter as OF  # noqa: E402

NO_CELLS = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}
CELLS = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}


def main():
    H, W = 96, 112
    st = S.synthetic_stack(2, H, W, base_seed=7, cells_every=2)
    flat, dark = S.synthetic_flat_dark(H, W)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    out = {"planes": st, "flat": flat, "dark": dark}
    for z in range(2):
        img = st[z].astype(np.float32)
        trace = {}
        ls = OF.log_space_fft_filtering(img, _trace=trace, **NO_CELLS)
        out[f"logspace_nocells_{z}"] = np.clip(ls, 0, 65535).astype(np.uint16)
        out[f"thresholds_nocells_{z}"] = np.array([float(l["threshold"]) for l in trace["levels"]], np.float32)
        out[f"filter_stripes_shadow_{z}"] = OF.filter_stripes(img, "0_0", NO_CELLS, CELLS, shadow, 2500)
        fg, bg, _ = OF.get_foreground_background_mean(img)
        out[f"fg_bg_{z}"] = np.array([fg, bg], np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plane_96x112_seed7.npz"), **out)
    print("written", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
