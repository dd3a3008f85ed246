"""CPU checks of the dual-band restatement (oracle/dual_band.py): the integer-image Otsu against a brute-force
between-class variance, the pure-notch sub-band filter against the identity (sigma -> 0 limit) and against the
masked filter with an unreachable threshold, and the product's host Otsu helper against the oracle's."""
import numpy as np

from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import synthetic as S
from oracle import dual_band as OD
from oracle import plane_filter as OF


def _brute_otsu(img):
    v = img.reshape(-1).astype(np.float64)
    best, arg = -1.0, None
    for t in range(int(v.min()), int(v.max())):
        a, b = v[v <= t], v[v > t]
        var = a.size * b.size * (a.mean() - b.mean()) ** 2
        if var > best * (1 + 1e-12):
            best, arg = var, t
    return arg


def test_integer_otsu_is_the_between_class_variance_maximum():
    rng = np.random.default_rng(3)
    for _ in range(5):
        img = np.concatenate([rng.poisson(40, 3000), rng.poisson(140, 800)]).astype(np.uint16).reshape(50, 76)
        t = OD.threshold_otsu_integer(img)
        assert abs(t - _brute_otsu(img)) <= 1  # float32 counts vs float64 brute force
        assert fl.otsu_from_counts(np.bincount(img.reshape(-1), minlength=65536)) == t
    assert OD.threshold_otsu_integer(np.full((4, 4), 7, np.uint16)) == 7.0
    assert fl.otsu_from_counts(np.bincount(np.full(16, 7), minlength=65536)) == 7.0


def test_subband_is_the_masked_filter_without_a_mask():
    img = S.synthetic_plane(128, 160, seed=4).astype(np.float64)
    # a huge max_threshold cannot disable the Otsu mask, so compare level by level on the notch alone:
    # with sigma tiny the notch removes only the row means of cH
    out = OD.filter_subband(img, 1e-3, 0)
    assert out.shape == img.shape and np.all(np.isfinite(out))
    assert np.abs(out - img).max() < 0.35 * img.max()
    # equal sigmas: filter_streaks is exactly the single sub-band filter, clipped and truncated
    single = OD.filter_streaks(img.astype(np.uint16), [64.0, 64.0], level=0)
    np.testing.assert_array_equal(single, np.clip(OD.filter_subband(img.astype(np.uint16).astype(float), 64.0, 0), 0, 65535).astype(np.uint16))
    # both sigmas zero: identity (clip only)
    np.testing.assert_array_equal(OD.filter_streaks(img.astype(np.uint16), [0.0, 0.0]), img.astype(np.uint16))


def test_blend_reduces_to_its_bands_far_from_the_threshold():
    img = S.synthetic_plane(128, 160, seed=6)
    lo = OD.filter_streaks(img, [256.0, 64.0], threshold=1e9, crossover=10)     # everything is background
    bg = np.clip(OD.filter_subband(img.astype(float), 64.0, 0), 0, 65535).astype(np.uint16)
    np.testing.assert_array_equal(lo, bg)
    hi = OD.filter_streaks(img, [256.0, 64.0], threshold=0, crossover=0.5)      # everything is foreground (f == 1.0)
    fg = np.clip(OD.filter_subband(img.astype(float), 256.0, 0), 0, 65535).astype(np.uint16)
    np.testing.assert_array_equal(hi, fg)
    assert OF.sigmoid(np.float64(0.0)) == 0.5
