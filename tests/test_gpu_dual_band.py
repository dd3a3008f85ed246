"""Classic dual-band mode (pystripe `filter_streaks`) against `oracle/dual_band.py`.

The mode is not in the reference snapshot (parity unpinned, stated in the oracle); the bar is the north star's:
uint16 outputs within +-1 count on >= 99.99 % of the pixels, max abs error printed."""
import numpy as np
import pytest

from _parity import U16_FRACTION, u16_agreement
from aind_smartspim_destripe_b200 import engine as E
from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import synthetic as S
from oracle import dual_band as OD

pytestmark = pytest.mark.gpu


def _plane(shape, seed):
    return S.synthetic_plane(shape[0], shape[1], seed=seed, n_cells=(shape[0] * shape[1]) // 4000, cell_peak=6000.0)


def test_histogram_and_otsu_threshold_match_the_oracle():
    st = np.stack([_plane((300, 420), s) for s in (1, 2, 3)])
    st[2, :7, :9] = 65535  # the last bin and a pair that shares a counter word
    st[2, 8, :5] = 65534
    eng = E.DestripeEngine(300, 420, max_planes=3)
    hist = eng.histogram_u16(st)
    for z in range(3):
        np.testing.assert_array_equal(hist[z], np.bincount(st[z].reshape(-1), minlength=65536))
        assert fl.otsu_from_counts(hist[z]) == OD.threshold_otsu_integer(st[z])
    eng.close()


@pytest.mark.parametrize("shape", [(256, 320), (402, 518), (1600, 2000)])
@pytest.mark.parametrize("sigma", [(256.0, 64.0), (128.0, 128.0), (256.0, 0.0), (0.0, 64.0), (0.0, 0.0)])
def test_filter_streaks_matches_oracle(shape, sigma):
    if shape == (1600, 2000) and sigma != (256.0, 64.0):
        pytest.skip("full-size plane: the dual-band case only")
    img = _plane(shape, seed=5)
    ref = OD.filter_streaks(img, list(sigma), level=0, crossover=10, threshold=-1)
    out = fl.filter_streaks(img, list(sigma), level=0, crossover=10, threshold=-1)
    assert out.dtype == np.uint16 and out.shape == img.shape
    frac, mx, exact = u16_agreement(out, ref)
    print(f"dual-band {shape} sigma {sigma}: within+-1 {frac:.6f} exact {exact:.4f} max abs {mx}")
    assert frac >= U16_FRACTION
    if sigma != (0.0, 0.0):
        assert np.any(out != img)


def test_filter_streaks_stack_flat_dark_and_explicit_threshold():
    shape = (256, 320)
    st = np.stack([_plane(shape, s) for s in range(10, 15)])
    flat = (1.0 + 0.3 * np.linspace(0, 1, shape[1], dtype=np.float32))[None, :].repeat(shape[0], 0)
    eng = E.DestripeEngine(shape[0], shape[1], max_planes=2)  # several batches per call
    out = fl.filter_streaks(st, [200.0, 50.0], level=3, crossover=15, threshold=400.0, flat=flat, dark=90, engine=eng)
    for z in range(st.shape[0]):
        ref = OD.filter_streaks(st[z], [200.0, 50.0], level=3, crossover=15, threshold=400.0, flat=flat, dark=90)
        frac, mx, exact = u16_agreement(out[z], ref)
        print(f"plane {z}: within+-1 {frac:.6f} exact {exact:.4f} max abs {mx}")
        assert frac >= U16_FRACTION
    # float32 planes need an explicit threshold
    with pytest.raises(ValueError):
        fl.filter_streaks(st[0].astype(np.float32), [200.0, 50.0], engine=eng)
    out_f = fl.filter_streaks(st[0].astype(np.float32), [200.0, 50.0], level=3, crossover=15, threshold=400.0, flat=flat,
                              dark=90, engine=eng)
    np.testing.assert_array_equal(out_f, out[0])
    eng.close()
