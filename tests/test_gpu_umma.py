"""GPU parity of the tcgen05 row filter (csrc/dstr_notch_umma.cuh): the same oracle checks as the
CUDA-core kernel, plus agreement between the two paths.  The tensor-core path is opt-in
(`DestripeEngine.set_umma(True)` / DSTR_UMMA=1) because it measures slower on B200."""
import numpy as np
import pytest

from _parity import U16_FRACTION, oracle_level_filter, rel_err, u16_agreement
from aind_smartspim_destripe_b200 import engine as E
from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import synthetic as S
from oracle import plane_filter as OF

pytestmark = pytest.mark.gpu


def _levels_dh(eng, img, p):
    Z = 1 if img.ndim == 2 else img.shape[0]
    eng.set_debug_stop(E.STAGE_FILTER)
    eng.set_subchunk(Z)  # debug_fetch sees the last sub-chunk only: keep the whole stack in one
    eng.filter_chunk(img[None] if img.ndim == 2 else img, p, out_dtype=np.float32)
    return [eng.debug_fetch(E.FETCH_CH, l, Z) for l in range(1, eng.max_level + 1)]


@pytest.mark.parametrize("shape,sigma", [((403, 517), 64), ((1600, 2000), 128), ((2048, 2048), 64), ((2048, 2048), 128)])
def test_umma_row_filter_matches_oracle_on_device_coefficients(shape, sigma):
    img = S.synthetic_plane(shape[0], shape[1], seed=2)
    eng = E.DestripeEngine(shape[0], shape[1], max_planes=2)
    eng.set_umma(True)
    p = E.make_params(dict(level=None, sigma=sigma, max_threshold=12))
    eng.set_debug_stop(E.STAGE_OTSU)
    eng.filter_chunk(img[None], p, out_dtype=np.float32)
    ch = [eng.debug_fetch(E.FETCH_CH, l, 1)[0] for l in range(1, eng.max_level + 1)]
    thr = [eng.debug_fetch(E.FETCH_STATS, l, 1)[0][3] for l in range(1, eng.max_level + 1)]
    dh = _levels_dh(eng, img, p)
    wf = sigma / min(shape)
    for l in range(1, eng.max_level + 1):
        ref, mask, med = oracle_level_filter(ch[l - 1], thr[l - 1], ch[l - 1].shape[0] * wf)
        err = np.abs(dh[l - 1][0] - ref).max() / max(np.abs(ch[l - 1]).max(), 1e-30)
        print(f"umma shape {shape} level {l} (W_l {ch[l - 1].shape[1]}): dH err/max|cH| {err:.2e}")
        assert np.all(dh[l - 1][0][mask] == 0)
        assert err < 2e-5
    eng.close()


def test_umma_and_cuda_core_row_filters_agree_on_a_stack():
    """Nine planes (more items than the two operand buffers of a CTA, partial last item) through both kernels."""
    st = S.synthetic_stack(9, 512, 640, base_seed=40)
    p = E.make_params(dict(level=None, sigma=128, max_threshold=12))
    eng = E.DestripeEngine(512, 640, max_planes=9)
    ref = _levels_dh(eng, st, p)
    eng.set_umma(True)
    got = _levels_dh(eng, st, p)
    for l, (a, b) in enumerate(zip(got, ref), start=1):
        scale = max(np.abs(b).max(), 1e-30)
        print(f"level {l}: umma vs cuda-core dH {np.abs(a - b).max() / scale:.2e} of max|dH|")
        assert np.abs(a - b).max() <= 2e-5 * max(scale, 1e-3)
    eng.close()


@pytest.mark.parametrize("shape", [(402, 518), (1600, 2000)])  # even sizes: the reference rejects odd planes with a flatfield
def test_umma_end_to_end_dispatch_uint16(shape, production_configs):
    """filter_stripes semantics (both configs in one chunk: the kernel's two table sweeps) within the
    north-star tolerance of the oracle."""
    no_cells, cells = production_configs
    Z = 6
    st = S.synthetic_stack(Z, shape[0], shape[1], base_seed=70, cells_every=2)
    flat, dark = S.synthetic_flat_dark(*shape)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    eng = E.DestripeEngine(shape[0], shape[1], max_planes=Z)
    eng.set_umma(True)
    out = fl.filter_planes(st, "0_0", no_cells, cells, shadow, 2500, engine=eng)
    used = set()
    for z in range(Z):
        ref = OF.filter_stripes(st[z].astype(np.float32), "0_0", no_cells, cells, shadow, 2500)
        fg, bg, _ = OF.get_foreground_background_mean(st[z].astype(np.float32))
        used.add(bool(fg > bg and fg > 2500))
        frac, mx, exact = u16_agreement(out[z], ref)
        print(f"umma dispatch {shape} plane {z}: within+-1 {frac:.6f} exact {exact:.4f} max abs {mx}")
        assert frac >= U16_FRACTION
    assert used == {True, False}  # both configs ran
    eng.close()
