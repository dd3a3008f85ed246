"""oracle/otsu.py: threshold_otsu restatement (skimage 0.24.0 over numpy 1.26.4 histogram)."""
import numpy as np

from oracle import otsu


def test_constant_image_returns_value():
    assert otsu.threshold_otsu(np.full((8, 8), 3.5, np.float32)) == np.float32(3.5)


def test_histogram_matches_installed_numpy_up_to_edge_rounding():
    rng = np.random.default_rng(0)
    q = (rng.standard_normal(200_000).astype(np.float32)) ** 2
    counts, edges = otsu.histogram_f32(q)
    ref_counts, ref_edges = np.histogram(q, bins=256)
    assert counts.sum() == q.size == ref_counts.sum()
    assert edges.dtype == np.float32 and edges[0] == q.min() and edges[-1] == q.max()
    # numpy >= 2 evaluates linspace in float32, 1.26.4 in float64: edges agree to 1 ulp,
    # counts may differ only for values sitting on such an edge
    np.testing.assert_allclose(edges, ref_edges.astype(np.float32), rtol=3e-7)
    assert np.abs(counts - ref_counts).sum() <= 8


def test_histogram_edge_semantics_last_bin_closed():
    q = np.array([0.0, 1.0, 2.0, 255.9, 256.0], dtype=np.float32)
    counts, edges = otsu.histogram_f32(q)
    assert counts[0] == 1 and counts[1] == 1 and counts[2] == 1 and counts[255] == 2


def test_bimodal_threshold_between_modes_and_is_bin_centre():
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.normal(1.0, 0.1, 50_000), rng.normal(5.0, 0.2, 20_000)]).astype(np.float32)
    thr = otsu.threshold_otsu(a)
    assert 1.2 < thr < 4.5  # flat variance between the modes: first arg-max
    counts, edges = otsu.histogram_f32(a)
    centers = (edges[:-1] + edges[1:]) / 2.0
    assert thr in centers


def test_two_values():
    a = np.array([0.0] * 10 + [1.0] * 5, dtype=np.float32)
    thr = otsu.threshold_otsu(a)
    # variance12 is flat between the two populated bins; first arg-max -> bin 0 centre
    counts, edges = otsu.histogram_f32(a)
    assert thr == (edges[0] + edges[1]) / 2.0


def test_float64_path_uses_numpy_histogram():
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.normal(1.0, 0.1, 5000), rng.normal(5.0, 0.2, 2000)])
    thr = otsu.threshold_otsu(a)
    assert 1.2 < thr < 4.5 and isinstance(thr, np.floating)
