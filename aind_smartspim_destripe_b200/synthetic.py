"""Seeded synthetic SmartSPIM-like planes with streaks (SURVEY.md §8d plane model).

Streaks are (nearly) constant along axis -1 (X), the axis the reference's row FFT runs
along (``/root/reference/code/aind_smartspim_destripe/filtering.py:206``).
Plane ``z`` of a stack uses ``seed = base_seed + z``.
"""

from __future__ import annotations

import numpy as np


def _smooth_rows(rng: np.random.Generator, n: int, sigma: float) -> np.ndarray:
    """randn(n) smoothed by a Gaussian of ``sigma`` samples (reflect borders)."""
    r = int(max(1, np.ceil(4 * sigma)))
    t = np.arange(-r, r + 1)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k /= k.sum()
    v = rng.standard_normal(n)
    return np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid")


def synthetic_plane(
    H: int,
    W: int,
    seed: int = 0,
    n_cells: int = None,
    cell_peak: float = 3000.0,
    streak_gain: float = 0.15,
    partial_streaks: int = 8,
) -> np.ndarray:
    """One (H, W) uint16 plane: smooth tissue + Poisson noise + cells + streaks."""
    if n_cells is None:
        n_cells = max(1, (H * W) // 7000)  # 600 cells on a 2048 x 2048 plane
    rng = np.random.default_rng(seed)
    y = np.arange(H, dtype=np.float32)[:, None]
    x = np.arange(W, dtype=np.float32)[None, :]
    img = 120.0 + 60.0 * np.sin(y / 200.0) * np.cos(x / 300.0)
    img = img + rng.poisson(30.0, size=(H, W)).astype(np.float32)

    # cells: Gaussian blobs, sigma 3 px, stamped as 25x25 patches
    if n_cells > 0:
        t = np.arange(-12, 13, dtype=np.float32)
        blob = np.exp(-0.5 * (t[:, None] ** 2 + t[None, :] ** 2) / 9.0).astype(np.float32)
        cy = rng.integers(0, H, size=n_cells)
        cx = rng.integers(0, W, size=n_cells)
        amp = cell_peak * (0.5 + rng.random(n_cells)).astype(np.float32)
        for k in range(n_cells):
            y0, y1 = max(cy[k] - 12, 0), min(cy[k] + 13, H)
            x0, x1 = max(cx[k] - 12, 0), min(cx[k] + 13, W)
            img[y0:y1, x0:x1] += amp[k] * blob[
                y0 - cy[k] + 12 : y1 - cy[k] + 12, x0 - cx[k] + 12 : x1 - cx[k] + 12
            ]

    # multiplicative streaks, constant along axis -1
    gain = 1.0 + streak_gain * _smooth_rows(rng, H, 2.0)
    img = img * gain[:, None].astype(np.float32)
    # a few partial-length streaks starting at a random x
    for _ in range(partial_streaks):
        r0 = int(rng.integers(0, H))
        h = int(rng.integers(1, 6))
        x0 = int(rng.integers(0, W))
        img[r0 : r0 + h, x0:] *= np.float32(1.0 + 0.2 * rng.standard_normal())

    return np.clip(img, 0, 65535).astype(np.uint16)


def _plane_job(job):
    H, W, seed, kw = job
    return synthetic_plane(H, W, seed=seed, **kw)


def synthetic_stack(
    Z: int,
    H: int,
    W: int,
    base_seed: int = 0,
    cells_every: int = 0,
    n_unique: int = 0,
    workers: int = 0,
    **plane_kwargs,
) -> np.ndarray:
    """(Z, H, W) uint16 stack.

    ``cells_every=k`` (k > 0) makes every k-th plane a dense-bright-cells plane (mean of
    pixels >= 384 above 2500) so that both branches of the reference's per-plane dispatch
    (``filtering.py:462``) are exercised.  ``n_unique`` > 0 generates only that many
    distinct planes and repeats them cyclically (large benchmark stacks).
    """
    n_gen = Z if n_unique <= 0 else min(Z, n_unique)
    jobs = []
    for z in range(n_gen):
        kw = dict(plane_kwargs)
        if cells_every > 0 and z % cells_every == cells_every - 1:
            kw.update(n_cells=max(kw.get("n_cells") or 0, (H * W) // 2000), cell_peak=30000.0)
        jobs.append((H, W, base_seed + z, kw))
    if workers and workers > 1 and n_gen >= 8:
        # planes are independent and seeded individually: generate them in worker processes
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        with ProcessPoolExecutor(max_workers=min(workers, n_gen), mp_context=mp.get_context("fork")) as ex:
            planes = list(ex.map(_plane_job, jobs, chunksize=max(1, n_gen // (4 * workers))))
    else:
        planes = [_plane_job(j) for j in jobs]
    out = np.empty((Z, H, W), dtype=np.uint16)
    for z in range(Z):
        out[z] = planes[z % n_gen]
    return out


def synthetic_flat_dark(H: int, W: int, seed: int = 1234):
    """Smooth flat field in [1, 2] (float32) and a constant ~100-count dark (uint16)."""
    rng = np.random.default_rng(seed)
    y = np.linspace(-1.0, 1.0, H, dtype=np.float32)[:, None]
    x = np.linspace(-1.0, 1.0, W, dtype=np.float32)[None, :]
    flat = (1.0 + 0.5 * (1.0 + np.cos(1.3 * y) * np.cos(0.9 * x + 0.2))).astype(np.float32)
    flat = np.clip(flat, 1.0, 2.0)
    dark = (100 + rng.integers(-3, 4, size=(H, W))).astype(np.uint16)
    return flat, dark
