#!/usr/bin/env python
"""BASELINE.json config 4: a SmartSPIM tile stack (Z x 1600 x 2000 uint16) streamed from host
memory through `destripe_volume` (reader thread -> pinned buffers -> 3-stream H2D/compute/D2H ->
writer thread), Z-slab sharded when launched with torchrun (one process per GPU, no collective).

    python tools/stream_tile.py --planes 512
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/stream_tile.py --planes 2000

Prints one JSON line with the read (decode stand-in: host memcpy from the source volume), device
(H2D + kernels + D2H) and write times and the wall-clock throughput of the whole job.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aind_smartspim_destripe_b200 import distributed as D  # noqa: E402
from aind_smartspim_destripe_b200 import synthetic as S  # noqa: E402
from aind_smartspim_destripe_b200 import zarr_destriper as zd  # noqa: E402

NO_CELLS = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}
CELLS = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=512)
    ap.add_argument("--height", type=int, default=1600)
    ap.add_argument("--width", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--unique", type=int, default=16)
    ap.add_argument("--pyramid", action="store_true")
    args = ap.parse_args()
    rank, world, local = D.init()
    Z, H, W = args.planes, args.height, args.width
    z0, z1 = zd.z_slab(Z, rank, world, align=args.chunk)
    # every rank materialises only its slab (the "tile on disk" stand-in)
    vol = S.synthetic_stack(z1 - z0, H, W, base_seed=7000 + z0 % args.unique, cells_every=4, n_unique=args.unique)
    out = np.zeros_like(vol)
    flat, dark = S.synthetic_flat_dark(H, W)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    pyr = None
    if args.pyramid:
        n = z1 - z0
        pyr = (np.zeros((n // 2, H // 2, W // 2), np.uint16), np.zeros((n // 4, H // 4, W // 4), np.uint16))
    D.barrier()
    t0 = time.perf_counter()
    t = zd.destripe_volume(vol, out, NO_CELLS, CELLS, shadow, chunk_planes=args.chunk, device=local,
                           pyramid_outputs=pyr)
    total = D.max_over_ranks(time.perf_counter() - t0)  # includes engine / pinned-buffer set-up
    wall = D.max_over_ranks(t["wall_s"])                 # the streaming pipeline itself
    planes = D.sum_over_ranks(float(z1 - z0))
    # plane independence: twins of the cyclic source give identical results
    ok = all(np.array_equal(out[z], out[z % args.unique]) for z in range(args.unique, z1 - z0, 37))
    ok = D.sum_over_ranks(0.0 if ok else 1.0) == 0.0
    if rank == 0:
        print(json.dumps({
            "workload": f"tile stack {Z}x{H}x{W} uint16 streamed, dual config + dark/flat, chunks of {args.chunk}",
            "n_gpus": world, "planes": int(planes), "pipeline_wall_s": wall, "total_with_setup_s": total,
            "Mpixel_per_s": planes * H * W / wall / 1e6,
            "rank0": {k: round(v, 3) for k, v in t.items()},
            "twin_planes_identical": bool(ok), "pyramid": bool(args.pyramid),
        }), flush=True)
    D.shutdown()


if __name__ == "__main__":
    main()
