#!/usr/bin/env python
"""End-to-end tile driver run: Zarr tile on disk -> destripe_zarr -> destriped Zarr + 2 multiscale
levels (reference zarr_destriper.py:909-1211), with the chunk codec in the loop.

    python tools/zarr_tile.py --planes 512 --codec zlib --workdir /dev/shm/dstr_tile
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/zarr_tile.py --planes 1024

Prints one JSON line: decode (read), device (pinned H2D + kernels + D2H) and encode (write) seconds of
the streaming pipeline and the whole-tile throughput.  Default codec: the reference's blosc-zstd-3 with
byte shuffle (zarr_destriper.py:1066-1074) for the input tile and the output (blosc1.py / dstr_blosc.cpp).
"""
import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aind_smartspim_destripe_b200 import destriper as DS  # noqa: E402
from aind_smartspim_destripe_b200 import distributed as D  # noqa: E402
from aind_smartspim_destripe_b200 import synthetic as S  # noqa: E402
from aind_smartspim_destripe_b200 import zarr_destriper as zd  # noqa: E402
from aind_smartspim_destripe_b200 import zarr_store as zs  # noqa: E402

NO_CELLS = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}
CELLS = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=512)
    ap.add_argument("--height", type=int, default=1600)
    ap.add_argument("--width", type=int, default=2000)
    ap.add_argument("--codec", choices=["none", "zlib", "blosc"], default="blosc")
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--workdir", default="/dev/shm/dstr_tile")
    args = ap.parse_args()
    rank, world, _ = D.init()
    Z, H, W = args.planes, args.height, args.width
    codec = {"none": None, "zlib": {"id": "zlib", "level": 1},
             "blosc": {"id": "blosc", "cname": "zstd", "clevel": 3, "shuffle": 1, "blocksize": 0}}[args.codec]
    work = Path(args.workdir)
    tile = work / "SPIM.ome.zarr" / "Ex_488_Em_525" / "471320_304840.zarr"
    deriv = work / "derivatives"
    flat, dark = S.synthetic_flat_dark(H, W)
    if rank == 0:
        shutil.rmtree(work, ignore_errors=True)
        deriv.mkdir(parents=True)
        DS._tiff_write(str(deriv / "DarkMaster_cropped.tif"), dark)
        zs.create_group(tile)
        src = zs.ZarrArray.create(tile / "0", (1, 1, Z, H, W), (1, 1, 128, 128, 128), np.uint16, codec, "/",
                                  threads=args.threads)
        unique = S.synthetic_stack(16, H, W, base_seed=7000, cells_every=4)
        for a in range(0, Z, 128):
            b = min(a + 128, Z)
            src[0, 0, a:b] = unique[np.arange(a, b) % 16]
        src.close()
    D.barrier()
    out = work / "results" / "Ex_488_Em_525" / tile.name
    t0 = time.perf_counter()
    t = zd.destripe_zarr(tile, "0", out, (64, H, W), 3072, args.threads, 1, None, work, deriv, [1.8, 1.8, 2.0],
                         {"no_cells_config": NO_CELLS, "cells_config": CELLS}, flatfield=flat, compressor=codec)
    D.barrier()
    total = D.max_over_ranks(time.perf_counter() - t0)
    wall = D.max_over_ranks(t["wall_s"])
    if rank == 0:
        lv = [zs.ZarrArray.open(out / str(k)) for k in range(3)]
        a, b = lv[0][0, 0, 5], lv[0][0, 0, 5 + 16]  # twins of the cyclic source
        nbytes = sum(f.stat().st_size for f in out.rglob("*") if f.is_file())
        print(json.dumps({
            "workload": f"zarr tile {Z}x{H}x{W} uint16 (chunks 128^3, codec {args.codec}) -> destripe_zarr -> "
                        f"zarr (1,1,64,128,128) + 2 multiscale levels",
            "n_gpus": world, "codec_threads": args.threads, "pipeline_wall_s": wall, "total_with_setup_s": total,
            "Mpixel_per_s": Z * H * W / wall / 1e6, "rank0": {k: round(float(v), 3) for k, v in t.items()},
            "levels": [list(x.shape) for x in lv], "output_bytes": nbytes, "twin_planes_identical": bool(np.array_equal(a, b)),
        }), flush=True)
        shutil.rmtree(work, ignore_errors=True)
    D.shutdown()


if __name__ == "__main__":
    main()
