"""world_size-2 test (gloo, CPU) of the one-process-per-GPU plumbing: rendezvous, barrier,
max/sum reductions used by bench.py, and the Z-slab partition (no data-path collective)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_planes, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from aind_smartspim_destripe_b200 import distributed as D
    from aind_smartspim_destripe_b200.zarr_destriper import tiles_of_rank, z_slab

    r, w, local = D.init(backend="gloo")
    assert (r, w, local) == (rank, world, rank)
    z0, z1 = z_slab(n_planes, r, w)
    D.barrier()
    assert D.max_over_ranks(float(rank + 1)) == float(world)
    assert D.sum_over_ranks(float(z1 - z0)) == float(n_planes)
    # every rank "processes" only its slab of a shared volume (here: marks it)
    np.save(os.path.join(out_dir, f"slab_{rank}.npy"), np.array([z0, z1]))
    # channel of 5 tiles: whole tiles dealt round-robin (BASELINE config 5), every tile owned exactly once
    tiles = [f"tile_{i}.zarr" for i in range(5)]
    mine, per_tile = tiles_of_rank(tiles, r, w)
    assert per_tile and mine == tiles[r::w]
    assert D.sum_over_ranks(float(len(mine))) == float(len(tiles))
    # fewer tiles than ranks: all ranks share every tile as Z-slabs
    assert tiles_of_rank(tiles[:1], r, w) == (tiles[:1], False)
    assert tiles_of_rank(tiles, r, w, tile_parallel=False) == (tiles, False)
    D.barrier()
    D.shutdown()


def test_two_ranks_gloo(tmp_path):
    world, n_planes = 2, 2000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_planes, str(tmp_path)), nprocs=world, join=True)
    slabs = [np.load(tmp_path / f"slab_{r}.npy") for r in range(world)]
    assert slabs[0][0] == 0 and slabs[0][1] == slabs[1][0] and slabs[1][1] == n_planes
    assert slabs[0][1] % 64 == 0
