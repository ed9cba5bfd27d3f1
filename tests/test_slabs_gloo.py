"""Host-side logic of the N>1 path on CPU: slab partitioning, neighbour ring, the blob exchange and the
read-back gather over torch.distributed with the gloo backend, world_size 2 and 3 (no GPU, no compute)."""
import os
import socket

import numpy as np
import pytest

from simuverse_b200.slabs import neighbours, slab_bounds


def test_slab_bounds_partition_every_lattice():
    for ny in (3, 77, 375, 4096, 16384, 131072):
        for world in (1, 2, 3, 4, 5, 8):
            if ny // world < 2 and world > 1:
                continue
            edges = [slab_bounds(ny, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == ny
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_neighbour_ring():
    assert neighbours(0, 1) == (0, 0)
    assert neighbours(0, 2) == (1, 1) and neighbours(1, 2) == (0, 0)
    assert [neighbours(r, 4) for r in range(4)] == [(3, 1), (0, 2), (1, 3), (2, 0)]
    for world in (2, 3, 8):  # my up-neighbour's down-neighbour is me
        for r in range(world):
            up, down = neighbours(r, world)
            assert neighbours(up, world)[1] == r and neighbours(down, world)[0] == r


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ny, nx):
    import torch.distributed as dist

    from simuverse_b200.slabs import exchange_blobs, gather_rows, neighbours, slab_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 256-byte blobs as lbm_ipc_export produces them (content is opaque to the transport)
        mine = bytes([rank]) * 200 + bytes(range(56))
        allb, up, down = exchange_blobs(dist, mine)
        u, d = neighbours(rank, world)
        assert len(allb) == world and all(len(b) == 256 for b in allb)
        assert up == bytes([u]) * 200 + bytes(range(56)) and down == bytes([d]) * 200 + bytes(range(56))
        # read-back gather: each rank contributes its rows of a known global array
        y0, y1 = slab_bounds(ny, rank, world)
        glob = np.arange(9 * ny * nx, dtype=np.float32).reshape(9, ny, nx)
        got = gather_rows(dist, glob[:, y0:y1, :].copy(), ny)
        assert got.shape == glob.shape and np.array_equal(got, glob)
        info = np.arange(ny * nx, dtype=np.int32).reshape(1, ny, nx)
        assert np.array_equal(gather_rows(dist, info[:, y0:y1, :].copy(), ny), info)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ny", [(2, 16), (2, 77), (3, 31)])
def test_blob_exchange_and_gather_over_gloo(world, ny):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), ny, 13), nprocs=world, join=True)
