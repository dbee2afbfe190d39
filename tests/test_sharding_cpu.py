"""Host-side logic of the x-slab sharding, including the count exchange over torch.distributed (gloo,
world_size 2, CPU).  The oracle plays the role of each rank's local extractor: MC has no cross-voxel state,
so running it on a slab with a halo plane and rebasing the faces must reproduce the unsharded mesh."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_all_voxel_planes(pkg):
    sb = pkg.sharding.slab_bounds
    for nx in (2, 3, 9, 128, 1025):
        for world in (1, 2, 3, 8):
            planes = []
            for r in range(world):
                xa, xb = sb(nx, world, r)
                planes += list(range(xa, xb - 1))
                assert xb <= nx
            assert planes == list(range(nx - 1)), (nx, world)
    assert sb(1025, 8, 0) == (0, 129) and sb(1025, 8, 7) == (896, 1025)


def test_exclusive_bases(pkg):
    counts = np.array([[10, 5], [0, 0], [7, 3]])
    assert pkg.sharding.exclusive_bases(counts, 0) == (0, 0)
    assert pkg.sharding.exclusive_bases(counts, 2) == (10, 5)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import harness as oracle

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    shape = (37, 20, 26)
    s = pkg.synth.gyroid(shape)
    xa, xb = pkg.sharding.slab_bounds(shape[0], world, rank)
    # rank-local extraction of voxel planes [xa, xb-1) (the oracle stands in for the GPU on this CPU box)
    v, f = oracle.isosurface(s, 0, iso_is_f32=True, xrange=(xa, xb - 1))
    counts = pkg.sharding.allgather_counts(len(v), len(f))
    vb, fb = pkg.sharding.exclusive_bases(counts, rank)
    q.put((rank, v, f + vb, vb, fb, counts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_stitch_equals_unsharded(pkg, oracle):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    v, f = pkg.sharding.stitch([(r[1], r[2]) for r in res])
    s = pkg.synth.gyroid((37, 20, 26))
    vo, fo = oracle.isosurface(s, 0, iso_is_f32=True)
    assert np.array_equal(v, vo) and np.array_equal(f, fo)
    assert res[0][5] == res[1][5] and res[1][3] == res[0][5][0][0]
