"""CPU tests (no GPU): host-side logic -- descriptors, synthetic workloads, sharding, and the
N>1 path's collectives on the gloo backend with world_size 2."""
import os
import sys

import numpy as np
import pytest

from axom_b200 import synth
from axom_b200.bvh import make_desc
from axom_b200.dist import morton_partition, slab_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_make_desc_aos_and_soa():
    a = np.arange(12, dtype=np.float64).reshape(2, 6)
    k = make_desc(a, 6)
    assert k.count == 2 and k.desc.stride_bytes == 48 and k.desc.ncomp == 6
    assert k.desc.comp[1] - k.desc.comp[0] == 8
    k = make_desc(tuple(np.ascontiguousarray(a[:, c]) for c in range(6)), 6)
    assert k.count == 2 and k.desc.stride_bytes == 8
    with pytest.raises(ValueError):
        make_desc((a[:, 0], a[:, 1]), 6)


def test_icosphere_is_closed_and_outward():
    for n in (1, 2, 7):
        x, y, z, conn = synth.icosphere(n)
        assert len(conn) == 20 * n * n and len(x) == 10 * n * n + 2
        e = np.concatenate([conn[:, [0, 1]], conn[:, [1, 2]], conn[:, [2, 0]]])
        keys = set(map(tuple, e.tolist()))
        assert len(keys) == len(e) and all((b, a) in keys for a, b in keys)  # watertight, consistent winding
        P = np.stack([x, y, z], 1)
        assert np.allclose(np.linalg.norm(P, axis=1), 0.5)
        nrm = np.cross(P[conn[:, 1]] - P[conn[:, 0]], P[conn[:, 2]] - P[conn[:, 0]])
        assert ((nrm * P[conn].sum(axis=1)).sum(axis=1) > 0).all()


def test_sublattice_is_bit_identical_to_full_grid():
    sys.path.insert(0, ROOT)
    import bench
    full = synth.uniform_grid_points(-1, 1, bench.GRID).reshape(bench.GRID, bench.GRID, bench.GRID, 3)
    sub = bench.sublattice(64)
    assert np.array_equal(sub, full[::64, ::64, ::64].reshape(-1, 3))


def test_slab_ranges_tile_the_grid():
    for world in (1, 2, 3, 4, 8):
        r = [slab_range(256, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 256
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_morton_partition_covers_all_cells():
    c = np.random.default_rng(0).random((1000, 3))
    parts = morton_partition(c, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(1000))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from axom_b200.dist import allreduce_max_scalar, allreduce_min_
    from oracle import oracle as O
    # partitioned-surface closest point (config 5): each rank owns a Morton range of the surface,
    # evaluates ALL queries against its part, then one elementwise MIN
    x, y, z, conn = synth.icosphere(6)
    P = np.stack([x, y, z], 1)
    parts = morton_partition(P[conn].mean(axis=1), world)
    q = synth.uniform_grid_points(-1, 1, 7)
    phi, _, _ = O.SignedDistance(x, y, z, conn[parts[rank]], compute_sign=False).compute(q)
    t = torch.from_numpy(phi * phi)
    allreduce_min_(t)
    full, _, _ = O.SignedDistance(x, y, z, conn, compute_sign=False).compute(q)
    ok = np.allclose(np.sqrt(t.numpy()), full, rtol=1e-15, atol=0)
    # max-over-ranks timing reduction used by bench.py
    m = allreduce_max_scalar(float(rank + 1))
    out.put((rank, bool(ok), m))
    dist.destroy_process_group()


def test_gloo_world2_min_reduce_and_max_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert all(m == 2.0 for _, _, m in res)
