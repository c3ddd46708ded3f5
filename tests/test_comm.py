"""The exchange steps inside the library (include/axb200.h axb_comm_*, csrc/comm.cuh): NCCL communicator,
partitioned-surface MIN (BASELINE config C5) and DistributedClosestPoint::computeClosestPoints as one C call
(quest/detail/DistributedClosestPointImpl.hpp:687-693, :737-851).

CPU: argument checks, loud failure without a device.  GPU, one device: the whole protocol on a one-rank communicator
(every kernel and every NCCL call runs), from Python against the oracle and from plain host C++ (tests/cpp/dcp_nccl_test.cpp)
against brute force.  GPU, two or more devices: the C++ test forks one process per rank -- no torchrun, no mpirun, the NCCL
id goes through a file -- and a torch.multiprocessing run compares the Python mirror with the reference's sequential ring."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DBL_MAX = float(np.finfo(np.float64).max)


def _gpu_count():
    from axom_b200 import _lib
    return max(_lib.lib().axb_device_count(), 0)


def _nccl_env():
    import torch
    env = dict(os.environ)
    sp = os.path.dirname(os.path.dirname(torch.__file__))
    cand = os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2")
    if os.path.exists(cand):
        env.setdefault("AXB_NCCL_LIB", cand)
    return env


def test_comm_argument_checks_and_no_device():
    from axom_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    idb = (C.c_uint8 * 128)()
    assert L.axb_comm_create(C.byref(h), 0, 0, idb, 0) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_comm_create(C.byref(h), 2, 2, idb, 0) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_comm_create(C.byref(h), 2, 0, None, 0) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_comm_create(None, 1, 0, idb, 0) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_comm_get_rank(None, None, None) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_sd_compute_distances_minreduce(None, None, None, 0, None, 0) == _lib.AXB_ERR_BAD_ARG
    assert L.axb_dcp_compute_closest_points(None, None, None, 0, 0, None, None, None, None, None) == _lib.AXB_ERR_BAD_ARG
    if L.axb_device_count() <= 0:
        assert L.axb_comm_create(C.byref(h), 1, 0, idb, 0) == _lib.AXB_ERR_NO_DEVICE  # no CPU fallback for the exchange either


def test_cpp_nccl_test_builds():
    import __graft_entry__ as g
    g.build_cpp_tests()
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "bin", "dcp_nccl_test"))


def _run_cpp(nranks, tmp_path):
    import __graft_entry__ as g
    g.build_cpp_tests()
    exe = os.path.join(ROOT, "tests", "cpp", "bin", "dcp_nccl_test")
    r = subprocess.run([exe, str(nranks), str(tmp_path)], capture_output=True, text=True, timeout=600, env=_nccl_env())
    assert r.returncode == 0 and ("dcp_nccl_test: OK (%d ranks)" % nranks) in r.stdout, r.stdout + r.stderr
    return r.stdout


@pytest.mark.gpu
def test_cpp_host_drives_the_whole_protocol_on_one_rank(tmp_path):
    out = _run_cpp(1, tmp_path)
    assert "NCCL:" in out


@pytest.mark.gpu
def test_cpp_host_forks_one_process_per_gpu(tmp_path):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs (NCCL does not place two ranks on one device); run with gpurun --gpus 2")
    for k in sorted({2, min(n, 4), n}):
        out = _run_cpp(k, tmp_path)
        assert "cross-rank ties" in out


@pytest.mark.gpu
@pytest.mark.parametrize("nd", [3, 2])
def test_one_rank_communicator_matches_oracle(oracle, nd):
    """every kernel of the collective path and every NCCL call, on a communicator of one rank"""
    import torch
    from axom_b200 import DistributedClosestPoint
    from axom_b200.comm import Comm, unique_id
    rng = np.random.default_rng(5)
    pts = rng.random((5000, nd))
    pts[100:104] = pts[100]  # ties inside the rank: first in traversal order
    q = rng.random((6000, nd)) * 1.4 - 0.2
    q[:10] = pts[100]
    comm = Comm(1, 0, unique_id(), 0)
    assert "version" in comm.library()
    for th in (None, 0.05, 0.0):
        d = DistributedClosestPoint(nd, device=0)
        d.setComm(comm)
        if th is not None:
            d.setDistanceThreshold(th)
        d.setObjectMesh([(pts[:2000], 7), (pts[2000:], 9)])
        d.generateBVHTree()
        got = d.computeClosestPoints(torch.from_numpy(q).cuda())
        dom = np.where(np.arange(len(pts)) < 2000, 7, 9).astype(np.int32)
        want = oracle.DistributedClosestPointRank(pts, dom, nd).compute_local(0, q, None, _DBL_MAX if th is None else th * th)
        for k in ("cp_index", "cp_domain_index", "cp_rank", "cp_coords", "cp_distance"):
            assert np.array_equal(got[k].cpu().numpy(), want[k], equal_nan=True), (k, th)
        empty = d.computeClosestPoints(np.empty((0, nd)))
        assert all(v.shape[0] == 0 for v in empty.values())
    b, n = comm.traffic()
    assert b > 0 and n > 0


@pytest.mark.gpu
def test_minreduce_on_one_rank_equals_compute_distances():
    import torch
    from axom_b200 import SignedDistance, synth
    from axom_b200.comm import Comm, unique_id
    x, y, z, conn = synth.icosphere(12)
    q = synth.random_points(30000, seed=4) * 2.4 - 1.2
    sd = SignedDistance(x, y, z, conn, computeSign=False, device=0)
    comm = Comm(1, 0, unique_id(), 0)
    want, _, _ = sd.computeDistances(q)
    got_h = sd.computeDistancesMinReduce(comm, q)
    got_d = sd.computeDistancesMinReduce(comm, torch.from_numpy(q).cuda())
    assert np.array_equal(want, got_h) and np.array_equal(want, got_d.cpu().numpy())
    signed = SignedDistance(x, y, z, conn, computeSign=True, device=0)
    from axom_b200._lib import AxbError
    with pytest.raises(AxbError):
        signed.computeDistancesMinReduce(comm, q)


def _mp_worker(rank, world, id_file, nd, threshold, out_dir):
    import torch
    from axom_b200 import DistributedClosestPoint, SignedDistance, synth
    from axom_b200.comm import Comm
    from test_distributed_closest_point import _cloud_parts
    torch.cuda.set_device(rank)
    comm = Comm.from_file(id_file, world, rank, rank)
    parts, q = _cloud_parts(nd)
    parts = parts[:world] if world <= len(parts) else parts + [parts[0][:0]] * (world - len(parts))
    blocks = np.array_split(q, world)
    d = DistributedClosestPoint(nd, device=rank)
    d.setComm(comm)
    if threshold is not None:
        d.setDistanceThreshold(threshold)
    d.setObjectMesh([parts[rank]])
    d.generateBVHTree()
    got = d.computeClosestPoints(torch.from_numpy(blocks[rank]).cuda(rank))
    np.savez(os.path.join(out_dir, "dcp_%d.npz" % rank), **{k: v.cpu().numpy() for k, v in got.items()})
    # C5: the icosphere's cells dealt to the ranks in Morton ranges, the same queries everywhere
    from axom_b200.dist import morton_partition
    x, y, z, conn = synth.icosphere(10)
    cent = np.stack([x[conn].mean(axis=1), y[conn].mean(axis=1), z[conn].mean(axis=1)], axis=1)
    mine = morton_partition(cent, world)[rank]
    sd = SignedDistance(x, y, z, conn[mine], computeSign=False, device=rank)
    qq = synth.random_points(20000, seed=9) * 2.4 - 1.2
    dist = sd.computeDistancesMinReduce(comm, qq)
    np.save(os.path.join(out_dir, "c5_%d.npy" % rank), dist)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,threshold", [(3, None), (2, 0.04)])
def test_python_mirror_over_nccl_equals_reference_ring(oracle, tmp_path, nd, threshold):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs; run with gpurun --gpus 2")
    import torch.multiprocessing as mp
    from test_distributed_closest_point import _cloud_parts, _ring
    world = min(n, 4)
    id_file = str(tmp_path / "nccl_id")
    os.environ.update({k: v for k, v in _nccl_env().items() if k == "AXB_NCCL_LIB"})
    mp.spawn(_mp_worker, args=(world, id_file, nd, threshold, str(tmp_path)), nprocs=world, join=True)
    parts, q = _cloud_parts(nd)
    parts = parts[:world]
    blocks = np.array_split(q, world)
    ranks = [oracle.DistributedClosestPointRank(p, None, nd) for p in parts]
    sq = _DBL_MAX if threshold is None else threshold * threshold
    for owner in range(world):
        want = _ring(lambda r: ranks[r], parts, blocks[owner], owner, sq)
        got = np.load(os.path.join(str(tmp_path), "dcp_%d.npz" % owner))
        for k in ("cp_index", "cp_domain_index", "cp_rank", "cp_coords", "cp_distance"):
            assert np.array_equal(got[k], want[k], equal_nan=True), (owner, k)
    # C5 against the oracle on the WHOLE surface
    from axom_b200 import synth
    x, y, z, conn = synth.icosphere(10)
    qq = synth.random_points(20000, seed=9) * 2.4 - 1.2
    want, _, _ = oracle.SignedDistance(x, y, z, conn, compute_sign=False).compute(qq)
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "c5_%d.npy" % r)), want), r
