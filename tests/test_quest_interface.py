"""The reference's legacy process-global signed-distance surface (QUEST_signed_distance_*, include/axb200_quest.h) and the
input side of the path (STL reader, vertex welding): SURVEY.md 8(f) rank 2.

CPU part (no GPU): quest::STLReader / quest::weldTriMeshVertices host logic against the real reference and the golden
fixture; the error behaviour of the process-global API (death tests of quest_signed_distance_interface.cpp:245-330).
GPU part: init(file) / init(mesh) -> evaluate x3 -> bounds -> finalize against the oracle and the reference's golden phi."""
import os

import numpy as np
import pytest

from axom_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _stl_case(tmp_path, binary, jitter_scale=0.0, freq=6):
    x, y, z, conn = synth.icosphere(freq)
    jit = None
    if jitter_scale:
        jit = np.random.default_rng(3).uniform(-jitter_scale, jitter_scale, (len(conn), 3, 3))
    p = str(tmp_path / ("s_%d_%d.stl" % (binary, freq)))
    synth.write_stl(p, x, y, z, conn, binary=binary, jitter=jit)
    return p, (x, y, z, conn)


@pytest.mark.parametrize("binary", [False, True])
def test_stl_reader_matches_reference(oracle, have_ref, tmp_path, binary):
    from axom_b200 import quest_interface as Q
    p, (x, y, z, conn) = _stl_case(tmp_path, binary)
    gx, gy, gz, gc = Q.read_stl(p)
    assert gc.shape == (len(conn), 3) and np.array_equal(gc.reshape(-1), np.arange(3 * len(conn)))
    want = np.stack([x, y, z], 1)[conn].reshape(-1, 3)
    if binary:
        want = want.astype(np.float32).astype(np.float64)  # binary STL stores float32
    assert np.array_equal(np.stack([gx, gy, gz], 1), want)
    if have_ref:
        r = oracle.ref_stl_read_weld(p, 0.0)
        assert all(np.array_equal(a, b) for a, b in zip((gx, gy, gz, gc), r))
    with pytest.raises(Q.QuestError):
        Q.read_stl(str(tmp_path / "missing.stl"))


@pytest.mark.parametrize("eps", [1e-7, 1e-3, 0.1])
def test_weld_matches_reference(oracle, have_ref, tmp_path, eps):
    """STL soup with per-copy jitter of 2e-8: small eps merges most copies of a vertex, 0.1 also collapses whole
    triangles, which are dropped; numbering and kept coordinates are the reference's (first appearance)"""
    from axom_b200 import quest_interface as Q
    p, (x, y, z, conn) = _stl_case(tmp_path, False, jitter_scale=2e-8)
    sx, sy, sz, sc = Q.read_stl(p)
    wx, wy, wz, wc = Q.weldTriMeshVertices(sx, sy, sz, sc, eps)
    assert len(wx) < len(sx)  # copies were merged (lattice welding is not a full clustering: cell borders split some)
    if eps == 0.1:
        assert len(wc) < len(conn)
    g = np.load(os.path.join(G, "stl_weld.npz"))
    key = "eps_%g" % eps
    assert np.array_equal(wc, g[key + "_conn"]) and np.array_equal(np.stack([wx, wy, wz], 1), g[key + "_xyz"])
    if have_ref:
        r = oracle.ref_stl_read_weld(p, eps)
        assert all(np.array_equal(a, b) for a, b in zip((wx, wy, wz, wc), r))


def test_process_global_api_error_behaviour():
    """quest_signed_distance_interface.cpp:245-330: evaluate / bounds before init and setters with a bad value are
    SLIC_ERRORs; here they surface as QuestError through the error handler (C default: print + abort)"""
    from axom_b200 import quest_interface as Q
    Q.signed_distance_finalize()
    assert not Q.signed_distance_initialized()
    with pytest.raises(Q.QuestError):
        Q.signed_distance_evaluate(0.0, 0.0, 0.0)
    with pytest.raises(Q.QuestError):
        Q.signed_distance_get_mesh_bounds()
    with pytest.raises(Q.QuestError):
        Q.signed_distance_set_dimension(2)
    Q.signed_distance_set_dimension(3)
    Q.signed_distance_set_closed_surface(True)
    assert Q.signed_distance_init("/nonexistent/file.stl") == -1
    assert not Q.signed_distance_initialized()


@pytest.mark.gpu
@pytest.mark.parametrize("binary", [False, True])
def test_legacy_api_end_to_end(oracle, tmp_path, binary):
    from axom_b200 import quest_interface as Q
    p, _ = _stl_case(tmp_path, binary, freq=8)
    sx, sy, sz, sc = Q.read_stl(p)
    q = synth.uniform_grid_points(-1, 1, 12)
    want, _, _ = oracle.SignedDistance(sx, sy, sz, sc).compute(q)
    Q.signed_distance_finalize()
    Q.signed_distance_set_closed_surface(True)
    Q.signed_distance_set_compute_signs(True)
    Q.signed_distance_set_execution_space(Q.SignedDistExec.GPU)
    assert Q.signed_distance_init(p) == 0
    try:
        assert Q.signed_distance_initialized()
        with pytest.raises(Q.QuestError):  # setters after init (signed_distance.cpp:253-300)
            Q.signed_distance_set_closed_surface(False)
        with pytest.raises(Q.QuestError):  # double init (:165-166)
            Q.signed_distance_init(p)
        phi = Q.signed_distance_evaluate(q[:, 0], q[:, 1], q[:, 2])
        assert np.array_equal(phi, want)
        for i in (0, 77, 500, 1727):
            assert Q.signed_distance_evaluate(*q[i]) == want[i]
            v, cp, n = Q.signed_distance_evaluate(*q[i], with_closest_point=True)
            assert v == want[i] and abs(np.linalg.norm(q[i] - cp) - abs(v)) < 1e-12 and abs(np.linalg.norm(n) - 1) < 1e-12
        lo, hi = Q.signed_distance_get_mesh_bounds()
        assert np.array_equal(lo, [sx.min(), sy.min(), sz.min()]) and np.array_equal(hi, [sx.max(), sy.max(), sz.max()])
        import torch
        qd = torch.from_numpy(q).cuda()
        pd = Q.signed_distance_evaluate(qd[:, 0].contiguous(), qd[:, 1].contiguous(), qd[:, 2].contiguous())
        assert np.array_equal(pd.cpu().numpy(), want)
        if not binary:
            g = np.load(os.path.join(G, "stl_weld.npz"))
            assert np.array_equal(phi, g["legacy_phi_ascii_f8"])  # the real reference's legacy API on the same file
    finally:
        Q.signed_distance_finalize()
    assert not Q.signed_distance_initialized()


@pytest.mark.gpu
def test_legacy_api_init_from_mesh_open_surface_unsigned(oracle):
    from axom_b200 import quest_interface as Q
    x, y, z, conn = synth.icosphere(7)
    conn = conn[: len(conn) // 2]  # open surface
    q = synth.uniform_grid_points(-1, 1, 9)
    Q.signed_distance_finalize()
    for closed, sign in ((False, True), (True, False)):
        Q.signed_distance_set_closed_surface(closed)
        Q.signed_distance_set_compute_signs(sign)
        assert Q.signed_distance_init((x, y, z, conn)) == 0
        try:
            want, _, _ = oracle.SignedDistance(x, y, z, conn, 3, closed, sign).compute(q)
            assert np.array_equal(Q.signed_distance_evaluate(q[:, 0], q[:, 1], q[:, 2]), want)
        finally:
            Q.signed_distance_finalize()
    Q.signed_distance_set_closed_surface(True)
    Q.signed_distance_set_compute_signs(True)
