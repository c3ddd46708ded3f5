"""quest::MarchingCubes: the oracle restatement against the real reference and its golden
fixtures (CPU), and the CUDA path behind axb_mc_* against the oracle (GPU, bit-exact: ids, parents, domain ids AND the
interpolated coordinates -- linear_interp is restated operation by operation).

The result checks mirror the reference's own test driver quest/examples/quest_marching_cubes_example.cpp
(checkContourSurface :1152, checkContourCellLimits :1250, checkCellsContainingContour :1380)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from axom_b200 import synth
from axom_b200.marching_cubes import MarchingCubes, MarchingCubesDataParallelism, domain_views
import mc_cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mc_contours.npz")


def _same(a, b):
    return all(x.size == y.size and np.array_equal(x.reshape(-1), y.reshape(-1)) for x, y in zip(a, b))


def _golden(name):
    z = np.load(GOLD)
    return [z[name + "/" + k] for k in ("ids", "xyz", "par", "dom")], z[name + "/fcn_sha"]


def _fcn_sha(mesh):
    sha = hashlib.sha256()
    for d in mesh.values():
        sha.update(np.ascontiguousarray(d["fields"]["dist"]["values"]).tobytes())
    return np.frombuffer(sha.digest(), np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# the reference driver's three result checks, on (mesh, contour arrays)
# ---------------------------------------------------------------------------------------------------------------------
def _cell_corners(view, parent):
    """node coordinates (ncorner, D) and values of parent cell `parent` (flat index in the function's stride order)"""
    nd = len(view.cell_shape)
    order = sorted(range(nd), key=lambda d: -view.fcn_strides[d])  # slowest first (unique strides)
    idx, rest = [0] * nd, int(parent)
    strides, t = {}, 1
    for d in reversed(order):
        strides[d] = t
        t *= view.cell_shape[d]
    for d in order:
        idx[d], rest = divmod(rest, strides[d])
    pts, vals = [], []
    for corner in np.ndindex(*([2] * nd)):
        node = [idx[d] + corner[d] for d in range(nd)]
        xo = view.coords_offset + sum(n * s for n, s in zip(node, view.coords_strides))
        fo = view.fcn_offset + sum(n * s for n, s in zip(node, view.fcn_strides))
        pts.append([np.asarray(view.coords[d]).reshape(-1)[xo] for d in range(nd)])
        vals.append(np.asarray(view.fcn).reshape(-1)[fo])
    return np.array(pts), np.array(vals)


def check_contour(mesh, mask_field, mask_val, contour, ids, xyz, par, dom, center, warped=False, tol=None):
    views = domain_views(mesh, "mesh", "dist", mask_field)
    by_id = {v.domain_id: v for v in views}
    nd = len(views[0].cell_shape)
    n = par.size
    assert ids.size == n * nd and np.array_equal(ids.reshape(-1), np.arange(n * nd))  # facet f owns nodes nd*f .. nd*f+nd-1
    if n == 0:
        return
    xyz = xyz.reshape(n, nd, nd)
    # checkContourSurface: every contour node sits on the iso-surface of the analytic function, up to the linear
    # interpolation error of one cell (straight grids only: a warped grid's field is not radial in the output coordinates)
    if not warped:
        r = np.sqrt(((xyz - np.asarray(center)) ** 2).sum(-1))
        h = max(np.abs(np.diff(np.asarray(views[0].coords[0]).reshape(-1)[:2]))[0], 1e-12)
        assert np.all(np.abs(r - contour) <= (tol if tol is not None else 2.0 * h * h / max(contour, h) + 1e-12))
    # checkContourCellLimits: the nodes of a facet lie inside the bounding box of its parent cell
    # checkCellsContainingContour: every parent cell straddles the contour value
    for f in range(0, n, max(1, n // 400)):
        pts, vals = _cell_corners(by_id[int(dom[f])], par[f])
        lo, hi = pts.min(0) - 1e-12, pts.max(0) + 1e-12
        assert np.all(xyz[f] >= lo) and np.all(xyz[f] <= hi), f
        assert vals.min() < contour + 1e-8 and vals.max() >= contour - 1e-8, f


def check_parents_complete(mesh, mask_field, mask_val, contour, par, dom):
    """checkCellsContainingContour, the other direction: a cell whose corners straddle the value (some >= value, some <)
    and that the mask keeps is the parent of at least one facet"""
    views = domain_views(mesh, "mesh", "dist", mask_field)
    for v in views:
        nd = len(v.cell_shape)
        shape_n = [c + 1 for c in v.cell_shape]
        f = np.asarray(v.fcn).reshape(-1)
        node = np.indices(shape_n).reshape(nd, -1)
        vals = f[v.fcn_offset + sum(node[d] * v.fcn_strides[d] for d in range(nd))].reshape(shape_n)
        ge = vals >= contour
        sl = [slice(0, -1)] * nd
        cnt = np.zeros(v.cell_shape, int)
        for corner in np.ndindex(*([2] * nd)):
            cnt += ge[tuple(slice(c, c + s) for c, s in zip(corner, v.cell_shape))]
        cross = (cnt > 0) & (cnt < 2 ** nd)
        if v.mask is not None:
            cell = np.indices(v.cell_shape).reshape(nd, -1)
            m = np.asarray(v.mask).reshape(-1)[v.mask_offset + sum(cell[d] * v.mask_strides[d] for d in range(nd))].reshape(v.cell_shape)
            cross &= (m == mask_val)
        order = sorted(range(nd), key=lambda d: -v.fcn_strides[d])
        strides, t = [0] * nd, 1
        for d in reversed(order):
            strides[d] = t
            t *= v.cell_shape[d]
        flat = sum(np.indices(v.cell_shape)[d] * strides[d] for d in range(nd))
        want = np.sort(flat[cross])
        got = np.unique(par[dom == v.domain_id])
        assert np.array_equal(want, got)


# ---------------------------------------------------------------------------------------------------------------------
# CPU: the oracle against the real reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(mc_cases.CASES))
def test_oracle_matches_golden_reference_contours(oracle, name):
    mesh, mask_field, mask_val, contours = mc_cases.build(name)
    want, sha = _golden(name)
    assert np.array_equal(_fcn_sha(mesh), sha), "synthetic input differs from the one the fixture was generated from"
    got = mc_cases.oracle_contour(oracle, mesh, mask_field, mask_val, contours)
    assert _same(got, want)


@pytest.mark.parametrize("name", ["mc3d_single", "mc3d_multi_ghost_row_mask", "mc3d_row_domains", "mc2d_row_mask", "mc3d_mask0"])
def test_reference_driver_checks_hold_for_the_oracle(oracle, name):
    mesh, mask_field, mask_val, contours = mc_cases.build(name)
    kw = mc_cases.CASES[name][0]
    views = domain_views(mesh, "mesh", "dist", mask_field)
    ids, xyz, par, dom = oracle.mc_isocontour(views, contours[0], mask_val)
    nd = len(kw["cells"])
    check_contour(mesh, mask_field, mask_val, contours[0], ids, xyz, par, dom, kw.get("center", [0.0] * nd))
    check_parents_complete(mesh, mask_field, mask_val, contours[0], par, dom)


def test_oracle_equals_live_reference_on_random_meshes(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(17)
    for trial in range(12):
        nd = 2 if trial % 3 == 2 else 3
        cells = tuple(int(c) for c in rng.integers(3, 14, nd))
        domains = tuple(int(min(c // 2, d)) for c, d in zip(cells, rng.integers(1, 3, nd)))
        kw = dict(cells=cells, domains=domains, ghosts=int(rng.integers(0, 3)), order=["column", "row"][trial % 2],
                  mask_every=int(rng.choice([0, 3, 5])), warp=float(rng.choice([0.0, 0.03])), center=tuple(rng.uniform(-0.3, 0.3, nd)))
        mesh = synth.blueprint_structured_mesh(**kw)
        # perturb the field so that the contour is not a sphere (random cases of the table)
        for d in mesh.values():
            v = d["fields"]["dist"]["values"]
            v += rng.normal(0, 0.08, v.shape)
        mask_field = "mask" if kw["mask_every"] else ""
        contours = (0.5, 0.35)
        got = mc_cases.oracle_contour(oracle, mesh, mask_field, 1, contours)
        for dp in (MarchingCubesDataParallelism.hybridParallel, MarchingCubesDataParallelism.fullParallel):
            want = oracle.ref_mc_isocontour(mesh, "mesh", "dist", mask_field, 1, contours, int(dp))
            assert _same(got, want), (kw, dp)


def test_mc_tables_match_reference(oracle, have_ref):
    L = oracle.mc_lib()
    # anchors quoted from the published table (any marching-cubes text): one corner inside -> one triangle on its three edges
    assert [L.axo_mc_table(3, 1, e) for e in range(4)] == [0, 8, 3, -1]
    assert [L.axo_mc_table(3, 254, e) for e in range(4)] == [0, 3, 8, -1]
    assert L.axo_mc_num_contour_cells(3, 0) == 0 and L.axo_mc_num_contour_cells(3, 255) == 0
    assert [L.axo_mc_table(2, 5, e) for e in range(4)] == [2, 1, 0, 3] and L.axo_mc_num_contour_cells(2, 5) == 2
    assert max(L.axo_mc_num_contour_cells(3, c) for c in range(256)) == 5
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for dim, ncase, width in ((2, 16, 4), (3, 256, 16)):
        for c in range(ncase):
            assert [L.axo_mc_table(dim, c, e) for e in range(width)] == [oracle.ref_mc_table(dim, c, e) for e in range(width)], (dim, c)
            assert L.axo_mc_num_contour_cells(dim, c) == oracle.ref_mc_num_contour_cells(dim, c), (dim, c)


def test_mdmapping_restatement_matches_reference(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L, R = oracle.mc_lib(), oracle.lib("reference").lib
    R.axref_mc_mapping.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
    rng = np.random.default_rng(3)
    for trial in range(200):
        nd = 2 + trial % 2
        shape = rng.integers(1, 9, nd).astype(np.int64)
        perm = rng.permutation(nd)
        strides, t = np.zeros(nd, np.int64), int(rng.integers(1, 3))
        for d in perm:  # unique strides in a random direction order, with padding
            strides[d] = t
            t *= int(shape[d] + 1 + rng.integers(0, 3))
        flat = int(rng.integers(0, int(np.prod(shape))))
        s1, c1, i1 = np.zeros(3, np.int32), np.zeros(3, np.int64), np.zeros(3, np.int64)
        s2, c2, i2 = np.zeros(3, np.int32), np.zeros(3, np.int64), np.zeros(3, np.int64)
        L.axo_mc_mapping(nd, strides.ctypes.data, shape.ctypes.data, s1.ctypes.data, c1.ctypes.data)
        L.axo_mc_to_multi_index(nd, strides.ctypes.data, shape.ctypes.data, flat, i1.ctypes.data)
        R.axref_mc_mapping(nd, strides.ctypes.data, shape.ctypes.data, s2.ctypes.data, c2.ctypes.data, flat, i2.ctypes.data)
        assert np.array_equal(s1, s2) and np.array_equal(c1, c2) and np.array_equal(i1, i2), (shape, strides)


def test_domain_views_follow_meshviewutil_defaults():
    mesh = synth.blueprint_structured_mesh(cells=(4, 3, 2))
    v = domain_views(mesh, "mesh", "dist")[0]
    assert v.cell_shape == [4, 3, 2] and v.coords_strides == [1, 5, 20] and v.fcn_strides == [1, 5, 20]
    assert v.coords_offset == 0 and v.fcn_offset == 0 and v.mask is None and v.domain_id == 0
    mesh = synth.blueprint_structured_mesh(cells=(4, 3, 2), ghosts=1, order="row", mask_every=2, domain_id_base=9, domains=(2, 1, 1))
    v = domain_views(mesh, "mesh", "dist", "mask")
    assert [x.domain_id for x in v] == [9, 10] and v[0].cell_shape == [2, 3, 2]
    assert v[0].coords_strides == [1, 5, 30] and v[0].coords_offset == 1 + 5 + 30      # padded node shape (5, 6, 5), column-major
    assert v[0].fcn_strides == [30, 5, 1] and v[0].fcn_offset == 30 + 5 + 1            # row-major
    assert v[0].mask_strides == [20, 4, 1] and v[0].mask_offset == 20 + 4 + 1          # padded cell shape (4, 5, 4)


def test_single_domain_tree_is_rejected_like_the_reference():
    mesh = synth.blueprint_structured_mesh(cells=(3, 3, 3))
    with pytest.raises(ValueError, match="multidomain"):
        domain_views(mesh["domain_000000"], "mesh", "dist")


def test_mc_fails_loudly_without_a_device():
    """no CPU fallback: on a box without a GPU the C ABI reports AXB_ERR_NO_DEVICE"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from axom_b200 import _lib
    mc = MarchingCubes()
    mc.setMesh(synth.blueprint_structured_mesh(cells=(3, 3, 3)), "mesh")
    mc.setFunctionField("dist")
    with pytest.raises(_lib.AxbError) as e:
        mc.computeIsocontour(0.5)
    assert e.value.status == _lib.AXB_ERR_NO_DEVICE


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the oracle / golden fixtures
# ---------------------------------------------------------------------------------------------------------------------
def _gpu_contour(mesh, mask_field, mask_val, contours, device_out=False):
    mc = MarchingCubes(device=0)
    mc.setMesh(mesh, "mesh", mask_field)
    mc.setFunctionField("dist")
    mc.setMaskValue(mask_val)
    for c in contours:
        mc.computeIsocontour(c)
    out = (mc.getContourFacetCorners(device_out), mc.getContourNodeCoords(device_out), mc.getContourFacetParents(device_out),
           mc.getContourFacetDomainIds(device_out))
    assert mc.getContourCellCount() == out[2].shape[0] and mc.getContourNodeCount() in (0, out[1].shape[0])
    return [o.cpu().numpy() if device_out else o for o in out], mc


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mc_cases.CASES))
def test_gpu_matches_golden_and_oracle_host_inputs(oracle, name):
    mesh, mask_field, mask_val, contours = mc_cases.build(name)
    got, _ = _gpu_contour(mesh, mask_field, mask_val, contours)
    assert _same(got, _golden(name)[0])
    assert _same(got, mc_cases.oracle_contour(oracle, mesh, mask_field, mask_val, contours))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mc_cases.CASES))
def test_gpu_plain_mark_kernel_matches_too(oracle, name, monkeypatch):
    """AXB_MC_MARK_PLAIN=1 selects the one-cell-per-thread mark kernel (the A/B reference of the row kernel)"""
    monkeypatch.setenv("AXB_MC_MARK_PLAIN", "1")
    mesh, mask_field, mask_val, contours = mc_cases.build(name)
    got, _ = _gpu_contour(mesh, mask_field, mask_val, contours)
    assert _same(got, _golden(name)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mc3d_multi_ghost_row_mask", "mc3d_warp_ghost", "mc2d_multi_ghost_warp", "mc3d_empty"])
def test_gpu_device_resident_inputs_and_outputs(oracle, name):
    mesh, mask_field, mask_val, contours = mc_cases.build(name)
    got, _ = _gpu_contour(synth.blueprint_to_device(mesh), mask_field, mask_val, contours, device_out=True)
    assert _same(got, mc_cases.oracle_contour(oracle, mesh, mask_field, mask_val, contours))


@pytest.mark.gpu
def test_gpu_random_fields_bit_exact(oracle):
    """noisy fields reach every row of the case table; several tiles per domain, ragged last tile"""
    rng = np.random.default_rng(23)
    seen = set()
    for trial, (cells, domains, order, ghosts) in enumerate([((37, 29, 23), (2, 1, 2), "column", 1), ((64, 33, 17), (1, 1, 1), "row", 0),
                                                              ((130, 97), (2, 3), "row", 2), ((1024, 3), (1, 1), "column", 0),
                                                              ((3, 2, 700), (1, 1, 3), "row", 1)]):
        mesh = synth.blueprint_structured_mesh(cells=cells, domains=domains, order=order, ghosts=ghosts, mask_every=5)
        for d in mesh.values():
            v = d["fields"]["dist"]["values"]
            v += rng.normal(0, 0.15, v.shape)
        want = mc_cases.oracle_contour(oracle, mesh, "mask", 1, (0.5, 0.8))
        got, _ = _gpu_contour(mesh, "mask", 1, (0.5, 0.8))
        assert want[2].size > 100 and _same(got, want), cells
    # the 3-D noisy cases above must have exercised (nearly) the whole table
    mesh = synth.blueprint_structured_mesh(cells=(40, 40, 40))
    v = mesh["domain_000000"]["fields"]["dist"]["values"]
    v[:] = rng.random(v.shape)
    want = mc_cases.oracle_contour(oracle, mesh, "", 1, (0.5,))
    got, _ = _gpu_contour(mesh, "", 1, (0.5,))
    assert _same(got, want)
    f = v.reshape(41, 41, 41, order="F") >= 0.5
    corners = [(1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 0), (1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 0, 1)]
    case = sum(f[a:a + 40, b:b + 40, c:c + 40].astype(int) << n for n, (a, b, c) in enumerate(corners))
    seen = set(np.unique(case))
    assert len(seen) == 256


@pytest.mark.gpu
def test_gpu_accumulate_clear_relinquish(oracle):
    mesh, mask_field, mask_val, _ = mc_cases.build("mc3d_row_domains")
    views = domain_views(mesh, "mesh", "dist", mask_field)
    mc = MarchingCubes(device=0)
    mc.setMesh(mesh, "mesh", mask_field)
    mc.setFunctionField("dist")
    assert mc.getContourCellCount() == 0
    mc.computeIsocontour(0.4)
    n1 = mc.getContourCellCount()
    mc.computeIsocontour(0.6)
    n2 = mc.getContourCellCount()
    a, b = oracle.mc_isocontour(views, 0.4), oracle.mc_isocontour(views, 0.6, first_facet=n1)
    assert n1 == a[2].size and n2 == n1 + b[2].size
    out = mc.relinquishContourData()
    assert _same(out, [np.concatenate([x, y]) for x, y in zip(a, b)])
    assert mc.getContourCellCount() == 0 and mc.getContourNodeCount() == 0
    mc.computeIsocontour(0.6)  # node ids restart at 0 after the contour was given away
    assert _same((mc.getContourFacetCorners(), mc.getContourNodeCoords(), mc.getContourFacetParents(), mc.getContourFacetDomainIds()),
                 oracle.mc_isocontour(views, 0.6))
    mc.clearOutput()
    assert mc.getContourCellCount() == 0
    m = mc.populateContourMesh("cellIds", "domainIds")
    assert m["cells"].shape == (0, 3) and set(m["fields"]) == {"cellIds", "domainIds"}


@pytest.mark.gpu
def test_gpu_field_updated_in_place_between_contours(oracle):
    """the reference re-reads its live field / mask views on every computeIsocontour (m_fcnView, m_maskView): an
    in-place update of the field followed by clearOutput() + computeIsocontour() must contour the NEW field, for host
    arrays (re-staged per contour) and for device tensors (read in place, after torch's stream has drained)"""
    import copy
    import torch
    mesh, mask_field, mask_val, _ = mc_cases.build("mc3d_row_domains")
    for on_device in (False, True):
        ref = copy.deepcopy(mesh)  # host twin for the oracle
        m = copy.deepcopy(mesh)
        if on_device:
            m = synth.blueprint_to_device(m)
        mc = MarchingCubes(device=0)
        mc.setMesh(m, "mesh", mask_field)
        mc.setFunctionField("dist")
        mc.setMaskValue(mask_val)
        mc.computeIsocontour(0.5)
        first = (mc.getContourFacetCorners().copy(), mc.getContourNodeCoords().copy())
        for name in m:
            v = m[name]["fields"]["dist"]["values"]
            if on_device:
                v.mul_(0.5).add_(0.1)  # queued on torch's stream, not synchronised by the caller
            else:
                v *= 0.5
                v += 0.1
            r = ref[name]["fields"]["dist"]["values"]
            r *= 0.5
            r += 0.1
        mc.clearOutput()
        mc.computeIsocontour(0.5)
        want = oracle.mc_isocontour(domain_views(ref, "mesh", "dist", mask_field), 0.5, mask_val=mask_val)
        got = (mc.getContourFacetCorners(), mc.getContourNodeCoords(), mc.getContourFacetParents(), mc.getContourFacetDomainIds())
        assert _same(got, want), on_device
        assert got[1].shape != first[1].shape or not np.array_equal(got[1], first[1])  # the contour did move


@pytest.mark.gpu
def test_gpu_rejects_non_unique_strides():
    from axom_b200 import _lib
    mesh = synth.blueprint_structured_mesh(cells=(4, 4, 4))
    mesh["domain_000000"]["fields"]["dist"]["strides"] = np.array([1, 5, 5], np.int32)
    mc = MarchingCubes(device=0)
    mc.setMesh(mesh, "mesh")
    mc.setFunctionField("dist")
    with pytest.raises(_lib.AxbError) as e:
        mc.computeIsocontour(0.5)
    assert e.value.status == _lib.AXB_ERR_BAD_ARG


@pytest.mark.gpu
def test_gpu_full_size_distance_field_properties(oracle):
    """the C2 consumer: iso-contour of a 256^3 nodal distance field (255^3 = 16.6 M cells) -- equal to the oracle, and the
    reference driver's checks hold"""
    mesh = synth.blueprint_structured_mesh(cells=(255, 255, 255), center=(0.01, -0.02, 0.03))
    want = mc_cases.oracle_contour(oracle, mesh, "", 1, (0.5,))
    got, mc = _gpu_contour(synth.blueprint_to_device(mesh), "", 1, (0.5,), device_out=True)
    assert want[2].size > 100000 and _same(got, want)
    check_contour(mesh, "", 1, 0.5, *got, center=(0.01, -0.02, 0.03))
    assert np.all(np.diff(got[2]) >= 0)  # facets come sorted by parent cell


@pytest.mark.gpu
def test_gpu_signed_distance_field_to_contour_chain(oracle):
    """quest::SignedDistance on a grid -> quest::MarchingCubes at phi = 0.05: both stages on the device, equal to the oracle chain"""
    import torch
    from axom_b200 import SignedDistance
    x, y, z, conn = synth.icosphere(12)
    n = 40
    mesh = synth.blueprint_structured_mesh(cells=(n, n, n))
    cs = mesh["domain_000000"]["coordsets"]["coords"]["values"]
    q = np.stack([cs["x"], cs["y"], cs["z"]], 1)
    rphi, _, _ = oracle.SignedDistance(x, y, z, conn).compute(q)
    sd = SignedDistance(x, y, z, conn, device=0)
    phi, _, _ = sd.computeDistances(torch.from_numpy(q).cuda())
    assert np.array_equal(rphi, phi.cpu().numpy())
    dmesh = synth.blueprint_to_device(mesh)
    dmesh["domain_000000"]["fields"]["dist"]["values"] = phi
    mesh["domain_000000"]["fields"]["dist"]["values"] = rphi
    got, _ = _gpu_contour(dmesh, "", 1, (0.05,), device_out=True)
    assert got[2].size > 1000 and _same(got, mc_cases.oracle_contour(oracle, mesh, "", 1, (0.05,)))


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: a field sharded by z-slabs, one halo plane exchanged per rank (the flow of bench.py at N > 1), under gloo on CPU
# with the oracle as the per-rank contouring step: the slabs' facets in rank order == the single-domain contour
# ---------------------------------------------------------------------------------------------------------------------
def _slab_field(n):
    ax = np.linspace(-1.0, 1.0, n)
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
    rng = np.random.default_rng(77)
    return ax, (np.sqrt((xx - 0.05) ** 2 + (yy + 0.02) ** 2 + zz ** 2) + rng.normal(0, 0.03, xx.shape)).reshape(-1)


def _slab_worker(rank, world, port, n, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from axom_b200.marching_cubes import slab_domain
        from oracle import oracle as O
        ax, field = _slab_field(n)
        plane = n * n
        p0, p1 = (n * rank) // world, (n * (rank + 1)) // world  # ragged slabs when world does not divide n
        mine = torch.from_numpy(field[p0 * plane:p1 * plane].copy())
        # the halo exchange: every rank contributes its first node plane
        firsts = [torch.empty(plane, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(firsts, mine[:plane].contiguous())
        dom, ncell_planes = slab_domain(mine.numpy(), firsts[rank + 1].numpy() if rank < world - 1 else None, (ax, ax, ax), p0, rank)
        ids, xyz, par, did = O.mc_isocontour(domain_views(dom, "mesh", "phi"), 0.5)
        gathered = [None] * world
        dist.all_gather_object(gathered, (p0, ncell_planes, xyz, par, did))
        if rank == 0:
            full, _ = slab_domain(field, None, (ax, ax, ax), 0, 0)
            _, fxyz, fpar, _ = O.mc_isocontour(domain_views(full, "mesh", "phi"), 0.5)
            cpp = (n - 1) ** 2
            cat_xyz = np.concatenate([g[2] for g in gathered])
            cat_par = np.concatenate([g[3] + g[0] * cpp for g in gathered])
            ok = (np.array_equal(cat_xyz, fxyz) and np.array_equal(cat_par, fpar) and sum(g[1] for g in gathered) == n - 1
                  and all((g[4] == r).all() for r, g in enumerate(gathered)))
            out.put((bool(ok), int(fpar.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 20), (3, 23)])
def test_sharded_slabs_with_halo_equal_single_domain_under_gloo(oracle, world, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + 11 * world
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, nfacets = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok and nfacets > 500
