"""Marching-cubes parity cases shared by tests/golden/make_golden.py, the CPU oracle tests and the GPU tests.
Each case: kwargs of synth.blueprint_structured_mesh, the mask field ("" = none), the mask value and the contour values
computed in turn into one accumulating contour mesh (quest/MarchingCubes.hpp:160-164)."""

CASES = {
    # name: (mesh kwargs, mask field, mask value, contour values)
    "mc3d_single": (dict(cells=(12, 10, 8)), "", 1, (0.55,)),
    "mc3d_multi_ghost_row_mask": (dict(cells=(12, 10, 8), domains=(2, 1, 2), ghosts=1, order="row", mask_every=7, domain_id_base=5),
                                  "mask", 1, (0.55, 0.3)),
    "mc3d_row_domains": (dict(cells=(33, 17, 9), domains=(2, 2, 1), order="row"), "", 1, (0.4,)),
    "mc3d_warp_ghost": (dict(cells=(16, 16, 16), domains=(1, 2, 1), ghosts=2, warp=0.04, center=(0.1, -0.05, 0.02)), "", 1, (0.5, 0.7)),
    "mc3d_mask0": (dict(cells=(14, 9, 11), mask_every=2), "mask", 0, (0.45,)),
    # contour value exactly on lattice nodes: the isNearlyEqual branches of linear_interp (centre on a node, axis distances k*h)
    "mc3d_on_nodes": (dict(cells=(8, 8, 8), lo=-1.0, hi=1.0), "", 1, (0.5, 0.25)),
    "mc2d_multi_ghost_warp": (dict(cells=(20, 15), domains=(2, 2), ghosts=2, warp=0.05), "", 1, (0.55, 0.3)),
    "mc2d_row_mask": (dict(cells=(40, 31), order="row", mask_every=3), "mask", 1, (0.6,)),
    "mc2d_on_nodes": (dict(cells=(8, 8)), "", 1, (0.5,)),
    # no crossing at all / everything above / a one-cell domain
    "mc3d_empty": (dict(cells=(5, 4, 3)), "", 1, (9.0, -1.0)),
    "mc3d_one_cell": (dict(cells=(1, 1, 1), lo=0.0, hi=1.0, center=(0.0, 0.0, 0.0)), "", 1, (0.5,)),
}


def build(name):
    from axom_b200 import synth
    kw, mask_field, mask_val, contours = CASES[name]
    return synth.blueprint_structured_mesh(**kw), mask_field, mask_val, contours


def oracle_contour(O, mesh, mask_field, mask_val, contours):
    """the restatement run as the reference accumulates contours: -> the four concatenated arrays"""
    import numpy as np
    from axom_b200.marching_cubes import domain_views
    views = domain_views(mesh, "mesh", "dist", mask_field)
    parts, first = [], 0
    for c in contours:
        r = O.mc_isocontour(views, c, mask_val, first_facet=first)
        first += r[0].shape[0]
        parts.append(r)
    return [np.concatenate([p[k] for p in parts]) for k in range(4)]
