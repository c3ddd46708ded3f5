"""Synthetic workloads for the BASELINE.json configs.

Host-side numpy only; nothing here is on the hot path.  The same arrays are handed to the
CUDA path and to the CPU oracle, so the generator's RNG only has to be seeded, not bit-stable
across numpy versions.
"""
import numpy as np


def triangle_aabbs(n, seed=12345, shift=(0.0, 0.0, 0.0), ndims=3):
    """C1/C3: n triangle AABBs; centres uniform in [0,1)^D, vertices = centre + h*U(-1/2,1/2)^D,
    h = n^(-1/D).  Returns (n, 2*D) float64 [min..., max...] (the primal::BoundingBox layout)."""
    rng = np.random.default_rng(seed)
    h = float(n) ** (-1.0 / ndims)
    c = rng.random((n, 1, ndims))
    v = c + h * (rng.random((n, 3, ndims)) - 0.5)
    v = v + np.asarray(shift, np.float64)[:ndims]
    return np.ascontiguousarray(np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1))


def random_points(q, seed=12345 + 1, lo=0.0, hi=1.0, ndims=3):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(lo + (hi - lo) * rng.random((q, ndims)))


def random_rays(q, seed=777, lo=-1.0, hi=1.0, ndims=3):
    """C4: origins uniform in [lo,hi]^D, directions uniform on the sphere (normalised Gaussian)."""
    rng = np.random.default_rng(seed)
    o = lo + (hi - lo) * rng.random((q, ndims))
    d = rng.standard_normal((q, ndims))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


_ICO_FACES = None


def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    return v, f


def icosphere(freq, radius=0.5, center=(0.0, 0.0, 0.0)):
    """Geodesic icosphere of frequency `freq`: 20*freq^2 triangles, 10*freq^2+2 welded vertices,
    consistent outward (CCW from outside) winding.  C2 uses freq=316 (1 997 120 triangles),
    C4 freq=1000 (20 000 000 triangles).
    Returns x, y, z (float64) and conn (ntri,3) int32."""
    n = int(freq)
    V, F = _icosahedron()
    # global vertex ids: 12 corners | 30 edges x (n-1) | 20 faces x (n-1)(n-2)/2
    edges = {}
    for f in F:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            key = (min(a, b), max(a, b))
            if key not in edges:
                edges[key] = len(edges)
    n_edge_pts = n - 1
    n_face_pts = (n - 1) * (n - 2) // 2
    nverts = 12 + 30 * n_edge_pts + 20 * n_face_pts
    P = np.empty((nverts, 3), np.float64)
    P[:12] = V
    ts = np.arange(1, n, dtype=np.float64)[:, None]
    for (a, b), e in edges.items():
        base = 12 + e * n_edge_pts
        P[base:base + n_edge_pts] = ((n - ts) * V[a] + ts * V[b]) / n

    # per-face barycentric lattice: point (i,j) with i+j<=n is ((n-i-j)*A + i*B + j*C)/n
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    lattice_valid = (ii + jj) <= n
    # interior numbering (i>=1, j>=1, i+j<=n-1), row-major in i then j
    interior = (ii >= 1) & (jj >= 1) & ((ii + jj) <= n - 1)
    interior_id = -np.ones((n + 1, n + 1), np.int64)
    interior_id[interior] = np.arange(n_face_pts)

    def edge_point(a, b, t):
        """id of the point at parameter t (from a towards b) on edge (a,b), t in 0..n (vectorised)"""
        key = (min(a, b), max(a, b))
        e = edges[key]
        tt = t if a < b else n - t
        out = 12 + e * n_edge_pts + (tt - 1)
        out = np.where(tt == 0, key[0], out)
        out = np.where(tt == n, key[1], out)
        return out

    conn_all = []
    for fi, (A, B, C) in enumerate(F):
        gid = -np.ones((n + 1, n + 1), np.int64)
        fbase = 12 + 30 * n_edge_pts + fi * n_face_pts
        gid[interior] = fbase + interior_id[interior]
        if n_face_pts:
            i_in, j_in = ii[interior].astype(np.float64)[:, None], jj[interior].astype(np.float64)[:, None]
            P[fbase:fbase + n_face_pts] = ((n - i_in - j_in) * V[A] + i_in * V[B] + j_in * V[C]) / n
        t = np.arange(n + 1)
        gid[t, 0] = edge_point(A, B, t)          # j == 0: from A (i=0) to B (i=n)
        gid[0, t] = edge_point(A, C, t)          # i == 0: from A (j=0) to C (j=n)
        gid[t, n - t] = edge_point(C, B, t)      # i + j == n: from C (i=0) to B (i=n)
        # upward triangles (i,j),(i+1,j),(i,j+1) for i+j<=n-1; downward (i+1,j),(i+1,j+1),(i,j+1) for i+j<=n-2
        iu, ju = np.nonzero((ii + jj) <= n - 1)
        up = np.stack([gid[iu, ju], gid[iu + 1, ju], gid[iu, ju + 1]], axis=1)
        idn, jdn = np.nonzero((ii + jj) <= n - 2)
        dn = np.stack([gid[idn + 1, jdn], gid[idn + 1, jdn + 1], gid[idn, jdn + 1]], axis=1)
        conn_all.append(up)
        conn_all.append(dn)
        assert lattice_valid[iu, ju].all()
    conn = np.concatenate(conn_all, axis=0)
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    P = P * radius + np.asarray(center, np.float64)
    # make sure the winding is outward
    a, b, c = P[conn[:, 0]], P[conn[:, 1]], P[conn[:, 2]]
    nrm = np.cross(b - a, c - a)
    flip = (nrm * (a + b + c - 3 * np.asarray(center))).sum(axis=1) < 0
    conn[flip] = conn[flip][:, [0, 2, 1]]
    return (np.ascontiguousarray(P[:, 0]), np.ascontiguousarray(P[:, 1]), np.ascontiguousarray(P[:, 2]),
            np.ascontiguousarray(conn.astype(np.int32)))


def latlong_sphere(radius=0.5, theta_res=25, phi_res=25, center=(0.0, 0.0, 0.0)):
    """The reference test's sphere mesh (quest/tests/quest_test_utilities.hpp:51-151), restated:
    same node order, same connectivity (including its seam/pole quirks), so the golden norms of
    quest/tests/quest_signed_distance.cpp:88-91 apply."""
    deg = np.pi / 180.0
    cx, cy, cz = center
    xs, ys, zs = [cx, cx], [cy, cy], [cz + radius, cz - radius]
    dphi = (180 * deg) / float(phi_res - 1)
    dtheta = (360 * deg) / float(theta_res - 1)
    for i in range(theta_res):
        theta = i * dtheta
        for j in range(phi_res - 2):
            phi = j * dphi
            r = radius * np.sin(phi)
            xs.append(r * np.cos(theta) + cx)
            ys.append(r * np.sin(theta) + cy)
            zs.append(radius * np.cos(phi) + cz)
    pr = phi_res - 2
    stride = pr * theta_res
    cells = []
    for i in range(theta_res):
        cells.append([0, (pr * (i + 1) % stride) + 2, pr * i + 2])
    off = phi_res - 1
    for i in range(theta_res):
        cells.append([1, (pr * (i + 1) % stride) + off, pr * i + off])
    for i in range(theta_res):
        for j in range(phi_res - 3):
            c0 = pr * i + j + 2
            c2 = ((pr * (i + 1) + j) % stride) + 3
            cells.append([c0, c0 + 1, c2])
            cells.append([c0, c2, c2 - 1])
    return (np.array(xs, np.float64), np.array(ys, np.float64), np.array(zs, np.float64), np.array(cells, np.int32))


def uniform_grid_points(lo, hi, n):
    """n^3 (or nx,ny,nz) lattice nodes spanning [lo,hi], x fastest (mint::UniformMesh node order)."""
    lo = np.broadcast_to(np.asarray(lo, np.float64), (3,))
    hi = np.broadcast_to(np.asarray(hi, np.float64), (3,))
    nn = np.broadcast_to(np.asarray(n, np.int64), (3,))
    axes = [lo[d] + np.arange(nn[d], dtype=np.float64) * ((hi[d] - lo[d]) / (nn[d] - 1)) for d in range(3)]
    zz, yy, xx = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    return np.ascontiguousarray(np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1))


def mesh_cell_boxes(x, y, z, conn):
    """per-cell AABB over the cell's nodes (quest/SignedDistance.hpp:608-633), (ncells, 6)."""
    P = np.stack([x, y, z], axis=1)[conn]
    return np.ascontiguousarray(np.concatenate([P.min(axis=1), P.max(axis=1)], axis=1))


def write_stl(path, x, y, z, conn, binary=False, jitter=None):
    """write a triangle mesh as an STL file (ASCII with 17 significant digits, or binary float32).  `jitter`
    (n, 3, 3) is added to the per-triangle vertex copies (to exercise vertex welding)."""
    P = np.stack([x, y, z], 1)[np.asarray(conn)]
    if jitter is not None:
        P = P + jitter
    n = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
    if binary:
        rec = np.zeros(len(P), dtype=[("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")])
        rec["n"] = n
        rec["v"] = P.reshape(-1, 9)
        with open(path, "wb") as f:
            f.write(b"axom_b200 synthetic surface".ljust(80, b" "))
            f.write(np.int32(len(P)).tobytes())
            f.write(rec.tobytes())
        return
    with open(path, "w") as f:
        f.write("solid synth\n")
        for t, nn in zip(P, n):
            f.write(" facet normal %.9g %.9g %.9g\n  outer loop\n" % tuple(nn))
            for v in t:
                f.write("   vertex %.17g %.17g %.17g\n" % tuple(v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid synth\n")


def blueprint_structured_mesh(cells, lo=-1.0, hi=1.0, domains=(1, 1, 1), fcn="dist", center=None, ghosts=0, order="column",
                              mask_every=0, warp=0.0, domain_id_base=None):
    """A multi-domain Blueprint-shaped structured mesh (dict tree, numpy leaves) like the one
    quest/examples/quest_marching_cubes_example.cpp builds: `cells` = total cells per direction (2 or 3 entries) split into
    `domains` blocks per direction; a nodal field `fcn` = distance to `center` (the example's "round" contour), so its
    iso-contours are circles / spheres.  ghosts > 0 pads coordinates and fields with that many ghost layers on every side
    (elements/dims/offsets+strides and fields/*/offsets+strides are then present); order = "column" (direction 0 fastest,
    Conduit's default) or "row" (last direction fastest) applies to the fields; mask_every = m > 0 adds an int32 cell field
    "mask" that is 1 except on every m-th cell (0 there); warp > 0 bends the grid (curvilinear coordinates)."""
    nd = len(cells)
    cells = [int(c) for c in cells]
    nper = [int(d) for d in domains[:nd]]
    center = [0.0] * nd if center is None else list(center)
    lo = [lo] * nd if np.isscalar(lo) else list(lo)
    hi = [hi] * nd if np.isscalar(hi) else list(hi)
    # global node lattice (exactly as one global linspace, so domain boundaries share bit-identical coordinates)
    axes = [np.linspace(lo[d], hi[d], cells[d] + 1) for d in range(nd)]
    splits = [np.linspace(0, cells[d], nper[d] + 1).astype(int) for d in range(nd)]
    mesh = {}
    blocks = [(a, b, c) for c in range(nper[2] if nd == 3 else 1) for b in range(nper[1]) for a in range(nper[0])]
    for pos, blk in enumerate(blocks):
        c0 = [int(splits[d][blk[d]]) for d in range(nd)]
        c1 = [int(splits[d][blk[d] + 1]) for d in range(nd)]
        shape = [c1[d] - c0[d] for d in range(nd)]
        g = int(ghosts)
        # padded node index ranges (ghost nodes extrapolate the lattice)
        idx = [np.arange(c0[d] - g, c1[d] + 1 + g) for d in range(nd)]
        ax = []
        for d in range(nd):
            h = (hi[d] - lo[d]) / cells[d]
            a = np.where((idx[d] >= 0) & (idx[d] <= cells[d]), axes[d][np.clip(idx[d], 0, cells[d])], lo[d] + idx[d] * h)
            ax.append(a)
        grids = np.meshgrid(*ax, indexing="ij")  # arrays indexed [i, j(, k)]
        if warp:
            w = [grids[d] + warp * np.sin(np.pi * grids[(d + 1) % nd]) for d in range(nd)]
            grids = w
        dist = np.sqrt(sum((grids[d] - center[d]) ** 2 for d in range(nd)))
        pshape = [len(i) for i in idx]

        def flat(a, how):
            # "column": direction 0 fastest = Fortran order of the [i,j,k] array
            return np.ascontiguousarray(a.ravel(order="F" if how == "column" else "C"))

        def strides_of(shp, how):
            s, t = [0] * nd, 1
            for d in (range(nd) if how == "column" else range(nd - 1, -1, -1)):
                s[d] = t
                t *= shp[d]
            return s
        dims = {k: shape[d] for d, k in enumerate("ijk"[:nd])}
        if g:
            dims["offsets"] = np.array([g] * nd, np.int32)
            dims["strides"] = np.array(strides_of(pshape, "column"), np.int32)
        dom = {
            "coordsets": {"coords": {"type": "explicit", "values": {k: flat(grids[d], "column") for d, k in enumerate("xyz"[:nd])}}},
            "topologies": {"mesh": {"type": "structured", "coordset": "coords", "elements": {"dims": dims}}},
            "fields": {fcn: {"association": "vertex", "topology": "mesh", "values": flat(dist, order)}},
        }
        if g or order != "column":
            dom["fields"][fcn]["strides"] = np.array(strides_of(pshape, order), np.int32)
            dom["fields"][fcn]["offsets"] = np.array([g] * nd, np.int32)
        if mask_every:
            cshape = [s + 2 * g for s in shape]
            m = np.ones(cshape, np.int32)
            m.ravel()[::mask_every] = 0
            dom["fields"]["mask"] = {"association": "element", "topology": "mesh", "values": flat(m, order)}
            if g or order != "column":
                dom["fields"]["mask"]["strides"] = np.array(strides_of(cshape, order), np.int32)
                dom["fields"]["mask"]["offsets"] = np.array([g] * nd, np.int32)
        if domain_id_base is not None:
            dom["state"] = {"domain_id": domain_id_base + pos}
        mesh["domain_%06d" % pos] = dom
    return mesh


def blueprint_to_device(mesh, device=0):
    """the same tree with every float64 / int32 data array moved to cuda:<device> (index vectors stay on the host)"""
    import torch

    def walk(n, key=None):
        if isinstance(n, dict):
            return {k: walk(v, k) for k, v in n.items()}
        if isinstance(n, np.ndarray) and key not in ("offsets", "strides"):
            return torch.from_numpy(n).to("cuda:%d" % device)
        return n
    return walk(mesh)
