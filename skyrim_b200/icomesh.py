"""Icosahedral multimesh and the grid<->mesh edge sets of the GraphCast step (host side, built once at init).

What the reference reaches through ``graphcast.load_time_loop_operational`` (/root/reference/skyrim/core/models/
graphcast.py:51-54) builds these tables inside deepmind's un-vendored JAX package; the published construction
 is restated here with numpy / scipy:

* nodes: an icosahedron refined ``levels`` times (every face split in four, new vertices pushed to the unit sphere):
  level 6 has 40,962 nodes; ``build_graph`` renumbers them north-to-south / west-to-east (gather locality on the GPU);
* mesh edges: the union over ALL refinement levels of the face edges, both directions (level 6: 327,660);
* grid2mesh: every grid point within 0.6 x (longest edge of the finest mesh) of a mesh node sends to it;
* mesh2grid: the three vertices of the finest-mesh triangle containing a grid point send to it;
* edge features: (|d|, d) / max|d| with d = sender - receiver expressed in the receiver's local frame (the rotation
  that takes the receiver to latitude 0, longitude 0); node features: (cos lat, sin lon, cos lon).

Orderings are chosen for the CUDA engine: mesh and grid2mesh edges are sorted by receiver (CSR segments for the
deterministic aggregation); mesh2grid edges are k-major ``[3][n_grid]`` (edge k of grid point g is row k*n_grid + g),
so the aggregation over a grid point's three edges becomes a K-concatenation in the grid-node GEMM.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def icosahedron():
    """12 unit vertices, 20 faces with outward (counter-clockwise seen from outside) orientation."""
    phi = (1.0 + np.sqrt(5.0)) / 2.0
    v = []
    for a in (-1.0, 1.0):
        for b in (-phi, phi):
            v += [(0.0, a, b), (a, b, 0.0), (b, 0.0, a)]
    v = np.array(v, dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    # faces = all vertex triples at mutual distance = edge length
    d = np.linalg.norm(v[:, None] - v[None], axis=-1)
    el = np.min(d[d > 1e-9])
    adj = np.abs(d - el) < 1e-9
    faces = []
    for i in range(12):
        for j in range(i + 1, 12):
            for k in range(j + 1, 12):
                if adj[i, j] and adj[j, k] and adj[i, k]:
                    a, b, c = v[i], v[j], v[k]
                    faces.append((i, j, k) if np.dot(np.cross(b - a, c - a), a + b + c) > 0 else (i, k, j))
    faces = np.array(faces, dtype=np.int64)
    assert faces.shape == (20, 3)
    return v, faces


def refine(vertices: np.ndarray, faces: np.ndarray):
    """Split every face in four; midpoints are shared between neighbouring faces and normalised to the sphere."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    key = np.sort(e, axis=1)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    mid = vertices[uniq[:, 0]] + vertices[uniq[:, 1]]
    mid /= np.linalg.norm(mid, axis=1, keepdims=True)
    n0, nf = len(vertices), len(faces)
    m = n0 + inv.reshape(3, nf).T        # m[f] = midpoints of edges (01, 12, 20)
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
    new_faces = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1),
                                np.stack([m20, m12, c], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return np.concatenate([vertices, mid], axis=0), new_faces


def multimesh(levels: int):
    """-> vertices (N, 3) fp64, finest faces (F, 3), directed edge list (senders, receivers) of all levels."""
    v, f = icosahedron()
    pairs = []
    for lvl in range(levels + 1):
        if lvl:
            v, f = refine(v, f)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        pairs.append(np.concatenate([e, e[:, ::-1]], axis=0))
    e = np.unique(np.concatenate(pairs, axis=0), axis=0)
    return v, f, e[:, 0].copy(), e[:, 1].copy()


def latlon_to_xyz(lat_deg, lon_deg):
    lat, lon = np.deg2rad(lat_deg), np.deg2rad(lon_deg)
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def xyz_to_latlon(p):
    lat = np.arcsin(np.clip(p[..., 2], -1.0, 1.0))
    lon = np.arctan2(p[..., 1], p[..., 0])
    return lat, lon


def grid_points(nlat: int, nlon: int):
    """Grid nodes in the state's own order: latitude 90 -> -90 (row-major), longitude 0 -> 360 (exclusive)."""
    lat = np.linspace(90.0, -90.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    return latlon_to_xyz(la.reshape(-1), lo.reshape(-1)), lat, lon


def _local_frame_delta(p_send, p_recv):
    """sender - receiver, both rotated so that the receiver sits at (1, 0, 0)."""
    lat, lon = xyz_to_latlon(p_recv)
    cl, sl = np.cos(lon), np.sin(lon)
    # rotate by -lon about z
    x = cl * p_send[:, 0] + sl * p_send[:, 1]
    y = -sl * p_send[:, 0] + cl * p_send[:, 1]
    z = p_send[:, 2]
    # rotate by +lat about y (takes (cos lat, 0, sin lat) to (1, 0, 0))
    ca, sa = np.cos(lat), np.sin(lat)
    x2 = ca * x + sa * z
    z2 = -sa * x + ca * z
    return np.stack([x2 - 1.0, y, z2], axis=-1)


def edge_features(p_send, p_recv):
    d = _local_frame_delta(p_send, p_recv)
    ln = np.linalg.norm(d, axis=1, keepdims=True)
    f = np.concatenate([ln, d], axis=1)
    return (f / ln.max()).astype(np.float32)


def node_features(p):
    lat, lon = xyz_to_latlon(p)
    return np.stack([np.cos(lat), np.sin(lon), np.cos(lon)], axis=-1).astype(np.float32)


def _sort_by_receiver(s, r, n_recv):
    order = np.lexsort((s, r))
    s, r = s[order], r[order]
    ptr = np.zeros(n_recv + 1, dtype=np.int64)
    np.add.at(ptr, r + 1, 1)
    return s, r, np.cumsum(ptr)


def containing_faces(points, vertices, faces):
    """Index of the finest-mesh triangle containing each unit point (ties on edges: the lowest face index)."""
    from scipy.spatial import cKDTree
    nv = len(vertices)
    # faces incident to every vertex (5 or 6)
    inc = np.full((nv, 6), -1, dtype=np.int64)
    cnt = np.zeros(nv, dtype=np.int64)
    for col in range(3):
        for fi, vi in enumerate(faces[:, col]):
            inc[vi, cnt[vi]] = fi
            cnt[vi] += 1
    a, b, c = vertices[faces[:, 0]], vertices[faces[:, 1]], vertices[faces[:, 2]]
    nab, nbc, nca = np.cross(a, b), np.cross(b, c), np.cross(c, a)
    tree = cKDTree(vertices)
    out = np.full(len(points), -1, dtype=np.int64)
    best = np.full(len(points), -np.inf)
    for knn in (1, 2, 3):
        todo = np.nonzero(best < -1e-12)[0]
        if not len(todo):
            break
        _, near = tree.query(points[todo], k=knn)
        near = near if knn == 1 else near[:, knn - 1]
        cand = inc[near]                                   # (n, 6)
        p = points[todo][:, None, :]
        fc = np.where(cand < 0, 0, cand)
        m = np.minimum(np.minimum((p * nab[fc]).sum(-1), (p * nbc[fc]).sum(-1)), (p * nca[fc]).sum(-1))
        m = np.where(cand < 0, -np.inf, m)
        j = np.argmax(m, axis=1)
        mm = m[np.arange(len(todo)), j]
        better = mm > best[todo]
        out[todo[better]] = cand[np.arange(len(todo)), j][better]
        best[todo[better]] = mm[better]
    assert (best >= -1e-9).all(), "a grid point fell outside every candidate triangle"
    return out


def build_graph(nlat: int, nlon: int, levels: int, radius_frac: float = 0.6) -> "OrderedDict[str, np.ndarray]":
    """All graph tables of one (grid, mesh) pair.  Index arrays are int64; the engine receives them as fp32 arena
    entries (every index < 2**24 is exact in fp32: 721 x 1440 = 1,038,240 grid points)."""
    from scipy.spatial import cKDTree
    gp, lat, lon = grid_points(nlat, nlon)
    ng = len(gp)
    assert ng < (1 << 24)
    v, faces, ms, mr = multimesh(levels)
    nm = len(v)
    # Renumber the mesh nodes in the GRID's order (latitude bands north to south, longitude west to east inside a band).
    # multimesh() numbers them by refinement level, i.e. spatial neighbours are ~random rows of the node tables; the
    # engine gathers table rows by edge index, and with this numbering the rows touched by consecutive edges (sorted by
    # receiver) and by consecutive grid points fall into a moving window of a few bands that stays in L1 / L2.
    lat_v, lon_v = xyz_to_latlon(v)
    band_h = np.pi / max(2.0, np.sqrt(nm / 2.0))            # ~ one row of nodes per band
    order = np.lexsort((lon_v, np.floor((np.pi / 2 - lat_v) / band_h)))
    new_id = np.empty(nm, dtype=np.int64)
    new_id[order] = np.arange(nm)
    v, faces, ms, mr = v[order], new_id[faces], new_id[ms], new_id[mr]
    g = OrderedDict()
    g["n_grid"], g["n_mesh"] = ng, nm
    # ---- mesh edges, sorted by receiver
    ms, mr, mptr = _sort_by_receiver(ms, mr, nm)
    g["mesh.senders"], g["mesh.receivers"], g["mesh.ptr"] = ms, mr, mptr
    g["mesh.edge_feat"] = edge_features(v[ms], v[mr])
    g["mesh.node_feat"] = node_features(v)
    # ---- grid2mesh: radius query around every mesh node
    ff = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    max_len = np.linalg.norm(v[ff[:, 0]] - v[ff[:, 1]], axis=1).max()
    tree = cKDTree(gp)
    hits = tree.query_ball_point(v, r=radius_frac * max_len)
    recv = np.repeat(np.arange(nm), [len(h) for h in hits])
    send = np.concatenate([np.asarray(h, dtype=np.int64) for h in hits]) if len(recv) else np.zeros(0, np.int64)
    send, recv, gptr = _sort_by_receiver(send, recv, nm)
    g["g2m.senders"], g["g2m.receivers"], g["g2m.ptr"] = send, recv, gptr
    g["g2m.edge_feat"] = edge_features(gp[send], v[recv])
    # ---- mesh2grid: containing triangle, k-major rows
    fidx = containing_faces(gp, v, faces)
    m2g_s = faces[fidx].T.copy()                          # (3, ng): sender mesh node of edge k of grid point g
    g["m2g.senders"] = m2g_s.reshape(-1)
    g["m2g.receivers"] = np.tile(np.arange(ng), 3)
    g["m2g.edge_feat"] = edge_features(v[g["m2g.senders"]], gp[g["m2g.receivers"]])
    # ---- grid node structural features (cos lat | sin lon | cos lon), stored per axis
    g["grid.coslat"] = np.cos(np.deg2rad(lat)).astype(np.float32)
    g["grid.sinlon"] = np.sin(np.deg2rad(lon)).astype(np.float32)
    g["grid.coslon"] = np.cos(np.deg2rad(lon)).astype(np.float32)
    g["mesh.xyz"] = v
    g["mesh.faces"] = faces
    return g


_ARENA_KEYS = ("mesh.senders", "mesh.receivers", "mesh.ptr", "mesh.edge_feat", "mesh.node_feat", "g2m.senders",
               "g2m.receivers", "g2m.ptr", "g2m.edge_feat", "m2g.senders", "m2g.edge_feat", "grid.coslat", "grid.sinlon",
               "grid.coslon")


def graph_arena_entries(graph) -> "OrderedDict[str, np.ndarray]":
    """The tables the CUDA engine consumes, as fp32 arena entries named ``graph.*``."""
    out = OrderedDict()
    for k in _ARENA_KEYS:
        a = np.asarray(graph[k])
        if a.dtype.kind in "iu":
            assert a.size == 0 or a.max() < (1 << 24)
        out["graph." + k] = np.ascontiguousarray(a, dtype=np.float32)
    return out
