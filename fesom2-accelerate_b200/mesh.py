"""Seeded synthetic FESOM2-style meshes, tracer fields and node partitions.

Stand-in for the unseeded random inputs of the reference's kernel_tuner scripts
(/root/reference/kernels/fct_ale_a1.py:75-85, fct_ale_a2.py:142-146) and for the real FESOM2
mesh files the Fortran model reads.  Array conventions follow the reference exactly
(/root/reference/src/reference.cpp:309-334, :361, :396, :408-415):

* int32, 1-based connectivity: elem2D_nodes[3e+k], edges[2g+k], edge_tri[2g+k]
  (edge_tri[2g+1] <= 0 marks a boundary edge), nod_in_elem2D[n*dim+k], nod_in_elem2D_num[n];
* node-major, level-contiguous fields: item = node*(nl-1) + level; fct_adf_v / area / area_inv use
  stride nl per node; fct_adf_h uses stride nl-1 per edge;
* owned nodes first (myDim_nod2D), halo nodes after (eDim_nod2D).

Unlike the reference's random connectivity these are real manifold triangulations that honour the
mesh invariant the model guarantees: nlevels_nod2D[n] = max over the ring elements of nlevels_elem.
`adversarial_case` additionally reproduces the kernel_tuner-style random inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# BASELINE.json configs -> grid sizes
WORKLOADS = {
    "tiny": dict(nx=14, ny=11, nl=9),
    "pi": dict(nx=60, ny=52, nl=48),
    "core2": dict(nx=400, ny=317, nl=48),
    "dart": dict(nx=2048, ny=1560, nl=80),
    "ng5": dict(nx=3072, ny=2408, nl=70),
}


@dataclass
class Mesh:
    nl: int
    myDim_nod2D: int
    eDim_nod2D: int
    myDim_elem2D: int
    myDim_edge2D: int
    nlevels_nod2D: np.ndarray       # int32 [N+H]
    nlevels_elem: np.ndarray        # int32 [E]
    elem2D_nodes: np.ndarray        # int32 [E,3], 1-based
    nod_in_elem2D_num: np.ndarray   # int32 [N+H] (entries past N are 0)
    nod_in_elem2D: np.ndarray       # int32 [N+H, dim], 1-based, 0-padded
    nod_in_elem2D_dim: int
    edges: np.ndarray               # int32 [G,2], 1-based
    edge_tri: np.ndarray            # int32 [G,2], 1-based; second entry 0 on boundary edges
    xy: Optional[np.ndarray] = None           # float64 [N+H,2] node coordinates (partitioning only)
    node_gid: Optional[np.ndarray] = None     # int64 [N+H] global node id (partitions)
    elem_gid: Optional[np.ndarray] = None
    edge_gid: Optional[np.ndarray] = None

    @property
    def L(self) -> int:
        return self.nl - 1

    @property
    def nnod(self) -> int:
        return self.myDim_nod2D + self.eDim_nod2D

    # --- units of work ---
    def edge_depth(self) -> np.ndarray:
        """Active level count of each edge: max(nlev_e1, nlev_e2) - 1 (reference.cpp:412-414)."""
        e1 = self.edge_tri[:, 0] - 1
        e2 = self.edge_tri[:, 1] - 1
        nl1 = self.nlevels_elem[e1] - 1
        nl2 = np.where(e2 >= 0, self.nlevels_elem[np.maximum(e2, 0)] - 1, 0)
        return np.maximum(nl1, nl2).astype(np.int32)

    def S_n(self) -> int:
        return int((self.nlevels_nod2D[: self.myDim_nod2D].astype(np.int64) - 1).sum())

    def S_g(self) -> int:
        return int(self.edge_depth().astype(np.int64).sum())

    def bytes_alg(self) -> int:
        """Compulsory HBM bytes of one fct_ale tracer step."""
        return 8 * (21 * self.S_n() + 3 * self.S_g()) + 16 * self.myDim_nod2D


@dataclass
class Fields:
    """One tracer's inputs/outputs of fct_ale (reference.cpp:289, docs/refactoring.md:13-315)."""
    ttf: np.ndarray
    fct_LO: np.ndarray
    fct_adf_v: np.ndarray
    fct_adf_h: np.ndarray
    area: np.ndarray
    area_inv: np.ndarray
    hnode: np.ndarray
    hnode_new: np.ndarray
    del_ttf_advvert: np.ndarray
    del_ttf_advhoriz: np.ndarray
    fct_ttf_max: np.ndarray
    fct_ttf_min: np.ndarray
    fct_plus: np.ndarray
    fct_minus: np.ndarray
    UV_rhs: Optional[np.ndarray]
    dt: float = 0.5
    flux_eps: float = 1e-16
    bignumber: float = 1e3
    vlimit: int = 1
    # iterative branch (docs/refactoring.md:226-290): rejected flux parts, inputs of the next pass
    iter_yn: bool = False
    fct_adf_v2: Optional[np.ndarray] = None
    fct_adf_h2: Optional[np.ndarray] = None

    def copy(self) -> "Fields":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.copy() if isinstance(v, np.ndarray) else v
        return Fields(**kw)


def hilbert_index(ix: np.ndarray, iy: np.ndarray, order: int) -> np.ndarray:
    """Vectorised Hilbert-curve index of integer points on a 2^order grid."""
    x = ix.astype(np.int64).copy()
    y = iy.astype(np.int64).copy()
    n = np.int64(1) << order
    d = np.zeros_like(x)
    s = n >> 1
    while s > 0:
        rx = (x & s) > 0
        ry = (y & s) > 0
        d += s * s * ((3 * rx.astype(np.int64)) ^ ry.astype(np.int64))
        flip = (~ry) & rx
        x = np.where(flip, n - 1 - x, x)
        y = np.where(flip, n - 1 - y, y)
        x, y = np.where(~ry, y, x), np.where(~ry, x, y)
        s >>= 1
    return d


def _bathymetry(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Smooth synthetic depth in [0,1]: a deep basin with ridges and a shelf towards the rim."""
    d = 0.80 + 0.22 * np.sin(2 * np.pi * (1.3 * x + 0.15)) * np.cos(2 * np.pi * (0.9 * y - 0.1))
    d += 0.12 * np.sin(2 * np.pi * (3.1 * x + 2.3 * y))
    rim = np.minimum(np.minimum(x, 1 - x), np.minimum(y, 1 - y))
    d *= np.clip(rim * 12.0, 0.08, 1.0)
    return np.clip(d, 0.0, 1.0)


def make_mesh(nx: int, ny: int, nl: int, seed: int = 0, land: bool = True,
              order: str = "hilbert") -> Mesh:
    """Triangulated nx x ny grid (2 triangles per quad, alternating diagonal), optional land mask
    (removed elements -> interior boundary edges with edge_tri[:,1] = 0), nodes renumbered along a
    Hilbert curve so that consecutive nodes are spatial neighbours."""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    ii = ii.ravel()
    jj = jj.ravel()
    a = jj * nx + ii
    b = a + 1
    c = a + nx + 1
    d = a + nx
    even = ((ii + jj) & 1) == 0
    t1 = np.where(even[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t2 = np.where(even[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    tri = np.empty((2 * a.size, 3), dtype=np.int64)
    tri[0::2] = t1
    tri[1::2] = t2

    gx = (np.arange(nx) + 0.5) / nx
    gy = (np.arange(ny) + 0.5) / ny
    X = np.tile(gx, ny)
    Y = np.repeat(gy, nx)
    cx = X[tri].mean(1)
    cy = Y[tri].mean(1)
    if land:
        keep = np.ones(tri.shape[0], dtype=bool)
        nblob = 5
        for _ in range(nblob):
            bx, by = rng.uniform(0.15, 0.85, 2)
            ra, rb = rng.uniform(0.02, 0.07, 2)
            th = rng.uniform(0, np.pi)
            ux = (cx - bx) * np.cos(th) + (cy - by) * np.sin(th)
            uy = -(cx - bx) * np.sin(th) + (cy - by) * np.cos(th)
            keep &= (ux / ra) ** 2 + (uy / rb) ** 2 > 1.0
        tri = tri[keep]
        cx = cx[keep]
        cy = cy[keep]

    # drop orphan nodes, renumber along the space-filling curve
    used = np.zeros(nx * ny, dtype=bool)
    used[tri.ravel()] = True
    old_ids = np.flatnonzero(used)
    oi = old_ids % nx
    oj = old_ids // nx
    if order == "hilbert":
        k = max(1, int(np.ceil(np.log2(max(nx, ny)))))
        key = hilbert_index(oi, oj, k)
        perm = np.argsort(key, kind="stable")
    elif order == "random":
        perm = rng.permutation(old_ids.size)
    else:
        perm = np.arange(old_ids.size)
    new_of_old = np.full(nx * ny, -1, dtype=np.int64)
    new_of_old[old_ids[perm]] = np.arange(old_ids.size)
    N = old_ids.size
    xy = np.stack([X[old_ids[perm]], Y[old_ids[perm]]], 1)
    tri = new_of_old[tri]
    # elements ordered by their lowest node id (keeps element ids local to node ids)
    eorder = np.lexsort((tri.sum(1), tri.min(1)))
    tri = tri[eorder]
    cx = cx[eorder]
    cy = cy[eorder]

    depth = _bathymetry(cx, cy)
    nlev_e = np.clip(np.rint(3 + depth * (nl - 3)), 3, nl).astype(np.int32)
    return build_mesh(tri, nlev_e, nl, N, xy=xy)


def build_mesh(tri: np.ndarray, nlev_e: np.ndarray, nl: int, N: int, xy=None,
               nlev_n: Optional[np.ndarray] = None, H: int = 0) -> Mesh:
    """Derive edges / edge_tri / nod_in_elem2D / nlevels_nod2D from 0-based triangles."""
    E = tri.shape[0]
    NT = N + H
    # --- edges: unique undirected node pairs; oriented like the first adjacent element winds ---
    he_a = tri[:, [0, 1, 2]].ravel()
    he_b = tri[:, [1, 2, 0]].ravel()
    he_e = np.repeat(np.arange(E, dtype=np.int64), 3)
    lo = np.minimum(he_a, he_b)
    hi = np.maximum(he_a, he_b)
    key = lo * np.int64(NT) + hi
    o = np.argsort(key, kind="stable")
    ks = key[o]
    first = np.ones(ks.size, dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    start = np.flatnonzero(first)
    cnt = np.diff(np.append(start, ks.size))
    if cnt.max() > 2:
        raise ValueError("non-manifold edge")
    G = start.size
    edges = np.stack([he_a[o[start]], he_b[o[start]]], 1) + 1
    e1 = he_e[o[start]]
    e2 = np.where(cnt == 2, he_e[o[np.minimum(start + 1, ks.size - 1)]], -1)
    edge_tri = np.stack([e1 + 1, e2 + 1], 1)

    # --- node -> elements table (ascending element id) ---
    nn = tri.ravel()
    ne = np.repeat(np.arange(E, dtype=np.int64), 3)
    o2 = np.argsort(nn, kind="stable")
    nns = nn[o2]
    num = np.bincount(nns, minlength=NT).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(num)])
    pos = np.arange(nns.size) - offs[nns]
    dim = int(num.max())
    nie = np.zeros((NT, dim), dtype=np.int32)
    nie[nns, pos] = ne[o2] + 1
    if nlev_n is None:
        nlev_n = np.zeros(NT, dtype=np.int32)
        np.maximum.at(nlev_n, nn, np.repeat(nlev_e, 3))
    num32 = num.astype(np.int32)
    if H:
        # the reference only walks the rings of owned nodes (reference.cpp:358-361)
        num32[N:] = 0
        nie[N:] = 0
    return Mesh(nl=nl, myDim_nod2D=N, eDim_nod2D=H, myDim_elem2D=E, myDim_edge2D=G,
                nlevels_nod2D=np.ascontiguousarray(nlev_n, dtype=np.int32),
                nlevels_elem=np.ascontiguousarray(nlev_e, dtype=np.int32),
                elem2D_nodes=np.ascontiguousarray(tri + 1, dtype=np.int32),
                nod_in_elem2D_num=num32, nod_in_elem2D=nie, nod_in_elem2D_dim=dim,
                edges=np.ascontiguousarray(edges, dtype=np.int32),
                edge_tri=np.ascontiguousarray(edge_tri, dtype=np.int32), xy=xy)


def make_delaunay_mesh(n_points: int, nl: int, seed: int = 0, land: bool = True) -> Mesh:
    """UNSTRUCTURED triangulation: Delaunay of jittered-random points in the unit square (node degrees
    3..12 instead of the 4 / 8 of the criss-cross grids), optional land blobs, the same bathymetry and
    Hilbert renumbering as make_mesh.  The irregular counterpart of the structured workloads: ragged
    per-node edge lists, tiles of uneven size, irregular halos once partitioned."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    # blue-noise-ish: one point per cell of a sqrt(n) grid, jittered by up to 0.45 of a cell, plus 15 % fully random points
    g = max(2, int(np.sqrt(n_points * 0.85)))
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="xy")
    pts = np.stack([(ii.ravel() + 0.5 + rng.uniform(-0.45, 0.45, g * g)) / g,
                    (jj.ravel() + 0.5 + rng.uniform(-0.45, 0.45, g * g)) / g], 1)
    extra = max(0, n_points - g * g)
    pts = np.concatenate([pts, rng.uniform(0.0, 1.0, (extra, 2))])
    tri = Delaunay(pts).simplices.astype(np.int64)
    # drop the sliver triangles of the convex hull (needle-shaped, degenerate areas)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    area2 = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    longest = np.maximum.reduce([np.linalg.norm(b - a, axis=1), np.linalg.norm(c - b, axis=1), np.linalg.norm(a - c, axis=1)])
    keep = area2 > 0.05 * longest ** 2
    cx, cy = pts[tri].mean(1).T
    if land:
        for _ in range(5):
            bx, by = rng.uniform(0.15, 0.85, 2)
            ra, rb = rng.uniform(0.02, 0.07, 2)
            th = rng.uniform(0, np.pi)
            ux = (cx - bx) * np.cos(th) + (cy - by) * np.sin(th)
            uy = -(cx - bx) * np.sin(th) + (cy - by) * np.cos(th)
            keep &= (ux / ra) ** 2 + (uy / rb) ** 2 > 1.0
    tri, cx, cy = tri[keep], cx[keep], cy[keep]
    used = np.zeros(pts.shape[0], dtype=bool)
    used[tri.ravel()] = True
    old_ids = np.flatnonzero(used)
    k = 16
    q = np.minimum((pts[old_ids] * (1 << k)).astype(np.int64), (1 << k) - 1)
    perm = np.argsort(hilbert_index(q[:, 0], q[:, 1], k), kind="stable")
    new_of_old = np.full(pts.shape[0], -1, dtype=np.int64)
    new_of_old[old_ids[perm]] = np.arange(old_ids.size)
    tri = new_of_old[tri]
    eorder = np.lexsort((tri.sum(1), tri.min(1)))
    tri, cx, cy = tri[eorder], cx[eorder], cy[eorder]
    nlev_e = np.clip(np.rint(3 + _bathymetry(cx, cy) * (nl - 3)), 3, nl).astype(np.int32)
    return build_mesh(tri, nlev_e, nl, old_ids.size, xy=pts[old_ids[perm]])


def make_workload(name: str, seed: int = 0) -> Mesh:
    w = WORKLOADS[name]
    return make_mesh(w["nx"], w["ny"], w["nl"], seed=seed)


SENTINEL = -7.0e0   # value of never-written output cells; parity tests compare these too


def make_fields(mesh: Mesh, seed: int = 1, with_uv: bool = True, poison: bool = True,
                realistic_area: bool = False) -> Fields:
    """Seeded tracer fields.  Inactive input levels are poisoned with huge
    values when `poison` so an out-of-depth read shows up in every result."""
    rng = np.random.default_rng(seed)
    NT, L, nl, G = mesh.nnod, mesh.L, mesh.nl, mesh.myDim_edge2D
    z = np.arange(L)[None, :]
    act = z < (mesh.nlevels_nod2D[:, None] - 1)
    prof = 10.0 + 15.0 * np.exp(-z / max(4.0, L / 5.0))
    ttf = prof + 0.1 * rng.standard_normal((NT, L))
    lo = ttf + 0.01 * rng.standard_normal((NT, L))
    zv = np.arange(nl)[None, :]
    adf_v = rng.standard_normal((NT, nl))
    adf_v[zv >= (mesh.nlevels_nod2D[:, None] - 1)] = 0.0      # bottom flux is zero (md:232)
    adf_h = rng.standard_normal((G, L))
    adf_h[z >= mesh.edge_depth()[:, None]] = 0.0
    if realistic_area:
        area = 10.0 ** rng.uniform(8.0, 10.0, (NT, nl))
    else:
        area = rng.uniform(1.0, 2.0, (NT, nl))
    hnode = rng.uniform(5.0, 50.0, (NT, L))
    hnode_new = hnode * (1.0 + 1e-3 * rng.standard_normal((NT, L)))
    del_v = rng.standard_normal((NT, L))
    del_h = rng.standard_normal((NT, L))
    if poison:
        big = 1.0e30
        ttf = np.where(act, ttf, big)
        lo = np.where(act, lo, -big)
        hnode = np.where(act, hnode, big)
        hnode_new = np.where(act, hnode_new, big)
    f = Fields(
        ttf=ttf, fct_LO=lo, fct_adf_v=adf_v, fct_adf_h=adf_h, area=area, area_inv=1.0 / area,
        hnode=hnode, hnode_new=hnode_new, del_ttf_advvert=del_v, del_ttf_advhoriz=del_h,
        fct_ttf_max=np.full((NT, L), SENTINEL), fct_ttf_min=np.full((NT, L), SENTINEL),
        fct_plus=np.full((NT, L), SENTINEL), fct_minus=np.full((NT, L), SENTINEL),
        UV_rhs=np.full((mesh.myDim_elem2D, L, 2), SENTINEL) if with_uv else None)
    for k, v in f.__dict__.items():
        if isinstance(v, np.ndarray):
            setattr(f, k, np.ascontiguousarray(v, dtype=np.float64))
    return f


def adversarial_case(nodes: int, nl: int, seed: int = 0, max_ring: int = 6):
    """kernel_tuner-style inputs (fct_ale_a1.py:75-85, fct_ale_a2.py:142-146,
    fct_ale_b1_horizontal.py:119-123): pure randn fields, random depths in [3, nl-1], random
    non-manifold connectivity, right element possibly 0.  The only constraint kept is the one
    the reference itself needs to stay in bounds and to read initialised data: a node is at least
    as deep as every element that references it."""
    rng = np.random.default_rng(seed)
    N = nodes
    E = 2 * N
    G = 3 * N
    L = nl - 1
    tri = rng.integers(0, N, (E, 3))
    nlev_e = rng.integers(3, nl, E).astype(np.int32)
    nlev_n = rng.integers(3, nl, N).astype(np.int32)
    np.maximum.at(nlev_n, tri.ravel(), np.repeat(nlev_e, 3))
    num = rng.integers(1, max_ring + 1, N).astype(np.int32)
    nie = np.zeros((N, max_ring), dtype=np.int32)
    for k in range(max_ring):
        nie[:, k] = np.where(k < num, rng.integers(1, E + 1, N), 0)
    # a3 reads UV_rhs up to the node's depth, which is always written (fill) -> any element is fine
    edges = rng.integers(1, N + 1, (G, 2)).astype(np.int32)
    edge_tri = np.stack([rng.integers(1, E + 1, G), rng.integers(0, E + 1, G)], 1).astype(np.int32)
    m = Mesh(nl=nl, myDim_nod2D=N, eDim_nod2D=0, myDim_elem2D=E, myDim_edge2D=G,
             nlevels_nod2D=nlev_n, nlevels_elem=nlev_e,
             elem2D_nodes=np.ascontiguousarray(tri + 1, dtype=np.int32),
             nod_in_elem2D_num=num, nod_in_elem2D=nie, nod_in_elem2D_dim=max_ring,
             edges=edges, edge_tri=edge_tri)
    # b1h / b3h / c_h touch node columns down to the edge depth: deepen the end nodes accordingly
    ed = m.edge_depth() + 1
    np.maximum.at(m.nlevels_nod2D, (edges - 1).ravel(), np.repeat(ed, 2))
    f = make_fields(m, seed=seed + 100, poison=False)
    r = np.random.default_rng(seed + 200)
    for k in ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h"):
        getattr(f, k)[...] = r.standard_normal(getattr(f, k).shape)
    return m, f


# ----------------------------------------------------------------------------------------------
# METIS-style node partitioning: contiguous chunks of the space-filling
# curve order, balanced on active node-levels.  Local numbering follows FESOM: owned nodes first,
# halo nodes after (grouped by owner rank, ascending global id inside a group); local elements =
# all elements touching an owned node; local edges = all edges touching an owned node.  Local
# element/edge ids keep the global order, so every per-node edge list is visited in the same order
# as in the single-domain run and the partitioned sums are bit-identical to it.
# ----------------------------------------------------------------------------------------------
@dataclass
class Partition:
    rank: int
    nparts: int
    mesh: Mesh
    owner_of_halo: np.ndarray                # int32 [H] owning rank of each halo node
    send_lists: Dict[int, np.ndarray]        # peer -> local owned node ids (0-based) to send
    recv_ranges: Dict[int, tuple]            # peer -> (first local halo node id, count)
    boundary_nodes: np.ndarray = field(default=None)   # owned nodes with a halo neighbour
    interior_nodes: np.ndarray = field(default=None)   # owned nodes with only owned neighbours


def partition_bounds(mesh: Mesh, nparts: int) -> np.ndarray:
    w = np.cumsum(mesh.nlevels_nod2D[: mesh.myDim_nod2D].astype(np.int64) - 1)
    tgt = w[-1] * np.arange(1, nparts) / nparts
    cuts = np.searchsorted(w, tgt)
    return np.concatenate([[0], cuts, [mesh.myDim_nod2D]]).astype(np.int64)


def grow_partition(mesh: Mesh, nparts: int, seed: int = 0) -> np.ndarray:
    """Greedy graph growing (the initial-partitioning heuristic of METIS-style k-way partitioners): part r
    grows breadth-first over the node graph from a peripheral seed until it holds 1/nparts of the
    active node-levels; what is left (possibly several disconnected pieces) joins the last part.  The
    parts are NOT runs of the node numbering: ragged boundaries, uneven halos, several peers per part."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import breadth_first_order
    N = mesh.myDim_nod2D
    e = mesh.edges.astype(np.int64) - 1
    adj = coo_matrix((np.ones(2 * e.shape[0], np.int8), (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])), shape=(N, N)).tocsr()
    w = mesh.nlevels_nod2D[:N].astype(np.int64) - 1
    target = w.sum() / nparts
    owner = np.full(N, -1, np.int32)
    rng = np.random.default_rng(seed)
    for r in range(nparts - 1):
        free = np.flatnonzero(owner < 0)
        sub = adj[free][:, free]
        # peripheral seed: the last node of a breadth-first sweep from a random free node
        start = int(rng.integers(free.size))
        order = breadth_first_order(sub, start, directed=False, return_predecessors=False)
        order = breadth_first_order(sub, int(order[-1]), directed=False, return_predecessors=False)
        cum = np.cumsum(w[free[order]])
        take = int(np.searchsorted(cum, target)) + 1
        owner[free[order[:take]]] = r
    owner[owner < 0] = nparts - 1
    return owner


def partition_mesh(mesh: Mesh, nparts: int, ranks: Optional[List[int]] = None, owner: Optional[np.ndarray] = None) -> List[Partition]:
    """owner: int32 [N] owning part of every node (e.g. grow_partition); default: contiguous runs of the
    space-filling-curve numbering, balanced on active node-levels."""
    if mesh.eDim_nod2D:
        raise ValueError("partition a single-domain mesh")
    N = mesh.myDim_nod2D
    if owner is None:
        bounds = partition_bounds(mesh, nparts)
        owner = (np.searchsorted(bounds, np.arange(N), side="right") - 1).astype(np.int32)
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    tri = mesh.elem2D_nodes.astype(np.int64) - 1
    edg = mesh.edges.astype(np.int64) - 1
    parts = []
    halo_gids = {}
    wanted = range(nparts) if ranks is None else ranks
    tri_owner = owner[tri]

    def halo_of(r):
        """halo nodes of rank r, grouped by owner, ascending gid inside a group (cached)."""
        if r not in halo_gids:
            emask = (tri_owner == r).any(1)
            nodes_r = np.unique(tri[emask].ravel())
            hg = nodes_r[owner[nodes_r] != r]
            halo_gids[r] = hg[np.lexsort((hg, owner[hg]))]
        return halo_gids[r]

    for r in wanted:
        own_g = np.flatnonzero(owner == r)          # ascending global ids: the local order of the owned nodes
        n_own = own_g.size
        hg = halo_of(r)
        H = hg.size
        gids = np.concatenate([own_g, hg])
        loc = np.full(N, -1, dtype=np.int64)
        loc[gids] = np.arange(gids.size)
        emask = (tri_owner == r).any(1)
        eg = np.flatnonzero(emask)
        gmask = (owner[edg] == r).any(1)
        gg = np.flatnonzero(gmask)
        eloc = np.full(mesh.myDim_elem2D + 1, 0, dtype=np.int64)   # 1-based, 0 stays 0
        eloc[eg + 1] = np.arange(eg.size) + 1
        ltri = loc[tri[eg]]
        et = mesh.edge_tri[gg].astype(np.int64)
        ledge_tri = eloc[et]
        # an edge keeps its (left, right) roles; if the left element is not local the roles cannot
        # be kept -- impossible here because both adjacent elements contain an owned node.
        assert (ledge_tri[:, 0] > 0).all()
        assert ((et[:, 1] > 0) == (ledge_tri[:, 1] > 0)).all()
        nlev_n = mesh.nlevels_nod2D[gids]
        lm = build_mesh(ltri, mesh.nlevels_elem[eg], mesh.nl, n_own,
                        xy=None if mesh.xy is None else mesh.xy[gids], nlev_n=nlev_n, H=H)
        # build_mesh re-derives edges from all local elements, which also yields halo-halo edges;
        # FESOM's myDim_edge2D only holds edges touching an owned node, in global order.
        ledges = loc[edg[gg]] + 1
        lm.edges = np.ascontiguousarray(ledges, dtype=np.int32)
        lm.edge_tri = np.ascontiguousarray(ledge_tri, dtype=np.int32)
        lm.myDim_edge2D = gg.size
        lm.node_gid = gids
        lm.elem_gid = eg
        lm.edge_gid = gg
        own_h = owner[hg]
        recv = {}
        for p in np.unique(own_h):
            idx = np.flatnonzero(own_h == p)
            recv[int(p)] = (int(n_own + idx[0]), int(idx.size))
        send = {}
        # node adjacency is symmetric: the ranks holding my nodes in their halo are the owners of
        # my halo nodes
        for p in [int(q) for q in np.unique(own_h)]:
            hp = halo_of(p)
            mine = hp[owner[hp] == r]
            if mine.size:
                send[int(p)] = loc[mine].astype(np.int32)
        nb = np.zeros(n_own, dtype=bool)
        le = lm.edges.astype(np.int64) - 1
        cut = (le >= n_own).any(1)
        ends = le[cut].ravel()
        nb[ends[ends < n_own]] = True
        parts.append(Partition(rank=r, nparts=nparts, mesh=lm, owner_of_halo=own_h.astype(np.int32),
                               send_lists=send, recv_ranges=recv,
                               boundary_nodes=np.flatnonzero(nb).astype(np.int32),
                               interior_nodes=np.flatnonzero(~nb).astype(np.int32)))
    return parts


def slice_fields(f: Fields, part: Partition) -> Fields:
    """Local view (copies) of single-domain fields for one partition."""
    g = part.mesh.node_gid
    kw = {}
    for k, v in f.__dict__.items():
        if not isinstance(v, np.ndarray):
            kw[k] = v
        elif k in ("fct_adf_h", "fct_adf_h2"):
            kw[k] = v[part.mesh.edge_gid].copy()
        elif k == "UV_rhs":
            kw[k] = v[part.mesh.elem_gid].copy()
        else:
            kw[k] = v[g].copy()
    return Fields(**kw)


def fast_fields(mesh: Mesh, seed: int = 1, alloc=None, with_uv: bool = False) -> Fields:
    """Benchmark-sized synthetic fields in seconds rather than minutes: every array is filled by
    tiling one random block (different phase and affine map per array).  Same value ranges as
    make_fields; no poisoning and no zeroing below the sea floor (those cells are never used).
    `alloc(shape)` lets the caller place the arrays in page-locked memory.

    The value of a cell is a function of its GLOBAL row id (mesh.node_gid / edge_gid / elem_gid when
    the mesh is a partition, else the row index), its column, the array and the seed only: the
    partitions of a mesh hold exactly the rows of the single-domain arrays, whatever the number of
    partitions (what bench.py's digest of the outputs relies on)."""
    rng = np.random.default_rng(seed)
    alloc = alloc or (lambda shape: np.empty(shape, dtype=np.float64))
    B = 1 << 22
    base = rng.standard_normal(B + 4096)
    NT, L, nl, G = mesh.nnod, mesh.L, mesh.nl, mesh.myDim_edge2D

    def fill(shape, scale, offset, phase, clip=None, kind="node"):
        a = alloc(shape)
        blk = base[phase: phase + B] * scale + offset
        if clip is not None:
            np.clip(blk, clip[0], clip[1], out=blk)
        gid = {"node": mesh.node_gid, "edge": mesh.edge_gid, "elem": mesh.elem_gid}[kind]
        if gid is None:
            flat = a.reshape(-1)
            for i in range(0, flat.size, B):
                n = min(B, flat.size - i)
                flat[i:i + n] = blk[:n]
            return a
        # partition: row r of the local array is row gid[r] of the single-domain array
        W = int(np.prod(shape[1:]))
        a2 = a.reshape(shape[0], W)
        cols = np.arange(W, dtype=np.int64)[None, :]
        step = max(1, (1 << 22) // max(W, 1))
        g64 = np.asarray(gid, dtype=np.int64)
        for r in range(0, shape[0], step):
            idx = (g64[r:r + step, None] * W + cols) & (B - 1)
            a2[r:r + step] = blk[idx]
        return a

    ttf = fill((NT, L), 0.1, 10.0, 0)
    lo = fill((NT, L), 0.1, 10.0, 0)
    lo += fill((NT, L), 0.01, 0.0, 17)
    area = fill((NT, nl), 0.25, 1.5, 101, clip=(1.0, 2.0))
    area_inv = alloc((NT, nl))
    np.divide(1.0, area, out=area_inv)
    hnode = fill((NT, L), 10.0, 27.0, 211, clip=(5.0, 50.0))
    hnode_new = fill((NT, L), 10.0, 27.0, 211, clip=(5.0, 50.0))
    hnode_new *= 1.0 + 1e-3 * base[307]
    f = Fields(
        ttf=ttf, fct_LO=lo, fct_adf_v=fill((NT, nl), 1.0, 0.0, 401), fct_adf_h=fill((G, L), 1.0, 0.0, 503, kind="edge"),
        area=area, area_inv=area_inv, hnode=hnode, hnode_new=hnode_new,
        del_ttf_advvert=fill((NT, L), 1.0, 0.0, 601), del_ttf_advhoriz=fill((NT, L), 1.0, 0.0, 701),
        fct_ttf_max=fill((NT, L), 0.0, SENTINEL, 0), fct_ttf_min=fill((NT, L), 0.0, SENTINEL, 0),
        fct_plus=fill((NT, L), 0.0, SENTINEL, 0), fct_minus=fill((NT, L), 0.0, SENTINEL, 0),
        UV_rhs=fill((mesh.myDim_elem2D, L, 2), 0.0, SENTINEL, 0, kind="elem") if with_uv else None)
    return f


# ----------------------------------------------------------------------------------------------
# Order-independent digest of a node array over the OWNED, ACTIVE cells of a (partition of a) mesh:
# the sum modulo 2^64 of a 64-bit mix of (global node id, level, value bits).  The digests of the
# partitions of a mesh add up (mod 2^64) to the digest of the single-domain array exactly when every
# owned cell holds the same bits, whatever the number of partitions (bench.py: `digest`).
# ----------------------------------------------------------------------------------------------
_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def digest_node_array(mesh: Mesh, a: np.ndarray, salt: int = 0) -> int:
    n = mesh.myDim_nod2D
    W = a.shape[1]
    gid = np.arange(n, dtype=np.int64) if mesh.node_gid is None else np.asarray(mesh.node_gid[:n], dtype=np.int64)
    cols = np.arange(W, dtype=np.uint64)[None, :]
    depth = (mesh.nlevels_nod2D[:n].astype(np.int64) - 1)[:, None]
    total = 0
    step = max(1, (1 << 22) // max(W, 1))
    with np.errstate(over="ignore"):
        for r in range(0, n, step):
            v = (a[r:min(r + step, n)] + 0.0).view(np.uint64)          # + 0.0: -0 and +0 are the same value
            key = (gid[r:r + step, None].astype(np.uint64) * np.uint64(W) + cols + np.uint64(salt)) * _M1
            h = (v ^ key) * _M2
            h ^= h >> np.uint64(29)
            h *= _M3
            h ^= h >> np.uint64(32)
            act = np.arange(W)[None, :] < depth[r:r + step]
            total = (total + int(h[act].sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    return total
