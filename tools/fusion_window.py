"""How far apart, in the node numbering, are the tiles a single-pass (phase A -> phase B through L2) schedule
would have to keep in flight?  CPU-only analysis behind DESIGN.md section 6 ("wavefront fusion").

Phase B of a tile needs fct_plus / fct_minus of every neighbour of its nodes, i.e. phase A of every tile that
holds such a neighbour.  With tiles = runs of TN consecutive nodes, tile(t) can run its phase B once phase A
has passed tile m(t) = max tile index of a neighbour; until then the 8 re-usable S_n-doubles per update of tile
t (fct_adf_h x3, fct_adf_v, ttf, fct_LO, fct_plus, fct_minus: 64 B per node-level) must survive in L2 while
all other traffic (about twice as much) streams through it.
usage: fusion_window.py [NXxNYxNL] [TN]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")

nx, ny, nl = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1536x1204x70").split("x")]
TN = int(sys.argv[2]) if len(sys.argv) > 2 else 35
L2 = 126e6


def analyse(label, m):
    N = m.myDim_nod2D
    e = m.edges.astype(np.int64) - 1
    tile = np.arange(N) // TN
    ntile = int(tile[-1]) + 1
    ta, tb = tile[e[:, 0]], tile[e[:, 1]]
    ahead = np.zeros(ntile, np.int64)
    np.maximum.at(ahead, ta, tb - ta)
    np.maximum.at(ahead, tb, ta - tb)
    halo = np.zeros(ntile, np.int64)
    cross = ta != tb
    # staged halo rows per tile = distinct foreign neighbours
    key = np.unique(np.concatenate([ta[cross] * N + e[cross, 1], tb[cross] * N + e[cross, 0]]))
    np.add.at(halo, key // N, 1)
    lev = (m.nlevels_nod2D[:N].astype(np.int64) - 1)
    per_node = 64.0 * lev.mean()                      # re-usable bytes per node
    stream = 3.0                                       # all traffic through L2 per re-usable byte (240 B vs 64 + the rest)
    window_nodes = L2 / (per_node * stream)
    window_tiles = window_nodes / TN
    pct = np.percentile(ahead, [50, 90, 99, 99.9, 100])
    inside = (ahead <= window_tiles).mean()
    lev_t = np.add.reduceat(lev, np.arange(0, N, TN))
    inside_w = (lev_t[ahead <= window_tiles].sum() / lev_t.sum())
    print(f"{label}: {N} nodes, {ntile} tiles of {TN}; staged rows per node {1 + halo.sum() / N:.2f}; "
          f"look-ahead (tiles) p50/p90/p99/p99.9/max = {pct[0]:.0f}/{pct[1]:.0f}/{pct[2]:.0f}/{pct[3]:.0f}/{pct[4]:.0f}; "
          f"L2 window ~{window_nodes:.0f} nodes = {window_tiles:.0f} tiles; tiles inside it {100 * inside:.1f} % "
          f"({100 * inside_w:.1f} % of the node-levels)")
    return inside_w


m = mesh.make_mesh(nx, ny, nl)
analyse("Hilbert numbering (what the plan uses)", m)
# bounded-bandwidth alternative: strips of h grid rows, column-major inside a strip (xy are the grid coordinates)
for h in (2, 4, 8, 16):
    x = np.rint(m.xy[:, 0] * nx - 0.5).astype(np.int64)
    y = np.rint(m.xy[:, 1] * ny - 0.5).astype(np.int64)
    order = np.lexsort((y % h, x, y // h))
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    tri = inv[m.elem2D_nodes.astype(np.int64) - 1]
    eorder = np.lexsort((tri.sum(1), tri.min(1)))
    m2 = mesh.build_mesh(tri[eorder], m.nlevels_elem[eorder], m.nl, m.myDim_nod2D, xy=m.xy[order])
    analyse(f"strips of {h} grid rows", m2)
