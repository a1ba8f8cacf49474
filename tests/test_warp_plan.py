"""Host logic of the warp-item kernels, checked without a GPU: the inspector's per-tile tables
(fct_ale_plan_inspect_, a host-only entry point of the product library) are interpreted by a small
numpy/Python executor that follows fct_warp_kernels.cuh step by step -- bulk copies of the listed
rows into a byte image of the CTA's shared memory, the in-place a1 pass, the warp-item schedule,
the neighbour lookups through precomputed byte offsets, the virtual-lane stencil -- and the result
is compared with the oracle.  A wrong offset, depth, role flag, schedule or summation order in the
tables shows up here before any kernel runs."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal

INF = float("inf")


def inspect(abi, m, tile_nodes=64, smem_cap=74 * 1024, which=0, packed=0):
    lib = abi.load()
    cap_words = 64 * 1024 * 1024 // 4
    blob = np.zeros(cap_words, np.uint32)
    off = np.zeros(m.myDim_nod2D + 2, np.uint32)
    nt, sm, st = C.c_int(), C.c_int(), C.c_int()
    u32p = C.POINTER(C.c_uint32)
    lib.fct_ale_plan_inspect_(
        abi.ci(m.myDim_nod2D), abi.ci(m.eDim_nod2D), abi.ci(m.myDim_elem2D), abi.ci(m.myDim_edge2D),
        abi.ci(m.nl), abi.iptr(m.nlevels_nod2D), abi.iptr(m.nlevels_elem),
        abi.iptr(m.elem2D_nodes.reshape(-1)), abi.iptr(m.nod_in_elem2D_num),
        abi.iptr(m.nod_in_elem2D.reshape(-1)), abi.ci(m.nod_in_elem2D_dim), abi.iptr(m.edges.reshape(-1)),
        abi.iptr(m.edge_tri.reshape(-1)), abi.ci(tile_nodes), abi.ci(smem_cap), abi.ci(which), abi.ci(packed),
        C.byref(C.c_longlong(cap_words * 4)), blob.ctypes.data_as(u32p), abi.ci(off.size),
        off.ctypes.data_as(u32p), C.byref(nt), C.byref(sm), C.byref(st))
    return st.value, nt.value, sm.value, blob, off


class Tile:
    """Decoded blob of one tile + the shared-memory image the kernels build from it."""

    def __init__(self, words, P):
        b = words.tobytes()
        h = np.frombuffer(b, np.int32, 16)
        (self.n_copies, self.n_erows, self.n_nodes, self.n_witems, self.n_rows, off_hdr, off_ent, off_sched,
         blob_bytes, self.rows_bytes, self.erows_bytes, self.tx) = [int(x) for x in h[:12]]
        assert blob_bytes == len(b)
        self.copies = np.frombuffer(b, np.uint32, 2 * self.n_copies, 64).reshape(-1, 2)
        self.hdr = np.frombuffer(b, np.uint32, 4 * self.n_nodes, off_hdr).reshape(-1, 4)
        n_ent = (off_sched - off_ent) // 16
        self.ent = np.frombuffer(b, np.uint32, 4 * n_ent, off_ent).reshape(-1, 4)
        self.sched = np.frombuffer(b, np.uint16, self.n_witems * 32, off_sched).reshape(self.n_witems, 32)
        self.P = P

    def stage(self, src_a, src_b, src_e):
        """the bulk copies of the issuer warps: returns (rowsA, rowsB, erows) as float64 images (NaN = never
        written); the three regions are consecutive in the stage"""
        ra, re = self.rows_bytes // 8, self.erows_bytes // 8
        img = np.full(2 * ra + re + 2 * self.P, np.nan)
        src = (src_a, src_b, src_e)
        lim = (ra, 2 * ra, 2 * ra + re)
        tx = 0
        for goff, pk in self.copies:
            pk = int(pk)
            so, sz, arr = (pk & 0x3fff) * 2, ((pk >> 14) & 0x3fff) * 2, (pk >> 28) & 3     # in doubles
            assert sz > 0 and arr < 3 and so + sz <= lim[arr] and so >= (0, ra, 2 * ra)[arr]
            assert np.isnan(img[so:so + sz]).all(), "two copies overlap"
            img[so:so + sz] = src[arr][goff:goff + sz]
            tx += sz * 8
        assert tx == self.tx
        return img[:ra + 2 * self.P].copy(), img[ra:2 * ra + 2 * self.P].copy(), img[2 * ra:].copy()


def pmax(a, b):
    return b if a < b else a


def pmin(a, b):
    return b if b < a else a


def pad(a, P):
    out = np.zeros((a.shape[0], P))
    out[:, :a.shape[1]] = a
    return out.reshape(-1)


class Packed:
    """The packed level storage: node n owns round_up_even(nlev-1 + 1) slots, edge g round_up_even(depth)."""

    def __init__(self, m):
        nz = np.maximum(m.nlevels_nod2D.astype(np.int64) - 1, 0)
        self.ncap = (nz + 2) & ~1
        self.ecap = (m.edge_depth().astype(np.int64) + 1) & ~1
        self.ncol = np.concatenate([[0], np.cumsum(self.ncap)])
        self.ecol = np.concatenate([[0], np.cumsum(self.ecap)])

    def pack(self, a, edge=False):
        cap, col = (self.ecap, self.ecol) if edge else (self.ncap, self.ncol)
        out = np.full(int(col[-1]), np.nan)
        for r in range(a.shape[0]):
            w = min(int(cap[r]), a.shape[1])
            out[col[r]:col[r] + w] = a[r, :w]
        return out

    def unpack(self, v, like, edge=False):
        cap, col = (self.ecap, self.ecol) if edge else (self.ncap, self.ncol)
        out = np.array(like)
        for r in range(like.shape[0]):
            w = min(int(cap[r]), like.shape[1])
            out[r, :w] = v[col[r]:col[r] + w]
        return out


def emulate(m, f, blob, off, ntiles, packed=None, vlimit=1):
    """Both fused phases on the padded (or packed) layout; returns the dict of device-layout result arrays.
    vlimit 2 / 3: the variant of phase A that reads the own a1 maxima of levels z-1..z+1 from the
    staged own row (wt_item_a<false> in fct_warp_kernels.cuh)."""
    P = (m.nl + 7) & ~7
    L = m.L
    names = ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "area", "area_inv", "hnode", "hnode_new", "del_ttf_advvert",
             "del_ttf_advhoriz", "fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus")
    if packed is None:
        g = {k: pad(getattr(f, k), P) for k in names}
    else:
        g = {k: packed.pack(getattr(f, k), edge=(k == "fct_adf_h")) for k in names}
    g["adf_v_out"] = g["fct_adf_v"].copy()
    g["adf_h_out"] = g["fct_adf_h"].copy()
    dt, eps, big = f.dt, f.flux_eps, f.bignumber
    W = 32
    seen_slots = set()
    for phase in "AB":
        for t in range(ntiles):
            T = Tile(blob[off[t] * 4: off[t + 1] * 4], P)
            if phase == "A":
                RA, RB, RE = T.stage(g["fct_LO"], g["ttf"], g["fct_adf_h"])
                lo_, tt_ = RA.copy(), RB.copy()
                with np.errstate(invalid="ignore"):
                    RA = np.where(lo_ < tt_, tt_, lo_)           # pick_max(lo, ttf)
                    RB = np.where(tt_ < lo_, tt_, lo_)           # pick_min(lo, ttf)
            else:
                RA, RB, RE = T.stage(g["fct_plus"], g["fct_minus"], g["fct_adf_h"])
            for wi in range(T.n_witems):
                lanes = []
                for vl in range(W):
                    d = int(T.sched[wi, vl])
                    if d == 0xffff:
                        lanes.append(None)
                        continue
                    ln, z0, ghost = d & 0xff, ((d >> 8) & 0x7f) * 2, bool(d >> 15)
                    hx, hy, hz, hw = [int(x) for x in T.hdr[ln]]
                    lanes.append(dict(ln=ln, z0=z0, ghost=ghost, grow=hx + z0, nz=hy & 0xff, fm=(hy >> 8) & 0xff,
                                      sd=(hy >> 16) & 0xff, own=hz // 8, e0=hw & 0xffff, cnt=hw >> 16))
                    if phase == "A" and not ghost:
                        assert (t, ln, z0) not in seen_slots        # every slot is computed exactly once
                        seen_slots.add((t, ln, z0))
                # the stencil neighbours of every real slot are the neighbouring lanes (ghosts included)
                for vl, s in enumerate(lanes):
                    if s is None or s["ghost"]:
                        continue
                    assert s["z0"] < s["nz"]
                    if s["z0"] > 0:
                        q = lanes[vl - 1]
                        assert vl > 0 and q is not None and q["ln"] == s["ln"] and q["z0"] == s["z0"] - 2
                    if s["z0"] + 2 < s["nz"]:
                        q = lanes[vl + 1] if vl + 1 < W else None
                        assert q is not None and q["ln"] == s["ln"] and q["z0"] == s["z0"] + 2
                if phase == "A":
                    phase_a_item(T, lanes, RA, RB, RE, g, dt, eps, big, vlimit)
                else:
                    phase_b_item(T, lanes, RA, RB, RE, g, dt)
            if phase == "A":
                # all active slots of the tile's nodes are scheduled
                for ln in range(T.n_nodes):
                    nz = int(T.hdr[ln][1]) & 0xff
                    assert all((t, ln, z) in seen_slots for z in range(0, nz, 2))
    return g, P


def phase_a_item(T, lanes, RA, RB, RE, g, dt, eps, big, vlimit=1):
    tv = []
    for s in lanes:
        if s is None:
            tv.append(None)
            continue
        z0, nz = s["z0"], s["nz"]
        hi = [(-big if z0 + v >= s["fm"] else -INF) for v in range(2)]
        lw = [(big if z0 + v >= s["fm"] else INF) for v in range(2)]
        fv = g["fct_adf_v"]
        f0, f1 = fv[s["grow"]], fv[s["grow"] + 1]
        f2 = fv[s["grow"] + 2] if z0 + 2 <= nz else 0.0
        for v in range(2):
            if z0 + v < s["sd"]:
                hi[v] = pmax(hi[v], RA[s["own"] + z0 + v])
                lw[v] = pmin(lw[v], RB[s["own"] + z0 + v])
        p = [pmax(0., f0) + pmax(0., -f1), pmax(0., f1) + pmax(0., -f2)]
        mm = [pmin(0., f0) + pmin(0., -f1), pmin(0., f1) + pmin(0., -f2)]
        for k in range(s["cnt"]):
            ex, ey, ez, ew = [int(x) for x in T.ent[s["e0"] + k]]
            dg, second = ez & 0xffff, bool(ez >> 31)
            for v in range(2):
                if z0 + v < dg:
                    x, y, h = RA[ey // 8 + z0 + v], RB[ey // 8 + z0 + v], RE[ex // 8 + z0 + v]
                    assert not (np.isnan(x) or np.isnan(y) or np.isnan(h)), "read of a byte that was never staged"
                    hi[v] = pmax(hi[v], x)
                    lw[v] = pmin(lw[v], y)
                    q = -h if second else h
                    p[v] += pmax(0., q)
                    mm[v] += pmin(0., q)
        s.update(hi=hi, lw=lw, p=p, m=mm)
        tv.append(s)
    for vl, s in enumerate(tv):
        if s is None or s["ghost"]:
            continue
        z0, nz = s["z0"], s["nz"]
        for v in range(2):
            z = z0 + v
            if z >= nz:
                continue
            x, y = s["hi"][v], s["lw"][v]
            if vlimit != 1:
                if 0 < z < nz - 1:
                    own = [RA[s["own"] + z + d] for d in (-1, 0, 1)]
                    assert not any(np.isnan(a) for a in own), "own row level that was never staged"
                    vmax = pmax(pmax(own[0], own[1]), own[2])
                    vmin = pmin(pmin(own[0], own[1]), own[2])
                    x, y = (pmax(x, vmax), pmin(y, vmin)) if vlimit == 2 else (pmin(x, vmax), pmax(y, vmin))
            elif 0 < z < nz - 1:
                if v == 0:
                    pv = tv[vl - 1]
                    x = pmax(pmax(pv["hi"][1], x), s["hi"][1])
                    y = pmin(pmin(pv["lw"][1], y), s["lw"][1])
                else:
                    nv = tv[vl + 1]
                    x = pmax(pmax(s["hi"][0], x), nv["hi"][0])
                    y = pmin(pmin(s["lw"][0], y), nv["lw"][0])
            o = s["grow"] + v
            l, ai = g["fct_LO"][o], g["area_inv"][o]
            bm, bn = x - l, y - l
            with np.errstate(all="ignore"):
                pf = pmin(1., float(np.float64(bm) / np.float64(s["p"][v] * dt * ai + eps)))
                mf = pmin(1., float(np.float64(bn) / np.float64(s["m"][v] * dt * ai - eps)))
            g["fct_ttf_max"][o], g["fct_ttf_min"][o], g["fct_plus"][o], g["fct_minus"][o] = bm, bn, pf, mf


def phase_b_item(T, lanes, RA, RB, RE, g, dt):
    for s in lanes:
        if s is None or s["ghost"]:
            continue
        z0, nz, own = s["z0"], s["nz"], s["own"]
        fv = g["fct_adf_v"]

        def lim(z):
            fz = fv[s["grow"] - z0 + z]
            if z >= nz:
                return fz
            ae = 1.
            if z == 0:
                ae = pmin(ae, RA[own] if fz >= 0. else RB[own])
            elif fz >= 0.:
                ae = pmin(pmin(ae, RB[own + z - 1]), RA[own + z])
            else:
                ae = pmin(pmin(ae, RA[own + z - 1]), RB[own + z])
            return ae * fz
        fl = [lim(z0), lim(z0 + 1), lim(z0 + 2) if z0 + 2 <= nz else 0.]
        out = {}
        for v in range(2):
            if z0 + v >= nz:
                continue
            o = s["grow"] + v
            ar = dt / g["area"][o]
            dv = (g["del_ttf_advvert"][o] - g["ttf"][o] * g["hnode"][o] + g["fct_LO"][o] * g["hnode_new"][o]
                  + (fl[v] - fl[v + 1]) * ar)
            out[v] = [ar, dv, g["del_ttf_advhoriz"][o]]
        for k in range(s["cnt"]):
            ex, ey, ez, ew = [int(x) for x in T.ent[s["e0"] + k]]
            dg, second, writer = ez & 0xffff, bool(ez >> 31), bool((ez >> 30) & 1)
            for v in range(2):
                z = z0 + v
                if z >= dg:
                    continue
                h, po, mo = RE[ex // 8 + z], RA[ey // 8 + z], RB[ey // 8 + z]
                pn, mn = RA[own + z], RB[own + z]
                assert not (np.isnan(h) or np.isnan(po) or np.isnan(mo) or np.isnan(pn) or np.isnan(mn))
                p1, m1, p2, m2 = (po, mo, pn, mn) if second else (pn, mn, po, mo)
                ae = pmin(pmin(1., p1), m2) if h >= 0. else pmin(pmin(1., m1), p2)
                hl = ae * h
                x = hl * out[v][0]
                out[v][2] = out[v][2] - x if second else out[v][2] + x
                if writer:
                    g["adf_h_out"][ew + z] = hl
        for v, (ar, dv, dh) in out.items():
            o = s["grow"] + v
            g["adf_v_out"][o] = fl[v]
            g["del_ttf_advvert"][o] = dv
            g["del_ttf_advhoriz"][o] = dh


def unpad(a, P, width):
    return a.reshape(-1, P)[:, :width]


def compare_packed(m, f, g, pk, want):
    """Active cells from the packed arrays, everything else as the inputs had it (what the harness does)."""
    for k, wk in (("fct_ttf_max", "fct_ttf_max"), ("fct_ttf_min", "fct_ttf_min"), ("fct_plus", "fct_plus"),
                  ("fct_minus", "fct_minus"), ("adf_v_out", "fct_adf_v"), ("del_ttf_advvert", "del_ttf_advvert"),
                  ("del_ttf_advhoriz", "del_ttf_advhoriz"), ("adf_h_out", "fct_adf_h")):
        edge = wk == "fct_adf_h"
        got = pk.unpack(g[k], getattr(f, wk), edge=edge)
        w = getattr(want, wk)
        z = np.arange(w.shape[1])[None, :]
        depth = m.edge_depth() if edge else m.nlevels_nod2D - 1
        act = z < depth[:, None]
        assert bits_equal(np.where(act, got, w), w), k


def compare(m, f, g, P, want, owned=None):
    L = m.L
    pairs = [("fct_ttf_max", "fct_ttf_max", L), ("fct_ttf_min", "fct_ttf_min", L), ("fct_plus", "fct_plus", L),
             ("fct_minus", "fct_minus", L), ("adf_v_out", "fct_adf_v", m.nl), ("del_ttf_advvert", "del_ttf_advvert", L),
             ("del_ttf_advhoriz", "del_ttf_advhoriz", L)]
    for k, wk, w in pairs:
        a, b = unpad(g[k], P, w), getattr(want, wk)
        if owned is not None:
            a, b = a[:owned], b[:owned]
        assert bits_equal(a, b), k
    if owned is None:
        assert bits_equal(unpad(g["adf_h_out"], P, L), want.fct_adf_h)


@pytest.mark.parametrize("name,tn,cap", [("tiny", 64, 74 * 1024), ("tiny", 3, 74 * 1024),
                                         ("pi", 96, 74 * 1024), ("pi", 24, 24 * 1024)])
def test_tables_reproduce_the_oracle(mesh_mod, abi, oracle_mod, name, tn, cap):
    m = mesh_mod.make_workload(name)
    f = mesh_mod.make_fields(m)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st, nt, smem, blob, off = inspect(abi, m, tn, cap)
    assert st == 0 and nt >= 1 and smem <= cap
    g, P = emulate(m, f, blob, off, nt)
    compare(m, f, g, P, want)


@pytest.mark.parametrize("vlimit", [2, 3])
@pytest.mark.parametrize("packed", [0, 1])
def test_tables_serve_the_vlimit_variant(mesh_mod, abi, oracle_mod, vlimit, packed):
    """vlimit 2 / 3 on the fused path need nothing the tables do not already stage: the own row holds
    every level the vertical neighbourhood reads (docs/refactoring.md:113-148)."""
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    f.vlimit = vlimit
    want = f.copy()
    oracle_mod.fct_ale_general(m, want)
    st, nt, smem, blob, off = inspect(abi, m, 96, 74 * 1024, packed=packed)
    assert st == 0
    if packed:
        pk = Packed(m)
        g, P = emulate(m, f, blob, off, nt, packed=pk, vlimit=vlimit)
        compare_packed(m, f, g, pk, want)
    else:
        g, P = emulate(m, f, blob, off, nt, vlimit=vlimit)
        compare(m, f, g, P, want)


@pytest.mark.parametrize("nx,ny,nl,seed,tn,cap", [(23, 17, 13, 5, 11, 20 * 1024), (31, 9, 30, 6, 64, 74 * 1024),
                                                  (12, 40, 6, 7, 200, 111 * 1024)])
def test_random_meshes_packed(mesh_mod, abi, oracle_mod, nx, ny, nl, seed, tn, cap):
    """Other shapes of the generator (aspect ratios, depths from 6 to 30 levels, land masks of other
    seeds) through the packed tables at three tile sizes."""
    m = mesh_mod.make_mesh(nx, ny, nl, seed=seed)
    f = mesh_mod.make_fields(m, seed=seed + 1)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st, nt, smem, blob, off = inspect(abi, m, tn, cap, packed=1)
    assert st == 0 and smem <= cap
    pk = Packed(m)
    g, P = emulate(m, f, blob, off, nt, packed=pk)
    compare_packed(m, f, g, pk, want)


TWO_STAGE_CAP = ((227 * 1024 - 256 - 4 * 1024) // 2) & ~127     # one stage of the two-stage ring (fct_driver.cu)


def test_two_stage_tiles_of_deep_columns(mesh_mod, abi, oracle_mod):
    """nl = 80 columns in the tiles of the two-stage ring (the default for deep packed meshes): larger
    tiles, still exact, and every copy list fits the 1 KB slot that travels ahead of its blob
    (WT_PRE_MAX_COPIES = 120 in fct_warp_kernels.cuh)."""
    m = mesh_mod.make_mesh(48, 37, 80, seed=3)
    f = mesh_mod.make_fields(m, seed=4)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st3, nt3, _, _, _ = inspect(abi, m, 96, 74 * 1024, packed=1)
    st, nt, smem, blob, off = inspect(abi, m, 96, TWO_STAGE_CAP, packed=1)
    assert st == 0 and st3 == 0 and smem <= TWO_STAGE_CAP and nt < 0.8 * nt3
    pk = Packed(m)
    g, P = emulate(m, f, blob, off, nt, packed=pk)
    compare_packed(m, f, g, pk, want)
    P = (m.nl + 7) & ~7
    assert max(Tile(blob[off[t] * 4: off[t + 1] * 4], P).n_copies for t in range(nt)) <= 120


@pytest.mark.parametrize("name,tn,cap", [("tiny", 64, 74 * 1024), ("pi", 96, 74 * 1024), ("pi", 7, 30 * 1024)])
def test_packed_level_storage(mesh_mod, abi, oracle_mod, name, tn, cap):
    """Columns back to back (active levels only): the tile's own columns and its edge rows in ascending
    id become runs that travel as single bulk copies."""
    m = mesh_mod.make_workload(name)
    f = mesh_mod.make_fields(m)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st, nt, smem, blob, off = inspect(abi, m, tn, cap, packed=1)
    assert st == 0 and nt >= 1 and smem <= cap
    pk = Packed(m)
    g, P = emulate(m, f, blob, off, nt, packed=pk)
    compare_packed(m, f, g, pk, want)
    # far fewer copies than staged rows
    P = (m.nl + 7) & ~7
    ncopies = sum(Tile(blob[off[t] * 4: off[t + 1] * 4], P).n_copies for t in range(nt))
    nrows = sum(2 * Tile(blob[off[t] * 4: off[t + 1] * 4], P).n_rows + Tile(blob[off[t] * 4: off[t + 1] * 4], P).n_erows
                for t in range(nt))
    assert ncopies < 0.6 * nrows


def test_deep_columns_are_cut_with_ghost_slots(mesh_mod, abi, oracle_mod):
    """nl = 80: a full column has 40 level pairs, more than a 32-lane item -> cut, ghosts on both sides."""
    m = mesh_mod.make_mesh(12, 9, 80, seed=3)
    f = mesh_mod.make_fields(m, seed=4)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st, nt, smem, blob, off = inspect(abi, m, 64, 74 * 1024)
    assert st == 0 and smem <= 74 * 1024
    ghosts = 0
    P = (m.nl + 7) & ~7
    for t in range(nt):
        T = Tile(blob[off[t] * 4: off[t + 1] * 4], P)
        live = T.sched[T.sched != 0xffff]
        ghosts += int((live >> 15).sum())
    assert ghosts > 0
    g, P = emulate(m, f, blob, off, nt)
    compare(m, f, g, P, want)


def test_non_triangulation_is_refused(mesh_mod, abi):
    m, _ = mesh_mod.adversarial_case(120, 17, seed=5)
    assert inspect(abi, m)[0] == 2


def test_partitioned_tables(mesh_mod, abi, oracle_mod):
    """Boundary + interior tile sets of every partition, halo exchange emulated: owned results equal
    the single-domain oracle (writer flags, halo rows as neighbours, cut edges)."""
    m = mesh_mod.make_workload("tiny")
    f = mesh_mod.make_fields(m)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    parts = mesh_mod.partition_mesh(m, 3)
    lfs = [mesh_mod.slice_fields(f, p) for p in parts]
    # phase A on every rank, exchange of fct_plus / fct_minus, phase B: run the emulator twice per
    # rank (the second time with the exchanged factors) and keep phase B of the second run
    gplus = np.array(f.fct_plus)
    gminus = np.array(f.fct_minus)
    runs = []
    for p, lf in zip(parts, lfs):
        tabs = [inspect(abi, p.mesh, 16, 74 * 1024, which) for which in (1, 2)]
        assert all(t[0] == 0 for t in tabs)
        runs.append(tabs)
        for st, nt, smem, blob, off in tabs:
            g, P = emulate(p.mesh, lf, blob, off, nt)
            n = p.mesh.myDim_nod2D
            own = np.zeros(p.mesh.nnod, bool)
            # rows this tile set wrote = nodes whose fct_plus changed from the input
            pl, mi = unpad(g["fct_plus"], P, m.L), unpad(g["fct_minus"], P, m.L)
            ch = (pl != lf.fct_plus).any(1) | (mi != lf.fct_minus).any(1)
            own[:n] = ch[:n]
            gplus[p.mesh.node_gid[own]] = pl[own]
            gminus[p.mesh.node_gid[own]] = mi[own]
    assert bits_equal(gplus, want.fct_plus) and bits_equal(gminus, want.fct_minus)


def test_an_edge_between_two_halo_nodes_is_refused(abi, mesh_mod):
    """myDim_edge2D holds the edges that touch an owned node; an edge joining two halo
    nodes would have no node to store its limited flux in the fused phase B (the staged b3h limits
    every edge): the inspector refuses such a mesh instead of returning stale fluxes for it."""
    gm = mesh_mod.make_workload("pi")
    part = mesh_mod.partition_mesh(gm, 3, ranks=[1])[0]
    m = part.mesh
    assert inspect(abi, m, packed=1)[0] == 0
    n = m.myDim_nod2D
    e = m.edges.copy()
    # re-point one local edge at two halo nodes
    e[0] = (n + 1, n + 2)
    bad = mesh_mod.Mesh(**{**m.__dict__, "edges": np.ascontiguousarray(e)})
    assert inspect(abi, bad, packed=1)[0] == 1


@pytest.mark.parametrize("tn,cap", [(96, 74 * 1024), (9, 24 * 1024)])
def test_unstructured_delaunay_mesh_packed(mesh_mod, abi, oracle_mod, tn, cap):
    """An UNSTRUCTURED triangulation (Delaunay of jittered-random points: node degrees 2..11 instead of
    the 4 / 8 of the criss-cross grids, ragged edge lists, land holes) through the packed tables."""
    m = mesh_mod.make_delaunay_mesh(700, 21, seed=3)
    deg = np.bincount((m.edges - 1).ravel(), minlength=m.myDim_nod2D)
    assert deg.min() <= 3 and deg.max() >= 9
    f = mesh_mod.make_fields(m, seed=8)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    st, nt, smem, blob, off = inspect(abi, m, tn, cap, packed=1)
    assert st == 0 and smem <= cap
    pk = Packed(m)
    g, P = emulate(m, f, blob, off, nt, packed=pk)
    compare_packed(m, f, g, pk, want)


@pytest.mark.parametrize("packed", [0, 1])
def test_l2_prefetch_entries_cover_the_own_columns(mesh_mod, abi, packed):
    """Behind its bulk copies every tile's blob lists the runs of its OWN columns (array code 3, no shared-memory
    offset): what the issuer warps pull into L2 for the arrays the consumers load straight from global memory.
    The runs must be 16-byte granular, disjoint, inside the tile's columns and together cover them (up to the cap
    of eight runs per tile) -- for the contiguous tiles of a single domain and the scattered ones of a boundary set."""
    gm = mesh_mod.make_workload("pi")
    for m, which in ((gm, 0), (mesh_mod.partition_mesh(gm, 3, ranks=[1])[0].mesh, 1), (mesh_mod.partition_mesh(gm, 3, ranks=[1])[0].mesh, 2)):
        st, nt, smem, blob, off = inspect(abi, m, 40, 74 * 1024, which=which, packed=packed)
        assert st == 0 and nt > 0
        P = (m.nl + 7) & ~7
        nz = np.maximum(m.nlevels_nod2D.astype(np.int64) - 1, 0)
        slots = ((nz + 2) & ~1) if packed else np.full(nz.shape, P)
        ncol = np.concatenate([[0], np.cumsum(slots)])
        covered_all = 0
        for t in range(nt):
            b = blob[off[t] * 4: off[t + 1] * 4].tobytes()
            h = np.frombuffer(b, np.int32, 16)
            n_copies, n_nodes, n_pf, off_hdr = int(h[0]), int(h[2]), int(h[12]), int(h[5])
            assert 1 <= n_pf <= 8
            ent = np.frombuffer(b, np.uint32, 2 * (n_copies + n_pf), 64).reshape(-1, 2)[n_copies:]
            hdr = np.frombuffer(b, np.uint32, 4 * n_nodes, off_hdr).reshape(-1, 4)
            own0 = hdr[:, 0].astype(np.int64)                                  # global element offset of every own column
            node = np.searchsorted(ncol, own0, side="right") - 1
            assert (ncol[node] == own0).all() and (node < m.myDim_nod2D).all()
            cells = np.zeros(int(ncol[-1]) + P, bool)
            for g, pk in ent:
                pk = int(pk)
                assert (pk >> 28) & 3 == 3 and (pk & 0x3fff) == 0
                n = ((pk >> 14) & 0x3fff) * 2                                   # doubles
                assert n > 0 and int(g) % 2 == 0 and not cells[int(g): int(g) + n].any()
                cells[int(g): int(g) + n] = True
            own = np.zeros_like(cells)
            for k in node:
                own[ncol[k]: ncol[k + 1]] = True
            assert not (cells & ~own).any()                                     # nothing but own columns
            if n_pf < 8:
                assert (cells == own).all()                                     # all of them, unless the cap cut the list
            covered_all += int(n_pf < 8)
        assert covered_all >= (nt if which == 0 else 1)
