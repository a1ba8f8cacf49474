"""CPU: pin the oracle (oracle/fct_ale_oracle.c) against the reference's own code.

* golden vectors produced by src/reference.cpp (a1..a4) and by the reference's numpy reference()
  functions (b3 / c) -- tests/golden/make_golden.py;
* live comparison with oracle/_ref/libref.so (the reference compiled unmodified) when present.
"""
import os

import numpy as np
import pytest

from conftest import bits_equal, load_golden, rel_err


@pytest.mark.parametrize("case", ["ref_cpp_tiny", "ref_cpp_adversarial"])
def test_oracle_matches_reference_cpp_golden(oracle_mod, case):
    m, f, z = load_golden(case)
    g = f.copy()
    oracle_mod.a1(m, g)
    assert bits_equal(g.fct_ttf_max, z["a1_fct_ttf_max"])
    assert bits_equal(g.fct_ttf_min, z["a1_fct_ttf_min"])
    oracle_mod.a2(m, g)
    assert bits_equal(g.UV_rhs, z["a2_UV_rhs"])
    oracle_mod.a3(m, g)
    oracle_mod.b1_vertical(m, g)
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus"):
        assert bits_equal(getattr(g, k), z["a3_" + k]), k
    oracle_mod.b1_horizontal(m, g)
    oracle_mod.b2(m, g)
    assert bits_equal(g.fct_plus, z["a4_fct_plus"])
    assert bits_equal(g.fct_minus, z["a4_fct_minus"])


def test_oracle_b3_c_match_reference_numpy_golden(oracle_mod):
    """b3 and c bit-exact against the reference's numpy twins: the oracle groups the flux term
    x*(dt/area) like kernels/fct_ale_c_*.cu and the twins do.  (The Fortran listing's (x*dt)/area
    differs by a few ulp; checked below to stay inside the 1e-12 relative bar.)"""
    m, f, z = load_golden("ref_numpy_tiny")
    g = f.copy()      # state after the reference's a1..a4
    oracle_mod.b3_vertical(m, g)
    assert bits_equal(g.fct_adf_v, z["b3v_fct_adf_v"])
    oracle_mod.b3_horizontal(m, g)
    assert bits_equal(g.fct_adf_h, z["b3h_fct_adf_h"])
    oracle_mod.c_vertical(m, g)
    oracle_mod.c_horizontal(m, g)
    assert bits_equal(g.del_ttf_advvert, z["cv_del_ttf_advvert"])
    assert bits_equal(g.del_ttf_advhoriz, z["ch_del_ttf_advhoriz"])
    # Fortran grouping (docs/refactoring.md:297-298, :311-312) in numpy, same inputs
    L, N = m.L, m.myDim_nod2D
    act = np.arange(L)[None, :] < (m.nlevels_nod2D[:N, None] - 1)
    v = z["b3v_fct_adf_v"]
    fort = f.del_ttf_advvert - f.ttf * f.hnode + f.fct_LO * f.hnode_new + (v[:, :L] - v[:, 1:]) * f.dt / f.area[:, :L]
    assert rel_err(np.where(act, fort, 0), np.where(act, g.del_ttf_advvert, 0), floor=1.0) < 1e-12
    # something actually moved
    assert not bits_equal(g.del_ttf_advvert, f.del_ttf_advvert)
    assert not bits_equal(g.fct_adf_h, f.fct_adf_h)


@pytest.mark.parametrize("name", ["tiny", "pi"])
def test_oracle_matches_libref_live(oracle_mod, mesh_mod, name):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    m = mesh_mod.make_workload(name)
    f = mesh_mod.make_fields(m)
    a, b = f.copy(), f.copy()
    oracle_mod.pre_comm(m, a)
    oracle_mod.ref_pre_comm(m, b)
    for k in ("fct_ttf_max", "fct_ttf_min", "UV_rhs", "fct_plus", "fct_minus"):
        assert bits_equal(getattr(a, k), getattr(b, k)), k


def test_oracle_matches_libref_adversarial(oracle_mod, mesh_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    for seed in range(4):
        m, f = mesh_mod.adversarial_case(200, 20, seed=seed)
        a, b = f.copy(), f.copy()
        oracle_mod.pre_comm(m, a)
        oracle_mod.ref_pre_comm(m, b)
        for k in ("fct_ttf_max", "fct_ttf_min", "UV_rhs", "fct_plus", "fct_minus"):
            assert bits_equal(getattr(a, k), getattr(b, k)), (seed, k)


def test_limiter_properties(oracle_mod, mesh_mod):
    """Domain properties the GPU tests reuse at full size: factors in (-inf, 1], limited fluxes never
    grow and keep their sign, bottom vertical flux untouched, inactive cells untouched."""
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    g = f.copy()
    oracle_mod.fct_ale(m, g)
    L = m.L
    act = np.arange(L)[None, :] < (m.nlevels_nod2D[:, None] - 1)
    assert g.fct_plus[act].max() <= 1.0 and g.fct_minus[act].max() <= 1.0
    assert (g.fct_plus[act] >= 0).all() and (g.fct_minus[act] >= 0).all()
    assert (np.abs(g.fct_adf_v) <= np.abs(f.fct_adf_v)).all()
    assert (np.abs(g.fct_adf_h) <= np.abs(f.fct_adf_h)).all()
    assert (g.fct_adf_v * f.fct_adf_v >= 0).all()
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz"):
        assert bits_equal(getattr(g, k)[~act], getattr(f, k)[~act]), k


# ---- vlimit 2 / 3 and iter_yn (parity unpinned by the reference) -------
def listing_a3_vlimit(m, f, vlimit):
    """Second, independent restatement of docs/refactoring.md:113-148 in plain Python loops (1-based
    like the listing), used to cross-check the C oracle on small meshes."""
    tmax, tmin = f.fct_ttf_max.copy(), f.fct_ttf_min.copy()
    uv = f.UV_rhs.reshape(m.myDim_elem2D, m.L, 2)
    for n in range(m.myDim_nod2D):
        nlev = int(m.nlevels_nod2D[n])
        ring = m.nod_in_elem2D[n, :m.nod_in_elem2D_num[n]] - 1
        tv_max = {nz: max(uv[e, nz - 1, 0] for e in ring) for nz in range(1, nlev)}
        tv_min = {nz: min(uv[e, nz - 1, 1] for e in ring) for nz in range(1, nlev)}
        for nz in range(2, nlev - 1):
            col = f.fct_ttf_max[n, nz - 2:nz + 1]       # fct_ttf_max(nz-1:nz+1, n), a1 values
            if vlimit == 2:
                tv_max[nz] = max(tv_max[nz], col.max())
                tv_min[nz] = min(tv_min[nz], col.min())
            else:
                tv_max[nz] = min(tv_max[nz], col.max())
                tv_min[nz] = max(tv_min[nz], col.min())
        for nz in range(1, nlev):
            tmax[n, nz - 1] = tv_max[nz] - f.fct_LO[n, nz - 1]
            tmin[n, nz - 1] = tv_min[nz] - f.fct_LO[n, nz - 1]
    return tmax, tmin


@pytest.mark.parametrize("vlimit", [2, 3])
def test_oracle_vlimit_matches_listing(oracle_mod, mesh_mod, vlimit):
    for m, f in (mesh_mod.adversarial_case(60, 12, seed=3), (mesh_mod.make_workload("tiny"), None)):
        if f is None:
            f = mesh_mod.make_fields(m)
        f.vlimit = vlimit
        oracle_mod.a1(m, f)
        oracle_mod.a2(m, f)
        want_max, want_min = listing_a3_vlimit(m, f, vlimit)
        wide = f.copy()
        wide.vlimit = 1
        oracle_mod.a3_vlimit(m, f)
        assert bits_equal(f.fct_ttf_max, want_max) and bits_equal(f.fct_ttf_min, want_min)
        oracle_mod.a3(m, wide)
        assert not bits_equal(f.fct_ttf_max, wide.fct_ttf_max)      # the branches differ from vlimit 1


def test_oracle_iterative_branch(oracle_mod, mesh_mod):
    """iter_yn: limited + rejected parts restore the flux, the surface and bottom vertical levels of
    fct_adf_v2 are never written (md:228-230), the low-order update follows the listing's order."""
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    f.iter_yn = True
    rng = np.random.default_rng(5)
    f.fct_adf_v2 = rng.standard_normal(f.fct_adf_v.shape)
    f.fct_adf_h2 = rng.standard_normal(f.fct_adf_h.shape)
    g = f.copy()
    g.iter_yn = False
    oracle_mod.pre_comm(m, g)
    lim = g.copy()
    oracle_mod.b3_vertical(m, lim)
    oracle_mod.b3_horizontal(m, lim)
    h = g.copy()
    oracle_mod.b3_vertical_iter(m, h)
    oracle_mod.b3_horizontal_iter(m, h)
    assert bits_equal(h.fct_adf_v, lim.fct_adf_v) and bits_equal(h.fct_adf_h, lim.fct_adf_h)
    N, L = m.myDim_nod2D, m.L
    z = np.arange(m.nl)[None, :]
    inner = (z >= 1) & (z < (m.nlevels_nod2D[:N, None] - 1))
    assert rel_err((h.fct_adf_v + h.fct_adf_v2)[inner], g.fct_adf_v[inner], floor=1e-3) < 1e-14
    assert bits_equal(h.fct_adf_v2[~inner], f.fct_adf_v2[~inner])
    eact = np.arange(L)[None, :] < m.edge_depth()[:, None]
    assert rel_err((h.fct_adf_h + h.fct_adf_h2)[eact], g.fct_adf_h[eact], floor=1e-3) < 1e-14
    # low-order update in numpy, same order of operations (vertical, then edges ascending)
    lo = h.fct_LO.copy()
    act = np.arange(L)[None, :] < (m.nlevels_nod2D[:N, None] - 1)
    v = h.fct_adf_v
    lo[:N] = np.where(act, lo[:N] + (v[:N, :L] - v[:N, 1:]) * f.dt / f.area[:N, :L] / f.hnode_new[:N], lo[:N])
    for g_ in range(m.myDim_edge2D):
        n1, n2 = m.edges[g_] - 1
        d = int(m.edge_depth()[g_])
        lo[n1, :d] = lo[n1, :d] + h.fct_adf_h[g_, :d] * f.dt / f.area[n1, :d] / f.hnode_new[n1, :d]
        lo[n2, :d] = lo[n2, :d] - h.fct_adf_h[g_, :d] * f.dt / f.area[n2, :d] / f.hnode_new[n2, :d]
    oracle_mod.lo_update(m, h)
    assert bits_equal(h.fct_LO, lo)
    # the composed subroutine ends with fct_adf_* = fct_adf_*2
    full = f.copy()
    oracle_mod.fct_ale_general(m, full)
    assert bits_equal(full.fct_adf_v, h.fct_adf_v2) and bits_equal(full.fct_adf_h, h.fct_adf_h2)
    assert bits_equal(full.fct_LO, lo)


# ---- stress2rhs: pinned to the reference's compiled function ----------
def test_oracle_stress2rhs_matches_golden(oracle_mod):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cpp_stress2rhs.npz"))
    d = {k: z[k] for k in z.files}
    d["N"], d["E"] = int(z["N"]), int(z["E"])
    u, v = oracle_mod.stress2rhs(d)
    assert bits_equal(u, z["U_rhs_ice"]) and bits_equal(v, z["V_rhs_ice"])
    assert np.count_nonzero(u) > d["N"] // 2


def test_oracle_stress2rhs_matches_libref_live(oracle_mod, mesh_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    m = mesh_mod.make_workload("pi")
    tri = np.ascontiguousarray((m.elem2D_nodes - 1).T)          # 0-based [3][E], reference.cpp:457
    for d in (oracle_mod.stress_case(500, 1300, seed=1), oracle_mod.stress_case(m.myDim_nod2D, m.myDim_elem2D, seed=2, elem_nodes=tri)):
        a, b = oracle_mod.stress2rhs(d), oracle_mod.ref_stress2rhs(d)
        assert bits_equal(a[0], b[0]) and bits_equal(a[1], b[1])
