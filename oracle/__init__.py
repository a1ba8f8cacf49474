"""TEST INFRASTRUCTURE -- the CPU checker for the fct_ale path.

ctypes front-end for
  * oracle/libfct_oracle.so : our plain-C restatement (oracle/fct_ale_oracle.c), and
  * oracle/_ref/libref.so   : the reference's own src/reference.cpp compiled unmodified
                              (stages a1, a2, a3+b1v, a4=b1h+b2 only; reference.cpp:306-438).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (fesom2-accelerate_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfct_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> None:
    """Compile the restatement and, when /root/reference is present, the reference's own file."""
    if force or not os.path.exists(_LIB) or os.path.exists("/root/reference/src/reference.cpp"):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF)
    return _ref


# ------------------------------------------------------------------ restatement, stage by stage
def a1(m, f, n_nodes=None):
    n = m.nnod if n_nodes is None else n_nodes
    lib().oracle_a1(C.c_int(n), _i(m.nlevels_nod2D), C.c_int(m.nl), _d(f.fct_ttf_max),
                    _d(f.fct_ttf_min), _d(f.fct_LO), _d(f.ttf))


def a2(m, f):
    lib().oracle_a2(C.c_int(m.myDim_elem2D), _i(m.nlevels_elem), C.c_int(m.nl), _d(f.UV_rhs),
                    _i(m.elem2D_nodes), _d(f.fct_ttf_max), _d(f.fct_ttf_min), C.c_double(f.bignumber))


def a3(m, f):
    scratch = np.empty(2 * m.L)
    lib().oracle_a3(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl), _d(f.fct_ttf_max),
                    _d(f.fct_ttf_min), _d(f.fct_LO), _d(f.UV_rhs), _i(m.nod_in_elem2D),
                    _i(m.nod_in_elem2D_num), C.c_int(m.nod_in_elem2D_dim), _d(scratch))


def b1_vertical(m, f):
    lib().oracle_b1_vertical(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl),
                             _d(f.fct_plus), _d(f.fct_minus), _d(f.fct_adf_v))


def b1_horizontal(m, f):
    lib().oracle_b1_horizontal(C.c_int(m.myDim_edge2D), C.c_int(m.nl), _i(m.nlevels_elem),
                               _i(m.edges), _i(m.edge_tri), _d(f.fct_adf_h), _d(f.fct_plus),
                               _d(f.fct_minus))


def b2(m, f):
    lib().oracle_b2(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl), _d(f.fct_plus),
                    _d(f.fct_minus), _d(f.fct_ttf_max), _d(f.fct_ttf_min), _d(f.area_inv),
                    C.c_double(f.dt), C.c_double(f.flux_eps))


def b3_vertical(m, f):
    lib().oracle_b3_vertical(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl),
                             _d(f.fct_adf_v), _d(f.fct_plus), _d(f.fct_minus))


def b3_horizontal(m, f):
    lib().oracle_b3_horizontal(C.c_int(m.myDim_edge2D), C.c_int(m.nl), _i(m.nlevels_elem),
                               _i(m.edges), _i(m.edge_tri), _d(f.fct_adf_h), _d(f.fct_plus),
                               _d(f.fct_minus))


def c_vertical(m, f):
    lib().oracle_c_vertical(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl),
                            _d(f.del_ttf_advvert), _d(f.ttf), _d(f.hnode), _d(f.fct_LO),
                            _d(f.hnode_new), _d(f.fct_adf_v), _d(f.area), C.c_double(f.dt))


def c_horizontal(m, f):
    lib().oracle_c_horizontal(C.c_int(m.myDim_edge2D), C.c_int(m.nl), _i(m.nlevels_elem),
                              _i(m.edges), _i(m.edge_tri), _d(f.fct_adf_h), _d(f.area),
                              _d(f.del_ttf_advhoriz), C.c_double(f.dt))


STAGES = [("a1", a1), ("a2", a2), ("a3", a3), ("b1v", b1_vertical), ("b1h", b1_horizontal),
          ("b2", b2), ("b3v", b3_vertical), ("b3h", b3_horizontal), ("cv", c_vertical),
          ("ch", c_horizontal)]


def pre_comm(m, f):
    """a1 .. b2 (reference.cpp:289-304)."""
    for _, fn in STAGES[:6]:
        fn(m, f)


def post_comm(m, f):
    """b3 .. c (docs/refactoring.md:204-314, iter_yn = .false.)."""
    for _, fn in STAGES[6:]:
        fn(m, f)


def fct_ale(m, f, exchange=None):
    """Whole chain on one domain; `exchange(f)` stands for exchange_nod(fct_plus, fct_minus)."""
    pre_comm(m, f)
    if exchange is not None:
        exchange(f)
    post_comm(m, f)


# ------------------------------------------------------------------ vlimit 2 / 3 and iter_yn
# SURVEY.md section 8(f) row 2.  PARITY UNPINNED: the reference holds no executable form of these
# branches (reference.cpp:51-96 TODO stubs, kernels/fct_ale_a3.py:152-155 `pass`); the restatement
# follows docs/refactoring.md:113-148 and :226-290 line by line.
def a3_vlimit(m, f):
    if f.vlimit == 1:
        return a3(m, f)
    scratch = np.empty(2 * m.L)
    lib().oracle_a3_vlimit(C.c_int(f.vlimit), C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl),
                           _d(f.fct_ttf_max), _d(f.fct_ttf_min), _d(f.fct_LO), _d(f.UV_rhs),
                           _i(m.nod_in_elem2D), _i(m.nod_in_elem2D_num), C.c_int(m.nod_in_elem2D_dim),
                           _d(scratch))


def b3_vertical_iter(m, f):
    lib().oracle_b3_vertical_iter(C.c_int(m.myDim_nod2D), _i(m.nlevels_nod2D), C.c_int(m.nl),
                                  _d(f.fct_adf_v), _d(f.fct_adf_v2), _d(f.fct_plus), _d(f.fct_minus))


def b3_horizontal_iter(m, f):
    lib().oracle_b3_horizontal_iter(C.c_int(m.myDim_edge2D), C.c_int(m.nl), _i(m.nlevels_elem),
                                    _i(m.edges), _i(m.edge_tri), _d(f.fct_adf_h), _d(f.fct_adf_h2),
                                    _d(f.fct_plus), _d(f.fct_minus))


def lo_update(m, f):
    lib().oracle_lo_update(C.c_int(m.myDim_nod2D), C.c_int(m.myDim_edge2D), C.c_int(m.nl),
                           _i(m.nlevels_nod2D), _i(m.nlevels_elem), _i(m.edges), _i(m.edge_tri),
                           _d(f.fct_LO), _d(f.fct_adf_v), _d(f.fct_adf_h), _d(f.area), _d(f.hnode_new),
                           C.c_double(f.dt))


def fct_ale_general(m, f, exchange=None):
    """The whole subroutine of docs/refactoring.md:13-315 with its vlimit and iter_yn branches.
    iter_yn: ends after the low-order update with fct_adf_* = fct_adf_*2 (md:288-290)."""
    a1(m, f)
    a2(m, f)
    a3_vlimit(m, f)
    b1_vertical(m, f)
    b1_horizontal(m, f)
    b2(m, f)
    if exchange is not None:
        exchange(f)
    if not f.iter_yn:
        return post_comm(m, f)
    b3_vertical_iter(m, f)
    b3_horizontal_iter(m, f)
    lo_update(m, f)
    f.fct_adf_h[...] = f.fct_adf_h2
    f.fct_adf_v[...] = f.fct_adf_v2


# ------------------------------------------------------------------ stress2rhs
STRESS_KEYS = ("ice_strength", "elem_area", "sigma11", "sigma12", "sigma22", "gradient_sca", "metric_factor",
               "inv_areamass", "rhs_a", "rhs_m")


def stress_case(n_nodes, n_elems, seed=0, elem_nodes=None):
    """Seeded inputs of stress2rhs in the layout src/reference.cpp:440-480 indexes: 0-based
    elem2D_nodes[3][E], gradient_sca with 6*E entries (the reference reads up to index 29+E)."""
    rng = np.random.default_rng(seed)
    d = {"N": n_nodes, "E": n_elems}
    d["elem2D_nodes"] = (rng.integers(0, n_nodes, (3, n_elems)) if elem_nodes is None else elem_nodes).astype(np.int32)
    d["ice_strength"] = np.where(rng.random(n_elems) < 0.8, rng.random(n_elems) * 1e4, 0.0)
    d["elem_area"] = rng.uniform(1e7, 1e9, n_elems)
    for k in ("sigma11", "sigma12", "sigma22"):
        d[k] = rng.standard_normal(n_elems) * 1e3
    d["gradient_sca"] = rng.standard_normal(6 * max(n_elems, 6)) * 1e-5
    d["metric_factor"] = rng.standard_normal(n_elems) * 1e-7
    d["inv_areamass"] = np.where(rng.random(n_nodes) < 0.9, rng.uniform(1e-12, 1e-9, n_nodes), 0.0)
    d["rhs_a"] = rng.standard_normal(n_nodes)
    d["rhs_m"] = rng.standard_normal(n_nodes)
    return d


def stress2rhs(d):
    u, v = np.full(d["N"], np.nan), np.full(d["N"], np.nan)
    lib().oracle_stress2rhs(C.c_int(d["N"]), C.c_int(d["E"]), C.c_int(d["elem2D_nodes"].shape[1]), _d(u), _d(v),
                            _d(d["ice_strength"]), _i(d["elem2D_nodes"].reshape(-1)), _d(d["elem_area"]),
                            _d(d["sigma11"]), _d(d["sigma12"]), _d(d["sigma22"]), _d(d["gradient_sca"]),
                            _d(d["metric_factor"]), _d(d["inv_areamass"]), _d(d["rhs_a"]), _d(d["rhs_m"]))
    return u, v


def ref_stress2rhs(d):
    """The reference's own function (C++ linkage, by-value scalars: reference.cpp:440)."""
    fn = getattr(ref(), "_Z10stress2rhsiiiPdS_S_PiS_S_S_S_S_S_S_S_S_")
    u, v = np.full(d["N"], np.nan), np.full(d["N"], np.nan)
    fn(C.c_int(d["N"]), C.c_int(d["E"]), C.c_int(d["elem2D_nodes"].shape[1]), _d(u), _d(v),
       _d(d["ice_strength"]), _i(d["elem2D_nodes"].reshape(-1)), _d(d["elem_area"]), _d(d["sigma11"]),
       _d(d["sigma12"]), _d(d["sigma22"]), _d(d["gradient_sca"]), _d(d["metric_factor"]),
       _d(d["inv_areamass"]), _d(d["rhs_a"]), _d(d["rhs_m"]))
    return u, v


# ------------------------------------------------------------------ the reference's own code
def _ci(v):
    return C.byref(C.c_int(v))


def _cd(v):
    return C.byref(C.c_double(v))


def ref_a1(m, f, n_nodes=None):
    n = m.nnod if n_nodes is None else n_nodes
    ref().fct_ale_a1_reference_(_ci(n), _i(m.nlevels_nod2D), _ci(m.nl), _d(f.fct_ttf_max),
                                _d(f.fct_ttf_min), _d(f.fct_LO), _d(f.ttf))


def ref_a2(m, f):
    ref().fct_ale_a2_reference_(_ci(m.myDim_elem2D), _i(m.nlevels_elem), _ci(m.nl), _d(f.UV_rhs),
                                _i(m.elem2D_nodes), _d(f.fct_ttf_max), _d(f.fct_ttf_min),
                                _cd(f.bignumber))


def ref_a3(m, f):
    """a3 bounds + b1 vertical (reference.cpp:353-401)."""
    ref().fct_ale_a3_reference_(_ci(m.myDim_nod2D), _i(m.nlevels_nod2D), _ci(m.nl),
                                _d(f.fct_ttf_max), _d(f.fct_ttf_min), _d(f.fct_LO), _d(f.UV_rhs),
                                _d(f.fct_plus), _d(f.fct_minus), _d(f.fct_adf_v),
                                _i(m.nod_in_elem2D), _i(m.nod_in_elem2D_num),
                                _ci(m.nod_in_elem2D_dim))


def ref_a4(m, f):
    """b1 horizontal + b2 (reference.cpp:403-438)."""
    ref().fct_ale_a4_reference_(_ci(m.myDim_nod2D), _i(m.nlevels_nod2D), _i(m.nlevels_elem),
                                _ci(m.nl), _ci(m.myDim_edge2D), _d(f.fct_plus), _d(f.fct_minus),
                                _d(f.fct_adf_h), _d(f.area_inv), _d(f.fct_ttf_max),
                                _d(f.fct_ttf_min), _i(m.edges), _i(m.edge_tri), _cd(f.flux_eps),
                                _cd(f.dt))


def ref_pre_comm(m, f):
    """The composition of reference.cpp:289-304 (the symbol itself is exported C++-mangled with
    30 arguments, SURVEY.md section 8b, so it is composed here from its four C-linkage parts)."""
    ref_a1(m, f)
    ref_a2(m, f)
    ref_a3(m, f)
    ref_a4(m, f)
