"""Generate tests/golden/*.npz.  Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Two kinds of golden vectors, both produced by the REFERENCE's own code on seeded inputs:

1. `ref_cpp_<mesh>.npz` -- outputs of the reference's src/reference.cpp (compiled unmodified into
   oracle/_ref/libref.so) for a1, a2, a3(+b1 vertical), a4 (= b1 horizontal + b2), stage by stage.
2. `ref_numpy_tiny.npz` -- outputs of the numpy `reference()` functions of
   kernels/fct_ale_b3_vertical.py:163-181, fct_ale_b3_horizontal.py:76-101,
   fct_ale_c_vertical.py:41-44, fct_ale_c_horizontal.py:53-71.  Those modules import kernel_tuner
   (not installed) at the top, so the function bodies are extracted with `ast` and exec'd; nothing
   is copied into the repo.  The .py functions index fct_adf_v / area with the same stride as the
   other arrays, while src/reference.cpp:396,:431 use stride nl; the inputs are therefore repacked
   to one common stride before the call (documented per case below) -- the arithmetic per cell is
   what is pinned.

3. `ref_cpp_stress2rhs.npz` -- outputs of the reference's stress2rhs (src/reference.cpp:440-480, the
   same compiled library).

The inputs are stored next to the outputs, so the tests never depend on the generator staying
bit-stable.
"""
import ast
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

mesh_mod = importlib.import_module("fesom2-accelerate_b200.mesh")
import oracle  # noqa: E402


def ref_function(pyfile):
    """exec only the `reference` function of a kernel_tuner script."""
    src = open(os.path.join(REF, "kernels", pyfile)).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "reference"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"numpy": np}
    exec(compile(mod, pyfile, "exec"), ns)
    return ns["reference"]


MESH_KEYS = ("nlevels_nod2D", "nlevels_elem", "elem2D_nodes", "nod_in_elem2D_num", "nod_in_elem2D",
             "edges", "edge_tri")
FIELD_KEYS = ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "area", "area_inv", "hnode", "hnode_new",
              "del_ttf_advvert", "del_ttf_advhoriz", "fct_ttf_max", "fct_ttf_min", "fct_plus",
              "fct_minus", "UV_rhs")


def pack_case(m, f):
    d = {"nl": m.nl, "N": m.myDim_nod2D, "H": m.eDim_nod2D, "dim": m.nod_in_elem2D_dim,
         "dt": f.dt, "flux_eps": f.flux_eps, "bignumber": f.bignumber}
    for k in MESH_KEYS:
        d["mesh_" + k] = getattr(m, k)
    for k in FIELD_KEYS:
        d["in_" + k] = getattr(f, k)
    return d


def gen_ref_cpp(name, m, f):
    d = pack_case(m, f)
    g = f.copy()
    oracle.ref_a1(m, g)
    d["a1_fct_ttf_max"], d["a1_fct_ttf_min"] = g.fct_ttf_max.copy(), g.fct_ttf_min.copy()
    oracle.ref_a2(m, g)
    d["a2_UV_rhs"] = g.UV_rhs.copy()
    oracle.ref_a3(m, g)
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus"):
        d["a3_" + k] = getattr(g, k).copy()
    oracle.ref_a4(m, g)
    d["a4_fct_plus"], d["a4_fct_minus"] = g.fct_plus.copy(), g.fct_minus.copy()
    np.savez_compressed(os.path.join(OUT, f"ref_cpp_{name}.npz"), **d)
    print("wrote", name, {k: v.shape for k, v in d.items() if hasattr(v, "shape") and k.startswith("a")})


def gen_ref_numpy(m, f):
    """b3 / c through the reference's numpy functions, on the state after the reference's a1..a4."""
    g = f.copy()
    oracle.ref_pre_comm(m, g)
    d = pack_case(m, g)           # inputs = state after pre_comm (limiter factors in fct_plus/minus)
    N, L, nl, G = m.myDim_nod2D, m.L, m.nl, m.myDim_edge2D
    levels = m.nlevels_nod2D.copy()

    def widen(a):   # [N,L] -> [N,nl] (stride nl, last column unused)
        w = np.zeros((a.shape[0], nl))
        w[:, :L] = a
        return w

    # b3 vertical: all arrays at stride nl (fct_adf_v's true stride, reference.cpp:396)
    b3v = ref_function("fct_ale_b3_vertical.py")
    adf_v = g.fct_adf_v.copy().reshape(-1)
    b3v(N, levels, nl, adf_v, widen(g.fct_plus).reshape(-1), widen(g.fct_minus).reshape(-1))
    d["b3v_fct_adf_v"] = adf_v.reshape(N, nl)

    # b3 horizontal: stride L everywhere, as the function expects
    b3h = ref_function("fct_ale_b3_horizontal.py")
    adf_h = g.fct_adf_h.copy().reshape(-1)
    b3h(G, m.edges.reshape(-1), m.edge_tri.reshape(-1), m.nlevels_elem, L, adf_h,
        g.fct_plus.reshape(-1), g.fct_minus.reshape(-1), np.float64)
    d["b3h_fct_adf_h"] = adf_h.reshape(G, L)

    # c vertical: everything widened to stride nl; uses the LIMITED vertical fluxes
    cv = ref_function("fct_ale_c_vertical.py")
    del_v = widen(g.del_ttf_advvert).reshape(-1)
    cv(N, levels, nl, del_v, widen(g.ttf).reshape(-1), widen(g.hnode).reshape(-1),
       widen(g.fct_LO).reshape(-1), widen(g.hnode_new).reshape(-1), adf_v, g.dt, g.area.reshape(-1))
    d["cv_del_ttf_advvert"] = del_v.reshape(N, nl)[:, :L].copy()

    # c horizontal: stride L everywhere (area narrowed); uses the LIMITED horizontal fluxes
    ch = ref_function("fct_ale_c_horizontal.py")
    del_h = g.del_ttf_advhoriz.copy().reshape(-1)
    ch(G, m.edges.reshape(-1), m.edge_tri.reshape(-1), m.nlevels_elem, L, del_h, adf_h, g.dt,
       np.ascontiguousarray(g.area[:, :L]).reshape(-1), np.float64)
    d["ch_del_ttf_advhoriz"] = del_h.reshape(N, L)
    np.savez_compressed(os.path.join(OUT, "ref_numpy_tiny.npz"), **d)
    print("wrote ref_numpy_tiny")


def gen_ref_stress2rhs():
    """stress2rhs through the reference's own compiled function (src/reference.cpp:440-480) on the
    connectivity of the tiny mesh, 0-based [3][E] as that function indexes it."""
    m = mesh_mod.make_workload("tiny", seed=0)
    tri = np.ascontiguousarray((m.elem2D_nodes - 1).T)
    d = oracle.stress_case(m.myDim_nod2D, m.myDim_elem2D, seed=7, elem_nodes=tri)
    d["U_rhs_ice"], d["V_rhs_ice"] = oracle.ref_stress2rhs(d)
    np.savez_compressed(os.path.join(OUT, "ref_cpp_stress2rhs.npz"), **d)
    print("wrote stress2rhs", d["N"], d["E"])


def main():
    oracle.build()
    assert oracle.have_ref(), "oracle/_ref/libref.so missing"
    m = mesh_mod.make_workload("tiny", seed=0)
    f = mesh_mod.make_fields(m, seed=1)
    gen_ref_cpp("tiny", m, f)
    gen_ref_numpy(m, f)
    ma, fa = mesh_mod.adversarial_case(96, 12, seed=3)
    gen_ref_cpp("adversarial", ma, fa)
    gen_ref_stress2rhs()


if __name__ == "__main__":
    main()
