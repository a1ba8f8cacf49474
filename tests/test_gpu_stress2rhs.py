"""GPU parity of stress2rhs through the C ABI: bit-exact against the
oracle, which is pinned to the reference's compiled src/reference.cpp:440-480."""
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def mesh_case(oracle_mod, mesh_mod, name, seed):
    m = mesh_mod.make_workload(name)
    tri = np.ascontiguousarray((m.elem2D_nodes - 1).T)          # 0-based [3][E]
    return oracle_mod.stress_case(m.myDim_nod2D, m.myDim_elem2D, seed=seed, elem_nodes=tri)


def test_stress2rhs_golden(harness):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cpp_stress2rhs.npz"))
    d = {k: np.ascontiguousarray(z[k]) for k in z.files}
    d["N"], d["E"] = int(z["N"]), int(z["E"])
    u, v = harness.stress2rhs_host(d)
    assert bits_equal(u, z["U_rhs_ice"]) and bits_equal(v, z["V_rhs_ice"])


@pytest.mark.parametrize("name", ["tiny", "pi", "core2"])
def test_stress2rhs_device_resident(harness, oracle_mod, mesh_mod, name):
    d = mesh_case(oracle_mod, mesh_mod, name, seed=3)
    want = oracle_mod.stress2rhs(d)
    ch = harness.StressChain(d)
    ch.run()
    u, v = ch.fetch()
    assert bits_equal(u, want[0]) and bits_equal(v, want[1])
    ch.run()                     # idempotent: the sums start from zero every call (reference.cpp:447-451)
    u, v = ch.fetch()
    assert bits_equal(u, want[0]) and bits_equal(v, want[1])
    ch.free()


def test_stress2rhs_random_connectivity_and_edge_cases(harness, oracle_mod):
    """Non-manifold random connectivity (repeated corners, isolated nodes), no ice anywhere, zero
    inverse mass everywhere, an empty mesh."""
    d = oracle_mod.stress_case(700, 1500, seed=5)
    want = oracle_mod.stress2rhs(d)
    got = harness.stress2rhs_host(d)
    assert bits_equal(got[0], want[0]) and bits_equal(got[1], want[1])
    e = dict(d)
    e["ice_strength"] = np.zeros_like(d["ice_strength"])
    got, want = harness.stress2rhs_host(e), oracle_mod.stress2rhs(e)
    assert bits_equal(got[0], want[0]) and bits_equal(got[1], want[1])
    assert bits_equal(got[0], np.where(e["inv_areamass"] > 0, e["rhs_a"], 0.0))
    e = dict(d)
    e["inv_areamass"] = np.zeros_like(d["inv_areamass"])
    got = harness.stress2rhs_host(e)
    assert not got[0].any() and not got[1].any()
    z = oracle_mod.stress_case(0, 0, seed=1)
    u, v = harness.stress2rhs_host(z)
    assert u.size == 0 and v.size == 0


def test_stress2rhs_full_size_linearity(harness, oracle_mod, mesh_mod):
    """NG5/4-sized connectivity (1.8 M nodes): the divergence is linear in the stress tensor, so
    scaling sigma by 2 (exact in binary) scales U - rhs_a by exactly 2."""
    m = mesh_mod.make_mesh(1536, 1204, 3)
    tri = np.ascontiguousarray((m.elem2D_nodes - 1).T)
    d = oracle_mod.stress_case(m.myDim_nod2D, m.myDim_elem2D, seed=9, elem_nodes=tri)
    d["rhs_a"][:] = 0.0
    d["rhs_m"][:] = 0.0
    a = harness.stress2rhs_host(d)
    e = dict(d)
    for k in ("sigma11", "sigma12", "sigma22"):
        e[k] = 2.0 * d[k]
    b = harness.stress2rhs_host(e)
    assert bits_equal(b[0], 2.0 * a[0]) and bits_equal(b[1], 2.0 * a[1])
    assert np.count_nonzero(a[0]) > 0.5 * d["N"]
