import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "fesom2-accelerate_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # incremental build of the product library and the oracle (no-op when up to date; on the GPU box
    # the prebuilt .so files travel with the snapshot and nvcc is still available)
    try:
        subprocess.run(["make", "-s", "-C", ROOT], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, timeout=900)
    except Exception as exc:  # pragma: no cover
        print("WARNING: make failed:", getattr(exc, "stderr", exc), file=sys.stderr)


def pkg(sub=None):
    return importlib.import_module(PKG + ("." + sub if sub else ""))


@pytest.fixture(scope="session")
def mesh_mod():
    return pkg("mesh")


@pytest.fixture(scope="session")
def abi():
    return pkg("abi")


@pytest.fixture(scope="session")
def harness():
    return pkg("harness")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def load_golden(name):
    """-> (Mesh, Fields, npz dict) from tests/golden/<name>.npz"""
    mm = pkg("mesh")
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m = mm.Mesh(nl=int(z["nl"]), myDim_nod2D=int(z["N"]), eDim_nod2D=int(z["H"]),
                myDim_elem2D=z["mesh_elem2D_nodes"].shape[0], myDim_edge2D=z["mesh_edges"].shape[0],
                nlevels_nod2D=z["mesh_nlevels_nod2D"], nlevels_elem=z["mesh_nlevels_elem"],
                elem2D_nodes=z["mesh_elem2D_nodes"], nod_in_elem2D_num=z["mesh_nod_in_elem2D_num"],
                nod_in_elem2D=z["mesh_nod_in_elem2D"], nod_in_elem2D_dim=int(z["dim"]),
                edges=z["mesh_edges"], edge_tri=z["mesh_edge_tri"])
    kw = {k[3:]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("in_")}
    f = mm.Fields(dt=float(z["dt"]), flux_eps=float(z["flux_eps"]), bignumber=float(z["bignumber"]), **kw)
    return m, f, z


def bits_equal(a, b):
    """IEEE value equality cell by cell, untouched sentinels included (+0 == -0, no NaNs expected)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def has_gpu():
    try:
        pkg("abi").device_info()
        return True
    except Exception:
        return False
