"""Worker of tests/test_gpu_multi.py: one rank per GPU, partition r of a mesh on GPU r, the fused
step with the NVLink halo exchange (NCCL) inside, owned results checked against the single-domain
oracle bit for bit.  Modes: 1 = persistent warp-item kernels, overlapped schedule (also in the packed
level storage), 3 = tile-staged, 2 = untiled, 0 = staged."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402  (before the product library: one NCCL per process)

import oracle  # noqa: E402

mesh_mod = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")
comm = importlib.import_module("fesom2-accelerate_b200.hostcomm")

KEYS = ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "del_ttf_advvert", "del_ttf_advhoriz")


def main():
    out, name = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    abi.load().set_mpi_rank_(abi.ci(int(os.environ.get("LOCAL_RANK", rank))), abi.ci(world))
    # "delaunay": an unstructured triangulation cut by greedy graph growing (irregular parts, ragged halos)
    m = mesh_mod.make_delaunay_mesh(30000, 48, seed=1) if name == "delaunay" else mesh_mod.make_workload(name)
    T = 2
    fs = [mesh_mod.make_fields(m, seed=1 + t) for t in range(T)]
    for k in ("area", "area_inv", "hnode", "hnode_new"):
        setattr(fs[1], k, getattr(fs[0], k))
    wants = []
    for f in fs:
        w = f.copy()
        oracle.fct_ale(m, w)
        wants.append(w)
    owner = mesh_mod.grow_partition(m, world, seed=3) if name == "delaunay" else None
    part = mesh_mod.partition_mesh(m, world, ranks=[rank], owner=owner)[0]
    plan = harness.DevicePlan(part.mesh)
    uid = comm.broadcast_bytes(harness.HaloLink.unique_id() if rank == 0 else None)
    halo = harness.HaloLink(plan, part, uid)
    n = part.mesh.myDim_nod2D
    g = part.mesh.node_gid[:n]
    for mode, packed in ((1, True), (1, False), (3, False), (2, False), (0, False)):
        df = harness.DeviceFields(plan, T, with_uv=True, packed=packed)
        lfs = [mesh_mod.slice_fields(f, part) for f in fs]
        for t in range(T):
            df.upload(lfs[t], tracer=t, static=(t == 0))
        for rep in range(1):
            st = df.step(lfs[0], mode=mode, halo=halo)
            assert st == 10, (mode, st)
        for t in range(T):
            got = df.download(lfs[t], tracer=t, mode=mode)
            for k in KEYS:
                assert np.array_equal(getattr(got, k)[:n], getattr(wants[t], k)[g]), (rank, mode, t, k)
            assert np.array_equal(got.fct_adf_h, wants[t].fct_adf_h[part.mesh.edge_gid]), (rank, mode, t, "fct_adf_h")
        df.free()
    # vlimit 3 + one pass of the iterative branch (fct_LO halo rows exchanged over NVLink inside the
    # call), then the closing plain pass: docs/refactoring.md:131-148, :226-290
    f = fs[0].copy()
    f.vlimit, f.iter_yn = 3, True
    f.fct_adf_v2, f.fct_adf_h2 = np.zeros_like(f.fct_adf_v), np.zeros_like(f.fct_adf_h)
    want = f.copy()
    lf = mesh_mod.slice_fields(f, part)
    lf.vlimit = 3
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(lf)
    for it in (True, False):
        want.iter_yn = lf.iter_yn = it
        oracle.fct_ale_general(m, want)
        assert df.step_general(lf, halo=halo) == 10
    names = ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz",
             "fct_adf_v", "fct_adf_h", "fct_LO"]
    got = df.download(lf, names=names)
    for k in names:
        if k == "fct_adf_h":
            assert np.array_equal(got.fct_adf_h, want.fct_adf_h[part.mesh.edge_gid]), (rank, "general", k)
        else:
            assert np.array_equal(getattr(got, k)[:n], getattr(want, k)[g]), (rank, "general", k)
    # the exchanged halo rows of fct_LO equal their owners' values
    assert np.array_equal(got.fct_LO, want.fct_LO[part.mesh.node_gid]), (rank, "fct_LO halo")
    df.free()
    # vlimit 2 on the fused fast path (packed fields), overlapped schedule with the halo exchange
    f = fs[0].copy()
    f.vlimit = 2
    want = f.copy()
    oracle.fct_ale_general(m, want)
    lf = mesh_mod.slice_fields(f, part)
    lf.vlimit = 2
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(lf)
    assert df.step_general(lf, halo=halo) == 10
    got = df.download(lf, mode=1)
    for k in KEYS:
        assert np.array_equal(getattr(got, k)[:n], getattr(want, k)[g]), (rank, "fused vlimit 2", k)
    assert np.array_equal(got.fct_adf_h, want.fct_adf_h[part.mesh.edge_gid]), (rank, "fused vlimit 2", "fct_adf_h")
    df.free()
    # the iterative branch on the fused fast path: one iterative pass (fct_LO halo exchanged), one plain
    f = fs[0].copy()
    f.vlimit, f.iter_yn = 1, True
    f.fct_adf_v2, f.fct_adf_h2 = np.zeros_like(f.fct_adf_v), np.zeros_like(f.fct_adf_h)
    want = f.copy()
    lf = mesh_mod.slice_fields(f, part)
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(lf)
    for it in (True, False):
        want.iter_yn = lf.iter_yn = it
        oracle.fct_ale_general(m, want)
        assert df.step_general(lf, halo=halo) == 10
    got = df.download(lf, mode=1)
    act = np.arange(m.L)[None, :] < (part.mesh.nlevels_nod2D[:n, None] - 1)
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz"):
        assert np.array_equal(getattr(got, k)[:n][act], getattr(want, k)[g][act]), (rank, "fused iter", k)
    df.free()
    dist.barrier()
    halo.free()
    plan.free()
    open(os.path.join(out, f"ok_{rank}"), "w").close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
