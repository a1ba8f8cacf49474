"""Worker of test_world_size_2_gloo: rank r owns partition r of the pi mesh, runs the oracle's
pre_comm, exchanges fct_plus / fct_minus halo rows with its peers over gloo exactly as bench.py's
host-side exchange does, runs post_comm and checks its owned results against the single-domain
oracle."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402

mesh_mod = importlib.import_module("fesom2-accelerate_b200.mesh")
comm = importlib.import_module("fesom2-accelerate_b200.hostcomm")


def main():
    out = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    want = f.copy()
    oracle.fct_ale(m, want)
    part = mesh_mod.partition_mesh(m, world, ranks=[rank])[0]
    lf = mesh_mod.slice_fields(f, part)
    oracle.pre_comm(part.mesh, lf)
    comm.exchange_nod(part, [lf.fct_plus, lf.fct_minus])
    oracle.post_comm(part.mesh, lf)
    n = part.mesh.myDim_nod2D
    g = part.mesh.node_gid[:n]
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "del_ttf_advvert",
              "del_ttf_advhoriz"):
        assert np.array_equal(getattr(lf, k)[:n], getattr(want, k)[g]), k
    assert np.array_equal(lf.fct_adf_h, want.fct_adf_h[part.mesh.edge_gid])
    # max-over-ranks reduction used for timings
    assert comm.max_over_ranks(float(rank)) == float(world - 1)
    # digests: one Python integer per rank, summed modulo 2^64
    assert comm.sum_mod64((1 << 63) + 5 + rank) == (world * ((1 << 63) + 5) + world * (world - 1) // 2) % (1 << 64)
    tot = comm.sum_mod64(mesh_mod.digest_node_array(part.mesh, lf.fct_plus, 0))
    assert tot == mesh_mod.digest_node_array(m, want.fct_plus, 0)
    dist.barrier()
    open(os.path.join(out, f"ok_{rank}"), "w").close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
