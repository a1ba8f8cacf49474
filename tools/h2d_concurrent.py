"""Concurrent host->device (and device->host) copy bandwidth of all ranks of one box: what bounds the
end-to-end figure of a host-resident caller at N GPUs.  Run under torchrun (one rank per GPU) or alone.
Each rank copies a page-locked 2 GiB block to its GPU `reps` times, all ranks at once (barrier), then alone
(one rank at a time); prints per-rank and aggregate GB/s.
usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_concurrent.py"""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
abi = importlib.import_module("fesom2-accelerate_b200.abi")
hostcomm = importlib.import_module("fesom2-accelerate_b200.hostcomm")

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
lib = abi.load()
lib.set_mpi_rank_(abi.ci(local), abi.ci(world))
abi.device_info()
n = (2 << 30) // 8
host = abi.pinned_empty(n)
host[:] = 1.0
var = abi.Var(host)
up, dn = abi.Stream(), abi.Stream()
reps = 6


def run(both):
    t0 = time.perf_counter()
    for _ in range(reps):
        var.upload(stream=up)
        if both:
            var.download(stream=dn)     # same buffer both ways: bandwidth only
    up.sync()
    dn.sync()
    return reps * host.nbytes / (time.perf_counter() - t0) / 1e9


run(False)
res = {}
for name, both in (("h2d", False), ("h2d+d2h", True)):
    hostcomm.barrier()
    g = run(both)
    res[name + "_concurrent_per_rank"] = g
    res[name + "_concurrent_aggregate"] = hostcomm.sum_over_ranks(g)
    alone = 0.0
    for r in range(world):
        hostcomm.barrier()
        if r == rank:
            alone = run(both)
    res[name + "_alone_per_rank"] = alone
allres = [None] * world
if world > 1:
    dist.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    try:
        cpus = len(os.sched_getaffinity(0))
    except Exception:
        cpus = os.cpu_count()
    print(json.dumps({"ranks": world, "host_cpus": cpus, "block_bytes": host.nbytes, "reps": reps, "per_rank": allres,
                      "note": "GB/s of payload per direction (h2d+d2h: each direction moves this much)"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
