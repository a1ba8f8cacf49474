"""Host-side plumbing for multi-rank runs (torch.distributed is imported lazily and only here):
rendezvous of the NCCL unique id, barriers, max-over-ranks of timings, and a host exchange_nod
(docs/refactoring.md:200) for the reference-style call sequence that stages fct_plus / fct_minus
through the host.  The device-resident path exchanges over NVLink inside the library instead."""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world():
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1          # single rank: torch is never imported
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier():
    if world()[1] > 1:
        _dist().barrier()


def max_over_ranks(x: float) -> float:
    if world()[1] == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    _dist().all_reduce(t, op=_dist().ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    if world()[1] == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    _dist().all_reduce(t, op=_dist().ReduceOp.SUM)
    return float(t.item())


def broadcast_bytes(b: bytes | None, src: int = 0) -> bytes:
    if world()[1] == 1:
        return b
    obj = [b]
    _dist().broadcast_object_list(obj, src=src)
    return obj[0]


def exchange_nod(part, arrays):
    """Owned boundary rows -> the peers' halo rows, for each array in `arrays` ([rows, L] float64)."""
    rank, nranks = world()
    if nranks == 1:
        return
    import torch
    dist = _dist()
    reqs, recvs = [], []
    for peer in sorted(part.send_lists):
        nodes = part.send_lists[peer]
        buf = torch.from_numpy(np.ascontiguousarray(np.stack([a[nodes] for a in arrays])))
        reqs.append(dist.isend(buf, dst=peer))
    for peer in sorted(part.recv_ranges):
        first, cnt = part.recv_ranges[peer]
        buf = torch.empty((len(arrays), cnt, arrays[0].shape[1]), dtype=torch.float64)
        reqs.append(dist.irecv(buf, src=peer))
        recvs.append((first, cnt, buf))
    for r in reqs:
        r.wait()
    for first, cnt, buf in recvs:
        b = buf.numpy()
        for i, a in enumerate(arrays):
            a[first:first + cnt] = b[i]


def sum_mod64(x: int) -> int:
    """Sum of one Python integer per rank, modulo 2^64 (digests)."""
    if world()[1] == 1:
        return int(x) & 0xFFFFFFFFFFFFFFFF
    vals = [None] * world()[1]
    _dist().all_gather_object(vals, int(x))
    return sum(vals) & 0xFFFFFFFFFFFFFFFF
