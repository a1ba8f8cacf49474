#!/usr/bin/env python
"""bench.py -- FCT-ALE node-level updates/s per tracer step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ng5|dart|core2|pi]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU implementation, same metric

A "step" is one fct_ale pass a1..c over one tracer of the workload mesh.  Product arm:
  value     device-resident fused step (fct_ale_step_, mode 1), inputs already in HBM, timed with
            CUDA events on the launching stream, max over ranks.  N>1: the SAME mesh split into N
            partitions (strong scaling), halo of fct_plus/fct_minus over NVLink inside the step.
  e2e       the same step for a caller whose fields live on the HOST, through the C ABI
            (fct_ale_field_upload_ x8 -> fct_ale_step_ -> fct_ale_field_download_ x2 ->
            await_stream_) on page-locked arrays: every step uploads its inputs (ttf, fct_LO,
            fct_adf_v, fct_adf_h, hnode, hnode_new, del_ttf_adv*) and downloads the tendencies, all
            inside the timed region.  `e2e.reference_sequence` is the reference library's own call
            sequence (transfer_var_async_ -> fct_ale_pre_comm_acc_ -> await -> inter -> post ->
            fct_ale_c_acc_, five PCIe crossings per step; N>1: plus the host exchange_nod over gloo).
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event launch time vs the measured
            HBM peak (MEASURED_PEAKS.json).
  cpu_baseline  the reference's CPU code timed on this box (rank 0, N=1), bounded sample.
torch is imported only for N>1 (rendezvous / barrier / max-over-ranks); the oracle only in the
cpu_baseline / --impl reference legs.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "fesom2-accelerate_b200"
METRIC = "FCT-ALE node-level updates/s per tracer step"
UNIT = "node-level updates/s"


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        sel = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for _, r in self.rows]
        if not sel:
            return None
        sm = sorted(float(r[0]) for r in sel if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in sel):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(sel[0][1]) if sel[0][1] else None,
                "reasons": reasons, "samples": len(sel)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU code on this box's host cores
# ------------------------------------------------------------------------------------------------
def _cpu_chain_worker(args):
    """One process = one MPI-style rank of the CPU model: full chain on its own sample mesh."""
    nx, ny, nl, reps, seed = args
    import oracle
    mesh = importlib.import_module(PKG + ".mesh")
    m = mesh.make_mesh(nx, ny, nl, seed=seed)
    f0 = mesh.fast_fields(m, seed=seed + 1, with_uv=True)
    use_ref = oracle.have_ref()
    times = []
    for _ in range(reps):
        f = f0.copy()
        t0 = time.perf_counter()
        if use_ref:
            oracle.ref_pre_comm(m, f)      # a1, a2, a3+b1v, a4 = b1h+b2: src/reference.cpp unmodified
        else:
            oracle.pre_comm(m, f)
        oracle.post_comm(m, f)             # b3, c: restated (the reference has no working C++ for them)
        times.append(time.perf_counter() - t0)
    return m.S_n(), times, use_ref


def cpu_chain(nx, ny, nl, procs, reps):
    """-> (updates per pass over all procs, per-pass wall times (max over procs), kind)"""
    import multiprocessing as mp
    jobs = [(nx, ny, nl, reps, 100 + i) for i in range(procs)]
    if procs == 1:
        res = [_cpu_chain_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_chain_worker, jobs)
    upd = sum(r[0] for r in res)
    per_pass = [max(r[1][k] for r in res) for k in range(reps)]
    return upd, per_pass, ("reference" if res[0][2] else "port")


CPU_SAMPLE = (384, 301)    # 1/64 of the NG5 grid per process (~115 k nodes), same generator and depth statistics


def run_reference(args):
    """--impl reference: the reference's CPU implementation, all host cores (one MPI-style process
    per core, each on its own sample mesh of the workload's depth; no halo exchange between them)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    mesh = importlib.import_module(PKG + ".mesh")
    nl = mesh.WORKLOADS[args.workload]["nl"]
    cores = max(1, min(os.cpu_count() or 1, 64))
    nx, ny = CPU_SAMPLE
    reps = args.warmup + args.steps
    t0 = time.time()
    upd, per_pass, kind = cpu_chain(nx, ny, nl, cores, reps)
    timed = per_pass[args.warmup:]
    total = sum(timed)
    value = upd * len(timed) / total
    sample = (f"{cores} processes x one {nx}x{ny} grid ({upd // cores} node-level updates each, nl={nl}, same generator as "
              f"the {args.workload} mesh), full chain a1..c per step; a1..b2 = src/reference.cpp unmodified "
              f"(oracle/_ref/libref.so), b3/c = oracle/fct_ale_oracle.c (the reference has no working C++ for them)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "tracers": 1, "levels": nl, "cpu_sample": f"{cores}x{nx}x{ny}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    _RESULT.append(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def run_product(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    abi = importlib.import_module(PKG + ".abi")
    mesh = importlib.import_module(PKG + ".mesh")
    harness = importlib.import_module(PKG + ".harness")
    hostcomm = importlib.import_module(PKG + ".hostcomm")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    lib = abi.load()
    lib.set_mpi_rank_(abi.ci(local_rank), abi.ci(max(world, 1)))
    devname, cc, sms = abi.device_info()
    t_setup = time.time()

    w = mesh.WORKLOADS[args.workload]
    gm = mesh.make_mesh(w["nx"], w["ny"], w["nl"], seed=0)
    Sn_total, Sg_total, N_total = gm.S_n(), gm.S_g(), gm.myDim_nod2D
    log(f"{args.workload}: {N_total} nodes, {gm.myDim_elem2D} elements, {gm.myDim_edge2D} edges, nl={gm.nl}, "
        f"S_n={Sn_total} ({time.time() - t_setup:.1f}s)")
    if world > 1:
        part = mesh.partition_mesh(gm, world, ranks=[rank])[0]
        m = part.mesh
        del gm
    else:
        part, m = None, gm
    Sn, Sg = m.S_n(), m.S_g()
    f = mesh.fast_fields(m, seed=1 + rank, alloc=abi.pinned_empty)     # page-locked host arrays
    log(f"fields ready ({time.time() - t_setup:.1f}s)")

    # ---------------- value: device-resident fused step ----------------
    plan = harness.DevicePlan(m)
    # the packed level storage is the fast path's own layout (plain triangulations)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=plan.packed_ok)
    df.upload(f, outputs=False)
    halo = None
    if world > 1:
        uid = hostcomm.broadcast_bytes(harness.HaloLink.unique_id() if rank == 0 else None)
        halo = harness.HaloLink(plan, part, uid)
    log(f"plan + upload ready ({time.time() - t_setup:.1f}s)")
    e0, e1 = abi.Event(), abi.Event()
    for _ in range(args.warmup):
        st = df.step(f, mode=1, halo=halo, sync=False)
    df.stream.sync()
    assert args.warmup == 0 or st == 10, f"alg_state {st}"
    hostcomm.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = abi.launch_count()
    t0 = time.time()
    e0.record(df.stream)
    for _ in range(args.steps):
        df.step(f, mode=1, halo=halo, sync=False)
    e1.record(df.stream)
    ms = e1.ms_since(e0)                      # synchronises on e1
    df.stream.sync()
    t1 = time.time()
    launches = abi.launch_count() - n0
    hostcomm.barrier()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_max = hostcomm.max_over_ranks(ms)
    ms_step = ms_max / args.steps
    value = Sn_total / (ms_step * 1e-3)
    bytes_alg_total = 8 * (21 * Sn_total + 3 * Sg_total) + 16 * N_total

    # ---------------- roofline: the two fused phase kernels, per-launch CUDA-event time ----------------
    peak, peak_src = measured_peak()
    algA = 8 * (8 * Sn + Sg) + 16 * m.myDim_nod2D
    algB = 8 * (13 * Sn + 2 * Sg)
    kern = {}
    suffix = {"warp": "_warp", "tile": "_tile", "untiled": ""}["warp" if plan.packed_ok else plan.kernels]
    mode_name = {"warp": "persistent TMA-staged warp-item fused phases A+B, " + ("packed level storage" if plan.packed_ok else "padded rows"), "tile": "tile-staged fused phases A+B",
                 "untiled": "untiled fused phases A+B"}[plan.kernels]
    if world == 1:
        for base, alg in (("phaseA", algA), ("phaseB", algB)):
            name = base + suffix                  # the kernels fct_ale_step_ mode 1 runs on this plan
            df.stage(name, f, sync=True)
            reps = max(3, min(args.steps, 10))
            e0.record(df.stream)
            for _ in range(reps):
                df.stage(name, f, sync=False)
            e1.record(df.stream)
            kern[name] = (e1.ms_since(e0) / reps, alg)
        dom = max(kern, key=lambda k: kern[k][0])
        dms, dalg = kern[dom]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(dom)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": {"phaseA_warp": "k_phase_warp<true,...>", "phaseB_warp": "k_phase_warp<false,...>"}.get(dom, "k_" + dom), "achieved": dalg / dms / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": dalg / dms / 1e6 / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": dalg, "ms_per_launch": dms,
                    "kernels": {k: {"ms": v[0], "alg_GBs": v[1] / v[0] / 1e6, "frac": v[1] / v[0] / 1e6 / peak} for k, v in kern.items()}}
    else:
        roofline = {"bound": "hbm", "kernel": "fct_ale_step (both fused phases + halo)", "achieved": bytes_alg_total / world / ms_step / 1e6,
                    "peak": peak, "unit": "GB/s", "frac": bytes_alg_total / world / ms_step / 1e6 / peak, "traffic": None,
                    "peak_source": peak_src, "note": "per-GPU average over the whole step"}
    log(f"value done: {ms_step:.3f} ms/step ({time.time() - t_setup:.1f}s)")

    # ---------------- e2e: host-resident caller, device-resident step ----------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    assert df.host_step(f, f, mode=1, halo=halo) == 10       # warm-up
    hostcomm.barrier()
    ta = time.perf_counter()
    # every step uploads its inputs and downloads its tendencies; the download of step k (second
    # stream) overlaps the upload of step k+1; ends with await_stream_: the tendencies are on the host
    assert df.host_steps(f, f, e2e_steps, mode=1, halo=halo) == 10
    tb = time.perf_counter()
    e2e_s = hostcomm.max_over_ranks((tb - ta) / e2e_steps)
    hb, db = df.host_step_bytes(f)
    h2d, d2h = hostcomm.sum_over_ranks(hb), hostcomm.sum_over_ranks(db)
    if halo is not None:
        halo.free()
    df.free()
    plan.free()
    log(f"e2e done: {e2e_s * 1e3:.1f} ms/step ({time.time() - t_setup:.1f}s)")

    # ---------------- the reference library's own call sequence on host arrays (context) ----------------
    refseq = None
    if not args.no_refseq:
        exchange = None
        if world > 1:
            def exchange(ff):
                hostcomm.exchange_nod(part, [ff.fct_plus, ff.fct_minus])
        ch = harness.HandleChain(m, f, with_c=True)
        ch.step(exchange)                          # warm-up (also builds the cached plan)
        hostcomm.barrier()
        ta = time.perf_counter()
        ch.step(exchange)                          # ends with await_stream_: results are on the host
        tb = time.perf_counter()
        rs = hostcomm.max_over_ranks(tb - ta)
        refseq = {"value": Sn_total / rs, "unit": UNIT, "ms_per_step": rs * 1e3, "steps": 1,
                  "h2d_bytes_per_step": int(hostcomm.sum_over_ranks(ch.h2d_bytes())),
                  "d2h_bytes_per_step": int(hostcomm.sum_over_ranks(ch.d2h_bytes())),
                  "api": "transfer_var_async_ / fct_ale_pre_comm_acc_ / inter / post / fct_ale_c_acc_ (reference call sequence, stage kernels)"}
        ch.free()
        log(f"reference call sequence done: {rs * 1e3:.1f} ms/step ({time.time() - t_setup:.1f}s)")

    # ---------------- cpu baseline (rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        nx, ny = CPU_SAMPLE
        upd, per_pass, kind = cpu_chain(nx, ny, m.nl, 1, 3)
        best = min(per_pass[1:])
        cpu = {"value": upd / best, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"one {nx}x{ny} grid of the same generator and depth (nl={m.nl}, {upd} node-level updates), full chain "
                         f"a1..c, best of 2 after one warm-up; a1..b2 = src/reference.cpp unmodified, b3/c = oracle port; "
                         f"host has {os.cpu_count()} cores, the reference is single-threaded"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": args.workload, "nodes": N_total, "levels": w["nl"], "tracers": 1, "partitions": world,
                           "node_level_updates": Sn_total, "mode": mode_name, "device": devname,
                           "l2": "inputs (tens of GB) far larger than the 126 MB L2; no flush needed",
                           "alg_bytes_per_step": bytes_alg_total},
                "hbm": {"alg_GBs_per_gpu": bytes_alg_total / world / ms_step / 1e6, "frac_of_peak": bytes_alg_total / world / ms_step / 1e6 / peak},
                "clocks": clocks,
                "e2e": {"value": Sn_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                        "api": "per step: fct_ale_field_upload_ x8 / fct_ale_step_ / fct_ale_field_download_ x2 on page-locked host arrays (download of step k on a second stream, overlapping the upload of step k+1), await_stream_ at the end",
                        "reference_sequence": refseq},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "setup_s": time.time() - t_setup}
        _RESULT.append(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    # native libraries (NCCL's version banner, ...) write to file descriptor 1: route it to stderr
    # for the duration of the run so that stdout carries exactly the one JSON line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        _main()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if _RESULT:
        print(_RESULT[0], flush=True)


_RESULT = []


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FCT_BENCH_WORKLOAD", "ng5"))
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-refseq", action="store_true", help="skip the timing of the reference library's call sequence")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
