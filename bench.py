#!/usr/bin/env python
"""bench.py -- FCT-ALE node-level updates/s per tracer step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ng5|dart|core2|pi]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU implementation, same metric

A "step" is one fct_ale pass a1..c over one tracer of the workload mesh.  Product arm:
  parity    BEFORE anything is timed: the CORE2-size mesh split into N partitions, run through the
            same DevicePlan / DeviceFields / HaloLink objects (NCCL halo exchange included) and compared
            bit for bit with the CPU oracle on the single domain (the oracle is the checker here,
            never the thing measured).  A mismatch ends the run with a non-zero exit code.
  digest    order-independent 64-bit digest (keyed by global node id and level) of fct_plus,
            fct_minus, del_ttf_advvert, del_ttf_advhoriz after ONE step on the freshly uploaded
            workload fields; the input fields are a function of the global ids only, so the lines of
            N = 1, 2, 4, 8 must show the same digests.
  value     device-resident fused step (fct_ale_step_, mode 1), inputs already in HBM, timed with
            CUDA events on the launching stream, max over ranks, after `preroll_s` seconds of untimed
            steps (thermal / power-cap steady state at every N).  N>1: the SAME mesh split into N
            partitions (strong scaling), halo of fct_plus/fct_minus over NVLink inside the step;
            `halo` = {comm_ms: device time of the exchange on its own stream, exposed_ms: step time
            minus the time of the same launches without the exchange, floor_ms: the step on the
            CORE2-size mesh split N ways = launch + NCCL latency floor}.
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
--impl reference: the reference's CPU code on the SAME workload mesh, split into one partition per
host core (processes), host exchange_nod between pre- and post-comm -- what FESOM2 does under MPI.
torch is imported only for N>1 (rendezvous / barrier / max-over-ranks); the oracle only in the
parity gate, the cpu_baseline and the --impl reference legs.
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
        # samples inside the timed region; a region shorter than the 100 ms sampling interval may hold none: then
        # the samples nearest to it (the pre-roll right before it runs the same steps at the same load)
        sel = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15]
        inside = len(sel)
        if not sel:
            near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:3]
            sel = [r for ts, r in near if min(abs(ts - t0), abs(ts - t1)) <= 1.0]
        if not sel:
            return None
        sm = sorted(float(r[0]) for r in sel if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in sel):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(sel[0][1]) if sel[0][1] else None,
                "reasons": reasons, "samples": len(sel), "samples_inside_timed_region": inside}


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


CPU_SAMPLE = (384, 301)    # 1/64 of the NG5 grid (~115 k nodes), same generator and depth statistics: the 1-core cpu_baseline


# --impl reference, same configuration as the product arm: the workload mesh itself, one partition per
# host core, each process running the reference's a1..b2 (src/reference.cpp unmodified) + the restated
# b3/c on its partition with the host exchange_nod of fct_plus / fct_minus in between -- FESOM2 under MPI
# (reference.cpp:289-304 + docs/refactoring.md:200, :235).
_GM = None          # the global mesh, inherited by the forked workers (copy-on-write, built once)


def _ref_rank(rank, cores, port, reps, budget_s, conn):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(cores),
                          LOCAL_RANK=str(rank), OMP_NUM_THREADS="1")
        import oracle
        mesh = importlib.import_module(PKG + ".mesh")
        hostcomm = importlib.import_module(PKG + ".hostcomm")
        if cores > 1:
            import torch
            import torch.distributed as dist
            torch.set_num_threads(1)
            dist.init_process_group("gloo")
            part = mesh.partition_mesh(_GM, cores, ranks=[rank])[0]
            m = part.mesh
        else:
            part, m = None, _GM
        f0 = mesh.fast_fields(m, seed=1, with_uv=True)
        use_ref = oracle.have_ref()
        times, dig = [], {}
        f = f0
        for k in range(reps):
            # (no fresh copy per step: 25 partitions' worth of fields would not fit the host; later steps
            #  re-limit the already limited fluxes in place, which costs the same arithmetic)
            hostcomm.barrier()
            t0 = time.perf_counter()
            if use_ref:
                oracle.ref_pre_comm(m, f)      # a1, a2, a3+b1v, a4 = b1h+b2: src/reference.cpp unmodified
            else:
                oracle.pre_comm(m, f)
            if part is not None:
                hostcomm.exchange_nod(part, [f.fct_plus, f.fct_minus])
            oracle.post_comm(m, f)             # b3, c: restated (the reference has no working C++ for them)
            dt = hostcomm.max_over_ranks(time.perf_counter() - t0)
            times.append(dt)
            if k == 0:   # digest of the outputs of ONE step on the fresh fields: comparable with the product arm's
                dig = {name: mesh.digest_node_array(m, getattr(f, name), i) for i, name in enumerate(DIGEST_FIELDS)}
            # bounded: stop early (all ranks agree: dt is the max over ranks) when the budget is spent
            if sum(times) > budget_s and k >= 1:
                break
        conn.send((m.S_n(), times, use_ref, dig, None))
        if cores > 1:
            dist.barrier()
            dist.destroy_process_group()
    except Exception as exc:   # pragma: no cover
        import traceback
        conn.send((0, [], False, {}, traceback.format_exc()))
    finally:
        conn.close()


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host cores, same workload mesh."""
    global _GM
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    mesh = importlib.import_module(PKG + ".mesh")
    w = mesh.WORKLOADS[args.workload]
    t0 = time.time()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 64))
    _GM = mesh.make_mesh(w["nx"], w["ny"], w["nl"], seed=0)
    Sn_total, N_total = _GM.S_n(), _GM.myDim_nod2D
    config = workload_config(args.workload, _GM)
    log(f"reference arm: {args.workload} mesh {N_total} nodes on {cores} host processes ({time.time() - t0:.1f}s)")
    # memory: ~21 dense node-array equivalents per partition (14 node arrays, fct_adf_h = 3, UV_rhs = 4) + halo rows
    need = 1.2 * 24 * 8.0 * N_total * w["nl"]
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    reps = args.warmup + args.steps
    same = avail is None or need < 0.8 * avail
    if same:
        ctx = mp.get_context("fork")
        import socket
        with socket.socket() as sk:             # a free port for the workers' own gloo rendezvous
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        pipes, procs = [], []
        for r in range(cores):
            a, b = ctx.Pipe(duplex=False)
            pr = ctx.Process(target=_ref_rank, args=(r, cores, port, reps, args.reference_budget, b))
            pr.start()
            b.close()
            pipes.append(a)
            procs.append(pr)
        res = [c.recv() for c in pipes]
        for pr in procs:
            pr.join()
        errs = [r[4] for r in res if r[4]]
        if errs:
            raise RuntimeError("reference worker failed:\n" + errs[0])
        assert sum(r[0] for r in res) == Sn_total
        per_pass = res[0][1]
        kind = "reference" if res[0][2] else "port"
        digest = {k: "0x%016x" % (sum(r[3][k] for r in res) & 0xFFFFFFFFFFFFFFFF) for k in DIGEST_FIELDS}
        upd = Sn_total
        sample = (f"the {args.workload} mesh itself ({N_total} nodes, nl={w['nl']}, {Sn_total} node-level updates) split into {cores} "
                  f"partitions, one process per host core, host exchange_nod of fct_plus/fct_minus (gloo) between pre- and "
                  f"post-comm; a1..b2 = src/reference.cpp unmodified (oracle/_ref/libref.so), b3/c = oracle/fct_ale_oracle.c "
                  f"(the reference has no working C++ for them)")
    else:
        nx, ny = CPU_SAMPLE
        upd, per_pass, kind = cpu_chain(nx, ny, w["nl"], cores, min(reps, 4))
        digest = None
        sample = (f"host memory too small for the whole mesh: {cores} processes x one {nx}x{ny} grid of the same generator and depth")
    nw = min(args.warmup, max(len(per_pass) - 1, 0))
    timed = per_pass[nw:]
    total = sum(timed)
    value = upd * len(timed) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "details": {"partitions": cores if same else None, "same_mesh_as_product_arm": bool(same), "steps_timed": len(timed),
                        "note": "steps beyond the time budget (--reference-budget seconds of CPU steps) are not run"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "digest": digest,
            "gpu_launches": 0, "wall_s": time.time() - t0}
    _RESULT.append(json.dumps(line))


DIGEST_FIELDS = ("fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz")


def workload_config(name, gm):
    """`config` of the JSON line: the workload and nothing arm-specific, so that the product arm and the
    reference arm print the same object when they ran the same thing."""
    Sn, Sg, N = gm.S_n(), gm.S_g(), gm.myDim_nod2D
    return {"workload": name, "nodes": N, "elements": gm.myDim_elem2D, "edges": gm.myDim_edge2D, "levels": gm.nl, "tracers": 1,
            "node_level_updates": Sn, "edge_level_updates": Sg, "alg_bytes_per_step": 8 * (21 * Sn + 3 * Sg) + 16 * N,
            "fields": "synthetic, a function of (global row id, level, array) only: mesh.fast_fields(seed=1)",
            "l2": (f"one step streams {(8 * (21 * Sn + 3 * Sg)) / 1e9:.2f} GB, " +
                   ("far more than the 126 MB L2: no flush needed" if 8 * (21 * Sn + 3 * Sg) > 8 * 126e6 else
                    "comparable to the 126 MB L2: not a bench configuration"))}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
REF_GPU_SO = os.path.join(ROOT, "baseline", "_ref", "libref_gpu_kernels.so")
REF_GPU_STAGES = ["a1", "a2", "a3", "b1v", "b1h_grid_nodes_as_shipped", "b1h", "b2", "b3v", "b3h", "cv", "ch", "sequence"]


def gpu_reference(m, abi, reps=3):
    """The reference's OWN kernels (/root/reference/kernels/*.cu compiled unmodified for sm_100 by
    baseline/build_ref_gpu.sh) on the workload mesh, device-resident, CUDA events, launched with the
    reference driver's geometry (src/fesom2-accelerate.cu:294-335).  Timing only: the reference launches
    b1_horizontal over nodes instead of edges (:327), its results are not a parity oracle."""
    if not os.path.exists(REF_GPU_SO):
        return {"unavailable": "baseline/_ref/libref_gpu_kernels.so not built (needs /root/reference at build time)"}
    import ctypes as C
    try:
        lib = C.CDLL(REF_GPU_SO)
    except OSError as exc:
        return {"unavailable": f"baseline/_ref/libref_gpu_kernels.so does not load: {exc}"}
    ms = (C.c_double * 12)()
    st = C.c_int()
    ci, ip = abi.ci, abi.iptr
    lib.ref_gpu_bench_(ci(m.myDim_nod2D), ci(m.eDim_nod2D), ci(m.myDim_elem2D), ci(m.myDim_edge2D), ci(m.nl),
                       ip(m.nlevels_nod2D), ip(m.nlevels_elem), ip(m.elem2D_nodes.reshape(-1)), ip(m.nod_in_elem2D_num),
                       ip(m.nod_in_elem2D.reshape(-1)), ci(m.nod_in_elem2D_dim), ip(m.edges.reshape(-1)),
                       ip(m.edge_tri.reshape(-1)), ci(reps), ms, C.byref(st))
    if st.value != 0:
        return {"unavailable": "ref_gpu_bench_ failed (see stderr)"}
    return {"stages_ms": dict(zip(REF_GPU_STAGES[:11], [float(x) for x in ms][:11])), "step_ms": float(ms[11]),
            "what": "reference kernels rebuilt for sm_100 (baseline/_ref), one 32-thread block per node / element / edge, fp64 atomics in "
                    "b1h / c_h; sequence = a1..c with b1h over edges; device-resident, CUDA events; timing only"}


PARITY_KEYS = ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "del_ttf_advvert", "del_ttf_advhoriz")


def parity_gate(rank, world, abi, mesh, harness, hostcomm):
    """The CORE2-size mesh split `world` ways through the SAME DevicePlan / DeviceFields / HaloLink
    classes (fused warp-item kernels, packed storage, NCCL halo exchange overlapped with interior work)
    the timed run uses, compared bit for bit with the CPU oracle on the single domain -- stage
    comparison as /root/reference/src/fesom2-accelerate.cu:295-335 does against reference.cpp:289-304.
    Also times that small step: with ~15 k nodes per GPU at N = 8 it is the launch + NCCL latency floor."""
    import oracle                                   # the checker, never the thing measured
    oracle.build()
    gm = mesh.make_workload("core2")
    f = mesh.make_fields(gm, seed=1, with_uv=True)
    want = f.copy()
    oracle.fct_ale(gm, want)
    if world > 1:
        part = mesh.partition_mesh(gm, world, ranks=[rank])[0]
        m, lf = part.mesh, mesh.slice_fields(f, part)
    else:
        part, m, lf = None, gm, f
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=plan.packed_ok)
    df.upload(lf)
    halo = None
    if world > 1:
        uid = hostcomm.broadcast_bytes(harness.HaloLink.unique_id() if rank == 0 else None)
        halo = harness.HaloLink(plan, part, uid)
    st = df.step(lf, mode=1, halo=halo)
    got = df.download(lf, mode=1)
    n = m.myDim_nod2D
    g = np.arange(n) if part is None else m.node_gid[:n]
    eg = slice(None) if part is None else m.edge_gid
    bad = [k for k in PARITY_KEYS if not np.array_equal(getattr(got, k)[:n], getattr(want, k)[g])]
    if not np.array_equal(got.fct_adf_h, want.fct_adf_h[eg]):
        bad.append("fct_adf_h")
    if st != 10:
        bad.append(f"alg_state {st}")
    # latency floor: the same step, back to back
    e0, e1 = abi.Event(), abi.Event()
    for _ in range(5):
        df.step(lf, mode=1, halo=halo, sync=False)
    df.stream.sync()
    hostcomm.barrier()
    e0.record(df.stream)
    for _ in range(50):
        df.step(lf, mode=1, halo=halo, sync=False)
    e1.record(df.stream)
    floor_ms = hostcomm.max_over_ranks(e1.ms_since(e0) / 50)
    if halo is not None:
        halo.free()
    df.free()
    plan.free()
    nbad = hostcomm.sum_over_ranks(float(len(bad)))
    if bad:
        print(f"[bench] PARITY MISMATCH on rank {rank}: {bad}", file=sys.stderr, flush=True)
    return {"checked": True, "ok": nbad == 0, "mesh": "core2", "nodes": gm.myDim_nod2D, "ranks": world,
            "against": "oracle.fct_ale on the single domain (a1..b2 pinned to src/reference.cpp), bit for bit on owned rows",
            "fields": list(PARITY_KEYS) + ["fct_adf_h"],
            "path": "DevicePlan / DeviceFields(packed) / HaloLink / fct_ale_step_ mode 1" + (" + NCCL halo" if world > 1 else "")}, floor_ms


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
    config = workload_config(args.workload, gm)
    log(f"{args.workload}: {N_total} nodes, {gm.myDim_elem2D} elements, {gm.myDim_edge2D} edges, nl={gm.nl}, "
        f"S_n={Sn_total} ({time.time() - t_setup:.1f}s)")
    if world > 1:
        part = mesh.partition_mesh(gm, world, ranks=[rank])[0]
        m = part.mesh
        del gm
    else:
        part, m = None, gm
    Sn, Sg = m.S_n(), m.S_g()
    # every cell is a function of (global row id, level, array): all N see the same global fields
    f = mesh.fast_fields(m, seed=1, alloc=abi.pinned_empty)     # page-locked host arrays
    log(f"fields ready ({time.time() - t_setup:.1f}s)")

    # ---------------- parity gate (before anything is timed) ----------------
    parity, floor_ms = ({"checked": False, "ok": None}, None)
    if not args.no_parity:
        parity, floor_ms = parity_gate(rank, world, abi, mesh, harness, hostcomm)
        log(f"parity gate: {parity['ok']} (core2 split {world} ways, floor {floor_ms:.3f} ms/step) ({time.time() - t_setup:.1f}s)")
        if not parity["ok"]:
            if rank == 0:
                _RESULT.append(json.dumps({"metric": METRIC, "value": None, "n_gpus": world, "parity": parity,
                                           "error": "parity gate failed: nothing was timed"}))
            _EXIT.append(3)
            return

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

    # ---------------- digest of ONE step on the fresh fields (identical for every N) ----------------
    st = df.step(f, mode=1, halo=halo)
    assert st == 10, f"alg_state {st}"
    digest = {}
    scratch = f.fct_plus                        # page-locked, not an input of the step
    for i, name in enumerate(DIGEST_FIELDS):
        df.download_field(name, scratch, merge=False)
        df.stream.sync()
        digest[name] = "0x%016x" % hostcomm.sum_mod64(mesh.digest_node_array(m, scratch, i))
    log(f"digest {digest} ({time.time() - t_setup:.1f}s)")

    e0, e1 = abi.Event(), abi.Event()

    def timed_steps(k):
        hostcomm.barrier()
        e0.record(df.stream)
        for _ in range(k):
            df.step(f, mode=1, halo=halo, sync=False)
        e1.record(df.stream)
        ms_ = e1.ms_since(e0)                  # synchronises on e1
        df.stream.sync()
        return ms_

    # (started before the warm-up: nvidia-smi needs a moment before its first sample, and at N = 8 the timed region
    #  of 20 steps is shorter than its 100 ms sampling interval)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    est = 0.0
    if args.warmup > 0:
        est = hostcomm.max_over_ranks(timed_steps(args.warmup)) / args.warmup
    # pre-roll into the thermal / power-cap steady state: every N is then timed in the same clock regime
    # (the step count is derived from a max-over-ranks value, so all ranks run the same number of steps)
    n_pre = int(min(5000, max(0, np.ceil(args.preroll * 1e3 / est)))) if (est > 0 and args.preroll > 0) else 0
    preroll_s = hostcomm.max_over_ranks(timed_steps(n_pre)) * 1e-3 if n_pre else 0.0
    hostcomm.barrier()
    n0 = abi.launch_count()
    t0 = time.time()
    ms = timed_steps(args.steps)
    t1 = time.time()
    launches = abi.launch_count() - n0
    hostcomm.barrier()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_max = hostcomm.max_over_ranks(ms)
    ms_step = ms_max / args.steps
    # ---------------- how much of the halo exchange is hidden (N > 1) ----------------
    halo_info = None
    if halo is not None:
        comm = []
        for _ in range(5):
            df.step(f, mode=1, halo=halo, sync=True)
            comm.append(halo.comm_ms())
        comm_ms = hostcomm.max_over_ranks(sorted(comm)[len(comm) // 2])
        ms_with = hostcomm.max_over_ranks(timed_steps(args.steps)) / args.steps
        abi.tune("HALO_SKIP", 1)               # same launches and events, no pack / send / recv (stale halo rows)
        timed_steps(3)
        ms_without = hostcomm.max_over_ranks(timed_steps(args.steps)) / args.steps
        abi.tune("HALO_SKIP", 0)
        halo_info = {"comm_ms": comm_ms, "exposed_ms": ms_with - ms_without, "step_ms_with": ms_with, "step_ms_without_exchange": ms_without,
                     "floor_ms": floor_ms,
                     "how": "comm_ms: CUDA events on the halo stream around pack + grouped ncclSend/ncclRecv (median of 5 steps, max over ranks); "
                            "exposed_ms: step time minus the time of the same launches with the exchange skipped (knob HALO_SKIP); "
                            "floor_ms: the step on the CORE2-size mesh split the same way (launch + NCCL latency floor)"}
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
    # (1) one model TIME STEP of E2E_TRACERS tracers (T, S + 10 passive: BASELINE.json config 5's batch, the way
    #     FESOM2 calls fct_ale): hnode / hnode_new -- inputs of the time step, constant across its tracers --
    #     uploaded once, the six per-tracer arrays uploaded and the two tendencies downloaded for every tracer,
    #     all inside the timed region, dense page-locked host arrays.  This is `e2e.value`, per tracer step.
    # (2) the same for a caller that keeps its columns in the packed level storage on the host (70 % of the bytes).
    # (3) the round-1 pattern: every tracer step uploads all eight arrays (`single_tracer`).
    T_E2E = max(1, args.e2e_tracers)
    per_tracer = [k for k in df.STEP_INPUTS if k not in ("hnode", "hnode_new")]
    assert df.host_step(f, f, mode=1, halo=halo) == 10       # warm-up
    hostcomm.barrier()
    ta = time.perf_counter()
    assert df.host_steps_batch(f, f, T_E2E, mode=1, halo=halo) == 10     # ends with await_stream_: tendencies on the host
    tb = time.perf_counter()
    e2e_s = hostcomm.max_over_ranks((tb - ta) / T_E2E)
    hb = sum(df.link_bytes(k, getattr(f, k), True) for k in per_tracer) + sum(df.link_bytes(k, getattr(f, k), True) for k in ("hnode", "hnode_new")) / T_E2E
    db = sum(df.link_bytes(k, getattr(f, k), False) for k in df.STEP_RESULTS)
    h2d, d2h = hostcomm.sum_over_ranks(hb), hostcomm.sum_over_ranks(db)
    log(f"e2e (time step of {T_E2E} tracers, dense host arrays): {e2e_s * 1e3:.1f} ms per tracer step ({time.time() - t_setup:.1f}s)")
    # (3) single tracer steps, everything uploaded every step
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hostcomm.barrier()
    ta = time.perf_counter()
    assert df.host_steps(f, f, e2e_steps, mode=1, halo=halo) == 10
    tb = time.perf_counter()
    single_s = hostcomm.max_over_ranks((tb - ta) / e2e_steps)
    shb, sdb = df.host_step_bytes(f)
    single = {"value": Sn_total / single_s, "unit": UNIT, "ms_per_step": single_s * 1e3, "steps": e2e_steps,
              "h2d_bytes_per_step": int(hostcomm.sum_over_ranks(shb)), "d2h_bytes_per_step": int(hostcomm.sum_over_ranks(sdb)),
              "api": "per tracer step: fct_ale_field_upload_ x8 / fct_ale_step_ / fct_ale_field_download_ x2 (round-1 pattern)"}
    # (2) packed host arrays
    packed_host = None
    if df.packed:
        ph, nbytes_up, nbytes_dn = {}, 0, 0
        for k in df.STEP_INPUTS:
            # the packed host image of the array, produced by the device (upload dense, download packed):
            # set-up only, outside every timed region
            kind = "edge" if k == "fct_adf_h" else "node"
            ph[k] = abi.pinned_empty(int(plan.packed_columns(kind)[-1]))
            df.upload_field(k, getattr(f, k))
            df.download_packed(k, ph[k])
            df.stream.sync()
            nbytes_up += ph[k].nbytes / (T_E2E if k in ("hnode", "hnode_new") else 1)
        for k in df.STEP_RESULTS:
            ph["out_" + k] = ph[k]
            nbytes_dn += ph[k].nbytes
        assert df.host_steps_batch(f, f, 1, mode=1, halo=halo, packed_host=ph) == 10
        hostcomm.barrier()
        ta = time.perf_counter()
        assert df.host_steps_batch(f, f, T_E2E, mode=1, halo=halo, packed_host=ph) == 10
        tb = time.perf_counter()
        pk_s = hostcomm.max_over_ranks((tb - ta) / T_E2E)
        packed_host = {"value": Sn_total / pk_s, "unit": UNIT, "ms_per_step": pk_s * 1e3, "tracers": T_E2E,
                       "h2d_bytes_per_step": int(hostcomm.sum_over_ranks(nbytes_up)), "d2h_bytes_per_step": int(hostcomm.sum_over_ranks(nbytes_dn)),
                       "api": "the same time step for a caller whose host arrays are already in the packed level storage: "
                              "fct_ale_field_upload_packed_ / fct_ale_field_download_packed_ (one contiguous copy per array, no repack)"}
        del ph
        log(f"e2e packed host arrays: {pk_s * 1e3:.1f} ms per tracer step ({time.time() - t_setup:.1f}s)")
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

    # ---------------- the reference's own GPU kernels on the same mesh (N = 1) ----------------
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        gpu_ref = gpu_reference(m, abi)
        if "step_ms" in gpu_ref:
            gpu_ref["value"] = Sn_total / (gpu_ref["step_ms"] * 1e-3)
            gpu_ref["speedup_of_this_repo"] = gpu_ref["step_ms"] / ms_step
            log(f"reference GPU kernels: {gpu_ref['step_ms']:.2f} ms/step, this repo {ms_step:.2f} ms ({time.time() - t_setup:.1f}s)")

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
                "config": config,
                "details": {"partitions": world, "mode": mode_name, "device": devname},
                "hbm": {"alg_GBs_per_gpu": bytes_alg_total / world / ms_step / 1e6, "frac_of_peak": bytes_alg_total / world / ms_step / 1e6 / peak},
                "clocks": clocks,
                "e2e": {"value": Sn_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_s * 1e3, "steps": T_E2E, "tracers_per_time_step": T_E2E,
                        "api": "one model time step of 12 tracers on dense page-locked host arrays: fct_ale_field_upload_ of hnode, hnode_new once per time step, "
                               "then per tracer fct_ale_field_upload_ x6 (ttf, fct_LO, fct_adf_v, fct_adf_h, del_ttf_adv*) / fct_ale_step_ / fct_ale_field_download_ x2 "
                               "(download of tracer k on a second stream, overlapping the upload of tracer k+1), await_stream_ at the end; per tracer step",
                        "single_tracer": single, "packed_host": packed_host,
                        "reference_sequence": refseq},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "gpu_reference": gpu_ref, "parity": parity, "digest": digest, "preroll_s": preroll_s, "preroll_steps": n_pre, "halo": halo_info,
                "small_mesh_step_ms": floor_ms, "setup_s": time.time() - t_setup}
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
    if _EXIT:
        sys.exit(_EXIT[0])


_RESULT = []
_EXIT = []


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FCT_BENCH_WORKLOAD", "ng5"))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-tracers", type=int, default=12, help="tracers of the time step the e2e leg times (T, S + 10 passive)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the timing of the reference's own GPU kernels")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs only)")
    ap.add_argument("--preroll", type=float, default=1.5, help="seconds of untimed steps before the timed region")
    ap.add_argument("--reference-budget", type=float, default=150.0, help="--impl reference: seconds of CPU steps after which no further step is started")
    ap.add_argument("--no-refseq", action="store_true", help="skip the timing of the reference library's call sequence")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
