"""GPU-vs-GPU baseline: the reference's own kernels rebuilt for sm_100 (baseline/build_ref_gpu.sh ->
baseline/_ref/libref_gpu_kernels.so, launched with the reference driver's geometry) against this
repository's stage kernels and fused phases on the same synthetic mesh, device-resident, CUDA events.
Timing only: the reference's b1_horizontal launch (grid = nodes instead of edges,
/root/reference/src/fesom2-accelerate.cu:327) makes its results unusable for parity.
usage: gpu_reference_bench.py [core2|ng5|dart|NXxNYxNL] ...   (prints a markdown table per workload)"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

REF_SO = os.path.join(ROOT, "baseline", "_ref", "libref_gpu_kernels.so")
REF_NAMES = ["a1", "a2", "a3", "b1v", "b1h (grid N, as shipped)", "b1h", "b2", "b3v", "b3h", "cv", "ch", "sequence a1..c"]


def reference_kernels_ms(m, reps=5):
    """-> dict stage -> ms of the reference's kernels on mesh m (None when the baseline library is absent)."""
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    ms = (C.c_double * 12)()
    st = C.c_int()
    ci, ip = abi.ci, abi.iptr
    lib.ref_gpu_bench_(ci(m.myDim_nod2D), ci(m.eDim_nod2D), ci(m.myDim_elem2D), ci(m.myDim_edge2D), ci(m.nl),
                       ip(m.nlevels_nod2D), ip(m.nlevels_elem), ip(m.elem2D_nodes.reshape(-1)), ip(m.nod_in_elem2D_num),
                       ip(m.nod_in_elem2D.reshape(-1)), ci(m.nod_in_elem2D_dim), ip(m.edges.reshape(-1)),
                       ip(m.edge_tri.reshape(-1)), ci(reps), ms, C.byref(st))
    if st.value != 0:
        return None
    return dict(zip(REF_NAMES, list(ms)))


def main():
    abi.device_info()
    for w in sys.argv[1:] or ["core2"]:
        if w in mesh.WORKLOADS:
            m = mesh.make_workload(w)
        else:
            nx, ny, nl = [int(x) for x in w.split("x")]
            m = mesh.make_mesh(nx, ny, nl)
        Sn = m.S_n()
        ref = reference_kernels_ms(m)
        f = mesh.fast_fields(m, with_uv=True) if m.myDim_nod2D > 500000 else mesh.make_fields(m, poison=False)
        plan = harness.DevicePlan(m)
        e0, e1 = abi.Event(), abi.Event()

        def timeit(df, fn, reps=5):
            fn(); fn()
            df.stream.sync(); e0.record(df.stream)
            for _ in range(reps): fn()
            e1.record(df.stream)
            return e1.ms_since(e0) / reps
        df = harness.DeviceFields(plan, 1, with_uv=True)
        df.upload(f, outputs=False)
        ours = {s: timeit(df, lambda: df.stage(s, f, sync=False)) for s in ("a1", "a2", "a3", "b1v", "b1h", "b2", "b3v", "b3h", "cv", "ch")}
        ours["sequence a1..c"] = timeit(df, lambda: df.step(f, mode=0, sync=False))
        df.free()
        dp = harness.DeviceFields(plan, 1, packed=plan.packed_ok)
        dp.upload(f, outputs=False)
        fused = timeit(dp, lambda: dp.step(f, mode=1, sync=False), reps=10)
        dp.free(); plan.free()
        print(f"\n### {w}: {m.myDim_nod2D} nodes, {m.myDim_elem2D} elements, {m.myDim_edge2D} edges, nl = {m.nl}, {Sn} node-level updates\n")
        print("| stage | reference kernel @ sm_100 (ms) | this repo, stage kernel (ms) | ratio |")
        print("|---|---|---|---|")
        for k in REF_NAMES:
            r = ref.get(k) if ref else None
            o = ours.get(k)
            print(f"| {k} | {r:.3f} | " + (f"{o:.3f} | {r / o:.2f}x |" if o else "- | - |") if r is not None else f"| {k} | n/a | {o} | - |")
        if ref:
            print(f"\nfull step: reference kernels {ref['sequence a1..c']:.3f} ms ({Sn / ref['sequence a1..c'] / 1e6:.2f} G updates/s), "
                  f"this repo staged {ours['sequence a1..c']:.3f} ms, this repo fused (fct_ale_step_ mode 1, packed) {fused:.3f} ms "
                  f"({Sn / fused / 1e6:.2f} G updates/s) = {ref['sequence a1..c'] / fused:.1f}x the reference kernels")


if __name__ == "__main__":
    main()
