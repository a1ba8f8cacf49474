"""BASELINE.json config 5: T, S + 10 passive tracers on the CORE2-size mesh in ONE pair of launches,
packed level storage (the fast path), CUDA-event time of fct_ale_step_ mode 1.
usage: multi_tracer_bench.py [workload] [tracers] [reps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

wl = sys.argv[1] if len(sys.argv) > 1 else "core2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
m = mesh.make_workload(wl)
fs = [mesh.make_fields(m, seed=1 + t, with_uv=False, poison=False) for t in range(T)]
plan = harness.DevicePlan(m)
Sn, Sg = m.S_n(), m.S_g()
alg = (8 * (21 * Sn + 3 * Sg) + 16 * m.myDim_nod2D) * T
e0, e1 = abi.Event(), abi.Event()
for packed in (True, False):
    df = harness.DeviceFields(plan, T, with_uv=False, packed=packed)
    for t in range(T):
        df.upload(fs[t], tracer=t, static=(t == 0), outputs=False)
    for rnd in range(3):
        for _ in range(5): df.step(fs[0], mode=1, sync=False)
        df.stream.sync(); e0.record(df.stream)
        for _ in range(reps): df.step(fs[0], mode=1, sync=False)
        e1.record(df.stream)
        ms = e1.ms_since(e0) / reps
        print(f"{wl} x {T} tracers, {'packed' if packed else 'padded'}: {ms:.3f} ms per step of all tracers, {Sn*T/ms/1e6:.2f} G node-level updates/s, "
              f"{alg/ms/1e6:.0f} GB/s algorithmic = {alg/ms/1e6/65.472:.1f}% of 6547 GB/s", flush=True)
    df.free()
plan.free()
