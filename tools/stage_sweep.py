"""Per-stage sweep a1 -> c on the multi-tracer CORE2 batch (BASELINE.json config 5): every stage
kernel (one per reference kernel) and the fused phases, timed with CUDA events, with the bytes of
the reference's own per-kernel traffic models next to them.  Run it under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` for the DRAM bytes.
usage: stage_sweep.py [workload] [tracers] [reps]"""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

wl = sys.argv[1] if len(sys.argv) > 1 else "core2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
m = mesh.make_workload(wl)
fs = [mesh.make_fields(m, seed=1 + t, with_uv=False, poison=False) for t in range(T)]
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, T, with_uv=True)
for t in range(T):
    df.upload(fs[t], tracer=t, static=(t == 0), outputs=False)
f = fs[0]
N, E, G = m.myDim_nod2D, m.myDim_elem2D, m.myDim_edge2D
Sn, Sg, Se = m.S_n(), m.S_g(), int((m.nlevels_elem.astype("int64") - 1).sum())
ring = int(m.nod_in_elem2D_num.sum())
L = m.L
# the reference's per-kernel byte models (kernels/fct_ale_*.py), per tracer
model = {
    "a1": 4 * N + 32 * Sn, "a2": 16 * E + 64 * Se + 16 * (E * L - Se), "a3": 20 * (Sn + ring * Sn // max(N, 1)) + 32 * Sn,
    "b1v": 4 * N + 32 * Sn, "b1h": 40 * Sg, "b2": 4 * N + 56 * Sn, "b3v": 28 * N + 48 * Sn, "b3h": 48 * Sg,
    "cv": 4 * N + 72 * Sn, "ch": 56 * Sg,
}
algA, algB = 8 * (8 * Sn + Sg) + 16 * N, 8 * (13 * Sn + 2 * Sg)
e0, e1 = abi.Event(), abi.Event()
def timeit(fn):
    for _ in range(2): fn()
    df.stream.sync(); e0.record(df.stream)
    for _ in range(reps): fn()
    e1.record(df.stream)
    return e1.ms_since(e0) / reps
print(f"# {wl}: N={N} E={E} G={G} nl={m.nl} tracers={T} S_n={Sn} S_g={Sg}")
print("| stage | kernel | ms | reference byte model GB | model GB/s | % of 6547 GB/s |")
print("|---|---|---|---|---|---|")
tot = 0.0
for s in ["a1", "a2", "a3", "b1v", "b1h", "b2", "b3v", "b3h", "cv", "ch"]:
    ms = timeit(lambda: df.stage(s, f, sync=False)); tot += ms
    b = model[s] * T
    print(f"| {s} | k_{s}<2> | {ms:.3f} | {b/1e9:.3f} | {b/ms/1e6:.0f} | {b/ms/1e6/65.472:.1f} |")
for s, d in (("b1h_atomic", "b1h"), ("ch_atomic", "ch")):
    ms = timeit(lambda: df.stage(s, f, sync=False))
    b = model[d] * T
    print(f"| {s} (alternative: edge scatter, fp64 atomicAdd) | k_{s}<2> | {ms:.3f} | {b/1e9:.3f} | {b/ms/1e6:.0f} | {b/ms/1e6/65.472:.1f} |")
print(f"| staged chain | ten launches | {tot:.3f} | {sum(model.values())*T/1e9:.3f} | {sum(model.values())*T/tot/1e6:.0f} | {sum(model.values())*T/tot/1e6/65.472:.1f} |")
for s, alg in (("phaseA_warp", algA), ("phaseB_warp", algB)):
    ms = timeit(lambda: df.stage(s, f, sync=False))
    print(f"| {s} | k_phase_warp (algorithmic bytes) | {ms:.3f} | {alg*T/1e9:.3f} | {alg*T/ms/1e6:.0f} | {alg*T/ms/1e6/65.472:.1f} |")
ms = timeit(lambda: df.step(f, mode=1, sync=False))
print(f"| fused step (mode 1) | 2 launches | {ms:.3f} | {(algA+algB)*T/1e9:.3f} | {(algA+algB)*T/ms/1e6:.0f} | {(algA+algB)*T/ms/1e6/65.472:.1f} |")
print(f"\nfused step: {Sn*T/ms/1e6:.2f} G node-level updates/s for {T} tracers")
df.free(); plan.free()
