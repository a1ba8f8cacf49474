"""Per-tile timeline of the warp-item kernels' staging pipeline (knob WT_TRACE): SM-clock stamps of CTA 0.
usage: trace_pipeline.py NXxNYxNL [A|B] ["KNOB=V,KNOB=V"]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = [int(x) for x in sys.argv[1].split("x")]
phase = sys.argv[2] if len(sys.argv) > 2 else "A"
for kv in (sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[3] else []):
    abi.tune(kv.split("=")[0], int(kv.split("=")[1]))
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m) if m.myDim_nod2D > 500000 else mesh.make_fields(m, with_uv=False, poison=False)
abi.tune("VERBOSE", 1)
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, 1, with_uv=False, packed=True)
df.upload(f, outputs=False)
stage = "phaseA_warp" if phase == "A" else "phaseB_warp"
e0, e1 = abi.Event(), abi.Event()
for _ in range(3):
    df.stage("phaseA_warp", f); df.stage("phaseB_warp", f)
abi.tune("WT_TRACE", 1 if phase == "A" else 2)
e0.record(df.stream)
df.stage(stage, f, sync=False)
e1.record(df.stream)
ms = e1.ms_since(e0)
abi.tune("WT_TRACE", 0)
cap = 10 * 4096
buf = np.zeros(cap, np.int64)
slots, st = C.c_int(), C.c_int()
abi.load().fct_ale_trace_read_(buf.ctypes.data_as(C.POINTER(C.c_longlong)), abi.ci(cap), C.byref(slots), C.byref(st))
assert st.value == 0
t = buf.reshape(-1, slots.value).astype(np.float64)
n = int((t[:, 6] > 0).sum())
t = t[:n]
t0 = t[t > 0].min()
span = t.max() - t0
ghz = span / (ms * 1e6)           # cycles per ns, from the kernel's event time (includes launch overhead: approximate)
print(f"{stage}: {ms*1e3:.1f} us, CTA 0 ran {n} tiles, span {span:.0f} cycles -> ~{ghz:.2f} GHz; times below in ns (median / p90)")
ns = lambda c: c / ghz


def stat(name, v):
    v = v[np.isfinite(v)]
    if v.size:
        print(f"  {name:58s} {np.median(ns(v)):8.0f} / {np.percentile(ns(v), 90):8.0f}")


S = 2 if (nl >= 60) else 3
S = int(os.environ.get("TRACE_STAGES", S))
last_leave = np.maximum(t[:, 7], t[:, 9])
first_ready = np.minimum(t[:, 6], t[:, 8])
period = np.diff(t[:, 6])
stat("tile period (first consumer warp, ready -> next ready)", period)
stat("compute: first consumer warp in the tile (6 -> 7)", t[:, 7] - t[:, 6])
stat("compute: last consumer warp in the tile (8 -> 9)", t[:, 9] - t[:, 8])
stat("consumer wait: leaves tile k-1 -> tile k ready (warp 0)", t[1:, 6] - t[:-1, 7])
stat("consumer wait: leaves tile k-1 -> tile k ready (last warp)", t[1:, 8] - t[:-1, 9])
if n > S:
    stat("stage idle: last warp left tile k-S -> fetcher sees it empty", t[S:, 0] - last_leave[:-S])
stat("fetcher: stage empty -> blob copy issued (0 -> 1)", t[:, 1] - t[:, 0])
stat("issuer: stage empty (fetcher) -> row copies start (0 -> 2)", t[:, 2] - t[:, 0])
stat("issuer: row copies start -> all issued (2 -> 3)", t[:, 3] - t[:, 2])
if phase == "A":
    stat("rows in flight: copies issued -> rows landed (3 -> 4)", t[:, 4] - t[:, 3])
    stat("a1 conversion (4 -> 5)", t[:, 5] - t[:, 4])
    stat("REFILL: stage empty -> a1 done (0 -> 5)", t[:, 5] - t[:, 0])
    stat("slack: a1 done -> first consumer enters (5 -> min(6,8))", first_ready - t[:, 5])
else:
    stat("REFILL: stage empty -> first consumer enters (0 -> min(6,8))", first_ready - t[:, 0])
if n > S:
    stat("stage occupancy: ready -> last warp leaves (min(6,8) -> max(7,9))", last_leave - first_ready)
