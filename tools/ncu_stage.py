"""One pass over every stage + the fused phases on a mesh, for ncu (few launches, no timing)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
w = sys.argv[1] if len(sys.argv) > 1 else "core2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m = mesh.make_workload(w)
f = mesh.make_fields(m, with_uv=False, poison=False)
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, T, with_uv=True)
for t in range(T):
    df.upload(f, tracer=t, static=(t == 0), outputs=False)
for _ in range(reps):
    for s in ["a1", "a2", "a3", "b1v", "b1h", "b2", "b3v", "b3h", "cv", "ch", "phaseA", "phaseB"]:
        df.stage(s, f)
print("S_n", m.S_n(), "S_g", m.S_g(), "N", m.myDim_nod2D, "bytes_alg", m.bytes_alg())
