"""Few launches of chosen stages for ncu."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
w = sys.argv[1]
stages = sys.argv[2].split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T = int(sys.argv[4]) if len(sys.argv) > 4 else 1
m = mesh.make_workload(w) if w in mesh.WORKLOADS else mesh.make_mesh(*[int(x) for x in w.split("x")])
f = mesh.make_fields(m, with_uv=False, poison=False)
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, T, with_uv=("a2" in stages or "a3" in stages))
for t in range(T):
    df.upload(f, tracer=t, static=(t == 0), outputs=False)
for _ in range(reps):
    for s in stages:
        df.stage(s, f)
print("S_n", m.S_n(), "S_g", m.S_g(), "N", m.myDim_nod2D, "bytes_alg", m.bytes_alg())
