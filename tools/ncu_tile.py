"""Few launches of chosen stages for ncu.
usage: ncu_tile.py <workload | NXxNYxNL> <stage,stage,...> [reps] [tracers] [packed]"""
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
packed = len(sys.argv) > 5 and sys.argv[5] == "packed"
big = m.myDim_nod2D > 500000
f = mesh.fast_fields(m, seed=1) if big else mesh.make_fields(m, with_uv=False, poison=False)
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, T, with_uv=("a2" in stages or "a3" in stages), packed=packed)
for t in range(T):
    df.upload(f, tracer=t, static=(t == 0), outputs=False)
for _ in range(reps):
    for s in stages:
        df.stage(s, f)
print("S_n", m.S_n(), "S_g", m.S_g(), "N", m.myDim_nod2D, "bytes_alg", m.bytes_alg())
