"""Tuning sweep of the tile kernels on one mesh: (tile nodes, iters) x kernel variants."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024x780x70").split("x")]
tiles = [tuple(int(v) for v in c.split(":")) for c in (sys.argv[2] if len(sys.argv) > 2 else "12:1,16:1,24:2").split(",")]
va = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0,1,2,3,4,5").split(",")]
vb = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "0,1,2,3,4,5").split(",")]
reps = 5
m = mesh.make_mesh(nx, ny, nl)
f = mesh.make_fields(m, with_uv=False, poison=False)
Sn, Sg = m.S_n(), m.S_g()
algA, algB = 8 * (8 * Sn + Sg) + 16 * m.myDim_nod2D, 8 * (13 * Sn + 2 * Sg)
print(f"N={m.myDim_nod2D} nl={nl} S_n={Sn} S_g={Sg}", flush=True)
e0, e1 = abi.Event(), abi.Event()
for TN, it in tiles:
    abi.tune("TILE_NODES", TN); abi.tune("TILE_ITERS", it)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False)
    df.upload(f, outputs=False)
    def timeit(fn):
        for _ in range(2): fn()
        df.stream.sync(); e0.record(df.stream)
        for _ in range(reps): fn()
        e1.record(df.stream)
        return e1.ms_since(e0) / reps
    try:
        for key, stage, alg, vs in (("TILE_VARIANT_A", "phaseA_tile", algA, va), ("TILE_VARIANT_B", "phaseB_tile", algB, vb)):
            for v in vs:
                abi.tune(key, v)
                ms = timeit(lambda: df.stage(stage, f, sync=False))
                print(f"  TN={TN:2d} it={it} {stage} v{v}: {ms*1e3:8.1f} us  {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/65.472:5.1f}%", flush=True)
    except abi.AbiError as ex:
        print("  TN", TN, it, "failed:", ex)
    df.free(); plan.free()
