"""A/B of two builds of the library in one process, interleaved (the GPU runs under its power cap:
only interleaved repeats are comparable).  usage: ab_libs.py NXxNYxNL libA.so libB.so [rounds]"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = [int(x) for x in sys.argv[1].split("x")]
paths = [os.path.abspath(p) for p in sys.argv[2:4]]
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
libs = [C.CDLL(p) for p in paths]
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m) if m.myDim_nod2D > 500000 else mesh.make_fields(m, with_uv=False, poison=False)
Sn, Sg = m.S_n(), m.S_g()
algA, algB = 8 * (8 * Sn + Sg) + 16 * m.myDim_nod2D, 8 * (13 * Sn + 2 * Sg)
state = []
for lib in libs:
    abi._lib = lib
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=True)
    df.upload(f, outputs=False)
    state.append((plan, df, abi.Event(), abi.Event()))
tot = [[0.0, 0.0], [0.0, 0.0]]
for r in range(rounds + 1):
    for i, lib in enumerate(libs):
        abi._lib = lib
        plan, df, e0, e1 = state[i]
        res = []
        for stage in ("phaseA_warp", "phaseB_warp"):
            for _ in range(3): df.stage(stage, f, sync=False)
            df.stream.sync(); e0.record(df.stream)
            for _ in range(10): df.stage(stage, f, sync=False)
            e1.record(df.stream)
            res.append(e1.ms_since(e0) / 10)
        if r > 0:       # round 0 warms the GPU up
            tot[i][0] += res[0]; tot[i][1] += res[1]
        print(f"round {r} lib {'AB'[i]}: phase A {res[0]*1e3:8.1f} us ({algA/res[0]/1e6/65.472:5.1f}%)  phase B {res[1]*1e3:8.1f} us ({algB/res[1]/1e6/65.472:5.1f}%)", flush=True)
for i in range(2):
    a, b = tot[i][0] / rounds, tot[i][1] / rounds
    print(f"lib {'AB'[i]} {os.path.basename(paths[i])}: phase A {a*1e3:.1f} us ({algA/a/1e6/65.472:.1f}%), phase B {b*1e3:.1f} us ({algB/b/1e6/65.472:.1f}%), step {(a+b)*1e3:.1f} us ({(algA+algB)/(a+b)/1e6/65.472:.1f}%)")
