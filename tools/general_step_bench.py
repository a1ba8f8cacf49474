"""fct_ale_step_general_: fused fast path (packed fields, two launches) against the stage kernels
(padded fields) for the vlimit 3 and iterative branches, CUDA-event times on one mesh.
usage: general_step_bench.py [NXxNYxNL]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")
nx, ny, nl = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024x780x80").split("x")]
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m)
Sn = m.S_n()
plan = harness.DevicePlan(m)
e0, e1 = abi.Event(), abi.Event()
print(f"N={m.myDim_nod2D} nl={nl} S_n={Sn}")
for packed in (True, False):
    df = harness.DeviceFields(plan, 1, with_uv=not packed, packed=packed)
    df.upload(f, outputs=False)
    for vlimit, it in ((1, False), (3, False), (1, True)):
        f.vlimit, f.iter_yn = vlimit, it
        for _ in range(2): df.step_general(f, sync=False)
        df.stream.sync(); e0.record(df.stream)
        for _ in range(5): df.step_general(f, sync=False)
        e1.record(df.stream)
        ms = e1.ms_since(e0) / 5
        print(f"{'packed, fused (2 launches)' if packed else 'padded, stage kernels    '} vlimit={vlimit} iter_yn={int(it)}: {ms:.3f} ms  {Sn/ms/1e6:.2f} G node-level updates/s", flush=True)
    df.free()
plan.free()
