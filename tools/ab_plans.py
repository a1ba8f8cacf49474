"""Interleaved A/B of kernel configurations on one mesh: every configuration gets its own plan
(tile sizes follow the ring depth), the timing alternates between them in thermal steady state.
usage: ab_plans.py NXxNYxNL "STAGES:OPT[:ISSUERS],..." [rounds] [packed=1]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = [int(x) for x in sys.argv[1].split("x")]
cfgs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[2].split(",")]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
packed = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m) if m.myDim_nod2D > 500000 else mesh.make_fields(m, with_uv=False, poison=False)
Sn, Sg = m.S_n(), m.S_g()
algA, algB = 8 * (8 * Sn + Sg) + 16 * m.myDim_nod2D, 8 * (13 * Sn + 2 * Sg)
state = []
cfgs = [c if len(c) > 2 else c + (0,) for c in cfgs]
for stages, opt, iss in cfgs:
    abi.tune("WT_STAGES", stages)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=packed)
    df.upload(f, outputs=False)
    state.append((plan, df))
e0, e1 = abi.Event(), abi.Event()
tot = [[0.0, 0.0] for _ in cfgs]
for r in range(rounds + 1):
    for i, (stages, opt, iss) in enumerate(cfgs):
        abi.tune("WT_STAGES", stages); abi.tune("WT_OPT", opt); abi.tune("WT_ISSUERS", iss)
        plan, df = state[i]
        res = []
        for stage in ("phaseA_warp", "phaseB_warp"):
            for _ in range(3): df.stage(stage, f, sync=False)
            df.stream.sync(); e0.record(df.stream)
            for _ in range(10): df.stage(stage, f, sync=False)
            e1.record(df.stream)
            res.append(e1.ms_since(e0) / 10)
        if r > 0:
            tot[i][0] += res[0]; tot[i][1] += res[1]
        print(f"round {r} stages={stages} opt={opt} issuers={iss}: A {res[0]*1e3:8.1f} us ({algA/res[0]/1e6/65.472:5.1f}%)  B {res[1]*1e3:8.1f} us ({algB/res[1]/1e6/65.472:5.1f}%)", flush=True)
print(f"N={m.myDim_nod2D} nl={nl} packed={packed}")
for i, (stages, opt, iss) in enumerate(cfgs):
    a, b = tot[i][0] / rounds, tot[i][1] / rounds
    print(f"stages={stages} opt={opt} issuers={iss}: phase A {a*1e3:.1f} us ({algA/a/1e6/65.472:.1f}%), phase B {b*1e3:.1f} us ({algB/b/1e6/65.472:.1f}%), step {(a+b)*1e3:.1f} us ({(algA+algB)/(a+b)/1e6/65.472:.1f}%)")
