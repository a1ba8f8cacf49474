"""Tuning sweep of the warp-item kernels on one mesh: tile nodes x stage KB (0: default) x ring depth x consumer warps.
usage: sweep_warp.py NXxNYxNL "TN:CAPKB:STAGES:WARPS_A:WARPS_B:ISSUERS,..." """
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "400x317x48").split("x")]
cfgs = [tuple(int(v) for v in c.split(":")) for c in (sys.argv[2] if len(sys.argv) > 2 else "0:0:3:0:0:0").split(",")]
reps = 10
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m) if m.myDim_nod2D > 500000 else mesh.make_fields(m, with_uv=False, poison=False)
Sn, Sg = m.S_n(), m.S_g()
algA, algB = 8 * (8 * Sn + Sg) + 16 * m.myDim_nod2D, 8 * (13 * Sn + 2 * Sg)
print(f"N={m.myDim_nod2D} nl={nl} S_n={Sn} S_g={Sg}", flush=True)
e0, e1 = abi.Event(), abi.Event()
abi.tune("VERBOSE", 1)
for TN, cap, nch, wa, wb, npw in cfgs:
    abi.tune("WT_NODES", TN); abi.tune("WT_SMEM", cap * 1024); abi.tune("WT_STAGES", nch)
    abi.tune("WT_WARPS_A", wa); abi.tune("WT_WARPS_B", wb); abi.tune("WT_ISSUERS", npw)
    abi.tune("TILE", 0)
    t0 = time.time()
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=bool(int(os.environ.get('SWEEP_PACKED', '1'))))
    df.upload(f, outputs=False)
    def timeit(fn):
        for _ in range(3): fn()
        df.stream.sync(); e0.record(df.stream)
        for _ in range(reps): fn()
        e1.record(df.stream)
        return e1.ms_since(e0) / reps
    try:
        # SWEEP_OPTS: scheduling options of the kernels (knob WT_OPT, read at every launch)
        for opt in [int(v) for v in os.environ.get("SWEEP_OPTS", "-1").split(",")]:
            if opt >= 0:
                abi.tune("WT_OPT", opt)
            tot = 0
            for stage, alg in (("phaseA_warp", algA), ("phaseB_warp", algB)):
                ms = timeit(lambda: df.stage(stage, f, sync=False))
                tot += ms
                print(f"  TN={TN:3d} cap={cap:3d}K stages={nch} warps={wa}/{wb} issuers={npw} opt={opt} {stage}: {ms*1e3:8.1f} us  {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/65.472:5.1f}%", flush=True)
            print(f"  -> step {tot*1e3:8.1f} us  {Sn/tot/1e6:6.2f} G upd/s  {(algA+algB)/tot/1e6/65.472:5.1f}% of 6547 GB/s (plan {time.time()-t0:.1f}s)", flush=True)
    except abi.AbiError as ex:
        print("  cfg", TN, cap, nch, wa, wb, npw, "failed:", ex)
    df.free(); plan.free()
