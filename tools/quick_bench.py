"""Scratch timing of every stage and of the fused phases (CUDA events), device-resident."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

def run(nx, ny, nl, T=1, reps=10, label=""):
    t0 = time.time()
    m = mesh.make_mesh(nx, ny, nl)
    f = mesh.make_fields(m, with_uv=False, poison=False)
    t1 = time.time()
    plan = harness.DevicePlan(m)
    t2 = time.time()
    df = harness.DeviceFields(plan, T, with_uv=True)
    for t in range(T):
        df.upload(f, tracer=t, static=(t == 0), outputs=False)
    Sn, Sg = m.S_n(), m.S_g()
    balg = m.bytes_alg() * T
    print(f"== {label} N={m.myDim_nod2D} E={m.myDim_elem2D} G={m.myDim_edge2D} nl={nl} T={T} S_n={Sn} S_g={Sg} bytes_alg={balg/1e9:.3f} GB  (mesh {t1-t0:.1f}s plan {t2-t1:.1f}s)")
    e0, e1 = abi.Event(), abi.Event()
    def timeit(fn, reps=reps):
        for _ in range(3): fn()
        df.stream.sync()
        e0.record(df.stream)
        for _ in range(reps): fn()
        e1.record(df.stream)
        return e1.ms_since(e0) / reps
    tot = 0
    for s in ["a1","a2","a3","b1v","b1h","b2","b3v","b3h","cv","ch"]:
        ms = timeit(lambda: df.stage(s, f, sync=False))
        tot += ms
        print(f"   stage {s:4s} {ms*1e3:9.1f} us")
    print(f"   staged sum {tot*1e3:9.1f} us -> {Sn*T/tot/1e6:8.2f} G upd/s, alg {balg/tot/1e6:8.1f} GB/s")
    Sg8 = 8 * Sg * T
    Sn8 = 8 * Sn * T
    algA, algB = 8 * Sn8 + Sg8 + 16 * m.myDim_nod2D * T, 13 * Sn8 + 2 * Sg8
    for s in ["phaseA", "phaseB", "phaseA_tile", "phaseB_tile"]:
        ms = timeit(lambda: df.stage(s, f, sync=False))
        alg = algA if "A" in s else algB
        print(f"   fused {s:12s} {ms*1e3:9.1f} us   alg {alg/ms/1e6:8.1f} GB/s = {alg/ms/1e6/6547.2*100:5.1f}%")
    for mode in (0, 2, 1):
        ms = timeit(lambda: df.step(f, mode=mode, sync=False))
        print(f"   step mode {mode}: {ms*1e3:9.1f} us -> {Sn*T/ms/1e6:8.2f} G upd/s, alg {balg/ms/1e6:8.1f} GB/s = {balg/ms/1e6/6547.2*100:5.1f}% of 6547 GB/s")
    df.free(); plan.free()

if __name__ == "__main__":
    print(abi.device_info())
    which = sys.argv[1:] or ["core2", "mid"]
    for w in which:
        if w == "core2": run(400, 317, 48, label="core2")
        elif w == "core2x12": run(400, 317, 48, T=12, label="core2 x12 tracers")
        elif w == "mid": run(1536, 1204, 70, label="mid (NG5/4)")
        elif w == "dart": run(2048, 1560, 80, label="dart")
        elif w == "ng5": run(3072, 2408, 70, reps=5, label="ng5")
