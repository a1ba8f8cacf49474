"""Host link timing of the packed fields: per-field upload / download and the end-to-end host step,
direct (kernel over mapped page-locked memory, slots only) against staged (contiguous copy of the
dense array + repack kernel).  CUDA events for the single copies, wall clock incl. the final
await_stream_ for the host steps (the region bench.py's e2e times).
usage: e2e_copy.py [nx ny nl] [steps]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")

nx, ny, nl = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1536, 1204, 70)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
m = mesh.make_mesh(nx, ny, nl)
f = mesh.fast_fields(m, seed=1, alloc=abi.pinned_empty)
plan = harness.DevicePlan(m)
df = harness.DeviceFields(plan, 1, packed=True)
Sn = m.S_n()
print(f"# N={m.myDim_nod2D} G={m.myDim_edge2D} nl={nl} S_n={Sn} active share of a dense node array: {Sn / (m.nnod * m.L):.3f}")
e0, e1 = abi.Event(), abi.Event()
for direct in (1, 0):
    abi.tune("DIRECT_COPY", direct)
    df.upload(f, outputs=False)
    for name, up in (("ttf", True), ("fct_adf_h", True), ("del_ttf_advvert", False)):
        host = getattr(f, name)
        fn = (lambda: df.upload_field(name, host)) if up else (lambda: df.download_field(name, host, merge=False))
        fn(); df.stream.sync()
        e0.record(df.stream)
        for _ in range(3): fn()
        e1.record(df.stream)
        ms = e1.ms_since(e0) / 3
        b = df.link_bytes(name, host, up)
        print(f"direct={direct} {'upload  ' if up else 'download'} {name:16s} {ms:8.2f} ms  link {b/1e9:6.3f} GB  {b/ms/1e6:6.1f} GB/s  (dense {host.nbytes/ms/1e6:6.1f} GB/s)")
    df.host_step(f, f, mode=1)
    t0 = time.perf_counter()
    df.host_steps(f, f, steps, mode=1)
    dt = (time.perf_counter() - t0) / steps
    up, dn = df.host_step_bytes(f)
    print(f"direct={direct} host step {dt*1e3:8.1f} ms  {Sn/dt/1e9:.3f} G updates/s  up {up/1e9:.2f} GB dn {dn/1e9:.2f} GB  -> {up/dt/1e9:.1f} GB/s up")
abi.tune("DIRECT_COPY", 0)
df.free(); plan.free()
