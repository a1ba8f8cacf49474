"""stress2rhs on the device: CUDA-event time of stress2rhs_acc_ on a
mesh-sized connectivity, algorithmic bytes (every element scalar and node scalar once: 6 + 6 doubles
and 3 int32 per element, 5 doubles per node) against the copy peak, next to the oracle on one core.
usage: stress_bench.py [nx ny]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mesh = importlib.import_module("fesom2-accelerate_b200.mesh")
harness = importlib.import_module("fesom2-accelerate_b200.harness")
abi = importlib.import_module("fesom2-accelerate_b200.abi")
import oracle   # checker / CPU baseline only

nx, ny = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (3072, 2408)
m = mesh.make_mesh(nx, ny, 3)
tri = np.ascontiguousarray((m.elem2D_nodes - 1).T)
d = oracle.stress_case(m.myDim_nod2D, m.myDim_elem2D, seed=1, elem_nodes=tri)
N, E = d["N"], d["E"]
alg = E * (12 * 8 + 12) + N * 5 * 8
ch = harness.StressChain(d)
e0, e1 = abi.Event(), abi.Event()
for _ in range(3): ch.run(sync=False)
ch.stream.sync(); e0.record(ch.stream)
reps = 20
for _ in range(reps): ch.run(sync=False)
e1.record(ch.stream)
ms = e1.ms_since(e0) / reps
u, v = ch.fetch()
t0 = time.perf_counter(); want = oracle.stress2rhs(d); cpu = time.perf_counter() - t0
ok = np.array_equal(u, want[0]) and np.array_equal(v, want[1])
print(f"stress2rhs N={N} E={E}: {ms*1e3:.1f} us  {alg/ms/1e6:.0f} GB/s algorithmic = {alg/ms/1e6/65.472:.1f}% of 6547 GB/s; "
      f"{N/ms/1e6:.2f} G nodes/s; oracle on one core {cpu*1e3:.1f} ms ({cpu*1e3/ms:.0f}x); bit-identical: {ok}")
ch.free()
