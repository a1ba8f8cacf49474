"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Bar (BASELINE.json north_star): fct_ttf_max / fct_ttf_min bit-exact; limited fluxes and
tracer increments within 1e-12 relative -- the kernels keep the oracle's operation order and are
built with -fmad=false, so these tests demand (and get) bit equality everywhere."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12          # relative, fluxes and increments (north_star)
OUT_KEYS = ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "fct_adf_h",
            "del_ttf_advvert", "del_ttf_advhoriz")


def check(got, want, keys=OUT_KEYS, exact=True, owned=None):
    for k in keys:
        a, b = getattr(got, k), getattr(want, k)
        if owned is not None and k != "fct_adf_h":
            a, b = a[:owned], b[:owned]
        if k in ("fct_ttf_max", "fct_ttf_min") or exact:
            assert bits_equal(a, b), f"{k}: max rel err {rel_err(a, b, 1e-30):.3e}"
        else:
            assert rel_err(a, b, floor=1e-30) <= TOL, k


def cases(mesh_mod, name):
    if name == "delaunay":    # unstructured: node degrees 2..11, ragged edge lists, land holes; CORE2-size column depths
        m = mesh_mod.make_delaunay_mesh(30000, 48, seed=1)
        return m, mesh_mod.make_fields(m, seed=9)
    if name == "deep":        # DART-depth columns: 40 level pairs, cut across warp items with ghost slots
        m = mesh_mod.make_mesh(48, 37, 80, seed=3)
        return m, mesh_mod.make_fields(m, seed=4)
    if name == "adversarial":
        return mesh_mod.adversarial_case(300, 17, seed=5)
    if name == "adversarial_even":
        return mesh_mod.adversarial_case(257, 16, seed=6)
    m = mesh_mod.make_workload(name)
    return m, mesh_mod.make_fields(m)


@pytest.mark.parametrize("name", ["tiny", "pi", "adversarial", "adversarial_even"])
@pytest.mark.parametrize("fused", [False, True])
def test_handle_abi_chain(mesh_mod, harness, abi, oracle_mod, name, fused):
    """The reference's own call sequence (transfer_var_async_ -> pre_comm_acc -> inter -> post ->
    c_acc) on dense host arrays."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    got = f.copy()
    abi.set_fused(fused)
    try:
        ch = harness.HandleChain(m, got)
        assert ch.step() == 10
        dev = ch.fetch("fct_ttf_max", "fct_ttf_min", "UV_rhs")
        ch.free()
    finally:
        abi.set_fused(False)
    check(got, want)
    if not fused:
        assert bits_equal(dev["UV_rhs"], want.UV_rhs)


@pytest.mark.parametrize("name", ["tiny", "pi", "core2", "deep", "adversarial", "adversarial_even"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_device_resident_step(mesh_mod, harness, oracle_mod, name, mode):
    """mode 0: ten stage kernels; 1: fastest fused variant of the plan (TMA-staged warp-item kernels on
    triangulations); 2: untiled fused kernels; 3: tile-staged fused kernels."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    plan = harness.DevicePlan(m)
    assert plan.pitch % 8 == 0 and plan.pitch >= m.nl
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    assert df.step(f, mode=mode) == 10
    got = df.download(f, mode=mode)
    check(got, want)
    if mode == 0:
        assert bits_equal(got.UV_rhs, want.UV_rhs)
    df.free()
    plan.free()


@pytest.mark.parametrize("name", ["pi", "deep"])
@pytest.mark.parametrize("knobs", [dict(WT_STAGES=2), dict(WT_STAGES=3), dict(WT_NODES=5), dict(WT_NODES=200, WT_SMEM=110 * 1024),
                                   dict(WT_WARPS_A=10, WT_WARPS_B=11, WT_ISSUERS=2),
                                   dict(WT_WARPS_A=16, WT_WARPS_B=13, WT_ISSUERS=6, WT_SMEM=20 * 1024)])
def test_warp_item_kernel_variants(mesh_mod, harness, abi, oracle_mod, name, knobs):
    """The warp-item kernels must not depend on their tiling: ring depth, tile size (tiny tiles:
    almost every neighbour row is a halo row; huge tiles: two stages only), consumer warps."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    defaults = dict(WT_STAGES=0, WT_NODES=0, WT_SMEM=0, WT_WARPS_A=0, WT_WARPS_B=0, WT_ISSUERS=0)
    try:
        for k, v in knobs.items():
            abi.tune(k, v)
        plan = harness.DevicePlan(m)
        df = harness.DeviceFields(plan, 1, with_uv=False)
        df.upload(f)
        df.stage("phaseA_warp", f)       # fails loudly if the plan has no warp tiles
        df.stage("phaseB_warp", f)
        check(df.download(f, mode=1), want)
        df.free()
        plan.free()
    finally:
        for k, v in defaults.items():
            abi.tune(k, v)


@pytest.mark.parametrize("name", ["pi", "deep"])
@pytest.mark.parametrize("knobs", [dict(WT_CONV=2), dict(WT_CONV=3), dict(WT_CONV=4), dict(WT_CONV=0), dict(WT_CONV=-1, WT_ISSUERS=2),
                                   dict(WT_REGS=80), dict(WT_REGS=88), dict(WT_REGS=72), dict(WT_REGS=72, WT_CONV=3),
                                   dict(WT_CONV=-1, WT_WARPS_A=24, WT_WARPS_B=24, WT_ISSUERS=3),
                                   dict(WT_OPT=14), dict(WT_OPT=22), dict(WT_OPT=38), dict(WT_OPT=70), dict(WT_FLAGS=3)])
def test_round2_kernel_alternatives_are_bit_identical(mesh_mod, harness, abi, oracle_mod, name, knobs):
    """Every measured alternative of DESIGN.md 3.4 (who runs the a1 pass, setmaxnreg register re-allocation
    between the warp roles, the 32-warp CTA, how the producers wait, the prefetch switches) on the packed
    storage, nl = 48 (three-stage ring) and nl = 80 (two stages, copy lists ahead): the same bits as the oracle."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    defaults = dict(WT_CONV=-1, WT_REGS=0, WT_OPT=-1, WT_FLAGS=0, WT_WARPS_A=0, WT_WARPS_B=0, WT_ISSUERS=0)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, packed=True)
    try:
        for k, v in knobs.items():
            abi.tune(k, v)
        df.upload(f)
        assert df.step(f, mode=1) == 10
        check(df.download(f, mode=1), want)
    finally:
        for k, v in defaults.items():
            abi.tune(k, v)
        df.free()
        plan.free()


@pytest.mark.parametrize("name", ["tiny", "pi", "core2", "deep", "delaunay"])
def test_packed_level_storage(mesh_mod, harness, oracle_mod, name):
    """The fast path's own layout: columns hold their active levels only, back to back."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(f)
    assert df.step(f, mode=1) == 10
    check(df.download(f, mode=1), want)
    with pytest.raises(Exception):
        df.stage("a1", f)                    # stage kernels need the padded layout: refused, not emulated
    assert df.step(f, mode=0) == 0
    df.free()
    plan.free()


def test_packed_multi_tracer(mesh_mod, harness, oracle_mod):
    m = mesh_mod.make_workload("pi")
    T = 3
    fs = [mesh_mod.make_fields(m, seed=1 + t) for t in range(T)]
    for t in range(1, T):
        for k in ("area", "area_inv", "hnode", "hnode_new"):
            setattr(fs[t], k, fs[0].__dict__[k])
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, T, packed=True)
    for t in range(T):
        df.upload(fs[t], tracer=t, static=(t == 0))
    assert df.step(fs[0], mode=1) == 10
    for t in range(T):
        want = fs[t].copy()
        oracle_mod.fct_ale(m, want)
        check(df.download(fs[t], tracer=t, mode=1), want)
    df.free()
    plan.free()


@pytest.mark.parametrize("packed", [False, True])
def test_host_resident_steps(mesh_mod, harness, oracle_mod, packed):
    """The end-to-end path bench.py times: per step upload the inputs, run, download the tendencies, the
    download of step k on a second stream overlapping the upload of step k+1."""
    m, f = cases(mesh_mod, "pi")
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False, packed=packed)
    df.upload(f)                               # mesh-static fields + defined output cells
    out = f.copy()
    assert df.host_steps(f, out, 3, mode=1) == 10
    act = np.arange(m.L)[None, :] < (m.nlevels_nod2D[:, None] - 1)
    for k in ("del_ttf_advvert", "del_ttf_advhoriz"):
        assert bits_equal(getattr(out, k)[act], getattr(want, k)[act]), k
    assert df.host_step(f, out, mode=1) == 10
    for k in ("del_ttf_advvert", "del_ttf_advhoriz"):
        assert bits_equal(getattr(out, k)[act], getattr(want, k)[act]), k
    df.free()
    plan.free()


@pytest.mark.parametrize("direct", [1, 0])
def test_packed_copies_from_page_locked_memory(mesh_mod, harness, abi, oracle_mod, direct):
    """Page-locked host arrays: with DIRECT_COPY=1 packed uploads / downloads run as one kernel over
    mapped host memory (only the levels that have a slot cross PCIe); the default is the staged copy.
    Both give the oracle's bits, and a download writes zeros into the levels that have no slot."""
    m, f0 = cases(mesh_mod, "deep")
    want = f0.copy()
    oracle_mod.fct_ale(m, want)
    f = f0.copy()
    for k, v in list(f.__dict__.items()):
        if isinstance(v, np.ndarray) and v.dtype == np.float64:
            p = abi.pinned_empty(v.shape)
            p[...] = v
            setattr(f, k, p)
    try:
        abi.tune("DIRECT_COPY", direct)
        plan = harness.DevicePlan(m)
        df = harness.DeviceFields(plan, 1, packed=True)
        df.upload(f)
        up, dn = df.host_step_bytes(f)
        dense_up = sum(getattr(f, k).nbytes for k in df.STEP_INPUTS)
        assert (up < 0.9 * dense_up) if direct else (up == dense_up)
        assert dn == sum(getattr(f, k).nbytes for k in df.STEP_RESULTS)
        out = f.copy()
        for k in df.STEP_RESULTS:
            p = abi.pinned_empty(getattr(f, k).shape)
            p[...] = np.nan
            setattr(out, k, p)
        assert df.host_steps(f, out, 2, mode=1) == 10
        act = np.arange(m.L)[None, :] < (m.nlevels_nod2D[:, None] - 1)
        for k in df.STEP_RESULTS:
            got = getattr(out, k)
            assert bits_equal(got[act], getattr(want, k)[act]), k
            assert np.isfinite(got).all(), k
        check(df.download(f, mode=1), want)
        df.free()
        plan.free()
    finally:
        abi.tune("DIRECT_COPY", 0)


def test_stage_by_stage(mesh_mod, harness, oracle_mod):
    """Each stage kernel against the oracle stage it replaces (the reference's NUM_KERNELS staged
    execution, src/fesom2-accelerate.cu:256-335)."""
    m, f = cases(mesh_mod, "pi")
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    want = f.copy()
    for name, fn in oracle_mod.STAGES:
        fn(m, want)
        df.stage(name, f)
        got = df.download(f, mode=0)
        check(got, want)
        assert bits_equal(got.UV_rhs, want.UV_rhs), name
    df.free()
    plan.free()


def test_atomic_alternatives(mesh_mod, harness, oracle_mod):
    """The reference's edge-centric fp64 atomicAdd scatter for b1h / c_h, kept as a measured
    alternative (stages 24 / 25): order of summation is the schedule's, so 1e-12, not bits."""
    m, f = cases(mesh_mod, "pi")
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    want = f.copy()
    swap = {"b1h": "b1h_atomic", "ch": "ch_atomic"}
    for name, fn in oracle_mod.STAGES:
        fn(m, want)
        df.stage(swap.get(name, name), f)
    got = df.download(f, mode=0)
    check(got, want, keys=("fct_ttf_max", "fct_ttf_min"))
    n = m.myDim_nod2D
    for k in OUT_KEYS[2:]:
        a, b = getattr(got, k), getattr(want, k)
        if k != "fct_adf_h":
            a, b = a[:n], b[:n]
        assert rel_err(a, b, floor=1.0) <= TOL, k
    df.free()
    plan.free()


def general_case(mesh_mod, name, vlimit, iter_yn, seed=11):
    m, f = cases(mesh_mod, name)
    f.vlimit, f.iter_yn = vlimit, iter_yn
    if iter_yn:
        rng = np.random.default_rng(seed)      # adf_*2 cells the listing never writes keep their values
        f.fct_adf_v2 = rng.standard_normal(f.fct_adf_v.shape)
        f.fct_adf_h2 = rng.standard_normal(f.fct_adf_h.shape)
    return m, f


@pytest.mark.parametrize("name", ["tiny", "pi", "deep", "adversarial"])
@pytest.mark.parametrize("vlimit,iter_yn", [(2, False), (3, False), (1, True), (3, True)])
def test_vlimit_and_iterative_branches(mesh_mod, harness, oracle_mod, name, vlimit, iter_yn):
    """vlimit 2 / 3 (docs/refactoring.md:113-148) and iter_yn
    (md:226-290) against the restatement of the listing (parity unpinned: the reference has no
    executable form of these branches)."""
    m, f = general_case(mesh_mod, name, vlimit, iter_yn)
    want = f.copy()
    oracle_mod.fct_ale_general(m, want)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    if iter_yn:
        df.upload_field("fct_adf_v2", f.fct_adf_v2)
        df.upload_field("fct_adf_h2", f.fct_adf_h2)
    assert df.step_general(f) == 10
    names = ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz",
             "fct_adf_v", "fct_adf_h", "fct_LO", "UV_rhs"] + (["fct_adf_v2", "fct_adf_h2"] if iter_yn else [])
    got = df.download(f, names=names)
    for k in names:
        assert bits_equal(getattr(got, k), getattr(want, k)), f"{k}: {rel_err(getattr(got, k), getattr(want, k), 1e-30):.3e}"
    df.free()
    plan.free()


@pytest.mark.parametrize("name", ["tiny", "pi", "core2", "deep"])
@pytest.mark.parametrize("vlimit", [2, 3])
def test_vlimit_on_the_fused_fast_path(mesh_mod, harness, oracle_mod, name, vlimit):
    """Packed fields: fct_ale_step_general_ runs the persistent warp-item kernels, phase A in its
    vlimit 2 / 3 variant (own a1 maxima of levels z-1..z+1 from the staged own row)."""
    m, f = general_case(mesh_mod, name, vlimit, False)
    want = f.copy()
    oracle_mod.fct_ale_general(m, want)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(f)
    n0 = harness.abi.launch_count()
    assert df.step_general(f) == 10
    assert harness.abi.launch_count() - n0 == 2          # two fused launches
    check(df.download(f, mode=1), want)
    df.free()
    plan.free()


@pytest.mark.parametrize("name", ["tiny", "pi", "deep"])
@pytest.mark.parametrize("vlimit", [1, 3])
def test_iterative_branch_on_the_fused_fast_path(mesh_mod, harness, oracle_mod, name, vlimit):
    """Packed fields, iter_yn: phase A + the iterative variant of the warp-item phase B (rejected flux
    parts to fct_adf_*2, low-order update with exact divisions in the listing's order), then
    fct_adf_* = fct_adf_*2; two iterative passes and the closing plain pass, against the oracle.
    Cells without a slot in the packed storage are not the device's: compared on the slots."""
    m, f = general_case(mesh_mod, name, vlimit, True)
    want = f.copy()
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(f)
    df.upload_field("fct_adf_v2", f.fct_adf_v2)
    df.upload_field("fct_adf_h2", f.fct_adf_h2)
    nslot, eslot = df._nslot, df._eslot[:, :m.L]

    def compare(names, got):
        for k in names:
            a, b = getattr(got, k), getattr(want, k)
            mask = eslot if k.startswith("fct_adf_h") else nslot[:, :a.shape[1]]
            assert bits_equal(a[mask], b[mask]), f"{k}: {rel_err(a[mask], b[mask], 1e-30):.3e}"

    it_names = ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "fct_adf_h", "fct_LO",
                "fct_adf_v2", "fct_adf_h2", "del_ttf_advvert", "del_ttf_advhoriz"]
    for it in (True, True):
        want.iter_yn = f.iter_yn = it
        oracle_mod.fct_ale_general(m, want)
        n0 = harness.abi.launch_count()
        assert df.step_general(f) == 10
        assert harness.abi.launch_count() - n0 == 2          # two fused launches (+ two device copies)
        compare(it_names, df.download(f, names=it_names))
    want.iter_yn = f.iter_yn = False
    oracle_mod.fct_ale_general(m, want)
    assert df.step_general(f) == 10
    got = df.download(f, mode=1)
    compare(["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz"], got)
    # a plain fused pass leaves its limited fluxes in FCT_ADF_V_OUT / FCT_ADF_H_OUT for the active levels
    # (the bottom flux is not limited, docs/refactoring.md:232, and stays in fct_adf_v)
    act = np.arange(m.nl)[None, :] < (m.nlevels_nod2D[:, None] - 1)
    assert bits_equal(got.fct_adf_v[act], want.fct_adf_v[act])
    eact = np.arange(m.L)[None, :] < m.edge_depth()[:, None]
    assert bits_equal(got.fct_adf_h[eact], want.fct_adf_h[eact])
    df.free()
    plan.free()


@pytest.mark.parametrize("knobs", [dict(WT_STAGES=2), dict(WT_STAGES=3), dict(WT_STAGES=4, WT_SMEM=50 * 1024), dict(WT_ISSUERS=4), dict(WT_OPT=7)])
def test_vlimit_fast_path_variants(mesh_mod, harness, abi, oracle_mod, knobs):
    """The vlimit variant of phase A in its other compiled shapes (ring depth, issuer warps) and
    with every scheduling option switched on."""
    m, f = general_case(mesh_mod, "deep", 3, False)
    want = f.copy()
    oracle_mod.fct_ale_general(m, want)
    defaults = dict(WT_STAGES=0, WT_SMEM=0, WT_ISSUERS=0, WT_OPT=-1)
    try:
        for k, v in knobs.items():
            abi.tune(k, v)
        plan = harness.DevicePlan(m)
        df = harness.DeviceFields(plan, 1, packed=True)
        df.upload(f)
        assert df.step_general(f) == 10
        check(df.download(f, mode=1), want)
        f1 = f.copy()
        f1.vlimit = 1
        want1 = f1.copy()
        oracle_mod.fct_ale(m, want1)
        df.upload(f1)
        assert df.step(f1, mode=1) == 10
        check(df.download(f1, mode=1), want1)
        df.free()
        plan.free()
    finally:
        for k, v in defaults.items():
            abi.tune(k, v)


def test_iterative_passes_converge_to_the_plain_limiter_inputs(mesh_mod, harness, oracle_mod):
    """Two passes of the iterative branch followed by the closing non-iterative call (the way
    FESOM2 drives fct_ale with iter_yn), device against oracle."""
    m, f = general_case(mesh_mod, "pi", 1, True)
    want = f.copy()
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    df.upload_field("fct_adf_v2", f.fct_adf_v2)
    df.upload_field("fct_adf_h2", f.fct_adf_h2)
    for it in (True, True, False):
        want.iter_yn = f.iter_yn = it
        oracle_mod.fct_ale_general(m, want)
        assert df.step_general(f) == 10
    names = ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz",
             "fct_adf_v", "fct_adf_h", "fct_LO"]
    got = df.download(f, names=names)
    for k in names:
        assert bits_equal(getattr(got, k), getattr(want, k)), k
    df.free()
    plan.free()


def test_handle_abi_vlimit(mesh_mod, harness, abi, oracle_mod):
    """fct_ale_pre_comm_acc_ honours its vlimit argument (2 and 3 run the stage kernels)."""
    for vlimit in (2, 3):
        m, f = general_case(mesh_mod, "pi", vlimit, False)
        want = f.copy()
        oracle_mod.fct_ale_general(m, want)
        got = f.copy()
        ch = harness.HandleChain(m, got)
        assert ch.step() == 10
        ch.fetch("fct_ttf_max", "fct_ttf_min")
        ch.free()
        check(got, want)


def test_reference_named_entry_points(abi, oracle_mod):
    """fct_ale_a{1,2,3,4}_reference_ / fct_ale_pre_comm_ keep the reference's names and host-array
    signatures but run on the GPU; checked against the golden vectors of src/reference.cpp."""
    import ctypes as C
    lib = abi.load()
    for case in ("ref_cpp_tiny", "ref_cpp_adversarial"):
        m, f, z = load_golden(case)
        g = f.copy()
        st = C.c_int(-1)
        lib.fct_ale_pre_comm_(
            C.byref(st), abi.dptr(g.fct_ttf_max.reshape(-1)), abi.dptr(g.fct_ttf_min.reshape(-1)),
            abi.dptr(g.fct_plus.reshape(-1)), abi.dptr(g.fct_minus.reshape(-1)), abi.dptr(g.ttf.reshape(-1)),
            abi.dptr(g.fct_LO.reshape(-1)), abi.dptr(g.fct_adf_v.reshape(-1)), abi.dptr(g.fct_adf_h.reshape(-1)),
            abi.dptr(g.UV_rhs.reshape(-1)), abi.dptr(g.area_inv.reshape(-1)), abi.ci(m.myDim_nod2D),
            abi.ci(m.eDim_nod2D), abi.ci(m.myDim_elem2D), abi.ci(m.myDim_edge2D), abi.ci(m.nl),
            abi.iptr(m.nlevels_nod2D), abi.iptr(m.nlevels_elem), abi.iptr(m.elem2D_nodes.reshape(-1)),
            abi.iptr(m.nod_in_elem2D_num), abi.iptr(m.nod_in_elem2D.reshape(-1)), abi.ci(m.nod_in_elem2D_dim),
            abi.iptr(m.edges.reshape(-1)), abi.iptr(m.edge_tri.reshape(-1)), abi.ci(1), abi.cd(g.flux_eps),
            abi.cd(g.bignumber), abi.cd(g.dt))
        assert st.value == 5          # alg_state of reference.cpp:302
        assert bits_equal(g.UV_rhs, z["a2_UV_rhs"])
        assert bits_equal(g.fct_ttf_max, z["a3_fct_ttf_max"])
        assert bits_equal(g.fct_ttf_min, z["a3_fct_ttf_min"])
        assert bits_equal(g.fct_plus, z["a4_fct_plus"])
        assert bits_equal(g.fct_minus, z["a4_fct_minus"])


def test_golden_b3_c_on_gpu(harness):
    """b3 / c kernels against the golden vectors of the reference's numpy reference() functions."""
    m, f, z = load_golden("ref_numpy_tiny")
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=True)
    df.upload(f)
    for s in ("b3v", "b3h", "cv", "ch"):
        df.stage(s, f)
    got = df.download(f, mode=0)
    assert bits_equal(got.fct_adf_v, z["b3v_fct_adf_v"])
    assert bits_equal(got.fct_adf_h, z["b3h_fct_adf_h"])
    assert rel_err(got.del_ttf_advvert, z["cv_del_ttf_advvert"], floor=1.0) < TOL
    assert rel_err(got.del_ttf_advhoriz, z["ch_del_ttf_advhoriz"], floor=1.0) < TOL
    df.free()
    plan.free()


@pytest.mark.parametrize("mode", [0, 1])
def test_multi_tracer_batch(mesh_mod, harness, oracle_mod, mode):
    """T, S + passive tracers in one launch (BASELINE.json config 5, reduced to 3 tracers on pi)."""
    m = mesh_mod.make_workload("pi")
    T = 3
    fs = [mesh_mod.make_fields(m, seed=1 + t) for t in range(T)]
    for t in range(1, T):       # mesh-static fields are shared by all tracers
        for k in ("area", "area_inv", "hnode", "hnode_new"):
            setattr(fs[t], k, fs[0].__dict__[k])
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, T, with_uv=True)
    for t in range(T):
        df.upload(fs[t], tracer=t, static=(t == 0))
    assert df.step(fs[0], mode=mode) == 10
    for t in range(T):
        want = fs[t].copy()
        oracle_mod.fct_ale(m, want)
        check(df.download(fs[t], tracer=t, mode=mode), want)
    df.free()
    plan.free()


def test_step_is_repeatable_and_deterministic(mesh_mod, harness):
    """No atomics anywhere: two runs from the same inputs give identical bits."""
    m, f = cases(mesh_mod, "core2")
    plan = harness.DevicePlan(m)
    outs = []
    for _ in range(2):
        df = harness.DeviceFields(plan, 1, with_uv=False)
        df.upload(f)
        assert df.step(f, mode=1) == 10
        outs.append(df.download(f, mode=1))
        df.free()
    check(outs[0], outs[1])
    plan.free()


@pytest.mark.parametrize("nparts,tiled,name", [(n, t, "pi") for n in (2, 5) for t in (False, True, "warp", "packed")] +
                         [(3, "packed", "deep"), (3, "warp", "deep"), (4, "packed", "delaunay"), (6, "packed", "delaunay-grown")])
def test_partitioned_on_one_gpu(mesh_mod, harness, oracle_mod, nparts, tiled, name):
    """Every partition of a mesh run on this GPU with the halo exchange emulated through the host:
    owned results of all partitions must reproduce the single-domain oracle bit for bit (boundary /
    interior node lists and their tiles, halo numbering, cut edges duplicated on both sides).  "deep":
    nl = 80 columns, where the packed layout runs the two-stage ring with the copy lists ahead."""
    m, f = cases(mesh_mod, name.split("-")[0])
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    # "-grown": an irregular partition (greedy graph growing: ragged boundaries, uneven halos, parts that
    # are not runs of the node numbering) of the unstructured mesh
    owner = mesh_mod.grow_partition(m, nparts, seed=3) if name.endswith("-grown") else None
    parts = mesh_mod.partition_mesh(m, nparts, owner=owner)
    plans = [harness.DevicePlan(p.mesh) for p in parts]
    lfs = [mesh_mod.slice_fields(f, p) for p in parts]
    dfs = [harness.DeviceFields(pl, 1, with_uv=False, packed=(tiled == "packed")) for pl in plans]
    stA = ["phaseA_tile_boundary", "phaseA_tile_interior"] if tiled else ["phaseA"]
    stB = ["phaseB_tile_interior", "phaseB_tile_boundary"] if tiled else ["phaseB"]
    if tiled in ("warp", "packed"):
        stA = ["phaseA_warp_boundary", "phaseA_warp_interior"]
        stB = ["phaseB_warp_interior", "phaseB_warp_boundary"]
    for df, lf in zip(dfs, lfs):
        df.upload(lf)
        for s in stA:
            df.stage(s, lf)
    # exchange_nod(fct_plus, fct_minus) through the host
    gplus = np.empty_like(f.fct_plus)
    gminus = np.empty_like(f.fct_minus)
    for p, df, lf in zip(parts, dfs, lfs):
        n = p.mesh.myDim_nod2D
        tmp = df.download(lf, mode=1, names=["fct_plus", "fct_minus"])
        gplus[p.mesh.node_gid[:n]] = tmp.fct_plus[:n]
        gminus[p.mesh.node_gid[:n]] = tmp.fct_minus[:n]
    for p, df, lf in zip(parts, dfs, lfs):
        df.upload_field("fct_plus", np.ascontiguousarray(gplus[p.mesh.node_gid]))
        df.upload_field("fct_minus", np.ascontiguousarray(gminus[p.mesh.node_gid]))
        df.stream.sync()
        for s in stB:
            df.stage(s, lf)
    got = f.copy()
    for p, df, lf in zip(parts, dfs, lfs):
        n = p.mesh.myDim_nod2D
        o = df.download(lf, mode=1)
        for k in OUT_KEYS:
            if k == "fct_adf_h":
                # an edge is written by exactly one of its owned end nodes on each rank that holds it
                got.fct_adf_h[p.mesh.edge_gid] = o.fct_adf_h
            else:
                getattr(got, k)[p.mesh.node_gid[:n]] = getattr(o, k)[:n]
    check(got, want)
    for df in dfs:
        df.free()
    for pl in plans:
        pl.free()


def test_full_size_properties(mesh_mod, harness):
    """DART-depth columns (nl = 80) at a size the oracle would take long on: size-independent
    properties of the limiter + determinism."""
    m = mesh_mod.make_mesh(1024, 780, 80, seed=0)
    f = mesh_mod.make_fields(m, seed=2, with_uv=False, poison=False)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, with_uv=False)
    df.upload(f)
    assert df.step(f, mode=1) == 10
    g = df.download(f, mode=1)
    L = m.L
    act = np.arange(L)[None, :] < (m.nlevels_nod2D[:, None] - 1)
    assert g.fct_plus[act].max() <= 1.0 and g.fct_minus[act].max() <= 1.0
    assert (g.fct_plus[act] >= 0).all() and (g.fct_minus[act] >= 0).all()
    assert (np.abs(g.fct_adf_v) <= np.abs(f.fct_adf_v)).all()
    assert (np.abs(g.fct_adf_h) <= np.abs(f.fct_adf_h)).all()
    assert (g.fct_adf_v * f.fct_adf_v >= 0).all() and (g.fct_adf_h * f.fct_adf_h >= 0).all()
    for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "del_ttf_advvert", "del_ttf_advhoriz"):
        assert bits_equal(getattr(g, k)[~act], getattr(f, k)[~act]), k
    # conservation: the horizontal increments of an edge cancel when weighted with area/dt
    w = (g.del_ttf_advhoriz - f.del_ttf_advhoriz) * f.area[:, :L]
    assert abs(w[act].sum()) <= 1e-9 * np.abs(w[act]).sum()
    df.free()
    plan.free()


def test_benchmark_scale_mesh_against_the_full_oracle(mesh_mod, harness, abi, oracle_mod):
    """A 1.8 M-node, 70-level mesh (a quarter of the NG5 workload: 50 000 tiles, the two-stage ring with
    the copy lists ahead that bench.py's default picks for nl >= 60, 10^8-element edge offsets) through
    the packed warp-item path, every output compared bit for bit with the full CPU oracle -- the
    configuration that is benchmarked is also a configuration that is checked."""
    m = mesh_mod.make_mesh(1560, 1204, 70, seed=0)
    assert m.myDim_nod2D > 1_800_000
    f = mesh_mod.fast_fields(m, seed=7, with_uv=True)
    # an out-of-depth read must show up in the results: poison the inactive levels of the inputs
    dead = np.arange(m.L)[None, :] >= (m.nlevels_nod2D[:, None] - 1)
    for k, v in (("ttf", 1e30), ("fct_LO", -1e30), ("hnode", 1e30), ("hnode_new", 1e30)):
        getattr(f, k)[dead] = v
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    f.UV_rhs = None
    want.UV_rhs = None
    plan = harness.DevicePlan(m)
    assert plan.packed_ok
    df = harness.DeviceFields(plan, 1, packed=True)
    df.upload(f)
    for rep in range(2):                     # the second pass re-arms the device-wide tile counters
        if rep:
            for k in ("del_ttf_advvert", "del_ttf_advhoriz"):
                df.upload_field(k, getattr(f, k))
        assert df.step(f, mode=1) == 10
        check(df.download(f, mode=1), want)
    df.free()
    plan.free()


@pytest.mark.parametrize("name", ["pi", "deep"])
def test_packed_host_arrays(mesh_mod, harness, oracle_mod, name):
    """A caller that keeps its columns in the packed level storage on the HOST: one contiguous copy per
    array (fct_ale_field_upload_packed_ / _download_packed_), no repack kernel; same results."""
    m, f = cases(mesh_mod, name)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    plan = harness.DevicePlan(m)
    df = harness.DeviceFields(plan, 1, packed=True)
    names = ["ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "del_ttf_advvert", "del_ttf_advhoriz", "area", "area_inv", "hnode",
             "hnode_new", "fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus"]
    for k in names:
        kind = "edge" if k == "fct_adf_h" else "node"
        ph = plan.pack_host(getattr(f, k), kind)
        assert ph.size == plan.packed_columns(kind)[-1]
        df.upload_packed(k, ph)
    df.upload_packed("fct_adf_h_out", plan.pack_host(f.fct_adf_h, "edge"))
    df.upload_packed("fct_adf_v_out", plan.pack_host(f.fct_adf_v, "node"))
    assert df.step(f, mode=1) == 10
    got = f.copy()
    for k, src in (("fct_ttf_max", None), ("fct_ttf_min", None), ("fct_plus", None), ("fct_minus", None),
                   ("del_ttf_advvert", None), ("del_ttf_advhoriz", None), ("fct_adf_v", "fct_adf_v_out"), ("fct_adf_h", "fct_adf_h_out")):
        kind = "edge" if k == "fct_adf_h" else "node"
        ph = np.empty(int(plan.packed_columns(kind)[-1]))
        df.download_packed(src or k, ph)
        df.stream.sync()
        plan.unpack_host(ph, getattr(got, k), kind)
    check(got, want)
    # a dense upload and a packed upload of the same array leave the same device image
    a, b = np.empty(int(plan.packed_columns("node")[-1])), np.empty(int(plan.packed_columns("node")[-1]))
    df.upload_field("ttf", f.ttf)
    df.download_packed("ttf", a)
    df.upload_packed("ttf", plan.pack_host(f.ttf, "node"))
    df.download_packed("ttf", b)
    df.stream.sync()
    assert np.array_equal(a, b)
    # padded fields refuse packed host copies
    dd = harness.DeviceFields(plan, 1, with_uv=False)
    with pytest.raises(Exception):
        dd.upload_packed("ttf", a)
    dd.free()
    df.free()
    plan.free()


def test_host_resident_time_step_batch(mesh_mod, harness, oracle_mod):
    """host_steps_batch: the mesh-static inputs of a time step uploaded once, the per-tracer inputs per
    tracer (here the same tracer every time): the tendencies accumulate exactly like repeated steps."""
    m, f = cases(mesh_mod, "pi")
    want = f.copy()
    for _ in range(3):
        w2 = want.copy()
        oracle_mod.fct_ale(m, w2)
        want.del_ttf_advvert, want.del_ttf_advhoriz = w2.del_ttf_advvert, w2.del_ttf_advhoriz
    plan = harness.DevicePlan(m)
    for packed_host in (False, True):
        df = harness.DeviceFields(plan, 1, packed=True)
        df.upload(f)
        g = f.copy()
        ph = None
        if packed_host:
            ph = {k: plan.pack_host(getattr(g, k), "edge" if k == "fct_adf_h" else "node") for k in df.STEP_INPUTS}
            for k in df.STEP_RESULTS:
                ph["out_" + k] = ph[k]          # results land where the next tracer's inputs are read from
        for _ in range(3):
            # each "tracer" continues from the tendencies the previous one produced (g is both input and output)
            assert df.host_steps_batch(g, g, 1, packed_host=ph) == 10
        if packed_host:
            for k in df.STEP_RESULTS:
                plan.unpack_host(ph[k], getattr(g, k), "node")
        act = np.arange(m.L)[None, :] < (m.nlevels_nod2D[:, None] - 1)
        for k in df.STEP_RESULTS:
            assert np.array_equal(getattr(g, k)[act], getattr(want, k)[act]), (packed_host, k)
        df.free()
    plan.free()
