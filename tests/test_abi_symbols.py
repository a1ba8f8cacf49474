"""CPU: the C-ABI library loads and exports every symbol include/fesom2-accelerate.h declares; the
product never links the oracle; without a GPU the compute entry points fail loudly."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fesom2-accelerate.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bvoid\s+(\w+)\s*\(", src)))


def test_header_and_binding_agree(abi):
    hs = header_symbols()
    assert sorted(abi.SYMBOLS) == hs, set(abi.SYMBOLS) ^ set(hs)


def test_library_exports_every_symbol(abi):
    lib = abi.load()
    for s in abi.SYMBOLS:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(abi.SYMBOLS) <= exported
    # the reference's 18 C symbols + its CPU-oracle names
    for s in ("set_mpi_rank_", "transfer_mesh_", "alloc_var_", "reserve_var_", "allocate_pinned_doubles_",
              "transfer_var_", "transfer_var_async_", "make_stream_", "await_stream_",
              "fct_ale_pre_comm_acc_", "fct_ale_inter_comm_acc_", "fct_ale_post_comm_acc_",
              "fct_ale_a1_accelerated", "fct_ale_a2_accelerated", "fct_ale_a1_a2_accelerated",
              "fct_ale_a1_reference_", "fct_ale_a2_reference_", "fct_ale_a3_reference_", "fct_ale_a4_reference_",
              "fct_ale_pre_comm_"):
        assert s in exported, s


def test_product_does_not_depend_on_the_oracle(abi):
    out = subprocess.run(["ldd", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out
    pkg_dir = os.path.dirname(abi.LIB_PATH.replace("/lib/", "/"))
    for root, _, files in os.walk(os.path.join(ROOT, "fesom2-accelerate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libfct_oracle" not in txt and "libref" not in txt, f


def test_gpumemory_struct_prefix_matches_reference(abi):
    """Leading members of struct gpuMemory = the reference's (include/fesom2-accelerate.h:19-28)."""
    g = abi.GpuMemory
    assert [f[0] for f in g._fields_[:4]] == ["host_pointer", "device_pointer", "size", "event"]
    assert g.host_pointer.offset == 0 and g.device_pointer.offset == 8 and g.size.offset == 16 and g.event.offset == 24


def test_no_cpu_fallback_without_a_device(abi, harness, mesh_mod):
    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.AbiError):
        abi.device_info()
    m = mesh_mod.make_workload("tiny")
    with pytest.raises(abi.AbiError):
        harness.DevicePlan(m)
    with pytest.raises(abi.AbiError):
        harness.HandleChain(m, mesh_mod.make_fields(m))


def test_fortran_module_matches_header(abi):
    """fortran/fesom2_accelerate_b200.f90 (no Fortran compiler in this image): every bind(C) name is a
    symbol of the header, with as many dummy arguments as the C prototype has parameters, and every
    Part 2 symbol of the header has an interface."""
    f90 = open(os.path.join(ROOT, "fortran", "fesom2_accelerate_b200.f90")).read()
    f90 = re.sub(r"!.*", "", f90)
    f90 = re.sub(r"&\s*\n\s*&?", " ", f90)
    bound = {}
    for m in re.finditer(r"subroutine\s+\w+\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", f90, flags=re.I):
        bound[m.group(2)] = len([a for a in m.group(1).split(",") if a.strip()])
    hdr = open(os.path.join(ROOT, "include", "fesom2-accelerate.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
              for m in re.finditer(r"\bvoid\s+(\w+)\s*\(([^)]*)\)\s*;", hdr)}
    assert len(bound) >= 30
    for name, nargs in bound.items():
        assert name in protos, name
        assert nargs == protos[name], (name, nargs, protos[name])
    part2 = hdr[hdr.index("Part 2") if "Part 2" in hdr else 0:]
    part2_syms = set(re.findall(r"\bvoid\s+(\w+_)\s*\(", hdr[hdr.index("fct_ale_c_acc_") - 200:]))
    missing = part2_syms - set(bound) - {"fct_ale_plan_inspect_"}      # host-side introspection for the tests
    assert not missing, missing
    enum = dict(re.findall(r"(FCT_\w+)\s*=\s*(\d+)", hdr))
    for k, v in re.findall(r"(FCT_\w+)\s*=\s*(\d+)", f90):
        assert enum[k] == v, k


def _c_param_kind(decl):
    """C parameter declaration -> the ISO_C_BINDING type a by-reference Fortran dummy must have."""
    d = decl.replace("const", " ").strip()
    if "**" in d:
        return ("type", "c_ptr")
    base = d.split("*")[0].split()
    if "*" not in d:
        raise AssertionError(f"by-value parameter in a Fortran-callable prototype: {decl!r}")
    if base[:2] == ["long", "long"]:
        return ("integer", "c_long_long")
    if base[0] in ("int", "unsigned"):
        return ("integer", "c_int")
    if base[0] in ("real_type", "double"):
        return ("real", "c_double")
    if base[0] == "bool":
        return ("logical", "c_bool")
    if base[0] == "char":
        return ("character", "c_char")
    if base[0] == "void":
        return ("type", "c_ptr")        # void* seen from Fortran: an opaque address passed by reference is not used here
    raise AssertionError(f"unknown C type in {decl!r}")


def test_fortran_module_is_parsed_and_typed_like_the_header(abi):
    """No Fortran compiler exists in this image, but numpy.f2py's Fortran parser does: the module is
    PARSED (syntax: module / interface / subroutine / declarations all resolve), and for every
    bind(C) interface the type and kind of each dummy argument is compared with the parameter of
    the C prototype it is bound to (int* <-> integer(c_int), real_type* <-> real(c_double),
    void** <-> type(c_ptr), long long* <-> integer(c_long_long), bool* <-> logical(c_bool),
    char* <-> character(c_char)), in order."""
    import contextlib
    import io
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    with contextlib.redirect_stdout(io.StringIO()):
        blocks = cf.crackfortran([os.path.join(ROOT, "fortran", "fesom2_accelerate_b200.f90")])
    assert len(blocks) == 1 and blocks[0]["block"] == "module" and blocks[0]["name"] == "fesom2_accelerate_b200"
    hdr = open(os.path.join(ROOT, "include", "fesom2-accelerate.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {m.group(1): [a.strip() for a in m.group(2).replace("\n", " ").split(",") if a.strip()]
              for m in re.finditer(r"\bvoid\s+(\w+)\s*\(([^)]*)\)\s*;", hdr)}
    checked = 0
    subs = []
    for b in blocks[0]["body"]:
        if b["block"] == "interface":
            subs += b["body"]
    assert len(subs) >= 40
    for sub in subs:
        bind = sub.get("bindlang", {}).get(sub["name"])
        assert bind and bind["lang"] == "c", sub["name"]
        cname = bind["name"]
        assert cname in protos, cname
        cargs = protos[cname]
        assert len(cargs) == len(sub["args"]), (cname, len(cargs), len(sub["args"]))
        for decl, arg in zip(cargs, sub["args"]):
            v = sub["vars"][arg]
            want = _c_param_kind(decl)
            got_type = v.get("typespec")
            got_kind = (v.get("kindselector") or {}).get("kind") or v.get("typename")
            if want[0] == "character":
                assert got_type == "character", (cname, arg, decl, v)
            else:
                assert (got_type, got_kind) == want, (cname, arg, decl, got_type, got_kind)
            checked += 1
    assert checked > 200
    # the module procedure that replaces the body of fct_ale parses too
    assert any(b["block"] == "subroutine" and b["name"] == "fct_ale_device" for b in blocks[0]["body"])
