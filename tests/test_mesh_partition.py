"""CPU: synthetic mesh invariants, partitions, and the N>1 path of the host logic:
partitioned oracle runs (in-process and over a world_size-2 gloo group) must reproduce the
single-domain oracle bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, bits_equal

OUT_KEYS = ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "fct_adf_h",
            "del_ttf_advvert", "del_ttf_advhoriz")


def test_mesh_invariants(mesh_mod):
    m = mesh_mod.make_workload("pi")
    N, E, G = m.myDim_nod2D, m.myDim_elem2D, m.myDim_edge2D
    assert 2900 < N < 3200 and 5600 < E < 6100          # BASELINE.json config 0 (3,140 nodes, ~5.8k elements)
    tri = m.elem2D_nodes - 1
    assert tri.min() == 0 and tri.max() == N - 1
    # Euler characteristic of a planar triangulation with holes: V - E + F = 1 - holes <= 1
    assert N - G + E <= 1
    # node depth = max over its ring (the invariant the model guarantees)
    want = np.zeros(N, dtype=np.int32)
    np.maximum.at(want, tri.ravel(), np.repeat(m.nlevels_elem, 3))
    assert np.array_equal(want, m.nlevels_nod2D)
    # every edge: left element present, both elements contain both nodes
    e = m.edges - 1
    et = m.edge_tri - 1
    assert (et[:, 0] >= 0).all()
    for side in (0, 1):
        ok = et[:, side] >= 0
        t = tri[et[ok, side]]
        assert ((t == e[ok, 0:1]).any(1) & (t == e[ok, 1:2]).any(1)).all()
    assert (m.edge_tri[:, 1] == 0).sum() > 0             # the land mask produced boundary edges
    # ring table consistent with the element table
    for n in (0, N // 2, N - 1):
        ring = m.nod_in_elem2D[n, : m.nod_in_elem2D_num[n]] - 1
        assert sorted(ring) == sorted(np.flatnonzero((tri == n).any(1)))
    assert abs(m.bytes_alg() / m.S_n() - 240) < 8        # ~240 B per update


def test_hilbert_locality(mesh_mod):
    m = mesh_mod.make_workload("core2")
    e = m.edges.astype(np.int64) - 1
    assert np.median(np.abs(e[:, 0] - e[:, 1])) < 64     # neighbours stay close in memory


def _run_partitioned(mesh_mod, oracle, m, f, nparts, exchange):
    parts = mesh_mod.partition_mesh(m, nparts)
    lfs = [mesh_mod.slice_fields(f, p) for p in parts]
    for p, lf in zip(parts, lfs):
        oracle.pre_comm(p.mesh, lf)
    exchange(parts, lfs)
    got = f.copy()
    for p, lf in zip(parts, lfs):
        oracle.post_comm(p.mesh, lf)
        n = p.mesh.myDim_nod2D
        for k in OUT_KEYS:
            if k == "fct_adf_h":
                got.fct_adf_h[p.mesh.edge_gid] = lf.fct_adf_h
            else:
                getattr(got, k)[p.mesh.node_gid[:n]] = getattr(lf, k)[:n]
    return got, parts


def _host_exchange(parts, lfs):
    """exchange_nod(fct_plus, fct_minus): send lists -> the receivers' halo ranges."""
    for p, lf in zip(parts, lfs):
        for peer, nodes in p.send_lists.items():
            first, cnt = parts[peer].recv_ranges[p.rank]
            assert cnt == nodes.size
            lfs[peer].fct_plus[first:first + cnt] = lf.fct_plus[nodes]
            lfs[peer].fct_minus[first:first + cnt] = lf.fct_minus[nodes]


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_partitioned_oracle_equals_single_domain(mesh_mod, oracle_mod, nparts):
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    got, parts = _run_partitioned(mesh_mod, oracle_mod, m, f, nparts, _host_exchange)
    for k in OUT_KEYS:
        assert bits_equal(getattr(got, k), getattr(want, k)), k
    # partition bookkeeping
    assert sum(p.mesh.myDim_nod2D for p in parts) == m.myDim_nod2D
    for p in parts:
        n = p.mesh.myDim_nod2D
        assert p.boundary_nodes.size + p.interior_nodes.size == n
        assert set(p.send_lists) == set(p.recv_ranges)
        sent = np.unique(np.concatenate(list(p.send_lists.values()))) if p.send_lists else np.empty(0)
        assert np.array_equal(sent, p.boundary_nodes)       # send set == nodes with a halo neighbour
        w = (p.mesh.nlevels_nod2D[:n].astype(np.int64) - 1).sum()
        assert abs(w - m.S_n() / nparts) < 0.15 * m.S_n() / nparts   # balanced on node-levels


def test_lazy_partition_matches_full(mesh_mod):
    m = mesh_mod.make_workload("pi")
    full = mesh_mod.partition_mesh(m, 5)
    one = mesh_mod.partition_mesh(m, 5, ranks=[3])[0]
    ref = full[3]
    assert np.array_equal(one.mesh.node_gid, ref.mesh.node_gid)
    assert np.array_equal(one.mesh.edges, ref.mesh.edges)
    assert one.recv_ranges == ref.recv_ranges
    assert all(np.array_equal(one.send_lists[k], ref.send_lists[k]) for k in ref.send_lists)


def test_world_size_2_gloo(tmp_path):
    """Two processes, one partition each, halo exchange over torch.distributed (gloo, CPU)."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


@pytest.mark.parametrize("vlimit", [1, 3])
def test_partitioned_iterative_branch_equals_single_domain(mesh_mod, oracle_mod, vlimit):
    """N>1 logic of fct_ale_step_general_ on the CPU: per-partition oracle passes with the halo rows of
    fct_plus / fct_minus exchanged inside a pass and those of fct_LO after an iterative pass (the
    low-order update touches owned rows only; docs/refactoring.md:265-290) reproduce the single
    domain bit for bit over an iterative pass followed by the closing plain pass."""
    m = mesh_mod.make_workload("pi")
    f = mesh_mod.make_fields(m)
    f.vlimit = vlimit
    f.fct_adf_v2, f.fct_adf_h2 = np.zeros_like(f.fct_adf_v), np.zeros_like(f.fct_adf_h)
    want = f.copy()
    parts = mesh_mod.partition_mesh(m, 4)
    lfs = [mesh_mod.slice_fields(f, p) for p in parts]

    def exchange(names):
        for p, lf in zip(parts, lfs):
            for peer, nodes in p.send_lists.items():
                first, cnt = parts[peer].recv_ranges[p.rank]
                for k in names:
                    getattr(lfs[peer], k)[first:first + cnt] = getattr(lf, k)[nodes]

    for it in (True, False):
        want.iter_yn = it
        oracle_mod.fct_ale_general(m, want)
        for p, lf in zip(parts, lfs):
            lf.iter_yn = it
        # a pass = pre_comm part, exchange of the factors, post part; run the oracle's general
        # subroutine per partition with the exchange hooked in between
        pend = []
        for p, lf in zip(parts, lfs):
            oracle_mod.a1(p.mesh, lf)
            oracle_mod.a2(p.mesh, lf)
            oracle_mod.a3_vlimit(p.mesh, lf)
            oracle_mod.b1_vertical(p.mesh, lf)
            oracle_mod.b1_horizontal(p.mesh, lf)
            oracle_mod.b2(p.mesh, lf)
        exchange(["fct_plus", "fct_minus"])
        for p, lf in zip(parts, lfs):
            if it:
                oracle_mod.b3_vertical_iter(p.mesh, lf)
                oracle_mod.b3_horizontal_iter(p.mesh, lf)
                oracle_mod.lo_update(p.mesh, lf)
                lf.fct_adf_h[...] = lf.fct_adf_h2
                lf.fct_adf_v[...] = lf.fct_adf_v2
            else:
                oracle_mod.post_comm(p.mesh, lf)
        if it:
            exchange(["fct_LO"])
    for p, lf in zip(parts, lfs):
        n = p.mesh.myDim_nod2D
        g = p.mesh.node_gid[:n]
        for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "fct_LO", "del_ttf_advvert",
                  "del_ttf_advhoriz"):
            assert bits_equal(getattr(lf, k)[:n], getattr(want, k)[g]), (p.rank, k)
        assert bits_equal(lf.fct_adf_h, want.fct_adf_h[p.mesh.edge_gid]), p.rank


def test_graph_grown_partition_of_an_unstructured_mesh_equals_single_domain(mesh_mod, oracle_mod):
    """Irregular partition (greedy graph growing: parts are not runs of the node numbering, ragged
    boundaries, uneven halos) of an unstructured Delaunay mesh: oracle pre_comm per part, host exchange
    of fct_plus / fct_minus, oracle post_comm -- owned results bit-identical to the single domain, and
    the plan's halo contract holds (send nodes = owned nodes with a halo neighbour, receive ranges
    inside the halo rows, no edge between two halo nodes)."""
    m = mesh_mod.make_delaunay_mesh(6000, 30, seed=1)
    f = mesh_mod.make_fields(m, seed=5)
    want = f.copy()
    oracle_mod.fct_ale(m, want)
    owner = mesh_mod.grow_partition(m, 5, seed=2)
    assert not np.all(np.diff(owner) >= 0)                 # not contiguous runs of the numbering
    parts = mesh_mod.partition_mesh(m, 5, owner=owner)
    assert sum(p.mesh.myDim_nod2D for p in parts) == m.myDim_nod2D
    lfs = [mesh_mod.slice_fields(f, p) for p in parts]
    for p, lf in zip(parts, lfs):
        oracle_mod.pre_comm(p.mesh, lf)
    for p, lf in zip(parts, lfs):
        n, H = p.mesh.myDim_nod2D, p.mesh.eDim_nod2D
        le = p.mesh.edges.astype(np.int64) - 1
        assert ((le < n).any(1)).all()                     # every local edge touches an owned node
        bnd = np.zeros(n, bool)
        cut = (le >= n).any(1)
        ends = le[cut].ravel()
        bnd[ends[ends < n]] = True
        for peer, nodes in p.send_lists.items():
            assert bnd[nodes].all()
        for peer, (first, cnt) in p.recv_ranges.items():
            assert n <= first and first + cnt <= n + H
            nodes = parts[peer].send_lists[p.rank]
            assert nodes.size == cnt
            for k in ("fct_plus", "fct_minus"):
                getattr(lf, k)[first:first + cnt] = getattr(lfs[peer], k)[nodes]
    for p, lf in zip(parts, lfs):
        oracle_mod.post_comm(p.mesh, lf)
        n = p.mesh.myDim_nod2D
        g = p.mesh.node_gid[:n]
        for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus", "fct_adf_v", "del_ttf_advvert", "del_ttf_advhoriz"):
            assert np.array_equal(getattr(lf, k)[:n], getattr(want, k)[g]), k
        assert np.array_equal(lf.fct_adf_h, want.fct_adf_h[p.mesh.edge_gid])


def test_fast_fields_and_digest_are_functions_of_global_ids(mesh_mod):
    """bench.py's determinism claim rests on two properties checked here on the CPU: the synthetic
    fields of a partition are exactly the rows of the single-domain fields (every cell a function of
    its global row id, level and array), and the per-partition digests of a node array add up (mod
    2^64) to the single-domain digest -- for contiguous and for graph-grown partitions -- while a
    one-ulp change of one owned cell changes it."""
    m = mesh_mod.make_mesh(90, 70, 25, seed=2)
    f = mesh_mod.fast_fields(m, seed=1)
    want = mesh_mod.digest_node_array(m, f.ttf, 3)
    for owner in (None, mesh_mod.grow_partition(m, 4, seed=1)):
        tot = 0
        for p in mesh_mod.partition_mesh(m, 4, owner=owner):
            lf = mesh_mod.fast_fields(p.mesh, seed=1)
            sf = mesh_mod.slice_fields(f, p)
            for k in ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "area", "area_inv", "hnode", "hnode_new",
                      "del_ttf_advvert", "del_ttf_advhoriz"):
                assert np.array_equal(getattr(lf, k), getattr(sf, k)), k
            tot = (tot + mesh_mod.digest_node_array(p.mesh, lf.ttf, 3)) & (2 ** 64 - 1)
        assert tot == want
    g = f.ttf.copy()
    g[7, 2] = np.nextafter(g[7, 2], np.inf)
    assert mesh_mod.digest_node_array(m, g, 3) != want
    assert mesh_mod.digest_node_array(m, f.ttf, 4) != want           # the salt separates the arrays
    # inactive levels are not part of the digest
    dead = np.arange(m.L)[None, :] >= (m.nlevels_nod2D[:, None] - 1)
    h = f.ttf.copy()
    h[dead] = -123.0
    assert mesh_mod.digest_node_array(m, h, 3) == want


def test_reference_arm_of_bench_runs_the_same_mesh_and_its_digest_is_partition_independent():
    """`bench.py --impl reference` on the pi mesh: the CPU arm splits the workload mesh over the host
    cores (forked processes, host exchange_nod over gloo); whatever the number of partitions, the
    digests of its outputs are those of the single process."""
    import json
    outs = []
    for cores in ("0", "0-2"):
        r = subprocess.run(["taskset", "-c", cores, sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                            "--workload", "pi", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["details"]["same_mesh_as_product_arm"] is True
        assert line["config"]["workload"] == "pi" and line["e2e"]["value"] == line["value"]
        outs.append(line)
    assert outs[0]["cpu_baseline"]["cores"] == 1 and outs[1]["cpu_baseline"]["cores"] == 3
    assert outs[0]["digest"] == outs[1]["digest"] and len(outs[0]["digest"]) == 4
    assert outs[0]["config"] == outs[1]["config"]
