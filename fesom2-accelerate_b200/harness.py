"""Host-side mirror of the reference's calling sequences, over the C ABI (ctypes).

* `HandleChain` replays exactly what FESOM2's Fortran does per tracer step with the reference
  library (/root/reference/src/fesom2-accelerate.cu:258-379):
  transfer_mesh_ x7, alloc_var_, transfer_var_async_, fct_ale_pre_comm_acc_, await_stream_,
  fct_ale_inter_comm_acc_, fct_ale_post_comm_acc_, plus the new fct_ale_c_acc_.
* `DevicePlan` / `DeviceFields` / `HaloLink` drive the device-resident plan / fields / step ABI.

This module plays the role of the reference's kernels/*.py drivers (argument set-up, run, fetch
results) without kernel_tuner; checking against an answer is left to the tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import abi
from .abi import cb, cd, ci, dptr, iptr
from .mesh import Fields, Mesh, Partition


class HandleChain:
    NODE_L = ("ttf", "fct_LO", "hnode", "hnode_new", "del_ttf_advvert", "del_ttf_advhoriz",
              "fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus")

    def __init__(self, mesh: Mesh, fields: Fields, with_c: bool = True):
        self.lib = abi.load()
        abi.device_info()
        self.m = mesh
        self.f = fields
        self.stream = abi.Stream()
        m = mesh
        self.mesh_vars = {k: abi.Var(np.ascontiguousarray(getattr(m, k)).reshape(-1), mesh=True)
                          for k in ("nlevels_nod2D", "nlevels_elem", "elem2D_nodes",
                                    "nod_in_elem2D_num", "nod_in_elem2D", "edges", "edge_tri")}
        ev = ("ttf", "fct_adf_v", "fct_adf_h")     # pre_comm_acc waits on these upload events
        names = ["ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "area_inv", "fct_ttf_max", "fct_ttf_min",
                 "fct_plus", "fct_minus"]
        if with_c:
            names += ["area", "hnode", "hnode_new", "del_ttf_advvert", "del_ttf_advhoriz"]
        self.vars: Dict[str, abi.Var] = {}
        for k in names:
            self.vars[k] = abi.Var(getattr(fields, k).reshape(-1), event=k in ev)
        # UV_rhs is a device-only scratch in the reference call sequence (reserve_var_)
        self.vars["UV_rhs"] = abi.Var(size=mesh.myDim_elem2D * mesh.L * 2)
        self.with_c = with_c
        self.alg_state = C.c_int(0)
        # outputs that only live on the device need one upload so untouched cells are defined
        for k in ("fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus"):
            self.vars[k].upload()

    def pre_comm(self):
        m, f, v, mv, s = self.m, self.f, self.vars, self.mesh_vars, self.stream
        for k in ("ttf", "fct_adf_v", "fct_adf_h"):
            v[k].upload(stream=s, record=True)
        v["area_inv"].upload(stream=s)
        self.lib.fct_ale_pre_comm_acc_(
            C.byref(self.alg_state), s.ref, v["fct_ttf_max"].ref, v["fct_ttf_min"].ref,
            v["fct_plus"].ref, v["fct_minus"].ref, v["ttf"].ref, v["fct_LO"].ref, v["fct_adf_v"].ref,
            v["fct_adf_h"].ref, v["UV_rhs"].ref, v["area_inv"].ref, ci(m.myDim_nod2D),
            ci(m.eDim_nod2D), ci(m.myDim_elem2D), ci(m.myDim_edge2D), ci(m.nl),
            mv["nlevels_nod2D"].ref, mv["nlevels_elem"].ref, mv["elem2D_nodes"].ref,
            mv["nod_in_elem2D_num"].ref, mv["nod_in_elem2D"].ref, ci(m.nod_in_elem2D_dim),
            mv["edges"].ref, mv["edge_tri"].ref, ci(f.vlimit), cd(f.flux_eps), cd(f.bignumber),
            cd(f.dt))
        return self.alg_state.value

    def inter_comm(self):
        m, v, mv, s = self.m, self.vars, self.mesh_vars, self.stream
        self.lib.fct_ale_inter_comm_acc_(C.byref(self.alg_state), s.ref, v["fct_plus"].ref,
                                         v["fct_minus"].ref, v["fct_adf_v"].ref, ci(m.myDim_nod2D),
                                         ci(m.nl), mv["nlevels_nod2D"].ref)
        return self.alg_state.value

    def post_comm(self):
        m, v, mv, s = self.m, self.vars, self.mesh_vars, self.stream
        self.lib.fct_ale_post_comm_acc_(C.byref(self.alg_state), s.ref, v["fct_plus"].ref,
                                        v["fct_minus"].ref, v["fct_adf_h"].ref, ci(m.myDim_edge2D),
                                        ci(m.nl), mv["nlevels_elem"].ref, ci(m.nod_in_elem2D_dim),
                                        mv["edges"].ref, mv["edge_tri"].ref)
        return self.alg_state.value

    def stage_c(self):
        m, f, v, mv, s = self.m, self.f, self.vars, self.mesh_vars, self.stream
        self.lib.fct_ale_c_acc_(C.byref(self.alg_state), s.ref, v["del_ttf_advvert"].ref,
                                v["del_ttf_advhoriz"].ref, v["ttf"].ref, v["fct_LO"].ref,
                                v["hnode"].ref, v["hnode_new"].ref, v["fct_adf_v"].ref,
                                v["fct_adf_h"].ref, v["area"].ref, ci(m.myDim_nod2D),
                                ci(m.myDim_edge2D), ci(m.nl), mv["nlevels_nod2D"].ref,
                                mv["nlevels_elem"].ref, mv["edges"].ref, mv["edge_tri"].ref, cd(f.dt))
        return self.alg_state.value

    def step(self, exchange=None):
        """One tracer step the way the Fortran drives the reference library; `exchange(fields)` is
        the host-side MPI exchange_nod of fct_plus / fct_minus between the two awaits."""
        st = self.pre_comm()
        self.stream.sync()
        if st != 6:
            raise abi.AbiError(f"fct_ale_pre_comm_acc_ stopped at alg_state {st}")
        self.inter_comm()           # overlaps the host exchange in the Fortran (md:203-235)
        if exchange is not None:
            exchange(self.f)
        st = self.post_comm()
        if st != 8:
            raise abi.AbiError(f"fct_ale_post_comm_acc_ stopped at alg_state {st}")
        if self.with_c:
            st = self.stage_c()
            if st != 10:
                raise abi.AbiError(f"fct_ale_c_acc_ stopped at alg_state {st}")
        self.stream.sync()
        return st

    def fetch(self, *names):
        """Synchronous download of device-only results (fct_ttf_max/min, UV_rhs)."""
        out = {}
        for k in names:
            if k == "UV_rhs":
                host = np.empty((self.m.myDim_elem2D, self.m.L, 2))
                out[k] = self.vars[k].download(host.reshape(-1)).reshape(host.shape)
            else:
                out[k] = self.vars[k].download().reshape(getattr(self.f, k).shape)
        return out

    def h2d_bytes(self):
        v = self.vars
        n = sum(v[k].host.nbytes for k in ("ttf", "fct_adf_v", "fct_adf_h", "area_inv", "fct_LO",
                                           "fct_plus", "fct_minus"))
        if self.with_c:
            n += sum(v[k].host.nbytes for k in ("area", "hnode", "hnode_new", "del_ttf_advvert",
                                                "del_ttf_advhoriz"))
        return n

    def d2h_bytes(self):
        v = self.vars
        n = sum(v[k].host.nbytes for k in ("fct_plus", "fct_minus", "fct_adf_v", "fct_adf_h"))
        if self.with_c:
            n += sum(v[k].host.nbytes for k in ("del_ttf_advvert", "del_ttf_advhoriz"))
        return n

    def free(self):
        for v in list(self.vars.values()) + list(self.mesh_vars.values()):
            v.free()
        self.stream.free()


class DevicePlan:
    def __init__(self, mesh: Mesh):
        self.lib = abi.load()
        abi.device_info()
        self.m = mesh
        self.h = C.c_void_p()
        st = C.c_int()
        m = mesh
        self.lib.fct_ale_plan_create_(
            C.byref(self.h), ci(m.myDim_nod2D), ci(m.eDim_nod2D), ci(m.myDim_elem2D),
            ci(m.myDim_edge2D), ci(m.nl), iptr(m.nlevels_nod2D), iptr(m.nlevels_elem),
            iptr(m.elem2D_nodes.reshape(-1)), iptr(m.nod_in_elem2D_num),
            iptr(m.nod_in_elem2D.reshape(-1)), ci(m.nod_in_elem2D_dim), iptr(m.edges.reshape(-1)),
            iptr(m.edge_tri.reshape(-1)), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("fct_ale_plan_create_ failed (see stderr)")
        p = C.c_int()
        self.lib.fct_ale_plan_pitch_(C.byref(self.h), C.byref(p))
        self.pitch = p.value
        wt, tt, pk = C.c_int(), C.c_int(), C.c_int()
        self.lib.fct_ale_plan_kernels_(C.byref(self.h), C.byref(wt), C.byref(tt), C.byref(pk))
        self.kernels = "warp" if wt.value else ("tile" if tt.value else "untiled")
        self.packed_ok = bool(pk.value)

    def packed_columns(self, kind: str = "node") -> np.ndarray:
        """Column offsets of the packed level storage: row r owns [col[r], col[r+1]) (uint32, in doubles)."""
        m = self.m
        rows = m.myDim_edge2D if kind == "edge" else m.nnod
        col = np.zeros(rows + 1, np.uint32)
        st = C.c_int()
        self.lib.fct_ale_plan_packed_columns_(C.byref(self.h), ci(1 if kind == "edge" else 0),
                                              col.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("this plan has no packed level storage")
        return col

    def pack_host(self, dense: np.ndarray, kind: str = "node", out: Optional[np.ndarray] = None, alloc=None) -> np.ndarray:
        """Dense host array [rows, W] -> packed host array (what a caller that keeps packed columns holds)."""
        col = self.packed_columns(kind).astype(np.int64)
        if out is None:
            out = (alloc or np.empty)(int(col[-1]))
        out[...] = 0.0
        W = dense.shape[1]
        cnt = np.minimum(np.diff(col), W)
        step = max(1, (1 << 22) // max(W, 1))
        cols = np.arange(W, dtype=np.int64)[None, :]
        for r in range(0, dense.shape[0], step):
            mask = cols < cnt[r:r + step, None]
            dst = (col[:-1][r:r + step, None] + cols)[mask]
            out[dst] = dense[r:r + step][mask]
        return out

    def unpack_host(self, packed: np.ndarray, dense: np.ndarray, kind: str = "node") -> np.ndarray:
        col = self.packed_columns(kind).astype(np.int64)
        W = dense.shape[1]
        cnt = np.minimum(np.diff(col), W)
        step = max(1, (1 << 22) // max(W, 1))
        cols = np.arange(W, dtype=np.int64)[None, :]
        for r in range(0, dense.shape[0], step):
            mask = cols < cnt[r:r + step, None]
            dense[r:r + step][mask] = packed[(col[:-1][r:r + step, None] + cols)[mask]]
        return dense

    def free(self):
        st = C.c_int()
        self.lib.fct_ale_plan_destroy_(C.byref(self.h), C.byref(st))


class HaloLink:
    def __init__(self, plan: DevicePlan, part: Partition, unique_id: bytes):
        self.lib = abi.load()
        peers = sorted(set(part.send_lists) | set(part.recv_ranges))
        send_counts = np.array([part.send_lists.get(p, np.empty(0, np.int32)).size for p in peers], np.int32)
        send_nodes = np.concatenate([part.send_lists.get(p, np.empty(0, np.int32)) for p in peers]
                                    + [np.empty(0, np.int32)]).astype(np.int32)
        recv_first = np.array([part.recv_ranges.get(p, (0, 0))[0] for p in peers], np.int32)
        recv_counts = np.array([part.recv_ranges.get(p, (0, 0))[1] for p in peers], np.int32)
        peers_a = np.array(peers, np.int32)
        if send_nodes.size == 0:
            send_nodes = np.zeros(1, np.int32)
        pad = lambda a: a if a.size else np.zeros(1, np.int32)
        self.h = C.c_void_p()
        st = C.c_int()
        idbuf = C.create_string_buffer(unique_id, 128)
        self.lib.fct_ale_halo_create_(C.byref(self.h), C.byref(plan.h), idbuf, ci(part.rank),
                                      ci(part.nparts), ci(len(peers)), iptr(pad(peers_a)),
                                      iptr(pad(send_counts)), iptr(send_nodes), iptr(pad(recv_first)),
                                      iptr(pad(recv_counts)), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("fct_ale_halo_create_ failed (see stderr)")

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = C.c_int()
        abi.load().fct_ale_comm_unique_id_(buf, C.byref(st))
        if st.value != 0:
            raise abi.AbiError("fct_ale_comm_unique_id_ failed")
        return buf.raw

    def comm_ms(self) -> float:
        """Device time of the exchange inside the last overlapped step (synchronises on it)."""
        ms, st = C.c_double(), C.c_int()
        self.lib.fct_ale_halo_comm_ms_(C.byref(self.h), C.byref(ms), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("fct_ale_halo_comm_ms_ failed (no overlapped step yet?)")
        return ms.value

    def free(self):
        st = C.c_int()
        self.lib.fct_ale_halo_destroy_(C.byref(self.h), C.byref(st))


class DeviceFields:
    """A batch of tracers resident on the device."""
    PER_TRACER = ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "del_ttf_advvert", "del_ttf_advhoriz",
                  "fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus")
    STATIC = ("area", "area_inv", "hnode", "hnode_new")

    def __init__(self, plan: DevicePlan, ntracers: int = 1, with_uv: bool = True, packed: bool = False):
        """packed: the packed level storage of the fast path (active levels only, mode 1 only)."""
        self.lib = abi.load()
        self.plan = plan
        self.T = ntracers
        self.packed = packed
        self.with_uv = with_uv and not packed
        self.h = C.c_void_p()
        st = C.c_int()
        if packed:
            self.lib.fct_ale_fields_create_packed_(C.byref(self.h), C.byref(plan.h), ci(ntracers), C.byref(st))
            m = plan.m
            z = np.arange(m.nl)[None, :]
            # levels that own a slot on the device (a download returns zeros elsewhere)
            self._nslot = z < ((np.maximum(m.nlevels_nod2D.astype(np.int64) - 1, 0) + 2) & ~1)[:, None]
            self._eslot = z < ((m.edge_depth().astype(np.int64) + 1) & ~1)[:, None]
        else:
            self.lib.fct_ale_fields_create_(C.byref(self.h), C.byref(plan.h), ci(ntracers),
                                            ci(1 if with_uv else 0), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("fct_ale_fields_create_ failed (out of device memory?)")
        self.stream = abi.Stream()

    def _copy(self, fn, name, tracer, host):
        st = C.c_int()
        fn(C.byref(self.h), ci(abi.FIELD_IDS[name]), ci(tracer), dptr(host.reshape(-1)), self.stream.ref,
           C.byref(st))
        if st.value != 0:
            raise abi.AbiError(f"transfer of {name} failed")

    def upload_field(self, name: str, host: np.ndarray, tracer: int = 0):
        self._copy(self.lib.fct_ale_field_upload_, name, tracer, host)

    def download_field(self, name: str, host: np.ndarray, tracer: int = 0, merge: bool = True):
        """Asynchronous download.  Packed fields: `merge` keeps the host's values in the cells that
        have no slot on the device (synchronises); without it those cells come back as zeros."""
        if self.packed and merge:
            tmp = np.empty_like(host)
            self._copy(self.lib.fct_ale_field_download_, name, tracer, tmp)
            self.stream.sync()
            slot = (self._eslot if name.startswith("fct_adf_h") else self._nslot)[:, :host.shape[1]]
            np.copyto(host, tmp, where=slot)
        else:
            self._copy(self.lib.fct_ale_field_download_, name, tracer, host)

    def upload(self, f: Fields, tracer: int = 0, static: bool = True, outputs: bool = True):
        names = ["ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "del_ttf_advvert", "del_ttf_advhoriz"]
        if outputs:
            names += ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus"]
            if self.with_uv and f.UV_rhs is not None:
                names.append("UV_rhs")
        if static:
            names += list(self.STATIC)
        for k in names:
            self.upload_field(k, getattr(f, k), tracer)
        if outputs:
            # cells the kernels never write keep the input values, like the reference's in-place update
            self.upload_field("fct_adf_h_out", f.fct_adf_h, tracer)
            self.upload_field("fct_adf_v_out", f.fct_adf_v, tracer)
        self.stream.sync()

    def download(self, f: Fields, tracer: int = 0, mode: int = 1, names=None) -> Fields:
        """Fetch results into (a copy of) `f`; the limited fluxes come from the OUT buffers in the
        fused modes."""
        out = f.copy()
        if names is None:
            names = ["fct_ttf_max", "fct_ttf_min", "fct_plus", "fct_minus",
                     "del_ttf_advvert", "del_ttf_advhoriz"]
            if mode == 0 and self.with_uv and f.UV_rhs is not None:
                names.append("UV_rhs")
            self.download_field("fct_adf_h_out" if mode >= 1 else "fct_adf_h", out.fct_adf_h, tracer)
            self.download_field("fct_adf_v_out" if mode >= 1 else "fct_adf_v", out.fct_adf_v, tracer)
        for k in names:
            self.download_field(k, getattr(out, k), tracer)
        self.stream.sync()
        return out

    def step(self, f: Fields, mode: int = 1, halo: Optional[HaloLink] = None, sync: bool = True) -> int:
        st = C.c_int()
        hp = C.byref(halo.h) if halo is not None else None
        self.lib.fct_ale_step_(C.byref(self.h), hp, self.stream.ref, ci(mode), cd(f.dt),
                               cd(f.flux_eps), cd(f.bignumber), C.byref(st))
        if sync:
            self.stream.sync()
        return st.value

    def step_general(self, f: Fields, halo: Optional[HaloLink] = None, sync: bool = True) -> int:
        """The subroutine with its vlimit (f.vlimit) and iter_yn (f.iter_yn) branches, stage kernels."""
        st = C.c_int()
        hp = C.byref(halo.h) if halo is not None else None
        self.lib.fct_ale_step_general_(C.byref(self.h), hp, self.stream.ref, ci(f.vlimit), ci(1 if f.iter_yn else 0),
                                       cd(f.dt), cd(f.flux_eps), cd(f.bignumber), C.byref(st))
        if sync:
            self.stream.sync()
        return st.value

    def exchange_field(self, halo: HaloLink, name: str):
        st = C.c_int()
        self.lib.fct_ale_halo_exchange_field_(C.byref(self.h), C.byref(halo.h), self.stream.ref,
                                              ci(abi.FIELD_IDS[name]), C.byref(st))
        if st.value != 0:
            raise abi.AbiError(f"halo exchange of {name} failed")

    # the arrays that change every tracer step (hnode / hnode_new move with the ALE surface) and the
    # results FESOM2 consumes after fct_ale (the advective tendencies, docs/refactoring.md:292-314)
    STEP_INPUTS = ("ttf", "fct_LO", "fct_adf_v", "fct_adf_h", "hnode", "hnode_new", "del_ttf_advvert",
                   "del_ttf_advhoriz")
    STEP_RESULTS = ("del_ttf_advvert", "del_ttf_advhoriz")

    def host_step(self, f: Fields, out: Fields, mode: int = 1, halo: Optional[HaloLink] = None, tracer: int = 0) -> int:
        """One end-to-end tracer step for a caller whose fields live on the HOST: upload this step's
        inputs from `f`, run a1..c on the device, download the tendencies into `out`, wait."""
        for k in self.STEP_INPUTS:
            self.upload_field(k, getattr(f, k), tracer)
        st = self.step(f, mode=mode, halo=halo, sync=False)
        for k in self.STEP_RESULTS:
            self.download_field(k, getattr(out, k), tracer, merge=False)
        self.stream.sync()
        return st

    def host_steps(self, f: Fields, out: Fields, steps: int, mode: int = 1, halo: Optional[HaloLink] = None,
                   tracer: int = 0) -> int:
        """`steps` end-to-end tracer steps back to back for a host-resident caller, with the download
        of step k (its own stream) overlapping the upload of step k+1: PCIe is full duplex.  Every
        step uploads all its inputs and downloads its tendencies; del_ttf_adv* (inputs AND results)
        are uploaded last, after the previous step's download has left the device."""
        if not hasattr(self, "_dn"):
            self._dn = abi.Stream()
            self._ev_step, self._ev_dn = abi.Event(), abi.Event()
        st = 0
        first = [k for k in self.STEP_INPUTS if k not in self.STEP_RESULTS]
        for k in range(steps):
            for name in first:
                self.upload_field(name, getattr(f, name), tracer)
            if k > 0:
                self.stream.wait(self._ev_dn)
            for name in self.STEP_RESULTS:
                self.upload_field(name, getattr(f, name), tracer)
            st = self.step(f, mode=mode, halo=halo, sync=False)
            self._ev_step.record(self.stream)
            self._dn.wait(self._ev_step)
            for name in self.STEP_RESULTS:
                stt = C.c_int()
                self.lib.fct_ale_field_download_(C.byref(self.h), ci(abi.FIELD_IDS[name]), ci(tracer),
                                                 dptr(getattr(out, name).reshape(-1)), self._dn.ref, C.byref(stt))
                if stt.value != 0:
                    raise abi.AbiError(f"download of {name} failed")
            self._ev_dn.record(self._dn)
        self._dn.sync()
        self.stream.sync()
        return st

    # ---- host arrays already in the packed level storage ----
    def upload_packed(self, name: str, host_packed: np.ndarray, tracer: int = 0, stream=None):
        st = C.c_int()
        self.lib.fct_ale_field_upload_packed_(C.byref(self.h), ci(abi.FIELD_IDS[name]), ci(tracer), dptr(host_packed),
                                              (stream or self.stream).ref, C.byref(st))
        if st.value != 0:
            raise abi.AbiError(f"packed upload of {name} failed")

    def download_packed(self, name: str, host_packed: np.ndarray, tracer: int = 0, stream=None):
        st = C.c_int()
        self.lib.fct_ale_field_download_packed_(C.byref(self.h), ci(abi.FIELD_IDS[name]), ci(tracer), dptr(host_packed),
                                                (stream or self.stream).ref, C.byref(st))
        if st.value != 0:
            raise abi.AbiError(f"packed download of {name} failed")

    def host_steps_batch(self, f: Fields, out: Fields, tracers: int, mode: int = 1, halo: Optional[HaloLink] = None,
                         packed_host: Optional[dict] = None) -> int:
        """One model TIME STEP for a host-resident caller: `tracers` tracer steps back to back (T, S and
        the passive tracers of docs/refactoring.md's caller), the mesh-static inputs of the step (hnode,
        hnode_new: they move with the ALE surface once per time step, not per tracer) uploaded ONCE, the
        per-tracer inputs (ttf, fct_LO, fct_adf_v, fct_adf_h, del_ttf_adv*) uploaded and the tendencies
        downloaded for EVERY tracer; downloads on their own stream overlap the next tracer's uploads.
        packed_host: name -> host array already in the packed level storage (then no repack, 70 % of
        the bytes); else the dense arrays of `f` / `out`."""
        if not hasattr(self, "_dn"):
            self._dn = abi.Stream()
            self._ev_step, self._ev_dn = abi.Event(), abi.Event()
        ph = packed_host

        def up(name):
            if ph is not None:
                self.upload_packed(name, ph[name])
            else:
                self.upload_field(name, getattr(f, name))

        st = 0
        for name in ("hnode", "hnode_new"):
            up(name)
        first = [k for k in self.STEP_INPUTS if k not in self.STEP_RESULTS and k not in ("hnode", "hnode_new")]
        for k in range(tracers):
            for name in first:
                up(name)
            if k > 0:
                self.stream.wait(self._ev_dn)
            for name in self.STEP_RESULTS:
                up(name)
            st = self.step(f, mode=mode, halo=halo, sync=False)
            self._ev_step.record(self.stream)
            self._dn.wait(self._ev_step)
            for name in self.STEP_RESULTS:
                if ph is not None:
                    self.download_packed(name, ph["out_" + name], stream=self._dn)
                else:
                    stt = C.c_int()
                    self.lib.fct_ale_field_download_(C.byref(self.h), ci(abi.FIELD_IDS[name]), ci(0),
                                                     dptr(getattr(out, name).reshape(-1)), self._dn.ref, C.byref(stt))
                    if stt.value != 0:
                        raise abi.AbiError(f"download of {name} failed")
            self._ev_dn.record(self._dn)
        self._dn.sync()
        self.stream.sync()
        return st

    def link_bytes(self, name: str, host: np.ndarray, upload: bool) -> int:
        """Bytes that cross PCIe for this field and host array (the library's own count: the dense
        array, or only the slots of the packed storage when the host array is page-locked)."""
        n = C.c_longlong()
        self.lib.fct_ale_field_link_bytes_(C.byref(self.h), ci(abi.FIELD_IDS[name]), dptr(host.reshape(-1)),
                                           ci(1 if upload else 0), C.byref(n))
        return n.value

    def host_step_bytes(self, f: Fields):
        return (sum(self.link_bytes(k, getattr(f, k), True) for k in self.STEP_INPUTS),
                sum(self.link_bytes(k, getattr(f, k), False) for k in self.STEP_RESULTS))

    def stage(self, name: str, f: Fields, sync: bool = True):
        st = C.c_int()
        self.lib.fct_ale_stage_(C.byref(self.h), self.stream.ref, ci(abi.STAGE_IDS[name]), cd(f.dt),
                                cd(f.flux_eps), cd(f.bignumber), C.byref(st))
        if st.value != 0:
            raise abi.AbiError(f"stage {name} failed")
        if sync:
            self.stream.sync()

    def exchange(self, halo: HaloLink):
        st = C.c_int()
        self.lib.fct_ale_halo_exchange_(C.byref(self.h), C.byref(halo.h), self.stream.ref, C.byref(st))
        if st.value != 0:
            raise abi.AbiError("halo exchange failed")

    def free(self):
        st = C.c_int()
        self.lib.fct_ale_fields_destroy_(C.byref(self.h), C.byref(st))
        self.stream.free()


# --------------------------------------------------------------------------------------------
# stress2rhs: device-resident through handles, or host arrays in / out
# --------------------------------------------------------------------------------------------
STRESS_INPUTS = ("ice_strength", "elem_area", "sigma11", "sigma12", "sigma22", "gradient_sca", "metric_factor",
                 "inv_areamass", "rhs_a", "rhs_m")


class StressChain:
    """stress2rhs_plan_create_ + handles of alloc_var_ + stress2rhs_acc_ (asynchronous)."""

    def __init__(self, d):
        self.lib = abi.load()
        self.d = d
        self.plan = C.c_void_p()
        st = C.c_int()
        en = np.ascontiguousarray(d["elem2D_nodes"], dtype=np.int32)
        self.lib.stress2rhs_plan_create_(C.byref(self.plan), ci(d["N"]), ci(d["E"]), ci(en.shape[1]),
                                         iptr(en.reshape(-1)), C.byref(st))
        if st.value != 0:
            raise abi.AbiError("stress2rhs_plan_create_ failed")
        self.stream = abi.Stream()
        self.u, self.v = np.full(d["N"], np.nan), np.full(d["N"], np.nan)
        self.vars = {k: abi.Var(d[k]) for k in STRESS_INPUTS}
        self.vars["U"], self.vars["V"] = abi.Var(self.u), abi.Var(self.v)
        for k in STRESS_INPUTS:
            self.vars[k].upload()

    def run(self, sync=True):
        v, st = self.vars, C.c_int()
        self.lib.stress2rhs_acc_(C.byref(self.plan), self.stream.ref, v["U"].ref, v["V"].ref, *[v[k].ref for k in STRESS_INPUTS],
                                 C.byref(st))
        if st.value != 0:
            raise abi.AbiError("stress2rhs_acc_ failed")
        if sync:
            self.stream.sync()

    def fetch(self):
        self.vars["U"].download()
        self.vars["V"].download()
        return self.u, self.v

    def free(self):
        for v in self.vars.values():
            v.free()
        st = C.c_int()
        self.lib.stress2rhs_plan_destroy_(C.byref(self.plan), C.byref(st))
        self.stream.free()


def stress2rhs_host(d):
    """stress2rhs_ on host arrays (the argument order of src/reference.cpp:440)."""
    lib = abi.load()
    u, v = np.full(d["N"], np.nan), np.full(d["N"], np.nan)
    st = C.c_int()
    en = np.ascontiguousarray(d["elem2D_nodes"], dtype=np.int32)
    lib.stress2rhs_(ci(d["N"]), ci(d["E"]), ci(en.shape[1]), dptr(u), dptr(v), dptr(d["ice_strength"]),
                    iptr(en.reshape(-1)), dptr(d["elem_area"]), dptr(d["sigma11"]), dptr(d["sigma12"]),
                    dptr(d["sigma22"]), dptr(d["gradient_sca"]), dptr(d["metric_factor"]),
                    dptr(d["inv_areamass"]), dptr(d["rhs_a"]), dptr(d["rhs_m"]), C.byref(st))
    if st.value != 0:
        raise abi.AbiError("stress2rhs_ failed")
    return u, v
