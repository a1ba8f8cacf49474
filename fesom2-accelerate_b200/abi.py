"""ctypes binding of include/fesom2-accelerate.h -- the stub a maintainer of the reference would
write instead of the Fortran ISO_C_BINDING interface block, and the harness
that stands in for the reference's kernel_tuner scripts.  No torch, no oracle, no CPU fallback: if
libfesom2-accelerate.so is missing or no CUDA device is present, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfesom2-accelerate.so")

# every symbol include/fesom2-accelerate.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "set_mpi_rank_", "transfer_mesh_", "alloc_var_", "reserve_var_", "allocate_pinned_doubles_",
    "transfer_var_", "transfer_var_async_", "make_stream_", "await_stream_",
    "fct_ale_pre_comm_acc_", "fct_ale_inter_comm_acc_", "fct_ale_post_comm_acc_",
    "fct_ale_a1_accelerated", "fct_ale_a2_accelerated", "fct_ale_a1_a2_accelerated",
    "fct_ale_a1_reference_", "fct_ale_a2_reference_", "fct_ale_a3_reference_",
    "fct_ale_a4_reference_", "fct_ale_pre_comm_",
    "fct_ale_c_acc_", "transfer_var_back_", "transfer_var_back_async_", "free_var_",
    "free_pinned_doubles_", "free_stream_", "fct_ale_set_fused_", "fct_ale_tune_", "fct_ale_launch_count_",
    "fct_ale_device_info_", "fct_ale_event_create_", "fct_ale_event_record_", "fct_ale_stream_wait_event_",
    "fct_ale_event_elapsed_ms_", "fct_ale_event_destroy_", "fct_ale_mem_info_",
    "fct_ale_plan_create_", "fct_ale_plan_destroy_", "fct_ale_plan_pitch_", "fct_ale_plan_inspect_", "fct_ale_plan_kernels_",
    "fct_ale_fields_create_", "fct_ale_fields_create_packed_", "fct_ale_fields_destroy_", "fct_ale_field_upload_",
    "fct_ale_field_download_", "fct_ale_field_link_bytes_", "fct_ale_step_", "fct_ale_step_general_", "fct_ale_halo_exchange_field_",
    "stress2rhs_plan_create_", "stress2rhs_plan_destroy_", "stress2rhs_acc_", "stress2rhs_", "fct_ale_stage_",
    "fct_ale_comm_unique_id_", "fct_ale_halo_create_", "fct_ale_halo_destroy_",
    "fct_ale_halo_exchange_", "fct_ale_halo_comm_ms_", "fct_ale_trace_read_",
    "fct_ale_plan_packed_size_", "fct_ale_plan_packed_columns_", "fct_ale_field_upload_packed_", "fct_ale_field_download_packed_",
]

# enum fct_field_id
FIELD_IDS = dict(ttf=0, fct_LO=1, fct_adf_v=2, fct_adf_h=3, area=4, area_inv=5, hnode=6,
                 hnode_new=7, del_ttf_advvert=8, del_ttf_advhoriz=9, fct_ttf_max=10, fct_ttf_min=11,
                 fct_plus=12, fct_minus=13, UV_rhs=14, fct_adf_h_out=15, fct_adf_v_out=16,
                 fct_adf_v2=17, fct_adf_h2=18)
STAGE_IDS = dict(a1=0, a2=1, a3=2, b1v=3, b1h=4, b2=5, b3v=6, b3h=7, cv=8, ch=9, phaseA=10,
                 phaseB=11, phaseA_tile=12, phaseB_tile=13, phaseA_tile_boundary=14,
                 phaseA_tile_interior=15, phaseB_tile_boundary=16, phaseB_tile_interior=17,
                 phaseA_warp=18, phaseB_warp=19, phaseA_warp_boundary=20, phaseA_warp_interior=21,
                 phaseB_warp_boundary=22, phaseB_warp_interior=23, b1h_atomic=24, ch_atomic=25,
                 a3_vlimit2=26, a3_vlimit3=27, b3v_iter=28, b3h_iter=29, lo_update=30)


class GpuMemory(C.Structure):
    """struct gpuMemory of the header (leading members = the reference's struct)."""
    _fields_ = [("host_pointer", C.c_void_p), ("device_pointer", C.c_void_p), ("size", C.c_size_t),
                ("event", C.c_void_p), ("has_event", C.c_int), ("event_recorded", C.c_int),
                ("magic", C.c_uint)]


class AbiError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the product library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AbiError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()) first; "
                           "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def ci(v):
    return C.byref(C.c_int(int(v)))


def cd(v):
    return C.byref(C.c_double(float(v)))


def cb(v):
    return C.byref(C.c_bool(bool(v)))


def dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_int))


def device_info():
    """(name, (major, minor), sm_count); raises AbiError when no CUDA device is usable."""
    lib = load()
    name = C.create_string_buffer(64)
    maj, mnr, sms, st = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.fct_ale_device_info_(name, C.byref(maj), C.byref(mnr), C.byref(sms), C.byref(st))
    if st.value != 0:
        raise AbiError("no CUDA device: the fct_ale path has no CPU fallback")
    return name.value.decode(), (maj.value, mnr.value), sms.value


def launch_count() -> int:
    n = C.c_longlong()
    load().fct_ale_launch_count_(C.byref(n))
    return n.value


def set_fused(flag: bool) -> None:
    load().fct_ale_set_fused_(ci(1 if flag else 0))


def tune(name: str, value: int) -> None:
    """Tuning knob FCT_<name> (see the header); never changes results."""
    load().fct_ale_tune_(C.c_char_p(name.encode()), ci(value))


# --------------------------------------------------------------------------------------------
# Reference-style handle ABI (what FESOM2's Fortran calls)
# --------------------------------------------------------------------------------------------
class Stream:
    def __init__(self):
        self.h = C.c_void_p()
        st = C.c_int()
        load().make_stream_(C.byref(self.h), C.byref(st))
        if st.value != 0:
            raise AbiError("make_stream_ failed")

    def sync(self):
        st = C.c_int()
        load().await_stream_(C.byref(self.h), C.byref(st))
        if st.value != 0:
            raise AbiError("await_stream_ failed (see stderr)")

    def wait(self, event: "Event"):
        """Work issued to this stream from now on starts after `event` (recorded on another stream)."""
        st = C.c_int()
        load().fct_ale_stream_wait_event_(C.byref(self.h), C.byref(event.h), C.byref(st))
        if st.value != 0:
            raise AbiError("fct_ale_stream_wait_event_ failed")

    def free(self):
        st = C.c_int()
        load().free_stream_(C.byref(self.h), C.byref(st))

    @property
    def ref(self):
        return C.byref(self.h)


class Var:
    """alloc_var_ / reserve_var_ / transfer_mesh_ handle."""

    def __init__(self, host: np.ndarray | None = None, size: int | None = None, event: bool = False,
                 mesh: bool = False):
        lib = load()
        self.h = C.c_void_p()
        self.host = host
        st = C.c_int()
        if mesh:
            lib.transfer_mesh_(C.byref(self.h), iptr(host), ci(host.size), C.byref(st))
        elif host is not None:
            lib.alloc_var_(C.byref(self.h), dptr(host), ci(host.size), cb(event), C.byref(st))
        else:
            lib.reserve_var_(C.byref(self.h), ci(size), cb(event), C.byref(st))
        if st.value != 0 or not self.h.value:
            raise AbiError("device allocation failed (see stderr)")

    @property
    def ref(self):
        return C.byref(self.h)

    def upload(self, host: np.ndarray | None = None, stream: Stream | None = None, record=False):
        host = self.host if host is None else host
        self.host = host
        if stream is None:
            load().transfer_var_(self.ref, dptr(host))
        else:
            load().transfer_var_async_(self.ref, dptr(host), stream.ref, cb(record))

    def download(self, host: np.ndarray | None = None, stream: Stream | None = None):
        host = self.host if host is None else host
        self.host = host
        if stream is None:
            load().transfer_var_back_(self.ref, dptr(host))
        else:
            load().transfer_var_back_async_(self.ref, dptr(host), stream.ref)
        return host

    def free(self):
        st = C.c_int()
        load().free_var_(self.ref, C.byref(st))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Page-locked numpy array through allocate_pinned_doubles_ (8-byte units)."""
    n = int(np.prod(shape))
    nbytes = n * np.dtype(dtype).itemsize
    nd = (nbytes + 7) // 8
    p = C.c_void_p()
    st = C.c_int()
    load().allocate_pinned_doubles_(C.byref(p), ci(nd), C.byref(st))
    if not p.value:
        raise AbiError("allocate_pinned_doubles_ failed")
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    return arr


class Event:
    """CUDA event on a harness stream (timing on the device, never by wall clock)."""

    def __init__(self):
        self.h = C.c_void_p()
        st = C.c_int()
        load().fct_ale_event_create_(C.byref(self.h), C.byref(st))
        if st.value != 0:
            raise AbiError("event creation failed")

    def record(self, stream: Stream):
        st = C.c_int()
        load().fct_ale_event_record_(C.byref(self.h), stream.ref, C.byref(st))
        if st.value != 0:
            raise AbiError("event record failed")

    def ms_since(self, start: "Event") -> float:
        ms = C.c_double()
        st = C.c_int()
        load().fct_ale_event_elapsed_ms_(C.byref(start.h), C.byref(self.h), C.byref(ms), C.byref(st))
        if st.value != 0:
            raise AbiError("event elapsed failed (kernel fault? see stderr)")
        return ms.value

    def free(self):
        st = C.c_int()
        load().fct_ale_event_destroy_(C.byref(self.h), C.byref(st))


def mem_info():
    fr, tot, st = C.c_longlong(), C.c_longlong(), C.c_int()
    load().fct_ale_mem_info_(C.byref(fr), C.byref(tot), C.byref(st))
    if st.value != 0:
        raise AbiError("cudaMemGetInfo failed")
    return fr.value, tot.value
