"""Loader + typed prototypes for libhalotrace_b200.so (the C ABI in include/halotrace_b200.h).

The library is built in-tree by `ice_halo_sim_b200/csrc/Makefile` (see __graft_entry__.build). There is
no fallback: if the shared object is missing or a compute call cannot reach an sm_100 device, an
exception is raised — nothing in this package computes on the CPU.
"""
import ctypes as C
import os

from . import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("HALOTRACE_LIB", os.path.join(HERE, "libhalotrace_b200.so"))  # override: A/B builds

_vp = C.c_void_p

# name -> (restype, argtypes); every function declared in include/halotrace_b200.h
PROTOTYPES = {
    "hb_abi_version": (C.c_uint32, []),
    "hb_last_error": (C.c_char_p, [_vp]),
    "hb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "hb_destroy": (None, [_vp]),
    "hb_set_scene": (C.c_int, [_vp, _vp]),
    "hb_set_render": (C.c_int, [_vp, _vp]),
    "hb_set_renders": (C.c_int, [_vp, C.c_uint32, _vp]),
    "hb_begin_session": (C.c_int, [_vp, _vp]),
    "hb_trace_layer": (C.c_int, [_vp, C.c_uint64, _vp]),
    "hb_recombine": (C.c_int, [_vp, C.c_int, _vp]),
    "hb_end_session": (C.c_int, [_vp]),
    "hb_readback_xyz": (C.c_int, [_vp, _vp, _vp]),
    "hb_readback_xyz_render": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "hb_export_root_masks": (C.c_int, [_vp, C.c_uint64, _vp, _vp]),
    "hb_readback_class_lanes": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "hb_snapshot": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "hb_drain_exits": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "hb_inject_rays": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "hb_export_roots": (C.c_int, [_vp, C.c_uint64] + [_vp] * 8),
    "hb_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "hb_resample_shapes": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32, _vp]),
    "hb_auto_resample": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32]),
    "hb_export_shapes": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "hb_pyramid_slope": (C.c_double, [C.c_float]),
    "hb_get_counters": (C.c_int, [_vp, _vp]),
    "hb_synchronize": (C.c_int, [_vp]),
    "hb_selftest_arith": (C.c_int, [_vp, C.c_uint32, C.c_uint64, C.c_uint32, _vp]),
    "hb_image_device_ptr": (C.c_int, [_vp, _vp, _vp]),
    "hb_stream": (_vp, [_vp]),
    "hb_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "hb_comm_unique_id": (C.c_int, [_vp]),
    "hb_allreduce_image": (C.c_int, [_vp]),
    "hb_reduce_image": (C.c_int, [_vp, C.c_int]),
    "hb_merge_from_peer": (C.c_int, [_vp, _vp]),
    "hb_make_prism": (C.c_int, [C.c_float, _vp, _vp]),
    "hb_make_pyramid": (C.c_int, [C.c_float] * 5 + [_vp, _vp]),
    "hb_make_axis_sampler": (C.c_int, [C.c_uint32, C.c_float, C.c_float] * 3 + [_vp]),
    "hb_ice_refractive_index": (C.c_double, [C.c_double]),
    "hb_make_proj_params": (C.c_int, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_int, C.c_int, C.c_float, _vp]),
    "hb_partition_rays": (C.c_int, [_vp, C.c_uint32, C.c_uint64, _vp, _vp]),
    "hb_build_scene": (C.c_int, [_vp, C.c_uint32, C.POINTER(_vp)]),
    "hb_scene_tables_get": (_vp, [_vp]),
    "hb_free_scene": (None, [_vp]),
    "hb_build_render": (C.c_int, [_vp, _vp]),
    "hb_make_wl_entry": (C.c_int, [C.c_float, C.c_float, _vp]),
    "hb_make_wl_pool_illuminant": (C.c_int, [C.c_int, C.c_uint32, _vp]),
    "hb_illuminant_spd": (C.c_float, [C.c_int, C.c_float]),
}

_lib = None


class HaloTraceError(RuntimeError):
    """Any non-zero HbStatus. `status` carries the code; HB_ERR_NO_DEVICE / HB_ERR_CUDA correspond to the
    reference's BackendUnavailableError (trace_backend.hpp:139-158)."""

    def __init__(self, status, message):
        super().__init__(f"{A.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class BackendUnavailableError(HaloTraceError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the engine has no CPU fallback)")
        lib = C.CDLL(os.environ.get("HALOTRACE_B200_LIB", SO_PATH))   # override: A/B builds in experiments
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        if lib.hb_abi_version() != 2:
            raise ImportError("libhalotrace_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(status, handle=None):
    if status == A.HB_OK:
        return
    msg = load().hb_last_error(handle)
    msg = msg.decode() if msg else ""
    if status in (-2, -3):
        raise BackendUnavailableError(status, msg)
    raise HaloTraceError(status, msg)
