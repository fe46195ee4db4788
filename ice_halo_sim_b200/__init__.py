"""halotrace-b200: Blackwell-native ice-halo trace engine behind Lumice's TraceBackend seam.

Only the trace hot path lives here: `csrc/` holds the sm_100a kernels and the
C ABI (include/halotrace_b200.h); `backend.py` mirrors the reference's TraceBackend interface over that
ABI; `config.py` reads Lumice JSON configs. There is no CPU implementation in this package.
"""
from . import _abi  # noqa: F401
from .backend import (B200TraceBackend, BackendUnavailableError, HaloTraceError, LayerHandle, RootRaySource,  # noqa: F401
                      SceneTables, SessionSpec, comm_unique_id, make_proj_params, make_wl_entry, make_wl_pool, simulate)
from .config import SceneConfig, load_config  # noqa: F401
from .driver import Frame, render_config, trace_session  # noqa: F401

__version__ = "0.1.0"
