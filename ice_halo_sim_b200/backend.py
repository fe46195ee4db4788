"""B200TraceBackend — host-side mirror of the reference's TraceBackend seam over the C ABI.

Method names, argument meaning, call order and error behaviour follow
`lumice::TraceBackend` (reference: src/core/backend/trace_backend.hpp:367-641):

    BeginSession -> (TraceLayer -> DrainExits -> Recombine)* -> TraceLayer -> DrainExits -> EndSession

and `simulate()` follows the reference driver `Simulator::SimulateOneWavelengthWithBackend`
(src/core/simulator.cpp:1479-1694). All compute happens in libhalotrace_b200.so on the GPU; this
module only marshals POD structs. The C++ adapter a Lumice maintainer would compile in is in adapter/.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi as A
from .lib import BackendUnavailableError, HaloTraceError, check, load  # noqa: F401

EXIT_DTYPE = np.dtype([("dir", np.float32, 3), ("weight", np.float32), ("path_len", np.uint8),
                       ("path", np.uint8, 64), ("pad0", np.uint8), ("crystal_id", np.uint16),
                       ("ms_layer_idx", np.uint8), ("wl_idx", np.uint8), ("pad1", np.uint8, 2),
                       ("component_mask", np.uint64)])
assert EXIT_DTYPE.itemsize == 96  # lumice::ExitRayRecord, exit_seam.hpp:40-53


@dataclass
class SessionSpec:
    """trace_backend.hpp:197-215. `wl` is the wavelength pool: [(n_idx, spd_weight, cmf_x, cmf_y, cmf_z)]
    (one entry for a discrete-wavelength session)."""
    seed: int
    wl: List[tuple]
    ray_num: int = 0
    record_exits: int = 0           # 1/True: ExitRayRecords + exported roots (parity); 2: records only (exit-seam
                                    # egress, TraceBackend::DrainExits); 0: production = fused accumulate
    accumulate: bool = True
    ray_base: Optional[int] = None  # multi-GPU sharding: global index of this session's first root


@dataclass
class LayerHandle:
    """trace_backend.hpp:309-332 + LayerStats (:297-300)."""
    root_count: int = 0
    continuation_count: int = 0
    exit_count: int = 0
    exit_w_sum: float = 0.0

    def ContinuationCount(self):
        return self.continuation_count

    def GetLayerStats(self):
        return self.exit_count, self.exit_w_sum


@dataclass
class RootRaySource:
    """trace_backend.hpp:259-276: FromHost{count} (engine generates roots) or FromDevice (continuations)."""
    is_device: bool = False
    count: int = 0
    # HostRayBatch ray injection (parity only; crystal-local rays, cpu_trace_backend.cpp:121-144)
    d: Optional[np.ndarray] = None
    p: Optional[np.ndarray] = None
    w: Optional[np.ndarray] = None
    tf: Optional[np.ndarray] = None
    rot: Optional[np.ndarray] = None

    @staticmethod
    def FromHost(count, d=None, p=None, w=None, tf=None, rot=None):
        return RootRaySource(False, int(count), d, p, w, tf, rot)

    @staticmethod
    def FromDevice(count):
        return RootRaySource(True, int(count))


class SceneTables:
    """Owns an HbSceneTables* built by hb_build_scene (host-side MakeCrystal / LatLut / filter tables)."""

    def __init__(self, desc: A.HbSceneDesc, geometry_seed=1):
        self._lib = load()
        self._h = C.c_void_p()
        check(self._lib.hb_build_scene(C.byref(desc), int(geometry_seed), C.byref(self._h)))
        self.desc = desc

    @property
    def scene_ptr(self):
        return self._lib.hb_scene_tables_get(self._h)

    def scene(self) -> A.HbScene:
        return C.cast(self.scene_ptr, C.POINTER(A.HbScene)).contents

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.hb_free_scene(self._h)
            self._h = C.c_void_p()


def make_wl_entry(wavelength_nm, weight=1.0):
    e = A.HbWlEntry()
    check(load().hb_make_wl_entry(float(wavelength_nm), float(weight), C.byref(e)))
    return (e.n_idx, e.spd_weight, e.cmf_x, e.cmf_y, e.cmf_z)


ILLUMINANTS = {"D50": 0, "D55": 1, "D65": 2, "D75": 3, "A": 4, "E": 5}  # util/illuminant_data.hpp:12-28


def make_wl_pool(illuminant, m=64):
    """Illuminant-mode wavelength pool (ComputeWlPool, backend/wl_pool.hpp:73-84): `m` mid-point samples of
    [380, 780] nm weighted by the illuminant's SPD; m defaults to the reference's kWlPoolSizeDefault."""
    if isinstance(illuminant, str):
        if illuminant not in ILLUMINANTS:
            raise ValueError(f"unknown illuminant {illuminant!r}")
        illuminant = ILLUMINANTS[illuminant]
    pool = (A.HbWlEntry * int(m))()
    check(load().hb_make_wl_pool_illuminant(int(illuminant), int(m), pool))
    return [(e.n_idx, e.spd_weight, e.cmf_x, e.cmf_y, e.cmf_z) for e in pool]


def make_proj_params(render: A.HbRenderDesc) -> A.HbProjParams:
    p = A.HbProjParams()
    check(load().hb_build_render(C.byref(render), C.byref(p)))
    return p


class B200TraceBackend:
    """One engine instance on one GPU (the reference creates one backend per Simulator::Run thread)."""

    def __init__(self, device_ordinal=0):
        self._lib = load()
        self._h = C.c_void_p()
        check(self._lib.hb_create(int(device_ordinal), C.byref(self._h)))
        self._keep = []
        self._in_session = False
        self._pending: Optional[LayerHandle] = None
        self._proj: Optional[A.HbProjParams] = None

    # ---- lifetime ----
    def close(self):
        if self._h and self._h.value:
            self._lib.hb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def _check(self, status):
        check(status, self._h)

    # ---- scene / render (captured by BeginSession in the reference; set explicitly here) ----
    def SetScene(self, scene):
        """scene: SceneTables, or a ctypes pointer/address of an HbScene built by the caller."""
        ptr = scene.scene_ptr if isinstance(scene, SceneTables) else scene
        self._keep = [scene]
        self._check(self._lib.hb_set_scene(self._h, ptr))

    def SetRender(self, render):
        """render: HbRenderDesc (built with the library's BuildProjParams twin) or HbProjParams."""
        self.SetRenders([render])

    def SetRenders(self, renders):
        """N projections of one trace (hb_set_renders): every exit lands in each render's own accumulator."""
        projs = [make_proj_params(r) if isinstance(r, A.HbRenderDesc) else r for r in renders]
        arr = (A.HbProjParams * len(projs))(*projs)
        self._check(self._lib.hb_set_renders(self._h, len(projs), arr))
        self._projs = projs
        self._proj = projs[0]

    def IsCompatible(self, render) -> bool:  # trace_backend.hpp:511-514: all 11 lens types supported
        return 0 <= int(render.lens_type if isinstance(render, A.HbRenderDesc) else render.proj_type) <= 10

    # ---- the seam ----
    def SupportsDeviceXyzAccum(self) -> bool:
        return True

    def SupportsThirdClockDrain(self) -> bool:
        return True

    def WlPoolSize(self) -> int:
        return 64  # kWlPoolSizeDefault, backend/wl_pool.hpp:38

    def BeginSession(self, spec: SessionSpec):
        wl = (A.HbWlEntry * len(spec.wl))(*[A.HbWlEntry(*e) for e in spec.wl])
        s = A.HbSessionSpec()
        s.seed = int(spec.seed) & 0xFFFFFFFF
        s.wl_cnt = len(spec.wl)
        s.wl = wl
        s.ray_num = int(spec.ray_num)
        s.record_exits = int(spec.record_exits)
        s.accumulate = 1 if spec.accumulate else 0
        if spec.ray_base is not None:
            s.ray_base = int(spec.ray_base)
            s.use_ray_base = 1
        self._check(self._lib.hb_begin_session(self._h, C.byref(s)))
        self._in_session = True
        self._spec = spec

    def TraceLayer(self, roots: RootRaySource, want_stats=True) -> LayerHandle:
        n = 0 if roots.is_device else roots.count
        if not roots.is_device and roots.d is not None:
            d = np.ascontiguousarray(roots.d, np.float32)
            p = np.ascontiguousarray(roots.p, np.float32)
            w = np.ascontiguousarray(roots.w, np.float32)
            tf = np.ascontiguousarray(roots.tf, np.uint16)
            rot = None if roots.rot is None else np.ascontiguousarray(roots.rot, np.float32)
            self._check(self._lib.hb_inject_rays(self._h, len(w), d.ctypes.data, p.ctypes.data, w.ctypes.data,
                                                 tf.ctypes.data, None if rot is None else rot.ctypes.data))
            n = len(w)
        st = A.HbLayerStats()
        self._check(self._lib.hb_trace_layer(self._h, int(n), C.byref(st) if want_stats else None))
        h = LayerHandle(st.root_count, st.continuation_count, st.exit_count, st.exit_w_sum)
        self._pending = h
        return h

    def Recombine(self, handle: LayerHandle, shuffle=True) -> RootRaySource:
        cnt = C.c_uint64()
        self._check(self._lib.hb_recombine(self._h, 1 if shuffle else 0, C.byref(cnt)))
        self._pending = None
        return RootRaySource.FromDevice(cnt.value)

    def DrainExits(self, with_roots=False):
        cnt = C.c_uint64()
        self._check(self._lib.hb_drain_exits(self._h, None, None, 0, C.byref(cnt)))
        n = cnt.value
        out = np.zeros(n, EXIT_DTYPE)
        roots = np.zeros(n, np.uint32)
        if n:
            self._check(self._lib.hb_drain_exits(self._h, out.ctypes.data, roots.ctypes.data, n, C.byref(cnt)))
        return (out, roots) if with_roots else out

    def ExportRoots(self):
        """Parity helper: the crystal-local roots the engine generated for the last traced layer."""
        cnt = C.c_uint64()
        self._check(self._lib.hb_export_roots(self._h, 0, None, None, None, None, None, None, None, C.byref(cnt)))
        n = cnt.value
        r = dict(d=np.zeros((n, 3), np.float32), p=np.zeros((n, 3), np.float32), w=np.zeros(n, np.float32),
                 face=np.zeros(n, np.uint16), rot=np.zeros((n, 9), np.float32), shape=np.zeros(n, np.uint32),
                 wl=np.zeros(n, np.uint32))
        if n:
            self._check(self._lib.hb_export_roots(self._h, n, r["d"].ctypes.data, r["p"].ctypes.data,
                                                  r["w"].ctypes.data, r["face"].ctypes.data, r["rot"].ctypes.data,
                                                  r["shape"].ctypes.data, r["wl"].ctypes.data, C.byref(cnt)))
        return r

    def ExportRootMasks(self):
        """Parity only: component masks the exported roots carried in (raypath colour), or an empty array."""
        cnt = C.c_uint64(0)
        self._check(self._lib.hb_export_root_masks(self._h, 0, None, C.byref(cnt)))
        out = np.zeros(cnt.value, np.uint64)
        if cnt.value:
            self._check(self._lib.hb_export_root_masks(self._h, cnt.value, out.ctypes.data, C.byref(cnt)))
        return out

    def EndSession(self):
        self._check(self._lib.hb_end_session(self._h))
        self._in_session = False

    def ReadbackXyzAccum(self, xyz: Optional[np.ndarray] = None, render=0):
        """Returns (xyz[H, W, 3] float32, landed_weight). Drains + zeroes that render's device accumulators."""
        if self._proj is None:
            raise HaloTraceError(-4, "ReadbackXyzAccum before SetRender")
        if not 0 <= render < len(self._projs):
            raise HaloTraceError(-1, "ReadbackXyzAccum: no such render")
        h, w = self._projs[render].img_h, self._projs[render].img_w
        if xyz is None:
            xyz = np.empty((h, w, 3), np.float32)
        landed = C.c_float(0.0)
        self._check(self._lib.hb_readback_xyz_render(self._h, render, xyz.ctypes.data, C.byref(landed)))
        return xyz, landed.value

    def ReadbackClassLanes(self):
        """TraceBackend::ReadbackClassLanes (trace_backend.hpp:471-493): [class, H, W] float32 Y lanes of render 0,
        drained + zeroed; None when the scene has no colour classes."""
        if self._proj is None:
            raise HaloTraceError(-4, "ReadbackClassLanes before SetRender")
        h, w = self._projs[0].img_h, self._projs[0].img_w
        buf = np.zeros((A.HB_MAX_COLOR_CLASSES, h, w), np.float32)
        cnt = C.c_uint32(0)
        self._check(self._lib.hb_readback_class_lanes(self._h, buf.ctypes.data, buf.size, C.byref(cnt)))
        return buf[: cnt.value].copy() if cnt.value else None

    def Snapshot(self, render=0, intensity_factor=1.0, ray_color=(-1.0, -1.0, -1.0), background=(0.0, 0.0, 0.0),
                 want_xyz=False):
        """Device display sink (RenderConsumer::PrepareSnapshot + PostSnapshot, server/render.cpp:463-577):
        returns (rgb8[H, W, 3] uint8, xyz[H, W, 3] float32 or None, snapshot_intensity). Non-destructive."""
        if self._proj is None:
            raise HaloTraceError(-4, "Snapshot before SetRender")
        if not 0 <= render < len(self._projs):
            raise HaloTraceError(-1, "Snapshot: no such render")
        h, w = self._projs[render].img_h, self._projs[render].img_w
        desc = A.HbSnapshotDesc(float(intensity_factor), (C.c_float * 3)(*ray_color), (C.c_float * 3)(*background))
        rgb = np.empty((h, w, 3), np.uint8)
        xyz = np.empty((h, w, 3), np.float32) if want_xyz else None
        inten = C.c_float(0.0)
        self._check(self._lib.hb_snapshot(self._h, render, C.byref(desc), rgb.ctypes.data,
                                          xyz.ctypes.data if want_xyz else None, C.byref(inten)))
        return rgb, xyz, inten.value

    # ---- stochastic geometry pool on the device ----
    def ResampleShapes(self, layer, population, crystal: A.HbCrystalDesc, seed, draw_base=0):
        """Redraw every shape of one population's pool on the device (hb_resample_shapes): shape scalars from
        `crystal`'s height / face-distance distributions, tables built by the code hb_make_prism/pyramid run.
        Returns the number of shapes the builder rejected (they become empty crystals)."""
        rejected = C.c_uint32()
        self._check(self._lib.hb_resample_shapes(self._h, int(layer), int(population), C.byref(crystal), int(seed),
                                                 int(draw_base), C.byref(rejected)))
        return rejected.value

    def AutoResample(self, layer, population, crystal: Optional[A.HbCrystalDesc], seed=0, draw_base=0):
        """Geometry clock run by the engine (hb_auto_resample): a fresh device-built pool for every session,
        prepared one session ahead. crystal=None switches it off."""
        self._check(self._lib.hb_auto_resample(self._h, int(layer), int(population),
                                               C.byref(crystal) if crystal is not None else None, int(seed), int(draw_base)))

    def ExportShapes(self, layer, population):
        """Parity helper: (HbCrystalTables array, scalars [n, 10] = h1, h2, h3, d0..d5, builder status)."""
        cnt = C.c_uint32()
        self._check(self._lib.hb_export_shapes(self._h, int(layer), int(population), 0, None, None, C.byref(cnt)))
        n = cnt.value
        tables = (A.HbCrystalTables * n)()
        scalars = np.zeros((n, 10), np.float32)
        self._check(self._lib.hb_export_shapes(self._h, int(layer), int(population), n, tables, scalars.ctypes.data,
                                               C.byref(cnt)))
        return tables, scalars

    # ---- tuning / measurement ----
    def SetOption(self, key, value):
        self._check(self._lib.hb_set_option(self._h, key.encode(), int(value)))

    def Synchronize(self):
        self._check(self._lib.hb_synchronize(self._h))

    def SelftestArith(self, mode: int, n: int, seed: int = 1):
        """hb_selftest_arith: (mismatches, a_bits, b_bits, result_bits) of the unchecked exact division / sqrt."""
        out = (C.c_uint64 * 4)()
        self._check(self._lib.hb_selftest_arith(self._h, int(mode), int(n), int(seed), out))
        return tuple(int(v) for v in out)

    def Counters(self) -> A.HbCounters:
        c = A.HbCounters()
        self._check(self._lib.hb_get_counters(self._h, C.byref(c)))
        return c

    def ImageDevicePtr(self):
        ptr = C.c_void_p()
        n = C.c_uint64()
        self._check(self._lib.hb_image_device_ptr(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def AllReduceImage(self):
        """Sum of all ranks' accumulators on EVERY rank (refuses to run twice on the same accumulation)."""
        self._check(self._lib.hb_allreduce_image(self._h))

    def ReduceImage(self, root=0):
        """Frame end: fp32 ncclReduce of the accumulators (all renders + colour lanes) to `root`; the other ranks'
        accumulators are zero afterwards, so only the root reads back."""
        self._check(self._lib.hb_reduce_image(self._h, int(root)))

    def MergeFromPeer(self, other: "B200TraceBackend"):
        """Same process, another device: add `other`'s accumulators into this engine's (P2P loads) and zero them."""
        self._check(self._lib.hb_merge_from_peer(self._h, other._h))

    def HasComm(self) -> bool:
        return getattr(self, "_comm_ranks", 1) > 1

    def CommInit(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._lib.hb_comm_init(self._h, buf, rank, nranks))
        self._comm_ranks = int(nranks)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(load().hb_comm_unique_id(buf))
    return buf.raw


def simulate(backend: B200TraceBackend, layer_cnt: int, spec: SessionSpec, ray_num: int, want_stats=False,
             drain_exits=False):
    """One session over all scattering layers — the reference driver loop
    (SimulateOneWavelengthWithBackend, simulator.cpp:1498-1560)."""
    backend.BeginSession(spec)
    exits = []
    handles = []
    try:
        roots = RootRaySource.FromHost(ray_num)
        for mi in range(layer_cnt):
            last = mi + 1 == layer_cnt
            handle = backend.TraceLayer(roots, want_stats=want_stats or not last)
            handles.append(handle)
            if drain_exits:
                exits.append(backend.DrainExits(with_roots=True))
            if last:
                break
            roots = backend.Recombine(handle, shuffle=True)
    finally:
        backend.EndSession()
    return handles, exits
