"""Host driver above the TraceBackend mirror: one Lumice config -> XYZ images of all its renderers.

The reference splits this job between `Simulator::Run` (SimBatch loop: one BeginSession ... EndSession
bracket per batch and wavelength, simulator.cpp:1498-1560), the server's dispatch/drain cadence
(server.cpp:140-151,1447-1457) and `RenderConsumer` (server/render.cpp). With a device that traces
~4 G rays/s the per-batch host work is what limits throughput, so this driver

* issues sessions of `session_rays` rays (default 16 Mi, not 262 144) and never synchronises inside a frame:
  the accumulators persist across sessions (third-clock drain, trace_backend.hpp:495-506);
* projects every exit through ALL renderers of the config in the same trace (hb_set_renders) instead of
  refusing multi-renderer configs (server.cpp:402-437);
* shards the global ray-index range over ranks (sharding.session_plan) and sums the accumulators with one
  NCCL reduce to rank 0 at frame end; only rank 0 returns frames;
* hands back either raw XYZ (ReadbackXyzAccum semantics) or the 8-bit sRGB frame rendered on the device
  (hb_snapshot = PrepareSnapshot + PostSnapshot).
"""
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _abi as A
from .backend import (B200TraceBackend, RootRaySource, SceneTables, SessionSpec, make_wl_entry, make_wl_pool)
from .config import SceneConfig
from .sharding import session_plan


@dataclass
class Frame:
    xyz: np.ndarray                  # [H, W, 3] float32 accumulated XYZ (sum over ranks if all-reduced)
    landed_weight: float             # snapshot intensity
    rgb: Optional[np.ndarray] = None  # [H, W, 3] uint8 sRGB, when requested


def trace_session(backend: B200TraceBackend, layer_cnt: int, spec: SessionSpec, ray_num: int):
    """BeginSession -> (TraceLayer -> Recombine)* -> TraceLayer -> EndSession without exit materialisation
    (SimulateOneWavelengthWithBackend, simulator.cpp:1498-1560, device-fused branch)."""
    backend.BeginSession(spec)
    try:
        roots = RootRaySource.FromHost(ray_num)
        for li in range(layer_cnt):
            last = li + 1 == layer_cnt
            handle = backend.TraceLayer(roots, want_stats=False)
            if last:
                break
            roots = backend.Recombine(handle, shuffle=True)
    finally:
        backend.EndSession()


def ensure_comm(backend: B200TraceBackend, rank: int, world: int):
    """Create the engine's NCCL communicator for a `world`-rank frame (hb_comm_init) unless it has one. The unique id
    travels over torch.distributed, whose process group the launcher (torchrun) has initialised."""
    if world <= 1 or backend.HasComm():
        return
    import torch.distributed as dist
    from .backend import comm_unique_id
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("render_config(world > 1): initialise torch.distributed first (the NCCL unique id is "
                           "broadcast over it), or call backend.CommInit yourself")
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    backend.CommInit(ids[0], rank, world)


def stochastic_populations(desc: A.HbSceneDesc):
    """(layer, population) of every population whose shape scalars are random (IsDeterministic, simulator.cpp:453-471)."""
    out = []
    for li in range(desc.layer_cnt):
        layer = desc.layers[li]
        for pi in range(layer.population_cnt):
            c = layer.populations[pi].crystal
            heights = c.height[:1] if c.kind == 0 else c.height[:3]
            if any(d.type != 0 for d in list(heights) + list(c.face_dist)):
                out.append((li, pi))
    return out


def render_config(cfg: SceneConfig, backend: Optional[B200TraceBackend] = None, seed: int = 42,
                  session_rays: int = 1 << 24, rank: int = 0, world: int = 1, geometry_seed: int = 1,
                  wl_pool_size: int = 64, srgb: bool = False, intensity_factor: float = 1.0,
                  allreduce: bool = True, device_geometry: bool = True) -> Dict[int, Frame]:
    """Trace `cfg` (from config.load_config) and return {renderer id: Frame}.

    Discrete spectrum: `rays_per_wavelength()` roots per wavelength, one single-entry pool per session
    (ray_num_semantics.hpp:12-16). Illuminant spectrum: all rays in sessions that carry the M-entry pool and
    draw a per-ray wavelength index (wl_pool.hpp:73-84). Ray indices are global: rank r of `world` traces its
    contiguous share of each wavelength's index range, so any world size traces the same set of rays.
    Stochastic crystal shapes: with `device_geometry` every session gets a fresh pool (`cfg.desc.geom_pool_size`
    shapes per population, one shape per 32 consecutive rays) drawn and built on the device one session ahead
    (hb_auto_resample); ranks draw from disjoint ranges of the shape stream.
    """
    if not cfg.renders:
        raise ValueError("config has no renderer")
    if len(cfg.renders) > A.HB_MAX_RENDERS:
        raise ValueError(f"more than {A.HB_MAX_RENDERS} renderers")
    own = backend is None
    be = backend or B200TraceBackend(rank if own and world > 1 else 0)
    try:
        tables = SceneTables(cfg.desc, geometry_seed)
        be.SetScene(tables)
        ids = sorted(cfg.renders)
        be.SetRenders([cfg.renders[i] for i in ids])
        for r in range(len(ids)):
            be.ReadbackXyzAccum(render=r)            # a frame starts from zero accumulators
        if cfg.illuminant is not None:
            jobs = [(make_wl_pool(cfg.illuminant, wl_pool_size), cfg.ray_num_total)]
        else:
            if not cfg.spectrum:
                raise ValueError("config has neither a discrete spectrum nor an illuminant")
            n = cfg.rays_per_wavelength()
            jobs = [([make_wl_entry(wl, w)], n) for wl, w in cfg.spectrum]
        index_base = 0
        if device_geometry and cfg.desc.geom_pool_size > 1:
            for k, (li, pi) in enumerate(stochastic_populations(cfg.desc)):   # engine-run geometry clock
                be.AutoResample(li, pi, cfg.desc.layers[li].populations[pi].crystal, geometry_seed,
                                ((rank * 64 + k) << 22) & 0xFFFFFFFF)
        for pool, total in jobs:
            for first, count in session_plan(total, rank, world, session_rays, index_base):
                trace_session(be, cfg.desc.layer_cnt,
                              SessionSpec(seed=seed, wl=pool, ray_num=count, accumulate=True, ray_base=first), count)
            index_base += total
        if world > 1 and allreduce:   # frame end: one fp32 reduce to rank 0, which alone reads the frame back
            ensure_comm(be, rank, world)
            be.ReduceImage(0)
            if rank != 0:
                be.Synchronize()
                return {}
        frames = {}
        for r, rid in enumerate(ids):
            rgb = None
            if srgb:
                rgb, _, _ = be.Snapshot(r, intensity_factor)
            xyz, landed = be.ReadbackXyzAccum(render=r)
            frames[rid] = Frame(xyz, landed, rgb)
        return frames
    finally:
        if own:
            be.close()
