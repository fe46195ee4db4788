#!/usr/bin/env python
"""bench.py — Mrays/s of the trace hot path on B200 (BASELINE.json metric), contract in the task brief.

One "step" = one full pass of the workload per GPU: config 2 of BASELINE.json
(examples/config_example.json as shipped: prism h=1.3, zenith gauss(90, 0.3), max_hits 7, sun alt 20,
9 wavelengths x 50 M root rays, render id 4 = fisheye_equal_area 1920x1080) = 450 M root rays.
Rays are generated on the device by the engine's counter-based RNG (data: synthetic); weak scaling:
every rank traces its own 450 M rays (disjoint global ray-index ranges); at frame end the per-GPU XYZ
accumulators are reduced (NCCL, fp32) to rank 0, which alone reads the frame back.

  value : root rays / s with the scene tables resident on the device (kernels + frame-end reduce)
  e2e   : same, through the TraceBackend-shaped public API with HOST buffers: per step the scene /
          wavelength tables are uploaded (hb_set_scene, hb_begin_session) and the XYZ image is read back
          into (pinned) host memory (hb_readback_xyz) inside the timed region
  configs : BASELINE configs[2..4] (plate + raypath filter, two-layer multi-scatter, stochastic prisms) at
          their per-GPU ray counts, one timed pass each, device-resident and e2e
  strong : (N > 1) the fixed 9 x 50 M frame split over the N ranks
  gpu_reference : (N = 1) the reference's own CudaTraceBackend (cuda_trace_backend.cu compiled for sm_100a,
          oracle/_ref/libhalo_refcuda.so) driven by its unmodified Simulator::Run on the same GPU, and the same
          driver on this repo's engine through the adapter (oracle/_ref/libhalo_refb200.so)
  --impl reference : the reference's own multi-threaded legacy CPU path (oracle/_ref perf build when
          present, else the oracle port) on a bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WAVELENGTHS = [450.0, 490.0, 530.0, 570.0, 610.0, 650.0, 690.0, 730.0, 770.0]
RAYS_PER_WL = 50_000_000
SESSION_RAYS = 1 << 24          # rays per BeginSession/EndSession bracket (the driver's SimBatch size)
# Algorithmic HBM bytes per ray-bounce of each kernel (DESIGN.md "kernels"): the state is P, D, Q = three float4 per ray.
BYTES_OPTICS = 64               # split pipeline: read P,D,Q (48) + write D (16); + 16 per emitted exit (one v4 reduction)
BYTES_OPTICS_LAST = 48          # final interaction writes no state
BYTES_EXIT = 16
BYTES_INTERSECT = 48            # read P,D (32) + write P (16)
BYTES_GEN = 48
# fused bounce kernel (one launch per interaction): read P,D (32) + write P,D (32); per emitted exit the
# orientation quaternion is read (16) and one v4 reduction goes to the image (16); the final interaction writes no state
BYTES_BOUNCE = 64
BYTES_BOUNCE_LAST = 32
BYTES_BOUNCE_EXIT = 32
# root generation fused with the entry interaction: writes P, D, Q once (48); the external reflection of every
# root leaves at hit 0 (1 exit per root: Q is still in flight, only the 16-byte reduction reaches the image)
BYTES_GENBOUNCE = 48
BYTES_GENBOUNCE_EXIT = 16
EXITS_AT_ENTRY = 1.0
EXITS_PER_ROOT = 4.7            # config 2, measured (reference CPU: 4.69-4.71)
TRAFFIC_FILE = "traffic_r2b.json"

# BASELINE.json configs -> scene (ice_halo_sim_b200.scenes), ray budget per wavelength, GPUs the config names,
# session size. config2 is the headline the bench line is quoted on; the others are reported as sub-records of the
# same line at the per-GPU share of their BASELINE ray count.
WORKLOADS = {
    "config2": dict(case="column_config2", rays_per_wl=RAYS_PER_WL, gpus=1, session=SESSION_RAYS,
                    what="BASELINE configs[1]: config_example.json as shipped, prism h=1.3 zenith gauss(90,0.3) "
                         "max_hits 7, fisheye_equal_area 1920x1080"),
    "config3": dict(case="plate_filter_config3", rays_per_wl=200_000_000, gpus=1, session=SESSION_RAYS,
                    what="BASELINE configs[2]: plate h=0.3 zenith gauss(0,0.8), raypath filter [3,5] symmetry P, "
                         "max_hits 7, fisheye_equal_area 1920x1080, 9 x 200 M rays"),
    "config4": dict(case="two_layer_config4", rays_per_wl=500_000_000, gpus=4, session=1 << 22,
                    what="BASELINE configs[3]: two layers, plate (prob 1.0) over full-sphere column, max_hits 7, "
                         "fisheye_equal_area 1920x1080; every exit of layer 0 re-enters layer 1; 9 x 500 M roots over 4 GPUs"),
    "config5": dict(case="stoch_config5", rays_per_wl=1_000_000_000, gpus=8, session=SESSION_RAYS,
                    what="BASELINE configs[4]: bench_config_stoch.json prism h=1 d_i~gauss(1,0.15) full-sphere axis, "
                         "max_hits 8, rectangular 2048x1024 full sky; 256-shape pool redrawn on the device every session, "
                         "one shape per 32 consecutive rays; 9 x 1 G rays over 8 GPUs"),
    # not a BASELINE config: the generic (non-prism) slab kernels on a 20-face pyramid, for the kernel notes in DESIGN.md
    "pyramid": dict(case="pyramid", rays_per_wl=RAYS_PER_WL, gpus=1, session=SESSION_RAYS,
                    what="parity case 'pyramid': 20-face hexagonal pyramid (10 paired axes, generic slab kernels), "
                         "max_hits 8, dual_fisheye_equal_area 1024x512 full sky"),
}


def workload_desc(name="config2"):
    from ice_halo_sim_b200 import scenes
    case = scenes.CASES[WORKLOADS[name]["case"]]
    return case["scene"](), case["render"]()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
                for nme, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def cpu_baseline(steps_budget_s=15.0, threads=None):
    """Reference legacy CPU path on a bounded sample of the workload; returns the cpu_baseline object.
    (The only place besides tests/ and smoke() that executes anything under oracle/.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    desc, rdesc = workload_desc()
    perf_so = os.path.join(ROOT, "oracle", "_ref", "libhalo_ref_perf.so")
    cores = os.cpu_count() or 1
    if os.path.exists(perf_so):
        lib = C.CDLL(perf_so)
        vp = C.c_void_p
        lib.ref_legacy_bench.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
        lib.ref_physical_cores.restype = C.c_uint32
        phys = int(lib.ref_physical_cores()) or cores
        nthreads = threads or phys
        import numpy as np
        wl = np.array(WAVELENGTHS, np.float32)
        ww = np.ones(len(wl), np.float32)
        rate, sec, exits = C.c_double(), C.c_double(), C.c_uint64()
        # calibrate on a small sample, then size the run for ~steps_budget_s
        per_wl = 2000 * nthreads
        lib.ref_legacy_bench(C.byref(desc), C.byref(rdesc), wl.ctypes.data, ww.ctypes.data, len(wl), per_wl, nthreads,
                             128, C.byref(rate), C.byref(sec), C.byref(exits))
        per_wl = max(per_wl, int(rate.value * steps_budget_s / len(wl)))
        lib.ref_legacy_bench(C.byref(desc), C.byref(rdesc), wl.ctypes.data, ww.ctypes.data, len(wl), per_wl, nthreads,
                             128, C.byref(rate), C.byref(sec), C.byref(exits))
        return {"value": rate.value / 1e6, "unit": "Mrays/s", "cores": nthreads, "kind": "reference",
                "sample": f"{len(wl)} wavelengths x {per_wl} root rays, legacy Simulator::Run x{nthreads} threads + "
                          f"host ScatterOutgoingToXyz consumer, 128-ray dispatch, {sec.value:.1f} s",
                "physical_cores": phys, "logical_cpus": cores}
    # oracle port, single thread
    import numpy as np
    import harness as H
    import parity
    from ice_halo_sim_b200 import backend as B
    A = H.A
    tables = B.SceneTables(desc, 7)
    sc = tables.scene()
    wl = [B.make_wl_entry(WAVELENGTHS[0], 1.0)]
    wl_arr = (A.HbWlEntry * 1)(*[A.HbWlEntry(*e) for e in wl])
    orc = H.oracle()
    n = 200000
    t0 = time.time()
    done = 0
    while time.time() - t0 < steps_budget_s:
        r = dict(d=np.zeros((n, 3), np.float32), p=np.zeros((n, 3), np.float32), w=np.zeros(n, np.float32),
                 face=np.zeros(n, np.uint16), rot=np.zeros((n, 9), np.float32), shape=np.zeros(n, np.uint32),
                 wl=np.zeros(n, np.uint32))
        orc.orc_gen_roots(tables.scene_ptr, 0, 0, 0, C.addressof(wl_arr), 1, 42, done, n, H.ptr(r["d"]), H.ptr(r["p"]),
                          H.ptr(r["w"]), H.ptr(r["face"]), None, H.ptr(r["rot"]), H.ptr(r["shape"]), H.ptr(r["wl"]))
        lp, keep = parity.layer_params(sc, 0, wl_arr, 42, done)
        ex, er, _ = parity.oracle_trace(lp, r, n * 9)
        parity.oracle_image(B.make_proj_params(rdesc), wl_arr, ex)
        done += n
    sec = time.time() - t0
    return {"value": done / sec / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": f"1 wavelength x {done} root rays, oracle gen+trace+accumulate, single thread, {sec:.1f} s"}


def reference_driver_run(lib_name, rays_per_wl, dispatch, devices=None):
    """The reference's unmodified Simulator::Run on its TraceBackend route (oracle/ref_driver.cpp: ref_backend_bench)
    over the config-2 scene; the backend is fixed by which oracle/_ref library is loaded. Runs in a child process
    (its own CUDA context). Returns a dict or None when the library is absent."""
    so = os.path.join(ROOT, "oracle", "_ref", lib_name)
    if not os.path.exists(so):
        return None
    code = f"""
import ctypes as C, json, sys
sys.path.insert(0, {ROOT!r})
import numpy as np
import bench
desc, rd = bench.workload_desc()
lib = C.CDLL({so!r})
vp = C.c_void_p
lib.ref_backend_bench.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, vp, vp, vp]
wl = np.array(bench.WAVELENGTHS, np.float32); ww = np.ones(len(wl), np.float32)
img = np.zeros((rd.img_h, rd.img_w, 3), np.float32)
landed, rate, sec, used = C.c_double(), C.c_double(), C.c_double(), C.c_uint32()
for rays in ({max(dispatch, rays_per_wl // 8)}, {rays_per_wl}):          # warm-up pass, timed pass
    lib.ref_backend_bench(C.byref(desc), C.byref(rd), wl.ctypes.data, ww.ctypes.data, len(wl), rays, {dispatch}, 42,
                          img.ctypes.data, C.byref(landed), C.byref(rate), C.byref(sec), C.byref(used))
print(json.dumps(dict(mrays_per_s=rate.value / 1e6, seconds=sec.value, backend_used=int(used.value),
                      landed_per_root=landed.value / ({rays_per_wl} * len(wl)),
                      image_sum_per_root=float(img.astype(np.float64).sum()) / ({rays_per_wl} * len(wl)))))
"""
    env = dict(os.environ)
    if devices:
        env["HALOTRACE_B200_DEVICES"] = ",".join(str(d) for d in devices)
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        line = [x for x in p.stdout.strip().splitlines() if x.startswith("{")]
        if p.returncode != 0 or not line:
            return {"error": (p.stderr or p.stdout)[-300:]}
        r = json.loads(line[-1])
        r.update(rays_per_wl=rays_per_wl, dispatch_rays=dispatch, wavelengths=len(WAVELENGTHS))
        return r
    except Exception as ex:  # the comparison leg is optional
        return {"error": str(ex)[:300]}


def run_reference(args, rank):
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    cb = None
    budget = max(5.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(budget)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    line = {"impl": "reference", "metric": "Mrays/sec (9λ×50M single-scatter)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (time.time() - t0) * 1e3 / max(1, args.steps + args.warmup), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config_example.json scene (9 wavelengths, prism h=1.3 zenith gauss(90,0.3), "
                                   "max_hits 7), fisheye_equal_area 1920x1080; bounded CPU sample per step"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Workload:
    """One BASELINE config on one engine: scene / render set, sessions sized, the geometry clock started."""

    def __init__(self, be, name, rank, world, np, A, B, driver, geom_pool=0, pinned=None):
        self.be, self.name, self.rank, self.world = be, name, rank, world
        self.B, self.driver = B, driver
        self.wk = WORKLOADS[name]
        self.desc, self.rdesc = workload_desc(name)
        if geom_pool:
            self.desc.geom_pool_size = geom_pool
        elif name == "config5":
            self.desc.geom_pool_size = 256   # the adapter's pool size; redrawn on the device at every session
        self.stochastic = driver.stochastic_populations(self.desc) if self.desc.geom_pool_size > 1 else []
        self.max_hits = int(self.desc.max_hits)
        self.layer_cnt = int(self.desc.layer_cnt)
        self.session_rays = self.wk["session"]
        self.tables = B.SceneTables(self.desc, 7)
        self.wl_entries = [B.make_wl_entry(x, 1.0) for x in WAVELENGTHS]
        h, w = self.rdesc.img_h, self.rdesc.img_w
        if pinned is not None and pinned.numel() >= h * w * 3:
            self.host_img = pinned[: h * w * 3].view(h, w, 3).numpy()
        else:
            self.host_img = np.empty((h, w, 3), np.float32)
        self.scene_bytes = C.sizeof(A.HbCrystalTables) + C.sizeof(A.HbAxisSampler) + C.sizeof(A.HbFilterDesc) + \
            C.sizeof(A.HbProjParams)
        self.step_ctr = 0
        self.activate()

    def activate(self):
        self.be.SetScene(self.tables)
        self.be.SetRender(self.rdesc)
        self.start_geometry_clock()

    def start_geometry_clock(self):   # stochastic geometry: a fresh shape pool per session, built on the device one session ahead
        for k, (li, pi) in enumerate(self.stochastic):
            self.be.AutoResample(li, pi, self.desc.layers[li].populations[pi].crystal, 7,
                                 ((self.rank * 64 + k) << 22) & 0xFFFFFFFF)

    def trace_step(self, rays_per_wl, e2e, shard=None):
        """One full pass: 9 wavelengths x rays_per_wl roots in session_rays-sized sessions. shard = (rank, world):
        this rank traces only its contiguous share of every wavelength's index range (strong scaling)."""
        be, B = self.be, self.B
        step_idx = self.step_ctr
        self.step_ctr += 1
        if e2e:  # host tables travel every step
            self.activate()
        for wi, wl in enumerate(self.wl_entries):
            if shard is None:
                begin, end = 0, rays_per_wl
                base0 = ((step_idx * self.world + self.rank) * len(WAVELENGTHS) + wi) * rays_per_wl
            else:
                r, wd = shard
                begin = r * (rays_per_wl // wd) + min(r, rays_per_wl % wd)
                end = begin + rays_per_wl // wd + (1 if r < rays_per_wl % wd else 0)
                base0 = (step_idx * len(WAVELENGTHS) + wi) * rays_per_wl
            done = begin
            while done < end:
                n = min(self.session_rays, end - done)
                self.driver.trace_session(be, self.layer_cnt, B.SessionSpec(seed=42, wl=[wl], ray_num=n, accumulate=True,
                                                                            ray_base=base0 + done), n)
                done += n
        if self.world > 1:
            be.ReduceImage(0)          # fp32 ncclReduce to rank 0; the other ranks' accumulators are zero afterwards
        if e2e and self.rank == 0:
            be.ReadbackXyzAccum(self.host_img)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--rays-per-wl", type=int, default=0)
    ap.add_argument("--geom-pool", type=int, default=0, help="shapes per stochastic population (config5)")
    ap.add_argument("--tile-rays", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs / strong / gpu_reference sub-records")
    ap.add_argument("--set", action="append", default=[], help="engine option key=value (experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from ice_halo_sim_b200 import _abi as A
    from ice_halo_sim_b200 import backend as B
    from ice_halo_sim_b200 import driver

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    be = B.B200TraceBackend(local_rank)
    if args.tile_rays:
        be.SetOption("tile_rays", args.tile_rays)
    for kv in args.set:
        k, v = kv.split("=")
        be.SetOption(k, int(v))
    pinned = torch.empty(2048 * 1080 * 3, dtype=torch.float32).pin_memory()   # e2e reads the frame back into pinned memory
    wk = Workload(be, args.workload, rank, world, np, A, B, driver, args.geom_pool, pinned)
    stream = torch.cuda.ExternalStream(be._lib.hb_stream(be._h), device=torch.device("cuda", local_rank))
    driver.ensure_comm(be, rank, world)

    rays_per_wl = args.rays_per_wl or max(1, wk.wk["rays_per_wl"] // wk.wk["gpus"])
    rays_per_step = rays_per_wl * len(WAVELENGTHS)
    h, w = wk.rdesc.img_h, wk.rdesc.img_w

    def barrier():
        be.Synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        c0 = be.Counters().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall_ms = (time.time() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        ms = max(dev_ms, 0.0) if dev_ms > 0 else wall_ms
        t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), be.Counters().kernel_launches - c0

    def drain(wl_obj=None):
        be.Synchronize()
        return be.ReadbackXyzAccum((wl_obj or wk).host_img)

    for _ in range(args.warmup):
        wk.trace_step(rays_per_wl, False)
    drain()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_ms, wall_ms, launches = timed(lambda: wk.trace_step(rays_per_wl, False), args.steps)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    img, landed = drain()
    img_sum = float(img.astype(np.float64).sum())

    # e2e through the public API with host buffers (tables up, image down, every step)
    wk.trace_step(rays_per_wl, True)
    e2e_ms, e2e_wall_ms, _ = timed(lambda: wk.trace_step(rays_per_wl, True), args.steps)
    e2e_ms = max(e2e_ms, e2e_wall_ms)  # the D2H read blocks the host: wall clock bounds the step
    drain()

    # per-kernel live timing for the roofline (CUDA events around every launch, short pass)
    def profile_pass(wl_obj, per_wl):
        be.SetOption("profile", 1)
        c0 = be.Counters()
        p_rays = min(per_wl, wl_obj.session_rays)
        driver.trace_session(be, wl_obj.layer_cnt, B.SessionSpec(seed=42, wl=[wl_obj.wl_entries[0]], ray_num=p_rays,
                                                                 accumulate=True, ray_base=1 << 40), p_rays)
        be.Synchronize()
        c1 = be.Counters()
        be.SetOption("profile", 0)
        be.ReadbackXyzAccum(wl_obj.host_img)
        return c0, c1

    def exits_per_root_of(wl_obj):  # filter-passing exits per layer-0 root, counted by the engine (LayerStats)
        be.BeginSession(B.SessionSpec(seed=42, wl=[wl_obj.wl_entries[0]], ray_num=1 << 20, accumulate=True, ray_base=1 << 41))
        hdl = be.TraceLayer(B.RootRaySource.FromHost(1 << 20), want_stats=True)
        be.EndSession()
        be.ReadbackXyzAccum(wl_obj.host_img)
        return hdl.exit_count / float(1 << 20)

    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "fallback (B200_PROFILING.md)"
    hbm_peak = 6650.0
    if os.path.exists(pk_path):
        hbm_peak = float(json.load(open(pk_path)).get("hbm_gbs", hbm_peak))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"

    def roofline_of(c0, c1, exits_per_root, hits):
        """hits = interactions traced per ray of the profiled layer (max_hits, or the filter's bound)."""
        hits = max(1, hits)
        epb = exits_per_root / hits          # exits per ray-bounce
        fam = {}
        with_genbounce = c1.genbounce_launches > c0.genbounce_launches and hits > 1
        for name, ms, n_launch, rays, bytes_per_ray in (
                ("bounce", c1.bounce_ms - c0.bounce_ms, c1.bounce_launches - c0.bounce_launches,
                 c1.bounce_rays - c0.bounce_rays,
                 ((BYTES_BOUNCE * (hits - 2) + BYTES_BOUNCE_LAST) / (hits - 1) +
                  BYTES_BOUNCE_EXIT * (exits_per_root - EXITS_AT_ENTRY) / (hits - 1)) if with_genbounce else
                 (BYTES_BOUNCE * (hits - 1) + BYTES_BOUNCE_LAST) / hits + BYTES_BOUNCE_EXIT * epb),
                ("genbounce", c1.genbounce_ms - c0.genbounce_ms, c1.genbounce_launches - c0.genbounce_launches,
                 c1.genbounce_rays - c0.genbounce_rays, BYTES_GENBOUNCE + BYTES_GENBOUNCE_EXIT * EXITS_AT_ENTRY),
                ("optics", c1.optics_ms - c0.optics_ms, c1.optics_launches - c0.optics_launches,
                 c1.optics_rays - c0.optics_rays,
                 (BYTES_OPTICS * (hits - 1) + BYTES_OPTICS_LAST) / hits + BYTES_EXIT * epb),
                ("intersect", c1.intersect_ms - c0.intersect_ms, c1.intersect_launches - c0.intersect_launches,
                 c1.intersect_rays - c0.intersect_rays, BYTES_INTERSECT)):
            if n_launch == 0:
                continue
            avg_ms = ms / n_launch
            tile = rays / n_launch
            ach = tile * bytes_per_ray / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
            fam[name] = {"avg_ms": avg_ms, "launches": int(n_launch), "rays_per_launch": tile,
                         "algorithmic_bytes_per_ray_bounce": bytes_per_ray, "achieved_gbs": ach,
                         "frac": ach / hbm_peak, "total_ms": ms}
        gen_ms = c1.gen_ms - c0.gen_ms
        gen_l = c1.gen_launches - c0.gen_launches
        gen = {"avg_ms": gen_ms / max(1, gen_l), "launches": int(gen_l), "algorithmic_bytes_per_root": BYTES_GEN}
        all_ms = sum(f["total_ms"] for f in fam.values()) + gen_ms
        dom = max(fam, key=lambda k: fam[k]["total_ms"])
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
        if os.path.exists(tr_path):  # DRAM bytes per ray-bounce from the committed ncu --set full capture
            tr = json.load(open(tr_path))["dram_bytes_per_ray_bounce"]
            if dom in tr:
                traffic = tr[dom] * fam[dom]["rays_per_launch"]
        return {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": fam[dom]["achieved_gbs"],
                "peak": hbm_peak, "unit": "GB/s", "frac": fam[dom]["frac"], "traffic": traffic,
                "algorithmic_bytes_per_launch": fam[dom]["algorithmic_bytes_per_ray_bounce"] * fam[dom]["rays_per_launch"],
                "peak_source": peak_src, "per_kernel": dict(fam, gen=gen),
                "share_of_step": dict({k: f["total_ms"] / max(1e-9, all_ms) for k, f in fam.items()},
                                      gen=gen_ms / max(1e-9, all_ms))}

    def hits_traced(c0, c1, wl_obj):
        per_layer = (c1.bounce_launches - c0.bounce_launches + c1.optics_launches - c0.optics_launches +
                     c1.genbounce_launches - c0.genbounce_launches)
        return wl_obj.max_hits if wl_obj.layer_cnt != 1 else max(1, min(wl_obj.max_hits, per_layer))

    c0, c1 = profile_pass(wk, rays_per_wl)
    exits_per_root = EXITS_PER_ROOT if args.workload == "config2" else exits_per_root_of(wk)
    roof = roofline_of(c0, c1, exits_per_root, hits_traced(c0, c1, wk))

    # the split pipeline's stand-alone intersect / optics kernels (the kernels BASELINE's north star names), same scene
    split_roof = None
    if not args.no_extras and args.workload == "config2":
        be.SetOption("fused_bounce", 0)
        driver.trace_session(be, wk.layer_cnt, B.SessionSpec(seed=42, wl=[wk.wl_entries[0]], ray_num=1 << 20, accumulate=True,
                                                             ray_base=1 << 42), 1 << 20)   # first launches load the kernels
        s0, s1 = profile_pass(wk, rays_per_wl)
        be.SetOption("fused_bounce", 1)
        sr = roofline_of(s0, s1, exits_per_root, wk.max_hits)
        split_roof = {k: {"avg_ms": v["avg_ms"], "achieved_gbs": v.get("achieved_gbs"), "frac": v.get("frac")}
                      for k, v in sr["per_kernel"].items() if k in ("optics", "intersect")}

    # ---- strong scaling: the fixed 9 x 50 M frame split over the ranks (N > 1) ----
    strong = None
    if world > 1 and not args.no_extras and args.workload == "config2":
        frame = WORKLOADS["config2"]["rays_per_wl"]
        wk.trace_step(frame, False, shard=(rank, world))
        drain()
        s_ms, _, _ = timed(lambda: wk.trace_step(frame, False, shard=(rank, world)), args.steps)
        drain()
        wk.trace_step(frame, True, shard=(rank, world))
        se_ms, se_wall, _ = timed(lambda: wk.trace_step(frame, True, shard=(rank, world)), args.steps)
        se_ms = max(se_ms, se_wall)
        drain()
        tot = frame * len(WAVELENGTHS) * args.steps
        strong = {"scaling": "strong", "frame_rays": frame * len(WAVELENGTHS), "value": tot / (s_ms * 1e-3) / 1e6,
                  "ms_per_step": s_ms / args.steps, "e2e": tot / (se_ms * 1e-3) / 1e6, "e2e_ms_per_step": se_ms / args.steps,
                  "unit": "Mrays/s"}

    # ---- the other BASELINE configs at their per-GPU ray counts (sub-records of this line) ----
    configs = {}
    if not args.no_extras and args.workload == "config2":
        for name in ("config3", "config4", "config5"):
            try:
                sub = Workload(be, name, rank, world, np, A, B, driver, 0, pinned)
                per_wl = max(1, sub.wk["rays_per_wl"] // sub.wk["gpus"])
                sub.trace_step(min(per_wl, 4 * sub.session_rays), False)      # warm-up pass (kernel variants, pools)
                drain(sub)
                v_ms, _, v_launch = timed(lambda: sub.trace_step(per_wl, False), 1)
                s_img, s_landed = drain(sub)
                ee_ms, ee_wall, _ = timed(lambda: sub.trace_step(per_wl, True), 1)
                ee_ms = max(ee_ms, ee_wall)
                drain(sub)
                p0, p1 = profile_pass(sub, per_wl)
                epr = exits_per_root_of(sub)
                rf = roofline_of(p0, p1, epr, hits_traced(p0, p1, sub))
                tot = per_wl * len(WAVELENGTHS) * world
                configs[name] = {
                    "workload": sub.wk["what"], "gpus_named": sub.wk["gpus"], "rays_per_wl_per_gpu": per_wl,
                    "value": tot / (v_ms * 1e-3) / 1e6, "ms_per_step": v_ms, "e2e": tot / (ee_ms * 1e-3) / 1e6,
                    "unit": "Mrays/s (layer-0 roots)", "exits_per_root": epr, "gpu_launches": int(v_launch),
                    "landed_per_root": s_landed / tot,
                    "per_kernel": {k: {kk: vv for kk, vv in v.items() if kk in ("avg_ms", "frac", "achieved_gbs", "launches")}
                                   for k, v in rf["per_kernel"].items()}}
            except Exception as ex:  # a sub-record must never take the headline down
                configs[name] = {"error": str(ex)[:300]}
        wk.activate()

    if rank == 0:
        total_rays = rays_per_step * world * args.steps
        value = total_rays / (dev_ms * 1e-3) / 1e6
        e2e_val = total_rays / (e2e_ms * 1e-3) / 1e6
        cb = None
        gpu_ref = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_baseline(15.0)
            except Exception as ex:  # the checker is optional at bench time
                cb = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        if world == 1 and not args.no_extras and args.workload == "config2":
            be.Synchronize()
            gpu_ref = {
                "what": "unmodified reference Simulator::Run (1 thread, TraceBackend route, third-clock drain) on the "
                        "config-2 scene, same GPU, child process; reference_cuda = the reference's CudaTraceBackend "
                        "(cuda_trace_backend.cu, -arch=sm_100a) at its default 262144-ray dispatch; b200_adapter = the "
                        "same driver on this engine through adapter/b200_trace_backend.hpp (oracle/shim), at the "
                        "reference's dispatch and at 16 Mi-ray dispatches (the reference's LUMICE_DISPATCH_RAY_NUM knob); "
                        "every run creates its backend inside the timed Simulator::Run, as the reference does; the reference "
                        "CUDA backend keeps the wavelength pool of its first session (450 nm) for the whole run, hence its "
                        "different image_sum_per_root",
                "reference_cuda": reference_driver_run("libhalo_refcuda.so", 256 * 262144, 262144),
                "b200_adapter_dispatch_262144": reference_driver_run("libhalo_refb200.so", 256 * 262144, 262144),
                "b200_adapter_dispatch_16Mi": reference_driver_run("libhalo_refb200.so", 8 * SESSION_RAYS, SESSION_RAYS)}
        line = {
            "metric": "Mrays/sec (9λ×50M single-scatter)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wk.wk['what']}; 9 wavelengths x {rays_per_wl} root rays per GPU per step",
                       "name": args.workload, "exits_per_root": exits_per_root,
                       "rays_per_step_per_gpu": rays_per_step, "session_rays": wk.session_rays,
                       "l2": "ray state per step (21.6 GB) >> 126 MB L2; no reuse across steps",
                       "parallelism": f"ray-index sharding x{world}, fp32 NCCL reduce of the image to rank 0 at frame end"},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": wk.scene_bytes +
                    len(WAVELENGTHS) * ((rays_per_wl + wk.session_rays - 1) // wk.session_rays) * C.sizeof(A.HbWlEntry),
                    "d2h_bytes_per_step": h * w * 3 * 4 + 8, "ms_per_step": e2e_ms / args.steps,
                    "host_buffer": "pinned"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
            "clocks": sampler.result() if sampler else None,
            "check": {"image_sum": img_sum, "landed_weight": landed,
                      "image_sum_per_gpu_step": img_sum / (world * args.steps),   # constant across N when the reduce is right
                      "landed_per_root": landed / (rays_per_step * world * args.steps),
                      "wall_ms_per_step": wall_ms / args.steps},
        }
        if split_roof:
            line["split_pipeline_kernels"] = split_roof
        if strong:
            line["strong"] = strong
        if configs:
            line["configs"] = configs
        if gpu_ref:
            line["gpu_reference"] = gpu_ref
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    be.close()


if __name__ == "__main__":
    main()
