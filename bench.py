#!/usr/bin/env python
"""bench.py — Mrays/s of the trace hot path on B200 (BASELINE.json metric), contract in the task brief.

One "step" = one full pass of the workload per GPU: config 2 of BASELINE.json
(examples/config_example.json as shipped: prism h=1.3, zenith gauss(90, 0.3), max_hits 7, sun alt 20,
9 wavelengths x 50 M root rays, render id 4 = fisheye_equal_area 1920x1080) = 450 M root rays.
Rays are generated on the device by the engine's counter-based RNG (data: synthetic); weak scaling:
every rank traces its own 450 M rays (disjoint global ray-index ranges) and the per-GPU XYZ images are
all-reduced (NCCL) at frame end.

  value : root rays / s with the scene tables resident on the device (kernels only + frame-end all-reduce)
  e2e   : same, through the TraceBackend-shaped public API with HOST buffers: per step the scene /
          wavelength tables are uploaded (hb_set_scene, hb_begin_session) and the XYZ image is read back
          into host memory (hb_readback_xyz) inside the timed region
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

WAVELENGTHS = [450.0, 490.0, 530.0, 570.0, 610.0, 650.0, 690.0, 730.0, 770.0]
RAYS_PER_WL = 50_000_000
MAX_HITS = 7
SESSION_RAYS = 1 << 24          # rays per BeginSession/EndSession bracket (the driver's SimBatch size)
# algorithmic HBM bytes per ray-bounce of each kernel (DESIGN.md "kernels"): state is three float4 per ray
BYTES_OPTICS = 64               # read P,D,Q (48) + write D (16); + 16 per emitted exit (one v4 reduction)
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
TRAFFIC_FILE = "traffic_r2.json"
EXITS_PER_ROOT = 4.7            # measured on this scene (reference CPU: 4.69-4.71)


# BASELINE.json configs -> parity case, ray budget per wavelength per GPU per step, session size, description.
# config2 is the default and the only one the driver's bench line is quoted on; the others are measured on
# request (`--workload`) at reduced ray counts per step (same scene, same per-ray work) for profiles/.
WORKLOADS = {
    "config2": dict(case="column_config2", rays_per_wl=RAYS_PER_WL, session=SESSION_RAYS,
                    what="BASELINE configs[1]: config_example.json as shipped, prism h=1.3 zenith gauss(90,0.3) "
                         "max_hits 7, fisheye_equal_area 1920x1080"),
    "config3": dict(case="plate_filter_config3", rays_per_wl=50_000_000, session=SESSION_RAYS,
                    what="BASELINE configs[2]: plate h=0.3 zenith gauss(0,0.8), raypath filter [3,5] symmetry P, "
                         "max_hits 7, fisheye_equal_area 1920x1080"),
    "config4": dict(case="two_layer_config4", rays_per_wl=16_000_000, session=1 << 22,
                    what="BASELINE configs[3]: two layers, plate (prob 1.0) over full-sphere column, max_hits 7, "
                         "fisheye_equal_area 1920x1080; every exit of layer 0 re-enters layer 1"),
    "config5": dict(case="stoch_config5", rays_per_wl=50_000_000, session=SESSION_RAYS,
                    what="BASELINE configs[4]: bench_config_stoch.json prism h=1 d_i~gauss(1,0.15) full-sphere axis, "
                         "max_hits 8, rectangular 2048x1024 full sky; 256-shape pool redrawn on the device every session, "
                         "one shape per 32 consecutive rays"),
}


def workload_desc(name="config2"):
    import parity
    case = parity.CASES[WORKLOADS[name]["case"]]
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
    """Reference legacy CPU path on a bounded sample of the workload; returns the cpu_baseline object."""
    import harness as H
    A = H.A
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
    import parity
    from ice_halo_sim_b200 import backend as B
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
    from ice_halo_sim_b200.driver import stochastic_populations, trace_session

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wk = WORKLOADS[args.workload]
    desc, rdesc = workload_desc(args.workload)
    if args.geom_pool:
        desc.geom_pool_size = args.geom_pool
    elif args.workload == "config5":
        desc.geom_pool_size = 256   # the adapter's pool size; redrawn on the device at every session (below)
    stochastic = stochastic_populations(desc) if desc.geom_pool_size > 1 else []
    max_hits = int(desc.max_hits)
    layer_cnt = int(desc.layer_cnt)
    session_rays = wk["session"]
    tables = B.SceneTables(desc, 7)
    wl_entries = [B.make_wl_entry(x, 1.0) for x in WAVELENGTHS]
    be = B.B200TraceBackend(local_rank)
    if args.tile_rays:
        be.SetOption("tile_rays", args.tile_rays)
    for kv in args.set:
        k, v = kv.split("=")
        be.SetOption(k, int(v))
    be.SetScene(tables)
    be.SetRender(rdesc)

    def start_geometry_clock():   # stochastic geometry: a fresh shape pool per session, built on the device one session ahead
        for k, (li, pi) in enumerate(stochastic):
            be.AutoResample(li, pi, desc.layers[li].populations[pi].crystal, 7, ((rank * 64 + k) << 22) & 0xFFFFFFFF)
    start_geometry_clock()
    stream = torch.cuda.ExternalStream(be._lib.hb_stream(be._h), device=torch.device("cuda", local_rank))
    if world > 1:
        ids = [B.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        be.CommInit(ids[0], rank, world)

    rays_per_wl = args.rays_per_wl or wk["rays_per_wl"]
    rays_per_step = rays_per_wl * len(WAVELENGTHS)
    h, w = rdesc.img_h, rdesc.img_w
    host_img = np.empty((h, w, 3), np.float32)
    scene_bytes = C.sizeof(A.HbCrystalTables) + C.sizeof(A.HbAxisSampler) + C.sizeof(A.HbFilterDesc) + \
        C.sizeof(A.HbProjParams)

    def trace_step(step_idx, e2e):
        """One full pass: 9 wavelengths x rays_per_wl roots in SESSION_RAYS-sized sessions."""
        if e2e:  # host tables travel every step
            be.SetScene(tables)
            be.SetRender(rdesc)
            start_geometry_clock()
        for wi, wl in enumerate(wl_entries):
            done = 0
            while done < rays_per_wl:
                n = min(session_rays, rays_per_wl - done)
                base = ((step_idx * world + rank) * len(WAVELENGTHS) + wi) * rays_per_wl + done
                trace_session(be, layer_cnt, B.SessionSpec(seed=42, wl=[wl], ray_num=n, accumulate=True,
                                                           ray_base=base), n)
                done += n
        if world > 1:
            be.AllReduceImage()
        if e2e:
            be.ReadbackXyzAccum(host_img)

    def barrier():
        be.Synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, first_idx):
        barrier()
        c0 = be.Counters().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for s in range(steps):
            fn(first_idx + s)
        e1.record(stream)
        barrier()
        wall_ms = (time.time() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        ms = max(dev_ms, 0.0) if dev_ms > 0 else wall_ms
        t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), be.Counters().kernel_launches - c0

    step_ctr = 0
    for _ in range(args.warmup):
        trace_step(step_ctr, False)
        step_ctr += 1
    be.ReadbackXyzAccum(host_img)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_ms, wall_ms, launches = timed(lambda i: trace_step(i, False), args.steps, step_ctr)
    step_ctr += args.steps
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    img, landed = be.ReadbackXyzAccum(host_img)
    img_sum = float(img.astype(np.float64).sum())

    # e2e through the public API with host buffers (tables up, image down, every step)
    trace_step(step_ctr, True)
    step_ctr += 1
    e2e_ms, e2e_wall_ms, _ = timed(lambda i: trace_step(i, True), args.steps, step_ctr)
    step_ctr += args.steps
    e2e_ms = max(e2e_ms, e2e_wall_ms)  # the D2H read blocks the host: wall clock bounds the step

    # per-kernel live timing for the roofline (CUDA events around every launch, short pass)
    be.SetOption("profile", 1)
    c0 = be.Counters()
    p_rays = min(rays_per_wl, session_rays)
    trace_session(be, layer_cnt, B.SessionSpec(seed=42, wl=[wl_entries[0]], ray_num=p_rays, accumulate=True,
                                               ray_base=1 << 40), p_rays)
    be.Synchronize()
    c1 = be.Counters()
    be.SetOption("profile", 0)
    be.ReadbackXyzAccum(host_img)
    exits_per_root = EXITS_PER_ROOT
    if args.workload != "config2":  # exits per layer-0 root of this scene, counted by the engine (LayerStats)
        be.BeginSession(B.SessionSpec(seed=42, wl=[wl_entries[0]], ray_num=1 << 20, accumulate=True, ray_base=1 << 41))
        hdl = be.TraceLayer(B.RootRaySource.FromHost(1 << 20), want_stats=True)
        be.EndSession()
        exits_per_root = hdl.exit_count / float(1 << 20)  # filter-passing exits (continuing ones included)
        be.ReadbackXyzAccum(host_img)

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_src = "fallback (B200_PROFILING.md)"
        hbm_peak = 6650.0
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
            hbm_peak = float(peaks.get("hbm_gbs", hbm_peak))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        epb = exits_per_root / max_hits          # exits per ray-bounce
        fam = {}
        for name, ms, launches, rays, bytes_per_ray in (
                ("bounce", c1.bounce_ms - c0.bounce_ms, c1.bounce_launches - c0.bounce_launches,
                 c1.bounce_rays - c0.bounce_rays,
                 # with the entry interaction inside genbounce the bounce kernels run hits 1 .. H-1
                 ((BYTES_BOUNCE * (max_hits - 2) + BYTES_BOUNCE_LAST) / (max_hits - 1) +
                  BYTES_BOUNCE_EXIT * (exits_per_root - EXITS_AT_ENTRY) / (max_hits - 1))
                 if c1.genbounce_launches > c0.genbounce_launches and max_hits > 1 else
                 (BYTES_BOUNCE * (max_hits - 1) + BYTES_BOUNCE_LAST) / max_hits + BYTES_BOUNCE_EXIT * epb),
                ("genbounce", c1.genbounce_ms - c0.genbounce_ms, c1.genbounce_launches - c0.genbounce_launches,
                 c1.genbounce_rays - c0.genbounce_rays, BYTES_GENBOUNCE + BYTES_GENBOUNCE_EXIT * EXITS_AT_ENTRY),
                ("optics", c1.optics_ms - c0.optics_ms, c1.optics_launches - c0.optics_launches,
                 c1.optics_rays - c0.optics_rays,
                 (BYTES_OPTICS * (max_hits - 1) + BYTES_OPTICS_LAST) / max_hits + BYTES_EXIT * epb),
                ("intersect", c1.intersect_ms - c0.intersect_ms, c1.intersect_launches - c0.intersect_launches,
                 c1.intersect_rays - c0.intersect_rays, BYTES_INTERSECT)):
            if launches == 0:
                continue
            avg_ms = ms / launches
            tile = rays / launches
            ach = tile * bytes_per_ray / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
            fam[name] = {"avg_ms": avg_ms, "launches": int(launches), "rays_per_launch": tile,
                         "algorithmic_bytes_per_ray_bounce": bytes_per_ray, "achieved_gbs": ach,
                         "frac": ach / hbm_peak, "total_ms": ms}
        gen_ms = c1.gen_ms - c0.gen_ms
        gen_l = c1.gen_launches - c0.gen_launches
        all_ms = sum(f["total_ms"] for f in fam.values()) + gen_ms
        dom = max(fam, key=lambda k: fam[k]["total_ms"])
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
        if os.path.exists(tr_path):  # DRAM bytes per ray-bounce from the committed ncu --set full capture
            tr = json.load(open(tr_path))["dram_bytes_per_ray_bounce"]
            if dom in tr:
                traffic = tr[dom] * fam[dom]["rays_per_launch"]
        roof = {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": fam[dom]["achieved_gbs"],
                "peak": hbm_peak, "unit": "GB/s", "frac": fam[dom]["frac"], "traffic": traffic,
                "algorithmic_bytes_per_launch": fam[dom]["algorithmic_bytes_per_ray_bounce"] * fam[dom]["rays_per_launch"],
                "peak_source": peak_src, "per_kernel": dict(fam, gen={"avg_ms": gen_ms / max(1, gen_l), "launches": int(gen_l)}),
                "share_of_step": {k: f["total_ms"] / max(1e-9, all_ms) for k, f in fam.items()}}
        total_rays = rays_per_step * world * args.steps
        value = total_rays / (dev_ms * 1e-3) / 1e6
        e2e_val = total_rays / (e2e_ms * 1e-3) / 1e6
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_baseline(15.0)
            except Exception as ex:  # the checker is optional at bench time
                cb = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        line = {
            "metric": "Mrays/sec (9λ×50M single-scatter)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wk['what']}; 9 wavelengths x {rays_per_wl} root rays per GPU per step",
                       "name": args.workload, "exits_per_root": exits_per_root,
                       "rays_per_step_per_gpu": rays_per_step, "session_rays": session_rays,
                       "l2": "ray state per step (21.6 GB) >> 126 MB L2; no reuse across steps",
                       "parallelism": f"ray-index sharding x{world}, NCCL image all-reduce at frame end"},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": scene_bytes +
                    len(WAVELENGTHS) * ((rays_per_wl + session_rays - 1) // session_rays) * C.sizeof(A.HbWlEntry),
                    "d2h_bytes_per_step": h * w * 3 * 4 + 8, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
            "clocks": sampler.result() if sampler else None,
            "check": {"image_sum": img_sum, "landed_weight": landed, "wall_ms_per_step": wall_ms / args.steps},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    be.close()


if __name__ == "__main__":
    main()
