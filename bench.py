#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + shadow + secondary) of the B200 render adapter on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sibenik|bunny|grid|incoherent|gyroid] [--impl reference]

A step = one frame of the hot path (camera rays -> instance routing -> two-level traversal -> shading -> shadow /
secondary rays -> framebuffer) on one synthetic scene.  rays = camera rays generated + shadow rays traced + secondary
rays traced (device counters).  Prints ONE JSON line (rank 0).

  value      device-timed throughput, scene + BVH resident in HBM, CUDA events around each frame, L2 flushed between
             frames (a buffer larger than L2 is rewritten)
  e2e        the same metric through the public frame call (gvtb200_render_frame == api::render): camera/film/light
             parameters go host->device and the RGBA float image comes back to pinned host memory inside the timed region
  roofline   dominant kernel = the k_trace launch with the largest share of the frame (for configs[1]: the closest-hit wave
             of the bounce): algorithmic bytes of THAT launch (ray stream + 128 B per wide node visited + 48 B per
             triangle tested, from per-launch counter snapshots of a counted frame) / its CUDA-event duration averaged over
             the timed frames, vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference semantics; Embree is absent) on the host cores, bounded sample

--impl reference times that CPU port alone (rank 0 only).

N > 1 (torchrun, one rank per GPU).  ImageScheduler workloads: the film rows of ONE frame are dealt to the ranks in 8-row
tiles, geometry replicated, and the rows are gathered into rank 0's framebuffer over the NVLink peer path (no collective;
--no-p2p falls back to an NCCL reduce).  gyroid (BASELINE configs[3]): Tracer<DomainScheduler>, brick k of the 2x2x2 split
lives on rank k mod N, the trace kernel stores migrating rays into the owner's queue over NVLink, one 8-byte all-reduce
per exchange round (strong scaling of one frame: the 100 M-triangle mesh is the fixed job).  ImageScheduler workloads default to
WEAK scaling (--scaling weak): the unit of work is a film tile, and with N ranks the film grows by sqrt(N) on each axis (same
scene, camera and aspect: 1920x1080 -> 2715x1527 -> 3840x2160 -> 5431x3055), so every rank keeps the N = 1 ray count and the
only cross-rank step is the row gather into rank 0.  --scaling strong deals the N = 1 film over the ranks instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from scenes import library  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: sibenik 1080p, shadow + 1 bounce (rayMaxDepth 2)
    "sibenik": dict(make=lambda w=1920, h=1080: library.sibenik_scene(w, h, depth=2), name="sibenik.obj 75,284 tris, 1920x1080, 1 point light, shadow + 1 bounce (configs[1])"),
    # north_star target: 70K bunny at 1080p, primary + shadow
    "bunny": dict(make=lambda w=1920, h=1080: library.bunny_scene(w, h), name="bun_zipper 69,451 tris, 1920x1080, 1 point light, primary+shadow"),
    # configs[2]: 32x32 instanced bunnies at 4K
    "grid": dict(make=lambda w=3840, h=2160: library.bunny_grid_scene(32, w, h), name="32x32 bunny instances (71.1M effective tris), 3840x2160, primary+shadow (configs[2])"),
    # configs[4]: incoherent stress
    "incoherent": dict(make=lambda w=3840, h=2160: library.incoherent_scene(w, h), name="sibenik.obj 3840x2160, 9 spp, rayMaxDepth 5 (4 diffuse bounces) (configs[4])"),
    # configs[3]: 100M-triangle gyroid isosurface in 8 bricks (built per rank in main(): geometry is generated on the GPU)
    "gyroid": dict(make=None, name="gyroid isosurface, 2x2x2 bricks, 3840x2160, primary+shadow (configs[3])"),
}
EXTRA_WORKLOADS = ("bunny", "grid", "incoherent")  # measured next to the default workload at N = 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: an NVML polling thread (1 ms period, so that even a 50 ms
    region is sampled), with `nvidia-smi -lms` as the fallback when NVML is not importable."""

    def __init__(self, index=0):
        import threading

        self.rows, self.stop_flag, self.thread, self.p, self.f = [], False, None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                     "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
            bits = {k: getattr(pynvml, v, None) or getattr(pynvml, v.replace("Event", "Throttle"), 0) for k, v in names.items()}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self.stop_flag:
                    try:
                        self.rows.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h))))
                    except Exception:
                        break
                    time.sleep(0.001)

            self.bits = bits
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            try:
                self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.rows:
                return None
            reasons = sorted(k for k, b in self.bits.items() if b and any(r & b for _, r in self.rows))
            return {"sm_mhz": statistics.median(c for c, _ in self.rows), "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(self.rows), "source": "nvml, 1 ms period"}
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


class cpu_frame_timer:
    """The CPU oracle rendering whole frames of a reduced film (same scene and camera), the film chosen so that one frame
    takes about `step_s` seconds on `threads` host threads; frame() returns (rays traced, seconds)."""

    def __init__(self, desc, threads, step_s):
        from oracle import pyoracle as po

        self.po, self.desc, self.threads = po, desc, threads
        d0 = desc()
        self.w, self.h = max(64, d0.width // 8), max(36, d0.height // 8)
        self._scene()
        for _ in range(3):  # the cost is not proportional to the film at small sizes (per-frame set-up): iterate
            rays, dt = self.frame()
            scale = min(8.0 * max(64, d0.width // 8) / self.w, (step_s / dt) ** 0.5)
            if scale < 1.2:
                break
            self.w, self.h = int(self.w * scale), int(self.h * scale)
            self._scene()

    def _scene(self):
        d = self.desc()
        d.width, d.height = self.w, self.h
        self.scene = library.apply(d, self.po.OracleScene(flags=d.flags))
        self.frame()  # first touch: acceleration-structure build, page faults, thread start-up

    def frame(self):
        t0 = time.perf_counter()
        _, st = self.scene.render(threads=self.threads)
        dt = time.perf_counter() - t0
        return st["camera_rays"] + st["shadow_generated"] + st["secondary_generated"], dt


def cpu_reference(desc, threads, target_s=12.0):
    """Times the CPU oracle on a bounded sample of the workload: the same scene and camera at a reduced film size chosen
    so that one frame takes a few seconds; rays counted exactly as on the GPU."""
    from oracle import pyoracle as po

    def run(w, h, budget_s=0.0):
        d = desc()
        d.width, d.height = w, h
        s = library.apply(d, po.OracleScene(flags=d.flags))
        rays_all, dts = 0, []
        while True:  # whole frames until the sample has used its budget; the rate is total rays / total time
            t0 = time.perf_counter()
            fb, st = s.render(threads=threads)
            dts.append(time.perf_counter() - t0)
            rays = st["camera_rays"] + st["shadow_generated"] + st["secondary_generated"]
            rays_all += rays
            if sum(dts) >= budget_s or len(dts) >= 64:
                break
        return rays_all, sum(dts), len(dts)

    d0 = desc()
    w, h = max(64, d0.width // 8), max(36, d0.height // 8)
    rays, dt, _ = run(w, h)
    rate = rays / dt
    scale = min(8.0, max(1.0, (target_s * rate / max(rays, 1)) ** 0.5))
    w2, h2 = int(w * scale), int(h * scale)
    if scale > 1.2:
        w, h = w2, h2
    rays, dt, frames = run(w, h, budget_s=target_s)
    return rays / dt / 1e6, f"{frames} frame(s) of the {w}x{h} film of the same scene/camera, {rays} rays in {dt:.2f} s, {threads} threads", rays, dt


def main():
    # the JSON line must be the only thing on stdout: libraries (NCCL's version banner) write to fd 1, so fd 1 is pointed at
    # stderr for the run and the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="sibenik", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tile-rows", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: NCCL reduce / all-to-all instead of the NVLink peer path")
    ap.add_argument("--owner", default="mod", choices=["mod", "balanced"], help="gyroid: brick k on rank k mod N (the reference's map) or the view-balanced placement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1, ImageScheduler workloads: weak = the film grows by sqrt(N) per axis so every rank keeps the N = 1 "
                    "ray count (same scene, camera and aspect); strong = the N = 1 film split over the ranks")
    ap.add_argument("--domain-schedule", default="rounds", choices=["async", "rounds"], help="gyroid, N > 1, peer path: free-running ranks with counter termination (scheduler::AsyncDomain) or "
                    "exchange rounds ended by an all-reduce (Tracer<DomainScheduler>); the other one is timed as well and reported in config")
    ap.add_argument("--gyroid-res", type=int, default=library.GYROID_RES_100M, help="grid cells per axis of the gyroid volume (946 = 100 M triangles)")
    a = ap.parse_args()
    a.warmup = max(3, a.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[a.workload]
    threads = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        if wl["make"] is None:  # a CPU-sized gyroid: the port builds its own BVH in seconds at this size
            wl = dict(wl, make=lambda: library.gyroid_scene(128, 2, 3840, 2160))
        # one step = one whole frame of a reduced film of the same scene and camera (the bounded sample), sized so that the
        # W + K frames take about half a minute; the rate is rays / time over the K timed frames
        step_s = min(4.0, max(0.25, 24.0 / max(1, a.steps)))
        timer = cpu_frame_timer(wl["make"], threads, step_s)
        for _ in range(a.warmup):
            timer.frame()
        rays, dts = 0, []
        for _ in range(a.steps):
            r, dt = timer.frame()
            rays += r
            dts.append(dt)
        v = rays / sum(dts) / 1e6
        sample = f"{a.steps} frame(s) of the {timer.w}x{timer.h} film of the same scene/camera, {rays} rays in {sum(dts):.2f} s, {threads} threads"
        d = wl["make"]()
        out = {
            "impl": "reference", "metric": "Mrays/s (primary+shadow+secondary)", "value": v, "unit": "Mrays/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts), "higher_is_better": True,
            "scaling": "strong" if (a.gpus > 1 and (a.scaling == "strong" or a.workload == "gyroid")) else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "film": [d.width, d.height], "depth": d.depth, "note": "CPU port of the reference path (oracle/); Embree 2.15 is absent so the Embree adapter itself cannot be built"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(out)
        return 0

    import torch
    import torch.distributed as dist

    import gravit_b200 as g
    from gravit_b200 import schedulers

    torch.cuda.set_device(local_rank)
    g.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    domain = a.workload == "gyroid"
    p2p = world > 1 and not a.no_p2p
    # ImageScheduler at N > 1: the unit of work is a film tile.  weak = N times the tiles of the N = 1 film (each axis x sqrt(N),
    # same scene / camera / aspect), every rank keeps the N = 1 ray count; strong = the N = 1 film dealt over the ranks.
    weak = world > 1 and not domain and a.scaling == "weak"
    build_s = [0.0]

    def setup(name):
        if name == "gyroid":
            # every rank generates (on its GPU) and builds only the bricks it owns: brick k -> rank k mod N (DomainTracer.h:140)
            t0 = time.perf_counter()
            owner = schedulers.domain_owner(8, world) if a.owner == "mod" else schedulers.gyroid_balanced_owner(world)
            mine = [k for k in range(8) if owner[k] == rank]
            d = library.gyroid_scene(a.gyroid_res, 2, 3840, 2160, only_bricks=mine, brick_fn=lambda *args: library.gyroid_brick_torch(*args, device=dev))
            boxes = schedulers.mesh_boxes_allgather(d, dist, world)
            e, _ = schedulers.build_domain_rank(d, rank, world, boxes=boxes, owner=owner)
            d.owner = [int(x) for x in owner]
            n_tris = torch.tensor([sum(len(m.tris) for m in d.meshes if m.verts is not None)], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(n_tris)
            d.n_tris = int(n_tris.item())
            for m in d.meshes:
                m.verts = m.tris = None  # the host copies are not needed any more
            build_s[0] = time.perf_counter() - t0
            return d, None, e
        if weak and name == a.workload:
            d0 = WORKLOADS[name]["make"]()
            f = world ** 0.5
            d = WORKLOADS[name]["make"](int(round(d0.width * f)), int(round(d0.height * f)))
        else:
            d = WORKLOADS[name]["make"]()
        s = library.apply(d, g.GvtScene(flags=d.flags))
        e = s.engine()
        return d, s, e

    # one explicit stream for the engine, the L2 flush, NCCL's entry/exit events and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    d, s, e = setup(a.workload)
    W, H = d.width, d.height
    e.set_stream(stream.cuda_stream)
    fb_ptr, _ = e.device_framebuffer()
    fb_t = schedulers.device_tensor(fb_ptr, (H, W, 4))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    if p2p:
        schedulers.p2p_connect(e, dist, rank, world, dev)
    rounds = [0]
    board = schedulers.AsyncBoard.connect(dist, rank, world) if (domain and p2p) else None  # control plane of the asynchronous schedule
    use_async = [a.domain_schedule == "async"]
    frame_no = [0]

    def step():
        if not domain:
            schedulers.render_image_schedule(e, W, H, dist if world > 1 else None, rank, world, a.tile_rows, fb_t, p2p=p2p)
        elif world == 1:
            e.render_async(H)  # one rank owns every brick: Tracer<DomainScheduler> degenerates to the local drain loop
        elif p2p and use_async[0]:
            frame_no[0] += 1
            rounds[0] = schedulers.render_domain_schedule_async(e, board, rank, world, frame_no[0])
        elif p2p:
            rounds[0] = len(schedulers.render_domain_schedule_p2p(e, dist, rank, world))
        else:
            rounds[0] = len(schedulers.render_domain_schedule(e, dist, rank, world, dev, fb_t))

    e.profile_launches(world == 1)  # an event pair around every kernel launch of the frame, on the launching stream
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    ctr = e.counters()
    # DomainScheduler: every rank generates all camera rays and keeps the ones that enter its bricks first (DomainTracer.h:148-181)
    rays_local = ctr["shadow_generated"] + ctr["secondary_generated"] + (0 if (domain and world > 1) else ctr["camera_rays"])
    ctr_keys = ("camera_rays", "routed_initial", "primary_hits", "shadow_generated", "shadow_unoccluded", "secondary_generated", "fb_adds", "rays_exported", "queue_overflows", "stack_overflows")
    rays_t = torch.tensor([rays_local] + [ctr[k] for k in ctr_keys], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_frame = int(rays_t[0].item()) + (W * H * d.samples * d.samples if (domain and world > 1) else 0)
    ctr_all = {k: int(v) for k, v in zip(ctr_keys, rays_t[1:].tolist())}
    if domain and world > 1:
        ctr_all["camera_rays"] = W * H * d.samples * d.samples

    launches0 = g.lib().gvtb200_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    trace_ms = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the per-step events)
        ev[i][0].record()
        step()
        ev[i][1].record()
        if world == 1:
            torch.cuda.synchronize()
            trace_ms.append([(k, ms) for k, ms, _ in e.launch_records()])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = g.lib().gvtb200_launch_count() - launches0
    ms_local = sum(x.elapsed_time(y) for x, y in ev)
    ms_t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    total_ms = float(ms_t.item())
    ms_per_step = total_ms / a.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6
    polls_main = rounds[0]

    other_sched = None
    if domain and p2p:  # the other DomainScheduler schedule on the same scene, same protocol (reported beside the headline)
        use_async[0] = not use_async[0]
        for _ in range(2):
            step()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        dist.barrier()
        torch.cuda.synchronize()
        for i in range(a.steps):
            flush.fill_(i & 0xFF)
            ev2[i][0].record()
            step()
            ev2[i][1].record()
        dist.barrier()
        torch.cuda.synchronize()
        o_t = torch.tensor([sum(x.elapsed_time(y) for x, y in ev2)], dtype=torch.float64, device=dev)
        dist.all_reduce(o_t, op=dist.ReduceOp.MAX)
        other_sched = {"schedule": "async" if use_async[0] else "rounds", "ms_per_step": float(o_t.item()) / a.steps, "rounds_or_polls": rounds[0]}
        use_async[0] = not use_async[0]

    # ---- end to end through the public frame call, host buffers (N = 1: gvtb200_render_frame; N > 1: + gather) ----
    host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    e2e_ms = []
    for i in range(a.warmup + a.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world == 1:
            e.render_into(host.data_ptr(), H)  # parameters H2D, frame, image D2H, synchronous
        else:
            step()
            if rank == 0:
                host.copy_(fb_t, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= a.warmup:
            e2e_ms.append(dt)
    e2e_t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = rays_frame / (float(e2e_t.item()) / a.steps * 1e-3) / 1e6

    out = None
    if rank == 0:
        pk, pk_kind = peaks()
        roof = None
        cpu = None
        extra = {}
        if world == 1:
            # counted frame: per-launch snapshots of the device counters -> algorithmic bytes of every launch
            e.profile_counters(True)
            e.render_async(H)
            e.sync()
            recs = e.launch_records()
            e.profile_counters(False)
            names = {0: "k_camera_route", 1: "k_trace (closest hit)", 2: "k_shade", 3: "k_trace (any hit, shadow rays)"}
            zero = {k: 0 for k in g.COUNTER_FIELDS}
            launches_out = []
            for i, (kind, _, after) in enumerate(recs):
                before = recs[i - 1][2] if i else zero
                dlt = {k: after[k] - before[k] for k in after}
                ms_i = statistics.mean(t[i][1] for t in trace_ms)
                # ray stream: 68 B read per queued ray (generated camera rays: 0), 84 B per hit record written + read by k_shade,
                # 68 B per ray emitted, 16 B per framebuffer add; scene: 128 B per wide node, 48 B per triangle record
                n_in = dlt["instance_visits"] if kind in (1, 3) else dlt["primary_hits"]
                if kind == 1:
                    ray_b = (0 if (i == 0) else 68) * n_in + 84 * (recs[i + 1][2]["primary_hits"] - after["primary_hits"] if i + 1 < len(recs) else 0)
                elif kind == 3:
                    ray_b = 68 * n_in + 16 * dlt["fb_adds"]
                elif kind == 2:
                    ray_b = 84 * dlt["primary_hits"] + 68 * (dlt["shadow_generated"] + dlt["secondary_generated"])
                else:
                    ray_b = 68 * dlt["routed_initial"]
                scene_b = dlt["nodes_visited"] * 128 + dlt["tris_tested"] * 48
                launches_out.append({"kernel": names[kind], "ms": ms_i, "share": ms_i / ms_per_step, "rays": n_in, "nodes_visited": dlt["nodes_visited"],
                                     "tris_tested": dlt["tris_tested"], "algorithmic_bytes": ray_b + scene_b, "GBps": (ray_b + scene_b) / (ms_i * 1e-3) / 1e9})
            dom = max(range(len(launches_out)), key=lambda i: launches_out[i]["ms"])
            dl = launches_out[dom]
            traffic = None
            tp = os.path.join(ROOT, "profiles", "r1_ncu_dominant_kernel.json")
            if os.path.exists(tp):
                try:
                    tj = json.load(open(tp))
                    if tj.get("workload") == a.workload:
                        traffic = tj.get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            roof = {
                "bound": "hbm", "achieved": dl["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dl["GBps"] / pk["hbm_gbs"], "traffic": traffic,
                "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "kernel": f"{dl['kernel']}, launch {dom} of the frame ({100 * dl['share']:.0f}% of the step)",
                "algorithmic_bytes_per_launch": dl["algorithmic_bytes"],
                "per_ray": {"nodes_visited": dl["nodes_visited"] / max(1, dl["rays"]), "tris_tested": dl["tris_tested"] / max(1, dl["rays"])},
                "launches": launches_out,
                "note": "nodes + triangles are L1/L2 resident at this scene size: the kernel is bound by instruction issue, L1 wavefronts and warp divergence (profiles/), "
                        "not by HBM; traffic = dram bytes of the same launch from the committed ncu capture",
            }
            if not a.no_cpu:
                cpu_make = WORKLOADS[a.workload]["make"] or (lambda: library.gyroid_scene(128, 2, 3840, 2160))
                v, sample, _, _ = cpu_reference(cpu_make, threads)
                if domain:
                    sample += "; gyroid at res 128 (1.8 M triangles): the CPU port cannot build the 100 M-triangle BVH inside a bench run"
                cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample}
            if not a.no_extra:
                for name in EXTRA_WORKLOADS:
                    if name == a.workload:
                        continue
                    try:
                        d2, s2, e2 = setup(name)
                        for _ in range(3):
                            e2.render_async(d2.height)
                        e2.sync()
                        c2 = e2.counters()
                        r2 = c2["camera_rays"] + c2["shadow_generated"] + c2["secondary_generated"]
                        ms = []
                        for i in range(10):
                            flush.fill_(i)
                            torch.cuda.synchronize()
                            e2.render_async(d2.height)
                            e2.sync()
                            ms.append(e2.last_timing()[0])
                        m = statistics.median(ms)
                        extra[name] = {"workload": WORKLOADS[name]["name"], "Mrays_per_s": r2 / (m * 1e-3) / 1e6, "ms_per_frame": m, "rays_per_frame": r2}
                    except Exception as ex:  # noqa: BLE001
                        extra[name] = {"error": str(ex)}
        if domain:
            wl = dict(wl, name=f"gyroid isosurface {d.n_tris:,} tris in 2x2x2 bricks (res {a.gyroid_res}), 3840x2160, 1 point light, primary+shadow (configs[3])")
            if world == 1:
                sched_desc = "DomainScheduler, 1 rank owns all bricks"
            elif not p2p:
                sched_desc = f"DomainScheduler, {world} ranks, brick -> rank {d.owner}, NCCL all-to-all of 80-byte ray records, {polls_main} rounds, NCCL reduce"
            elif a.domain_schedule == "async":
                sched_desc = (f"asynchronous DomainScheduler (scheduler::AsyncDomain), {world} ranks, brick -> rank {d.owner}, rays stored into the owner's inbound region over NVLink "
                              f"(CUDA IPC) by the trace kernel, no exchange rounds: free-running ranks, sent/received counters on a shared-memory board for termination ({polls_main} polls on rank 0), "
                              "framebuffer accumulated into rank 0 over NVLink")
            else:
                sched_desc = (f"DomainScheduler, {world} ranks, brick -> rank {d.owner}, rays stored into the owner's queue over NVLink (CUDA IPC) by the trace kernel, "
                              f"8-byte all-reduce per round, {polls_main} rounds, framebuffer accumulated into rank 0 over NVLink")
        else:
            sched_desc = "ImageScheduler" + (f", {world} ranks, row tiles of {a.tile_rows} dealt round-robin, " + ("rows stored into rank 0's framebuffer over NVLink (CUDA IPC), no collective" if p2p else "NCCL reduce to rank 0") if world > 1 else "")
        out = {
            "metric": "Mrays/s (primary+shadow+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": wl["name"] + (f"; weak scaling: film {W}x{H} = {world} x the N=1 film's pixels, same camera and aspect" if weak else ""), "film": [W, H], "depth": d.depth, "samples": d.samples, "rays_per_frame": rays_frame, "bvh": "binned SAH, 8-wide, 128 B nodes",
                "scheduler": sched_desc,
                "l2": "256 MiB buffer rewritten between timed frames (L2 flush)", "timing": "CUDA events per frame on the launching stream, max over ranks",
            },
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 496, "d2h_bytes_per_step": W * H * 16,
                    "call": "gvtb200_render_frame (== api::render) into pinned host memory" if world == 1 else "scheduler step + framebuffer to pinned host memory on rank 0"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
            "counters": ctr_all,
        }
        if domain:
            out["config"]["scene_build_s"] = round(build_s[0], 2)
        if other_sched:
            out["config"]["other_schedule"] = other_sched
        if roof:
            out["roofline"] = roof
        if cpu:
            out["cpu_baseline"] = cpu
        if extra:
            out["other_workloads"] = extra
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
