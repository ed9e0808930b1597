#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + shadow + secondary) of the B200 render adapter on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sibenik|bunny|grid|incoherent|gyroid] [--impl reference]

A step = one frame of the hot path (camera rays -> instance routing -> two-level traversal -> shading -> shadow /
secondary rays -> framebuffer) on one synthetic scene.  rays = camera rays generated + shadow rays traced + secondary
rays traced (device counters).  Prints ONE JSON line (rank 0).

  value      device-timed throughput, scene + BVH resident in HBM, CUDA events around each frame, L2 flushed between
             frames (a buffer larger than L2 is rewritten), max over ranks
  e2e        the same metric through the public frame calls with HOST buffers: N = 1 gvtb200_render_frame (== api::render:
             camera/film/light parameters host->device, the RGBA float image back to pinned host memory); N > 1 every rank
             renders its rows and brings them to ONE shared pinned host image over its own PCIe link
             (gvtb200_framebuffer_read_rows), so the image is never funnelled through one GPU
  roofline   SURVEY 8(d): dominant kernel = the launch with the largest share of the frame; bound = the larger of the issue
             term (smsp__inst_executed of that launch / (SMs x 4 schedulers x SM clock x its duration)) and the HBM term
             (dram bytes of that launch / its duration / measured HBM peak).  Instruction and DRAM counts come from an ncu
             capture of the SAME kernel sources (profiles/r2_ncu_metrics_*.json carries the hash of gravit_b200/csrc; a
             capture of other sources is refused), durations and the SM clock are measured live in this run
  cpu_baseline  the CPU oracle (a port of the reference semantics; Embree is absent) on the host cores, bounded sample
  parity     N > 1: max |difference| between the N-GPU image and the same film rendered by rank 0 alone
  other_workloads  at EVERY N: BASELINE configs[2] (1,024-instance grid, ImageScheduler, 4K, strong scaling) and configs[3]
             (100 M-triangle gyroid, DomainScheduler, 4K, strong scaling: persistent scheduler with and without shared
             BLAS, round-synchronous schedule, per-rank busy times); N = 1 also bunny 1080p and configs[4]

--impl reference times that CPU port alone (rank 0 only).

N > 1 (torchrun, one rank per GPU).  Headline: sibenik ImageScheduler, WEAK scaling (--scaling weak): the unit of work is a
film tile, and with N ranks the film grows by sqrt(N) on each axis (same scene, camera and aspect: 1920x1080 -> 2715x1527 ->
3840x2160 -> 5431x3055), so every rank keeps the N = 1 ray count; rows are dealt in 8-row tiles and gathered into rank 0's
framebuffer over the NVLink peer path (no collective; --no-p2p falls back to an NCCL reduce).  --scaling strong deals the
N = 1 film over the ranks instead.  gyroid (configs[3]): brick k of the 2x2x2 split is OWNED by rank k mod N
(DomainTracer.h:140); the persistent scheduler lets idle ranks serve copies of the busiest bricks (a measured load plan).
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
EXTRA_N1 = ("bunny", "incoherent")  # measured next to the default workload at N = 1 only
EXTRA_ALL_N = ("grid", "gyroid")   # BASELINE configs[2] and configs[3]: measured at every N (strong scaling)


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


def ncu_capture(workload, mode="waves"):
    """The committed ncu counters of `workload`, or (None, why) when they were taken from other kernel sources."""
    import gravit_b200 as g

    path = os.path.join(ROOT, "profiles", f"r2_ncu_metrics_{workload}_{mode}.json")
    if not os.path.exists(path):
        return None, f"no capture ({os.path.relpath(path, ROOT)} is missing)"
    j = json.load(open(path))
    if j.get("source_hash") != g.source_hash():
        return None, f"stale capture: taken from kernel sources {j.get('source_hash')}, this run uses {g.source_hash()}"
    return j, os.path.relpath(path, ROOT)


def shared_host_image(dist, rank, world, H, W):
    """One RGBA float image in host memory that every rank can write: a file under /dev/shm mapped by all ranks and pinned
    (cudaHostRegister) by each of them.  Returns (torch tensor view, numpy array, cleanup)."""
    import torch

    path = [None]
    if rank == 0:
        root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        fd, path[0] = tempfile.mkstemp(prefix="gvtb200_img_", dir=root)
        os.ftruncate(fd, H * W * 16)
        os.close(fd)
    if world > 1:
        dist.broadcast_object_list(path, src=0)
    arr = np.memmap(path[0], dtype=np.float32, mode="r+", shape=(H, W, 4))
    t = torch.from_numpy(arr)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * 4, 0)
    pinned = int(rc) == 0

    def cleanup():
        if pinned:
            torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
        if world > 1:
            dist.barrier()
        if rank == 0:
            try:
                os.unlink(path[0])
            except OSError:
                pass

    return t, arr, cleanup, pinned


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
    ap.add_argument("--extra", default=None, help="comma-separated subset of the secondary workloads (bunny,incoherent,grid,gyroid)")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: NCCL reduce / all-to-all instead of the NVLink peer path")
    ap.add_argument("--owner", default="mod", choices=["mod", "balanced"], help="gyroid: brick k on rank k mod N (the reference's map) or the view-balanced placement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1, ImageScheduler headline: weak = the film grows by sqrt(N) per axis so every rank keeps the N = 1 "
                    "ray count (same scene, camera and aspect); strong = the N = 1 film split over the ranks")
    ap.add_argument("--replica-sweep", default="1,2,3,99", help="gyroid, persistent scheduler, local-first policy: replica counts per rank to time (99 = everything everywhere); '' = none")
    ap.add_argument("--gyroid-samples", type=int, default=2, help="gyroid: also time the same scene and film with samples^2 rays per pixel (the frame of a GPU at N = 8 holds "
                    "15 rays per lane at 1 spp: the latency of the longest ray chain, not throughput, is what is measured there); 1 = skip")
    ap.add_argument("--replicas", type=int, default=1, help="gyroid, persistent scheduler: BLAS copies of other ranks' bricks a rank may hold (0 = the owner alone serves a brick)")
    ap.add_argument("--shade-sms", type=int, default=0, help="persistent scheduler: SMs that shade instead of tracing (0 = default)")
    ap.add_argument("--gyroid-res", type=int, default=library.GYROID_RES_100M, help="grid cells per axis of the gyroid volume (946 = 100 M triangles)")
    a = ap.parse_args()
    a.warmup = max(3, a.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[a.workload]
    threads = os.cpu_count() or 1
    vs_note = "CPU port of the reference path (oracle/), NOT Embree: Embree 2.15 / TBB / MPI / IceT are absent, the Embree adapter cannot be built here"

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
            "config": {"workload": wl["name"], "film": [timer.w, timer.h], "film_of_the_workload": [d.width, d.height], "depth": d.depth,
                       "note": vs_note + "; the bounded sample is the same scene and camera on the reduced film given in `film`"},
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
    p2p = world > 1 and not a.no_p2p
    # one explicit stream for the engine, the L2 flush, NCCL's entry/exit events and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    ctr_keys = ("camera_rays", "routed_initial", "primary_hits", "shadow_generated", "shadow_unoccluded", "secondary_generated", "fb_adds", "rays_exported", "queue_overflows", "stack_overflows")

    def allsum(vals, op=None):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return t.tolist()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(step, steps, warmup, sync_each=None):
        """W untimed steps, then K steps bracketed by CUDA events on the launching stream; max over ranks of the total."""
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the per-step events)
            ev[i][0].record()
            step()
            ev[i][1].record()
            if sync_each:
                torch.cuda.synchronize()
                sync_each(i)
        barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = allsum([sum(x.elapsed_time(y) for x, y in ev)], dist.ReduceOp.MAX if world > 1 else None)[0]
        return ms / steps, wall

    def frame_rays(e, domain_sharded_camera=False):
        c = e.counters()
        local = c["camera_rays"] + c["shadow_generated"] + c["secondary_generated"]
        tot = allsum([local] + [c[k] for k in ctr_keys])
        return int(tot[0]), {k: int(v) for k, v in zip(ctr_keys, tot[1:])}

    # ================= ImageScheduler workloads (sibenik, bunny, grid, incoherent) =========================================
    def run_image(name, film=None, steps=None, warmup=None, headline=False, scaling="strong"):
        steps, warmup = steps or a.steps, warmup or a.warmup
        d = WORKLOADS[name]["make"](*film) if film else WORKLOADS[name]["make"]()
        s = library.apply(d, g.GvtScene(flags=d.flags))
        e = s.engine()
        W, H = d.width, d.height
        e.set_stream(stream.cuda_stream)
        fb_ptr, _ = e.device_framebuffer()
        fb_t = schedulers.device_tensor(fb_ptr, (H, W, 4))
        if p2p:
            schedulers.p2p_connect(e, dist, rank, world, dev)

        def step():
            schedulers.render_image_schedule(e, W, H, dist if world > 1 else None, rank, world, a.tile_rows, fb_t, p2p=p2p)

        res = {"workload": WORKLOADS[name]["name"], "film": [W, H], "depth": d.depth, "samples": d.samples}
        e.profile_launches(world == 1 and headline)
        for _ in range(warmup):
            step()
        e.sync()
        rays, ctr_all = frame_rays(e)
        trace_ms = []
        launches0 = g.lib().gvtb200_launch_count()
        sampler = ClockSampler(local_rank) if (rank == 0 and headline) else None
        ms, wall = timed(step, steps, 0, sync_each=(lambda i: trace_ms.append([(k, t) for k, t, _ in e.launch_records()])) if (world == 1 and headline) else None)
        res["clocks"] = sampler.stop() if sampler else None
        res["launches"] = int(g.lib().gvtb200_launch_count() - launches0)
        e.sync()  # raises if a queue overflowed or a peer timed out
        res.update(ms_per_frame=ms, rays_per_frame=rays, Mrays_per_s=rays / (ms * 1e-3) / 1e6, counters=ctr_all, wall_s=wall, trace_ms=trace_ms)
        entered = ctr_all["routed_initial"] + ctr_all["shadow_generated"] + ctr_all["secondary_generated"]
        res["Mrays_per_s_entered_an_instance"] = entered / (ms * 1e-3) / 1e6  # rays that were actually traced in a BLAS (background camera rays excluded)
        # ---- end to end with host buffers -------------------------------------------------------------------------------
        if world == 1:
            host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
            cleanup = lambda: None  # noqa: E731
        else:
            host, host_np, cleanup, pinned = shared_host_image(dist, rank, world, H, W)
            res["e2e_host_image"] = "one /dev/shm image mapped by all ranks, " + ("cudaHostRegister'ed by each" if pinned else "NOT pinned (cudaHostRegister failed)")
        e2e_ms = []
        for i in range(warmup + steps):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                e.render_into(host.data_ptr(), H)  # parameters H2D, frame, image D2H, synchronous
            else:
                e.render_tiles(H, a.tile_rows, rank, world)  # this rank's rows ...
                e.fb_read_rows(host.data_ptr(), a.tile_rows, rank, world)  # ... straight to their place in the host image
                e.sync()
            dt = (time.perf_counter() - t0) * 1e3
            if i >= warmup:
                e2e_ms.append(dt)
        e2e = allsum([sum(e2e_ms)], dist.ReduceOp.MAX if world > 1 else None)[0] / steps
        res["e2e"] = {"value": rays / (e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e, "h2d_bytes_per_step": 496 * world, "d2h_bytes_per_step": W * H * 16,
                      "call": "gvtb200_render_frame (== api::render) into pinned host memory" if world == 1 else
                      f"gvtb200_render_frame_tiles + gvtb200_framebuffer_read_rows on every rank: {world} PCIe links write one shared pinned host image"}
        # ---- parity: the N-GPU image against the same film rendered by rank 0 alone ---------------------------------------
        if world > 1:
            barrier()
            img_host = np.array(host_np) if rank == 0 else None  # what the last e2e frame left in the shared image
            step()
            e.sync()
            barrier()
            if rank == 0:
                img_gather = e.fb_read(W, H)
                e.render_async(H)
                e.sync()
                ref = e.fb_read(W, H)
                res["parity"] = {"max_abs_err_gather": float(np.abs(img_gather - ref).max()), "max_abs_err_host_image": float(np.abs(img_host - ref).max()),
                                 "reference": "the same film rendered by rank 0 alone (gvtb200_render_frame_async)", "lit_pixels": int((ref[..., 3] > 0).sum())}
            barrier()
        cleanup()
        if p2p:
            barrier()
            e.p2p_detach()
            barrier()
        return res, e, d

    # ================= BASELINE configs[3]: gyroid, Tracer<DomainScheduler> ==================================================
    def run_gyroid(steps=None, warmup=None):
        steps, warmup = steps or a.steps, warmup or a.warmup
        t0 = time.perf_counter()
        owner = schedulers.domain_owner(8, world) if a.owner == "mod" else schedulers.gyroid_balanced_owner(world)
        mine = [k for k in range(8) if owner[k] == rank]
        d = library.gyroid_scene(a.gyroid_res, 2, 3840, 2160, only_bricks=mine, brick_fn=lambda *args: library.gyroid_brick_torch(*args, device=dev))
        gen_s = time.perf_counter() - t0
        boxes = schedulers.mesh_boxes_allgather(d, dist, world)
        e, _ = schedulers.build_domain_rank(d, rank, world, boxes=boxes, owner=owner)
        n_tris = int(allsum([sum(len(m.tris) for m in d.meshes if m.verts is not None)])[0])
        blas_ms = [e._keep[k].stats()["build_us"] * 1e-3 for k in range(8) if e._keep[k] is not None]
        for m in d.meshes:
            m.verts = m.tris = None  # the host copies are not needed any more
        torch.cuda.empty_cache()
        build_s = time.perf_counter() - t0
        W, H = d.width, d.height
        e.set_stream(stream.cuda_stream)
        if world > 1:
            schedulers.p2p_connect(e, dist, rank, world, dev)
        if a.shade_sms:
            e.domain_persist_config(0, a.shade_sms)
        res = {"workload": f"gyroid isosurface {n_tris:,} tris in 2x2x2 bricks (res {a.gyroid_res}), 3840x2160, 1 point light, primary+shadow (configs[3])",
               "film": [W, H], "n_tris": n_tris, "owner": [int(x) for x in owner], "scene_build_s": round(build_s, 2), "geometry_generation_s": round(gen_s, 2),
               "blas_build_ms_per_brick_this_rank": [round(x, 1) for x in blas_ms],
               "blas_build_Mtris_per_s_this_rank": round(n_tris / 8.0 * len(mine) / max(1e-9, sum(blas_ms) * 1e-3) / 1e6, 1),
               "blas_build_note": "the first build on a rank also grows the stream-ordered memory pool (several GB of scratch) and loads the kernels; the later ones show the build itself"}
        if len(blas_ms) > 1:
            res["blas_build_Mtris_per_s_warm_median"] = round(n_tris / 8.0 / (statistics.median(blas_ms[1:]) * 1e-3) / 1e6, 1)

        def persist():
            schedulers.render_domain_persist(e, 0, a.tile_rows)

        def persist_stats():
            busy, inbox, _ = e.domain_persist_stats()
            b = torch.zeros(world, dtype=torch.float64, device=dev)
            b[rank] = busy
            r = torch.zeros(world, dtype=torch.float64, device=dev)
            r[rank] = inbox
            if world > 1:
                dist.all_reduce(b)
                dist.all_reduce(r)
            b, r = b.tolist(), r.tolist()
            return {"busy_ms_per_rank": [round(x, 3) for x in b], "imbalance_max_over_mean": round(max(b) / max(1e-9, sum(b) / len(b)), 3), "inbox_rays_per_rank": [int(x) for x in r]}

        def progress_curve():
            """The root's progress samples of the last persistent frame, thinned: [ms, fraction of the rays retired]."""
            if rank != 0:
                return None
            p = e.domain_persist_progress()
            if not len(p):
                return []
            total = float(p[-1, 2]) or 1.0
            keep = p[:: max(1, len(p) // 24)]
            return [[round(float(us) * 1e-3, 3), round(float(c) / total, 4)] for us, c, _ in keep] + [[round(float(p[-1, 0]) * 1e-3, 3), round(float(p[-1, 1]) / total, 4)]]

        spp = [1]  # rays per pixel of the camera installed now

        def measure(label, step):
            ms, _ = timed(step, steps, warmup)
            e.sync()
            _, ctr_all = frame_rays(e)
            # persistent scheduler: every rank generates its own shard of the camera rays; round-synchronous: every rank
            # generates all of them and keeps its own (DomainTracer.h:148-181) -- the film's rays are counted once either way
            assert ctr_all["camera_rays"] in (W * H * spp[0], W * H * spp[0] * world), ctr_all
            ctr_all["camera_rays"] = W * H * spp[0]
            rays = W * H * spp[0] + ctr_all["shadow_generated"] + ctr_all["secondary_generated"]
            out = {"schedule": label, "ms_per_frame": ms, "rays_per_frame": rays, "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "rays_exported": ctr_all["rays_exported"],
                   "queue_overflows": ctr_all["queue_overflows"]}
            return out, ctr_all

        def calibrate(label):
            """How many SMs shade instead of tracing is a property of the scene (hits per ray, cost of a ray): a few frames per
            candidate, the fastest stays (a scene-setup step like the server plan; --shade-sms fixes it instead)."""
            if a.shade_sms:
                return {"shade_sms": a.shade_sms, "fixed_by": "--shade-sms"}
            best, tried, split, tried_split = schedulers.calibrate_shader_sms(e, persist, lambda f: timed(f, 6, 2)[0])
            return {"shade_sms": best, "schedule": label, "ms_per_frame_by_shade_sms": {str(k): round(v, 3) for k, v in tried.items()},
                    "tail_split": split, "ms_per_frame_by_tail_split": {str(k): round(v, 3) for k, v in tried_split.items()}}

        # (1) the persistent scheduler, every brick served by its owner alone
        res["shader_sm_calibration"] = [calibrate("owner only")]
        base, ctr_all = measure("persistent kernel, brick served by its owner", persist)
        base.update(persist_stats())
        base["progress_ms_vs_fraction_retired"] = progress_curve()
        if world == 1:
            e.profile_counters(True)
            persist()
            e.sync()
            res["cycle_profile_counted_frame"] = [e.domain_persist_profile()]
            e.profile_counters(False)
        res["persistent_owner_only"] = base
        headline = base
        descs = None
        install = lambda: e.domain_set_servers(None)  # noqa: E731  (re-installs the headline schedule's server table)
        ref_img = [None]

        def parity(label):
            """The N-rank image (root's framebuffer) of the schedule installed now against ONE rank tracing every brick (the
            others' BLAS mapped over NVLink; rendered once)."""
            persist()
            e.sync()
            barrier()
            out = None
            if rank == 0:
                img = e.fb_read(W, H)
                if ref_img[0] is None:
                    full, imported_ref = schedulers.whole_scene_on_one_rank(e, descs, owner, rank, mode=g.IMPORT_MAP)
                    t1 = time.perf_counter()
                    ref_img[0], _ = full.render(W, H)
                    res["parity_reference"] = {"what": "rank 0 alone renders the whole scene through the wave path, the other ranks' bricks traversed through NVLink-mapped BLAS",
                                               "frame_s": round(time.perf_counter() - t1, 3), "lit_pixels": int((ref_img[0][..., 3] > 0).sum())}
                    full.close()
                    del imported_ref
                out = {"schedule": label, "max_abs_err": float(np.abs(img - ref_img[0]).max())}
            barrier()
            return out

        if world > 1:
            # (2) shared BLAS: a counted frame gives the load per brick, the plan spreads the hot bricks over idle ranks
            e.profile_counters(True)
            persist()
            e.sync()
            load = e.domain_persist_stats(8)[2].astype(np.float64)
            prof = [None] * world
            dist.all_gather_object(prof, e.domain_persist_profile())
            res["cycle_profile_owner_only_counted_frame"] = prof
            e.profile_counters(False)
            load = np.array(allsum(load))
            res["node_visits_per_brick"] = [int(x) for x in load]
            descs = schedulers.exchange_mesh_descs(dist, e._keep, e.recipe["inst"], owner, rank, world, dev)
            parities = [parity("owner only")]
            imported = {}
            brick_bytes = [0.0]

            def counted_profile():
                e.profile_counters(True)
                persist()
                e.sync()
                prof = [None] * world
                dist.all_gather_object(prof, e.domain_persist_profile())
                e.profile_counters(False)
                return prof

            if a.replicas > 0:
                tab_r, tab_w = schedulers.plan_servers(load, owner, world, max_replicas=a.replicas)
                t1 = time.perf_counter()
                free0 = torch.cuda.mem_get_info()[0]
                imported.update(schedulers.apply_server_plan(e, tab_r, tab_w, descs, rank, mode=g.IMPORT_COPY))
                torch.cuda.synchronize()
                if imported:
                    brick_bytes[0] = (free0 - torch.cuda.mem_get_info()[0]) / len(imported)
                barrier()
                res["shader_sm_calibration"].append(calibrate("shared BLAS, by hash"))
                shared, _ = measure(f"persistent kernel, <= {a.replicas} BLAS replica(s) of other ranks' bricks per rank (copied from the owner over NVLink), every ray of a brick goes to one of its servers by hash", persist)
                shared.update(persist_stats())
                shared["progress_ms_vs_fraction_retired"] = progress_curve()
                shared["cycle_profile_counted_frame"] = counted_profile()
                shared.update(servers=[[int(x) for x in row if x >= 0] for row in tab_r], shares=[[round(float(w), 3) for w, x in zip(wr, row) if x >= 0] for wr, row in zip(tab_w, tab_r)],
                              replica_setup_s=round(time.perf_counter() - t1, 3))
                res["persistent_shared_blas"] = shared
                parities.append(parity("shared BLAS, by hash"))
                if shared["ms_per_frame"] < headline["ms_per_frame"]:
                    headline = shared
                    install = lambda: e.domain_set_servers(tab_r, tab_w)  # noqa: E731
            # (2b) local first: a rank that holds a brick keeps the rays that meet it there; replicas per rank swept up to
            # "everything everywhere" (no ray leaves its rank: the ImageScheduler's data layout under the DomainScheduler)
            full_r = 8 - min(int((np.asarray(owner) == r).sum()) for r in range(world))
            sweep = sorted({min(int(x), full_r) for x in a.replica_sweep.split(",") if x.strip()}) if a.replica_sweep else []
            res["replication_sweep"] = []
            per_brick = allsum([brick_bytes[0]], dist.ReduceOp.MAX)[0]
            for R in sweep:
                if R <= 0:
                    continue
                holders, share = schedulers.plan_holders(load, owner, world, max_replicas=R)
                t1 = time.perf_counter()
                free0 = torch.cuda.mem_get_info()[0]
                n0 = len(imported)
                imported = schedulers.apply_holder_plan(e, holders, share, descs, rank, owner, mode=g.IMPORT_COPY, imported=imported)
                torch.cuda.synchronize()
                if len(imported) > n0 and not per_brick:
                    per_brick = (free0 - torch.cuda.mem_get_info()[0]) / (len(imported) - n0)
                barrier()
                setup_s = time.perf_counter() - t1
                rl, moved = schedulers.planned_rank_load_local(load, holders, share, world)
                m, _ = measure(f"persistent kernel, local first, <= {R} BLAS replica(s) of other ranks' bricks per rank", persist)
                m.update(persist_stats())
                held = [sum(1 for h in holders if r in h) for r in range(world)]
                m.update(replicas_per_rank=R, bricks_held_per_rank=held, everything_everywhere=all(len(h) == world for h in holders),
                         holders=[[int(x) for x in h] for h in holders], planned_max_rank_load_over_mean=round(float(rl.max() * world), 3),
                         planned_fraction_of_brick_visits_that_start_with_a_transfer=round(float(moved), 3), replica_setup_s=round(setup_s, 3))
                if R == sweep[-1] or R == a.replicas:
                    m["progress_ms_vs_fraction_retired"] = progress_curve()
                    m["cycle_profile_counted_frame"] = counted_profile()
                p_ = parity(f"local first, {R} replica(s)")
                parities.append(p_)
                m["parity_max_abs_err"] = p_["max_abs_err"] if p_ else None
                res["replication_sweep"].append(m)
                # the headline stays a PARTITIONED scene: at most --replicas copies per rank
                if R <= a.replicas and m["ms_per_frame"] < headline["ms_per_frame"]:
                    headline = m
                    install = (lambda hh, ss: (lambda: e.domain_set_servers(*schedulers.server_table_for_rank(hh, ss, rank))))(holders, share)
            per_brick = allsum([per_brick], dist.ReduceOp.MAX)[0]
            res["blas_GB_per_brick"] = round(per_brick / 1e9, 3)
            if rank == 0:
                res["parity"] = {"max_abs_err": max(p_["max_abs_err"] for p_ in parities), "per_schedule": parities, **res.get("parity_reference", {})}
                ref = ref_img[0]
        else:
            persist()
            e.sync()
            barrier()
            img = e.fb_read(W, H)
            ref, _ = e.render(W, H)  # the wave path on the same scene
            res["parity"] = {"max_abs_err": float(np.abs(img - ref).max()), "lit_pixels": int((ref[..., 3] > 0).sum()), "reference": "the wave path (k_trace / k_shade launches) on the same scene"}
        # (3) the other schedules, for comparison
        if world > 1:
            e.domain_set_servers(None)
            rounds = [0]

            def rounds_step():
                rounds[0] = len(schedulers.render_domain_schedule_p2p(e, dist, rank, world))

            rs, _ = measure("round-synchronous (Tracer<DomainScheduler>): one trace/shade/trace wave and an 8-byte all-reduce per exchange round, NVLink peer queues", rounds_step)
            rs["rounds"] = rounds[0]
            res["round_synchronous"] = rs
        else:
            e.profile_launches(True)
            wave_ms = []

            def wave_step():
                e.render_async(H)
                e.sync()
                wave_ms.append([t for _, t, _ in e.launch_records()])

            ws, _ = measure("wave path (k_trace / k_shade launches per bounce level), all bricks on one GPU", wave_step)
            e.profile_launches(False)
            res["waves"] = ws
            if ws["ms_per_frame"] < headline["ms_per_frame"]:
                headline = ws
            # roofline of the launches (scene >> L2: the one config where HBM could bind): instruction and DRAM counts from the
            # committed ncu capture of these kernel sources, durations measured here
            cap, cap_src = ncu_capture("gyroid", "waves")
            pk, _ = peaks()
            sms = g.device_info()["sm_count"]
            issue_peak = sms * 4 * pk.get("sm_max_mhz", 1965.0) * 1e6
            if cap and wave_ms and len(cap["launches"]) == len(wave_ms[-1]) and a.gyroid_res == library.GYROID_RES_100M:
                ms_l = [statistics.mean(x[i] for x in wave_ms[-steps:]) for i in range(len(wave_ms[-1]))]
                rl = []
                for ms_i, c in zip(ms_l, cap["launches"]):
                    dram = c.get("dram__bytes_read.sum", 0.0) + c.get("dram__bytes_write.sum", 0.0)
                    rl.append({"kernel": c.get("kernel"), "ms": ms_i, "warp_inst": c.get("smsp__inst_executed.sum"), "issue_frac": c.get("smsp__inst_executed.sum", 0.0) / (ms_i * 1e-3) / issue_peak,
                               "dram_bytes": dram, "hbm_frac": dram / (ms_i * 1e-3) / 1e9 / pk["hbm_gbs"], "simt_lanes_per_inst": c.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                               "l2_hit_pct": c.get("lts__t_sector_hit_rate.pct"), "l1_wavefronts_pct_of_peak": c.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")})
                dom = max(rl, key=lambda x: x["ms"])
                res["roofline"] = {"bound": "issue" if dom["issue_frac"] >= dom["hbm_frac"] else "hbm", "frac": max(dom["issue_frac"], dom["hbm_frac"]), "dominant": dom, "launches": rl,
                                   "counters_source": cap_src, "note": "even with 4.8 GB of triangle records the L2 serves most node fetches: DRAM traffic is a fraction of the algorithmic bytes and the kernel is issue / latency bound like the L2-resident scenes"}
            else:
                res["roofline"] = {"frac": None, "note": "no ncu capture of these kernel sources for this workload (" + str(cap_src) + ")"}
        # ---- end to end: frame + parallel read-back of the root's image ---------------------------------------------------
        if world > 1:
            install()
        host, host_np, cleanup, pinned = shared_host_image(dist, rank, world, H, W)
        use_waves = world == 1 and headline is res.get("waves")
        e2e_ms = []
        for i in range(warmup + steps):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            barrier()
            t1 = time.perf_counter()
            if use_waves:
                e.render_into(host.data_ptr(), H)
            else:
                persist()
                e.fb_read_rows(host.data_ptr(), a.tile_rows, rank, world, src_rank=0 if world > 1 else -1)
                e.sync()
            dt = (time.perf_counter() - t1) * 1e3
            if i >= warmup:
                e2e_ms.append(dt)
        e2e = allsum([sum(e2e_ms)], dist.ReduceOp.MAX if world > 1 else None)[0] / steps
        barrier()
        if rank == 0 and "parity" in res:
            res["parity"]["max_abs_err_host_image"] = float(np.abs(np.array(host_np) - ref).max())
        res["e2e"] = {"value": headline["rays_per_frame"] / (e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e, "h2d_bytes_per_step": 496 * world, "d2h_bytes_per_step": W * H * 16,
                      "call": "gvtb200_render_frame into pinned host memory" if use_waves else "gvtb200_domain_frame_persist + gvtb200_framebuffer_read_rows on every rank (rows of the root's image over NVLink, then this rank's PCIe link)"}
        cleanup()
        res.update(ms_per_frame=headline["ms_per_frame"], Mrays_per_s=headline["Mrays_per_s"], rays_per_frame=headline["rays_per_frame"], schedule=headline["schedule"], counters=ctr_all)
        if world > 1:
            barrier()
            e.p2p_detach()
            barrier()
        # ---- the same scene, film and schedule with samples^2 rays per pixel: a frame with enough rays per lane to show the
        # scheduler's throughput (at 1 spp and 8 GPUs a lane sees ~15 rays and the frame is as long as its longest ray chain)
        if a.gyroid_samples > 1:
            ns = a.gyroid_samples
            eye, focus, up, fov, depth, _, jitter = e.recipe["cam"]
            g.check(g.lib().gvtb200_scene_set_camera(e.h, g._fp(g.f32(eye)), g._fp(g.f32(focus)), g._fp(g.f32(up)), fov, depth, ns, jitter))
            spp[0] = ns * ns
            if world > 1:
                schedulers.p2p_connect(e, dist, rank, world, dev)
                install()
            more, _ = measure(headline["schedule"] if world > 1 else "persistent kernel, all bricks on one GPU", persist)
            more["samples_per_pixel"] = ns * ns
            if world > 1:
                more.update(persist_stats())
                persist()
                e.sync()
                barrier()
                if rank == 0:
                    img = e.fb_read(W, H)
                    e.recipe["cam"] = (eye, focus, up, fov, depth, ns, jitter)
                    full, imported_ref = schedulers.whole_scene_on_one_rank(e, descs, owner, rank, mode=g.IMPORT_MAP)
                    ref4, _ = full.render(W, H)
                    more["parity_max_abs_err"] = float(np.abs(img - ref4).max())
                    full.close()
                    del imported_ref
                barrier()
                e.p2p_detach()
                barrier()
            else:
                def wave4():
                    e.render_async(H)

                w4, _ = measure("wave path", wave4)
                more["wave_path_ms_per_frame"] = w4["ms_per_frame"]
                more["one_gpu_ms_per_frame"] = min(w4["ms_per_frame"], more["ms_per_frame"])
            res["more_rays_per_pixel"] = more
        return res, e, d

    # ================= the headline workload ========================================================================================
    domain = a.workload == "gyroid"
    weak = world > 1 and not domain and a.scaling == "weak"
    if domain:
        hres, e, d = run_gyroid()
    else:
        film = None
        if weak:
            d0 = WORKLOADS[a.workload]["make"]()
            f = world ** 0.5
            film = (int(round(d0.width * f)), int(round(d0.height * f)))
        hres, e, d = run_image(a.workload, film, headline=True, scaling="weak" if weak else "strong")
    W, H = d.width, d.height
    ms_per_step, rays_frame = hres["ms_per_frame"], hres["rays_per_frame"]

    out = None
    roof = None
    cpu = None
    if rank == 0 and world == 1 and not domain:
        pk, pk_kind = peaks()
        # counted frame: per-launch snapshots of the device counters -> algorithmic bytes of every launch
        e.profile_launches(True)
        e.profile_counters(True)
        e.render_async(H)
        e.sync()
        recs = e.launch_records()
        e.profile_counters(False)
        names = {0: "k_camera_route", 1: "k_trace (closest hit)", 2: "k_shade", 3: "k_trace (any hit, shadow rays)"}
        zero = {k: 0 for k in g.COUNTER_FIELDS}
        cap, cap_src = ncu_capture(a.workload, "waves")
        cap_l = cap["launches"] if cap else []
        if cap and len(cap_l) != len(recs):
            cap, cap_src, cap_l = None, f"capture has {len(cap_l)} launches, the frame has {len(recs)}", []
        clk = (hres["clocks"] or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
        sms = g.device_info()["sm_count"]
        issue_peak = sms * 4 * clk * 1e6  # warp instructions per second: one per scheduler per clock
        launches_out = []
        for i, (kind, _, after) in enumerate(recs):
            before = recs[i - 1][2] if i else zero
            dlt = {k: after[k] - before[k] for k in after}
            ms_i = statistics.mean(t[i][1] for t in hres["trace_ms"])
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
            lo = {"kernel": names[kind], "ms": ms_i, "share": ms_i / ms_per_step, "rays": n_in, "nodes_visited": dlt["nodes_visited"], "tris_tested": dlt["tris_tested"],
                  "algorithmic_bytes": ray_b + scene_b, "algorithmic_GBps_through_L1_L2": (ray_b + scene_b) / (ms_i * 1e-3) / 1e9}
            if cap:
                c = cap_l[i]
                inst, dram = c.get("smsp__inst_executed.sum"), (c.get("dram__bytes_read.sum", 0.0) + c.get("dram__bytes_write.sum", 0.0))
                lo.update(warp_inst=inst, issue_frac=inst / (ms_i * 1e-3) / issue_peak, dram_bytes=dram, hbm_frac=dram / (ms_i * 1e-3) / 1e9 / pk["hbm_gbs"],
                          simt_lanes_per_inst=c.get("smsp__thread_inst_executed_per_inst_executed.ratio"), l1_wavefronts_pct_of_peak=c.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                          ncu_ms=c.get("gpu__time_duration.sum", 0.0) * 1e3, ncu_kernel=c.get("kernel"))
            launches_out.append(lo)
        dom = max(range(len(launches_out)), key=lambda i: launches_out[i]["ms"])
        dl = launches_out[dom]
        roof = {"kernel": f"{dl['kernel']}, launch {dom} of the frame ({100 * dl['share']:.0f}% of the step)",
                "per_ray": {"nodes_visited": dl["nodes_visited"] / max(1, dl["rays"]), "tris_tested": dl["tris_tested"] / max(1, dl["rays"])},
                "algorithmic_bytes_per_launch": dl["algorithmic_bytes"], "launches": launches_out, "counters_source": cap_src}
        if cap:
            issue, hbm = dl["issue_frac"], dl["hbm_frac"]
            wi = sum(l["warp_inst"] for l in launches_out)
            roof.update({
                "bound": "issue" if issue >= hbm else "hbm", "frac": max(issue, hbm),
                "achieved": dl["warp_inst"] / (dl["ms"] * 1e-3) / 1e9 if issue >= hbm else dl["dram_bytes"] / (dl["ms"] * 1e-3) / 1e9,
                "peak": issue_peak / 1e9 if issue >= hbm else pk["hbm_gbs"], "unit": "Gwarp-inst/s" if issue >= hbm else "GB/s",
                "traffic": dl["dram_bytes"], "issue_frac": issue, "hbm_frac": hbm, "simt_lanes_per_inst": dl["simt_lanes_per_inst"], "l1_wavefronts_pct_of_peak": dl["l1_wavefronts_pct_of_peak"],
                "peak_source": f"issue: {sms} SMs x 4 schedulers x {clk:.0f} MHz (SM clock sampled during the timed region); hbm: {pk_kind} MEASURED_PEAKS.json hbm_gbs (burst copy)",
                "frame": {"issue_frac": wi / (ms_per_step * 1e-3) / issue_peak, "hbm_frac": sum(l["dram_bytes"] for l in launches_out) / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "lane_issue_frac": sum(l["warp_inst"] * l["simt_lanes_per_inst"] for l in launches_out) / 32.0 / (ms_per_step * 1e-3) / issue_peak},
                "note": "SURVEY 8(d): frac = max(issue term, HBM term) of the dominant launch; nodes + triangles are L1/L2 resident at this scene size, so DRAM carries the ray stream only",
            })
        else:
            roof.update({"bound": "issue", "frac": None, "achieved": None, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s", "traffic": None,
                         "note": "no ncu capture of THESE kernel sources is committed (" + str(cap_src) + "): the issue and HBM terms are not reported rather than taken from a stale file"})
        if not a.no_cpu:
            v, sample, _, _ = cpu_reference(WORKLOADS[a.workload]["make"], threads)
            cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample, "note": vs_note}
    if rank == 0 and world == 1 and domain and not a.no_cpu:
        v, sample, _, _ = cpu_reference(lambda: library.gyroid_scene(128, 2, 3840, 2160), threads)
        cpu = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "note": vs_note,
               "sample": sample + "; gyroid at res 128 (1.8 M triangles): the CPU port cannot build the 100 M-triangle BVH inside a bench run"}
    del e

    # ================= the other BASELINE workloads ================================================================================
    extra = {}
    wanted = None if a.extra is None else set(x for x in a.extra.split(",") if x)
    if not a.no_extra:
        names = [n for n in EXTRA_ALL_N + (EXTRA_N1 if world == 1 else ()) if n != a.workload and (wanted is None or n in wanted)]
        if world > 1 and a.workload == "sibenik" and weak and (wanted is None or "sibenik_strong" in wanted):
            names.append("sibenik_strong")
        for name in names:
            try:
                if name == "gyroid":
                    r2, e2, _ = run_gyroid(steps=10, warmup=3)
                elif name == "sibenik_strong":
                    r2, e2, _ = run_image("sibenik", None, steps=10, warmup=3)
                    r2["scaling"] = "strong: the 1920x1080 film dealt over the ranks"
                else:
                    r2, e2, _ = run_image(name, None, steps=10, warmup=3)
                del e2
                r2.pop("trace_ms", None), r2.pop("clocks", None), r2.pop("wall_s", None)
                if world > 1:
                    r2.setdefault("scaling", "strong")
                extra[name] = r2
            except Exception as ex:  # noqa: BLE001
                extra[name] = {"error": f"{type(ex).__name__}: {ex}"}
                if world > 1:
                    raise  # a rank that left a collective sequence cannot rejoin it

    if rank == 0:
        if domain:
            sched_desc = hres["schedule"]
            wl = dict(wl, name=hres["workload"])
        else:
            sched_desc = "ImageScheduler" + (f", {world} ranks, row tiles of {a.tile_rows} dealt round-robin, " + ("rows stored into rank 0's framebuffer over NVLink (CUDA IPC), no collective" if p2p else "NCCL reduce to rank 0") if world > 1 else "")
        out = {
            "metric": "Mrays/s (primary+shadow+secondary)", "value": hres["Mrays_per_s"], "unit": "Mrays/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": wl["name"] + (f"; weak scaling: film {W}x{H} = {world} x the N=1 film's pixels, same camera and aspect" if weak else ""), "film": [W, H], "depth": d.depth, "samples": d.samples, "rays_per_frame": rays_frame, "bvh": "binned SAH, 8-wide, 128 B nodes",
                "scheduler": sched_desc,
                "l2": "256 MiB buffer rewritten between timed frames (L2 flush)", "timing": "CUDA events per frame on the launching stream, max over ranks",
                "reference_arm": vs_note,
            },
            "e2e": hres["e2e"],
            "gpu_launches": int(hres.get("launches", a.steps)), "clocks": hres.get("clocks"), "wall_s_timed_region": hres.get("wall_s"),
            "counters": hres["counters"], "Mrays_per_s_entered_an_instance": hres.get("Mrays_per_s_entered_an_instance"),
        }
        if "parity" in hres:
            out["parity"] = hres["parity"]
        if domain:
            out["config"]["gyroid"] = {k: v for k, v in hres.items() if k not in ("counters", "e2e")}
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
