#!/usr/bin/env python
"""Multi-GPU correctness + timing check (run under torchrun on the GPU box, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

ImageScheduler: the N-rank composited frame must equal rank 0's own single-GPU frame bit for bit (disjoint pixels).
DomainScheduler: the N-rank frame (NCCL ray exchange, reduce + clamp composite) must equal the single-GPU frame to
fp32 reassociation (a pixel can collect light on several ranks).  Prints one JSON line per check from rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    import gravit_b200 as g
    from gravit_b200 import schedulers
    from scenes import library

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    g.init(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def say(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    # ---- ImageScheduler: sibenik 640x360 depth 2 and the 25-instance golden scene ----
    for name, d in (("sibenik", library.sibenik_scene(640, 360, depth=1)), ("simple", library.simple_scene(512, 512))):
        s = library.apply(d, g.GvtScene(flags=d.flags if name == "sibenik" else 1))
        e = s.engine()
        e.set_stream(stream.cuda_stream)
        ref, _ = e.render(d.width, d.height)
        ptr, _ = e.device_framebuffer()
        fb_t = schedulers.device_tensor(ptr, (d.height, d.width, 4))
        schedulers.render_image_schedule(e, d.width, d.height, dist if world > 1 else None, rank, world, 8, fb_t)
        torch.cuda.synchronize()
        if rank == 0:
            out = fb_t.cpu().numpy()
            say(check="image_scheduler", scene=name, world=world, bit_identical=bool(np.array_equal(out, ref)), lit_pixels=int((ref[..., 3] > 0).sum()))
            assert np.array_equal(out, ref)
        if world > 1:  # the same frame over the NVLink peer path (CUDA IPC), and both variants timed
            schedulers.p2p_connect(e, dist, rank, world, dev)
            for _ in range(2):
                schedulers.render_image_schedule(e, d.width, d.height, dist, rank, world, 8, fb_t, p2p=True)
            e.sync()
            dist.barrier()
            if rank == 0:
                out = fb_t.cpu().numpy()
                say(check="image_scheduler_p2p", scene=name, world=world, bit_identical=bool(np.array_equal(out, ref)))
                assert np.array_equal(out, ref)
            ms = {}
            for variant in ("nccl_reduce", "p2p"):
                dist.barrier()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    schedulers.render_image_schedule(e, d.width, d.height, dist, rank, world, 8, fb_t, p2p=(variant == "p2p"))
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms[variant] = float(t.item())
            say(check="image_scheduler_timing", scene=name, world=world, ms_per_frame=ms)
            e.sync()
            e.p2p_detach()
        e.set_stream(None)
        del e, s

    # ---- DomainScheduler: the golden scene (instances dealt round-robin) and a gyroid split into 8 bricks ----
    for name, d, flags in (("simple", library.simple_scene(512, 512), 1), ("gyroid64", library.gyroid_scene(64, 2, 960, 540), 0)):
        single = library.apply(d, g.GvtScene(flags=flags))
        ref, cref = single.engine().render(d.width, d.height)
        scene, owner = schedulers.build_domain_rank(d, rank, world, flags=flags)
        scene.set_stream(stream.cuda_stream)
        ptr, _ = scene.device_framebuffer()
        fb_t = schedulers.device_tensor(ptr, (d.height, d.width, 4))
        if world > 1:
            log = schedulers.render_domain_schedule(scene, dist, rank, world, dev, fb_t)
        else:
            log = []
            scene.domain_begin_frame(), scene.domain_drain_local(), scene.sync(), scene.fb_clamp()
        torch.cuda.synchronize()
        c = scene.counters()
        tot = torch.tensor([c["rays_exported"], c["primary_hits"], c["fb_adds"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        if rank == 0:
            out = fb_t.cpu().numpy()
            err = float(np.abs(out - ref).max())
            say(check="domain_scheduler", scene=name, world=world, max_abs_err=err, rounds=len(log), rays_exchanged=int(tot[0]),
                hits=int(tot[1]), hits_single=cref["primary_hits"], fb_adds=int(tot[2]), fb_adds_single=cref["fb_adds"])
            assert err < 1e-5 and int(tot[1]) == cref["primary_hits"] and int(tot[2]) == cref["fb_adds"]
        if world > 1:  # the same frame with the trace kernel storing migrating rays into the owner's queue over NVLink
            schedulers.p2p_connect(scene, dist, rank, world, dev)
            for _ in range(2):
                logp = schedulers.render_domain_schedule_p2p(scene, dist, rank, world)
            scene.sync()
            dist.barrier()
            c = scene.counters()
            tot = torch.tensor([c["rays_exported"], c["primary_hits"], c["fb_adds"], c["queue_overflows"]], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            if rank == 0:
                out = fb_t.cpu().numpy()
                err = float(np.abs(out - ref).max())
                say(check="domain_scheduler_p2p", scene=name, world=world, max_abs_err=err, rounds=len(logp), rays_exchanged=int(tot[0]),
                    hits=int(tot[1]), hits_single=cref["primary_hits"], fb_adds=int(tot[2]), fb_adds_single=cref["fb_adds"], queue_overflows=int(tot[3]))
                assert err < 1e-5 and int(tot[1]) == cref["primary_hits"] and int(tot[2]) == cref["fb_adds"] and int(tot[3]) == 0
            ms = {}
            for variant in ("nccl_all_to_all", "p2p"):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    if variant == "p2p":
                        schedulers.render_domain_schedule_p2p(scene, dist, rank, world)
                    else:
                        schedulers.render_domain_schedule(scene, dist, rank, world, dev, fb_t)
                scene.sync()
                torch.cuda.synchronize()
                t = torch.tensor([(time.perf_counter() - t0) * 100.0], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms[variant] = float(t.item())
            say(check="domain_scheduler_timing", scene=name, world=world, ms_per_frame_wall=ms)
            scene.p2p_detach()
        scene.set_stream(None)
        del scene, single
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
