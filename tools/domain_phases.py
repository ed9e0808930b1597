#!/usr/bin/env python
"""Per-phase timing of one DomainScheduler frame over the NVLink peer path (run under torchrun, one rank per GPU):
begin_frame (camera rays + first routing), every exchange round (trace kernels | all-reduce), composite.  Each phase is
bracketed by a device synchronize, so the sum is an upper bound of the pipelined frame time."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import gravit_b200 as g
    from gravit_b200 import schedulers
    from scenes import library

    res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    owner_mode = sys.argv[2] if len(sys.argv) > 2 else "mod"
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    g.init(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    owner = schedulers.domain_owner(8, world) if owner_mode == "mod" else schedulers.gyroid_balanced_owner(world)
    mine = [k for k in range(8) if owner[k] == rank]
    d = library.gyroid_scene(res, 2, 3840, 2160, only_bricks=mine, brick_fn=lambda *a: library.gyroid_brick_torch(*a, device=dev))
    boxes = schedulers.mesh_boxes_allgather(d, dist, world)
    e, _ = schedulers.build_domain_rank(d, rank, world, boxes=boxes, owner=owner)
    e.set_stream(stream.cuda_stream)
    schedulers.p2p_connect(e, dist, rank, world, dev)
    for _ in range(3):
        schedulers.render_domain_schedule_p2p(e, dist, rank, world)
    torch.cuda.synchronize()

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r

    rows = []
    for _ in range(3):
        dist.barrier()
        ph = {}
        ph["begin"], _ = timed(e.domain_begin_frame)
        log = []
        for rnd in range(100):
            t_round, ptr = timed(lambda: e.domain_round_p2p(rnd))
            total = schedulers.device_tensor(ptr, (1,), "<i8").clone()
            t_ar, _ = timed(lambda: dist.all_reduce(total))
            log.append(int(total.item()))
            ph[f"round{rnd}"] = round(t_round, 3)
            ph[f"allreduce{rnd}"] = round(t_ar, 3)
            if schedulers.domain_terminated(log):
                break
        ph["composite"], _ = timed(lambda: e.domain_composite_p2p(0))
        ph["begin"] = round(ph["begin"], 3)
        ph["composite"] = round(ph["composite"], 3)
        ph["exported_cumulative"] = log
        rows.append(ph)
    c = e.counters()
    out = {"rank": rank, "world": world, "res": res, "owner": [int(x) for x in owner], "bricks": mine, "phases_ms": rows[-1],
           "instance_visits": c["instance_visits"], "primary_hits": c["primary_hits"], "rays_exported": c["rays_exported"]}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        for o in gathered:
            print(json.dumps(o), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
