"""Multi-GPU frame drivers: one process per GPU (torchrun), torch.distributed for the plumbing.

  ImageScheduler   (reference: algorithm/ImageTracer.h:111-125,127-269): every rank holds all geometry and renders a
                   shard of the film rows; the framebuffer gather replaces icetCompositeImage
                   (composite/IceTComposite.cpp:96).  No data-path collective: pixels are independent.
  DomainScheduler  (reference: algorithm/DomainTracer.h:115-496): instances are owned by ranks, rays migrate.  The MPI
                   count exchange (:409-415) and payload exchange (:433-463) become one all_to_all_single of counts and
                   one all_to_all_single of 80-byte wire records over NCCL/NVLink; the gather/scatter termination test
                   (:339-349) becomes an all_reduce(sum); IceT compositing becomes reduce(sum) + clamp on the device.

The host logic (row shards, split sizes, termination) is backend-agnostic and is covered on CPU with gloo, world_size 2.
"""
import numpy as np

WIRE_BYTES = 80  # the reference's Ray record (actor/Ray.h:95); the first pad word carries the target instance id


def image_shard(height, rank, world, tile_rows=8):
    """Row tiles owned by `rank`: tiles of `tile_rows` rows dealt round-robin (interleaving balances the load; the
    reference's contiguous slices are tile_rows = ceil(height / world))."""
    return dict(row0=0, row1=height, tile_rows=tile_rows, index=rank, stride=world)


def shard_rows(height, rank, world, tile_rows=8):
    """The explicit list of film rows of a shard (host mirror of the kernel's RowTiles arithmetic)."""
    rows = []
    b = 0
    while True:
        start = (b * world + rank) * tile_rows
        if start >= height:
            break
        rows.extend(range(start, min(height, start + tile_rows)))
        b += 1
    return rows


def domain_owner(n_instances, world):
    """mpiInstanceMap (DomainTracer.h:115-144) for one mesh location list per instance: instance k -> rank k % world."""
    return np.arange(n_instances, dtype=np.int32) % world


def exchange_counts(dist, send_counts, device):
    """SendRays phase 1: every rank learns how many rays each peer sends it."""
    import torch

    s = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    return r.cpu().numpy()


def exchange_rays(dist, send_buf, send_counts, recv_counts):
    """SendRays phase 2: send_buf is a (n, 80) uint8 tensor packed peer-major; returns the (m, 80) received records."""
    import torch

    m = int(np.sum(recv_counts))
    recv = torch.empty((m, WIRE_BYTES), dtype=torch.uint8, device=send_buf.device)
    dist.all_to_all_single(recv, send_buf, output_split_sizes=[int(c) for c in recv_counts], input_split_sizes=[int(c) for c in send_counts])
    return recv


class _DevArray:
    """__cuda_array_interface__ view of engine-owned device memory, so that torch (NCCL) can read/write it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape, typestr="<f4"):
    import torch

    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


def render_image_schedule(engine, width, height, dist=None, rank=0, world=1, tile_rows=8, fb_tensor=None):
    """One ImageScheduler frame.  Leaves the composited image in rank 0's device framebuffer."""
    engine.render_tiles(height, tile_rows, rank, world)
    if world > 1:
        # every pixel is written by exactly one rank, the others hold zeros: sum == gather
        dist.reduce(fb_tensor, dst=0, op=dist.ReduceOp.SUM)


def render_domain_schedule(engine, dist, rank, world, device, fb_tensor, max_rounds=10000):
    """One DomainScheduler frame (DomainTracer.h:185-368).  Returns per-round [sent-to-peer] count matrices rows."""
    import torch

    log = []
    engine.domain_begin_frame()
    for _ in range(max_rounds):
        engine.domain_drain_local()
        counts, d_ptr, n = engine.domain_export(world)
        log.append(counts.copy())
        recv_counts = exchange_counts(dist, counts, device)
        total = torch.as_tensor([int(n)], device=device, dtype=torch.int64)
        dist.all_reduce(total)  # termination: no rank has rays in flight (DomainTracer.h:339-349)
        if int(total.item()) == 0:
            break
        send = device_tensor(d_ptr, (int(n), WIRE_BYTES), "|u1") if n else torch.empty((0, WIRE_BYTES), dtype=torch.uint8, device=device)
        recv = exchange_rays(dist, send, counts, recv_counts)
        torch.cuda.current_stream().synchronize()
        if recv.shape[0]:
            engine.domain_import(recv.data_ptr(), recv.shape[0])
    engine.sync()
    dist.reduce(fb_tensor, dst=0, op=dist.ReduceOp.SUM)  # a pixel can collect light on several ranks
    if rank == 0:
        engine.fb_clamp()
    return log
