"""Multi-GPU frame drivers: one process per GPU (torchrun), torch.distributed for the plumbing.

  ImageScheduler   (reference: algorithm/ImageTracer.h:111-125,127-269): every rank holds all geometry and renders a
                   shard of the film rows; the framebuffer gather replaces icetCompositeImage
                   (composite/IceTComposite.cpp:96).  No data-path collective: pixels are independent.
  DomainScheduler  (reference: algorithm/DomainTracer.h:115-496): instances are owned by ranks, rays migrate.  The MPI
                   count exchange (:409-415) and payload exchange (:433-463) become one all_to_all_single of counts and
                   one all_to_all_single of 80-byte wire records over NCCL/NVLink; the gather/scatter termination test
                   (:339-349) becomes an all_reduce(sum); IceT compositing becomes reduce(sum) + clamp on the device.

Both schedulers also have an NVLink peer path (`p2p_connect`, then `p2p=True` / `render_domain_schedule_p2p`): the ranks map
each other's framebuffer and inbound ray queues with CUDA IPC; the trace kernel stores migrating rays straight into the
owner's queue and the film rows / touched pixels go straight into the root's framebuffer, so the only collective left on
the data path is the 8-byte termination all-reduce of the DomainScheduler.

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


def gather_descriptors(dist, desc, world, device="cpu"):
    """all-gather of one fixed-size byte string per rank (works with gloo on CPU tensors and with nccl on CUDA tensors)."""
    import torch

    mine = torch.frombuffer(bytearray(desc), dtype=torch.uint8).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def p2p_connect(engine, dist, rank, world, device):
    """Maps every peer's framebuffer, inbound ray queues and mailbox into this rank (CUDA IPC over NVLink)."""
    desc = engine.p2p_export(world)
    engine.p2p_attach(rank, world, gather_descriptors(dist, desc, world, device))
    dist.barrier()  # nobody stores into a peer before every peer has attached


def render_image_schedule(engine, width, height, dist=None, rank=0, world=1, tile_rows=8, fb_tensor=None, p2p=False):
    """One ImageScheduler frame.  Leaves the composited image in rank 0's device framebuffer.
    p2p=True (after p2p_connect): the ranks store their clamped rows into rank 0's framebuffer over NVLink, no collective."""
    if p2p and world > 1:
        engine.image_frame_p2p(height, tile_rows, 0)
        return
    engine.render_tiles(height, tile_rows, rank, world)
    if world > 1:
        # every pixel is written by exactly one rank, the others hold zeros: sum == gather
        dist.reduce(fb_tensor, dst=0, op=dist.ReduceOp.SUM)


def domain_terminated(log):
    """Termination rule of the peer path: `log` holds the all-reduced cumulative exported-ray totals after each round; the
    frame is over when a round exported nothing anywhere (DomainTracer.h:339-349: no rank has rays in flight)."""
    return len(log) >= 1 and (log[-1] == (log[-2] if len(log) > 1 else 0))


def render_domain_schedule_p2p(engine, dist, rank, world, max_rounds=10000, root=0):
    """One DomainScheduler frame over the NVLink peer path (after p2p_connect).  The engine must launch on torch's current
    stream (engine.set_stream) so that the counter read below is ordered after the round's kernels.  Returns the
    all-reduced cumulative exported-ray total after each round."""
    log = []
    engine.domain_begin_frame()
    for rnd in range(max_rounds):
        ptr = engine.domain_round_p2p(rnd)
        total = device_tensor(ptr, (1,), "<i8").clone()
        dist.all_reduce(total)  # the only collective: 8 bytes per round (and the barrier that ends the round)
        log.append(int(total.item()))
        if domain_terminated(log):
            break
    engine.domain_composite_p2p(root)
    return log


class AsyncTermination:
    """Termination rule of the asynchronous DomainScheduler (Mattern's four-counter method; stands in for the two-phase
    vote of core/comm/vote/vote.cpp).  Feed it the all-reduced (busy, sent, received) of every poll, in order; a poll's
    values must have been read after the previous poll's all-reduce ended.  The frame is over when two consecutive polls
    find no rank busy and the same totals with sent == received: nothing was in flight during the second poll."""

    def __init__(self):
        self.prev = None

    def update(self, busy, sent, received):
        quiet = busy == 0 and sent == received
        done = quiet and self.prev == (sent, received)
        self.prev = (sent, received) if quiet else None
        return done


class AsyncBoard:
    """Control plane of the asynchronous DomainScheduler on one box: a few words per rank in POSIX shared memory (a file under
    /dev/shm mapped by every rank).  It carries no rays -- those go over NVLink -- only what the reference's vote messages
    carry (core/comm/vote/vote.cpp): who is idle and how many rays each rank has sent and received.  A collective cannot do
    this job: an NCCL kernel queues behind a rank's persistent trace kernel (every SM is taken), so one busy rank would stall
    the polls of all the others.  Words are int64; each rank writes only its own entry (seqlock: the sequence word is odd while
    the entry is being written), rank 0 alone writes the header."""

    HDR = 8                 # header words: [0] wave (coordinator), [1] frame whose rendering is over (coordinator)
    ENT = 8                 # entry words: seq, wave_seen, busy, sent, received, frame, ready_frame, -

    def __init__(self, path, rank, world, create):
        self.path, self.rank, self.world = path, rank, world
        n = self.HDR + self.ENT * world
        if create:
            np.zeros(n, np.int64).tofile(path)
        self.a = np.memmap(path, dtype=np.int64, mode="r+", shape=(n,))
        self.wave = 0       # coordinator only: the wave being collected
        self.term = AsyncTermination()

    @classmethod
    def connect(cls, dist, rank, world):
        """Rank 0 creates the board, the others map it (the path travels through torch.distributed)."""
        import os
        import tempfile

        box = [None]
        if rank == 0:
            root = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            fd, box[0] = tempfile.mkstemp(prefix="gvtb200_board_", dir=root)
            os.close(fd)
            board = cls(box[0], rank, world, True)
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        if rank != 0:
            board = cls(box[0], rank, world, False)
        if world > 1:
            dist.barrier()
        return board

    def close(self):
        import os

        del self.a
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def _e(self, r):
        return self.HDR + self.ENT * r

    def publish(self, wave_seen, busy, sent, received, frame):
        a, e = self.a, self._e(self.rank)
        a[e] += 1  # odd: being written
        a[e + 1], a[e + 2], a[e + 3], a[e + 4], a[e + 5] = wave_seen, int(busy), sent, received, frame
        a[e] += 1

    def read(self, r):
        a, e = self.a, self._e(r)
        while True:
            s1 = int(a[e])
            v = [int(x) for x in a[e + 1:e + 6]]
            if not (s1 & 1) and int(a[e]) == s1:
                return v

    def frame_barrier(self, frame, spin=None):
        """Every rank has finished gvtb200_domain_async_reset for `frame` (nobody publishes tails into a rank that has not
        cleared them yet)."""
        self.a[self._e(self.rank) + 6] = frame
        while any(int(self.a[self._e(r) + 6]) < frame for r in range(self.world)):
            if spin:
                spin()

    def current_wave(self):
        return int(self.a[0])

    def finished(self, frame):
        return int(self.a[1]) >= frame

    def coordinate(self, frame):
        """Rank 0, once per poll: when every rank has published a snapshot taken after the current wave started, apply the
        four-counter rule to the wave and start the next one."""
        ents = [self.read(r) for r in range(self.world)]
        if any(e[4] != frame or e[0] != self.wave for e in ents):
            return
        if self.term.update(sum(e[1] for e in ents), sum(e[2] for e in ents), sum(e[3] for e in ents)):
            self.a[1] = frame
            self.term = AsyncTermination()
        self.wave += 1
        self.a[0] = self.wave


def render_domain_schedule_async(engine, board, rank, world, frame, root=0, timeout_s=120.0):
    """One DomainScheduler frame without exchange rounds (scheduler::AsyncDomain, tracer/Domain/DomainTracer.cpp:109-192)
    over the NVLink peer path (after p2p_connect).  Every rank free-runs: each poll it asks its engine to trace whatever its
    peers have published (non-blocking: nothing happens while its previous kernels are still running) and posts (busy, sent,
    received) on the shared-memory board; rank 0 also plays the coordinator of the termination waves.  `frame` is a counter
    that is the same on every rank and grows by one per call (starting at 1).  Returns the number of polls.  A rank that
    waits longer than timeout_s for the others (one of them died) raises instead of spinning for ever."""
    import time

    t0 = time.monotonic()

    def check_time():
        if time.monotonic() - t0 > timeout_s:
            raise RuntimeError(f"asynchronous DomainScheduler: rank {rank} gave up after {timeout_s:.0f} s (frame {frame})")

    engine.domain_async_reset()
    board.frame_barrier(frame, spin=check_time)
    engine.domain_begin_frame()
    if rank == 0:
        board.wave += 1  # a wave that starts inside this frame: snapshots of the previous frame cannot satisfy it
        board.a[0] = board.wave
    poll = 0
    while True:
        poll += 1
        if (poll & 1023) == 0:
            check_time()
        wave = board.current_wave()  # read BEFORE the snapshot: a snapshot tagged w was taken after wave w started
        busy, sent, received = engine.domain_async_step()
        board.publish(wave, busy, sent, received, frame)
        if rank == 0:
            board.coordinate(frame)
        if board.finished(frame):
            engine.domain_composite_p2p(root)
            return poll


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


def gyroid_balanced_owner(world):
    """Brick -> rank map for BASELINE config 4 (2x2x2 bricks, index = bx + 2 by + 4 bz, camera on the (1,1,1) diagonal).
    In GraviT the owner of a domain is whichever rank loaded it (`Locations`, DomainTracer.h:115-144) -- data placement is
    the application's choice.  Seen from the diagonal, brick 7 shows three quarter-faces, bricks 3/5/6 two, bricks 1/2/4 one
    and brick 0 none, so `k mod N` puts 2/3 of the first-hit work on one of two ranks; this placement evens it out for
    N = 2 and 4 (N = 8 is one brick per rank either way)."""
    table = {
        1: [0] * 8,
        2: [0, 0, 0, 1, 0, 1, 1, 0],      # {7,0,1,2,4} = 3+0+1+1+1 | {3,5,6} = 2+2+2
        4: [0, 1, 2, 1, 3, 2, 3, 0],      # {7,0} | {3,1} | {5,2} | {6,4}: 3 units each
        8: list(range(8)),
    }
    if world not in table:
        return domain_owner(8, world)
    return np.asarray(table[world], dtype=np.int32)


def build_domain_rank(desc, rank, world, flags=None, boxes=None, owner=None):
    """The scene one DomainScheduler rank holds: every instance (matrices + boxes, for routing) but only the meshes of
    the instances it owns (DomainTracer.h:115-144: instance k lives on rank k % world).  `desc` is a scenes.library
    description.  A mesh this rank does not own may have verts=None; then boxes[k] = its (lo.xyz, hi.xyz) bounding box
    (ranks exchange them, see mesh_boxes_allgather) and a two-corner placeholder stands in for it in the host layer, so
    that api::addInstance computes the same instance box on every rank.  Returns (StandaloneScene, owner table)."""
    import gravit_b200 as g

    owner = domain_owner(len(desc.instances), world) if owner is None else np.asarray(owner, dtype=np.int32)
    flags = desc.flags if flags is None else flags
    needed = [any(owner[i] == rank and mi == k for i, (mi, _) in enumerate(desc.instances)) for k in range(len(desc.meshes))]
    host = g.GvtScene(flags=flags)  # host layer: Mesh::addFace filter, normals, addInstance matrices and boxes
    ids = []
    for k, m in enumerate(desc.meshes):
        if m.verts is not None:
            ids.append(host.add_mesh(m.verts, m.tris, one_based=m.one_based, kd=m.kd))
        else:
            if needed[k]:
                raise ValueError(f"mesh {k} is owned by rank {rank} but has no geometry")
            lo, hi = np.asarray(boxes[k][:3], np.float32), np.asarray(boxes[k][3:], np.float32)
            # Mesh::addFace drops the degenerate face, Mesh::computeBoundingBox still sees the two corners
            ids.append(host.add_mesh(np.stack([lo, hi, lo]), np.array([[1, 2, 3]], np.uint32), one_based=True, kd=m.kd))
    for mesh, m16 in desc.instances:
        host.add_instance(ids[mesh], m16)
    lights = np.array([g.point_light(p, c) for p, c in desc.lights], g.LIGHT_DTYPE)
    inst = [(mi,) + tuple(host.instance_get(i)) for i, (mi, _) in enumerate(desc.instances)]
    meshes = []
    for k, m in enumerate(desc.meshes):
        if not needed[k]:
            meshes.append(None)
            continue
        nv = len(np.asarray(m.verts).reshape(-1, 3))
        faces, normals, _ = host.mesh_get(k, nv)
        meshes.append(g.Mesh(m.verts, faces, normals=normals, material=(0, tuple(m.kd or (0.5, 0.5, 0.5)), (0.5, 0.5, 0.5), 1.0)))
    cam = (desc.eye, desc.focus, desc.up, desc.fov, desc.depth, desc.samples, desc.jitter)
    scene = g.StandaloneScene(meshes, inst, lights, cam, (desc.width, desc.height), flags=flags, owner=owner, rank=rank, world=world)
    return scene, owner


def mesh_boxes_allgather(desc, dist, world):
    """Bounding boxes of all meshes of a description whose geometry is spread over the ranks (verts=None where another
    rank holds it): each rank contributes the boxes of the meshes it holds; max/min-reduced across ranks."""
    import torch

    n = len(desc.meshes)
    lo = np.full((n, 3), np.inf, np.float32)
    hi = np.full((n, 3), -np.inf, np.float32)
    for k, m in enumerate(desc.meshes):
        if m.verts is not None and len(m.verts):
            v = np.asarray(m.verts, np.float32).reshape(-1, 3)
            lo[k], hi[k] = v.min(0), v.max(0)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        tl, th = torch.as_tensor(lo, device=dev), torch.as_tensor(hi, device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MIN)
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
        lo, hi = tl.cpu().numpy(), th.cpu().numpy()
    return np.concatenate([lo, hi], axis=1)
