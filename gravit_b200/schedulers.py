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
the data path is the 8-byte termination all-reduce of the round-synchronous DomainScheduler.  `render_domain_persist` removes
that one too: one persistent kernel per rank and frame, termination found on the device (gvtb200_domain_frame_persist), and
`plan_servers` / `apply_server_plan` let several ranks serve a hot domain (replicas or NVLink-mapped BLAS).

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


def render_domain_persist(engine, root=0, tile_rows=8):
    """One DomainScheduler frame as ONE persistent kernel per rank (after p2p_connect; world 1 needs no peers): camera rays
    are sharded 1/world per rank, migrating rays go into the serving rank's inbound queue over NVLink and are picked up by
    waiting lanes, termination and the composite happen on the device.  No collective, no host round trip inside the frame
    (the reference's nearest relative is scheduler::AsyncDomain, tracer/Domain/DomainTracer.cpp:109-192).  Asynchronous on
    the engine's stream; engine.sync() reports time-outs and full queues."""
    engine.domain_frame_persist(root, tile_rows)


def calibrate_shader_sms(engine, frame, time_frames, candidates=(2, 3, 4, 6, 8, 10, 14, 20), splits=(16, 32)):
    """Two knobs of the persistent kernel are properties of the scene and of how many rays a GPU gets per frame, and are
    set by timing a few frames per candidate (a scene-setup step like the server plan):
      * how many SMs shade instead of tracing (hits per ray against the cost of tracing one);
      * up to how many traversing lanes a warp shares its rays' stacks out to its free lanes (gvtb200_set_tail_split): with
        rays to spare a free lane is better off waiting for the next refill (16), with few rays per lane -- small films,
        many GPUs -- every free lane should help (32).
    frame(): renders one frame (asynchronous); time_frames(frame) -> ms per frame, the same number on every rank (the
    caller's max over ranks).  Returns (best shader SMs, {candidate: ms}, best split, {candidate: ms})."""
    import gravit_b200 as g

    tried = {}
    for sm in candidates:
        engine.domain_persist_config(0, sm)
        tried[sm] = float(time_frames(frame))
        engine.sync()
    best = min(tried, key=tried.get)
    engine.domain_persist_config(0, best)
    tried_split = {}
    for sp in splits:
        g.check(g.lib().gvtb200_set_tail_split(sp))
        tried_split[sp] = float(time_frames(frame))
        engine.sync()
    best_split = min(tried_split, key=tried_split.get)
    if splits and tried_split[best_split] > 0.98 * tried_split[splits[0]]:
        best_split = splits[0]  # within the noise of a few frames: stay with the first (the default)
    g.check(g.lib().gvtb200_set_tail_split(best_split))
    return best, tried, best_split, tried_split


def plan_servers(load, owner, world, max_replicas=1, max_servers=4, slack=1.02):
    """Which ranks serve which instance, and with what share, so that the ranks' loads even out.

    load[k]   measured cost of instance k in the last frame (wide BVH nodes visited, all ranks summed)
    owner[k]  the rank that owns instance k (mpiInstanceMap, DomainTracer.h:115-144)
    Every rank may hold up to `max_replicas` BLAS copies of instances it does not own (GraviT's hybrid schedulers load hot
    domains on idle ranks; here the copy comes from the owner over NVLink).  Greedy: the least loaded rank with room takes a
    copy of the most loaded rank's heaviest instance; after every step the shares of all multi-served instances are
    re-balanced by water-filling.  Returns (ranks, weights): (n, 4) int32 (-1 = unused) and (n, 4) float32."""
    load = np.asarray(load, np.float64)
    n = len(load)
    servers = [[int(owner[k])] for k in range(n)]
    share = [{int(owner[k]): 1.0} for k in range(n)]
    held = [0] * world

    def rank_load():
        rl = np.zeros(world)
        for k in range(n):
            for r, x in share[k].items():
                rl[r] += load[k] * x
        return rl

    def rebalance(sweeps=30):
        for _ in range(sweeps):
            rl = rank_load()
            for k in range(n):
                if len(servers[k]) < 2 or load[k] <= 0:
                    continue
                others = {r: rl[r] - load[k] * share[k][r] for r in servers[k]}
                # water level L: sum(max(0, L - others)) == load[k]
                lv = sorted(others.values())
                level = None
                for m in range(len(lv), 0, -1):
                    cand = (load[k] + sum(lv[:m])) / m
                    if cand >= lv[m - 1]:
                        level = cand
                        break
                for r in servers[k]:
                    new = max(0.0, level - others[r]) / load[k]
                    rl[r] += load[k] * (new - share[k][r])
                    share[k][r] = new

    total = load.sum()
    if total > 0:
        for _ in range(world * max_replicas):
            rl = rank_load()
            r_max = int(np.argmax(rl))
            if rl[r_max] <= slack * total / world:
                break
            room = [r for r in range(world) if held[r] < max_replicas and r != r_max]
            if not room:
                break
            r_min = min(room, key=lambda r: rl[r])
            cands = sorted((k for k in range(n) if r_max in servers[k] and r_min not in servers[k] and len(servers[k]) < max_servers),
                           key=lambda k: -load[k] * share[k][r_max])
            if not cands or rl[r_min] >= rl[r_max]:
                break
            servers[cands[0]].append(r_min)
            share[cands[0]][r_min] = 0.0
            held[r_min] += 1
            rebalance()
    ranks = np.full((n, 4), -1, np.int32)
    weights = np.zeros((n, 4), np.float32)
    for k in range(n):
        live = [r for r in servers[k] if share[k][r] > 1e-4 or r == owner[k]]
        for j, r in enumerate(live):
            ranks[k, j], weights[k, j] = r, max(share[k][r], 1e-4 if r == owner[k] else 0.0)
    return ranks, weights


def planned_rank_load(load, ranks, weights, world):
    """Per-rank load a server plan predicts (the planner's objective; reported beside the measured busy times)."""
    rl = np.zeros(world)
    for k in range(len(load)):
        w = weights[k] / max(weights[k].sum(), 1e-30)
        for j in range(4):
            if ranks[k, j] >= 0:
                rl[ranks[k, j]] += load[k] * w[j]
    return rl


def plan_holders(load, owner, world, max_replicas=1, slack=1.02, candidates=8):
    """Replica placement for the *local first* policy: a rank that holds a copy of a domain traces every ray that meets that
    domain on this rank itself; only rays that meet a domain the rank does not hold are sent away, to one of the domain's
    holders chosen by `share`.  Every replica therefore removes ray traffic as well as load from the owner -- with a copy
    of everything on every rank the DomainScheduler degenerates into the ImageScheduler (no ray leaves its rank), which is
    what 180 GB of HBM per GPU allow for scenes of a few GB; `max_replicas` (copies of other ranks' domains per rank) is the
    memory knob in between.  GraviT's `Locations` list (DomainTracer.h:115-144) and its hybrid schedulers ("load the hot
    domain where the rays are") are the reference's form of this.

    Load model: the demand for domain k arises uniformly over the ranks (camera rays are sharded by interleaved row tiles);
    rank r gets load[k]/world of it if it holds k, and share[k][r] of the (world - holders)/world part that arises elsewhere.
    Greedy: the step (rank with room, domain it does not hold) that gives the smallest predicted maximum rank load is taken,
    ties -- everything within `slack` of the mean counts as balanced -- go to the heaviest domain (most traffic removed);
    after every step the shares are re-balanced by water-filling.
    Returns (holders, share): holders[k] = ranks holding k, owner first; share[k] = {rank: share of the non-local demand}."""
    load = np.asarray(load, np.float64)
    n = len(load)
    total = float(load.sum())
    holders = [[int(owner[k])] for k in range(n)]
    share = [{int(owner[k]): 1.0} for k in range(n)]
    if total <= 0 or world < 2:
        return holders, share
    L = load / total

    def away(k):  # the part of k's demand that arises on ranks which do not hold it
        return L[k] * (world - len(holders[k])) / world

    def rank_load():
        rl = np.zeros(world)
        for k in range(n):
            for r in holders[k]:
                rl[r] += L[k] / world + away(k) * share[k][r]
        return rl

    def rebalance(sweeps=20):
        rl = rank_load()
        for _ in range(sweeps):
            for k in range(n):
                a = away(k)
                if len(holders[k]) < 2 or a <= 0:
                    continue
                others = {r: rl[r] - a * share[k][r] for r in holders[k]}
                lv = sorted(others.values())
                level = lv[0] + a
                for m in range(len(lv), 0, -1):  # water level: sum(max(0, level - others)) == a
                    cand = (a + sum(lv[:m])) / m
                    if cand >= lv[m - 1]:
                        level = cand
                        break
                for r in holders[k]:
                    new = max(0.0, level - others[r]) / a
                    rl[r] += a * (new - share[k][r])
                    share[k][r] = new
        return rl

    held = [0] * world
    heavy = sorted(range(n), key=lambda k: -load[k])
    floor = slack / world
    for _ in range(world * max_replicas):
        best = None
        for r in range(world):
            if held[r] >= max_replicas:
                continue
            for k in [k for k in heavy if r not in holders[k] and load[k] > 0][:candidates]:
                saved = [list(holders[j]) for j in range(n)], [dict(share[j]) for j in range(n)]
                holders[k].append(r)
                share[k][r] = 0.0
                worst = max(float(rebalance().max()), floor)
                key = (round(worst, 6), -load[k], r)
                if best is None or key < best[0]:
                    best = (key, r, k)
                holders[:], share[:] = saved
        if best is None:
            break
        _, r, k = best
        holders[k].append(r)
        share[k][r] = 0.0
        held[r] += 1
        rebalance()
    return holders, share


def planned_rank_load_local(load, holders, share, world):
    """Per-rank load (fraction of the frame's) and the fraction of instance visits that start with a ray transfer, as
    plan_holders' model predicts them."""
    load = np.asarray(load, np.float64)
    L = load / max(load.sum(), 1e-30)
    rl = np.zeros(world)
    moved = 0.0
    for k in range(len(load)):
        a = L[k] * (world - len(holders[k])) / world
        moved += a
        for r in holders[k]:
            rl[r] += L[k] / world + a * share[k].get(r, 0.0)
    return rl, moved


def server_table_for_rank(holders, share, rank, max_servers=4):
    """The server table rank `rank` installs under the local-first policy (every rank's table is different): a domain the
    rank holds is served by the rank itself; any other domain by up to `max_servers` of its holders, by share."""
    n = len(holders)
    ranks = np.full((n, 4), -1, np.int32)
    weights = np.zeros((n, 4), np.float32)
    for k in range(n):
        if rank in holders[k]:
            ranks[k, 0], weights[k, 0] = rank, 1.0
            continue
        top = sorted(holders[k], key=lambda r: -share[k].get(r, 0.0))[:max_servers]
        live = [r for r in top if share[k].get(r, 0.0) > 1e-4] or [holders[k][0]]
        s = sum(share[k].get(r, 0.0) for r in live) or 1.0
        for j, r in enumerate(live):
            ranks[k, j], weights[k, j] = r, max(share[k].get(r, 0.0) / s, 1e-4)
    return ranks, weights


def apply_holder_plan(engine, holders, share, descs, rank, owner, mode=0, imported=None):
    """Import the BLAS of every domain this rank holds without owning it (what is already in `imported` is kept), attach
    it, re-commit and install this rank's local-first server table.  Returns the dict of imported meshes (keep it alive)."""
    import gravit_b200 as g

    imported = {} if imported is None else imported
    fresh = False
    for k in range(len(holders)):
        if rank in holders[k] and int(owner[k]) != rank and k not in imported:
            m = g.Mesh.from_desc(descs[k], mode)
            engine.set_instance_mesh(k, m)
            imported[k] = m
            fresh = True
    if fresh:
        engine.commit()
    engine.domain_set_servers(*server_table_for_rank(holders, share, rank))
    return imported


def exchange_mesh_descs(dist, scene_meshes, instances, owner, rank, world, device="cpu"):
    """Every rank exports the meshes of the instances it owns; returns desc[k] (bytes) for every instance k."""
    import gravit_b200 as g

    blank = bytes(g.MESH_DESC_BYTES)
    mine = b"".join((scene_meshes[mi].export() if (owner[k] == rank and scene_meshes[mi] is not None) else blank) for k, (mi, *_rest) in enumerate(instances))
    if world == 1:
        allb = mine
    else:
        allb = gather_descriptors(dist, mine, world, device)
    n, sz = len(instances), g.MESH_DESC_BYTES
    return [allb[(int(owner[k]) * n + k) * sz:(int(owner[k]) * n + k + 1) * sz] for k in range(n)]


def apply_server_plan(engine, ranks, weights, descs, rank, mode=0):
    """Import (replica: mode 0, NVLink mapping: mode 1) the BLAS of every instance this rank serves without owning it, attach
    it to the instance, re-commit, and install the server table.  Returns the imported meshes (keep them alive)."""
    import gravit_b200 as g

    imported = {}
    for k in range(len(ranks)):
        if rank in [int(r) for r in ranks[k] if r >= 0][1:]:  # entry 0 is the owner
            m = g.Mesh.from_desc(descs[k], mode)
            engine.set_instance_mesh(k, m)
            imported[k] = m
    if imported:
        engine.commit()
    engine.domain_set_servers(ranks, weights)
    return imported


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
    scene.recipe = dict(inst=inst, lights=lights, cam=cam, film=(desc.width, desc.height), flags=flags)  # what a second scene over the same meshes needs
    return scene, owner


def whole_scene_on_one_rank(scene, descs, owner, rank, mode=1):
    """A single-rank scene over ALL instances of a domain-partitioned scene: this rank's own meshes plus the other ranks'
    BLAS imported through their descriptors (mode 1: mapped over NVLink, no copy).  Used as the parity reference of a
    multi-GPU DomainScheduler frame: one GPU tracing everything, however slowly.  Returns (scene, imported meshes)."""
    import gravit_b200 as g

    rc = scene.recipe
    meshes, imported = [], []
    for k, (mi, *_rest) in enumerate(rc["inst"]):
        if owner[k] == rank:
            meshes.append(scene._keep[mi])
        else:
            m = g.Mesh.from_desc(descs[k], mode)
            imported.append(m)
            meshes.append(m)
    inst = [(k,) + tuple(x[1:]) for k, x in enumerate(rc["inst"])]
    full = g.StandaloneScene(meshes, inst, rc["lights"], rc["cam"], rc["film"], flags=rc["flags"])
    return full, imported


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
