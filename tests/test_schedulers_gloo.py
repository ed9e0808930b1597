"""Host logic of the multi-GPU drivers on CPU: world_size 2, gloo backend, no GPU.

Covers the ImageScheduler row sharding (every film row owned exactly once, matches the kernel's RowTiles arithmetic),
the DomainScheduler ownership map, and the SendRays replacement (count exchange, 80-byte wire-record exchange,
termination all-reduce) with synthetic ray records standing in for the engine's export/import buffers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gravit_b200 import schedulers


def test_row_shards_partition_the_film():
    for H, world, tile in ((1080, 8, 8), (1080, 3, 16), (17, 4, 8), (5, 8, 8), (2160, 2, 1)):
        seen = np.zeros(H, np.int32)
        for r in range(world):
            rows = schedulers.shard_rows(H, r, world, tile)
            seen[rows] += 1
            # the kernel's closed form (kernels.cuh tiles_local_rows)
            period, full, rem = tile * world, H // (tile * world), H % (tile * world)
            extra = min(tile, max(0, rem - r * tile))
            assert len(rows) == full * tile + extra
        assert np.all(seen == 1)


def test_domain_owner_round_robin():
    assert schedulers.domain_owner(8, 8).tolist() == list(range(8))
    assert schedulers.domain_owner(5, 2).tolist() == [0, 1, 0, 1, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        # three exchange rounds with shrinking traffic, then termination
        inbox_total = 0
        rounds = 0
        n_out = [7 + 3 * rank, 2, 0]
        for k in range(10):
            n = n_out[k] if k < len(n_out) else 0
            dest = rng.integers(0, world, n)
            dest = dest[dest != rank] if n else dest
            recs = np.zeros((len(dest), schedulers.WIRE_BYTES), np.uint8)
            recs[:, 0] = rank  # sender tag
            recs[:, 1] = dest
            recs[:, 64] = 40 + dest  # "target instance id" in the first pad word
            order = np.argsort(dest, kind="stable")  # peer-major, as gvtb200_domain_export packs them
            recs, dest = recs[order], dest[order]
            counts = np.bincount(dest, minlength=world)
            recv_counts = schedulers.exchange_counts(dist, counts, "cpu")
            total = torch.tensor([len(dest)], dtype=torch.int64)
            dist.all_reduce(total)
            if int(total) == 0:
                break
            got = schedulers.exchange_rays(dist, torch.from_numpy(recs), counts, recv_counts).numpy()
            assert got.shape == (int(recv_counts.sum()), schedulers.WIRE_BYTES)
            assert np.all(got[:, 1] == rank) and np.all(got[:, 64] == 40 + rank)  # every record reached its owner
            # peer-major arrival order
            assert np.all(np.diff(got[:, 0].astype(int)) >= 0)
            inbox_total += len(got)
            rounds += 1
        sent = torch.tensor([sum(1 for _ in range(0))], dtype=torch.int64)
        out.put((rank, rounds, inbox_total))
    finally:
        dist.destroy_process_group()


def test_exchange_and_termination_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 2  # both ranks agree on the number of exchange rounds
    assert res[0][2] + res[1][2] > 0


def test_plan_servers_evens_out_the_ranks():
    """The replica planner (GraviT's hybrid schedulers load hot domains on idle ranks; here a rank imports the owner's BLAS):
    BASELINE configs[3] seen from the cube diagonal puts 3/12 of the first hits on brick 7, 2/12 on bricks 3/5/6, 1/12 on
    bricks 1/2/4 and nothing on brick 0."""
    load = np.array([0.5, 8.3, 8.3, 16.7, 8.3, 16.7, 16.7, 25.0])
    for world in (2, 4, 8):
        owner = schedulers.domain_owner(8, world)
        r0, w0 = schedulers.plan_servers(load, owner, world, max_replicas=0)
        assert np.all(r0[:, 0] == owner) and np.all(r0[:, 1:] == -1)  # no room: the owner serves its domains
        base = schedulers.planned_rank_load(load, r0, w0, world)
        r1, w1 = schedulers.plan_servers(load, owner, world, max_replicas=1)
        rl = schedulers.planned_rank_load(load, r1, w1, world)
        assert np.all(r1[:, 0] == owner)  # the owner always stays first in the list
        assert rl.max() / rl.mean() < 1.03 < base.max() / base.mean()
        assert np.isclose(rl.sum(), load.sum())
        held = np.zeros(world, int)
        for k in range(8):
            for r in r1[k, 1:]:
                if r >= 0:
                    held[r] += 1
                    assert r != owner[k]
        assert held.max() <= 1
    # nothing to balance: nothing is replicated
    r, w = schedulers.plan_servers(np.ones(8), np.arange(8), 8, max_replicas=2)
    assert np.all(r[:, 1:] == -1)
    # all the load in one domain, room for everybody: it is shared four ways (the table holds four servers)
    r, w = schedulers.plan_servers(np.array([100.0, 0, 0, 0, 0, 0, 0, 0]), np.arange(8), 8, max_replicas=1)
    assert (r[0] >= 0).sum() == 4 and np.allclose(w[0], 0.25, atol=0.01)


def test_persistent_termination_rule_model():
    """The termination rule of the persistent DomainScheduler kernel (kernels.cuh persist_control), as a randomised model: each
    rank holds two counters that only grow -- `given` (slots reserved in its inbox, bumped by the PRODUCER before the parent
    ray is retired) and `retired`.  The observer reads every `retired`, then every `given`, at arbitrary moments between the
    ranks' steps; equal sums must imply that nothing was in flight at the instant between the two passes, and once nothing is
    in flight the rule must fire."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        P = int(rng.integers(1, 5))
        given, retired = np.zeros(P, int), np.zeros(P, int)
        flying = []  # rays that exist: (rank that holds it, remaining hops)
        for r in range(P):
            for _ in range(int(rng.integers(0, 4))):
                given[r] += 1
                flying.append([r, int(rng.integers(0, 4))])
        fired = False
        steps = 0
        while not fired:
            steps += 1
            assert steps < 10000
            # the observer's two passes with arbitrary activity in between
            snap_retired = retired.copy()
            for _ in range(int(rng.integers(0, 3))):
                if flying:
                    i = int(rng.integers(0, len(flying)))
                    r, hops = flying[i]
                    if hops > 0:  # forward: the child is counted as given BEFORE the parent is counted as retired
                        dst = int(rng.integers(0, P))
                        given[dst] += 1
                        flying.append([dst, hops - 1])
                        if rng.random() < 0.3:  # ... and may spawn a second ray (a shadow ray)
                            given[r] += 1
                            flying.append([r, 0])
                    retired[r] += 1
                    flying.pop(i)
            in_flight_between = given.sum() - retired.sum()  # an upper bound of what was in flight at the first pass's end
            snap_given = given.copy()
            if snap_retired.sum() == snap_given.sum():
                assert not flying and in_flight_between == 0
                fired = True
        assert given.sum() == retired.sum()


class _FakeMesh:
    def __init__(self, tag):
        self.tag = tag

    def export(self):
        import gravit_b200 as g

        return bytes([self.tag]) * g.MESH_DESC_BYTES


def _plan_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 8
        owner = schedulers.domain_owner(n, world)
        # what a counted frame gives every rank: the node visits it made itself, per instance
        mine = np.array([0.5, 8.3, 8.3, 16.7, 8.3, 16.7, 16.7, 25.0]) * (owner == rank)
        t = torch.from_numpy(mine.copy())
        dist.all_reduce(t)
        ranks_tab, weights = schedulers.plan_servers(t.numpy(), owner, world, max_replicas=1)
        meshes = [_FakeMesh(100 + k) if owner[k] == rank else None for k in range(n)]
        descs = schedulers.exchange_mesh_descs(dist, meshes, [(k,) for k in range(n)], owner, rank, world)
        assert all(d[0] == 100 + k for k, d in enumerate(descs))  # every instance's descriptor came from its owner
        # local first: the same holders on every rank, but a table of its own per rank
        holders, share = schedulers.plan_holders(t.numpy(), owner, world, max_replicas=2)
        tab_r, tab_w = schedulers.server_table_for_rank(holders, share, rank)
        for k, h in enumerate(holders):
            live = [int(x) for x in tab_r[k] if x >= 0]
            assert live == [rank] if rank in h else (live and rank not in live and set(live) <= set(h))
        out.put((rank, ranks_tab.tolist(), np.round(weights, 4).tolist(), [list(map(int, h)) for h in holders]))
    finally:
        dist.destroy_process_group()


def test_server_plan_and_mesh_descriptors_world2():
    """Host side of the shared-BLAS setup on two ranks (gloo): the loads are all-reduced, every rank derives the same plan,
    and the owners' mesh descriptors reach every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]  # the hash plan and the holders are the same on both ranks
    assert any(row[1] >= 0 for row in res[0][1])
    assert any(len(h) == 2 for h in res[0][3])  # somebody holds a replica


def test_local_first_replica_plan():
    """plan_holders: more replicas never move more rays, the predicted load stays near the mean, everything everywhere
    means no transfer at all, and every rank's table serves the domains it holds itself."""
    from gravit_b200 import schedulers as s

    load = [550148, 7170928, 6964897, 24289854, 7175614, 18528249, 24262832, 62267103]  # the 100M gyroid's bricks (node visits)
    for world, owner in ((8, list(range(8))), (4, [0, 1, 2, 1, 3, 2, 3, 0]), (2, [0, 0, 0, 1, 0, 1, 1, 0])):
        own = max(owner.count(r) for r in range(world))
        moved_prev = 1.0
        for R in (0, 1, 2, 3, 8):
            holders, share = s.plan_holders(load, owner, world, max_replicas=R)
            rl, moved = s.planned_rank_load_local(load, holders, share, world)
            assert abs(rl.sum() - 1.0) < 1e-9
            assert moved <= moved_prev + 1e-12
            moved_prev = moved
            for k, h in enumerate(holders):
                assert h[0] == owner[k] and len(set(h)) == len(h)
                assert abs(sum(share[k].values()) - 1.0) < 1e-9
            for r in range(world):
                assert sum(1 for k, h in enumerate(holders) if r in h and owner[k] != r) <= R
                tr, tw = s.server_table_for_rank(holders, share, r)
                for k, h in enumerate(holders):
                    live = [int(x) for x in tr[k] if x >= 0]
                    assert live and set(live) <= set(h)
                    if r in h:
                        assert live == [r]
                    assert abs(sum(float(w) for w, x in zip(tw[k], tr[k]) if x >= 0) - 1.0) < 1e-3
            if R == 0:
                assert all(len(h) == 1 for h in holders)
            if R == 1 and world == 8:
                assert rl.max() * world < 1.1  # one replica per rank is enough to balance the eight bricks
            if R == 8:
                assert moved == 0.0 and all(len(h) == world for h in holders) and abs(rl.max() * world - 1.0) < 1e-9
        assert own >= 1


def test_persistent_kernel_calibration_picks_the_fastest():
    """calibrate_shader_sms: the fastest shader-SM count stays; the stack-sharing width only moves off its default for a
    gain beyond the noise of a few frames.  (Host logic only: a stand-in engine, no GPU.)"""
    import gravit_b200 as g
    from gravit_b200 import schedulers as s

    class Lib:
        def __init__(self):
            self.split = None

        def gvtb200_set_tail_split(self, n):
            self.split = n
            return 0

    class Engine:
        def __init__(self):
            self.sms = None

        def domain_persist_config(self, grid, sms):
            self.sms = sms

        def sync(self):
            pass

    lib, e = Lib(), Engine()
    old_lib, old_check = g.lib, g.check
    g.lib, g.check = (lambda: lib), (lambda rc: None)
    try:
        for gain, want_split in ((0.995, 16), (0.9, 32)):
            cost = lambda: (abs(e.sms - 6) * 0.1 + 3.0) * (gain if lib.split == 32 else 1.0)  # noqa: E731
            best, tried, split, tried_split = s.calibrate_shader_sms(e, lambda: None, lambda f: cost())
            assert best == 6 and e.sms == 6 and set(tried) == {2, 3, 4, 6, 8, 10, 14, 20}
            assert split == want_split and lib.split == want_split and set(tried_split) == {16, 32}
            lib.split = None
    finally:
        g.lib, g.check = old_lib, old_check
