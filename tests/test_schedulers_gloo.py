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


def test_async_termination_rule():
    t = schedulers.AsyncTermination()
    assert not t.update(1, 0, 0)       # somebody is busy
    assert not t.update(0, 10, 7)      # rays in flight
    assert not t.update(0, 10, 10)     # first quiet poll
    assert not t.update(0, 12, 12)     # totals moved between the polls: something was in flight after all
    assert not t.update(1, 0, 0)
    assert not t.update(0, 12, 12)
    assert t.update(0, 12, 12)         # second consecutive quiet poll with the same totals


class _ScriptedEngine:
    """Stands in for the engine in render_domain_schedule_async: two ranks pass 'rays' through a shared directory (one file
    per published tail), so sent/received follow the real protocol: a rank is busy while it 'traces' for two polls after
    new rays arrived, and forwards a shrinking share of them to its peer."""

    def __init__(self, rank, world, root):
        os.makedirs(root, exist_ok=True)
        self.rank, self.world, self.root = rank, world, root
        self.sent = self.received = 0
        self.sent_to = [0] * world  # what travels goes to the next rank
        self.work = 0  # polls of tracing left
        self.pending_out = 0
        self.first = True
        self.composited = False

    def _tail_path(self, src, dst):
        return os.path.join(self.root, f"tail_{src}_{dst}")

    def _publish(self):
        for dst in range(self.world):
            if dst != self.rank:
                tmp = self._tail_path(self.rank, dst) + ".tmp"
                with open(tmp, "w") as f:
                    f.write(str(self.sent_to[dst]))
                os.replace(tmp, self._tail_path(self.rank, dst))

    def domain_async_reset(self):
        self._publish()

    def domain_begin_frame(self):
        pass

    def domain_async_step(self):
        if self.first:  # the camera rays this rank kept: rank r forwards 8 + 4 r rays
            self.first = False
            self.work, self.pending_out = 2, 8 + 4 * self.rank
            return True, 0, 0
        if self.work:
            self.work -= 1
            if self.work == 0:
                self.sent += self.pending_out
                self.sent_to[(self.rank + 1) % self.world] += self.pending_out
                self._publish()
            return True, 0, 0
        tails = 0
        for src in range(self.world):
            if src != self.rank:
                try:
                    tails += int(open(self._tail_path(src, self.rank)).read())
                except (FileNotFoundError, ValueError):
                    pass
        if tails > self.received:
            new = tails - self.received
            self.received = tails
            self.work, self.pending_out = 2, new // 2  # half of what arrives moves on
            return True, 0, 0
        return False, self.sent, self.received

    def domain_composite_p2p(self, root):
        self.composited = True


def _async_worker(rank, world, port, root, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        board = schedulers.AsyncBoard.connect(dist, rank, world)
        res = []
        for frame in (1, 2):  # twice: the board is reused from frame to frame
            e = _ScriptedEngine(rank, world, os.path.join(root, f"f{frame}"))
            polls = schedulers.render_domain_schedule_async(e, board, rank, world, frame)
            res.append((polls, e.sent, e.received, e.composited))
        dist.barrier()
        board.close()
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_async_domain_schedule(tmp_path, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_async_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for f in range(2):
        per_rank = [res[r][1][f] for r in range(world)]
        assert all(x[3] for x in per_rank)                                       # every rank left the loop and composited
        assert sum(x[1] for x in per_rank) == sum(x[2] for x in per_rank)        # everything sent was received: not cut off early
        assert all(x[1] >= 8 for x in per_rank)
