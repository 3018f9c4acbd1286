"""CPU, world_size 2 over gloo: the host-side multi-rank logic (unique-id exchange, sharding,
the exactness of the integer partial-sum exchange that the NCCL allreduce performs)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "3d-point-cloud-multiday-imagery_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = importlib.import_module(PKG + ".dist")
        # 1. unique-id broadcast: every rank ends with rank 0's 128 bytes
        uid = d.exchange_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128))
        # 1b. all-gather of equal-length byte strings in rank order (CUDA IPC handles travel this way)
        got = d.gather_bytes(bytes([rank + 1]) * 64, rank, world)
        assert got == b"".join(bytes([r + 1]) * 64 for r in range(world))
        # 2. sharding + exchange of fixed-point partial sums is exact and order independent:
        #    each rank quantises its shard like the kernels do, int64 sums are all-reduced.
        rs = np.random.RandomState(0)
        X = rs.uniform(-1000, 1000, size=(10001, 3)).astype(np.float32)
        labels = rs.randint(0, 5, size=X.shape[0])
        b, e = d.shard_range(X.shape[0], rank, world)
        scale = 2.0 ** 12
        q_local = np.rint(X[b:e].astype(np.float64) * scale).astype(np.int64)
        acc = np.zeros((5, 4), dtype=np.int64)
        for j in range(5):
            m = labels[b:e] == j
            acc[j, :3] = q_local[m].sum(axis=0)
            acc[j, 3] = m.sum()
        t = torch.from_numpy(acc)
        dist.all_reduce(t)
        q_all = np.rint(X.astype(np.float64) * scale).astype(np.int64)
        ref = np.zeros((5, 4), dtype=np.int64)
        for j in range(5):
            m = labels == j
            ref[j, :3] = q_all[m].sum(axis=0)
            ref[j, 3] = m.sum()
        assert np.array_equal(t.numpy(), ref)  # bit-identical for any number of ranks
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


class _FakeEngine:
    """Stands in for Engine in init_engine_comm: records the calls; p2p_open fails on `bad_rank`."""

    def __init__(self, rank, bad_rank):
        self.device, self.rank_, self.bad = None, rank, bad_rank
        self.p2p = False
        self.calls = []

    def make_unique_id(self):
        return bytes(range(128))

    def init_comm(self, world, rank, uid):
        self.calls.append(("init_comm", world, rank, uid))

    def p2p_handle(self):
        return bytes([self.rank_]) * 64

    def p2p_open(self, handles):
        self.calls.append(("p2p_open", len(handles)))
        self.p2p = self.rank_ != self.bad
        return self.p2p

    def p2p_close(self):
        self.calls.append(("p2p_close",))
        self.p2p = False


def _worker_fallback(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = importlib.import_module(PKG + ".dist")
        # peer access missing on rank 1 only: EVERY rank must take the same (collective) way back
        # to the NCCL exchange -- one p2p_close each, no second unique-id broadcast / comm init --
        # and nobody may be left with p2p switched on
        eng = _FakeEngine(rank, bad_rank=1)
        d.init_engine_comm(eng, rank, world)
        names = [c[0] for c in eng.calls]
        assert names == ["init_comm", "p2p_open", "p2p_close"], names
        assert eng.p2p is False
        # the next collective still pairs up on every rank (nothing is left half-entered)
        t = torch.tensor([rank + 1])
        dist.all_reduce(t)
        assert int(t.item()) == world * (world + 1) // 2
        # all ranks fine: the fused exchange stays on, nothing is closed
        eng = _FakeEngine(rank, bad_rank=-1)
        d.init_engine_comm(eng, rank, world)
        assert [c[0] for c in eng.calls] == ["init_comm", "p2p_open"] and eng.p2p is True
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def _run(target, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    return sorted(results)


def test_p2p_open_failure_on_one_rank_falls_back_collectively():
    assert _run(_worker_fallback) == [(0, "ok"), (1, "ok")]


def test_two_rank_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
