"""Multi-GPU plumbing: one process per GPU, points sharded by contiguous pixel ranges.

torch.distributed is used only to ship the 128-byte NCCL unique id and for barriers; the
per-iteration exchange (K x 4 int64 partial sums + the changed-label count, one
ncclAllReduce on the kernels' own stream) lives inside libmdkm.so.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``total`` items for ``rank`` of ``world``.

    Items are split as evenly as possible in units of ``align`` (e.g. one raster row, so that
    a rank's slice starts on a row boundary; SURVEY.md section 8(e) "whole days / row bands").
    The concatenation over ranks is exactly [0, total).
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (total + align - 1) // align
    base, rem = divmod(units, world)
    b = rank * base + min(rank, rem)
    e = b + base + (1 if rank < rem else 0)
    return min(b * align, total), min(e * align, total)


def exchange_unique_id(make_id, rank: int, world: int, device=None) -> bytes:
    """Rank 0 calls ``make_id()``; the bytes are broadcast through torch.distributed."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return b""
    backend = dist.get_backend()
    dev = torch.device("cuda", device) if (backend == "nccl" and device is not None) else torch.device("cpu")
    if rank == 0:
        raw = make_id()
        t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def gather_bytes(raw: bytes, rank: int, world: int, device=None) -> bytes:
    """All-gather of equal-length byte strings, concatenated in rank order."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", device) if (backend == "nccl" and device is not None) else torch.device("cpu")
    mine = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().tolist()) for p in parts)


def init_engine_comm(engine, rank: int, world: int, p2p: bool = True):
    """Give ``engine`` a communicator spanning the torch.distributed world: NCCL for the
    one-off collectives, and (``p2p``) peer-mapped exchange buffers so that the Lloyd step
    kernel completes the K x 4 partial sums itself over NVLink."""
    import torch
    import torch.distributed as dist

    if world == 1:
        engine.init_comm(1, 0, None)
        return
    uid = exchange_unique_id(engine.make_unique_id, rank, world, device=engine.device)
    engine.init_comm(world, rank, uid)
    if p2p and world <= 8:
        handles = gather_bytes(engine.p2p_handle(), rank, world, device=engine.device)
        ok = engine.p2p_open(handles)
        # all ranks or none: one rank without peer access sends everybody back to NCCL.  The
        # fallback is itself collective -- EVERY rank (whether its own open worked or not) drops
        # its peer mappings and keeps the communicator it already has, so no rank is left
        # waiting in a broadcast / ncclCommInitRank the others never enter.
        backend = dist.get_backend()
        dev = torch.device("cuda", engine.device) if backend == "nccl" else torch.device("cpu")
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            engine.p2p_close()


def torch_gather_mappings(engine):
    """The collective ``api._is_same_clustering`` needs on a sharded engine, over
    torch.distributed: all-gather of the ranks' label maps (k int64 + a flag).  None when
    torch.distributed does not span the engine's ranks (the test is then skipped)."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == engine.n_ranks):
        return None
    backend = dist.get_backend()
    dev = torch.device("cuda", engine.device) if backend == "nccl" else torch.device("cpu")

    def gather(mapping, ok):
        mine = torch.tensor(list(mapping) + [1 if ok else 0], dtype=torch.int64, device=dev)
        parts = [torch.zeros_like(mine) for _ in range(engine.n_ranks)]
        dist.all_gather(parts, mine)
        parts = [p.cpu().numpy() for p in parts]
        return [p[:-1] for p in parts], [bool(p[-1]) for p in parts]

    return gather
