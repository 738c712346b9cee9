"""Multi-GPU plumbing: the batch is partitioned across ranks (one process per GPU); the solve path
has NO collective — the only exchange is an allgather of the 32-byte per-problem result records
(J, c_max, iteration counts, status) after the solve."""
import numpy as np

from .api import RESULT_DTYPE


def shard_range(B_total, rank, world):
    """Contiguous slice [b0, b1) of the global batch owned by `rank` (ceil(B/world) per rank)."""
    per = (B_total + world - 1) // world
    b0 = min(B_total, rank * per)
    return b0, min(B_total, b0 + per)


def allgather_results(local_records, dist, device=None):
    """All-gather the structured result records of every rank (ragged shards are padded).
    `dist` is torch.distributed (nccl on GPU, gloo in the CPU tests)."""
    import torch
    world = dist.get_world_size()
    cnt = torch.tensor([len(local_records)], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    per = max(cnts)
    buf = np.zeros(per, dtype=RESULT_DTYPE)
    buf[: len(local_records)] = local_records
    t = torch.from_numpy(buf.view(np.uint8).copy()).to(device) if device is not None else torch.from_numpy(buf.view(np.uint8).copy())
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy().view(RESULT_DTYPE)[:c] for o, c in zip(outs, cnts)]
    return np.concatenate(parts)
