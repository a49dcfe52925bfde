"""Host-side data-parallel logic (replaces the reference's torch.nn.DataParallel wrapper,
main_dgl.py:244): one process per GPU, contiguous batch shards in DataParallel `chunk` order,
per-replica BatchNorm, gradient SUM all-reduce of shards whose losses are scaled by 1/B_global
(== the reference's full-batch mean over gathered logits, main_dgl.py:102-104)."""
import torch
import torch.distributed as dist


def shard_range(rank, world, batch):
    """Rows [lo, hi) of the global batch owned by `rank` (torch.chunk order: rank 0 == reference
    device 0, which owns the persisted BN running statistics)."""
    per = (batch + world - 1) // world
    lo = min(rank * per, batch)
    return lo, min(lo + per, batch)


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of a flat gradient arena (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def flatten_grads(grads, names):
    return torch.cat([grads[k].reshape(-1) for k in names])
