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


def gradient_buckets(grad, cut_a, end_a, cut_v):
    """Views of the flat gradient arena (order: head | audio_net | visual_net | 3 losses) for the two-bucket
    all-reduce: `late` = layer3 + layer4 of both encoders (their gradients are final first, 94 % of the
    bytes) plus the loss tail, `early` = head + stem / layer1 / layer2.  cut_a / cut_v: arena offsets where
    layer3 of the audio / visual encoder starts; end_a: where the audio group ends."""
    assert 0 < cut_a < end_a < cut_v < grad.numel()
    return dict(late=[grad[cut_a:end_a], grad[cut_v:]], early=[grad[:cut_a], grad[end_a:cut_v]])


def allreduce_async(tensors, group=None):
    """Start SUM all-reduces that run on the backend's own stream beside whatever the caller enqueues next;
    returns the work handles (wait() makes the current stream — not the host — wait on NCCL)."""
    return [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tensors]


def allreduce_finish(works, tensors, group=None):
    """Join the asynchronous bucket, then all-reduce the remaining (small) tensors on the critical path."""
    for w in works:
        w.wait()
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
