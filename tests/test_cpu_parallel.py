"""world_size-2 gloo test of the data-parallel host logic: contiguous shards, CE scaled by
1/B_global on every rank, SUM all-reduce == the single-process computation over both shards
(per-replica BatchNorm, like the reference's DataParallel)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
    from gdl_b200.parallel import allreduce_sum_, flatten_grads, shard_range
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    B = 4
    spec, image, label = make_batch(B, 6, "tiny", seed=5)
    lo, hi = shard_range(rank, world, B)
    sd = O.init_state("concat", "CREMAD", 0)
    res = O.dgl_step(sd, {}, spec[lo:hi], image[lo:hi], label[lo:hi], fusion="concat", alpha=4.0,
                     max_norm=1e9, inv_batch=1.0 / B, apply_update=False)
    names = sorted(res["grads"].keys())
    flat = flatten_grads(res["grads"], names)
    losses = torch.tensor(res["losses"], dtype=torch.float64)
    allreduce_sum_(flat)
    allreduce_sum_(losses)
    if rank == 0:
        torch.save({"flat": flat, "losses": losses, "names": names}, os.path.join(out_dir, "dp.pt"))
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_sharded_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from gdl_b200.parallel import flatten_grads, shard_range
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    # single-process simulation of the two replicas
    torch.set_num_threads(4)
    B = 4
    spec, image, label = make_batch(B, 6, "tiny", seed=5)
    total, losses = None, torch.zeros(3, dtype=torch.float64)
    for r in range(2):
        lo, hi = shard_range(r, 2, B)
        assert (lo, hi) == (2 * r, 2 * r + 2)
        sd = O.init_state("concat", "CREMAD", 0)
        res = O.dgl_step(sd, {}, spec[lo:hi], image[lo:hi], label[lo:hi], fusion="concat", alpha=4.0,
                         max_norm=1e9, inv_batch=1.0 / B, apply_update=False)
        flat = flatten_grads(res["grads"], got["names"])
        total = flat if total is None else total + flat
        losses += torch.tensor(res["losses"], dtype=torch.float64)
    assert torch.allclose(got["flat"], total, rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["losses"], losses, rtol=1e-6)


def test_shard_range_covers_batch():
    from gdl_b200.parallel import shard_range
    for B in (1, 7, 64, 256):
        for world in (1, 2, 4, 8):
            rows = []
            for r in range(world):
                lo, hi = shard_range(r, world, B)
                rows += list(range(lo, hi))
            assert rows == list(range(B))
