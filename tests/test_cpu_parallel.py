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


def _bucket_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
    from gdl_b200.parallel import allreduce_async, allreduce_finish, gradient_buckets
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(5000 + 64, generator=g)
    whole = flat.clone()
    b = gradient_buckets(flat, 700, 2100, 3900)
    works = allreduce_async(b["late"])          # bucket 1 in flight ...
    flat[:700].mul_(1.0)                        # ... while the "early" gradients are still being produced
    allreduce_finish(works, b["early"])
    dist.all_reduce(whole)
    if rank == 0:
        torch.save({"bucketed": flat, "whole": whole}, os.path.join(out_dir, "buckets.pt"))
    dist.destroy_process_group()


def test_two_bucket_allreduce_equals_one_allreduce(tmp_path):
    """The overlapped two-bucket exchange covers the arena exactly once."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_bucket_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "buckets.pt"))
    assert torch.equal(got["bucketed"], got["whole"])


def test_late_bucket_is_a_contiguous_tail_of_each_encoder():
    """ParamArena.late_split: layer3 + layer4 form the tail of each encoder's arena group, and the four
    bucket views partition the gradient arena (incl. the loss tail) without overlap."""
    import argparse
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
    import gdl_b200
    from gdl_b200.parallel import gradient_buckets
    from gdl_b200.step import ParamArena
    args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modality="full")
    gdl_b200.setup_seed(0)
    model = gdl_b200.AVClassifier_DGL(args)
    arena = ParamArena(model, torch.device("cpu"))
    cuts = []
    for gid, net in ((0, model.audio_net), (1, model.visual_net)):
        late = [p for n, p in net.named_parameters() if n.startswith(("layer3.", "layer4."))]
        cut = arena.late_split(gid, late)
        start, end = arena.group_ranges[gid]
        frac = (end - cut) / (end - start)
        assert 0.9 < frac < 0.96, frac  # layer3 + layer4 hold ~94 % of a ResNet-18 encoder
        cuts.append(cut)
    b = gradient_buckets(arena.grad, cuts[0], arena.group_ranges[0][1], cuts[1])
    arena.grad.zero_()
    for t in b["late"] + b["early"]:
        t.add_(1.0)
    assert torch.equal(arena.grad, torch.ones_like(arena.grad))
    # a set that is not a tail is rejected
    try:
        arena.late_split(0, [p for n, p in model.audio_net.named_parameters() if n.startswith("layer2.")])
    except RuntimeError:
        pass
    else:
        raise AssertionError("late_split accepted a non-tail parameter set")


def test_checkpoint_momentum_is_adopted_by_the_arena():
    """Resume path: momentum buffers that `optimizer.load_state_dict` restored are copied
    into the fused optimizer's arena and `optimizer.state_dict()` keeps carrying the arena's values."""
    import argparse
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
    import gdl_b200
    from gdl_b200.step import ParamArena
    from gdl_b200.train import adopt_momentum

    def build():
        args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modality="full")
        gdl_b200.setup_seed(0)
        m = gdl_b200.AVClassifier_DGL(args)
        return m, torch.optim.SGD(m.parameters(), lr=0.001, momentum=0.9, weight_decay=1e-4)

    # "previous run": an arena whose momentum the fused kernel has been updating, saved the reference way
    m0, opt0 = build()
    a0 = ParamArena(m0, torch.device("cpu"))
    assert adopt_momentum(a0, opt0) is False            # fresh run: nothing to adopt, buffers alias the arena
    a0.momentum.copy_(torch.arange(a0.numel, dtype=torch.float32) * 1e-6)
    saved = {"model": m0.state_dict(), "optimizer": opt0.state_dict()}
    # never-trained parameters (fc_auxi) carry no optimizer state, exactly like torch.optim.SGD
    assert len(saved["optimizer"]["state"]) == len(a0.params) < len(list(m0.parameters()))

    # "resumed run"
    m1, opt1 = build()
    m1.load_state_dict(saved["model"])
    opt1.load_state_dict(saved["optimizer"])
    a1 = ParamArena(m1, torch.device("cpu"))

    class _Step:
        momentum_loaded = False
    st = _Step()
    assert adopt_momentum(a1, opt1, st) is True and st.momentum_loaded

    def views(a):
        return [a.momentum[o:o + p.numel()] for p, o in zip(a.params, a.offsets)]
    for v0, v1 in zip(views(a0), views(a1)):            # (the alignment padding between tensors is not state)
        assert torch.equal(v0, v1)
    p = a1.params[3]
    assert opt1.state[p]["momentum_buffer"].data_ptr() == a1.momentum[a1.offsets[3]:].data_ptr()
    for v in views(a1):                                 # the fused kernel keeps updating the arena ...
        v.mul_(2.0)
    again = opt1.state_dict()["state"]                  # ... and the next checkpoint carries those values
    total = sum(v["momentum_buffer"].double().sum().item() for v in again.values())
    want = sum(v.double().sum().item() for v in views(a1))
    assert abs(total - want) <= 1e-9 * max(1.0, abs(want))


def test_chunk_shard_sampler_reproduces_dataparallel_split():
    """ChunkShardBatchSampler: the N ranks' index lists of step k, concatenated in rank order, are exactly global
    batch k of the reference's DataLoader(shuffle=True, drop_last=True) (same torch RNG seed => same RandomSampler
    permutation), i.e. rank r holds the rows nn.DataParallel's chunk would give device r (main_dgl.py:244,284-288);
    a second epoch reshuffles (the ADVICE item: DistributedSampler without set_epoch replayed one order)."""
    sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
    from torch.utils.data import BatchSampler, RandomSampler
    from gdl_b200.parallel import ChunkShardBatchSampler, shard_range
    data = list(range(203))
    B, world = 16, 4
    torch.manual_seed(7)
    ref_sampler = BatchSampler(RandomSampler(data), B, drop_last=True)
    ref_epochs = [list(ref_sampler), list(ref_sampler)]
    per_rank = []
    for r in range(world):
        torch.manual_seed(7)  # every rank runs the same script from the same seed
        s = ChunkShardBatchSampler(RandomSampler(data), B, r, world)
        assert len(s) == 203 // B
        e0 = list(s)
        d0 = s.order_digest
        e1 = list(s)
        per_rank.append(((e0, e1), (d0, s.order_digest)))
    for ep in range(2):
        for k, gb in enumerate(ref_epochs[ep]):
            cat = sum((per_rank[r][0][ep][k] for r in range(world)), [])
            assert cat == gb
            for r in range(world):
                lo, hi = shard_range(r, world, B)
                assert per_rank[r][0][ep][k] == gb[lo:hi]
    assert ref_epochs[0] != ref_epochs[1]
    assert len({pr[1] for pr in per_rank}) == 1  # identical digests on every rank, both epochs
    import pytest
    with pytest.raises(ValueError):
        ChunkShardBatchSampler(RandomSampler(data), 10, 0, 4)
