"""CPU oracle for the DGL training step — TEST INFRASTRUCTURE, NOT THE PRODUCT.

A restatement, in plain PyTorch fp32 functional ops on the CPU, of the one hot path of
shicaiwei123/ICCV2025-GDL: `train_epoch` of main_dgl.py.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` leg may import this module; the shipped path
(iccv2025-gdl_b200/) never does.

Why torch and not numpy/C: the reference's arithmetic lives in third-party PyTorch (not under
/root/reference; README.md:7-10 pins "PyTorch 1.11", the container has 2.11.0) — conv2d,
batch_norm, max_pool2d, linear, cross_entropy.  This file restates the *algorithm the reference
composes from those ops* and is pinned against the unmodified reference run in the build
container: tests/golden/make_golden.py imports /root/reference (with stubs for the absent
timm/librosa/skimage), runs main_dgl.train_epoch, and (a) asserts this oracle reproduces it
bit-for-bit in fp32 and (b) writes the golden vectors under tests/golden/.  The reference has
no tests, golden vectors or fixtures of its own, so those generated vectors
are the pin.

What is restated (reference file:line):
  * ResNet-18 encoders without avgpool/fc, audio Cin=1 / visual Cin=3 with frames folded into
    the batch                                   models/backbone.py:52-68 (BasicBlock), :160-201
  * pooling + fusion                            models/basic_model.py:65-86
  * the four *_DGL heads                        models/fusion_modules.py:22-30,51-59,140-178,230-250
  * 3x CrossEntropy, alpha*(La+Lv) backward, fusion-grad wipe, Lf backward
                                                main_dgl.py:102-122
    restated as ONE backward of alpha*(La+Lv)+Lf in which the unimodal branch sees detached
    head parameters and the multimodal branch sees detached features (the reference's own
    .detach() calls) — same effective gradients, proven equal by make_golden.py
  * clip_grad_norm_(40, 2)                      main_dgl.py:129
  * sum_p mean|grad_p| diagnostics (clipped)    main_dgl.py:132-143
  * SGD momentum 0.9, weight decay 1e-4; params without a gradient are skipped
                                                main_dgl.py:154,249
  * parameter initialisation                    models/backbone.py:117-122, utils/utils.py:15-23

`quantize="bf16"` rounds activations, weights and back-propagated activations gradients to
bf16 at the points where the CUDA path stores bf16 tensors (fp32 accumulation everywhere);
it is the "FP32-accumulate check mode" comparison target.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

N_CLASSES = {"VGGSound": 309, "KineticSound": 34, "kinect400": 400, "CREMAD": 6, "AVE": 28}


# ----------------------------------------------------------------------------------------------
# parameter construction (same constructors, same order => same RNG stream as the reference)
# ----------------------------------------------------------------------------------------------
def _resnet18_modules(cin):
    """Modules in the registration order of reference models/backbone.py:97-116."""
    mods = OrderedDict()
    mods["conv1"] = nn.Conv2d(cin, 64, 7, 2, 3, bias=False)
    mods["bn1"] = nn.BatchNorm2d(64)
    inplanes = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            stride = 2 if (li > 1 and bi == 0) else 1
            pre = "layer%d.%d." % (li, bi)
            down = None
            if stride != 1 or inplanes != planes:
                # reference _make_layer builds the downsample BEFORE the block (backbone.py:141-148)
                down = (nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
            mods[pre + "conv1"] = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
            mods[pre + "bn1"] = nn.BatchNorm2d(planes)
            mods[pre + "conv2"] = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            mods[pre + "bn2"] = nn.BatchNorm2d(planes)
            if down is not None:
                mods[pre + "downsample.0"] = down[0]
                mods[pre + "downsample.1"] = down[1]
            inplanes = planes
    # backbone.py:117-122: kaiming-normal convs, N(1, 0.02) BN weights, in self.modules() order
    for m in mods.values():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        else:
            nn.init.normal_(m.weight, mean=1, std=0.02)
            nn.init.constant_(m.bias, 0)
    return mods


def _fusion_modules(fusion, n):
    mods = OrderedDict()
    if fusion == "sum":
        mods["fc_x"] = nn.Linear(512, n)
        mods["fc_y"] = nn.Linear(512, n)
    elif fusion == "concat":
        mods["fc_out"] = nn.Linear(1024, n)
        mods["fc_auxi"] = nn.Linear(1024, n)
    elif fusion == "film":
        mods["fc"] = nn.Linear(512 * 512, 512)
        mods["fc_out"] = nn.Linear(512, n)
    elif fusion == "gated":
        mods["fc_x"] = nn.Linear(512, 512)
        mods["fc_y"] = nn.Linear(512, 512)
        mods["fc_out"] = nn.Linear(512, n)
    else:
        raise NotImplementedError("Incorrect fusion method: {}!".format(fusion))
    return mods


def _weight_init(m):
    """reference utils/utils.py:15-23."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)


def init_state(fusion="concat", dataset="CREMAD", seed=0):
    """state_dict (reference key names, no 'module.' prefix) after setup_seed(seed),
    AVClassifier_DGL(args) and model.apply(weight_init) (main_dgl.py:230-238)."""
    if dataset not in N_CLASSES:
        raise NotImplementedError("Incorrect dataset name {}".format(dataset))
    torch.manual_seed(seed)
    groups = OrderedDict()
    groups["fusion_module"] = _fusion_modules(fusion, N_CLASSES[dataset])  # basic_model.py:28-40
    groups["audio_net"] = _resnet18_modules(1)                             # basic_model.py:43
    groups["visual_net"] = _resnet18_modules(3)                            # basic_model.py:44
    # model.apply visits children depth-first in registration order.  Inside a BasicBlock the
    # registration order is conv1,bn1,conv2,bn2,downsample (backbone.py:44-50), which is the
    # order of the dict above.
    sd = OrderedDict()
    for gname, mods in groups.items():
        for m in mods.values():
            _weight_init(m)
        for name, m in mods.items():
            for k, v in m.state_dict().items():
                sd["%s.%s.%s" % (gname, name, k)] = v.detach().clone()
    return sd


# ----------------------------------------------------------------------------------------------
# bf16 rounding points
# ----------------------------------------------------------------------------------------------
class _RoundFB(torch.autograd.Function):
    """bf16 round-trip in forward AND on the gradient in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundF(torch.autograd.Function):
    """bf16 round-trip in forward only (weights: their gradient stays fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _Force(torch.autograd.Function):
    """Teacher forcing: the forward value is REPLACED by a recorded tensor (the activation the CUDA path
    stored in bf16); the gradient passes straight through, rounded to bf16 like the CUDA path stores it.
    With every rounding point forced, ReLU masks / max-pool arg-maxes / BN statistics of the oracle are
    exactly those of the CUDA forward, so what remains is a test of the backward kernels alone."""

    @staticmethod
    def forward(ctx, x, forced):
        assert forced.shape == x.shape, (forced.shape, x.shape)
        return forced.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32), None


FORCED = []  # activations consumed in call order by _qa when quantize == "forced" (tests/ only)


def _qa(x, quantize):
    if quantize == "bf16":
        return _RoundFB.apply(x)
    if quantize == "forced":
        t = FORCED.pop(0)
        return _RoundFB.apply(x) if t is None else _Force.apply(x, t)
    return x


def _qw(w, quantize):
    return _RoundF.apply(w) if quantize in ("bf16", "forced") else w


# ----------------------------------------------------------------------------------------------
# functional model
# ----------------------------------------------------------------------------------------------
def _bn(x, p, pre, training, quantize, new_buffers):
    rm, rv = p[pre + ".running_mean"], p[pre + ".running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, p[pre + ".weight"], p[pre + ".bias"], True, 0.1, 1e-5)
        new_buffers[pre + ".running_mean"] = rm
        new_buffers[pre + ".running_var"] = rv
        new_buffers[pre + ".num_batches_tracked"] = p[pre + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, p[pre + ".weight"], p[pre + ".bias"], False, 0.1, 1e-5)


def _basic_block(x, p, pre, stride, has_down, training, quantize, nb):
    """reference models/backbone.py:52-68."""
    identity = x
    out = _qa(F.conv2d(x, _qw(p[pre + ".conv1.weight"], quantize), None, stride, 1), quantize)
    out = _qa(F.relu(_bn(out, p, pre + ".bn1", training, quantize, nb)), quantize)
    out = _qa(F.conv2d(out, _qw(p[pre + ".conv2.weight"], quantize), None, 1, 1), quantize)
    out = _bn(out, p, pre + ".bn2", training, quantize, nb)
    if has_down:
        identity = _qa(F.conv2d(x, _qw(p[pre + ".downsample.0.weight"], quantize), None, stride, 0), quantize)
        identity = _qa(_bn(identity, p, pre + ".downsample.1", training, quantize, nb), quantize)
    return _qa(F.relu(out + identity), quantize)


def resnet18_features(x, p, pre, modality, training=True, quantize=None, new_buffers=None):
    """reference models/backbone.py:160-201 — returns the layer4 map."""
    nb = new_buffers if new_buffers is not None else {}
    if modality == "visual":
        B, C, T, H, W = x.shape
        x = x.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, H, W)
    x = _qa(x, quantize)
    x = _qa(F.conv2d(x, _qw(p[pre + ".conv1.weight"], quantize), None, 2, 3), quantize)
    x = _qa(F.relu(_bn(x, p, pre + ".bn1", training, quantize, nb)), quantize)
    x = F.max_pool2d(x, 3, 2, 1)
    inplanes = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            stride = 2 if (li > 1 and bi == 0) else 1
            has_down = stride != 1 or inplanes != planes
            x = _basic_block(x, p, "%s.layer%d.%d" % (pre, li, bi), stride, has_down, training, quantize, nb)
            inplanes = planes
    return x


def fusion_forward(fusion, p, a, v, detach_head=False):
    """The *_DGL heads (reference models/fusion_modules.py).  Returns (x_out, y_out, out).
    With detach_head=True the unimodal logits are computed with detached head parameters
    (the restatement of the gradient wipe at main_dgl.py:114-119); `out` always uses live
    head parameters and detached features, as in the reference."""
    f = "fusion_module."
    hp = (lambda k: p[f + k].detach()) if detach_head else (lambda k: p[f + k])
    lp = lambda k: p[f + k]
    if fusion == "concat":
        out = F.linear(torch.cat((a, v), 1).detach(), lp("fc_out.weight"), lp("fc_out.bias"))
        x_out = F.linear(torch.cat((a, torch.zeros_like(v)), 1), hp("fc_out.weight"), hp("fc_out.bias"))
        y_out = F.linear(torch.cat((torch.zeros_like(a), v), 1), hp("fc_out.weight"), hp("fc_out.bias"))
    elif fusion == "sum":
        x_out = F.linear(a, hp("fc_x.weight"), hp("fc_x.bias"))
        y_out = F.linear(v, hp("fc_y.weight"), hp("fc_y.bias"))
        out = F.linear(a.detach(), lp("fc_x.weight"), lp("fc_x.bias")) + \
            F.linear(v.detach(), lp("fc_y.weight"), lp("fc_y.bias"))
    elif fusion == "film":
        x = a.unsqueeze(2)
        y = v.unsqueeze(1)
        z = torch.bmm(x.detach(), y.detach()).flatten(1)
        out = F.linear(F.linear(z, lp("fc.weight"), lp("fc.bias")), lp("fc_out.weight"), lp("fc_out.bias"))
        zx = torch.bmm(x, x.transpose(2, 1)).flatten(1)
        x_out = F.linear(F.linear(zx, hp("fc.weight"), hp("fc.bias")), hp("fc_out.weight"), hp("fc_out.bias"))
        zy = torch.bmm(y.transpose(2, 1), y).flatten(1)
        y_out = F.linear(F.linear(zy, hp("fc.weight"), hp("fc.bias")), hp("fc_out.weight"), hp("fc_out.bias"))
    elif fusion == "gated":
        hx = F.linear(a, hp("fc_x.weight"), hp("fc_x.bias"))
        hy = F.linear(v, hp("fc_y.weight"), hp("fc_y.bias"))
        # Lf sees hx.detach()/hy.detach() (fusion_modules.py:235-236): fc_x/fc_y never get an Lf grad
        out = F.linear(torch.sigmoid(hx.detach()) * hy.detach(), lp("fc_out.weight"), lp("fc_out.bias"))
        x_out = F.linear(torch.sigmoid(hx) * hx, hp("fc_out.weight"), hp("fc_out.bias"))
        y_out = F.linear(torch.sigmoid(hy) * hy, hp("fc_out.weight"), hp("fc_out.bias"))
    else:
        raise NotImplementedError("Incorrect fusion method: {}!".format(fusion))
    return x_out, y_out, out


def model_forward(p, spec, image, fusion, training=True, quantize=None, detach_head=False,
                  new_buffers=None):
    """reference models/basic_model.py:65-86; spec [B,F,T] (unsqueezed here, main_dgl.py:100),
    image [B,3,T,H,W].  Returns (out, out_a, out_v) — the reference's return order."""
    a = resnet18_features(spec.unsqueeze(1).float(), p, "audio_net", "audio", training, quantize, new_buffers)
    v = resnet18_features(image.float(), p, "visual_net", "visual", training, quantize, new_buffers)
    B = a.shape[0]
    _, C, H, W = v.shape
    v = v.view(B, -1, C, H, W).permute(0, 2, 1, 3, 4)
    a = torch.flatten(F.adaptive_avg_pool2d(a, 1), 1)
    v = torch.flatten(F.adaptive_avg_pool3d(v, 1), 1)
    a_out, v_out, out = fusion_forward(fusion, p, a, v, detach_head)
    return out, a_out, v_out


def trainable_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def dgl_step(sd, momentum, spec, image, label, fusion="concat", alpha=4.0, lr=0.001, mu=0.9,
             wd=1e-4, max_norm=40.0, quantize=None, inv_batch=None, apply_update=True):
    """One DGL training step (reference main_dgl.py:93-158).  `sd` and `momentum` (dict, may be
    empty on the first step) are updated IN PLACE when apply_update.  Returns a dict with
    losses (Lf, La, Lv), logits (out, out_a, out_v), grads (clipped, like the reference after
    main_dgl.py:129), grad_norm, clip_coef, audio_grad_sum, visual_grad_sum."""
    names = trainable_names(sd)
    p = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    new_buffers = {}
    out, out_a, out_v = model_forward(p, spec, image, fusion, True, quantize, True, new_buffers)
    if inv_batch is None:
        Lv, La, Lf = F.cross_entropy(out_v, label), F.cross_entropy(out_a, label), F.cross_entropy(out, label)
    else:  # data-parallel shard: sum over local rows / global batch
        Lv, La, Lf = (F.cross_entropy(t, label, reduction="sum") * inv_batch for t in (out_v, out_a, out))
    total = (La + Lv) * alpha + Lf
    live = [k for k in names]
    grads = torch.autograd.grad(total, [p[k] for k in live], allow_unused=True)
    grads = {k: g for k, g in zip(live, grads) if g is not None}
    # clip_grad_norm_(max_norm=40, norm_type=2) over every parameter that has a gradient
    sq = 0.0  # python double; float() keeps this device-agnostic (tests may run the restatement on cuda in fp32)
    for g in grads.values():
        sq += float(g.double().pow(2).sum())
    norm = sq ** 0.5
    coef = min(1.0, max_norm / (norm + 1e-6))
    grads = {k: g * coef for k, g in grads.items()}
    a_sum = sum(float(g.abs().mean()) for k, g in grads.items() if k.startswith("audio_net."))
    v_sum = sum(float(g.abs().mean()) for k, g in grads.items() if k.startswith("visual_net."))
    if apply_update:
        with torch.no_grad():
            for k, g in grads.items():
                d = g + wd * sd[k]
                if k not in momentum:
                    momentum[k] = d.clone()
                else:
                    momentum[k].mul_(mu).add_(d)
                sd[k].add_(momentum[k], alpha=-lr)
            for k, v in new_buffers.items():
                sd[k] = v if torch.is_tensor(v) else torch.tensor(v)
    return {"losses": (float(Lf.detach()), float(La.detach()), float(Lv.detach())),
            "logits": (out.detach(), out_a.detach(), out_v.detach()),
            "grads": grads, "grad_norm": norm, "clip_coef": coef,
            "audio_grad_sum": a_sum, "visual_grad_sum": v_sum}
