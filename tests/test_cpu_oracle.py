"""CPU suite: the oracle against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py), init parity of the drop-in modules, the C-ABI surface."""
import argparse
import glob
import os
import re

import pytest
import torch

from oracle import dgl_oracle as O
from oracle.synth import make_batch

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dgl_*.pt")))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_files_present():
    names = {os.path.basename(p) for p in GOLD}
    for f in ("concat", "sum", "gated", "film"):
        assert "dgl_%s_CREMAD.pt" % f in names
    assert "dgl_concat_KineticSound.pt" in names


@pytest.mark.parametrize("path", [p for p in GOLD if "film" not in p], ids=os.path.basename)
def test_oracle_matches_reference_golden(path):
    """Oracle == reference on the committed vectors (fp32, CPU): losses, diagnostics, per-tensor
    gradient norms, wiped gradients, eval logits after two steps."""
    g = torch.load(path)
    torch.set_num_threads(8)
    sd = O.init_state(g["fusion"], g["dataset"], 0)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - g["init_checksum"]) < 1e-6 * abs(g["init_checksum"])
    n = O.N_CLASSES[g["dataset"]]
    mom = {}
    for s in range(2):
        batch = make_batch(g["B"], n, g["shape"], seed=1 + s, label_max=31 if g["dataset"] == "KineticSound" else None)
        res = O.dgl_step(sd, mom, *batch, fusion=g["fusion"], alpha=g["alpha"], lr=g["lr"])
        ltol = 1e-6 if s == 0 else 2e-5
        for a, b in zip(res["losses"], g["losses"][s]):
            assert abs(a - b) <= ltol * max(1.0, abs(b))
        if s == 0:
            assert abs(res["audio_grad_sum"] - g["diag"][0][0]) <= 1e-5 * g["diag"][0][0]
            assert abs(res["visual_grad_sum"] - g["diag"][0][1]) <= 1e-5 * g["diag"][0][1]
            assert set(g["none_grads"]).isdisjoint(res["grads"].keys())
            assert set(g["grad_l2"][0].keys()) == set(res["grads"].keys())
            for k, v in g["grad_l2"][0].items():
                assert abs(float(res["grads"][k].norm()) - v) <= 1e-4 * v + 1e-9, k
            for k, v in g["small_grads"].items():
                assert torch.allclose(res["grads"][k], v, rtol=1e-4, atol=1e-6 * float(v.abs().max())), k
    # eval-mode logits of the updated model (running statistics included)
    spec, image, _ = make_batch(g["B"], n, g["shape"], seed=1, label_max=31 if g["dataset"] == "KineticSound" else None)
    with torch.no_grad():
        out, oa, ov = O.model_forward(sd, spec, image, g["fusion"], training=False)
    for got, want in zip((out, oa, ov), g["eval_logits"]):
        assert torch.allclose(got, want, rtol=2e-2, atol=2e-2)  # step-1 trajectories drift (make_golden.py)


def test_oracle_quirks():
    """fc_auxi and gated fc_x/fc_y never get a gradient; grads of the
    encoders come from the unimodal losses only, of the head from Lf only."""
    torch.set_num_threads(8)
    for fusion, dead in (("concat", ["fusion_module.fc_auxi.weight", "fusion_module.fc_auxi.bias"]),
                         ("gated", ["fusion_module.fc_x.weight", "fusion_module.fc_y.bias"])):
        sd = O.init_state(fusion, "CREMAD", 0)
        batch = make_batch(2, 6, "tiny", seed=3)
        before = {k: sd[k].clone() for k in dead}
        res = O.dgl_step(sd, {}, *batch, fusion=fusion, alpha=4.0, lr=0.1)
        for k in dead:
            assert k not in res["grads"]
            assert torch.equal(sd[k], before[k])  # skipped entirely: no weight decay, no momentum
    # alpha scales only the encoder gradients
    sd = O.init_state("concat", "CREMAD", 0)
    batch = make_batch(2, 6, "tiny", seed=3)
    r1 = O.dgl_step(sd, {}, *batch, fusion="concat", alpha=1.0, max_norm=1e9, apply_update=False)
    r2 = O.dgl_step(sd, {}, *batch, fusion="concat", alpha=2.0, max_norm=1e9, apply_update=False)
    k = "audio_net.conv1.weight"
    assert torch.allclose(r2["grads"][k], 2 * r1["grads"][k], rtol=1e-5, atol=1e-8)
    k = "fusion_module.fc_out.weight"
    assert torch.allclose(r2["grads"][k], r1["grads"][k], rtol=1e-6, atol=1e-9)


def test_oracle_bf16_mode_close_to_fp32():
    torch.set_num_threads(8)
    sd = O.init_state("concat", "CREMAD", 0)
    batch = make_batch(4, 6, "tiny", seed=1)
    r32 = O.dgl_step(sd, {}, *batch, fusion="concat", apply_update=False)
    rq = O.dgl_step(sd, {}, *batch, fusion="concat", apply_update=False, quantize="bf16")
    for a, b in zip(r32["losses"], rq["losses"]):
        assert abs(a - b) <= 3e-2 * abs(a)


@pytest.mark.parametrize("fusion", ["concat", "sum", "gated"])
def test_dropin_init_matches_reference_order(fusion):
    """gdl_b200.AVClassifier_DGL consumes the init RNG exactly like the reference (same names,
    shapes, order): state_dict == oracle.init_state, which make_golden.py proved == reference."""
    import gdl_b200
    args = argparse.Namespace(dataset="CREMAD", fusion_method=fusion, modality="full")
    gdl_b200.setup_seed(0)
    m = gdl_b200.AVClassifier_DGL(args)
    m.apply(gdl_b200.weight_init)
    sd, msd = O.init_state(fusion, "CREMAD", 0), m.state_dict()
    assert list(sd.keys()) == list(msd.keys())
    assert all(torch.equal(sd[k], msd[k]) for k in sd)
    assert hasattr(m, "args") and m.modality == "full"


def test_error_conventions():
    """reference basic_model.py:26,40 messages."""
    import gdl_b200
    with pytest.raises(NotImplementedError, match="Incorrect dataset name"):
        gdl_b200.AVClassifier_DGL(argparse.Namespace(dataset="nope", fusion_method="concat", modality="full"))
    with pytest.raises(NotImplementedError, match="Incorrect fusion method"):
        gdl_b200.AVClassifier_DGL(argparse.Namespace(dataset="CREMAD", fusion_method="nope", modality="full"))
    with pytest.raises(NotImplementedError):
        O.init_state("nope", "CREMAD")


def test_cabi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/gdl_b200.h declares."""
    from gdl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gdl_b200.h")).read()
    declared = set(re.findall(r"\b(gdl_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gdl_status"}
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
        assert name in _lib.SIGNATURES, "no ctypes signature for " + name
    assert set(_lib.SIGNATURES) == declared
    assert lib.gdl_version() >= 100


def test_cabi_host_side_validation():
    """Pure host entry points work without a GPU; bad descriptors are rejected with GDL_EINVAL."""
    import ctypes as C
    from gdl_b200 import _lib, ops
    lib = _lib.load()
    d = ops.conv_desc(4, 56, 56, 64, 64, 3, 3, 1, 1)
    assert (d.Ho, d.Wo) == (56, 56)
    assert ops.conv_packed_k(d) == 576
    assert ops.conv_wgrad_workspace_bytes(d) > 0
    stem = ops.conv_desc(2, 224, 224, 8, 64, 7, 7, 2, 3)
    assert (stem.Ho, stem.Wo) == (112, 112) and ops.conv_packed_k(stem) == 448
    bad = ops.conv_desc(4, 56, 56, 48, 64, 3, 3, 1, 1)  # 48 channels: not storable
    assert lib.gdl_conv_packed_k(C.byref(bad)) == _lib.GDL_EINVAL
    assert lib.gdl_conv_fwd(C.byref(bad), None, None, None, None) == _lib.GDL_EINVAL
    assert b"bad descriptor" in lib.gdl_last_error_string()
    assert ops.bn_partial_floats(1000, 64) > 0 and ops.head_scratch_floats(8, 6) == 8 * 6 + 24


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "iccv2025-gdl_b200", "gdl_b200")
    for f in glob.glob(os.path.join(pkg, "*.py")):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
        assert "dgl_oracle" not in src and "/root/reference" not in src, f
