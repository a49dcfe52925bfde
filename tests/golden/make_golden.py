"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from
/root/reference with stubs for the absent timm/librosa/skimage) on seeded
synthetic inputs, and pins oracle/dgl_oracle.py against it on the way.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
"""
import argparse
import csv
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = nn.Identity
    layers.to_2tuple = lambda x: (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm.models = timm_models
    timm_models.layers = layers
    for name, mod in (("timm", timm), ("timm.models", timm_models), ("timm.models.layers", layers),
                      ("librosa", types.ModuleType("librosa")), ("skimage", types.ModuleType("skimage"))):
        sys.modules.setdefault(name, mod)
    import main_dgl  # noqa
    return main_dgl


def run_reference(main_dgl, fusion, dataset, batches, alpha, lr, seed=0):
    """Two calls of the reference's train_epoch, one batch each; returns per-step records."""
    from utils.utils import setup_seed, weight_init
    from models.basic_model import AVClassifier_DGL
    args = argparse.Namespace(dataset=dataset, fusion_method=fusion, modality="full", alpha=alpha,
                              epochs=1, batch_size=batches[0][0].shape[0], drop=0)
    setup_seed(seed)
    model = AVClassifier_DGL(args)
    model.apply(weight_init)
    init_sd = {k: v.clone() for k, v in model.state_dict().items()}
    dp = nn.DataParallel(model, device_ids=[])
    opt = torch.optim.SGD(dp.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    recs = []
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for step, batch in enumerate(batches):
                devnull = open(os.devnull, "w")
                old = sys.stdout, sys.stderr
                sys.stdout, sys.stderr = devnull, devnull
                try:
                    out = main_dgl.train_epoch(args, step + 30, dp, torch.device("cpu"), [batch], opt, None)
                finally:
                    sys.stdout, sys.stderr = old
                rows = list(csv.reader(open("audio_visual_grad_vanilla.csv")))
                grads = {k: (p.grad.clone() if p.grad is not None else None)
                         for k, p in model.named_parameters()}
                recs.append({"losses": out[:3], "diag": [float(x) for x in rows[-1]], "grads": grads,
                             "state": {k: v.clone() for k, v in model.state_dict().items()}})
        finally:
            os.chdir(cwd)
    # logits of the updated model in eval mode on the first batch (pins running stats too)
    model.eval()
    with torch.no_grad():
        spec, image, _ = batches[0]
        ev = model(spec.unsqueeze(1).float(), image.float())
    return init_sd, recs, [t.clone() for t in ev]


def main():
    main_dgl = import_reference()
    from oracle import dgl_oracle as O
    from oracle.synth import make_batch
    torch.set_num_threads(8)
    alpha, lr = 4.0, 0.01
    cases = [("concat", "CREMAD"), ("sum", "CREMAD"), ("gated", "CREMAD"), ("film", "CREMAD"),
             ("concat", "KineticSound")]
    for fusion, dataset in cases:
        n = O.N_CLASSES[dataset]
        batches = [make_batch(4, n, "tiny", seed=1 + s, label_max=31 if dataset == "KineticSound" else None)
                   for s in range(2)]
        init_sd, recs, ev = run_reference(main_dgl, fusion, dataset, batches, alpha, lr)

        # ---- pin the oracle: identical init, then step-by-step agreement with the reference ----
        sd = O.init_state(fusion, dataset, seed=0)
        assert list(sd.keys()) == list(init_sd.keys()), "state_dict key order differs"
        for k in sd:
            assert torch.equal(sd[k], init_sd[k]), "init differs at " + k
        mom = {}
        worst = 0.0
        for step, (batch, rec) in enumerate(zip(batches, recs)):
            res = O.dgl_step(sd, mom, *batch, fusion=fusion, alpha=alpha, lr=lr)
            for a, b in zip(res["losses"], rec["losses"]):
                assert abs(a - b) <= (1e-6 if step == 0 else (5e-3 if fusion == "film" else 2e-5)) * max(1, abs(b)), (fusion, step, res["losses"], rec["losses"])
            dtol = 1e-5 if step == 0 else 2e-2  # step 1: ulp-level parameter drift, tiny-batch BN
            if fusion == "film":
                # torch's CPU fp32 clip-norm over the 134M-element fc gradient is ~1.4e-3 low
                # (the reference's clipped norm comes out as 40.06, not 40); the oracle sums in double
                dtol = max(dtol, 3e-3)
            da_ = abs(res["audio_grad_sum"] - rec["diag"][0]) / abs(rec["diag"][0])
            dv_ = abs(res["visual_grad_sum"] - rec["diag"][1]) / abs(rec["diag"][1])
            print("  step %d diag rel diff audio %.2e visual %.2e" % (step, da_, dv_))
            assert da_ <= dtol and dv_ <= dtol
            for k, g in rec["grads"].items():
                if g is None:
                    assert k not in res["grads"], "oracle produced a grad the reference wiped: " + k
                else:
                    d = (res["grads"][k] - g).abs().max().item() / (g.abs().max().item() + 1e-12)
                    worst = max(worst, d)
                    # step 0 starts from bit-identical parameters: only summation-order noise.
                    # step 1 starts from parameters that already differ in the last ulp (clip
                    # coefficient computed in double here), amplified through 20 BN layers.
                    if step == 0:
                        assert d < (3e-3 if fusion == "film" else 1e-4), (fusion, step, k, d)
                    else:  # ReLU/max-pool selections may flip under ulp-level drift: compare direction
                        cos = torch.nn.functional.cosine_similarity(res["grads"][k].flatten().double(),
                                                                    g.flatten().double(), dim=0).item()
                        assert cos > 0.99, (fusion, step, k, cos)
            for k, v in rec["state"].items():
                if step == 0 and fusion != "film":
                    assert torch.allclose(sd[k].float(), v.float(), atol=2e-6, rtol=1e-4), (fusion, step, k)
                else:  # drifted trajectories: whole-tensor relative error only
                    e = (sd[k].double() - v.double()).norm() / (v.double().norm() + 1e-12)
                    amax = (sd[k].double() - v.double()).abs().max().item()
                    # film at step 1 is in an exploding regime (loss ~31): state is not compared there
                    assert e < 2e-3 or amax < 1e-3 or (fusion == "film" and step == 1), (fusion, step, k, float(e), amax)
            print("  step %d worst grad rel-max diff %.2e" % (step, worst))
        print("oracle == reference for %s/%s (worst grad rel-max diff %.2e)" % (fusion, dataset, worst))

        # ---- golden vectors (small): losses, diagnostics, logits, per-tensor grad summaries ----
        gold = {"fusion": fusion, "dataset": dataset, "alpha": alpha, "lr": lr, "shape": "tiny", "B": 4,
                "init_checksum": float(sum(v.double().sum() for v in init_sd.values())),
                "losses": [[float(x) for x in r["losses"]] for r in recs],
                "diag": [r["diag"] for r in recs],
                "eval_logits": ev,
                "none_grads": [k for k, g in recs[0]["grads"].items() if g is None],
                "grad_l2": [{k: float(g.norm()) for k, g in r["grads"].items() if g is not None} for r in recs],
                "grad_absmean": [{k: float(g.abs().mean()) for k, g in r["grads"].items() if g is not None}
                                 for r in recs],
                "param_l2_after": {k: float(v.double().norm()) for k, v in recs[-1]["state"].items()},
                "small_grads": {k: g for k, g in recs[0]["grads"].items()
                                if g is not None and g.numel() <= 8192 and
                                (k.startswith("fusion_module") or k.endswith("bn1.weight") or "layer4.1.bn2" in k)},
                "small_params_after": {k: v for k, v in recs[-1]["state"].items()
                                       if v.numel() <= 1024 and ("layer1.0.bn1" in k or k.startswith("fusion_module"))}}
        torch.save(gold, os.path.join(HERE, "dgl_%s_%s.pt" % (fusion, dataset)))
        print("wrote golden for", fusion, dataset)


if __name__ == "__main__":
    main()
