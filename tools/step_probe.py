"""GPU probe: run DGLStep vs the CPU oracle (fp32 and bf16-quantised) and print per-tensor parity."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))

import gdl_b200  # noqa: E402
from gdl_b200.step import DGLStep  # noqa: E402
from oracle import dgl_oracle as O  # noqa: E402
from oracle.synth import SHAPES, make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fusion", default="concat")
    ap.add_argument("--dataset", default="CREMAD")
    ap.add_argument("--shape", default="tiny")
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/step_probe.json")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    n = O.N_CLASSES[a.dataset]
    Fq, Tt, T, H, W = SHAPES[a.shape]
    args = argparse.Namespace(dataset=a.dataset, fusion_method=a.fusion, modality="full")
    gdl_b200.setup_seed(0)
    model = gdl_b200.AVClassifier_DGL(args)
    model.apply(gdl_b200.weight_init)
    model.cuda().train()
    step = DGLStep(model, a.B, (Fq, Tt), (T, H, W), alpha=4.0, lr=a.lr, use_graph=bool(a.graph))
    sd32 = O.init_state(a.fusion, a.dataset, 0)
    sdq = {k: v.clone() for k, v in sd32.items()}
    m32, mq = {}, {}
    report = []
    for s in range(a.steps):
        batch = make_batch(a.B, n, a.shape, seed=1 + s)
        t0 = time.time()
        step.step(*[t.cuda() for t in batch])
        torch.cuda.synchronize()
        got = step.read_stats()
        t1 = time.time()
        r32 = O.dgl_step(sd32, m32, *batch, fusion=a.fusion, alpha=4.0, lr=a.lr)
        rq = O.dgl_step(sdq, mq, *batch, fusion=a.fusion, alpha=4.0, lr=a.lr, quantize="bf16")
        t2 = time.time()
        rec = {"step": s, "gpu_s": t1 - t0, "oracle_s": t2 - t1,
               "losses_gpu": got[:3], "losses_fp32": r32["losses"], "losses_q": rq["losses"],
               "norm": (got[3], r32["grad_norm"], rq["grad_norm"]),
               "diag_gpu": got[5:], "diag_fp32": (r32["audio_grad_sum"], r32["visual_grad_sum"]),
               "diag_q": (rq["audio_grad_sum"], rq["visual_grad_sum"])}
        names = dict(model.named_parameters())
        worst32, worstq = (1.0, None), (1.0, None)
        cos_list = {}
        for k, g32 in r32["grads"].items():
            gg = names[k].grad.detach().float().cpu().flatten().double()
            c32 = torch.nn.functional.cosine_similarity(gg, g32.flatten().double(), dim=0).item()
            cq = torch.nn.functional.cosine_similarity(gg, rq["grads"][k].flatten().double(), dim=0).item()
            ratio = (gg.norm() / (g32.double().norm() + 1e-30)).item()
            cqo = torch.nn.functional.cosine_similarity(rq["grads"][k].flatten().double(),
                                                        g32.flatten().double(), dim=0).item()
            cos_list[k] = (round(c32, 5), round(cq, 5), round(ratio, 4), round(cqo, 5))
            if c32 < worst32[0]:
                worst32 = (c32, k)
            if cq < worstq[0]:
                worstq = (cq, k)
        rec["worst_cos_fp32"], rec["worst_cos_q"] = worst32, worstq
        rec["cos"] = cos_list
        agree = [(step.logits[i].argmax(1).cpu() == r32["logits"][i].argmax(1)).float().mean().item()
                 for i in range(3)]
        rec["argmax_agree"] = agree
        perr = max(((names[k].detach().float().cpu() - sd32[k]).norm() / (sd32[k].norm() + 1e-12)).item()
                   for k in r32["grads"])
        rec["worst_param_rel_err_fp32"] = perr
        report.append(rec)
        print(json.dumps({k: v for k, v in rec.items() if k != "cos"}))
        low = sorted(cos_list.items(), key=lambda kv: kv[1][0])[:6]
        print("  lowest cos:", low)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
