"""Stem kernel timings at the bench geometry: python tools/stem_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
from gdl_b200 import ops
ops.init()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for name, B, C, T, H, W in (("visual", 256, 3, 3, 224, 224), ("audio", 256, 1, 1, 257, 188)):
    N = B * T
    Ho, Wo, Hp, Wp = ops.stem_geometry(H, W)
    src = torch.randn(B, C, T, H, W, device="cuda")
    x16 = torch.empty(N, Hp, Wp, 16, device="cuda", dtype=torch.bfloat16)
    ops.stem_layout(src, x16, B, C, T, H, W)
    w = torch.randn(64, C, 7, 7, device="cuda") * 0.1
    wp = torch.empty(64, 256, device="cuda", dtype=torch.bfloat16)
    ops.stem_pack_weights(w, wp, C)
    y = torch.empty(N, Ho, Wo, 64, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(N, Ho, Wo, 64, device="cuda").to(torch.bfloat16)
    ws = torch.empty(ops.stem_wgrad_workspace_bytes(N, H, W) // 4, device="cuda")
    dw = torch.empty_like(w)
    print(name, "layout %.3f fwd %.3f wgrad %.3f ms" % (timeit(lambda: ops.stem_layout(src, x16, B, C, T, H, W)),
          timeit(lambda: ops.stem_fwd(x16, wp, y, N, H, W, C)),
          timeit(lambda: ops.stem_wgrad(x16, dy, dw, C, N, H, W, ws))), flush=True)
