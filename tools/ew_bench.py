"""Stem-tail / BN kernel timings at the bench geometry (visual stem: N=768, 112x112x64): python tools/ew_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iccv2025-gdl_b200"))
from gdl_b200 import ops
ops.init()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
N, H, W, C = int(os.environ.get("EW_N", 768)), 112, 112, 64
Ho, Wo = 56, 56
P = N * H * W
x = (torch.randn(N, H, W, C, device="cuda") * 2 + 0.3).to(torch.bfloat16)
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
partial = torch.empty(ops.bn_partial_floats(P, C), device="cuda")
mean, invstd, scale, shift = (torch.empty(C, device="cuda") for _ in range(4))
ops.bn_stats(x, P, C, partial, gamma, beta, 1e-5, 0.1, None, None, mean, invstd, scale, shift)
y = torch.empty(N, Ho, Wo, C, device="cuda", dtype=torch.bfloat16)
am = torch.empty(N, Ho, Wo, C, device="cuda", dtype=torch.uint8)
gp = torch.randn(N, Ho, Wo, C, device="cuda").to(torch.bfloat16)
xm = torch.empty_like(gp)
dx = torch.empty_like(x)
dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
GB = 1e-6
t = timeit(lambda: ops.bn_stats(x, P, C, partial, gamma, beta, 1e-5, 0.1, None, None, mean, invstd, scale, shift))
print("bn_stats           %.3f ms  %.0f GB/s" % (t, 2.0 * P * C * GB / t))
t = timeit(lambda: ops.bn_relu_maxpool_fwd(x, scale, shift, y, am, xm, N, H, W, C, Ho, Wo))
print("stem_tail_fwd      %.3f ms  %.0f GB/s" % (t, (2.0 * P * C + 3.0 * N * Ho * Wo * C) * GB / t))
t = timeit(lambda: ops.bn_relu_maxpool_bwd(gp, am, xm, x, dx, N, H, W, C, Ho, Wo, gamma, mean, invstd, scale, shift, partial, dg, db))
print("stem_tail_bwd      %.3f ms  %.0f GB/s" % (t, (6.0 * P * C + 6.0 * N * Ho * Wo * C) * GB / t))
