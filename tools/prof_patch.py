"""Isolated timing of the patch embedding: one fused kernel vs gather + embed_init + reduce-add GEMM, fp32 and uint8 inputs
(B=8, T=30, 240x320, 3 queries per video)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcow_b200 import ops
from tcow_b200.ops import EPI_F32_ADD
d = torch.device('cuda')
g = torch.Generator(device=d).manual_seed(0)
B, T, Hf, Wf, P, D, Q = 8, 30, 240, 320, 16, 768, 3
N = (Hf // P) * (Wf // P); M = B * N * T
V = (B + Q - 1) // Q
fr32 = torch.rand(V, 3, T, Hf, Wf, device=d, generator=g)
q32 = (torch.rand(B, 1, T, Hf, Wf, device=d, generator=g) > 0.9).float()
fr8, q8 = (fr32 * 255).round().to(torch.uint8), q32.to(torch.uint8)
w = (torch.randn(D, 4 * P * P, device=d, generator=g) * 0.03).to(torch.bfloat16)
cb, pos, tim, cls = (torch.randn(s, device=d, generator=g) * 0.1 for s in [(D,), (N + 1, D), (T, D), (D,)])
X = torch.empty(M + B, D, device=d)
PM = torch.empty(M, 4 * P * P, device=d, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device=d, dtype=torch.uint8)

def fused(fr, q, sc): ops.patch_embed_fused(fr, q, w, cb, pos, tim, cls, X, P, False, Q, 0, sc)
def three(fr, q, sc):
    ops.patch_gather(fr, q, PM, P, False, Q, 0, sc)
    ops.embed_init(X, cb, pos, tim, cls, B, N, T, D)
    ops.gemm(PM, w, None, X[:M], EPI_F32_ADD)
for name, fn, args in (('fused fp32', fused, (fr32, q32, 1.0)), ('three fp32', three, (fr32, q32, 1.0)),
                       ('fused u8', fused, (fr8, q8, 1 / 255)), ('three u8', three, (fr8, q8, 1 / 255))):
    for _ in range(3): fn(*args)
    tot = 0.0
    for _ in range(20):
        flush.zero_()                       # cold L2, as inside the step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*args); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f'{name:12s} {tot / 20 * 1000:8.1f} us')
