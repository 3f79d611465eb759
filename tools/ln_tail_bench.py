import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcow_b200 import ops
d = torch.device('cuda'); M = 72008; N = 768
for name, K in (('proj', 768), ('fc2', 3072)):
    a = torch.randn(M, K, device=d).to(torch.bfloat16); w = (torch.randn(N, K, device=d) * 0.02).to(torch.bfloat16)
    bias = torch.randn(N, device=d); x = torch.randn(M, N, device=d); out = torch.empty(M, N, device=d, dtype=torch.bfloat16)
    gm = torch.ones(N, device=d); bt = torch.zeros(N, device=d)
    fns = {'add': lambda: ops.gemm(a, w, bias, x, ops.EPI_F32_ADD),
           'add_ln0': lambda: ops.gemm_add_ln(a, w, bias, x, gm, bt, out, 0),
           'add_lnM': lambda: ops.gemm_add_ln(a, w, bias, x, gm, bt, out, M),
           'add+ln': lambda: (ops.gemm(a, w, bias, x, ops.EPI_F32_ADD), ops.layernorm(x, gm, bt, out))}
    for tag, fn in fns.items():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): fn()
        e1.record(); torch.cuda.synchronize()
        print(f'{name} {tag:8s} {e0.elapsed_time(e1) / 30 * 1e3:8.1f} us', flush=True)
