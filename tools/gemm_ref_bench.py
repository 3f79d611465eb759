"""Library reference for the hot GEMM shapes: torch.matmul (cuBLAS) bf16 vs tcow_gemm_bf16, sustained loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcow_b200 import ops

d = torch.device('cuda')
M = 72008
shapes = [('qkv', 2304, 768, ops.EPI_BF16), ('fc1', 3072, 768, ops.EPI_BF16_GELU), ('fc2', 768, 3072, ops.EPI_F32_ADD),
          ('proj', 768, 768, ops.EPI_F32_ADD)]
for name, N, K, epi in shapes:
    a = torch.randn(M, K, device=d).to(torch.bfloat16)
    w = (torch.randn(N, K, device=d) * 0.02).to(torch.bfloat16)
    bias = torch.randn(N, device=d)
    out = torch.empty(M, N, device=d, dtype=torch.float32 if epi == ops.EPI_F32_ADD else torch.bfloat16)
    outb = torch.empty(M, N, device=d, dtype=torch.bfloat16)
    def ours():
        ops.gemm(a, w, bias, out, epi)
    def lib():
        torch.matmul(a, w.t(), out=outb)
    res = {}
    for tag, fn in (('ours', ours), ('cublas', lib)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[tag] = (ms, 2.0 * M * N * K / ms / 1e9)
    print(f"{name:5s} M={M} N={N} K={K}: ours {res['ours'][0]*1e3:7.1f} us {res['ours'][1]:7.1f} TFLOP/s (with epilogue) | "
          f"cuBLAS plain bf16 {res['cublas'][0]*1e3:7.1f} us {res['cublas'][1]:7.1f} TFLOP/s", flush=True)
