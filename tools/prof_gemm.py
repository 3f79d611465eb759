"""Launch one GEMM shape a few times (ncu target).  usage: prof_gemm.py N K epi [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcow_b200 import ops
N, K, epi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
M = 72008
d = torch.device('cuda')
a = torch.randn(M, K, device=d).to(torch.bfloat16)
w = (torch.randn(N, K, device=d) * 0.02).to(torch.bfloat16)
bias = torch.randn(N, device=d)
out = torch.empty(M, N, device=d, dtype=torch.float32 if epi >= 2 else torch.bfloat16)
for _ in range(reps):
    ops.gemm(a, w, bias, out, epi)
torch.cuda.synchronize()
