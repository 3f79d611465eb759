"""Launch one hot-path op a few times on north-star shapes (B=8, T=30, N=300) — a short target for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tcow_b200 import ops  # noqa: E402

op = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, N, T, D, H = 8, 300, 30, 768, 12
M = B * N * T
d = torch.device('cuda')
g = torch.Generator(device=d).manual_seed(0)
qkv = (torch.randn(M + B, 3 * D, device=d, generator=g)).to(torch.bfloat16)
out = torch.empty(M + B, D, device=d, dtype=torch.bfloat16)
ocls = torch.empty(B, T, D, device=d)
x = torch.randn(M + B, D, device=d, generator=g)
wq = (torch.randn(3 * D, D, device=d, generator=g) * 0.03).to(torch.bfloat16)
bq = torch.randn(3 * D, device=d, generator=g) * 0.1
a_ = x[:M].to(torch.bfloat16)
for _ in range(reps):
    if op == 'spatial':
        ops.attn_spatial(qkv, out, ocls, B, N, T, H, True, M)
    elif op == 'temporal':
        ops.attn_temporal(qkv, out, B * N, T, H, 0)
    elif op == 'ln':
        gm = torch.ones(D, device=d); bt = torch.zeros(D, device=d)
        ops.layernorm(x, gm, bt, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for _ in range(ITERS):
    if op == 'spatial':
        ops.attn_spatial(qkv, out, ocls, B, N, T, H, True, M)
    elif op == 'temporal':
        ops.attn_temporal(qkv, out, B * N, T, H, 0)
e1.record()
torch.cuda.synchronize()
print(op, os.path.basename(os.environ.get('TCOW_B200_LIB', 'default')), 'avg us', e0.elapsed_time(e1) * 1000 / ITERS)
