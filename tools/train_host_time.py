import logging, os, sys, time
sys.path.insert(0, os.getcwd())
import torch, tcow_b200
from tcow_b200 import synth
T, Hf, Wf, V, Q = 30, 240, 320, 2, 3
dev = 'cuda:0'
net = tcow_b200.Seeker(logging.getLogger('p'), num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                       tracker_pretrained=False, causal_attention=1, patch_size=16, drop_path_rate=0.1)
net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=Hf, frame_width=Wf))
net = net.to(dev).train()
rgb = torch.rand(V, 3, T, Hf, Wf, device=dev)
q = torch.zeros(V, Q, 1, T, Hf, Wf, device=dev); q[:, :, 0, 0, 10:50, 20:60] = 1
tm = (torch.rand(V * Q, 3, T, Hf, Wf, device=dev) > 0.7).float()
tf = (torch.rand(V * Q, T, 3, device=dev) > 0.5).float()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    mask, flags = net.forward_queries(rgb, q)
    t1 = time.perf_counter()
    loss = synth.training_loss(mask.flatten(0, 1), flags.flatten(0, 1), tm, tf)
    loss.backward()
    t2 = time.perf_counter()
    opt.step()
    t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f'host: fwd issue {1e3*(t1-t0):.1f} ms, bwd issue {1e3*(t2-t1):.1f} ms, opt {1e3*(t3-t2):.1f}, total issue {1e3*(t3-t0):.1f} ms; step wall {1e3*(t4-t0):.1f} ms; launches {net.seeker.train_engine().launches}')
