"""Developer / evidence: one forward+backward of a small case through the product's public route, for compute-sanitizer.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py T N L S [f64]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "torch-asg_b200")]
import torch
import torch_asg
T, N, L, S = [int(v) for v in sys.argv[1:5]]
dtype = torch.float64 if len(sys.argv) > 5 and sys.argv[5] == "f64" else torch.float32
g = torch.Generator().manual_seed(0)
x = torch.randn(T, N, L, generator=g, dtype=dtype).log_softmax(-1).cuda().requires_grad_(True)
y = torch.randint(0, L, (N, S), generator=g).cuda()
il = torch.randint(max(S, T // 2), T + 1, (N,), generator=g).cuda()
tl = torch.randint(1, S + 1, (N,), generator=g).cuda()
crit = torch_asg.ASGLoss(L).cuda().to(dtype)
loss = crit(x, y, il, tl)
loss.backward()
ws = torch_asg.GraphedASGLoss(crit, use_graph=False)
l2 = ws.step(x, y, il, tl)
res = torch_asg.forced_align(x.detach(), y, crit.transition.detach(), il, tl)
sc = res[0]
torch.cuda.synchronize()
print("loss", float(loss), float(l2), "flags", ws.check().tolist(), "viterbi", float(sc.sum()))
