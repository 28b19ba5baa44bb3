import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/torch-asg_b200")
import torch, torch_asg
T, N, L, S = 1000, 64, 40, 200
g = torch.Generator().manual_seed(0)
xs = [torch.randn(T, N, L, generator=g).log_softmax(-1).cuda().requires_grad_() for _ in range(6)]
y = torch.randint(0, L, (N, S), generator=g).cuda()
crit = torch_asg.ASGLoss(L).cuda()
for i in range(8):
    if i == 7: os.environ["ASG_GRAD_DRY"] = "2"
    x = xs[i % 6]; x.grad = None; crit.transition.grad = None
    crit(x, y).backward()
torch.cuda.synchronize()
