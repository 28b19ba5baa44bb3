for p in _ab/consonly _ab/gathonly; do for T in 1000 2000; do
echo "== $p T=$T $(PYTHONPATH=$PWD/$p:$PWD python - $T <<'PY'
import sys, torch
import torch_asg_native as nat
T = int(sys.argv[1]); N, L, S = 64, 40, 200
g = torch.Generator().manual_seed(0)
x = torch.randn(T, N, L, generator=g).log_softmax(-1).cuda(); A = (0.1 * torch.randn(L, L, generator=g)).cuda()
y = torch.randint(0, L, (N, S), generator=g).cuda()
il = torch.full((N,), T, dtype=torch.long, device="cuda"); tl = torch.full((N,), S, dtype=torch.long, device="cuda")
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print("fcc %.1f us  fac %.1f us" % (timeit(lambda: nat.fully_connected_forward(x, A, il, T, N, L)), timeit(lambda: nat.force_aligned_forward(x, y, A, il, tl, T, N, L, S))))
PY
)"
done; done
