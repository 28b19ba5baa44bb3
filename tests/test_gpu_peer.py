"""GPU tier: the exchange step of the batch-sharded path over NVLink peer memory (C ABI asg_peer_*, csrc/asg_peer.cu).

One GPU: world = 1 runs the same kernel (push to itself, flag, sum of one slot) -- identity, also inside a CUDA graph, and the
argument checks.  Two or more GPUs: tests/multi/peer_exchange.py under torchrun (peer vs NCCL, bit-identical across ranks, graph
replays, the sharded GraphedASGLoss.step against the whole batch)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_world_of_one_is_identity_and_replays_in_a_graph():
    import torch_asg_native as nat
    dev = torch.device("cuda", 0)
    comm = nat.PeerComm(0, 1, 1 << 14, 0)
    assert len(comm.handle()) == 64
    for dtype in (torch.float32, torch.float64):
        v = torch.randn(1602, device=dev, dtype=dtype)
        w = v.clone()
        comm.allreduce(w)
        assert torch.equal(v, w)
    src, buf = torch.zeros(1602, device=dev), torch.zeros(1602, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        buf.copy_(src); comm.allreduce(buf)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        buf.copy_(src)
        comm.allreduce(buf)
    for it in range(5):
        src.fill_(float(it + 1))
        g.replay()
        assert torch.equal(buf, src)
    epoch, timed_out = comm.status()
    assert epoch == 2 + 1 + 5 and timed_out == 0
    with pytest.raises(RuntimeError, match="do not fit"):
        comm.allreduce(torch.zeros(1 << 14, device=dev))
    with pytest.raises(RuntimeError):
        nat.PeerComm(3, 2, 1024, 0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_two_ranks_against_nccl_and_the_whole_batch():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi", "peer_exchange.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
