"""The exchange step of a batch-sharded ASG job over NVLink peer memory (C ABI ``asg_peer_*``, csrc/asg_peer.cu).

The reference has no multi-GPU path: its users wrap ``ASGLoss`` in DistributedDataParallel, which all-reduces
``transition.grad`` through NCCL.  A batch-sharded step has exactly one exchange -- the sum over ranks of the fused vector
[L*L transition gradient | loss sum | utterance count], 6.4 KB at L = 40 -- so here it is one single-CTA kernel per rank that
stores the vector into every rank's exchange buffer, raises a flag there, waits for all flags locally and sums the slots in
rank order.  ``torch.distributed`` is only the plumbing: it carries the 64-byte CUDA IPC handles once, at construction.

``PeerAllReduce.create(group)`` returns ``None`` (on every rank, decided collectively) when the ranks are not all on one host
or some pair of GPUs has no peer access; the caller keeps ``dist.all_reduce`` then.
"""
import socket

import torch
import torch.distributed as dist


class PeerAllReduce:
    def __init__(self, comm, rank, world, max_bytes):
        self._comm = comm
        self.rank, self.world, self.max_bytes = rank, world, max_bytes

    @classmethod
    def create(cls, native, group=None, device=None, max_bytes=1 << 16):
        """Collective over ``group`` (every rank must call it).  ``native``: the B200 build of torch_asg_native."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8 or not hasattr(native, "PeerComm"):
            return None
        device = torch.cuda.current_device() if device is None else torch.device(device).index
        comm, handle, err = None, b"", ""
        try:
            comm = native.PeerComm(rank, world, max_bytes, device)
            handle = comm.handle()
        except RuntimeError as exc:             # noqa: PERF203
            err = str(exc)
        infos = [None] * world
        dist.all_gather_object(infos, (socket.gethostname(), device, handle, err), group=group)
        ok = all(i[2] and not i[3] and i[0] == infos[0][0] for i in infos) and len({i[1] for i in infos}) == world
        if ok and world > 1:
            try:
                comm.connect(b"".join(i[2] for i in infos))
            except RuntimeError:
                ok = False
        flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", device), dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return None
        return cls(comm, rank, world, max_bytes)

    def all_reduce(self, buf):
        """In-place sum of a contiguous float32 / float64 CUDA tensor over all ranks, on the current stream; one kernel."""
        self._comm.allreduce(buf)
        return buf

    def status(self):
        """(exchanges done, epoch of the first exchange that timed out or 0).  Synchronises with the device."""
        return self._comm.status()
