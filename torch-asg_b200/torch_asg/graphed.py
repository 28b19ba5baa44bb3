"""``GraphedASGLoss`` -- the reference's ``ASGLoss(...)(inputs, ...)`` + ``loss.backward()`` with a persistent
workspace and CUDA graphs behind it, so that a training step costs the host a few tens of microseconds
instead of the ~0.2-0.7 ms the eager route needs (two pybind calls, ~8 allocations, a handful of tiny
reduction kernels, the autograd nodes between them and -- batch-sharded -- a Python ``all_reduce``).

    crit  = torch_asg.ASGLoss(L, reduction='mean').cuda()          # the reference's module, unchanged
    gcrit = torch_asg.GraphedASGLoss(crit)                          # same call signature, same gradients
    loss  = gcrit(inputs, targets, input_lengths, target_lengths)   # replays the forward graph
    loss.backward()                                                 # replays the backward graph
    gcrit.step(inputs, targets, input_lengths, target_lengths)      # or: forward+backward(+all-reduce) as ONE graph

What runs is the C ABI's ``asg_forward`` / ``asg_backward`` / ``asg_step`` (include/asg_b200.h) on the buffers of a
``torch_asg_native.Workspace``: no allocation, no host synchronisation, 3-5 kernel launches -- which is what makes
the call capturable.  Graphs are keyed by the addresses they were captured with (inputs, targets, lengths,
transition); a new address captures a new graph (a small LRU), ``use_graph=False`` runs the same zero-allocation
calls eagerly.

Batch-sharded jobs: pass ``group``; the backward graph then ends with the ONE exchange of the job, the sum
over ranks of the fused ``[L*L + 2]`` buffer (transition gradient, loss sum, utterance count), so ``transition.grad`` is
the global gradient on every rank and ``global_loss`` the global reduced loss.  The value ``forward`` returns is this
rank's share of the reduced loss (its sum over ranks is the global loss).  ``exchange``:
  'peer'  one kernel of this library per rank over NVLink peer memory (torch_asg/peer.py, C ABI ``asg_peer_*``): no
          library collective in the graph, the replay costs the host what a single-GPU replay costs;
  'nccl'  ``dist.all_reduce`` captured in the graph.  Set ``NCCL_GRAPH_MIXING_SUPPORT=0`` before ``init_process_group``
          then: with NCCL's default every launch of a graph that holds a collective waits on the host for the previous
          one (measured on 2 x B200: 0.10-0.25 ms per replay);
  'auto'  (default) 'peer' when all ranks sit on one host with peer access between every pair of GPUs, else 'nccl'.

Reference semantics kept: ``reduction`` 'mean' / 'sum' / anything else = none (asg.py:137-142), default lengths
(asg.py:113-117), ``forward_only`` / eval = the two beta scans only and no gradient (asg.py:129-131).
"""
import collections

import torch
import torch.distributed as dist
from torch import nn


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


class _Entry:
    """Graphs captured for one set of addresses."""
    __slots__ = ("ws", "fwd", "fwd_only", "bwd", "step", "gloss", "kernels")

    def __init__(self, ws):
        self.ws = ws
        self.fwd = self.fwd_only = self.bwd = self.step = None
        self.gloss = None
        self.kernels = {}


class _GraphedFn(torch.autograd.Function):
    """forward = the scans + the reduced loss, backward = the gradient assembly (+ the all-reduce)."""

    @staticmethod
    def forward(ctx, inputs, transition, owner, entry, targets, il, tl, reduce):
        owner._run(entry, "fwd", (inputs, transition, targets, il, tl))
        if reduce and owner.use_graph and entry.bwd is None:
            # capture the backward graph here, on the caller's thread (not lazily inside the autograd engine's thread);
            # its warm-up runs on the state the forward just wrote and only touches the gradient buffers
            entry.bwd = owner._capture(entry, "bwd", (None, None, targets, None, None))
        ctx.owner, ctx.entry, ctx.reduce, ctx.targets = owner, entry, reduce, targets
        ws = entry.ws
        return ws.loss_tail[2].detach() if reduce else ws.losses.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        owner, entry, ws = ctx.owner, ctx.entry, ctx.entry.ws
        L = ws.L
        if ctx.reduce:
            ws.gout.copy_(grad_out.reshape(1))
            owner._run(entry, "bwd", (None, None, ctx.targets, None, None))
        else:                               # reduction none: a per-utterance upstream gradient, eager (still allocation-free)
            g = grad_out.to(ws.grad_x.dtype).contiguous()
            ws.backward(g, 1.0, ctx.targets)
            owner._tail_and_reduce(ws)
        return ws.grad_x, ws.fused[:L * L].view(L, L), None, None, None, None, None, None


class GraphedASGLoss(nn.Module):
    def __init__(self, criterion, group=None, global_batch=None, use_graph=True, max_graphs=32, exchange="auto"):
        super().__init__()
        self.criterion = criterion
        self.group = group
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        self.exchange = exchange            # how the fused vector is summed over ranks; resolved on first use (collective)
        self._peer = None
        self._exchange_resolved = False
        self.global_batch = global_batch
        self.use_graph = use_graph
        self.max_graphs = max_graphs
        self._entries = collections.OrderedDict()
        self._ws = {}
        self._full_lengths = {}
        self._streams = {}
        self._global_n = {}
        self._hits = self._misses = 0
        self.graph_replays = 0
        self.kernel_launches = 0          # kernels of libasg_b200 launched through this object (graph replays included)
        self.nccl_in_graph = None
        native = getattr(type(criterion), "_native", None)
        if native is None or not hasattr(native, "Workspace"):
            raise RuntimeError("GraphedASGLoss needs the B200 build of torch_asg_native (no CPU path, no fallback)")
        self._native = native

    # ---- the reference's attributes, forwarded ----
    @property
    def transition(self):
        return self.criterion.transition

    @property
    def reduction(self):
        return self.criterion.reduction

    # ---- plumbing --------------------------------------------------------------------------------
    def _scale(self, n_local, device):
        """Upstream gradient of every utterance's loss: 1 ('sum'/none) or 1/global batch ('mean')."""
        if self.criterion.reduction != 'mean':
            return 1.0
        gb = self.global_batch
        if gb is None:
            world = _world(self.group)
            if world == 1:
                gb = n_local
            else:                           # shards may be uneven: count once per local size, never assume n_local * world
                gb = self._global_n.get(n_local)
                if gb is None:
                    t = torch.tensor([float(n_local)], device=device, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                    gb = self._global_n[n_local] = int(round(float(t.item())))
        return 1.0 / float(gb)

    def _lengths(self, inputs, targets, input_lengths, target_lengths):
        T, N, _ = inputs.shape
        S = targets.shape[1]
        if input_lengths is None or target_lengths is None:
            key = (inputs.device, T, N, S)
            full = self._full_lengths.get(key)
            if full is None:
                full = self._full_lengths[key] = (torch.full((N,), T, dtype=torch.long, device=inputs.device),
                                                  torch.full((N,), S, dtype=torch.long, device=inputs.device))
            input_lengths = full[0] if input_lengths is None else input_lengths
            target_lengths = full[1] if target_lengths is None else target_lengths
        return input_lengths, target_lengths

    def _entry(self, inputs, targets, il, tl):
        T, N, L = inputs.shape
        S = targets.shape[1]
        A = self.criterion.transition
        self._resolve_exchange(inputs.device)
        wkey = (inputs.dtype, inputs.device, T, N, L, S)
        ws = self._ws.get(wkey)
        if ws is None:
            ws = self._ws[wkey] = self._native.Workspace(inputs.dtype == torch.float64, inputs.get_device(), T, N, L, S)
        key = (wkey, inputs.data_ptr(), inputs.stride(), targets.data_ptr(), targets.stride(), il.data_ptr(), tl.data_ptr(),
               A.data_ptr(), A.stride())
        # A graph holds raw addresses and is replayed only when a call arrives with exactly those addresses (then the
        # memory is the caller's live tensors), so no reference to the tensors it was captured with is kept.
        e = self._entries.get(key)
        if e is None:
            self._misses += 1
            if self.use_graph and self._misses > self.max_graphs and self._hits < self._misses:
                import warnings
                warnings.warn("GraphedASGLoss: the tensors' addresses change from call to call (%d captures, %d reuses); "
                              "switching to eager zero-allocation calls (use_graph=False)" % (self._misses, self._hits))
                self.use_graph = False
            e = _Entry(ws)
            self._entries[key] = e
            while len(self._entries) > self.max_graphs:
                self._entries.popitem(last=False)
        else:
            self._hits += 1
            self._entries.move_to_end(key)
        return e

    def _resolve_exchange(self, device):
        """'auto' -> the peer-memory kernel (torch_asg/peer.py) when every pair of ranks has peer access, else NCCL.
        Collective; runs once, outside any capture."""
        if self._exchange_resolved:
            return
        self._exchange_resolved = True
        if self.exchange in ("peer", "auto") and self._peer is None and _world(self.group) > 1:
            from .peer import PeerAllReduce
            self._peer = PeerAllReduce.create(self._native, self.group, device)
            if self._peer is None:
                if self.exchange == "peer":
                    raise RuntimeError("GraphedASGLoss(exchange='peer'): the ranks have no peer access to each other")
                self.exchange = "nccl"
            else:
                self.exchange = "peer"
        elif self.exchange == "auto":
            self.exchange = "nccl"

    def _exchange(self, ws):
        """The one exchange of a batch-sharded job: fused <- sum over ranks."""
        if _world(self.group) > 1:
            if self._peer is not None:
                self._peer.all_reduce(ws.fused)
            else:
                dist.all_reduce(ws.fused, op=dist.ReduceOp.SUM, group=self.group)

    def _tail_and_reduce(self, ws):
        """fused = grad_A | loss sum | count, then the one exchange of a batch-sharded job."""
        ws.fused[-2:].copy_(ws.loss_tail[:2])
        self._exchange(ws)

    def _body(self, e, kind, args):
        ws = e.ws
        x, A, y, il, tl = args
        scale = self._scale(ws.N, ws.grad_x.device)
        if kind == "fwd":
            ws.forward(x, A, y, il, tl, False, scale)
        elif kind == "fwd_only":
            ws.forward(x, A, y, il, tl, True, scale)
        elif kind == "bwd":
            ws.backward(ws.gout, scale, y)
            self._tail_and_reduce(ws)
        else:                               # "step": the whole training step
            ws.step(x, A, y, il, tl, scale)
            self._exchange(ws)
            if e.gloss is None:
                e.gloss = torch.zeros((), dtype=ws.fused.dtype, device=ws.fused.device)
            if self.criterion.reduction == 'mean':
                torch.div(ws.fused[-2], ws.fused[-1], out=e.gloss)
            else:
                e.gloss.copy_(ws.fused[-2])

    def _run(self, e, kind, args):
        """Replay (capturing on first use) the graph of `kind` for this entry, or run it eagerly."""
        if not self.use_graph:
            before = self._native._launch_count()
            self._body(e, kind, args)
            self.kernel_launches += self._native._launch_count() - before
            return
        g = getattr(e, kind)
        if g is None:
            g = self._capture(e, kind, args)
            setattr(e, kind, g)
        if g is False:                      # capture was not possible (e.g. this NCCL cannot be captured): eager
            before = self._native._launch_count()
            self._body(e, kind, args)
            self.kernel_launches += self._native._launch_count() - before
            return
        g.replay()
        self.graph_replays += 1
        self.kernel_launches += e.kernels[kind]

    def _capture(self, e, kind, args):
        dev = e.ws.grad_x.device
        side = self._streams.get(dev)
        if side is None:
            side = self._streams[dev] = torch.cuda.Stream(device=dev)
        self._scale(e.ws.N, dev)             # resolves the global batch (a collective + sync) outside the capture
        # warm-up on the capture stream: creates the library's side stream/events for it, sets kernel attributes,
        # initialises NCCL's communicator -- none of which may happen inside a capture
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            before = self._native._launch_count()
            self._body(e, kind, args)
            e.kernels[kind] = self._native._launch_count() - before
            self._body(e, kind, args)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        try:
            # thread_local: other threads (NCCL's watchdog, pinned-memory copies) may call into CUDA during the capture
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                self._body(e, kind, args)
        except Exception as exc:            # noqa: BLE001 -- report and fall back to the eager zero-allocation calls
            import warnings
            warnings.warn("GraphedASGLoss: CUDA graph capture of '%s' failed (%s); running it eagerly" % (kind, exc))
            torch.cuda.synchronize(dev)
            if _world(self.group) > 1 and kind in ("bwd", "step"):
                self.nccl_in_graph = False
            return False
        if _world(self.group) > 1 and kind in ("bwd", "step"):
            self.nccl_in_graph = self._peer is None
        g.replay()                          # the first launch of an instantiated graph uploads it to the device: pay that here
        return g

    def prepare(self, inputs, targets, input_lengths=None, target_lengths=None, kinds=("fwd", "bwd")):
        """Capture the graphs for these tensors now (e.g. before a timed loop) instead of on first use."""
        il, tl = self._lengths(inputs, targets, input_lengths, target_lengths)
        e = self._entry(inputs, targets, il, tl)
        if self.use_graph:
            args = (inputs.detach(), self.criterion.transition.detach(), targets, il, tl)
            for kind in kinds:
                if getattr(e, kind) is None:
                    setattr(e, kind, self._capture(e, kind, args))
        return e

    def _graphable(self, inputs, targets):
        return (inputs.is_cuda and inputs.dim() == 3 and targets.dim() == 2 and targets.shape[1] <= inputs.shape[0]
                and inputs.dtype in (torch.float32, torch.float64) and inputs.dtype == self.criterion.transition.dtype
                and targets.dtype == torch.long and not getattr(self.criterion, "gpu_no_stream_impl", False))

    # ---- the reference's call ----------------------------------------------------------------------
    def forward(self, inputs, targets, input_lengths=None, target_lengths=None):
        crit = self.criterion
        if not self._graphable(inputs, targets):
            return crit(inputs, targets, input_lengths, target_lengths)
        il, tl = self._lengths(inputs, targets, input_lengths, target_lengths)
        e = self._entry(inputs, targets, il, tl)
        reduce = crit.reduction in ('mean', 'sum')
        if crit.forward_only or not crit.training or not torch.is_grad_enabled():
            self._run(e, "fwd_only", (inputs.detach(), crit.transition.detach(), targets, il, tl))
            return e.ws.loss_tail[2].detach() if reduce else e.ws.losses.detach().clone()
        return _GraphedFn.apply(inputs, crit.transition, self, e, targets, il, tl, reduce)

    def step(self, inputs, targets, input_lengths=None, target_lengths=None):
        """Forward + backward (+ the all-reduce) as ONE graph replay, no autograd graph.  On return ``inputs.grad`` IS the
        workspace's gradient buffer (overwritten by the next step of the same shape -- consume it before) when ``inputs``
        requires grad, ``transition.grad`` is a view of the fused buffer (the global gradient), and the returned 0-dim
        tensor is the global reduced loss."""
        crit = self.criterion
        if not self._graphable(inputs, targets) or crit.reduction not in ('mean', 'sum'):
            raise RuntimeError("GraphedASGLoss.step needs CUDA float tensors, S <= T and reduction 'mean' or 'sum'")
        il, tl = self._lengths(inputs, targets, input_lengths, target_lengths)
        e = self._entry(inputs, targets, il, tl)
        self._run(e, "step", (inputs.detach(), crit.transition.detach(), targets, il, tl))
        ws = e.ws
        if inputs.requires_grad:
            if inputs.grad_fn is None:
                inputs.grad = ws.grad_x
            else:
                inputs.backward(ws.grad_x)
        crit.transition.grad = ws.fused[:ws.L * ws.L].view(ws.L, ws.L)
        return e.gloss

    @property
    def range_flags(self):
        """int32 [N] range report of the most recent call (``asg_range_flags``): 0 everywhere unless the inputs are so peaked
        that fp32's ~87 nats of in-frame / in-block dynamic range were exceeded; rerun such a batch in float64."""
        return next(reversed(self._entries.values())).ws.range_flags()

    def check(self):
        """Synchronise and validate the most recent call: raises IndexError when a target label was outside [0, L) (the kernels
        clamped it; the reference gathers out of bounds), RuntimeError when a peer exchange timed out, and returns the int32 [N]
        range flags (all zero unless the emissions were too peaked for this dtype -- rerun that batch in float64)."""
        flags = self.range_flags
        f = flags.cpu()
        if bool((f & 8).any()):
            raise IndexError("GraphedASGLoss: utterance %d has a target label outside [0, %d)" % (int((f & 8).nonzero()[0]), self.criterion.num_labels))
        if self._peer is not None and self._peer.status()[1]:
            raise RuntimeError("GraphedASGLoss: a peer exchange timed out waiting for another rank (epoch %d)" % self._peer.status()[1])
        return flags

    @property
    def global_loss(self):
        """After a backward (or step): the reduced loss over ALL ranks' utterances, from the all-reduced fused buffer."""
        ws = next(reversed(self._entries.values())).ws
        if self.criterion.reduction == 'mean':
            return ws.fused[-2] / ws.fused[-1]
        return ws.fused[-2]
