"""Data parallelism over volumes (one process per GPU; reference: Lightning DDP, vqvae/train.py:26-27).

The path has exactly two exchanges per training step: the gradient all-reduce below and the flat
[counts | dw] EMA-statistics all-reduce inside `Quantizer._forward_impl`.  Both go through torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from typing import Iterable

import torch
import torch.distributed as dist

from . import _ops


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int = None, flat: torch.Tensor = None) -> None:
    """Average the gradients of `params` over the ranks with ONE all-reduce of a flat fp32 buffer (7.5 M floats = 30 MB
    for the Full model: a single latency-bound NVLink message instead of DDP's 25 MB buckets + hooks).  `flat`: the
    optimizer's flat gradient buffer (FusedAdamAMSGrad.flat_grad) -- the gradients already live there, nothing is copied."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world_size or dist.get_world_size()
    if world == 1:
        return
    if flat is not None:
        dist.all_reduce(flat)
        flat.div_(world)
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def training_step(model, optimizer, batch) -> torch.Tensor:
    """One data-parallel optimisation step of the reference's training loop (model.py:95-113 + DDP + Adam):
    forward in training mode (EMA statistics all-reduced inside the quantizers), backward, gradient average, step."""
    optimizer.zero_grad(set_to_none=True)
    ops = _ops.default()
    loss = model.training_step(batch, 0)
    ops.begin_step(loss.device)            # one launch zeroes the scalar-gradient scratch of every convolution backward of the step
    loss.backward()
    ops.end_step()
    allreduce_gradients(model.parameters(), flat=getattr(optimizer, "flat_grad", None))
    optimizer.step()
    return loss.detach()


class GraphedTrainingStep:
    """`training_step` captured in ONE CUDA graph (forward in training mode, backward through the library's kernels, the
    gradient all-reduce, the fused Adam step): a step of the published configurations is ~9 000 library calls, so replaying
    a graph removes the host from the loop.  Needs fixed batch shapes, a FusedAdamAMSGrad in flat mode (device-side step
    counter) and quantizers whose EMA init has run; the first `warmup` steps run eagerly on a side stream (they also do
    the data-dependent codebook init), then one step is captured and every call copies the batch into the static buffers
    and replays.  Multi-rank capture records the NCCL all-reduces into the graph as well."""

    def __init__(self, model, optimizer, example_batch, warmup: int = 3):
        if getattr(optimizer, "flat_grad", None) is None:
            raise RuntimeError("GraphedTrainingStep needs FusedAdamAMSGrad(flatten=True) on CUDA parameters")
        x, num_valid = example_batch
        self.model, self.optimizer = model, optimizer
        self.x = x.detach().clone()
        self.num_valid = torch.as_tensor(num_valid, dtype=torch.int32, device=x.device).reshape(-1).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                training_step(model, optimizer, (self.x, self.num_valid))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = training_step(model, optimizer, (self.x, self.num_valid))
        self.warmup_steps = max(1, warmup)           # optimisation steps already applied (capture itself does not execute)

    def __call__(self, batch) -> torch.Tensor:
        x, num_valid = batch
        self.x.copy_(x, non_blocking=True)
        self.num_valid.copy_(torch.as_tensor(num_valid, dtype=torch.int32).reshape(-1), non_blocking=True)
        self.graph.replay()
        return self.loss
