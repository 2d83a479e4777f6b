"""Data-parallel plumbing for the routed stack: one process per GPU, batch sharded across ranks, weights
replicated, ONE gradient all-reduce (mean) per step over NCCL / NVLink -- nothing else crosses GPUs
(SURVEY.md §8e).  The reference has no distributed code at all; this is the B200-side design.

Static plan: parameters that never receive a gradient in the reference (the dead ``CrossModalAlignment.
fc_1/fc_2``, ``path_mapping`` and ``bn``) are excluded up front, so no unused-parameter search
is needed.  Gradients are packed into one flat fp32 buffer (a single NVSwitch all-reduce is bandwidth-
optimal: every peer is one hop away) and ``p.grad`` is re-pointed at views of the reduced buffer.

Batch-coupled pieces stay per-rank exactly as they would in the reference under DP: the BatchNorm1d(1) of
the attention filtration uses the local batch's statistics and ``sim_paths`` is the local [b, b] Gram, so
N-rank DP equals the reference run independently on each shard, with averaged gradients.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

DEAD_SUFFIXES = (".CrossModalAlignment.fc_1.weight", ".CrossModalAlignment.fc_1.bias",
                 ".CrossModalAlignment.fc_2.weight", ".CrossModalAlignment.fc_2.bias")
DEAD_PREFIXES = ("path_mapping.", "bn.")


def is_live(name: str) -> bool:
    leaf = name.split("itr_module.")[-1]
    return not (name.endswith(DEAD_SUFFIXES) or leaf.startswith(DEAD_PREFIXES))


def shard_batch(n: int, rank: int, world: int):
    """Contiguous B/N samples per rank (the remainder goes to the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Flat-bucket mean all-reduce of the live parameter gradients of one or more modules."""

    def __init__(self, modules: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.params: List[torch.nn.Parameter] = []
        self.names: List[str] = []
        for mi, m in enumerate(modules):
            for n, p in m.named_parameters():
                if p.requires_grad and is_live(n):
                    self.params.append(p)
                    self.names.append(f"{mi}.{n}")
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None

    def pack(self) -> torch.Tensor:
        """Copy every live gradient into the flat bucket (one fused copy); missing gradients are a bug."""
        missing = [n for n, p in zip(self.names, self.params) if p.grad is None]
        if missing:
            raise RuntimeError(f"GradAllReducer: live parameters without gradient: {missing[:3]} ...")
        if self.flat is None:
            p0 = self.params[0]
            self.flat = torch.empty(self.numel, device=p0.device, dtype=torch.float32)
        views, off = [], 0
        for p in self.params:
            views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        torch._foreach_copy_(views, [p.grad for p in self.params])
        self._views = views
        return self.flat

    def all_reduce(self, async_op: bool = False):
        """Mean all-reduce of the packed bucket (call pack() first)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM,
                               group=self.group, async_op=async_op)

    def finish(self) -> None:
        """Point p.grad at the reduced bucket (gloo has no AVG: divide here)."""
        if not self.flat.is_cuda and dist.is_initialized():
            self.flat.div_(dist.get_world_size(self.group))
        for p, v in zip(self.params, self._views):
            p.grad = v

    def step(self) -> None:
        self.pack()
        self.all_reduce()
        self.finish()


class InputPrefetcher:
    """Host -> device input pipeline of one rank: the pinned host batch of step i+1 is copied to the GPU on a copy
    stream while step i computes, then moved into the (static, CUDA-graph visible) input tensors with a
    device-to-device copy.  Every step's inputs still cross PCIe exactly once; only the wait is hidden.

        pf = InputPrefetcher([d_text, d_image])
        pf.fetch([h_text0, h_image0])              # first batch
        for i in range(steps):
            pf.commit()                            # inputs of step i are now in d_text / d_image
            if i + 1 < steps:
                pf.fetch(next_host_batch)          # overlaps with the compute below
            ... forward / backward on the current stream ...
    """

    def __init__(self, static_inputs: List[torch.Tensor]):
        self.static = list(static_inputs)
        self.stage = [torch.empty_like(t) for t in self.static]
        self.copy_stream = torch.cuda.Stream(device=self.static[0].device)
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self.free.record(torch.cuda.current_stream(self.static[0].device))
        self.bytes_per_fetch = sum(t.numel() * t.element_size() for t in self.static)

    def fetch(self, host_batch: List[torch.Tensor]) -> None:
        """Start the H2D copy of the next batch (pinned host tensors) on the copy stream."""
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free)         # the previous commit has drained the staging buffers
            with torch.no_grad():
                for s, h in zip(self.stage, host_batch):
                    s.copy_(h, non_blocking=True)
            self.ready.record(self.copy_stream)

    def commit(self) -> None:
        """Make the fetched batch the current input (current stream waits for the copy, then D2D)."""
        cur = torch.cuda.current_stream(self.static[0].device)
        cur.wait_event(self.ready)
        with torch.no_grad():
            for d, s in zip(self.static, self.stage):
                d.copy_(s, non_blocking=True)
        self.free.record(cur)
