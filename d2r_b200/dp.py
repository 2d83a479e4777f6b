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


LAYER_HOOK_ATTR = "_d2r_layer_hook"


def _layer_of(name: str) -> str:
    """'dynamic_itr_l1.0.imrc...' -> 'dynamic_itr_l1.0'; 'dynamic_itr_l2.glac...' -> 'dynamic_itr_l2'."""
    parts = name.split(".")
    return ".".join(parts[:2]) if parts[0] == "dynamic_itr_l1" else parts[0]


class GradAllReducer:
    """Flat-bucket mean all-reduce of the live parameter gradients of one or more stack modules.

    Two ways to drive it:
      * ``step()`` (= ``pack(); all_reduce(); finish()``) after ``loss.backward()``: one collective per step;
      * ``install()``: the stack's backward calls back after every routing layer (last layer first) with that
        layer's finished gradients; they are copied into the layer's slice of the flat bucket and the slice's
        all-reduce is launched at once (async, NCCL's stream), so it runs under the backward of the earlier
        layers and of the other stack.  ``wait()`` joins them; the pattern is capturable in a CUDA graph.
    The flat bucket is ordered by completion (module, layer in backward order), so both ways end in the same
    state: every live ``p.grad`` is a view of one reduced buffer (``finish()`` / ``wait()`` re-point them, since
    autograd assigns the local gradients to ``p.grad`` during ``backward()``).
    """

    def __init__(self, modules: Iterable[torch.nn.Module], group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.modules = list(modules)
        self.params: List[torch.nn.Parameter] = []
        self.names: List[str] = []
        self.slices = {}                      # (module index, layer prefix) -> (first param idx, end idx, off, numel)
        for mi, m in enumerate(self.modules):
            live = [(n, p) for n, p in m.named_parameters() if p.requires_grad and is_live(n)]
            layers: List[str] = []
            for n, _ in live:
                if _layer_of(n) not in layers:
                    layers.append(_layer_of(n))
            for layer in reversed(layers):    # the backward finishes the last routing layer first
                i0, off = len(self.params), sum(p.numel() for p in self.params)
                for n, p in live:
                    if _layer_of(n) == layer:
                        self.params.append(p)
                        self.names.append(f"{mi}.{n}")
                self.slices[(mi, layer)] = (i0, len(self.params), off,
                                            sum(p.numel() for p in self.params[i0:]))
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self._views: Optional[List[torch.Tensor]] = None
        self._works: list = []
        self._filled = 0
        self._pending_div = False
        self.active = True                    # install()ed callbacks do nothing while this is False

    # ------------------------------------------------------------------ shared
    def _ensure_flat(self) -> None:
        if self.flat is None:
            p0 = self.params[0]
            self.flat = torch.empty(self.numel, device=p0.device, dtype=torch.float32)
            views, off = [], 0
            for p in self.params:
                views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._views = views

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, t: torch.Tensor, async_op: bool):
        if not t.is_cuda:
            self._pending_div = True          # gloo has no AVG: sum now, divide exactly once in _settle()
        return dist.all_reduce(t, op=dist.ReduceOp.AVG if t.is_cuda else dist.ReduceOp.SUM, group=self.group,
                               async_op=async_op)

    def _settle(self) -> None:
        """After the collective(s) of a step: finish the mean on gloo and make every live p.grad a view of the
        reduced bucket (autograd has set p.grad to the local, un-reduced tensors during backward)."""
        if self._pending_div:
            self.flat.div_(dist.get_world_size(self.group))
            self._pending_div = False
        for p, v in zip(self.params, self._views):
            p.grad = v

    # ------------------------------------------------------------------ one collective per step
    def pack(self) -> torch.Tensor:
        """Copy every live gradient into the flat bucket (one fused copy); missing gradients are a bug."""
        missing = [n for n, p in zip(self.names, self.params) if p.grad is None]
        if missing:
            raise RuntimeError(f"GradAllReducer: live parameters without gradient: {missing[:3]} ...")
        self._ensure_flat()
        torch._foreach_copy_(self._views, [p.grad for p in self.params])
        return self.flat

    def all_reduce(self, async_op: bool = False):
        """Mean all-reduce of the packed bucket (call pack() first)."""
        if not self._distributed():
            return None
        return self._reduce(self.flat, async_op)

    def finish(self) -> None:
        """Point p.grad at the reduced bucket (idempotent: the gloo division happens once per reduction)."""
        self._settle()

    def step(self) -> None:
        self.pack()
        self.all_reduce()
        self.finish()

    # ------------------------------------------------------------------ layer-wise, overlapped with the backward
    def install(self) -> None:
        """Register the per-layer callback on the modules (picked up by the stack's autograd node)."""
        for mi, m in enumerate(self.modules):
            m.__dict__[LAYER_HOOK_ATTR] = (lambda layer, grads, _mi=mi: self.on_layer(_mi, layer, grads))

    def uninstall(self) -> None:
        for m in self.modules:
            m.__dict__.pop(LAYER_HOOK_ATTR, None)

    def on_layer(self, mi: int, layer: str, grads) -> None:
        """``grads``: name -> finished gradient for (at least) every live parameter of ``layer`` of module ``mi``.
        Runs on the stream that computed them."""
        if not self.active:
            return
        i0, i1, off, numel = self.slices[(mi, layer)]
        self._ensure_flat()
        prefix = f"{mi}."
        src = []
        for n in self.names[i0:i1]:
            g = grads.get(n[len(prefix):])
            if g is None:
                raise RuntimeError(f"GradAllReducer: live parameter without gradient: {n}")
            src.append(g)
        torch._foreach_copy_(self._views[i0:i1], src)
        self._filled += 1
        if self._distributed():
            self._works.append(self._reduce(self.flat[off:off + numel], async_op=True))

    def wait(self) -> None:
        """Join the layer collectives (the current stream waits; no host block for NCCL)."""
        if self._filled != len(self.slices):
            raise RuntimeError(f"GradAllReducer: {self._filled} of {len(self.slices)} layer buckets were filled")
        for w in self._works:
            w.wait()
        self._works, self._filled = [], 0
        self._settle()                        # p.grad -> views of the reduced bucket, as after step()


class FlatGradReducer:
    """Mean all-reduce of the gradients of an arbitrary parameter list through ONE flat fp32 bucket -- the whole
    reference model (encoders on stock PyTorch + the routed stacks) under data parallelism.  The live set is fixed
    at the first ``step()``: parameters whose ``.grad`` is still None after a backward never receive one in this
    model (110 of the 1184 tensors of UnimoModelF, SURVEY §8e caveat 3) and are left out of the bucket, so no
    unused-parameter search runs per step."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None):
        self.all_params = [p for p in params if p.requires_grad]
        self.group = group
        self.params: Optional[List[torch.nn.Parameter]] = None
        self.flat: Optional[torch.Tensor] = None
        self._views: List[torch.Tensor] = []

    def _plan(self) -> None:
        self.params = [p for p in self.all_params if p.grad is not None]
        if not self.params:
            raise RuntimeError("FlatGradReducer: no parameter has a gradient (call after backward)")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=self.params[0].device, dtype=torch.float32)
        off = 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def dead(self) -> int:
        return len(self.all_params) - len(self.params or [])

    def pack(self) -> torch.Tensor:
        if self.params is None:
            self._plan()
        missing = sum(1 for p in self.params if p.grad is None)
        if missing:
            raise RuntimeError(f"FlatGradReducer: {missing} planned parameters have no gradient this step")
        torch._foreach_copy_(self._views, [p.grad for p in self.params])
        return self.flat

    def all_reduce(self) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            if self.flat.is_cuda:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(dist.get_world_size(self.group))

    def finish(self) -> None:
        for p, v in zip(self.params, self._views):
            p.grad = v

    def step(self) -> None:
        self.pack()
        self.all_reduce()
        self.finish()


class BucketedGradReducer:
    """Whole-model data parallelism with the gradient all-reduce OVERLAPPED with the backward (what DDP does, with a
    static plan instead of an unused-parameter search).

    After one ordinary backward (``plan()``: the live set is whatever received a gradient -- 1074 of the 1184 tensors
    of UnimoModelF), every live ``p.grad`` becomes a VIEW of one flat fp32 buffer; autograd accumulates into the views
    in place, so no packing copy exists.  The buffer is cut into buckets of about ``bucket_mb`` in reverse parameter
    order -- the order in which gradients become final -- and a post-accumulate hook launches a bucket's asynchronous
    mean all-reduce (NCCL over NVLink) the moment its last gradient has landed, under the rest of the backward.

        red = BucketedGradReducer(model.parameters()); loss.backward(); red.plan()       # once, after
                                                                                         # init_process_group
        for batch in data:
            red.begin_step()            # zero the flat buffer (p.grad stay views of it)
            loss = model(batch); loss.backward()
            red.finish()                # join the bucket collectives; p.grad now hold the rank-mean
            optimizer.step()
    One backward per step (no gradient accumulation over micro-batches: a bucket is reduced when its parameters have
    been written once)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 128.0,
                 group: Optional[dist.ProcessGroup] = None):
        self.all_params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.params: List[torch.nn.Parameter] = []
        self.flat: Optional[torch.Tensor] = None
        self.buckets: List[list] = []          # [first element, end element, parameters in the bucket]
        self._pending: List[int] = []
        self._works: list = []
        self._hooks: list = []

    @property
    def dead(self) -> int:
        return len(self.all_params) - len(self.params)

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def plan(self) -> None:
        live = [p for p in self.all_params if p.grad is not None]
        if not live:
            raise RuntimeError("BucketedGradReducer.plan(): call after a backward")
        self.params = list(reversed(live))                      # roughly the order gradients are produced in
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        off = 0
        cur = [0, 0, 0]
        index = {}
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            if cur[2] and off - cur[0] >= self.bucket_elems:
                cur[1] = off
                self.buckets.append(cur)
                cur = [off, 0, 0]
            index[p] = len(self.buckets)
            cur[2] += 1
            off += p.numel()
        cur[1] = off
        self.buckets.append(cur)
        self._pending = [b[2] for b in self.buckets]
        if not self._distributed():
            return      # single process: nothing to reduce, and 1074 Python hooks cost ~2.8 ms of a host-bound step
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(
                lambda _p, _b=index[p]: self._landed(_b)))

    def _landed(self, b: int) -> None:
        self._pending[b] -= 1
        if self._pending[b] == 0 and self._distributed():
            lo, hi, _ = self.buckets[b]
            t = self.flat[lo:hi]
            op = dist.ReduceOp.AVG if t.is_cuda else dist.ReduceOp.SUM
            self._works.append((dist.all_reduce(t, op=op, group=self.group, async_op=True), t))

    def begin_step(self) -> None:
        self.flat.zero_()
        self._pending = [b[2] for b in self.buckets]
        self._works = []

    def finish(self) -> None:
        if any(self._pending) and self._distributed():
            raise RuntimeError("BucketedGradReducer.finish(): some planned parameters received no gradient this step")
        for w, t in self._works:
            w.wait()
            if not t.is_cuda:
                t.div_(dist.get_world_size(self.group))          # gloo has no AVG
        self._works = []

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


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
