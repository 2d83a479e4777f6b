"""Glue between ``torch.autograd`` and the hand-scheduled passes in ``stack.py``.

``run_block(module, inputs, fwd, bwd)`` runs ``fwd(env, inputs_cd) -> (outputs, state)`` as ONE
autograd node whose backward is ``bwd(env, state, grad_outputs) -> input_grads``; parameter
gradients collected in ``env.G`` are handed back to autograd so ``.grad`` / optimizers / DDP work
unchanged.  Parameters stay ordinary fp32 ``nn.Parameter`` objects with the reference's names.

Precision: bf16 arithmetic (tcgen05) when the call runs under ``torch.autocast(bfloat16)`` or is
given bf16 inputs, fp32 arithmetic otherwise -- the two modes the reference supports.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import kernels as K
from . import lanes as LN
from . import stack as S

Tensor = torch.Tensor


def compute_dtype(*tensors: Tensor) -> torch.dtype:
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    if any(t is not None and t.dtype == torch.bfloat16 for t in tensors):
        return torch.bfloat16
    return torch.float32


def _stager_of(module: torch.nn.Module) -> S.Stager:
    st = module.__dict__.get("_d2r_stager")
    if st is None:
        st = S.Stager()
        module.__dict__["_d2r_stager"] = st
    return st


def _fwd_one(spec, tensors):
    """Run one block's forward on the CURRENT stream -> (outs, record for the backward)."""
    n_in, names, bufs, fwd, bwd, cd, training, stager, heads, out_nondiff = spec[:10]
    inputs = tensors[:n_in]
    P: Dict[str, Tensor] = dict(zip(names, tensors[n_in:]))
    P.update(bufs)
    env = S.Env(P, cd, training, stager, heads)
    env.layer_hook = spec[10]
    env.bn_modes = spec[11]
    x_cd = [None if t is None else t.detach().to(cd).contiguous() for t in inputs]
    outs, state = fwd(env, x_cd)
    # the state aliases inputs and parameters instead of going through ctx.save_for_backward: keep autograd's safety
    # net by hand -- their version counters are compared again when the backward starts
    rec = dict(spec=spec, state=state, in_dtypes=[None if t is None else t.dtype for t in inputs],
               out_meta=[(o.shape, o.dtype) for o in outs],
               versions=[(t, t._version) for t in tensors if t is not None])
    return tuple(outs), rec


def _bwd_one(rec, grads) -> List[Optional[Tensor]]:
    """Run one block's backward on the CURRENT stream -> [input grads..., parameter grads...]."""
    n_in, names, bufs, fwd, bwd, cd, training, stager, heads, out_nondiff = rec["spec"][:10]
    # parameters are only needed by name -> the forward's tensors are reachable through the state's env
    env: S.Env = rec["state"]["env"]
    env.G = {}
    gs = []
    for g, (shape, dtype) in zip(grads, rec["out_meta"]):
        gs.append(None if g is None else g.to(dtype).contiguous())
    in_grads = bwd(env, rec["state"], gs)
    env.aux_sync()                      # helper streams of this pass (bias-gradient sums) rejoin their parents
    rec["state"] = None
    res: List[Optional[Tensor]] = []
    for g, dt in zip(in_grads, rec["in_dtypes"]):
        res.append(None if (g is None or dt is None) else g.to(dt))
    for nme in names:
        res.append(env.G.get(nme))
    return res


def _block_lanes(dev, n):
    """Lanes of n concurrent blocks -> (lanes, index of block 0's lane).  With priorities, every block runs on a
    side stream (the caller's stream only forks and joins): block 0 -- the text stack under run_pair, twice the
    work of the image stack and therefore the critical chain -- gets high-priority streams, so its CTAs are
    scheduled first whenever SMs free up and the other block fills what is left."""
    if n > 1 and LN.PRIORITIZE_FIRST_BLOCK:
        return LN.fork(dev, n + 1, "blocks", priorities=[0, -1] + [0] * (n - 1)), 1
    return LN.fork(dev, n, "blocks"), 0


class _BlocksFn(torch.autograd.Function):
    """N independent blocks as ONE autograd node.  With N > 1 the blocks are issued on separate CUDA streams:
    their kernels fill each other's partial waves and sub-148-CTA launches.  Each lane allocates its temporaries
    while its own stream is current, every fork waits on the caller's stream and every call joins before it
    returns, so no block of memory is reused across lanes without an ordering edge."""

    @staticmethod
    def forward(ctx, specs, out_counts, *tensors):
        dev = next(t.device for t in tensors if t is not None)
        _same_device(dev, tensors)
        # every launch, stream lookup and allocation below refers to the operands' device, whatever the caller's
        # current device is (the C ABI launches on the stream it is handed and never calls cudaSetDevice)
        with _device_guard(dev):
            lanes, first = _block_lanes(dev, len(specs))
            recs, all_outs, off = [], [], 0
            for i, spec in enumerate(specs):
                n = spec[0] + len(spec[1])
                with lanes.lane(first + i):
                    outs, rec = _fwd_one(spec, tensors[off:off + n])
                off += n
                recs.append(rec)
                all_outs.append(outs)
            lanes.join()
        ctx.recs = recs
        ctx.dev = dev
        out_counts.extend(len(o) for o in all_outs)     # (caller-owned list: how to split the flat result)
        nd = [outs[j] for outs, spec in zip(all_outs, specs) for j in spec[9]]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(o for outs in all_outs for o in outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        recs = ctx.recs
        if recs is None:
            raise RuntimeError("d2r_b200: trying to backward through the stack a second time -- the saved "
                               "activations are freed by the first backward (retain_graph is not supported)")
        for rec in recs:                 # before any launch or stream fork
            for t, version in rec["versions"]:
                if t._version != version:
                    raise RuntimeError("d2r_b200: one of the variables needed for gradient computation (an input or a "
                                       f"parameter of the stack, shape {tuple(t.shape)}) has been modified by an "
                                       f"inplace operation: it is at version {t._version}, the forward saw version "
                                       f"{version}")
            rec["versions"] = None
        with _device_guard(ctx.dev):
            # atomically-accumulated gradients (bias, router head, SAF) are carved out of per-stream zeroed arenas:
            # start every backward -- of the whole stack or of a stand-alone cell / Block -- with fresh ones, so that
            # a CUDA graph of the pass contains the memsets and replays do not accumulate onto stale values
            K.zero_arena_reset()
            lanes, first = _block_lanes(ctx.dev, len(recs))
            res: List[Optional[Tensor]] = [None, None]
            off = 0
            for i, rec in enumerate(recs):
                n = len(rec["out_meta"])
                with lanes.lane(first + i):
                    res += _bwd_one(rec, grads[off:off + n])
                off += n
            lanes.join()
        ctx.recs = None
        return tuple(res)


def _device_guard(dev):
    """Make the operands' device current for the duration of a pass (CPU devices only occur in the emulated
    tests of tests/test_stack_emulated.py, where the kernels are replaced)."""
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def _same_device(dev, tensors) -> None:
    for t in tensors:
        if t is not None and t.device != dev:
            raise RuntimeError(f"d2r_b200: inputs and parameters of one call must live on one device "
                               f"(found {dev} and {t.device})")


def _require_cuda(inputs) -> None:
    for t in inputs:
        if t is not None and not t.is_cuda:
            raise RuntimeError("d2r_b200: the routed interaction stack runs on CUDA (sm_100a) only; "
                               "there is no CPU path")


def _bn_modes(module: torch.nn.Module, prefix: str) -> Dict[str, bool]:
    """Train/eval flag of every BatchNorm1d below ``module``, by prefixed name.  The reference's batch-norm layers each
    read their OWN flag (XModules.py:380-384 calls ``self.bn``), so statistics frozen with ``bn.eval()`` under
    ``model.train()`` -- or the reverse -- must behave the same here.  The module list is collected once."""
    bns = module.__dict__.get("_d2r_bns")
    if bns is None:
        bns = [(n, m) for n, m in module.named_modules() if isinstance(m, torch.nn.BatchNorm1d)]
        module.__dict__["_d2r_bns"] = bns
    return {prefix + n: m.training for n, m in bns}


def _make_spec(module: torch.nn.Module, inputs: Sequence[Optional[Tensor]], fwd: Callable, bwd: Callable, *,
               prefix: str = "", cd: Optional[torch.dtype] = None, heads: int = 16,
               out_nondiff: Tuple[int, ...] = ()):
    _require_cuda(inputs)
    cd = cd or compute_dtype(*inputs)
    names, params = [], []
    for n, p in module.named_parameters():
        names.append(prefix + n)
        params.append(p)
    bufs = {prefix + n: b for n, b in module.named_buffers()}

    def fwd_wrapped(env, xs):
        outs, state = fwd(env, xs)
        state = dict(state)
        state["env"] = env
        return outs, state

    # optional per-layer gradient callback (d2r_b200.dp.GradAllReducer.install): layer prefix, {name: grad}
    spec = (len(inputs), tuple(names), bufs, fwd_wrapped, bwd, cd, module.training, _stager_of(module), heads,
            tuple(out_nondiff), module.__dict__.get("_d2r_layer_hook"), _bn_modes(module, prefix))
    return spec, list(inputs) + params


def run_block(module: torch.nn.Module, inputs: Sequence[Optional[Tensor]],
              fwd: Callable, bwd: Callable, *, prefix: str = "", cd: Optional[torch.dtype] = None,
              heads: int = 16, out_nondiff: Tuple[int, ...] = ()) -> Tuple[Tensor, ...]:
    """See module docstring.  ``prefix`` is prepended to the module's parameter names so that helper
    code written against full stack names (e.g. 'L.glac.fc_1') can serve a stand-alone sub-module."""
    spec, tensors = _make_spec(module, inputs, fwd, bwd, prefix=prefix, cd=cd, heads=heads, out_nondiff=out_nondiff)
    return _BlocksFn.apply((spec,), [], *tensors)


def run_blocks(requests: Sequence[dict]) -> List[Tuple[Tensor, ...]]:
    """Several independent blocks (each a dict of run_block's arguments) as one autograd node, issued on
    concurrent CUDA streams.  Returns one output tuple per request."""
    specs, tensors = [], []
    for rq in requests:
        rq = dict(rq)
        spec, ts = _make_spec(rq.pop("module"), rq.pop("inputs"), rq.pop("fwd"), rq.pop("bwd"), **rq)
        specs.append(spec)
        tensors += ts
    counts: List[int] = []
    flat = _BlocksFn.apply(tuple(specs), counts, *tensors)
    # split by each block's number of outputs: not known before the forward ran, so the forward reports it
    res, off = [], 0
    for n in counts:
        res.append(tuple(flat[off:off + n]))
        off += n
    return res
