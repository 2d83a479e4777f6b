"""Glue between ``torch.autograd`` and the hand-scheduled passes in ``stack.py``.

``run_block(module, inputs, fwd, bwd)`` runs ``fwd(env, inputs_cd) -> (outputs, state)`` as ONE
autograd node whose backward is ``bwd(env, state, grad_outputs) -> input_grads``; parameter
gradients collected in ``env.G`` are handed back to autograd so ``.grad`` / optimizers / DDP work
unchanged.  Parameters stay ordinary fp32 ``nn.Parameter`` objects with the reference's names.

Precision: bf16 arithmetic (tcgen05) when the call runs under ``torch.autocast(bfloat16)`` or is
given bf16 inputs, fp32 arithmetic otherwise -- the two modes the reference supports.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

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


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, *tensors):
        n_in, names, bufs, fwd, bwd, cd, training, stager, heads, out_nondiff = spec
        inputs = tensors[:n_in]
        P: Dict[str, Tensor] = dict(zip(names, tensors[n_in:]))
        P.update(bufs)
        env = S.Env(P, cd, training, stager, heads)
        x_cd = [None if t is None else t.detach().to(cd).contiguous() for t in inputs]
        outs, state = fwd(env, x_cd)
        ctx.spec = spec
        ctx.state = state
        ctx.in_dtypes = [None if t is None else t.dtype for t in inputs]
        ctx.out_meta = [(o.shape, o.dtype) for o in outs]
        nd = [outs[i] for i in out_nondiff]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        n_in, names, bufs, fwd, bwd, cd, training, stager, heads, out_nondiff = ctx.spec
        P: Dict[str, Tensor] = {}
        # parameters are only needed by name -> the forward's tensors are reachable through the state's env
        env: S.Env = ctx.state["env"]
        env.G = {}
        gs = []
        for g, (shape, dtype) in zip(grads, ctx.out_meta):
            gs.append(None if g is None else g.to(dtype).contiguous())
        in_grads = bwd(env, ctx.state, gs)
        ctx.state = None
        res: List[Optional[Tensor]] = [None]
        for g, dt in zip(in_grads, ctx.in_dtypes):
            res.append(None if (g is None or dt is None) else g.to(dt))
        for nme in names:
            res.append(env.G.get(nme))
        return tuple(res)


def _require_cuda(inputs) -> None:
    for t in inputs:
        if t is not None and not t.is_cuda:
            raise RuntimeError("d2r_b200: the routed interaction stack runs on CUDA (sm_100a) only; "
                               "there is no CPU path")


def run_block(module: torch.nn.Module, inputs: Sequence[Optional[Tensor]],
              fwd: Callable, bwd: Callable, *, prefix: str = "", cd: Optional[torch.dtype] = None,
              heads: int = 16, out_nondiff: Tuple[int, ...] = ()) -> Tuple[Tensor, ...]:
    """See module docstring.  ``prefix`` is prepended to the module's parameter names so that helper
    code written against full stack names (e.g. 'L.glac.fc_1') can serve a stand-alone sub-module."""
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

    spec = (len(inputs), tuple(names), bufs, fwd_wrapped, bwd, cd, module.training, _stager_of(module), heads,
            tuple(out_nondiff))
    return _BlockFn.apply(spec, *inputs, *params)
