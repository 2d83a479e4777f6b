"""Mirror of reference models/InteractionModule.py: InteractionModule :9-55 and
Reversed_InteractionModule :61-108.  ``forward(text, image) -> ([out (B, Lq, D)], sim_paths (B, B))``.

The whole stack (all routing layers, both directions of every gradient) runs as ONE autograd node built
from ``stack.stack_forward`` / ``stack.stack_backward``.  ``num_layer_routing >= 2`` and
``num_cells in {4, 6}`` are accepted; the reference itself only runs with 6 cells and >= 3 layers
(SURVEY §0 #3/#4), so the other values are reference-derived generalisations.
"""
import torch.nn as nn

from .. import stack as S
from ..autograd import run_block, run_blocks
from .DynamicInteraction import (DynamicInteraction_Layer, DynamicInteraction_Layer0,
                                 Reversed_DynamicInteraction_Layer, Reversed_DynamicInteraction_Layer0)


def _check_inputs(mod, text, image):
    """The reference fails on these inputs somewhere inside its first matmul (RuntimeError); fail at the boundary, with
    the same exception type and a message that names the argument."""
    D = mod.args.embed_size
    for name, t in (("text", text), ("image", image)):
        if t.dim() != 3 or t.shape[-1] != D:
            raise RuntimeError(f"d2r_b200.{type(mod).__name__}: the {name} input must be [batch, tokens, {D}], "
                               f"got {tuple(t.shape)}")
        if t.shape[1] == 0:
            raise RuntimeError(f"d2r_b200.{type(mod).__name__}: the {name} input has no tokens")
    if text.shape[0] != image.shape[0] or text.shape[0] == 0:
        raise RuntimeError(f"d2r_b200.{type(mod).__name__}: batch sizes differ or are zero "
                           f"({text.shape[0]} and {image.shape[0]})")


def _stack_request(mod, own, ctx):
    R, Kc = mod.num_layer_routing, mod.num_cells
    heads = mod.dynamic_itr_l0.imrc.sa.h

    def fwd(env, ts):
        out, sim, probs, state = S.stack_forward(env, ts[0], ts[1], R, Kc)
        return (out, sim) + tuple(probs), dict(state=state, nprob=len(probs))

    def bwd(env, sd, grads):
        d_probs = list(grads[2:2 + sd["nprob"]])
        dx, dz = S.stack_backward(env, sd["state"], grads[0], grads[1],
                                  d_probs if any(g is not None for g in d_probs) else None)
        return dx, dz

    return dict(module=mod, inputs=[own, ctx], fwd=fwd, bwd=bwd, heads=heads)


def _stack_result(mod, res, want_probs):
    mod.last_path_probs = [p.detach() for p in res[2:]]   # per-layer routing probabilities (inspection only)
    if want_probs:
        return [res[0]], res[1], list(res[2:])
    return [res[0]], res[1]


def _stack_call(mod, own, ctx, want_probs=False):
    return _stack_result(mod, run_block(**_stack_request(mod, own, ctx)), want_probs)


def run_pair(itr_module, reversed_itr_module, text, image, return_path_probs=False):
    """Both branch stacks of one batch, concurrently.  Replaces the reference's back-to-back calls
    (models/modeling_unimo.py:842-843)

        sim_mat, sim_paths = self.itr_module(text, image)
        Reversed_sim_mat, Reversed_sim_paths = self.Reversed_itr_module(text, image)

    with ``(sim_mat, sim_paths), (Reversed_sim_mat, Reversed_sim_paths) = run_pair(self.itr_module,
    self.Reversed_itr_module, text, image)``.  The two stacks are independent until the loss, so they are
    issued on two CUDA streams inside one autograd node (forward and backward): on B200 the tail waves and the
    many small launches of one stack are filled by the other.  Results are identical to the two separate calls."""
    _check_inputs(itr_module, text, image)
    rt, ri = run_blocks([_stack_request(itr_module, text, image), _stack_request(reversed_itr_module, image, text)])
    return _stack_result(itr_module, rt, return_path_probs), _stack_result(reversed_itr_module, ri, return_path_probs)


class InteractionModule(nn.Module):
    def __init__(self, args, num_layer_routing=3, num_cells=4, path_hid=128):
        super(InteractionModule, self).__init__()
        if num_layer_routing < 2:
            raise ValueError("d2r_b200: num_layer_routing must be >= 2")
        self.args = args
        self.num_cells = num_cells
        self.num_layer_routing = num_layer_routing
        self.dynamic_itr_l0 = DynamicInteraction_Layer0(args, num_cells, num_cells)
        self.dynamic_itr_l1 = nn.ModuleList([DynamicInteraction_Layer(args, num_cells, num_cells)
                                             for i in range(num_layer_routing - 2)])
        self.dynamic_itr_l2 = DynamicInteraction_Layer(args, num_cells, 1)
        total_paths = num_cells ** 2 * (num_layer_routing - 1) + num_cells
        self.path_mapping = nn.Linear(total_paths, path_hid)   # constructed but unused upstream (:19)
        self.bn = nn.BatchNorm1d(args.embed_size)              # constructed but unused upstream (:20)

    def forward(self, text, image, return_path_probs=False):
        _check_inputs(self, text, image)
        return _stack_call(self, text, image, return_path_probs)


class Reversed_InteractionModule(nn.Module):
    def __init__(self, args, num_layer_routing=3, num_cells=4, path_hid=128):
        super(Reversed_InteractionModule, self).__init__()
        if num_layer_routing < 2:
            raise ValueError("d2r_b200: num_layer_routing must be >= 2")
        self.args = args
        self.num_cells = num_cells
        self.num_layer_routing = num_layer_routing
        self.dynamic_itr_l0 = Reversed_DynamicInteraction_Layer0(args, num_cells, num_cells)
        self.dynamic_itr_l1 = nn.ModuleList([Reversed_DynamicInteraction_Layer(args, num_cells, num_cells)
                                             for i in range(num_layer_routing - 2)])
        self.dynamic_itr_l2 = Reversed_DynamicInteraction_Layer(args, num_cells, 1)
        total_paths = num_cells ** 2 * (num_layer_routing - 1) + num_cells
        self.path_mapping = nn.Linear(total_paths, path_hid)
        self.bn = nn.BatchNorm1d(args.embed_size)

    def forward(self, text, image, return_path_probs=False):
        # the image stream is this branch's own stream, the text is the context (reference :157-165)
        _check_inputs(self, text, image)
        return _stack_call(self, image, text, return_path_probs)
