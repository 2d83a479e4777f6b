"""Drop the B200-native stack into an already constructed reference model, in place.

    from d2r_b200.integration import accelerate
    model = UnimoModelF(args, vision_config, text_config)        # the reference's own (models/unimo_model.py:138)
    accelerate(model)                                            # before or after .to(device) / optimizer creation

What is swapped (everything else -- encoders, embeddings, the extra self-attention layers, the classifier head --
stays the reference's stock PyTorch, as BASELINE.json:north_star asks):

  ``itr_module`` / ``Reversed_itr_module``   models/modeling_unimo.py:781-782  -> d2r_b200.interaction.InteractionModule /
                                             Reversed_InteractionModule; the two back-to-back calls of :842-843 run as
                                             ONE autograd node on two CUDA streams (``run_pair``) through a pair of thin
                                             subclasses: the first call computes both, the second returns its half
  ``text_pool`` / ``vision_pool``            :778-779, :871-872 (BertPooler on row 0 of the stack outputs)
  ``block_fusion``                           :776, :884 (XModules.Block bilinear fusion)
  ``js_div``                                 :849 (XModules.js_div on sim_paths / Reversed_sim_paths): the module-level
                                             symbol :849 resolves becomes a dispatcher that takes the fused kernel only
                                             while an ACCELERATED backbone's forward is running -- other models of the
                                             same class in the process keep the reference's own function

The swapped-in modules are bound to the SAME ``nn.Parameter`` / buffer objects as the modules they replace (no
copy): ``state_dict`` keys, optimizer parameter groups built by name (modules/train.py:293-320) and checkpoints
(train.py:215) are untouched, and ``accelerate`` may run at any point of the model's life.
"""
from __future__ import annotations

import sys
import threading

import torch
import torch.nn as nn

from .autograd import _bn_modes
from .interaction.Cells import BertPooler
from .interaction.InteractionModule import InteractionModule, Reversed_InteractionModule, run_pair
from .interaction.XModules import Block, js_div


class _PairModule(nn.Module):
    """Both stacks as one callable with tensor-only inputs / outputs (what torch.cuda.make_graphed_callables wants).
    A helper object: it is never attached to the model, so it does not show up in the model's state_dict."""

    def __init__(self, itr, rev):
        super().__init__()
        self.itr, self.rev = itr, rev

    def forward(self, text, image):
        (ot, st), (oi, si) = run_pair(self.itr, self.rev, text, image)
        return ot[0], st, oi[0], si


def _graphed_pair(itr, rev, text, image):
    """CUDA-graph the two stacks (forward graph + backward graph, torch.cuda.make_graphed_callables) for this input
    signature.  Inside an otherwise eager model the ~1400 launches of the stacks cost more host time than GPU time at
    small batches; replaying them as two graph launches removes that.  BatchNorm running statistics are restored
    after the warm-up / capture passes (they run the forward four times on the sample)."""
    import contextlib
    bufs = [(b, b.detach().clone()) for m in (itr, rev) for b in m.buffers()]
    bf16 = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
    ctx = torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False) if bf16 else contextlib.nullcontext()
    sample = (text.detach().clone().requires_grad_(True), image.detach().clone().requires_grad_(True))
    with ctx:
        fn = torch.cuda.make_graphed_callables(_PairModule(itr, rev), sample, allow_unused_input=True)
    with torch.no_grad():
        for b, saved in bufs:
            b.copy_(saved)
    return fn


def _portable_state(module: nn.Module) -> dict:
    """Module state for pickling / deep-copying: captured CUDA graphs and a parked result stay behind."""
    d = module.__dict__.copy()
    d.pop("_d2r_graph_cache", None)
    d.pop("_d2r_parked", None)
    return d


class _PairedInteraction(InteractionModule):
    """First of the two back-to-back stack calls: runs both stacks concurrently, parks the partner's result."""

    __getstate__ = _portable_state

    def forward(self, text, image, return_path_probs=False):
        partner = self.__dict__.get("_d2r_partner")
        if partner is None or return_path_probs:
            return super().forward(text, image, return_path_probs)
        if (self.__dict__.get("_d2r_graph") and torch.is_grad_enabled() and text.requires_grad and image.requires_grad
                and not torch.cuda.is_current_stream_capturing()):
            cache = self.__dict__.setdefault("_d2r_graph_cache", {})
            key = (tuple(text.shape), tuple(image.shape), text.dtype, image.dtype, self.training, partner.training,
                   tuple(_bn_modes(self, "").values()), tuple(_bn_modes(partner, "").values()),
                   torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda"))
            fn = cache.get(key)
            if fn is None:
                fn = cache[key] = _graphed_pair(self, partner, text, image)
            ot, st, oi, si = fn(text, image)
            partner.__dict__["_d2r_parked"] = (text, image, ([oi], si))
            return [ot], st
        mine, other = run_pair(self, partner, text, image)
        partner.__dict__["_d2r_parked"] = (text, image, other)
        return mine


class _PairedReversedInteraction(Reversed_InteractionModule):
    __getstate__ = _portable_state

    def forward(self, text, image, return_path_probs=False):
        parked = self.__dict__.pop("_d2r_parked", None)
        if parked is not None and not return_path_probs and parked[0] is text and parked[1] is image:
            return parked[2]
        return super().forward(text, image, return_path_probs)


def _rebind(new: nn.Module, old: nn.Module) -> nn.Module:
    """Make ``new`` own exactly the Parameter / buffer objects of ``old`` (same names, checked)."""
    new_keys, old_keys = list(new.state_dict().keys()), list(old.state_dict().keys())
    if new_keys != old_keys:
        diff = sorted(set(new_keys) ^ set(old_keys))[:5]
        raise RuntimeError(f"d2r_b200.accelerate: state_dict keys differ from the reference module ({diff} ...)")
    for name, p in old.named_parameters():
        mod, _, leaf = name.rpartition(".")
        sub = new.get_submodule(mod) if mod else new
        if tuple(sub._parameters[leaf].shape) != tuple(p.shape):
            raise RuntimeError(f"d2r_b200.accelerate: shape of {name} differs from the reference")
        sub._parameters[leaf] = p
    for name, b in old.named_buffers():
        mod, _, leaf = name.rpartition(".")
        sub = new.get_submodule(mod) if mod else new
        sub._buffers[leaf] = b
    new.train(old.training)
    for name, m in old.named_modules():          # per-module flags (e.g. batch-norm layers frozen with .eval())
        try:
            new.get_submodule(name).training = m.training
        except AttributeError:
            pass
    return new


_js_scope = threading.local()      # depth of accelerated backbone forwards running on this thread


def _js_dispatcher(original):
    """Stand-in for the reference module's global ``js_div``: fused kernel inside an accelerated backbone's forward,
    the reference's own function everywhere else."""
    def js_div_dispatch(*args, **kwargs):
        if getattr(_js_scope, "depth", 0) > 0:
            return js_div(*args, **kwargs)
        return original(*args, **kwargs)
    js_div_dispatch._d2r_original = original
    return js_div_dispatch


def _js_enter(module, args):
    _js_scope.depth = getattr(_js_scope, "depth", 0) + 1


def _js_exit(module, args, output):
    _js_scope.depth = max(0, getattr(_js_scope, "depth", 0) - 1)


def _find_backbone(model: nn.Module) -> nn.Module:
    for m in model.modules():
        if hasattr(m, "itr_module") and hasattr(m, "Reversed_itr_module"):
            return m
    raise RuntimeError("d2r_b200.accelerate: no module with itr_module / Reversed_itr_module found")


def accelerate(model: nn.Module, *, pair: bool = True, head: bool = True, graph: bool = False) -> nn.Module:
    """See module docstring.  ``pair=False`` keeps the two stack calls separate; ``head=False`` leaves the CLS
    poolers, the Block fusion and js_div on the reference's PyTorch code; ``graph=True`` (needs ``pair``) replays the
    two stacks from CUDA graphs in training (static shapes per graph; one graph pair per input signature; switch it on
    before the model's first forward -- the graphs are captured at the first training call of each signature)."""
    bb = _find_backbone(model)
    ref_t, ref_i = bb.itr_module, bb.Reversed_itr_module
    args = ref_t.args
    layers = len(ref_t.dynamic_itr_l1) + 2
    cells = ref_t.num_cells
    with torch.device("meta"):
        new_t = (_PairedInteraction if pair else InteractionModule)(args, layers, cells, ref_t.path_mapping.out_features)
        new_i = (_PairedReversedInteraction if pair else Reversed_InteractionModule)(args, layers, cells,
                                                                                    ref_i.path_mapping.out_features)
    bb.itr_module = _rebind(new_t, ref_t)
    bb.Reversed_itr_module = _rebind(new_i, ref_i)
    if pair:
        new_t.__dict__["_d2r_partner"] = new_i
        new_t.__dict__["_d2r_graph"] = bool(graph)
    if head:
        for name in ("text_pool", "vision_pool"):
            old = getattr(bb, name)
            cfg = type("Cfg", (), {"hidden_size": old.dense.in_features})()
            with torch.device("meta"):
                new = BertPooler(cfg)
            setattr(bb, name, _rebind(new, old))
        old = bb.block_fusion
        with torch.device("meta"):
            new = Block(list(old.input_dims), old.output_dim, mm_dim=old.mm_dim, chunks=old.chunks, rank=old.rank)
        bb.block_fusion = _rebind(new, old)
        mod = sys.modules.get(type(bb).__module__)
        if mod is not None and hasattr(mod, "js_div"):
            if not hasattr(mod.js_div, "_d2r_original"):
                mod.js_div = _js_dispatcher(mod.js_div)      # the symbol modeling_unimo.py:849 resolves at call time
            if "_d2r_js_hooks" not in bb.__dict__:
                bb.__dict__["_d2r_js_hooks"] = (bb.register_forward_pre_hook(_js_enter),
                                                bb.register_forward_hook(_js_exit, always_call=True))
    model.__dict__["_d2r_accelerated"] = dict(pair=pair, head=head, graph=bool(graph and pair))
    return model
