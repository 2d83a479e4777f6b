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
  ``js_div``                                 :849 (XModules.js_div on sim_paths / Reversed_sim_paths)

The swapped-in modules are bound to the SAME ``nn.Parameter`` / buffer objects as the modules they replace (no
copy): ``state_dict`` keys, optimizer parameter groups built by name (modules/train.py:293-320) and checkpoints
(train.py:215) are untouched, and ``accelerate`` may run at any point of the model's life.
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn

from .interaction.Cells import BertPooler
from .interaction.InteractionModule import InteractionModule, Reversed_InteractionModule, run_pair
from .interaction.XModules import Block, js_div


class _PairedInteraction(InteractionModule):
    """First of the two back-to-back stack calls: runs both stacks concurrently, parks the partner's result."""

    def forward(self, text, image, return_path_probs=False):
        partner = self.__dict__.get("_d2r_partner")
        if partner is None or return_path_probs:
            return super().forward(text, image, return_path_probs)
        mine, other = run_pair(self, partner, text, image)
        partner.__dict__["_d2r_parked"] = (text, image, other)
        return mine


class _PairedReversedInteraction(Reversed_InteractionModule):
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
    return new


def _find_backbone(model: nn.Module) -> nn.Module:
    for m in model.modules():
        if hasattr(m, "itr_module") and hasattr(m, "Reversed_itr_module"):
            return m
    raise RuntimeError("d2r_b200.accelerate: no module with itr_module / Reversed_itr_module found")


def accelerate(model: nn.Module, *, pair: bool = True, head: bool = True) -> nn.Module:
    """See module docstring.  ``pair=False`` keeps the two stack calls separate; ``head=False`` leaves the CLS
    poolers, the Block fusion and js_div on the reference's PyTorch code."""
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
            mod.js_div = js_div          # the symbol modeling_unimo.py:849 resolves at call time
    model.__dict__["_d2r_accelerated"] = dict(pair=pair, head=head)
    return model
