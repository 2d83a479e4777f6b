"""Mirror of reference models/DynamicInteraction.py: DynamicInteraction_Layer0 :20-69,
DynamicInteraction_Layer :72-134 and their Reversed_ twins :140-254 (identical arithmetic with the
roles of text and image swapped)."""
import torch
import torch.nn as nn

from .. import kernels as K
from .. import stack as S
from ..autograd import run_block
from .Cells import (ContextRichCrossModalCell, CrossModalRefinementCell, GlobalEnhancedSemanticCell,
                    GlobalLocalAlignmentCell, IntraModelReasoningCell, RectifiedIdentityCell)


def _check_cells(num_cell):
    if num_cell not in (4, 6):
        raise ValueError("d2r_b200: num_cell must be 6 (the reference's only runnable value) or 4 "
                         "(reference-derived: RIC, GLAC, IMRC, CMRC); got %r" % (num_cell,))


def _build_cells(mod, args, num_cell, num_out_path, layer0):
    """Registration order follows the reference (Layer0: ric, imrc, glac, ...; Layer: ric, glac, imrc, ...)
    so state_dict order and seeded initialisation match."""
    mod.ric = RectifiedIdentityCell(args, num_out_path)
    if layer0:
        mod.imrc = IntraModelReasoningCell(args, num_out_path)
        mod.glac = GlobalLocalAlignmentCell(args, num_out_path)
    else:
        mod.glac = GlobalLocalAlignmentCell(args, num_out_path)
        mod.imrc = IntraModelReasoningCell(args, num_out_path)
    mod.cmrc = CrossModalRefinementCell(args, num_out_path)
    if num_cell > 4:
        mod.crcmc = ContextRichCrossModalCell(args, num_out_path)
        mod.gesc = GlobalEnhancedSemanticCell(args, num_out_path)


def _run_layer(mod, xs, own_first, ctx, shared):
    """One routing layer as a single autograd node.  xs: list of K inputs (or one shared input)."""
    Kc, final = mod.num_cell, mod.num_out_path == 1
    n_in = 1 if shared else Kc
    heads = mod.imrc.sa.h

    def fwd(env, ts):
        own, z = list(ts[:n_in]), ts[n_in]
        xin = own * Kc if shared else own
        pooled = K.pool_mean([xin[0]] if shared else xin)
        outs, _, norm, st = S.layer_forward(env, "L", xin, z, pooled, Kc, final)
        return tuple(outs) + (norm,), dict(st=st, n_out=len(outs))

    def bwd(env, sd, grads):
        st, n_out = sd["st"], sd["n_out"]
        x0 = st["xs"][0]
        B, Lq, D = x0.shape
        d_outs = [g if g is not None else torch.zeros_like(x0) for g in grads[:n_out]]
        d_xs, d_pooled, dz = S.layer_backward(env, "L", st, d_outs, None, grads[n_out], None, shared)
        for j, dx in enumerate(d_xs):
            K.pool_mean_bwd_into(d_pooled[j], dx)
        return tuple(d_xs) + (dz,)

    res = run_block(mod, list(xs) + [ctx], fwd, bwd, prefix="L.", heads=heads)
    return list(res[:-1]), res[-1]


class DynamicInteraction_Layer0(nn.Module):
    def __init__(self, args, num_cell, num_out_path):
        super(DynamicInteraction_Layer0, self).__init__()
        _check_cells(num_cell)
        self.args, self.threshold, self.eps = args, 0.0001, 1e-8
        self.num_cell, self.num_out_path = num_cell, num_out_path
        _build_cells(self, args, num_cell, num_out_path, layer0=True)

    def forward(self, text, image):
        return _run_layer(self, [text], True, image, shared=True)


class DynamicInteraction_Layer(nn.Module):
    def __init__(self, args, num_cell, num_out_path):
        super(DynamicInteraction_Layer, self).__init__()
        _check_cells(num_cell)
        self.args, self.threshold, self.eps = args, 0.0001, 1e-8
        self.num_cell, self.num_out_path = num_cell, num_out_path
        _build_cells(self, args, num_cell, num_out_path, layer0=False)

    def forward(self, ref_wrd, text, image):
        return _run_layer(self, ref_wrd, True, image, shared=False)


class Reversed_DynamicInteraction_Layer0(nn.Module):
    def __init__(self, args, num_cell, num_out_path):
        super(Reversed_DynamicInteraction_Layer0, self).__init__()
        _check_cells(num_cell)
        self.args, self.threshold, self.eps = args, 0.0001, 1e-8
        self.num_cell, self.num_out_path = num_cell, num_out_path
        _build_cells(self, args, num_cell, num_out_path, layer0=True)

    def forward(self, text, image):
        return _run_layer(self, [image], False, text, shared=True)


class Reversed_DynamicInteraction_Layer(nn.Module):
    def __init__(self, args, num_cell, num_out_path):
        super(Reversed_DynamicInteraction_Layer, self).__init__()
        _check_cells(num_cell)
        self.args, self.threshold, self.eps = args, 0.0001, 1e-8
        self.num_cell, self.num_out_path = num_cell, num_out_path
        _build_cells(self, args, num_cell, num_out_path, layer0=False)

    def forward(self, ref_wrd, text, image):
        return _run_layer(self, ref_wrd, False, text, shared=False)
