"""Mirror of the live parts of reference models/XModules.py: l1norm/l2norm :14-24, CrossModalAlignment
:277-328 (live part :300-310; the reverse-attention / ContrastiveLoss branch :312-326 is dead work whose
result every caller discards -- its parameters fc_1/fc_2 are kept for state_dict parity, the returned
loss is a constant 0), AttentionFiltration :366-394, js_div :32-41 (fused kernel).  ``Block`` :478-555 (the
bilinear fusion the backbone applies after the stack, SURVEY §8f: batched merge GEMMs + one fused rank-sum /
signed-sqrt / normalise kernel)."""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import stack as S
from .. import kernels as K
from .. import autograd as _A
from ..autograd import run_block


def l2norm(X, dim=-1, eps=1e-8):
    norm = torch.pow(X, 2).sum(dim=dim, keepdim=True).sqrt() + eps
    return torch.div(X, norm)


def l1norm(X, dim, eps=1e-8):
    norm = torch.abs(X).sum(dim=dim, keepdim=True) + eps
    return torch.div(X, norm)


class _JsDivFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, q, get_softmax):
        pf, qf = p.detach().float().contiguous(), q.detach().float().contiguous()
        ctx.save_for_backward(pf, qf)
        ctx.get_softmax = get_softmax
        ctx.dtypes = (p.dtype, q.dtype)
        return K.js_div_fwd(pf, qf, get_softmax)

    @staticmethod
    def backward(ctx, d_loss):
        pf, qf = ctx.saved_tensors
        dp, dq = K.js_div_bwd(pf, qf, d_loss.detach().float().contiguous(), ctx.get_softmax)
        return dp.to(ctx.dtypes[0]), dq.to(ctx.dtypes[1]), None


def js_div(p_output, q_output, get_softmax=True):
    """reference models/XModules.py:32-41: JS divergence between the row-softmaxes of two logit matrices
    (``(sim_paths, sim_text)`` and ``(Reversed_sim_paths, sim_vision)`` at modeling_unimo.py:849); KLDivLoss
    'batchmean' = sum over all elements / first dimension.  One fused kernel each way,
    2-D CUDA inputs only, like the rest of the path."""
    if p_output.dim() != 2 or p_output.shape != q_output.shape:
        raise ValueError("d2r_b200.js_div: expects two [rows, cols] matrices of the same shape")
    _A._require_cuda([p_output, q_output])
    return _JsDivFn.apply(p_output, q_output, bool(get_softmax))


def hidden_size_of(config) -> int:
    return int(getattr(config, "hidden_size", 768))


def _cma_block(module, text_emb, image_emb):
    def fwd(env, xs):
        x, z = xs
        kv = S._KV(env, z, ["C"])
        out, sv = S._cma_fwd(env, "C", x, kv, 0)
        return (out,), dict(cma=sv, kv=kv, x=x)

    def bwd(env, st, grads):
        dx = S._cma_bwd(env, "C", st["x"], st["kv"], 0, st["cma"], grads[0], 1.0, None)
        dz = st["kv"].backward(env, None)
        return dx, dz

    return run_block(module, [text_emb, image_emb], fwd, bwd, prefix="C.")[0]


class CrossModalAlignment(nn.Module):
    def __init__(self, config, args):
        super(CrossModalAlignment, self).__init__()
        self.config, self.args = config, args
        hs = hidden_size_of(config)
        self.query = nn.Linear(hs, hs)
        self.key = nn.Linear(hs, hs)
        self.value = nn.Linear(hs, hs)
        self.fc_1 = nn.Linear(hs, hs)   # dead in the reference, kept for state_dict parity
        self.fc_2 = nn.Linear(hs, hs)

    def forward(self, text_emb, image_emb):
        out = _cma_block(self, text_emb, image_emb)
        return out, out.new_zeros(())   # (text_img_rep_init, text_img_loss); the loss is discarded upstream


class AttentionFiltration(nn.Module):
    def __init__(self, sim_dim):
        super(AttentionFiltration, self).__init__()
        self.attn_sim_w = nn.Linear(sim_dim, 1)
        self.bn = nn.BatchNorm1d(1)
        self.init_weights()

    def init_weights(self):
        for m in self.children():
            if isinstance(m, nn.Linear):
                r = np.sqrt(6.) / np.sqrt(m.in_features + m.out_features)
                m.weight.data.uniform_(-r, r)
                m.bias.data.fill_(0)
            elif isinstance(m, nn.BatchNorm1d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, sim_emb):
        """sim_emb (B, L+1, D) -> (B, D)."""
        def fwd(env, xs):
            s = xs[0]
            P = env.P
            sg, sl = s[:, 0].contiguous(), s[:, 1:].contiguous()
            out, saved = K.saf_fwd(sg, sl, P["attn_sim_w.weight"].detach().view(-1), P["attn_sim_w.bias"].detach(),
                                   P["bn.weight"].detach(), P["bn.bias"].detach(), P["bn.running_mean"],
                                   P["bn.running_var"], P.get("bn.num_batches_tracked"), env.bn_training("bn"))
            return (out,), dict(sg=sg, sl=sl, saved=saved)

        def bwd(env, st, grads):
            P = env.P
            d_sg, d_sl, d_w, d_b, d_bnw, d_bnb = K.saf_bwd(
                grads[0], st["sg"], st["sl"], P["attn_sim_w.weight"].detach().view(-1), P["attn_sim_w.bias"].detach(),
                P["bn.weight"].detach(), P["bn.bias"].detach(), P["bn.running_mean"], P["bn.running_var"],
                env.bn_training("bn"), st["saved"])
            env.grad("attn_sim_w.weight", d_w.view(1, -1))
            env.grad("attn_sim_w.bias", d_b)
            env.grad("bn.weight", d_bnw)
            env.grad("bn.bias", d_bnb)
            return (torch.cat([d_sg.unsqueeze(1), d_sl], 1),)

        return run_block(self, [sim_emb], fwd, bwd)[0]


class Block(nn.Module):
    """Mirror of reference models/XModules.py:478-555 (bilinear fusion of the two pooled branch outputs,
    ``modeling_unimo.py:776,884``): same constructor signature, parameter names / shapes / creation order.
    The drop-in covers what the reference instantiates -- ``shared=False``, no dropout, ``pos_norm='before_cat'``,
    ``mm_dim`` divisible by ``chunks`` -- and refuses the other (never used) option values at construction."""

    def __init__(self, input_dims, output_dim, mm_dim=1600, chunks=20, rank=15, shared=False, dropout_input=0.,
                 dropout_pre_lin=0., dropout_output=0., pos_norm='before_cat'):
        super(Block, self).__init__()
        size = mm_dim // chunks
        if shared or dropout_input or dropout_pre_lin or dropout_output or pos_norm != 'before_cat':
            raise ValueError("d2r_b200.Block: only shared=False, dropout 0 and pos_norm='before_cat' are supported")
        if mm_dim % chunks or size % 8 or size > 128:
            raise ValueError("d2r_b200.Block: mm_dim / chunks must be an integer multiple of 8, at most 128")
        self.input_dims, self.output_dim, self.mm_dim = input_dims, output_dim, mm_dim
        self.chunks, self.rank, self.shared = chunks, rank, shared
        self.dropout_input, self.dropout_pre_lin, self.dropout_output = dropout_input, dropout_pre_lin, dropout_output
        self.pos_norm = pos_norm
        self.linear0 = nn.Linear(input_dims[0], mm_dim)
        self.linear1 = nn.Linear(input_dims[1], mm_dim)
        self.sizes_list = [size] * chunks
        merge_linears0, merge_linears1 = [], []
        for s in self.sizes_list:                      # creation order alternates, as upstream (:507-515)
            merge_linears0.append(nn.Linear(s, s * rank))
            merge_linears1.append(nn.Linear(s, s * rank))
        self.merge_linears0 = nn.ModuleList(merge_linears0)
        self.merge_linears1 = nn.ModuleList(merge_linears1)
        self.linear_out = nn.Linear(mm_dim, output_dim)
        self.n_params = sum(p.numel() for p in self.parameters() if p.requires_grad)

    def forward(self, x):
        chunks, rank = self.chunks, self.rank

        def fwd(env, ts):
            out, st = S.block_forward(env, ts[0], ts[1], chunks, rank)
            return (out,), st

        def bwd(env, st, grads):
            return S.block_backward(env, st, grads[0])

        return run_block(self, [x[0], x[1]], fwd, bwd)[0]
