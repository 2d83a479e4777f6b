"""Mirror of reference models/SelfAttention.py (AttentionLayer :11-42, FeedForward :45-53, SelfAttention :56-70)."""
import copy

import torch.nn as nn

from .. import stack as S
from ..autograd import run_block


def clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class AttentionLayer(nn.Module):
    """Parameter container; the arithmetic lives in SelfAttention.forward (one fused block)."""

    def __init__(self, embed_size, h, is_share=False, drop=0.0):
        super(AttentionLayer, self).__init__()
        if is_share or drop > 0:
            raise NotImplementedError("d2r_b200: the stack uses is_share=False, drop=0.0 (reference Cells.py:47)")
        self.is_share, self.h, self.embed_size, self.d_k, self.drop_p = is_share, h, embed_size, embed_size // h, drop
        self.linears = clones(nn.Linear(embed_size, embed_size), 3)


class FeedForward(nn.Module):
    def __init__(self, embed_size, hidden, drop=0.0):
        super(FeedForward, self).__init__()
        self.fc1 = nn.Linear(embed_size, hidden)
        self.fc2 = nn.Linear(hidden, embed_size)
        self.dropout = nn.Dropout(drop)


class SelfAttention(nn.Module):
    def __init__(self, embed_size, hid_size, h, drop=0.0):
        super(SelfAttention, self).__init__()
        if drop > 0:
            raise NotImplementedError("d2r_b200: dropout p>0 is not part of the reference stack")
        self.h = h
        self.att_layer = AttentionLayer(embed_size, h, drop=drop)
        self.feed_forward_layer = FeedForward(embed_size, hid_size, drop=drop)
        self.dropout = nn.Dropout(drop)

    def forward(self, local_emb, mask=None):
        if mask is not None:
            raise NotImplementedError("d2r_b200: the stack never passes an attention mask (reference Cells.py:57)")

        def fwd(env, xs):
            out, sv = S._imrc_fwd(env, "SA", xs[0])
            sv["x"] = xs[0]
            return (out,), sv

        def bwd(env, sv, grads):
            return (S._imrc_bwd(env, "SA", sv["x"], sv, grads[0], None),)

        return run_block(self, [local_emb], fwd, bwd, prefix="SA.", heads=self.h)[0]
