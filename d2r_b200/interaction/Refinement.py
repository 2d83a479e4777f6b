"""Mirror of the live part of reference models/Refinement.py (CrossModalAlignment :83-117, Refinement :120-154)."""
import torch.nn as nn

from .. import stack as S
from ..autograd import run_block
from .XModules import hidden_size_of, _cma_block


class FallbackConfig:
    """Stand-in for a Hugging Face config that cannot be loaded (no network, no local copy): only ``hidden_size`` is
    ever read (reference Cells.py:93, Refinement.py:131).  Module-level so that ``torch.save(model)`` can pickle it."""
    hidden_size = 768


def _bert_config(name):
    try:
        from transformers import BertConfig
        return BertConfig.from_pretrained(name)
    except Exception:
        return FallbackConfig()


class CrossModalAlignment(nn.Module):
    def __init__(self, config):
        super(CrossModalAlignment, self).__init__()
        self.config = config
        hs = hidden_size_of(config)
        self.query = nn.Linear(hs, hs)
        self.key = nn.Linear(hs, hs)
        self.value = nn.Linear(hs, hs)
        self.fc_1 = nn.Linear(hs, hs)   # dead in the reference, kept for state_dict parity
        self.fc_2 = nn.Linear(hs, hs)

    def forward(self, text_emb, image_emb):
        return _cma_block(self, text_emb, image_emb)


class Refinement(nn.Module):
    def __init__(self, args, embed_size, raw_feature_norm, lambda_softmax):
        super(Refinement, self).__init__()
        self.raw_feature_norm = raw_feature_norm      # stored but dead in the reference (Refinement.py:123-124)
        self.lambda_softmax = lambda_softmax
        self.fc_scale = nn.Linear(embed_size, embed_size)
        self.fc_shift = nn.Linear(embed_size, embed_size)
        self.fc_1 = nn.Linear(embed_size, embed_size)
        self.fc_2 = nn.Linear(embed_size, embed_size)
        self.CrossModalAlignment = CrossModalAlignment(_bert_config(args.bert_name))

    def forward(self, text, image):
        def fwd(env, xs):
            x, z = xs
            kv = S._KV(env, z, ["RF.CrossModalAlignment"])
            out, sv = S._cmrc_fwd(env, "RF", x, kv, 0)
            sv.update(kv=kv, x=x)
            return (out,), sv

        def bwd(env, sv, grads):
            dx = S._cmrc_bwd(env, "RF", sv["x"], sv["kv"], 0, sv, grads[0], None)
            return dx, sv["kv"].backward(env, None)

        return run_block(self, [text, image], fwd, bwd, prefix="RF.")[0]
