"""Mirror of the live cells of reference models/Cells.py: RectifiedIdentityCell :30-40,
IntraModelReasoningCell :42-60, CrossModalRefinementCell :76-87, BertPooler :90-102,
GlobalLocalAlignmentCell :131-175, GlobalEnhancedSemanticCell :179-218, ContextRichCrossModalCell :222-255.
Each cell returns ``(emb, path_prob)`` exactly like the reference; stand-alone calls run the cell's fused
forward/backward from ``stack.py`` as one autograd node (inside a routing layer the layer runs as one node)."""
import torch
import torch.nn as nn

from .. import kernels as K
from .. import stack as S
from ..autograd import run_block
from .Refinement import FallbackConfig, Refinement, _bert_config
from .Router import Router
from .SelfAttention import SelfAttention
from .XModules import AttentionFiltration, CrossModalAlignment, hidden_size_of, l1norm, l2norm  # noqa: F401


def _clip_vision_config(name):
    try:
        from transformers import CLIPConfig
        return CLIPConfig.from_pretrained(name).vision_config
    except Exception:
        return FallbackConfig()


def _route(env, x, router_prefix):
    pooled = K.pool_mean([x])
    n_out = env.P[router_prefix + ".mlp.2.weight"].shape[0]
    _, _, sv = S._routers_fwd(env, [router_prefix], pooled, n_out, False)
    return sv["raw"].view(x.shape[0], n_out), sv


def _route_bwd(env, sv, d_prob, router_prefix, x_shape):
    B, Ln, D = x_shape
    d_raw = (d_prob if d_prob is not None else torch.zeros(B, sv["raw"].shape[1], device=sv["raw"].device))
    d_pooled = S._routers_bwd(env, [router_prefix], sv, d_raw.reshape(B, -1, 1), True)
    return d_pooled[0]


def _zeros_like_grad(g, like):
    return g if g is not None else torch.zeros_like(like)


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        hs = hidden_size_of(config)
        self.dense = nn.Linear(hs, hs)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        def fwd(env, xs):
            return (S._row0_fwd(env, xs[0], "dense"),), dict(x=xs[0])

        def bwd(env, st, grads):
            dx = torch.zeros_like(st["x"])
            S._row0_bwd(env, grads[0], st["y"], st["x"], "dense", dx)
            return (dx,)

        def fwd2(env, xs):
            (y,), st = fwd(env, xs)
            st["y"] = y
            return (y,), st

        return run_block(self, [hidden_states], fwd2, bwd)[0]


class RectifiedIdentityCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(RectifiedIdentityCell, self).__init__()
        self.keep_mapping = nn.ReLU()
        self.router = Router(num_out_path, args.embed_size, args.hid_router)

    def forward(self, x):
        path_prob = self.router(x)
        # stand-alone only; inside a routing layer relu() is fused into the aggregation kernel
        emb = self.keep_mapping(x)
        return emb, path_prob


class IntraModelReasoningCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(IntraModelReasoningCell, self).__init__()
        self.args = args
        self.router = Router(num_out_path, args.embed_size, args.hid_router)
        self.sa = SelfAttention(args.embed_size, args.hid_IMRC, args.num_head_IMRC)

    def forward(self, inp):
        if inp.dim() != 3:
            raise NotImplementedError("d2r_b200: the stack feeds 3-D (B, L, D) inputs (reference Cells.py:55-56)")
        return self.sa(inp), self.router(inp)


class CrossModalRefinementCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(CrossModalRefinementCell, self).__init__()
        self.refine = Refinement(args, args.embed_size, args.raw_feature_norm_CMRC, args.lambda_softmax_CMRC)
        self.router = Router(num_out_path, args.embed_size, args.hid_router)

    def forward(self, text, image):
        return self.refine(text, image), self.router(text)


class GlobalLocalAlignmentCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(GlobalLocalAlignmentCell, self).__init__()
        self.args = args
        self.router = Router(num_out_path, args.embed_size, args.hid_router)
        cfg = _bert_config(args.bert_name)
        self.CrossModalAlignment = CrossModalAlignment(cfg, args)
        self.SAF_module = AttentionFiltration(hidden_size_of(cfg))
        self.text_cls_pool = BertPooler(cfg)
        self.image_cls_pool = BertPooler(_clip_vision_config(args.vit_name))
        self.fc_sim_tranloc = nn.Linear(768, 768)
        self.fc_sim_tranglo = nn.Linear(768, 768)
        self.fc_1 = nn.Linear(768, 768)
        self.fc_2 = nn.Linear(768, 768)

    def forward(self, text, image):
        def fwd(env, xs):
            x, z = xs
            prob, rsv = _route(env, x, "G.router")
            kv = S._KV(env, z, ["G.CrossModalAlignment"])
            out, sv = S._glac_fwd(env, "G", x, z, kv, 0)
            return (out, prob), dict(sv=sv, rsv=rsv, kv=kv, x=x, z=z)

        def bwd(env, st, grads):
            x, z = st["x"], st["z"]
            dz_rows = torch.zeros_like(z)
            d_out = grads[0] if grads[0] is not None else \
                torch.zeros(x.shape[0], x.shape[2], device=x.device, dtype=torch.float32)
            dx = S._glac_bwd(env, "G", x, z, st["kv"], 0, st["sv"], d_out, None, dz_rows)
            dz = st["kv"].backward(env, dz_rows)
            d_pool = _route_bwd(env, st["rsv"], grads[1], "G.router", x.shape)
            K.pool_mean_bwd_into(d_pool, dx)
            return dx, dz

        sim_emb, path_prob = run_block(self, [text, image], fwd, bwd, prefix="G.")
        return sim_emb.unsqueeze(-2).expand(-1, text.size(1), -1), path_prob


class GlobalEnhancedSemanticCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(GlobalEnhancedSemanticCell, self).__init__()
        self.args = args
        self.router = Router(num_out_path, args.embed_size, args.hid_router)
        self.text_cls_pool = BertPooler(_bert_config(args.bert_name))
        self.image_cls_pool = BertPooler(_bert_config(args.bert_name))
        self.fc_mlp = nn.Sequential(nn.Linear(768, 768), nn.Tanh(), nn.Linear(768, 768))

    def forward(self, text, image):
        def fwd(env, xs):
            x, z = xs
            prob, rsv = _route(env, x, "E.router")
            out, sv = S._gesc_fwd(env, "E", x, z)
            return (out, prob), dict(sv=sv, rsv=rsv, x=x, z=z)

        def bwd(env, st, grads):
            x, z = st["x"], st["z"]
            dx, dz = torch.zeros_like(x), torch.zeros_like(z)
            S._gesc_bwd(env, "E", x, z, st["sv"], _zeros_like_grad(grads[0], st["sv"]["g"]), dx, dz)
            d_pool = _route_bwd(env, st["rsv"], grads[1], "E.router", x.shape)
            K.pool_mean_bwd_into(d_pool, dx)
            return dx, dz

        gate_out, path_prob = run_block(self, [text, image], fwd, bwd, prefix="E.")
        return gate_out.unsqueeze(-2).expand(-1, text.size(1), -1), path_prob


class ContextRichCrossModalCell(nn.Module):
    def __init__(self, args, num_out_path):
        super(ContextRichCrossModalCell, self).__init__()
        self.args = args
        self.router = Router(num_out_path, args.embed_size, args.hid_router)
        self.CrossModalAlignment = CrossModalAlignment(_bert_config(args.bert_name), args)
        self.fc_mlp_1 = nn.Sequential(nn.Linear(768, 768), nn.Tanh())
        self.fc_mlp_2 = nn.Sequential(nn.Linear(768, 768), nn.Tanh())
        self.fc_1 = nn.Linear(768, 768)
        self.fc_2 = nn.Linear(768, 768)

    def forward(self, text, image):
        def fwd(env, xs):
            x, z = xs
            prob, rsv = _route(env, x, "X.router")
            kv = S._KV(env, z, ["X.CrossModalAlignment"])
            out, sv = S._crcmc_fwd(env, "X", x, kv, 0)
            return (out, prob), dict(sv=sv, rsv=rsv, kv=kv, x=x, z=z)

        def bwd(env, st, grads):
            x = st["x"]
            d_out = _zeros_like_grad(grads[0], x)
            dx = S._crcmc_bwd(env, "X", x, st["kv"], 0, st["sv"], d_out, None)
            dz = st["kv"].backward(env, None)
            d_pool = _route_bwd(env, st["rsv"], grads[1], "X.router", x.shape)
            K.pool_mean_bwd_into(d_pool, dx)
            return dx, dz

        return run_block(self, [text, image], fwd, bwd, prefix="X.")
