"""CPU oracle for D2R's dual-branch routed interaction stack.

TEST INFRASTRUCTURE ONLY.  Nothing under ``d2r_b200/`` may import this file; it
is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the *checker* and the *timed CPU
baseline*, never as the product path.

What it is: a functional, plain-PyTorch (fp32, CPU, autograd) restatement of
the reference algorithm.  It takes a flat ``{name: tensor}`` parameter dict
using the reference's ``state_dict`` key names, so a reference checkpoint and
this oracle are interchangeable.  Every function cites the reference file:line
it follows (paths relative to the upstream repo root).

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md §4, §8c).  This oracle is instead pinned against the *unmodified
reference executed in the authoring container*: ``tests/golden/make_golden.py``
imports the reference modules, runs them on seeded inputs and stores the
outputs/gradient digests in ``tests/golden/*.npz``; ``tests/test_oracle.py``
checks this file against those fixtures.

Envelope: the reference only runs with ``num_cells == 6`` and
``num_layer_routing >= 3``.  ``num_cells == 4`` (the
first four cells in ``emb_lst`` order: RIC, GLAC, IMRC, CMRC) and
``num_layer_routing == 2`` are *reference-derived* generalisations.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# cell order inside ``emb_lst`` (models/DynamicInteraction.py:41-48)
CELL_ORDER = ("ric", "glac", "imrc", "cmrc", "crcmc", "gesc")
THRESHOLD = 1e-4   # models/DynamicInteraction.py:23
EPS = 1e-8         # models/DynamicInteraction.py:24
BN_EPS = 1e-5      # torch.nn.BatchNorm1d default, models/XModules.py:376
BN_MOMENTUM = 0.1


def _lin(x: torch.Tensor, P: Params, name: str) -> torch.Tensor:
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


def l1norm(x: torch.Tensor, dim: int, eps: float = 1e-8) -> torch.Tensor:
    """models/XModules.py:20-24 -- eps is added to the norm, not inside it."""
    return x / (x.abs().sum(dim=dim, keepdim=True) + eps)


def l2norm(x: torch.Tensor, dim: int = -1, eps: float = 1e-8) -> torch.Tensor:
    """models/XModules.py:14-18 -- sqrt(sum x^2) + eps."""
    return x / (x.pow(2).sum(dim=dim, keepdim=True).sqrt() + eps)


def router(x: torch.Tensor, P: Params, pre: str) -> torch.Tensor:
    """models/Router.py:22-26 with activateFunc :6-8 -> relu(tanh(MLP(mean_L x)))."""
    pooled = x.mean(-2)
    hid = F.relu(_lin(pooled, P, pre + ".mlp.0"))
    return F.relu(torch.tanh(_lin(hid, P, pre + ".mlp.2")))


def cls_pool(x: torch.Tensor, P: Params, pre: str) -> torch.Tensor:
    """BertPooler, models/Cells.py:96-102: tanh(dense(x[:, 0]))."""
    return torch.tanh(_lin(x[:, 0], P, pre + ".dense"))


def cross_modal_attention(q_in, ctx, P: Params, pre: str, hidden: int = 768,
                          dead_branch: bool = False):
    """Live part of CrossModalAlignment.

    models/XModules.py:300-310 and models/Refinement.py:105-115 (identical):
    softmax(100 * (Wq x)(Wk z)^T / sqrt(hidden)) (Wv z).

    ``dead_branch=True`` additionally executes models/XModules.py:312-326 (reverse
    attention, fc_1/fc_2, normalise, ContrastiveLoss with beta=0) whose result every
    caller discards -- only so that a timed CPU baseline pays the reference's real cost.
    """
    q = _lin(q_in, P, pre + ".query")
    k = _lin(ctx, P, pre + ".key")
    v = _lin(ctx, P, pre + ".value")
    score = torch.bmm(q, k.transpose(-1, -2)) / math.sqrt(hidden)
    attn = torch.softmax(100 * score, dim=-1)
    out = torch.bmm(attn, v)
    if dead_branch:
        rev = torch.softmax(100 * (1 - attn), dim=-1)
        rev_out = torch.bmm(rev, v)
        a = _lin(out, P, pre + ".fc_1").unsqueeze(-2)
        b = _lin(rev_out, P, pre + ".fc_2").unsqueeze(-2)
        tot = F.normalize(torch.cat((a, b), dim=-2))
        txt = F.normalize(q_in.unsqueeze(-2))
        # ContrastiveLoss.forward, models/XModules.py:206-244 (alpha=args.alpha, beta=0)
        s1 = (torch.matmul(tot, txt.transpose(-1, -2).contiguous()) / math.sqrt(tot.size(-1))).squeeze()
        c1 = (0.1 + s1 - s1[:, :, 0].unsqueeze(-1)).clamp(0)
        s2 = torch.matmul(tot[:, :, 0, :], txt.squeeze().transpose(-1, -2).contiguous()) / math.sqrt(tot.size(-1))
        d = torch.diagonal(s2, dim1=-2, dim2=-1).reshape(s2.size(0), -1, 1)
        c2 = (0.1 + s2 - d).clamp(min=0).max(-1)[0]
        _ = 0.0 * c1.sum() + 0.0 * c2.sum()
    return out


# --------------------------------------------------------------------------- cells
def cell_ric(x, P, pre):
    """RectifiedIdentityCell.forward, models/Cells.py:36-40."""
    return F.relu(x), router(x, P, pre + ".router")


def self_attention_block(x, P, pre, heads: int = 16):
    """SelfAttention.forward models/SelfAttention.py:64-70 (AttentionLayer :27-42,
    FeedForward :52-53).  No output projection, no LayerNorm, dropout p=0."""
    B, L, D = x.shape
    dk = D // heads
    q, k, v = [
        _lin(x, P, f"{pre}.att_layer.linears.{i}").view(B, L, heads, dk).transpose(1, 2)
        for i in range(3)
    ]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    p = F.softmax(scores, dim=-1)
    a = torch.matmul(p, v).transpose(1, 2).contiguous().view(B, L, D)
    y = x + a
    ff = _lin(F.relu(_lin(y, P, pre + ".feed_forward_layer.fc1")), P, pre + ".feed_forward_layer.fc2")
    return y + ff


def cell_imrc(x, P, pre, heads: int = 16):
    """IntraModelReasoningCell.forward, models/Cells.py:49-60."""
    return self_attention_block(x, P, pre + ".sa", heads), router(x, P, pre + ".router")


def cell_cmrc(x, ctx, P, pre, dead_branch=False):
    """CrossModalRefinementCell.forward models/Cells.py:82-87 ->
    Refinement.forward/refine models/Refinement.py:141-154, :133-139."""
    prob = router(x, P, pre + ".router")
    wctx = cross_modal_attention(x, ctx, P, pre + ".refine.CrossModalAlignment")
    scaling = torch.tanh(_lin(wctx, P, pre + ".refine.fc_scale"))
    shifting = _lin(wctx, P, pre + ".refine.fc_shift")
    modu = _lin(F.relu(_lin(x * scaling + shifting, P, pre + ".refine.fc_1")), P, pre + ".refine.fc_2")
    return modu + x, prob


def attention_filtration(sim_emb, P, pre, training: bool, bn_updates: Optional[dict]):
    """AttentionFiltration.forward, models/XModules.py:380-384.

    BatchNorm1d(1) over all B*(L+1) scalars: batch statistics in training mode
    (running stats updated with momentum 0.1 / unbiased variance), running statistics
    in eval mode.  ``bn_updates`` (if given) receives the new buffer values instead of
    mutating ``P``.
    """
    logit = _lin(sim_emb, P, pre + ".attn_sim_w").permute(0, 2, 1)   # (B, 1, L+1)
    rm = P[pre + ".bn.running_mean"].clone()
    rv = P[pre + ".bn.running_var"].clone()
    y = F.batch_norm(logit, rm, rv, P[pre + ".bn.weight"], P[pre + ".bn.bias"],
                     training, BN_MOMENTUM, BN_EPS)
    if training and bn_updates is not None:
        bn_updates[pre + ".bn.running_mean"] = rm
        bn_updates[pre + ".bn.running_var"] = rv
        bn_updates[pre + ".bn.num_batches_tracked"] = P[pre + ".bn.num_batches_tracked"] + 1
    attn = l1norm(torch.sigmoid(y), dim=-1)
    saf = torch.matmul(attn, sim_emb).squeeze(1)
    return l2norm(saf, dim=-1)


def cell_glac(x, ctx, P, pre, training=True, bn_updates=None, dead_branch=False):
    """GlobalLocalAlignmentCell.forward/alignment, models/Cells.py:145-175."""
    prob = router(x, P, pre + ".router")
    aware = cross_modal_attention(x, ctx, P, pre + ".CrossModalAlignment", dead_branch=dead_branch)
    sim_local = torch.pow(x - aware, 2)
    sim_local = l2norm(_lin(sim_local, P, pre + ".fc_sim_tranloc"), dim=-1)
    sim_local = _lin(sim_local, P, pre + ".fc_1")
    t_cls = cls_pool(x, P, pre + ".text_cls_pool")
    i_cls = cls_pool(ctx, P, pre + ".image_cls_pool")
    sim_global = torch.pow(t_cls - i_cls, 2)
    sim_global = l2norm(_lin(sim_global, P, pre + ".fc_sim_tranglo"), dim=-1)
    sim_global = _lin(sim_global, P, pre + ".fc_2")
    sim_emb = torch.cat([sim_global.unsqueeze(1), sim_local], 1)
    saf = attention_filtration(sim_emb, P, pre + ".SAF_module", training, bn_updates)
    return saf.unsqueeze(-2).expand(-1, x.size(1), -1), prob


def cell_crcmc(x, ctx, P, pre, dead_branch=False):
    """ContextRichCrossModalCell.forward/alignment, models/Cells.py:236-255."""
    prob = router(x, P, pre + ".router")
    aware = cross_modal_attention(x, ctx, P, pre + ".CrossModalAlignment", dead_branch=dead_branch)
    q_state = torch.tanh(_lin(aware, P, pre + ".fc_mlp_1.0"))
    k_state = torch.tanh(_lin(x, P, pre + ".fc_mlp_2.0"))
    q = _lin(q_state, P, pre + ".fc_1")
    k = _lin(k_state, P, pre + ".fc_2")
    scores = torch.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)   # un-scaled
    return q_state + torch.bmm(scores, k_state), prob


def cell_gesc(x, ctx, P, pre):
    """GlobalEnhancedSemanticCell.forward/global_gate_fusion, models/Cells.py:197-218."""
    prob = router(x, P, pre + ".router")
    t = cls_pool(x, P, pre + ".text_cls_pool")
    i = cls_pool(ctx, P, pre + ".image_cls_pool")
    g = _lin(torch.tanh(_lin(t + i, P, pre + ".fc_mlp.0")), P, pre + ".fc_mlp.2")
    g = torch.softmax(g, dim=-1)
    out = g * t + (1 - g) * i
    return out.unsqueeze(-2).expand(-1, x.size(1), -1), prob


# --------------------------------------------------------------------------- layers
def run_cells(inputs: List[torch.Tensor], ctx, P, pre, K, training, bn_updates, dead_branch):
    """The six cell calls of models/DynamicInteraction.py:41-48 / :93-102.
    ``inputs[j]`` feeds cell j; the context is always the raw other modality."""
    embs, probs = [None] * K, [None] * K
    embs[0], probs[0] = cell_ric(inputs[0], P, pre + ".ric")
    embs[1], probs[1] = cell_glac(inputs[1], ctx, P, pre + ".glac", training, bn_updates, dead_branch)
    embs[2], probs[2] = cell_imrc(inputs[2], P, pre + ".imrc")
    embs[3], probs[3] = cell_cmrc(inputs[3], ctx, P, pre + ".cmrc")
    if K > 4:
        embs[4], probs[4] = cell_crcmc(inputs[4], ctx, P, pre + ".crcmc", dead_branch)
        embs[5], probs[5] = cell_gesc(inputs[5], ctx, P, pre + ".gesc")
    return embs, probs


def aggregate_multi(embs, probs, K):
    """Non-final aggregation, models/DynamicInteraction.py:50-67 (= :118-132).

    gate_i = (sum_j p_j[:, i] < 1e-4); P[b, i, j] = p_j[b, i] / (sum_j p_j[b, i] + 1e-8);
    out_i = sum_j P[:, i, j] * emb_j + gate_i * emb_0.
    """
    gate = (sum(probs) < THRESHOLD).float()                 # (B, K_out)
    allp = torch.stack(probs, dim=2)                        # (B, K_out, K_cell)
    allp = allp / (allp.sum(dim=-1, keepdim=True) + EPS)
    outs = []
    for i in range(allp.size(1)):
        res = 0
        for j in range(K):
            res = res + allp[:, i, j].unsqueeze(-1).unsqueeze(-1) * embs[j]
        res = res + gate[:, i].unsqueeze(-1).unsqueeze(-1) * embs[0]
        outs.append(res)
    return outs, allp


def aggregate_final(embs, probs, inputs, K):
    """Final aggregation (num_out_path == 1), models/DynamicInteraction.py:104-117.

    g_j = (p_j < 1e-4 / K); out = sum_j (p_j emb_j + g_j input_j) / (sum_j g_j + sum_j p_j);
    returned probabilities are NOT normalised.
    """
    res, gates = 0, []
    for j in range(K):
        g = (probs[j] < THRESHOLD / K).float()              # (B, 1)
        gates.append(g)
        res = res + probs[j].unsqueeze(-1) * embs[j]
        res = res + g.unsqueeze(-1) * inputs[j]
    res = res / (sum(gates) + sum(probs)).unsqueeze(-1)
    return [res], torch.stack(probs, dim=2)                 # (B, 1, K)


def stack_forward(P: Params, text: torch.Tensor, image: torch.Tensor, num_layer_routing: int = 3,
                  num_cells: int = 6, reversed_branch: bool = False, training: bool = True,
                  bn_updates: Optional[dict] = None, dead_branch: bool = False
                  ) -> Tuple[List[torch.Tensor], torch.Tensor, List[torch.Tensor]]:
    """InteractionModule.forward (models/InteractionModule.py:22-55) and
    Reversed_InteractionModule.forward (:75-108).

    Returns ``([out], sim_paths, per_layer_probs)``; the third item is extra (the reference
    only exposes it through ``sim_paths``) and is what the routing-probability tolerance of
    the acceptance test is measured on.
    """
    R, K = num_layer_routing, num_cells
    assert R >= 2 and K in (4, 6)
    own, ctx = (image, text) if reversed_branch else (text, image)
    layer_probs = []
    # Layer0: every cell sees the branch's own raw stream (DynamicInteraction.py:37-69)
    embs, probs = run_cells([own] * K, ctx, P, "dynamic_itr_l0", K, training, bn_updates, dead_branch)
    cur, allp = aggregate_multi(embs, probs, K)
    layer_probs.append(allp)
    # middle layers (InteractionModule.py:27-29)
    for li in range(R - 2):
        embs, probs = run_cells(cur, ctx, P, f"dynamic_itr_l1.{li}", K, training, bn_updates, dead_branch)
        cur, allp = aggregate_multi(embs, probs, K)
        layer_probs.append(allp)
    # final layer, num_out_path = 1 (InteractionModule.py:31)
    embs, probs = run_cells(cur, ctx, P, "dynamic_itr_l2", K, training, bn_updates, dead_branch)
    out, allp = aggregate_final(embs, probs, cur, K)
    layer_probs.append(allp)
    B = own.size(0)
    paths = torch.cat([p.reshape(B, -1) for p in layer_probs], dim=-1)   # (B, K^2 (R-1) + K)
    sim_paths = paths @ paths.t()                                        # InteractionModule.py:53
    return out, sim_paths, layer_probs


# --------------------------------------------------------------------------- parameters
def _cell_param_shapes(pre: str, K_out: int, D: int = 768, hid_router: int = 768, hid_imrc: int = 768):
    """(name, shape, kind) for one routing layer, in the reference's construction order
    (models/DynamicInteraction.py:28-35 / :81-88 and the cell constructors)."""
    S = []

    def lin(n, o, i, kind="linear"):
        S.append((n + ".weight", (o, i), kind + "_w"))
        S.append((n + ".bias", (o,), kind + "_b"))

    def rout(n):
        lin(n + ".mlp.0", hid_router, D)
        S.append((n + ".mlp.2.weight", (K_out, hid_router), "linear_w"))
        S.append((n + ".mlp.2.bias", (K_out,), "router_b2"))

    def cma(n):
        for s in ("query", "key", "value", "fc_1", "fc_2"):
            lin(n + "." + s, D, D)

    return S, lin, rout, cma


def layer_param_spec(pre: str, K_out: int, layer0: bool, K: int = 6, D: int = 768):
    S, lin, rout, cma = _cell_param_shapes(pre, K_out, D)
    def ric():
        rout(pre + ".ric.router")
    def imrc():
        rout(pre + ".imrc.router")
        for i in range(3):
            lin(f"{pre}.imrc.sa.att_layer.linears.{i}", D, D)
        lin(pre + ".imrc.sa.feed_forward_layer.fc1", D, D)
        lin(pre + ".imrc.sa.feed_forward_layer.fc2", D, D)
    def glac():
        rout(pre + ".glac.router")
        cma(pre + ".glac.CrossModalAlignment")
        S.append((pre + ".glac.SAF_module.attn_sim_w.weight", (1, D), "saf_w"))
        S.append((pre + ".glac.SAF_module.attn_sim_w.bias", (1,), "zeros"))
        S.append((pre + ".glac.SAF_module.bn.weight", (1,), "ones"))
        S.append((pre + ".glac.SAF_module.bn.bias", (1,), "zeros"))
        S.append((pre + ".glac.SAF_module.bn.running_mean", (1,), "buf_zeros"))
        S.append((pre + ".glac.SAF_module.bn.running_var", (1,), "buf_ones"))
        S.append((pre + ".glac.SAF_module.bn.num_batches_tracked", (), "buf_long"))
        lin(pre + ".glac.text_cls_pool.dense", D, D)
        lin(pre + ".glac.image_cls_pool.dense", D, D)
        for s in ("fc_sim_tranloc", "fc_sim_tranglo", "fc_1", "fc_2"):
            lin(pre + ".glac." + s, D, D)
    def cmrc():
        for s in ("fc_scale", "fc_shift", "fc_1", "fc_2"):
            lin(pre + ".cmrc.refine." + s, D, D)
        cma(pre + ".cmrc.refine.CrossModalAlignment")
        rout(pre + ".cmrc.router")
    def crcmc():
        rout(pre + ".crcmc.router")
        cma(pre + ".crcmc.CrossModalAlignment")
        lin(pre + ".crcmc.fc_mlp_1.0", D, D)
        lin(pre + ".crcmc.fc_mlp_2.0", D, D)
        lin(pre + ".crcmc.fc_1", D, D)
        lin(pre + ".crcmc.fc_2", D, D)
    def gesc():
        rout(pre + ".gesc.router")
        lin(pre + ".gesc.text_cls_pool.dense", D, D)
        lin(pre + ".gesc.image_cls_pool.dense", D, D)
        lin(pre + ".gesc.fc_mlp.0", D, D)
        lin(pre + ".gesc.fc_mlp.2", D, D)
    # Layer0 registers imrc before glac; later layers glac before imrc
    order = [ric, imrc, glac, cmrc, crcmc, gesc] if layer0 else [ric, glac, imrc, cmrc, crcmc, gesc]
    for f in order:
        if K == 4 and f in (crcmc, gesc):
            continue
        f()
    return S


def stack_param_spec(num_layer_routing: int = 3, num_cells: int = 6, path_hid: int = 128, D: int = 768):
    """All state_dict entries of (Reversed_)InteractionModule, incl. dead ones
    (models/InteractionModule.py:10-20)."""
    R, K = num_layer_routing, num_cells
    S = layer_param_spec("dynamic_itr_l0", K, True, K, D)
    for li in range(R - 2):
        S += layer_param_spec(f"dynamic_itr_l1.{li}", K, False, K, D)
    S += layer_param_spec("dynamic_itr_l2", 1, False, K, D)
    total = K * K * (R - 1) + K
    S += [("path_mapping.weight", (path_hid, total), "linear_w"), ("path_mapping.bias", (path_hid,), "linear_b"),
          ("bn.weight", (D,), "ones"), ("bn.bias", (D,), "zeros"),
          ("bn.running_mean", (D,), "buf_zeros"), ("bn.running_var", (D,), "buf_ones"),
          ("bn.num_batches_tracked", (), "buf_long")]
    return S


def make_params(seed: int, num_layer_routing: int = 3, num_cells: int = 6, scale: float = 1.0,
                router_bias: float = 1.5) -> Params:
    """Deterministic synthetic parameters with the reference's default-init *distributions*
    (nn.Linear: U(-1/sqrt(in), 1/sqrt(in)); Router second bias 1.5, models/Router.py:19-20;
    AttentionFiltration: U(+-sqrt(6/(in+out))), bias 0, BN weight 1 bias 0,
    models/XModules.py:386-394).  One torch.Generator stream in spec order, so the same
    (seed, R, K) gives the same bits on any machine with the same torch build -- this is how
    fixtures avoid shipping 300 MB of weights."""
    g = torch.Generator().manual_seed(seed)
    P: Params = {}
    for name, shape, kind in stack_param_spec(num_layer_routing, num_cells):
        if kind in ("linear_w", "linear_b"):
            fan_in = shape[1] if len(shape) == 2 else 768
            if name.startswith("path_mapping") and kind == "linear_b":
                fan_in = num_cells ** 2 * (num_layer_routing - 1) + num_cells
            b = scale / math.sqrt(fan_in)
            P[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "router_b2":
            P[name] = torch.full(shape, float(router_bias))
        elif kind == "saf_w":
            r = math.sqrt(6.0) / math.sqrt(shape[1] + shape[0])
            P[name] = (torch.rand(shape, generator=g) * 2 - 1) * r
        elif kind in ("ones", "buf_ones"):
            P[name] = torch.ones(shape)
        elif kind in ("zeros", "buf_zeros"):
            P[name] = torch.zeros(shape)
        elif kind == "buf_long":
            P[name] = torch.zeros(shape, dtype=torch.long)
        else:
            raise AssertionError(kind)
    return P


def make_inputs(seed: int, B: int, Lt: int, Li: int, D: int = 768, realistic: bool = False):
    """Synthetic inputs of SURVEY.md §8(d): N(0,1) text/image from one seeded CPU generator;
    ``realistic=True`` gives row-LayerNormed text and 3*N(0,1) image with a few x20 outlier
    channels (un-normalised CLIP residual stream)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randn(B, Lt, D, generator=g)
    image = torch.randn(B, Li, D, generator=g)
    if realistic:
        text = F.layer_norm(text, (D,))
        image = 3.0 * image
        image[..., ::97] *= 20.0
    return text, image


DEAD_PARAM_SUFFIXES = (".CrossModalAlignment.fc_1.weight", ".CrossModalAlignment.fc_1.bias",
                       ".CrossModalAlignment.fc_2.weight", ".CrossModalAlignment.fc_2.bias")


ZERO_GRAD_SUFFIXES = (".CrossModalAlignment.key.bias", ".att_layer.linears.1.bias", ".crcmc.fc_2.bias")


def js_div(p_output: torch.Tensor, q_output: torch.Tensor, get_softmax: bool = True) -> torch.Tensor:
    """XModules.py:32-41 -- JS divergence as the reference writes it (KLDivLoss 'batchmean' on the log of the
    mean distribution)."""
    kl = torch.nn.KLDivLoss(reduction="batchmean")
    if get_softmax:
        p_output = torch.softmax(p_output, dim=-1)
        q_output = torch.softmax(q_output, dim=-1)
    log_mean_output = ((p_output + q_output) / 2).log()
    return (kl(log_mean_output, p_output) + kl(log_mean_output, q_output)) / 2


def block_param_spec(input_dims=(768, 768), output_dim=768, mm_dim=1600, chunks=20, rank=15):
    """Parameter names / shapes of XModules.Block in creation order (XModules.py:501-519, shared=False)."""
    size = mm_dim // chunks
    spec = [("linear0.weight", (mm_dim, input_dims[0])), ("linear0.bias", (mm_dim,)),
            ("linear1.weight", (mm_dim, input_dims[1])), ("linear1.bias", (mm_dim,))]
    for group in ("merge_linears0", "merge_linears1"):
        for c in range(chunks):
            spec += [(f"{group}.{c}.weight", (size * rank, size)), (f"{group}.{c}.bias", (size * rank,))]
    spec += [("linear_out.weight", (output_dim, mm_dim)), ("linear_out.bias", (output_dim,))]
    return spec


def make_block_params(seed: int, **kw) -> Params:
    g = torch.Generator().manual_seed(seed)
    P: Params = {}
    for name, shape in block_param_spec(**kw):
        fan_in = shape[-1] if len(shape) > 1 else 64
        P[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in) * (0.2 if name.endswith(".bias") else 1.0)
    return P


def block_fusion(P: Params, x0: torch.Tensor, x1: torch.Tensor, chunks: int = 20, rank: int = 15) -> torch.Tensor:
    """XModules.py:521-555 (Block.forward, pos_norm='before_cat', no dropout): bilinear fusion of the two pooled
    branch outputs (modeling_unimo.py:871-884).  x0, x1: [B, 768] -> [B, 768]."""
    a = F.linear(x0, P["linear0.weight"], P["linear0.bias"])                # :522
    b = F.linear(x1, P["linear1.weight"], P["linear1.bias"])                # :523
    size = a.shape[1] // chunks
    zs = []
    for c in range(chunks):                                                 # :531-545
        m = F.linear(a[:, c * size:(c + 1) * size], P[f"merge_linears0.{c}.weight"], P[f"merge_linears0.{c}.bias"]) * \
            F.linear(b[:, c * size:(c + 1) * size], P[f"merge_linears1.{c}.weight"], P[f"merge_linears1.{c}.bias"])
        z = m.view(a.shape[0], rank, -1).sum(1)                             # :539-540
        z = torch.sqrt(F.relu(z)) - torch.sqrt(F.relu(-z))                  # :542
        zs.append(F.normalize(z, p=2))                                      # :543
    z = torch.cat(zs, 1)
    return F.linear(z, P["linear_out.weight"], P["linear_out.bias"])        # :552


def is_zero_grad_param(name: str, training: bool) -> bool:
    """Parameters whose gradient is mathematically zero (softmax shift invariance of a key-projection bias;
    a bias in front of train-mode BatchNorm): both sides only hold rounding noise."""
    return name.endswith(ZERO_GRAD_SUFFIXES) or (training and name.endswith(".SAF_module.attn_sim_w.bias"))


def is_dead_param(name: str) -> bool:
    """Parameters that never receive a gradient in the reference."""
    return name.endswith(DEAD_PARAM_SUFFIXES) or name.startswith(("path_mapping.", "bn."))
