"""GPU tests of the HBM-bound kernels (router, aggregation, softmax, norms, FiLM, attention
filtration, GESC gate) through the C ABI, against plain torch fp32/fp64 on the same device."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def tol(dtype):
    return 2e-2 if dtype == torch.bfloat16 else 2e-5


def close(a, b, rel, abs_=1e-6):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() <= rel * b.abs().max().item() + abs_


@pytest.mark.parametrize("dtype", DTYPES)
def test_cast_axpby(dtype):
    from d2r_b200 import kernels as K
    x = torch.randn(1000, 37, device="cuda")
    y = K.cast(x, dtype)
    assert torch.equal(y, x.to(dtype))
    z = torch.randn_like(x).to(dtype)
    r = K.axpby(y, z, 2.0, -0.5)
    assert close(r, 2.0 * y.float() - 0.5 * z.float(), tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", ["none", "relu", "tanh"])
def test_bias_act_bwd(dtype, act):
    from d2r_b200 import kernels as K, _lib as L
    rows, cols = 777, 768
    yv = torch.randn(rows, cols, device="cuda")
    y = (torch.relu(yv) if act == "relu" else torch.tanh(yv) if act == "tanh" else yv).to(dtype)
    dy = torch.randn(rows, cols, device="cuda").to(dtype)
    dz, db = K.bias_act_bwd(dy, y, L.ACT[act], True, True)
    yf, g = y.float(), dy.float()
    ref = g * (yf > 0) if act == "relu" else g * (1 - yf * yf) if act == "tanh" else g
    assert close(dz, ref, tol(dtype))
    assert close(db, ref.sum(0), 5e-3 if dtype == torch.bfloat16 else 1e-4, 1e-3)


@pytest.mark.parametrize("cols", [50, 128, 197, 768])
def test_softmax(cols):
    from d2r_b200 import kernels as K
    rows, ld = 333, (cols + 7) // 8 * 8
    x = torch.randn(rows, ld, device="cuda") * 3
    scale = 100.0 / math.sqrt(768)
    ref = torch.softmax(scale * x[:, :cols].double(), -1)
    y = K.softmax_fwd(x, cols, scale, torch.float32)
    assert close(y[:, :cols], ref, 1e-5)
    yb = K.softmax_fwd(x, cols, scale, torch.bfloat16)
    assert close(yb[:, :cols], ref, 1e-2)
    dy = torch.randn(rows, ld, device="cuda")
    dx = K.softmax_bwd(y, dy, cols, scale, torch.float32)
    xr = x[:, :cols].double().requires_grad_(True)
    torch.softmax(scale * xr, -1).backward(dy[:, :cols].double())
    assert close(dx[:, :cols], xr.grad, 1e-4)
    dxb = K.softmax_bwd(yb, dy, cols, scale, torch.bfloat16)
    assert close(dxb[:, :cols], xr.grad, 3e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_l2norm(dtype):
    from d2r_b200 import kernels as K
    x = torch.randn(500, 768, device="cuda").to(dtype)
    y, rn = K.l2norm_fwd(x)
    xr = x.double().requires_grad_(True)
    ref = xr / (xr.pow(2).sum(-1, keepdim=True).sqrt() + 1e-8)
    assert close(y, ref, tol(dtype))
    dy = torch.randn_like(x)
    ref.backward(dy.double())
    dx = K.l2norm_bwd(y, dy, rn)
    assert close(dx, xr.grad, 2 * tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
def test_film_sqdiff(dtype):
    from d2r_b200 import kernels as K
    rows, D = 300, 768
    x = torch.randn(rows, D, device="cuda").to(dtype)
    st = torch.cat([torch.tanh(torch.randn(rows, D, device="cuda")), torch.randn(rows, D, device="cuda")], 1).to(dtype)
    m = K.film_fwd(x, st)
    s, t = st[:, :D].float(), st[:, D:].float()
    assert close(m, x.float() * s + t, tol(dtype))
    dm = torch.randn(rows, D, device="cuda").to(dtype)
    dx, dst = K.film_bwd(dm, x, st)
    assert close(dx, dm.float() * s, tol(dtype))
    assert close(dst[:, :D], dm.float() * x.float() * (1 - s * s), tol(dtype))
    assert close(dst[:, D:], dm.float(), tol(dtype))
    g = K.sqdiff_bwd(dm, x)
    assert close(g, 2 * dm.float() * x.float(), tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
def test_pool_mean(dtype):
    from d2r_b200 import kernels as K
    xs = [torch.randn(5, 50, 768, device="cuda").to(dtype) for _ in range(6)]
    p = K.pool_mean(xs)
    ref = torch.stack([x.float().mean(1) for x in xs])
    assert close(p, ref, 1e-5)
    dx = K.pool_mean_bwd(p[0].contiguous(), 50, dtype)
    assert close(dx, (p[0] / 50).unsqueeze(1).expand(5, 50, 768), tol(dtype))


@pytest.mark.parametrize("final", [False, True])
@pytest.mark.parametrize("Kc", [4, 6])
def test_router_head(final, Kc):
    from d2r_b200 import kernels as K
    B, H = 9, 768
    n_out = 1 if final else Kc
    hid = torch.relu(torch.randn(Kc, B, H, device="cuda"))
    w2 = [torch.randn(n_out, H, device="cuda") / 28 for _ in range(Kc)]
    b2 = [torch.full((n_out,), 0.7, device="cuda") + 0.3 * torch.randn(n_out, device="cuda") for _ in range(Kc)]
    b2[1][0] = -9.0   # a dead path
    hd = hid.double().requires_grad_(True)
    w2d = [w.double().requires_grad_(True) for w in w2]
    b2d = [b.double().requires_grad_(True) for b in b2]
    raw_ref = torch.stack([torch.relu(torch.tanh(hd[j] @ w2d[j].t() + b2d[j])) for j in range(Kc)], dim=2)
    norm_ref = raw_ref if final else raw_ref / (raw_ref.sum(-1, keepdim=True) + 1e-8)
    raw, norm, gate = K.router_head_fwd(hid, w2, b2, n_out, final)
    assert close(raw, raw_ref, 1e-5) and close(norm, norm_ref, 1e-5)
    gate_ref = (raw_ref[:, 0, :] < 1e-4 / Kc).double() if final else (raw_ref.sum(-1) < 1e-4).double()
    assert torch.equal(gate.double(), gate_ref)
    d_norm = torch.randn(B, n_out, Kc, device="cuda")
    norm_ref.backward(d_norm.double())
    d_hid, d_logit, d_w2, d_b2 = K.router_head_bwd(d_norm, raw, hid, w2, final)
    assert close(d_hid, hd.grad * (hid > 0), 1e-4)   # kernel applies the hidden ReLU mask
    for j in range(Kc):
        assert close(d_w2[j], w2d[j].grad, 1e-4) and close(d_b2[j], b2d[j].grad, 1e-4)


def _agg_ref(x0, embs, P, gate, final, inputs):
    Kc = len(embs)
    e = [torch.relu(x0)] + list(embs[1:])
    if not final:
        outs = []
        for i in range(Kc):
            r = sum(P[:, i, j].view(-1, 1, 1) * e[j] for j in range(Kc)) + gate[:, i].view(-1, 1, 1) * e[0]
            outs.append(r)
        return outs
    r = sum(P[:, 0, j].view(-1, 1, 1) * e[j] + gate[:, j].view(-1, 1, 1) * inputs[j] for j in range(Kc))
    return [r / (gate.sum(-1) + P[:, 0].sum(-1)).view(-1, 1, 1)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("final", [False, True])
@pytest.mark.parametrize("Kc", [4, 6])
def test_aggregate(dtype, final, Kc):
    from d2r_b200 import kernels as K
    B, Ln, D = 5, 19, 768
    n_out = 1 if final else Kc
    bcast = [1, 5] if Kc == 6 else [1]
    x0 = torch.randn(B, Ln, D, device="cuda").to(dtype)
    full = [x0] + [None if j in bcast else torch.randn(B, Ln, D, device="cuda").to(dtype) for j in range(1, Kc)]
    bvec = [torch.randn(B, D, device="cuda") if j in bcast else None for j in range(Kc)]
    inputs = [x0] + [torch.randn(B, Ln, D, device="cuda").to(dtype) for _ in range(1, Kc)]
    P = torch.rand(B, n_out, Kc, device="cuda")
    if final:
        P[0, 0, 2] = 0.0
        P[1, 0, 3] = 1e-6
        gate = (P[:, 0, :] < 1e-4 / Kc).float()
    else:
        P = P / P.sum(-1, keepdim=True)
        gate = torch.zeros(B, n_out, device="cuda")
        gate[2, 1] = 1.0
    # reference in fp64 with autograd
    x0d = x0.double().requires_grad_(True)
    embd = [None] + [(full[j].double() if full[j] is not None else bvec[j].double().unsqueeze(1).expand(B, Ln, D))
                     .detach().clone().requires_grad_(True) for j in range(1, Kc)]
    inpd = [x0d] + [t.double().requires_grad_(True) for t in inputs[1:]]
    Pd = P.double().requires_grad_(True)
    ref = _agg_ref(x0d, embd, Pd, gate.double(), final, inpd)
    outs, pooled = K.aggregate_fwd(full, bvec, P, gate, final, inputs if final else None)
    for o, r in zip(outs, ref):
        assert close(o, r, tol(dtype))
    d_outs = [torch.randn(B, Ln, D, device="cuda").to(dtype) for _ in range(n_out)]
    d_pooled = None
    loss = sum((r * g.double()).sum() for r, g in zip(ref, d_outs))
    if not final:
        assert close(pooled, torch.stack([r.mean(1) for r in ref]), 1e-2 if dtype == torch.bfloat16 else 1e-5)
        d_pooled = torch.randn(n_out, B, D, device="cuda")
        loss = loss + sum((r.mean(1) * d_pooled[i].double()).sum() for i, r in enumerate(ref))
    loss.backward()
    d_full, d_bvec, dP = K.aggregate_bwd(full, bvec, P, gate, final, d_outs, d_pooled, inputs if final else None)
    d_inputs = [None] * Kc
    if final:
        d_inputs = [None] + [torch.randn(B, Ln, D, device="cuda").to(dtype) for _ in range(1, Kc)]
        d_inputs[Kc - 1] = torch.full((B, Ln, D), float("nan"), device="cuda").to(dtype)   # overwritten (mask bit clear)
        before = [None if t is None else t.clone() for t in d_inputs]
        K.gate_skip_bwd(d_outs[0], P, gate, d_inputs, ((1 << Kc) - 1) & ~1 & ~(1 << (Kc - 1)))
        for j in range(1, Kc - 1):
            d_inputs[j] = d_inputs[j].float() - before[j].float()
    t2 = 3 * tol(dtype)
    assert close(d_full[0], x0d.grad, t2)
    for j in range(1, Kc):
        if full[j] is not None:
            assert close(d_full[j], embd[j].grad, t2), j
        else:
            assert close(d_bvec[j], embd[j].grad.sum(1), t2), j
        if final:
            assert close(d_inputs[j], inpd[j].grad, t2, 1e-5), j
    assert close(dP, Pd.grad, 2e-2 if dtype == torch.bfloat16 else 1e-4, 1e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("training", [True, False])
def test_attention_filtration(dtype, training):
    from d2r_b200 import kernels as K
    B, Ln, D = 6, 21, 768
    sg = torch.randn(B, D, device="cuda").to(dtype)
    sl = torch.randn(B, Ln, D, device="cuda").to(dtype)
    w = (torch.randn(D, device="cuda") / 10)
    bias = torch.tensor([0.1], device="cuda")
    bn_w, bn_b = torch.tensor([1.3], device="cuda"), torch.tensor([-0.2], device="cuda")
    rm, rv = torch.tensor([0.05], device="cuda"), torch.tensor([0.8], device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    rm0, rv0 = rm.clone(), rv.clone()
    out, saved = K.saf_fwd(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, training)
    # reference (fp64)
    S = torch.cat([sg.unsqueeze(1), sl], 1).double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    gw, gb = bn_w.double().requires_grad_(True), bn_b.double().requires_grad_(True)
    rmr, rvr = rm0.double().clone(), rv0.double().clone()
    logit = (S @ wd + bd).unsqueeze(1)                                # (B,1,L+1)
    y = F.batch_norm(logit, rmr, rvr, gw, gb, training, 0.1, 1e-5)
    a = torch.sigmoid(y)
    a = a / (a.abs().sum(-1, keepdim=True) + 1e-8)
    saf = (a @ S).squeeze(1)
    ref = saf / (saf.pow(2).sum(-1, keepdim=True).sqrt() + 1e-8)
    assert close(out, ref, tol(dtype))
    if training:
        assert close(rm, rmr, 1e-4) and close(rv, rvr, 1e-4) and int(nbt.item()) == 1
    else:
        assert torch.equal(rm, rm0) and torch.equal(rv, rv0)
    d_out = torch.randn(B, D, device="cuda")
    ref.backward(d_out.double())
    d_sg, d_sl, d_w, d_bias, d_bn_w, d_bn_b = K.saf_bwd(d_out, sg, sl, w, bias, bn_w, bn_b, rm, rv, training, saved)
    t2 = 3 * tol(dtype)
    assert close(d_sg, S.grad[:, 0], t2) and close(d_sl, S.grad[:, 1:], t2)
    assert close(d_w, wd.grad, t2, 1e-5)
    assert close(d_bn_w, gw.grad, t2, 1e-5) and close(d_bn_b, gb.grad, t2, 1e-5)
    if not training:
        assert close(d_bias, bd.grad, t2, 1e-5)


def test_gate_fuse():
    from d2r_b200 import kernels as K
    B, D = 7, 768
    gl, t, i = [torch.randn(B, D, device="cuda") for _ in range(3)]
    gd, td, idd = [v.double().requires_grad_(True) for v in (gl, t, i)]
    g_ref = torch.softmax(gd, -1)
    ref = g_ref * td + (1 - g_ref) * idd
    g, out = K.gate_fuse_fwd(gl, t, i)
    assert close(out, ref, 1e-5) and close(g, g_ref, 1e-5)
    d_out = torch.randn(B, D, device="cuda")
    ref.backward(d_out.double())
    d_gl, d_t, d_i = K.gate_fuse_bwd(d_out, g, t, i)
    assert close(d_gl, gd.grad, 1e-4) and close(d_t, td.grad, 1e-5) and close(d_i, idd.grad, 1e-5)


@pytest.mark.parametrize("rows,cols,get_softmax", [(256, 256, True), (8, 8, True), (33, 1000, True), (64, 64, False)])
def test_js_div_vs_reference_formula(rows, cols, get_softmax):
    """Fused JS-divergence kernels (XModules.py:32-41, the loss on sim_paths) vs the reference formula under torch
    autograd in fp64; logits at the scale of sim_paths Grams (tens)."""
    from d2r_b200.interaction.XModules import js_div
    from oracle import d2r_oracle as O
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    a = torch.randn(rows, cols, device="cuda", generator=g) * 4
    b = torch.randn(rows, cols, device="cuda", generator=g) * 4
    if not get_softmax:
        a, b = torch.softmax(a, -1), torch.softmax(b, -1)
    p1, q1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    p2, q2 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    l1 = js_div(p1, q1, get_softmax)
    l2 = O.js_div(p2, q2, get_softmax)
    (-0.7 * l1).backward()
    (-0.7 * l2).backward()
    assert abs(l1.item() - l2.item()) <= 2e-5 * max(1.0, abs(l2.item()))
    for got, ref in ((p1.grad, p2.grad), (q1.grad, q2.grad)):
        assert (got.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-9
    with pytest.raises(RuntimeError):
        js_div(a.cpu(), b.cpu())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_block_merge_kernels(dtype):
    """Rank-sum / signed sqrt / per-chunk L2 norm of XModules.Block (:538-543) and its backward against a torch fp64
    restatement.  m1 = m0 * positive factors keeps r = sum_k m0 m1 away from zero, where the map is well
    conditioned (its derivative 1 / (2 sqrt|r|) is unbounded at 0)."""
    from d2r_b200 import kernels as K
    B, C, R, S = 9, 20, 15, 80
    g = torch.Generator(device="cuda").manual_seed(3)
    m0 = torch.randn(B, C * R * S, device="cuda", generator=g).to(dtype)
    m1 = (m0.float() * (0.5 + torch.rand(B, C * R * S, device="cuda", generator=g))).to(dtype)
    sign = torch.where(torch.rand(B, C, 1, 1, device="cuda", generator=g) < 0.5, -1.0, 1.0)   # negative r too
    m1 = (m1.float().view(B, C, R, S) * sign).reshape(B, -1).to(dtype)
    a, b = m0.double().requires_grad_(True), m1.double().requires_grad_(True)
    r_ref = (a * b).view(B, C, R, S).sum(2)
    zs = torch.sqrt(torch.relu(r_ref)) - torch.sqrt(torch.relu(-r_ref))
    z_ref = torch.nn.functional.normalize(zs, p=2, dim=-1).reshape(B, C * S)
    dz = torch.randn(B, C * S, device="cuda", generator=g).to(dtype)
    (z_ref * dz.double()).sum().backward()
    z, r, inv = K.block_merge_fwd(m0, m1, C, R, S)
    dm0, dm1 = K.block_merge_bwd(dz, m0, m1, r, inv, C, R, S)
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert (z.double() - z_ref).abs().max().item() <= tol
    assert (r.double() - r_ref.reshape(B, -1)).abs().max().item() <= 1e-4 * r_ref.abs().max().item()
    for got, ref in ((dm0, a.grad), (dm1, b.grad)):
        assert (got.double() - ref).abs().max().item() <= tol * ref.abs().max().item() + 1e-7
