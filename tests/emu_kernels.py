"""TEST INFRASTRUCTURE ONLY: torch-CPU emulation of the wrappers in ``d2r_b200/kernels.py``.

The authoring container has no GPU, so ``tests/test_stack_emulated.py`` patches ``d2r_b200.kernels`` with
these functions to execute the *product's Python orchestration* (stack.py, autograd.py, the nn.Module
mirror: launch order, strides, saved tensors, the hand-written backward) on the CPU and compare it with
the oracle.  Nothing in ``d2r_b200/`` imports this file, and it never runs on the GPU box's product
path; the real kernels are tested against torch on the GPU in test_kernels_gpu.py / test_gemm_gpu.py.
The emulation honours the same raw-stride contract as the C ABI (pointers = tensor views, explicit
lds / batch strides), which is exactly what the orchestration can get wrong.
"""
import torch

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
EPI_STD, EPI_SQDIFF, EPI_SOFTMAX, EPI_SOFTMAX_BWD = 0, 1, 2, 3


def _view(t, sizes, strides):
    return torch.as_strided(t, sizes, strides, t.storage_offset())


def _act(v, act):
    return torch.relu(v) if act == ACT_RELU else torch.tanh(v) if act == ACT_TANH else v


def gemm(a, b, c, *, m, n, k, lda, ldb, ldc, a_mn=False, b_mn=False, batch=1, batch_inner=1, a_str=(0, 0),
         b_str=(0, 0), c_str=(0, 0), alpha=1.0, bias=None, bias_sz=0, act=ACT_NONE, residual=None, ldr=0,
         r_str=(0, 0), epilogue=EPI_STD, c2=None, accumulate=False, split_k=1, tile_n=0, act_cols=0):
    assert a.dtype == b.dtype
    if a.dtype == torch.bfloat16:
        assert lda % 8 == 0 and ldb % 8 == 0 and all(s % 8 == 0 for s in a_str + b_str), "TMA 16-byte rule"
        assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
    bo, bi = batch // batch_inner, batch_inner
    A = _view(a, (bo, bi, k, m), (a_str[0], a_str[1], lda, 1)).transpose(-1, -2) if a_mn else \
        _view(a, (bo, bi, m, k), (a_str[0], a_str[1], lda, 1))
    Bm = _view(b, (bo, bi, k, n), (b_str[0], b_str[1], ldb, 1)) if b_mn else \
        _view(b, (bo, bi, n, k), (b_str[0], b_str[1], ldb, 1)).transpose(-1, -2)
    v = alpha * (A.float() @ Bm.float())
    if bias is not None:
        if bias_sz:
            v = v + _view(bias, (bo, bi, 1, n), (bias_sz * bi, bias_sz, 0, 1))
        else:
            v = v + bias[:n]
    C = _view(c, (bo, bi, m, n), (c_str[0], c_str[1], ldc, 1))
    R = None if residual is None else _view(residual, (bo, bi, m, n), (r_str[0], r_str[1], ldr, 1)).float()
    if epilogue == EPI_SQDIFF:
        d = R - v
        _view(c2, (bo, bi, m, n), (c_str[0], c_str[1], ldc, 1)).copy_(d.to(c2.dtype))
        v = d * d
    elif epilogue == EPI_SOFTMAX:
        assert a.dtype == torch.bfloat16 and c.dtype == torch.bfloat16 and n <= 256 and bias is None
        v = torch.softmax(v, -1)
    elif epilogue == EPI_SOFTMAX_BWD:
        assert a.dtype == torch.bfloat16 and c.dtype == torch.bfloat16 and n <= 256 and R is not None
        v = R * (v - (v * R).sum(-1, keepdim=True))
    else:
        if act_cols:
            v = torch.cat([_act(v[..., :act_cols], act), v[..., act_cols:]], -1)
        else:
            v = _act(v, act)
        if R is not None:
            v = v + R
    if accumulate:
        for zo in range(bo):           # batches may alias the same C (atomic accumulation on the GPU)
            for zi in range(bi):
                C[zo, zi].add_(v[zo, zi].to(c.dtype))
    else:
        C.copy_(v.to(c.dtype))
    return c


def attn_fused_supported(dtype, Lc, hd):
    return dtype == torch.bfloat16 and Lc <= 128 and hd % 16 == 0 and (hd <= 64 or hd % 64 == 0)


def attn_fused_bwd_supported(dtype, Lq, Lc, hd):
    return attn_fused_supported(dtype, Lc, hd) and Lq <= 128


def _heads(t, B, rows, heads, dh, ld):
    """X[b, row, h*dh + c] with row stride ld -> view [B, heads, rows, dh]."""
    assert ld % 8 == 0 and t.data_ptr() % 16 == 0, "TMA 16-byte rule"
    return _view(t, (B, heads, rows, dh), (rows * ld, dh, ld, 1))


def attn_fused_fwd(q, q_ld, k, k_ld, v, v_ld, *, B, Lq, Lc, D, heads, alpha, p_ld, residual=None, mode=0, out2=None):
    dh = D // heads
    assert attn_fused_supported(q.dtype, Lc, dh) and p_ld % 8 == 0 and p_ld >= Lc
    Q, Kk, V = _heads(q, B, Lq, heads, dh, q_ld), _heads(k, B, Lc, heads, dh, k_ld), _heads(v, B, Lc, heads, dh, v_ld)
    Pf = torch.softmax(alpha * (Q.float() @ Kk.float().transpose(-1, -2)), -1)
    P = torch.zeros(B, heads, Lq, p_ld, dtype=torch.bfloat16)
    P[..., :Lc] = Pf.to(torch.bfloat16)
    o = (P[..., :Lc].float() @ V.float()).transpose(1, 2).reshape(B, Lq, D)
    if mode == 1:
        d = residual.float() - o
        if out2 is None:
            out2 = torch.empty(B, Lq, D, dtype=torch.bfloat16)
        out2.copy_(d.to(torch.bfloat16))
        o = d * d
    elif residual is not None:
        o = o + residual.float()
    return o.to(torch.bfloat16), P, out2


def attn_fused_bwd(dO, do_ld, sign, P, q, q_ld, k, k_ld, v, v_ld, dq, dq_ld, dk, dk_ld, dv, dv_ld, *, B, Lq, Lc, D,
                   heads, alpha):
    dh = D // heads
    assert attn_fused_bwd_supported(q.dtype, Lq, Lc, dh)
    G = sign * _heads(dO, B, Lq, heads, dh, do_ld).float()
    Q, Kk, V = (_heads(q, B, Lq, heads, dh, q_ld).float(), _heads(k, B, Lc, heads, dh, k_ld).float(),
                _heads(v, B, Lc, heads, dh, v_ld).float())
    Pf = P[..., :Lc].float()
    dP = G @ V.transpose(-1, -2)
    dS = (alpha * Pf * (dP - (dP * Pf).sum(-1, keepdim=True))).to(torch.bfloat16).float()
    _heads(dv, B, Lc, heads, dh, dv_ld).copy_((Pf.transpose(-1, -2) @ G).to(dv.dtype))
    _heads(dq, B, Lq, heads, dh, dq_ld).copy_((dS @ Kk).to(dq.dtype))
    _heads(dk, B, Lc, heads, dh, dk_ld).copy_((dS.transpose(-1, -2) @ Q).to(dk.dtype))


def softmax_fwd(x, cols, scale, out_dtype, ldy=None):
    ldy = ldy or x.shape[-1]
    y = torch.full(x.shape[:-1] + (ldy,), float("nan"), dtype=out_dtype)
    y[..., :cols] = torch.softmax(scale * x[..., :cols].float(), -1).to(out_dtype)
    return y


def softmax_bwd(y, dy, cols, scale, out_dtype):
    yv, g = y[..., :cols].float(), dy[..., :cols].float()
    dx = torch.full(y.shape, float("nan"), dtype=out_dtype)
    dx[..., :cols] = (scale * yv * (g - (yv * g).sum(-1, keepdim=True))).to(out_dtype)
    return dx


def cast(x, dtype, out=None):
    if out is None:
        return x.to(dtype).contiguous()
    out.copy_(x.to(dtype))
    return out


def bias_act_bwd(dy, y, act, want_dz, want_db):
    g = dy.float()
    if act == ACT_RELU:
        g = g * (y.float().reshape(dy.shape) > 0)
    elif act == ACT_TANH:
        g = g * (1 - y.float().reshape(dy.shape) ** 2)
    dz = g.to(dy.dtype) if (want_dz and act != ACT_NONE) else None
    db = g.reshape(-1, dy.shape[-1]).sum(0) if want_db else None
    return (dz if dz is not None else dy), db


def l2norm_fwd(x):
    xf = x.float()
    r = 1.0 / (xf.pow(2).sum(-1).sqrt() + 1e-8)
    return (xf * r.unsqueeze(-1)).to(x.dtype), r.reshape(-1)


def l2norm_bwd(y, dy, rn):
    yf, g = y.float(), dy.float()
    r = rn.view(y.shape[:-1]).unsqueeze(-1)
    n = (1.0 / r - 1e-8).clamp_min(1e-30)
    return (r * g - yf * (yf * g).sum(-1, keepdim=True) / n).to(y.dtype)


def film_fwd(x, st):
    D = x.shape[-1]
    return (x.float() * st[..., :D].float() + st[..., D:].float()).to(x.dtype)


def film_bwd(dm, x, st, add=None):
    D = x.shape[-1]
    g, s = dm.float().reshape(x.shape), st[..., :D].float()
    dx = g * s + (add.float().reshape(x.shape) if add is not None else 0)
    dst = torch.cat([g * x.float() * (1 - s * s), g], -1)
    return dx.to(x.dtype), dst.to(st.dtype)


def mul(x, z, alpha=1.0):
    return (alpha * x.float() * z.float()).to(x.dtype)


def axpby(x, z, a, b):
    return (a * x.float() + (b * z.float() if z is not None else 0)).to(x.dtype)


def sqdiff_bwd(dsq, d, add=None, want_gx=False):
    g = 2 * dsq.float().reshape(d.shape) * d.float()
    if not want_gx:
        return g.to(d.dtype)
    gx = g + (add.float().reshape(d.shape) if add is not None else 0)
    return g.to(d.dtype), gx.to(d.dtype)


def pool_mean(xs):
    return torch.stack([x.float().mean(1) for x in xs])


def pool_mean_bwd(d_pooled, Ln, dtype):
    return (d_pooled / Ln).unsqueeze(1).expand(-1, Ln, -1).to(dtype).contiguous()


def pool_mean_bwd_into(d_pooled, dx):
    dx.add_((d_pooled / dx.shape[1]).unsqueeze(1).to(dx.dtype))
    return dx


def gate_fuse_fwd(gl, t, i):
    g = torch.softmax(gl, -1)
    return g, g * t + (1 - g) * i


def _js_terms(p, q, get_softmax):
    P = torch.softmax(p, -1) if get_softmax else p
    Q = torch.softmax(q, -1) if get_softmax else q
    lm = (0.5 * (P + Q)).log()
    gp = torch.where(P > 0, P.clamp_min(1e-45).log() - lm, torch.zeros_like(P))
    gq = torch.where(Q > 0, Q.clamp_min(1e-45).log() - lm, torch.zeros_like(Q))
    return P, Q, gp, gq


def js_div_fwd(p, q, get_softmax=True):
    P, Q, gp, gq = _js_terms(p.float(), q.float(), get_softmax)
    return ((P * gp).sum() + (Q * gq).sum()) * (0.5 / p.shape[0])


def js_div_bwd(p, q, d_loss, get_softmax=True):
    P, Q, gp, gq = _js_terms(p.float(), q.float(), get_softmax)
    up = d_loss * (0.5 / p.shape[0])
    if get_softmax:
        return (up * P * (gp - (P * gp).sum(-1, keepdim=True)), up * Q * (gq - (Q * gq).sum(-1, keepdim=True)))
    return up * gp, up * gq


def block_merge_fwd(m0, m1, chunks, rank, size):
    B = m0.shape[0]
    r = (m0.float() * m1.float()).view(B, chunks, rank, size).sum(2)
    zs = torch.sign(r) * r.abs().sqrt()
    inv = 1.0 / zs.norm(dim=-1).clamp_min(1e-12)
    return (zs * inv.unsqueeze(-1)).reshape(B, chunks * size).to(m0.dtype), r.reshape(B, chunks * size), inv


def block_merge_bwd(dz, m0, m1, r, inv, chunks, rank, size):
    B = m0.shape[0]
    rv = r.view(B, chunks, size)
    zn = torch.sign(rv) * rv.abs().sqrt() * inv.unsqueeze(-1)
    g = dz.float().view(B, chunks, size)
    dzs = (g - zn * (zn * g).sum(-1, keepdim=True)) * inv.unsqueeze(-1)
    sq = rv.abs().sqrt()
    dr = torch.where(sq > 0, dzs * 0.5 / sq.clamp_min(1e-30), torch.zeros_like(sq)).unsqueeze(2)
    dm0 = (dr * m1.float().view(B, chunks, rank, size)).reshape(B, -1).to(m0.dtype)
    dm1 = (dr * m0.float().view(B, chunks, rank, size)).reshape(B, -1).to(m0.dtype)
    return dm0, dm1


def gate_fuse_bwd(d_out, g, t, i):
    dg = d_out * (t - i)
    return g * (dg - (dg * g).sum(-1, keepdim=True)), d_out * g, d_out * (1 - g)


def router_head_fwd(hid, w2, b2, n_out, final_layer):
    Kc = hid.shape[0]
    raw = torch.stack([torch.relu(torch.tanh(hid[j] @ w2[j].t() + b2[j])) for j in range(Kc)], dim=2)
    if final_layer:
        return raw, raw.clone(), (raw[:, 0, :] < 1e-4 / Kc).float()
    s = raw.sum(-1, keepdim=True)
    return raw, raw / (s + 1e-8), (s.squeeze(-1) < 1e-4).float()


def router_head_bwd(d_norm, raw, hid, w2, final_layer):
    Kc = hid.shape[0]
    if final_layer:
        d_raw = d_norm
    else:
        inv = 1.0 / (raw.sum(-1, keepdim=True) + 1e-8)
        d_raw = d_norm * inv - (d_norm * raw).sum(-1, keepdim=True) * inv * inv
    dl = torch.where(raw > 0, d_raw * (1 - raw * raw), torch.zeros_like(raw))      # [B,n_out,K]
    d_hid = torch.stack([(dl[:, :, j] @ w2[j]) * (hid[j] > 0) for j in range(Kc)])
    d_w2 = [dl[:, :, j].t() @ hid[j] for j in range(Kc)]
    d_b2 = [dl[:, :, j].sum(0) for j in range(Kc)]
    return d_hid, dl, d_w2, d_b2


def _embs(full, bvec, B, Ln, D):
    e = []
    for j, f in enumerate(full):
        v = f.float() if f is not None else bvec[j].float().unsqueeze(1).expand(B, Ln, D)
        e.append(torch.relu(v) if j == 0 else v)
    return e


def aggregate_fwd(full, bvec, P, gate, final_layer, inputs=None, want_pooled=True):
    B, Ln, D = full[0].shape
    Kc = len(full)
    e = _embs(full, bvec, B, Ln, D)
    dt = full[0].dtype
    if final_layer:
        N = sum(P[:, 0, j].view(B, 1, 1) * e[j] for j in range(Kc))
        N = N + sum(gate[:, j].view(B, 1, 1) * inputs[j].float() for j in range(Kc))
        S = (P[:, 0].sum(-1) + gate.sum(-1)).view(B, 1, 1)
        return [(N / S).to(dt)], None
    outs = [sum(P[:, i, j].view(B, 1, 1) * e[j] for j in range(Kc)) + gate[:, i].view(B, 1, 1) * e[0]
            for i in range(Kc)]
    pooled = torch.stack([o.mean(1) for o in outs]) if want_pooled else None
    return [o.to(dt) for o in outs], pooled


@torch.enable_grad()
def aggregate_bwd(full, bvec, P, gate, final_layer, d_outs, d_pooled, inputs=None):
    B, Ln, D = full[0].shape
    Kc = len(full)
    dt = full[0].dtype
    x0 = full[0].float().detach().requires_grad_(True)
    leaves = [x0] + [(f.float() if f is not None else bvec[j].float()).detach().requires_grad_(True)
                     for j, f in enumerate(full) if j > 0]
    e = [torch.relu(x0)] + [leaves[j] if full[j] is not None else leaves[j].unsqueeze(1).expand(B, Ln, D)
                            for j in range(1, Kc)]
    Pd = P.detach().clone().requires_grad_(True)
    ins = None
    if final_layer:
        ins = [x0] + [t.float().detach().requires_grad_(True) for t in inputs[1:]]
        N = sum(Pd[:, 0, j].view(B, 1, 1) * e[j] + gate[:, j].view(B, 1, 1) * ins[j] for j in range(Kc))
        outs = [N / (Pd[:, 0].sum(-1) + gate.sum(-1)).view(B, 1, 1)]
    else:
        outs = [sum(Pd[:, i, j].view(B, 1, 1) * e[j] for j in range(Kc)) + gate[:, i].view(B, 1, 1) * e[0]
                for i in range(Kc)]
    loss = sum((o * g.float()).sum() for o, g in zip(outs, d_outs))
    if d_pooled is not None:
        loss = loss + sum((o.mean(1) * d_pooled[i]).sum() for i, o in enumerate(outs))
    loss.backward()
    d_full = [leaves[j].grad.to(dt) if full[j] is not None else None for j in range(Kc)]
    d_bvec = [leaves[j].grad if full[j] is None else None for j in range(Kc)]
    return d_full, d_bvec, Pd.grad


def gate_skip_bwd(d_out, P, gate, dxs, accumulate_mask):
    B = d_out.shape[0]
    S = (P[:, 0].sum(-1) + gate.sum(-1)).view(B, 1, 1)
    for j, dx in enumerate(dxs):
        if j == 0 or dx is None:
            continue
        v = (gate[:, j].view(B, 1, 1) / S * d_out.float()).to(dx.dtype)
        if (accumulate_mask >> j) & 1:
            dx.add_(v)
        else:
            dx.copy_(v)


def _saf_ref(sg, sl, w, bias, bn_w, bn_b, rm, rv, training):
    import torch.nn.functional as F
    S = torch.cat([sg.unsqueeze(1), sl], 1).float()
    logit = (S @ w + bias).unsqueeze(1)
    y = F.batch_norm(logit, rm, rv, bn_w, bn_b, training, 0.1, 1e-5)
    a = torch.sigmoid(y)
    a = a / (a.abs().sum(-1, keepdim=True) + 1e-8)
    saf = (a @ S).squeeze(1)
    return saf / (saf.pow(2).sum(-1, keepdim=True).sqrt() + 1e-8)


def saf_fwd(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, training):
    rm0, rv0 = rm.clone(), rv.clone()
    out = _saf_ref(sg, sl, w, bias, bn_w, bn_b, rm, rv, training)      # updates rm / rv in place when training
    if training and nbt is not None:
        nbt.add_(1)
    return out.detach(), (rm0, rv0)


@torch.enable_grad()
def saf_bwd(d_out, sg, sl, w, bias, bn_w, bn_b, rm, rv, training, saved):
    rm0, rv0 = saved
    leaves = [t.detach().float().clone().requires_grad_(True) for t in (sg, sl, w, bias, bn_w, bn_b)]
    out = _saf_ref(*leaves, rm0.clone(), rv0.clone(), training)
    out.backward(d_out)
    g = [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]
    return g[0].to(sg.dtype), g[1].to(sl.dtype), g[2], g[3], g[4], g[5]


def zeros_f32(shape, device):
    return torch.zeros(shape, device=device, dtype=torch.float32)


def zero_arena_reset():
    pass
