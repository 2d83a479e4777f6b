"""Fused attention kernel (d2r_attn_fwd) against a torch fp32 reference of the same op on the same bf16 inputs,
and against the composed two-GEMM path it replaces (which the goldens already pin).  Tolerance: the kernel rounds
P to bf16 before the second product exactly like the composed path; outputs are bf16 -> 2^-8 relative."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def ref_attention(q, k, v, heads, alpha):
    B, Lq, D = q.shape
    Lc = k.shape[1]
    dh = D // heads
    qh = q.float().view(B, Lq, heads, dh).transpose(1, 2)
    kh = k.float().view(B, Lc, heads, dh).transpose(1, 2)
    vh = v.float().view(B, Lc, heads, dh).transpose(1, 2)
    P = torch.softmax(alpha * qh @ kh.transpose(-1, -2), dim=-1)
    out = (P.to(bf).float() @ vh).transpose(1, 2).reshape(B, Lq, D)
    return out, P


def close(a, b, tol):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() <= tol * (b.abs().max() + 1e-6)).item()


@pytest.mark.parametrize("B,heads,Lq,Lc,D,alpha", [
    (3, 1, 128, 50, 768, 100.0 / math.sqrt(768)),     # cross-modal attention, text branch
    (3, 1, 50, 128, 768, 100.0 / math.sqrt(768)),     # image branch
    (2, 1, 128, 128, 768, 1.0),                        # CRCMC second attention (unscaled)
    (2, 1, 50, 50, 768, 1.0),
    (2, 16, 128, 128, 768, 1.0 / math.sqrt(48)),       # 16-head self-attention
    (2, 16, 50, 50, 768, 1.0 / math.sqrt(48)),
    (150, 1, 128, 50, 768, 3.6),                       # more units than SMs: persistent loop, both buffer parities
    (9, 16, 72, 72, 768, 0.15),                        # > 148 units with heads
    (2, 1, 200, 100, 768, 0.05),                       # two query tiles per sample, ragged
    (1, 1, 7, 3, 64, 1.0),                             # tiny
])
@pytest.mark.parametrize("mode", ["plain", "residual", "sqdiff"])
def test_attn_fused_forward(B, heads, Lq, Lc, D, alpha, mode):
    from d2r_b200 import kernels as K
    torch.manual_seed(B * 1000 + Lq * 7 + Lc)
    scale = 0.25 if alpha > 1.5 else 1.0           # keep the temperature-100 logits in a sane range
    q = (torch.randn(B, Lq, D, device="cuda") * scale).to(bf)
    kv = (torch.randn(B, Lc, 2 * D + 64, device="cuda") * scale).to(bf)     # k / v are column slices (ld = 2D + 64)
    k, v = kv[:, :, :D], kv[:, :, D + 64:]
    res = torch.randn(B, Lq, D, device="cuda").to(bf) if mode != "plain" else None
    Lcp = (Lc + 7) // 8 * 8
    out, P, out2 = K.attn_fused_fwd(q, D, k, kv.shape[2], v, kv.shape[2], B=B, Lq=Lq, Lc=Lc, D=D, heads=heads,
                                    alpha=alpha, p_ld=Lcp, residual=res, mode=1 if mode == "sqdiff" else 0)
    torch.cuda.synchronize()
    ref, Pref = ref_attention(q, k.contiguous(), v.contiguous(), heads, alpha)
    assert close(P[..., :Lc], Pref, 1e-2), (P[..., :Lc].float() - Pref).abs().max()
    if mode == "plain":
        assert close(out, ref, 1.5e-2)
    elif mode == "residual":
        assert close(out, ref + res.float(), 1.5e-2)
    else:
        d = res.float() - ref
        assert close(out2, d, 1.5e-2) and close(out, d * d, 3e-2)


def test_attn_fused_matches_composed_path():
    """Same inputs through stack.attn_fwd with the fused kernel and with the two-GEMM path."""
    import d2r_b200.stack as S
    torch.manual_seed(0)
    B, Lq, D, H = 5, 128, 768, 16
    qkv = torch.randn(B, Lq, 3 * D, device="cuda").to(bf)
    x = torch.randn(B, Lq, D, device="cuda").to(bf)
    args = (qkv, 3 * D, qkv[:, :, D:], 3 * D, qkv[:, :, 2 * D:], 3 * D, B, Lq, Lq, D, H, 1.0 / math.sqrt(D // H), bf)
    try:
        S.FUSED_ATTN = True
        o1, p1 = S.attn_fwd(*args, residual=x)
        S.FUSED_ATTN = False
        o2, p2 = S.attn_fwd(*args, residual=x)
    finally:
        S.FUSED_ATTN = True
    torch.cuda.synchronize()
    assert close(p1, p2, 1e-2) and close(o1, o2, 1e-2)


def test_attn_fused_rejects_unsupported_shapes():
    from d2r_b200 import kernels as K
    q = torch.zeros(1, 8, 768, device="cuda", dtype=bf)
    kv = torch.zeros(1, 200, 768, device="cuda", dtype=bf)
    with pytest.raises(RuntimeError):
        K.attn_fused_fwd(q, 768, kv, 768, kv, 768, B=1, Lq=8, Lc=200, D=768, heads=1, alpha=1.0, p_ld=200)


@pytest.mark.parametrize("B,heads,Lq,Lc,D,alpha,sign", [
    (3, 1, 128, 50, 768, 100.0 / math.sqrt(768), 1.0),
    (3, 1, 50, 128, 768, 100.0 / math.sqrt(768), -1.0),     # the alignment cell folds a minus into d_out
    (2, 1, 128, 128, 768, 1.0, 1.0),
    (2, 1, 50, 50, 768, 1.0, 1.0),
    (2, 16, 128, 128, 768, 1.0 / math.sqrt(48), 1.0),
    (2, 16, 50, 50, 768, 1.0 / math.sqrt(48), 1.0),
    (150, 1, 128, 50, 768, 3.6, 1.0),
    (10, 16, 72, 56, 768, 0.15, 1.0),
    (1, 1, 7, 3, 64, 1.0, 1.0),
])
def test_attn_fused_backward(B, heads, Lq, Lc, D, alpha, sign):
    """dq / dk / dv of the fused backward against torch autograd through the same attention (fp32 math on the same
    bf16 operands and the same bf16 P)."""
    from d2r_b200 import kernels as K
    torch.manual_seed(B * 31 + Lq + 3 * Lc)
    scale = 0.25 if alpha > 1.5 else 1.0
    q = (torch.randn(B, Lq, D, device="cuda") * scale).to(bf)
    kv = (torch.randn(B, Lc, 2 * D, device="cuda") * scale).to(bf)
    k, v = kv[:, :, :D], kv[:, :, D:]
    dO = torch.randn(B, Lq, D, device="cuda").to(bf)
    Lcp = (Lc + 7) // 8 * 8
    out, P, _ = K.attn_fused_fwd(q, D, k, 2 * D, v, 2 * D, B=B, Lq=Lq, Lc=Lc, D=D, heads=heads, alpha=alpha, p_ld=Lcp)
    dq = torch.full((B, Lq, D), float("nan"), device="cuda", dtype=bf)
    dkv = torch.full((B, Lc, 2 * D), float("nan"), device="cuda", dtype=bf)
    K.attn_fused_bwd(dO, D, sign, P, q, D, k, 2 * D, v, 2 * D, dq, D, dkv, 2 * D, dkv[:, :, D:], 2 * D,
                     B=B, Lq=Lq, Lc=Lc, D=D, heads=heads, alpha=alpha)
    torch.cuda.synchronize()
    # reference: autograd through softmax attention in fp32 on the same bf16 values
    dh = D // heads
    qf = q.float().requires_grad_(True)
    kf = k.float().contiguous().requires_grad_(True)
    vf = v.float().contiguous().requires_grad_(True)
    qh = qf.view(B, Lq, heads, dh).transpose(1, 2)
    kh = kf.view(B, Lc, heads, dh).transpose(1, 2)
    vh = vf.view(B, Lc, heads, dh).transpose(1, 2)
    o = (torch.softmax(alpha * qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(B, Lq, D)
    o.backward(sign * dO.float())
    def l2(a, b):
        return ((a.float() - b).norm() / (b.norm() + 1e-12)).item()
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    # bf16 P / dS / outputs: a few 1e-3 .. 1e-2 relative in L2
    assert l2(dkv[:, :, D:], vf.grad) <= 2e-2, ("dv", l2(dkv[:, :, D:], vf.grad))
    assert l2(dq, qf.grad) <= 3e-2, ("dq", l2(dq, qf.grad))
    assert l2(dkv[:, :, :D], kf.grad) <= 3e-2, ("dk", l2(dkv[:, :, :D], kf.grad))


def test_attn_fused_backward_matches_composed_path():
    import d2r_b200.stack as S
    torch.manual_seed(1)
    B, Lq, D, H = 4, 128, 768, 16
    qkv = torch.randn(B, Lq, 3 * D, device="cuda").to(bf)
    dy = torch.randn(B, Lq, D, device="cuda").to(bf)
    alpha = 1.0 / math.sqrt(D // H)
    res = []
    try:
        for fused in (True, False):
            S.FUSED_ATTN = fused
            _, P = S.attn_fwd(qkv, 3 * D, qkv[:, :, D:], 3 * D, qkv[:, :, 2 * D:], 3 * D, B, Lq, Lq, D, H, alpha, bf)
            dqkv = torch.empty_like(qkv)
            S.attn_bwd(dy, D, 1.0, P, qkv, 3 * D, qkv[:, :, D:], 3 * D, qkv[:, :, 2 * D:], 3 * D, dqkv, 3 * D,
                       dqkv[:, :, D:], 3 * D, dqkv[:, :, 2 * D:], 3 * D, B, Lq, Lq, D, H, alpha, bf)
            res.append(dqkv)
    finally:
        S.FUSED_ATTN = True
    torch.cuda.synchronize()
    a, b = res[0].float(), res[1].float()
    assert ((a - b).norm() / b.norm()).item() <= 1e-2
