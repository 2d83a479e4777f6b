"""GPU tests of d2r_gemm (C ABI) against torch fp32/fp64 matmul on the same device.

bf16 operands -> tcgen05 path; fp32 operands -> CUDA-core path.  Covers the four operand
major-ness combinations (forward / dgrad / wgrad), ragged sizes (TMA OOB fill + masked
epilogue), head-strided batching, split-K and the fused epilogues.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.double().transpose(-1, -2) if a_mn else a.double()
    Bm = b.double() if b_mn else b.double().transpose(-1, -2)
    return A @ Bm


def _run(dtype, m, n, k, a_mn, b_mn, tile_n=0, batch=1, seed=0):
    from d2r_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = lambda v: (v + 7) // 8 * 8
    a_shape = (batch, k, pad(m)) if a_mn else (batch, m, pad(k))
    b_shape = (batch, k, pad(n)) if b_mn else (batch, n, pad(k))
    a = torch.randn(a_shape, device="cuda", generator=g).to(dtype)
    b = torch.randn(b_shape, device="cuda", generator=g).to(dtype)
    c = torch.full((batch, m, n), float("nan"), device="cuda", dtype=torch.float32)
    K.gemm(a, b, c, m=m, n=n, k=k, lda=a.shape[-1], ldb=b.shape[-1], ldc=n, a_mn=a_mn, b_mn=b_mn, batch=batch,
           a_str=(a.stride(0), 0), b_str=(b.stride(0), 0), c_str=(m * n, 0), tile_n=tile_n)
    av = a[:, :, :m] if a_mn else a[:, :, :k]
    bv = b[:, :, :n] if b_mn else b[:, :, :k]
    ref = _ref(av, bv, a_mn, b_mn)
    err = (c.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale


MAJORS = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_tc_majors(a_mn, b_mn, tile_n):
    err, scale = _run(torch.bfloat16, 256, 256, 192, a_mn, b_mn, tile_n)
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("m,n,k", [(512, 512, 192), (300, 333, 72), (1000, 768, 768)])
def test_tc_cta_pair(a_mn, b_mn, m, n, k):
    """tile_n=512 forces the cta_group::2 kernel (256x256 tiles on a CTA pair), incl. ragged edges."""
    err, scale = _run(torch.bfloat16, m, n, k, a_mn, b_mn, 512, batch=2)
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("m,n,k", [(1024, 512, 192), (300, 333, 72), (1000, 768, 768), (4096, 768, 1536)])
def test_tc_cta_quad(a_mn, b_mn, m, n, k):
    """tile_n=1024 forces the 4-CTA cluster kernel: two cta_group::2 pairs on a 512x256 tile, the B tile TMA-
    multicast across the pairs; ragged shapes leave whole CTAs of a cluster without rows."""
    err, scale = _run(torch.bfloat16, m, n, k, a_mn, b_mn, 1024, batch=2)
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
@pytest.mark.parametrize("m,n,k", [(200, 50, 72), (128, 48, 128), (130, 197, 768), (77, 300, 40)])
def test_tc_ragged(a_mn, b_mn, m, n, k):
    err, scale = _run(torch.bfloat16, m, n, k, a_mn, b_mn, batch=3)
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", MAJORS)
def test_simt_majors(a_mn, b_mn):
    err, scale = _run(torch.float32, 150, 70, 90, a_mn, b_mn, batch=2)
    assert err <= 1e-5 * scale + 1e-5, (err, scale)


def test_tc_large_linear_shapes():
    for (m, n, k) in [(4096, 768, 768), (2560, 2304, 768)]:
        err, scale = _run(torch.bfloat16, m, n, k, False, False)
        assert err <= 2e-3 * scale + 1e-3, (m, n, k, err, scale)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_heads_strided_batch(dtype):
    """scores[b,h] = Q_h K_h^T with Q/K living inside a packed [B, L, 3*D] buffer (head stride 48)."""
    from d2r_b200 import kernels as K
    B, H, Ln, dk = 3, 16, 40, 48
    D = H * dk
    qkv = torch.randn(B, Ln, 3 * D, device="cuda").to(dtype)
    s = torch.full((B, H, Ln, Ln), float("nan"), device="cuda", dtype=torch.float32)
    K.gemm(qkv, qkv[:, :, D:], s, m=Ln, n=Ln, k=dk, lda=3 * D, ldb=3 * D, ldc=Ln, batch=B * H, batch_inner=H,
           a_str=(Ln * 3 * D, dk), b_str=(Ln * 3 * D, dk), c_str=(H * Ln * Ln, Ln * Ln), alpha=0.125)
    q = qkv[:, :, :D].view(B, Ln, H, dk).transpose(1, 2).double()
    k = qkv[:, :, D:2 * D].view(B, Ln, H, dk).transpose(1, 2).double()
    ref = 0.125 * q @ k.transpose(-1, -2)
    tol = 2e-3 if dtype == torch.bfloat16 else 1e-5
    assert (s.double() - ref).abs().max().item() <= tol * ref.abs().max().item() + 1e-4
    # o[b, :, h*dk:(h+1)*dk] = P[b,h] V[b,h]  (B operand MN-major, output written head-strided)
    p = torch.softmax(s, -1).to(dtype)
    o = torch.full((B, Ln, D), float("nan"), device="cuda", dtype=torch.float32)
    pl = p.contiguous()
    if Ln % 8:
        pytest.skip("ld alignment")
    K.gemm(pl, qkv[:, :, 2 * D:], o, m=Ln, n=dk, k=Ln, lda=Ln, ldb=3 * D, ldc=D, b_mn=True, batch=B * H,
           batch_inner=H, a_str=(H * Ln * Ln, Ln * Ln), b_str=(Ln * 3 * D, dk), c_str=(Ln * D, dk))
    v = qkv[:, :, 2 * D:].view(B, Ln, H, dk).transpose(1, 2).double()
    ref = (pl.double() @ v).transpose(1, 2).reshape(B, Ln, D)
    assert (o.double() - ref).abs().max().item() <= tol * ref.abs().max().item() + 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_epilogues(dtype):
    from d2r_b200 import kernels as K
    from d2r_b200 import _lib as L
    m, n, k = 300, 768, 768
    x = torch.randn(m, k, device="cuda").to(dtype)
    w = (torch.randn(n, k, device="cuda") / 28).to(dtype)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda").to(dtype)
    base = x.double() @ w.double().t() + bias.double()
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    y = K.linear(x, w, bias, act=L.ACT_RELU, residual=res)
    ref = torch.relu(base) + res.double()
    assert (y.double() - ref).abs().max().item() <= tol * ref.abs().max().item() + tol
    y = K.linear(x, w, bias, act=L.ACT_TANH, out_dtype=torch.float32)
    assert (y.double() - torch.tanh(base)).abs().max().item() <= tol
    d = torch.empty(m, n, device="cuda", dtype=dtype)
    sq = K.linear(x, w, bias, residual=res, epilogue=L.EPI_SQDIFF, c2=d)
    dref = res.double() - base
    assert (d.double() - dref).abs().max().item() <= tol * dref.abs().max().item() + tol
    assert (sq.double() - dref ** 2).abs().max().item() <= 2 * tol * (dref ** 2).abs().max().item() + tol


@pytest.mark.parametrize("Lq,Lc,dk,H", [(32, 32, 48, 16), (36, 32, 48, 16), (50, 37, 64, 4), (130, 200, 256, 1),
                                        (256, 256, 64, 2)])
def test_softmax_epilogues(Lq, Lc, dk, H):
    """P = softmax(alpha Q K^T) and dS = P * (alpha dO V^T - rowsum(alpha dO V^T * P)) finished in the GEMM's
    epilogue (scores never written to HBM); ragged Lc exercises the column mask and the padded ld."""
    from d2r_b200 import kernels as K
    from d2r_b200 import _lib as L
    B = 3
    Lcp = (Lc + 7) // 8 * 8
    bf = torch.bfloat16
    q = torch.randn(B, H, Lq, dk, device="cuda").to(bf)
    k = torch.randn(B, H, Lc, dk, device="cuda").to(bf)
    alpha = 1.0 / dk ** 0.5
    P = torch.full((B, H, Lq, Lcp), float("nan"), device="cuda", dtype=bf)
    K.gemm(q, k, P, m=Lq, n=Lc, k=dk, lda=dk, ldb=dk, ldc=Lcp, batch=B * H, batch_inner=H,
           a_str=(H * Lq * dk, Lq * dk), b_str=(H * Lc * dk, Lc * dk), c_str=(H * Lq * Lcp, Lq * Lcp), alpha=alpha,
           epilogue=L.EPI_SOFTMAX)
    ref = torch.softmax(alpha * q.double() @ k.double().transpose(-1, -2), -1)
    assert torch.isfinite(P[..., :Lc]).all()
    assert (P[..., :Lc].double() - ref).abs().max().item() <= 8e-3
    assert (P[..., :Lc].double().sum(-1) - 1).abs().max().item() <= 2e-2
    do = torch.randn(B, H, Lq, dk, device="cuda").to(bf)
    v = torch.randn(B, H, Lc, dk, device="cuda").to(bf)
    dS = torch.full((B, H, Lq, Lcp), float("nan"), device="cuda", dtype=bf)
    K.gemm(do, v, dS, m=Lq, n=Lc, k=dk, lda=dk, ldb=dk, ldc=Lcp, batch=B * H, batch_inner=H,
           a_str=(H * Lq * dk, Lq * dk), b_str=(H * Lc * dk, Lc * dk), c_str=(H * Lq * Lcp, Lq * Lcp), alpha=-0.5,
           epilogue=L.EPI_SOFTMAX_BWD, residual=P, ldr=Lcp, r_str=(H * Lq * Lcp, Lq * Lcp))
    Pd = P[..., :Lc].double()
    dP = -0.5 * do.double() @ v.double().transpose(-1, -2)
    ref = Pd * (dP - (dP * Pd).sum(-1, keepdim=True))
    assert (dS[..., :Lc].double() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    # the fp32 path has no fused softmax
    with pytest.raises(RuntimeError):
        K.gemm(q.float(), k.float(), P, m=Lq, n=Lc, k=dk, lda=dk, ldb=dk, ldc=Lcp, batch=B * H, batch_inner=H,
               a_str=(H * Lq * dk, Lq * dk), b_str=(H * Lc * dk, Lc * dk), c_str=(H * Lq * Lcp, Lq * Lcp),
               epilogue=L.EPI_SOFTMAX)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_wgrad_split_k(dtype):
    """dW[n,k] = dY^T X: both operands MN-major, contraction over the long row dimension, split-K."""
    from d2r_b200 import kernels as K
    rows, n, k = 4000, 768, 768
    dy = torch.randn(rows, n, device="cuda").to(dtype)
    x = torch.randn(rows, k, device="cuda").to(dtype)
    dw = torch.full((n, k), float("nan"), device="cuda", dtype=torch.float32)
    K.gemm(dy, x, dw, m=n, n=k, k=rows, lda=n, ldb=k, ldc=k, a_mn=True, b_mn=True, split_k=8)
    ref = dy.double().t() @ x.double()
    tol = 2e-3 if dtype == torch.bfloat16 else 2e-5
    assert (dw.double() - ref).abs().max().item() <= tol * ref.abs().max().item()
    K.gemm(dy, x, dw, m=n, n=k, k=rows, lda=n, ldb=k, ldc=k, a_mn=True, b_mn=True, split_k=4, accumulate=True)
    assert (dw.double() - 2 * ref).abs().max().item() <= 2 * tol * ref.abs().max().item()


def test_bad_arguments_raise():
    from d2r_b200 import kernels as K
    a = torch.randn(16, 20, device="cuda", dtype=torch.bfloat16)   # ld 20 not a multiple of 8
    b = torch.randn(16, 20, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(16, 16, device="cuda", dtype=torch.float32)
    with pytest.raises(RuntimeError):
        K.gemm(a, b, c, m=16, n=16, k=20, lda=20, ldb=20, ldc=16)
    with pytest.raises(RuntimeError):
        K.gemm(a.cpu(), b, c, m=16, n=16, k=16, lda=24, ldb=24, ldc=16)
