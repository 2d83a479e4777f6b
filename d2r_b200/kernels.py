"""Tensor-level wrappers of the C ABI (no autograd here; see ``functional.py``).

Every function takes torch CUDA tensors, passes raw device pointers / sizes / the current
stream to ``libd2r_b200.so`` and returns the output tensors it allocated from torch's caching
allocator (the library itself never allocates).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor


class _ZeroArena:
    """Small fp32 gradients that kernels accumulate into atomically (bias gradients, router head gradients...)
    are carved out of ONE pre-zeroed buffer per backward pass instead of ~250 separate torch.zeros launches.
    A fresh buffer is taken for every pass (never re-zeroed in place), so gradients of an earlier step that
    still alias an old arena stay valid."""

    CHUNK = 1 << 20

    def __init__(self):
        self.buf: Optional[Tensor] = None
        self.off = 0

    def reset(self) -> None:
        self.buf, self.off = None, 0

    def take(self, n: int, device: torch.device) -> Tensor:
        n_al = (n + 63) // 64 * 64
        if n_al > self.CHUNK // 4:
            return torch.zeros(n, device=device, dtype=torch.float32)
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            self.buf, self.off = torch.zeros(self.CHUNK, device=device, dtype=torch.float32), 0
        v = self.buf[self.off:self.off + n]
        self.off += n_al
        return v


# one arena per CUDA stream: a chunk is zeroed on the stream that is current when it is taken, so it must only
# be handed to kernels launched on that same stream (concurrent blocks run on different streams)
_zero_arenas: dict = {}


def _arena() -> _ZeroArena:
    key = L.stream()
    a = _zero_arenas.get(key)
    if a is None:
        a = _zero_arenas[key] = _ZeroArena()
    return a


def zeros_f32(shape, device: torch.device) -> Tensor:
    n = 1
    for d in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
        n *= int(d)
    return _arena().take(n, device).view(shape)


def zero_arena_reset() -> None:
    """Start of a backward pass: drop every stream's arena so that the chunks this pass uses are zeroed inside
    it (a CUDA graph of the pass must contain the memsets).  Views handed out earlier keep their chunk alive."""
    _zero_arenas.clear()


def gemm(a: Tensor, b: Tensor, c: Tensor, *, m: int, n: int, k: int, lda: int, ldb: int, ldc: int,
         a_mn: bool = False, b_mn: bool = False, batch: int = 1, batch_inner: int = 1,
         a_str: Tuple[int, int] = (0, 0), b_str: Tuple[int, int] = (0, 0), c_str: Tuple[int, int] = (0, 0),
         alpha: float = 1.0, bias: Optional[Tensor] = None, bias_sz: int = 0, act: int = L.ACT_NONE,
         residual: Optional[Tensor] = None, ldr: int = 0, r_str: Tuple[int, int] = (0, 0),
         epilogue: int = L.EPI_STD, c2: Optional[Tensor] = None, accumulate: bool = False, split_k: int = 1,
         tile_n: int = 0, act_cols: int = 0) -> Tensor:
    """C[z] = epilogue(alpha * A[z] (m x k) * B[z]^T (n x k)); strides in elements, (outer, inner)."""
    L.require_cuda(a, b, c, bias, residual, c2)
    if a.dtype != b.dtype:
        raise TypeError("gemm: A and B must have the same dtype")
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("gemm: bias must be fp32")
    g = L.GemmArgs()
    g.dtype, g.c_dtype = L.dt(a), L.dt(c)
    g.m, g.n, g.k = m, n, k
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.batch, g.batch_inner = batch, batch_inner
    g.act, g.epilogue = act, epilogue
    g.r_dtype = L.dt(residual) if residual is not None else 0
    g.accumulate, g.split_k = int(accumulate), split_k
    g.alpha, g.tile_n, g.act_cols = alpha, tile_n, act_cols
    g.a, g.b, g.c, g.c2 = a.data_ptr(), b.data_ptr(), c.data_ptr(), L.ptr(c2)
    g.bias, g.residual = L.ptr(bias), L.ptr(residual)
    g.lda, g.ldb, g.ldc, g.ldr = lda, ldb, ldc, ldr
    g.a_so, g.a_si = a_str
    g.b_so, g.b_si = b_str
    g.c_so, g.c_si = c_str
    g.r_so, g.r_si = r_str
    g.bias_sz = bias_sz
    L.check(L.lib.d2r_gemm(C.byref(g), L.stream()), "gemm")
    return c


def attn_fused_supported(dtype: torch.dtype, Lc: int, hd: int) -> bool:
    """Shapes the fused attention kernel serves (see d2r_attn_fwd in include/d2r_b200.h)."""
    return dtype == torch.bfloat16 and Lc <= 128 and hd % 16 == 0 and (hd <= 64 or hd % 64 == 0)


def attn_fused_fwd(q: Tensor, q_ld: int, k: Tensor, k_ld: int, v: Tensor, v_ld: int, *, B: int, Lq: int, Lc: int,
                   D: int, heads: int, alpha: float, p_ld: int, residual: Optional[Tensor] = None, mode: int = 0,
                   out2: Optional[Tensor] = None):
    """-> (out [B,Lq,D], P [B,heads,Lq,p_ld], out2 | None).  q/k/v may be column slices of wider buffers
    (row strides q_ld / k_ld / v_ld); residual is a contiguous [B,Lq,D] tensor; mode 1 = squared difference
    (out2 = residual - attention, out = out2^2), written into ``out2`` when given."""
    L.require_cuda(q, k, v, residual, out2)
    out = torch.empty(B, Lq, D, device=q.device, dtype=torch.bfloat16)
    if mode == 1 and out2 is None:
        out2 = torch.empty_like(out)
    P = torch.empty(B, heads, Lq, p_ld, device=q.device, dtype=torch.bfloat16)
    a = L.AttnArgs()
    a.B, a.heads, a.Lq, a.Lc, a.hd = B, heads, Lq, Lc, D // heads
    a.mode, a.alpha = mode, alpha
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.q_ld, a.k_ld, a.v_ld = q_ld, k_ld, v_ld
    a.p, a.p_ld = P.data_ptr(), p_ld
    a.out, a.out2, a.o_ld = out.data_ptr(), L.ptr(out2), D
    a.residual, a.r_ld = L.ptr(residual), D
    L.check(L.lib.d2r_attn_fwd(C.byref(a), L.stream()), "attn_fwd")
    return out, P, out2


def attn_fused_bwd_supported(dtype: torch.dtype, Lq: int, Lc: int, hd: int) -> bool:
    return attn_fused_supported(dtype, Lc, hd) and Lq <= 128


def attn_fused_bwd(dO: Tensor, do_ld: int, sign: float, P: Tensor, q: Tensor, q_ld: int, k: Tensor, k_ld: int,
                   v: Tensor, v_ld: int, dq: Tensor, dq_ld: int, dk: Tensor, dk_ld: int, dv: Tensor, dv_ld: int, *,
                   B: int, Lq: int, Lc: int, D: int, heads: int, alpha: float) -> None:
    """Backward of the attention core in one kernel; dq / dk / dv (possibly column slices, own row strides) are
    overwritten.  P [B, heads, Lq, p_ld] from the forward."""
    L.require_cuda(dO, P, q, k, v, dq, dk, dv)
    a = L.AttnBwdArgs()
    a.B, a.heads, a.Lq, a.Lc, a.hd = B, heads, Lq, Lc, D // heads
    a.alpha, a.sign = alpha, sign
    a.d_out, a.p, a.q, a.k, a.v = dO.data_ptr(), P.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.do_ld, a.p_ld, a.q_ld, a.k_ld, a.v_ld = do_ld, P.shape[-1], q_ld, k_ld, v_ld
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.dq_ld, a.dk_ld, a.dv_ld = dq_ld, dk_ld, dv_ld
    L.check(L.lib.d2r_attn_bwd(C.byref(a), L.stream()), "attn_bwd")


def linear(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, act: int = L.ACT_NONE,
           residual: Optional[Tensor] = None, out_dtype: Optional[torch.dtype] = None,
           out: Optional[Tensor] = None, epilogue: int = L.EPI_STD, c2: Optional[Tensor] = None,
           tile_n: int = 0) -> Tensor:
    """y = act(x w^T + bias) + residual for x [..., k] (last-dim contiguous rows, uniform row stride)."""
    k = x.shape[-1]
    n = w.shape[0]
    x2 = x.reshape(-1, k) if x.is_contiguous() else x
    assert x2.dim() == 2 and x2.stride(1) == 1, "linear: rows must be contiguous"
    m = x2.shape[0]
    if out is None:
        out = torch.empty(x.shape[:-1] + (n,), device=x.device, dtype=out_dtype or x.dtype)
    ldr = residual.shape[-1] if residual is not None else 0
    gemm(x2, w, out, m=m, n=n, k=k, lda=x2.stride(0), ldb=w.stride(0), ldc=n, bias=bias, act=act,
         residual=residual, ldr=ldr, epilogue=epilogue, c2=c2, tile_n=tile_n)
    return out


def softmax_fwd(x: Tensor, cols: int, scale: float, out_dtype: torch.dtype, ldy: Optional[int] = None) -> Tensor:
    """x [..., ldx] (only the first `cols` of each row are used) -> y [..., ldy]."""
    ldx = x.shape[-1]
    rows = x.numel() // ldx
    ldy = ldy or ldx
    y = torch.empty(x.shape[:-1] + (ldy,), device=x.device, dtype=out_dtype)
    L.check(L.lib.d2r_softmax_fwd(x.data_ptr(), L.dt(x), ldx, y.data_ptr(), L.dt(y), ldy, rows, cols, scale,
                                  L.stream()), "softmax_fwd")
    return y


def softmax_bwd(y: Tensor, dy: Tensor, cols: int, scale: float, out_dtype: torch.dtype) -> Tensor:
    ldy, lddy = y.shape[-1], dy.shape[-1]
    rows = y.numel() // ldy
    dx = torch.empty(y.shape, device=y.device, dtype=out_dtype)
    L.check(L.lib.d2r_softmax_bwd(y.data_ptr(), L.dt(y), ldy, dy.data_ptr(), L.dt(dy), lddy, dx.data_ptr(),
                                  L.dt(dx), ldy, rows, cols, scale, L.stream()), "softmax_bwd")
    return dx


def cast(x: Tensor, dtype: torch.dtype, out: Optional[Tensor] = None) -> Tensor:
    x = x.contiguous()
    y = out if out is not None else torch.empty(x.shape, device=x.device, dtype=dtype)
    L.check(L.lib.d2r_cast(x.data_ptr(), L.dt(x), y.data_ptr(), L.dt(y), x.numel(), L.stream()), "cast")
    return y


def bias_act_bwd(dy: Tensor, y: Optional[Tensor], act: int, want_dz: bool, want_db: bool):
    """dz = dy * act'(y) (in a new tensor, or dy itself when act is none); db = column sums (fp32)."""
    cols = dy.shape[-1]
    rows = dy.numel() // cols
    assert dy.is_contiguous()
    dz = torch.empty_like(dy) if (want_dz and act != L.ACT_NONE) else None
    db = zeros_f32(cols, dy.device) if want_db else None
    if dz is not None or db is not None:
        L.check(L.lib.d2r_bias_act_bwd(dy.data_ptr(), L.ptr(y), L.dt(dy), act, L.ptr(dz), L.ptr(db), rows, cols,
                                       cols, L.stream()), "bias_act_bwd")
    return (dz if dz is not None else dy), db


def l2norm_fwd(x: Tensor):
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty_like(x)
    rn = torch.empty(rows, device=x.device, dtype=torch.float32)
    L.check(L.lib.d2r_l2norm_fwd(x.data_ptr(), L.dt(x), y.data_ptr(), rn.data_ptr(), rows, cols, L.stream()),
            "l2norm_fwd")
    return y, rn


def l2norm_bwd(y: Tensor, dy: Tensor, rn: Tensor) -> Tensor:
    cols = y.shape[-1]
    rows = y.numel() // cols
    dx = torch.empty_like(y)
    L.check(L.lib.d2r_l2norm_bwd(y.data_ptr(), dy.data_ptr(), L.dt(y), rn.data_ptr(), dx.data_ptr(), rows, cols,
                                 L.stream()), "l2norm_bwd")
    return dx


def film_fwd(x: Tensor, st: Tensor) -> Tensor:
    cols = x.shape[-1]
    rows = x.numel() // cols
    m = torch.empty_like(x)
    L.check(L.lib.d2r_film_fwd(x.data_ptr(), st.data_ptr(), L.dt(x), m.data_ptr(), rows, cols, L.stream()),
            "film_fwd")
    return m


def film_bwd(dm: Tensor, x: Tensor, st: Tensor, add: Optional[Tensor] = None):
    """dx = dm * s (+ add), d_st = [dm * x * (1 - s^2) | dm]."""
    cols = x.shape[-1]
    rows = x.numel() // cols
    dx = torch.empty_like(x)
    dst = torch.empty_like(st)
    L.check(L.lib.d2r_film_bwd(dm.data_ptr(), x.data_ptr(), st.data_ptr(), L.ptr(add), L.dt(x), dx.data_ptr(),
                               dst.data_ptr(), rows, cols, L.stream()), "film_bwd")
    return dx, dst


def mul(x: Tensor, z: Tensor, alpha: float = 1.0) -> Tensor:
    y = torch.empty_like(x)
    L.check(L.lib.d2r_mul(x.data_ptr(), z.data_ptr(), L.dt(x), alpha, y.data_ptr(), x.numel(), L.stream()), "mul")
    return y


def axpby(x: Tensor, z: Optional[Tensor], a: float, b: float) -> Tensor:
    y = torch.empty_like(x)
    L.check(L.lib.d2r_axpby(x.data_ptr(), L.ptr(z), L.dt(x), a, b, y.data_ptr(), x.numel(), L.stream()), "axpby")
    return y


def sqdiff_bwd(dsq: Tensor, d: Tensor, add: Optional[Tensor] = None, want_gx: bool = False):
    """g = 2 d dsq; optionally gx = g + add (gradient of the minuend with an accumulated term)."""
    g = torch.empty_like(d)
    gx = torch.empty_like(d) if want_gx else None
    L.check(L.lib.d2r_sqdiff_bwd(dsq.data_ptr(), d.data_ptr(), L.ptr(add), L.dt(d), g.data_ptr(), L.ptr(gx),
                                 d.numel(), L.stream()), "sqdiff_bwd")
    return (g, gx) if want_gx else g


def pool_mean(xs: Sequence[Tensor]) -> Tensor:
    """[groups] tensors [B,L,D] -> fp32 [groups,B,D] (mean over L)."""
    B, Ln, D = xs[0].shape
    out = torch.empty(len(xs), B, D, device=xs[0].device, dtype=torch.float32)
    L.check(L.lib.d2r_pool_mean(L.ptr8(xs), len(xs), L.dt(xs[0]), B, Ln, D, out.data_ptr(), L.stream()),
            "pool_mean")
    return out


def pool_mean_bwd(d_pooled: Tensor, Ln: int, dtype: torch.dtype) -> Tensor:
    B, D = d_pooled.shape
    dx = torch.empty(B, Ln, D, device=d_pooled.device, dtype=dtype)
    L.check(L.lib.d2r_pool_mean_bwd(d_pooled.data_ptr(), B, Ln, D, dx.data_ptr(), L.dt(dx), 0, L.stream()),
            "pool_mean_bwd")
    return dx


def pool_mean_bwd_into(d_pooled: Tensor, dx: Tensor) -> Tensor:
    """dx[b, l, :] += d_pooled[b, :] / L (in place)."""
    B, Ln, D = dx.shape
    L.check(L.lib.d2r_pool_mean_bwd(d_pooled.data_ptr(), B, Ln, D, dx.data_ptr(), L.dt(dx), 1, L.stream()),
            "pool_mean_bwd")
    return dx


def gate_fuse_fwd(gl: Tensor, t: Tensor, i: Tensor):
    B, D = gl.shape
    g = torch.empty_like(gl)
    out = torch.empty_like(gl)
    L.check(L.lib.d2r_gate_fuse_fwd(gl.data_ptr(), t.data_ptr(), i.data_ptr(), g.data_ptr(), out.data_ptr(), B, D,
                                    L.stream()), "gate_fuse_fwd")
    return g, out


def gate_fuse_bwd(d_out: Tensor, g: Tensor, t: Tensor, i: Tensor):
    B, D = g.shape
    d_gl, d_t, d_i = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
    L.check(L.lib.d2r_gate_fuse_bwd(d_out.data_ptr(), g.data_ptr(), t.data_ptr(), i.data_ptr(), d_gl.data_ptr(),
                                    d_t.data_ptr(), d_i.data_ptr(), B, D, L.stream()), "gate_fuse_bwd")
    return d_gl, d_t, d_i


def js_div_fwd(p: Tensor, q: Tensor, get_softmax: bool = True) -> Tensor:
    """JS divergence of two [rows, cols] fp32 logit matrices (XModules.py:32-41) -> fp32 scalar tensor."""
    L.require_cuda(p, q)
    rows, cols = p.shape
    loss = torch.zeros((), device=p.device, dtype=torch.float32)
    L.check(L.lib.d2r_js_div_fwd(p.data_ptr(), q.data_ptr(), rows, cols, int(get_softmax), loss.data_ptr(),
                                 L.stream()), "js_div_fwd")
    return loss


def js_div_bwd(p: Tensor, q: Tensor, d_loss: Tensor, get_softmax: bool = True):
    rows, cols = p.shape
    dp, dq = torch.empty_like(p), torch.empty_like(q)
    L.check(L.lib.d2r_js_div_bwd(p.data_ptr(), q.data_ptr(), rows, cols, int(get_softmax), d_loss.data_ptr(),
                                 dp.data_ptr(), dq.data_ptr(), L.stream()), "js_div_bwd")
    return dp, dq


def block_merge_fwd(m0: Tensor, m1: Tensor, chunks: int, rank: int, size: int):
    """Block fusion core (XModules.py:538-543): m0, m1 [B, chunks*rank*size] -> z [B, chunks*size] (same dtype),
    r [B, chunks*size] fp32, inv_norm [B, chunks] fp32."""
    L.require_cuda(m0, m1)
    B = m0.shape[0]
    z = torch.empty(B, chunks * size, device=m0.device, dtype=m0.dtype)
    r = torch.empty(B, chunks * size, device=m0.device, dtype=torch.float32)
    inv = torch.empty(B, chunks, device=m0.device, dtype=torch.float32)
    L.check(L.lib.d2r_block_merge_fwd(m0.data_ptr(), m1.data_ptr(), L.dt(m0), B, chunks, rank, size, z.data_ptr(),
                                      r.data_ptr(), inv.data_ptr(), L.stream()), "block_merge_fwd")
    return z, r, inv


def block_merge_bwd(dz: Tensor, m0: Tensor, m1: Tensor, r: Tensor, inv: Tensor, chunks: int, rank: int, size: int):
    B = m0.shape[0]
    dm0, dm1 = torch.empty_like(m0), torch.empty_like(m1)
    L.check(L.lib.d2r_block_merge_bwd(dz.data_ptr(), m0.data_ptr(), m1.data_ptr(), r.data_ptr(), inv.data_ptr(),
                                      L.dt(m0), B, chunks, rank, size, dm0.data_ptr(), dm1.data_ptr(), L.stream()),
            "block_merge_bwd")
    return dm0, dm1


def router_head_fwd(hid: Tensor, w2: Sequence[Tensor], b2: Sequence[Tensor], n_out: int, final_layer: bool):
    """hid [K,B,H] fp32 (post-ReLU) -> raw, norm [B,n_out,K], gate [B,n_out] ([B,K] in the final layer)."""
    K_, B, H = hid.shape
    raw = torch.empty(B, n_out, K_, device=hid.device, dtype=torch.float32)
    norm = torch.empty_like(raw)
    gate = torch.empty(B, K_ if final_layer else n_out, device=hid.device, dtype=torch.float32)
    L.check(L.lib.d2r_router_head_fwd(hid.data_ptr(), L.ptr8(w2), L.ptr8(b2), K_, n_out, B, H, int(final_layer),
                                      raw.data_ptr(), norm.data_ptr(), gate.data_ptr(), L.stream()),
            "router_head_fwd")
    return raw, norm, gate


def router_head_bwd(d_norm: Tensor, raw: Tensor, hid: Tensor, w2: Sequence[Tensor], final_layer: bool):
    """-> d_hid [K,B,H] (ReLU mask applied), d_logit [B,n_out,K], [dW2_j], [db2_j]."""
    K_, B, H = hid.shape
    n_out = raw.shape[1]
    d_hid = torch.empty_like(hid)
    d_logit = torch.empty_like(raw)
    d_w2 = [zeros_f32(w.shape, hid.device) for w in w2]
    d_b2 = [zeros_f32(n_out, hid.device) for _ in w2]
    L.check(L.lib.d2r_router_head_bwd(d_norm.data_ptr(), raw.data_ptr(), hid.data_ptr(), L.ptr8(w2), K_, n_out, B, H,
                                      int(final_layer), d_hid.data_ptr(), d_logit.data_ptr(), L.ptr8(d_w2),
                                      L.ptr8(d_b2), L.stream()), "router_head_bwd")
    return d_hid, d_logit, d_w2, d_b2


def _agg_args(full, bvec, inputs, outs, P, gate, pooled, K_, n_out, final_layer, B, Ln, D, dtype) -> L.AggArgs:
    a = L.AggArgs()
    a.K, a.n_out, a.final_layer, a.dtype = K_, n_out, int(final_layer), dtype
    a.B, a.L, a.D = B, Ln, D
    a.full, a.bvec, a.inputs, a.out = L.ptr8(full), L.ptr8(bvec), L.ptr8(inputs), L.ptr8(outs)
    a.P, a.gate, a.pooled = P.data_ptr(), gate.data_ptr(), L.ptr(pooled)
    return a


def aggregate_fwd(full: Sequence[Optional[Tensor]], bvec: Sequence[Optional[Tensor]], P: Tensor, gate: Tensor,
                  final_layer: bool, inputs: Optional[Sequence[Optional[Tensor]]] = None, want_pooled: bool = True):
    """full[j]: [B,L,D] or None; bvec[j]: fp32 [B,D] or None -> (outs, pooled [n_out,B,D] | None)."""
    K_ = len(full)
    x0 = full[0]
    B, Ln, D = x0.shape
    n_out = 1 if final_layer else K_
    outs = [torch.empty_like(x0) for _ in range(n_out)]
    pooled = (torch.empty(n_out, B, D, device=x0.device, dtype=torch.float32)
              if (want_pooled and not final_layer) else None)
    a = _agg_args(full, bvec, inputs or [None] * K_, outs, P, gate, pooled, K_, n_out, final_layer, B, Ln, D,
                  L.dt(x0))
    L.check(L.lib.d2r_aggregate_fwd(C.byref(a), L.stream()), "aggregate_fwd")
    return outs, pooled


def aggregate_bwd(full, bvec, P: Tensor, gate: Tensor, final_layer: bool, d_outs: Sequence[Tensor],
                  d_pooled: Optional[Tensor], inputs=None):
    """-> d_full (list, None for broadcast cells), d_bvec (fp32 [B,D] list), dP."""
    K_ = len(full)
    x0 = full[0]
    B, Ln, D = x0.shape
    n_out = 1 if final_layer else K_
    d_full = [torch.empty_like(x0) if f is not None else None for f in full]
    d_bvec = [torch.empty(B, D, device=x0.device, dtype=torch.float32) if v is not None else None for v in bvec]
    dP = torch.empty(B, n_out, K_, device=x0.device, dtype=torch.float32)
    b = L.AggBwdArgs()
    b.fwd = _agg_args(full, bvec, inputs or [None] * K_, [None] * 8, P, gate, None, K_, n_out, final_layer, B, Ln, D,
                      L.dt(x0))
    b.d_out = L.ptr8(d_outs)
    b.d_pooled = L.ptr(d_pooled)
    b.d_full, b.d_bvec = L.ptr8(d_full), L.ptr8(d_bvec)
    b.dP = dP.data_ptr()
    L.check(L.lib.d2r_aggregate_bwd(C.byref(b), L.stream()), "aggregate_bwd")
    return d_full, d_bvec, dP


def gate_skip_bwd(d_out: Tensor, P: Tensor, gate: Tensor, dxs: Sequence[Optional[Tensor]], accumulate_mask: int) -> None:
    """Final layer: dxs[j][b] (+)= gate[b,j] / S[b] * d_out[b] for j >= 1 (in place; None entries skipped)."""
    B, Ln, D = d_out.shape
    L.check(L.lib.d2r_gate_skip_bwd(d_out.data_ptr(), P.data_ptr(), gate.data_ptr(), L.ptr8(dxs), len(dxs), B, Ln, D,
                                    L.dt(d_out), accumulate_mask, L.stream()), "gate_skip_bwd")


def _saf_args(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, training, saved) -> L.SafArgs:
    B, Ln, D = sl.shape
    a = L.SafArgs()
    a.dtype, a.training = L.dt(sl), int(training)
    a.B, a.L, a.D = B, Ln, D
    a.sg, a.sl, a.w, a.bias = sg.data_ptr(), sl.data_ptr(), w.data_ptr(), bias.data_ptr()
    a.bn_w, a.bn_b = bn_w.data_ptr(), bn_b.data_ptr()
    a.running_mean, a.running_var, a.num_batches_tracked = rm.data_ptr(), rv.data_ptr(), L.ptr(nbt)
    logits, attn, stats, rnorm, out = saved
    a.logits, a.attn, a.stats, a.rnorm, a.out = (logits.data_ptr(), attn.data_ptr(), stats.data_ptr(),
                                                 rnorm.data_ptr(), out.data_ptr())
    return a


def saf_fwd(sg: Tensor, sl: Tensor, w: Tensor, bias: Tensor, bn_w: Tensor, bn_b: Tensor, rm: Tensor, rv: Tensor,
            nbt: Optional[Tensor], training: bool):
    """Attention filtration; sg [B,D], sl [B,L,D] (same dtype); returns (out fp32 [B,D], saved tuple)."""
    B, Ln, D = sl.shape
    dev = sl.device
    saved = (torch.empty(B, Ln + 1, device=dev), torch.empty(B, Ln + 1, device=dev), torch.empty(2, device=dev),
             torch.empty(B, device=dev), torch.empty(B, D, device=dev))
    a = _saf_args(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, training, saved)
    L.check(L.lib.d2r_saf_fwd(C.byref(a), L.stream()), "saf_fwd")
    return saved[4], saved


def saf_bwd(d_out: Tensor, sg, sl, w, bias, bn_w, bn_b, rm, rv, training: bool, saved):
    B, Ln, D = sl.shape
    dev = sl.device
    d_sg, d_sl = torch.empty_like(sg), torch.empty_like(sl)
    d_w = zeros_f32(D, dev)
    d_bias, d_bn_w, d_bn_b = zeros_f32(1, dev), zeros_f32(1, dev), zeros_f32(1, dev)
    scratch = torch.empty(B * D + B * (Ln + 1) + 8, device=dev)
    b = L.SafBwdArgs()
    b.fwd = _saf_args(sg, sl, w, bias, bn_w, bn_b, rm, rv, None, training, saved)
    b.d_out = d_out.data_ptr()
    b.d_sg, b.d_sl = d_sg.data_ptr(), d_sl.data_ptr()
    b.d_w, b.d_bias, b.d_bn_w, b.d_bn_b = d_w.data_ptr(), d_bias.data_ptr(), d_bn_w.data_ptr(), d_bn_b.data_ptr()
    b.scratch = scratch.data_ptr()
    L.check(L.lib.d2r_saf_bwd(C.byref(b), L.stream()), "saf_bwd")
    return d_sg, d_sl, d_w, d_bias, d_bn_w, d_bn_b
