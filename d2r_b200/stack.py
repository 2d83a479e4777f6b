"""Hand-scheduled forward and backward of the routed interaction stack.

Everything here is plain Python that sequences calls into ``libd2r_b200.so`` (through
``kernels.py``); there is no torch arithmetic on the hot tensors and no autograd inside a stack
call -- the backward pass is written out explicitly so that gradient accumulations are fused
into GEMM epilogues (``residual=``) and no per-cell intermediate is materialised twice.

Reference map (paths relative to the upstream repo):
  stack_forward / stack_backward   models/InteractionModule.py:22-55, :75-108
  layer_forward / layer_backward   models/DynamicInteraction.py:37-69, :90-134 (+ Reversed_ twins)
  _routers_*                       models/Router.py:22-26
  _cma_*                           models/XModules.py:300-310, models/Refinement.py:105-115
  _glac_*                          models/Cells.py:145-175, models/XModules.py:380-384
  _imrc_*                          models/Cells.py:49-60, models/SelfAttention.py:27-70
  _cmrc_*                          models/Cells.py:82-87, models/Refinement.py:133-154
  _crcmc_*                         models/Cells.py:236-255
  _gesc_*                          models/Cells.py:197-218

Precision modes (``Env.cd``): torch.bfloat16 -> activations/weights bf16, tcgen05 GEMMs with fp32
accumulation, fp32 softmax/norm/router statistics; torch.float32 -> fp32 CUDA-core GEMMs.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import kernels as K
from . import lanes as LN

Tensor = torch.Tensor
CELLS6 = ("ric", "glac", "imrc", "cmrc", "crcmc", "gesc")   # emb_lst order, DynamicInteraction.py:41-48
CMA_TEMPERATURE = 100.0
FUSED_ATTN = True      # attention forward as ONE kernel (d2r_attn_fwd) where the shape allows; False: two d2r_gemm launches


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


# ----------------------------------------------------------------------------- weight staging
class Stager:
    """Per-module cache of GEMM-ready weights: bf16 copies (bf16 mode) and row-concatenated weights
    that share an A operand (QKV, the K/V projections of all cross-modal cells, FiLM scale|shift).
    Entries are refreshed when a parameter's version counter or storage changes; inside a CUDA-graph
    capture they are refreshed once per stack call (``fresh``: the keys already re-staged by this call's Env), so
    that the cast kernels are part of the graph but the backward does not repeat the forward's casts."""

    def __init__(self):
        self._cache: Dict[tuple, tuple] = {}

    def __getstate__(self):
        # staged copies are derived data: torch.save(model) / copy.deepcopy(model) start with an empty cache
        return {"_cache": {}}

    def invalidate(self) -> None:
        """Forget every staged copy (after weights were changed behind autograd's back, e.g. through ``p.data``)."""
        self._cache.clear()

    def get(self, names: Tuple[str, ...], params: Dict[str, Tensor], cd: torch.dtype, suffix: str,
            fresh: Optional[set] = None) -> Tensor:
        ps = [params[n + suffix] for n in names]
        if len(ps) == 1 and ps[0].dtype == cd:
            return ps[0].detach()
        key = (names, cd, suffix)
        sig = tuple((p.data_ptr(), p._version) for p in ps)
        hit = self._cache.get(key)
        capturing = ps[0].is_cuda and torch.cuda.is_current_stream_capturing()
        if hit is not None and hit[0] == sig and (not capturing or (fresh is not None and key in fresh)):
            return hit[1]
        rows = sum(p.shape[0] for p in ps)
        shape = (rows,) + tuple(ps[0].shape[1:])
        buf = hit[1] if (hit is not None and hit[1].shape == shape and hit[1].device == ps[0].device) else \
            torch.empty(shape, device=ps[0].device, dtype=cd)
        r = 0
        for p in ps:
            K.cast(p.detach(), cd, out=buf[r:r + p.shape[0]])
            r += p.shape[0]
        self._cache[key] = (sig, buf)
        if fresh is not None:
            fresh.add(key)
        return buf


class Env:
    """One stack call: parameters (fp32 masters, reference state_dict names relative to the module that
    owns the call), compute dtype, train/eval flag, collected parameter gradients."""

    def __init__(self, params: Dict[str, Tensor], cd: torch.dtype, training: bool, stager: Stager,
                 heads: int = 16):
        self.P = params
        self.cd = cd
        self.training = training
        self.bn_modes: Dict[str, bool] = {}   # BatchNorm1d name -> its own train/eval flag (autograd._bn_modes)
        self.st = stager
        self.layer_hook = None         # callable(layer prefix, {name: grad}) after each layer's backward
        self._fresh: set = set()       # staged-weight keys this call has already refreshed (see Stager)
        self._aux: Dict[int, tuple] = {}   # parent stream handle -> (parent, helper stream, tensors kept alive)
        self.heads = heads
        self.G: Dict[str, Tensor] = {}

    def bn_training(self, name: str) -> bool:
        """Does the batch-norm layer ``name`` use batch statistics?  (Its own flag, like nn.BatchNorm1d.forward.)"""
        return self.bn_modes.get(name, self.training)

    @contextlib.contextmanager
    def aux(self, *keep: Tensor):
        """Issue the enclosed launches on the current stream's helper stream (after everything queued so far);
        nothing on the current stream waits for them until aux_sync().  ``keep``: inputs allocated on the current
        stream that the helper reads -- held until the sync so that their memory is not reused under it."""
        t0 = keep[0] if keep else None
        if t0 is None or not t0.is_cuda or not (LN.ENABLED and LN.AUX_BIAS):
            yield
            return
        cur = torch.cuda.current_stream(t0.device)
        ent = self._aux.get(cur.cuda_stream)
        if ent is None:
            ent = self._aux[cur.cuda_stream] = (cur, LN.aux_stream(cur), [])
        ent[2].extend(keep)
        ent[1].wait_stream(cur)
        with torch.cuda.stream(ent[1]):
            yield

    def aux_sync(self) -> None:
        """Every stream that used a helper waits for it (call before a lane is joined / a pass returns)."""
        for cur, st, _keep in self._aux.values():
            cur.wait_stream(st)
        self._aux = {}

    def W(self, *names: str) -> Tensor:
        return self.st.get(tuple(names), self.P, self.cd, ".weight", self._fresh)

    def Wf32(self, name: str) -> Tensor:
        return self.P[name + ".weight"].detach()

    def b(self, *names: str) -> Tensor:
        if len(names) == 1:
            return self.P[names[0] + ".bias"].detach()
        return self.st.get(tuple(names), self.P, torch.float32, ".bias", self._fresh)

    def grad(self, name: str, g: Tensor) -> None:
        self.G[name] = g if name not in self.G else self.G[name] + g


# ----------------------------------------------------------------------------- linear helpers
def _split_k(m_out: int, n_out: int, k: int, tc: bool) -> int:
    """Split-K factor of a weight-gradient GEMM (k = B*L rows).  Tensor-core path: the long-k products run on
    CTA pairs (256x256 tiles, 74 pairs); pick the factor that minimises waves x (k-blocks per split + the cost
    of one fp32 reduce-add epilogue, ~6 k-blocks)."""
    if tc:
        pair = k >= 1536
        tiles = ((m_out + 255) // 256 if pair else (m_out + 127) // 128) * ((n_out + 255) // 256)
        workers = 74 if pair else 148
        kb = (k + 63) // 64
        best, best_cost = 1, None
        for sk in range(1, min(kb, 64) + 1):
            waves = (tiles * sk + workers - 1) // workers
            cost = waves * ((kb + sk - 1) // sk + 6)
            if best_cost is None or cost < best_cost:
                best, best_cost = sk, cost
        return best
    tiles = ((m_out + 63) // 64) * ((n_out + 63) // 64)
    kb = (k + 15) // 16
    want = max(1, (2 * 148) // tiles)      # <= 2 full waves of 148 SMs
    return max(1, min(want, kb, 64))


def lin_bwd(env: Env, dy: Tensor, x: Tensor, ldx: int, W: Tensor, names: Sequence[str], *, y: Optional[Tensor] = None,
            act: int = L.ACT_NONE, residual: Optional[Tensor] = None, need_dx: bool = True,
            dx_out: Optional[Tensor] = None, dx_ld: Optional[int] = None) -> Optional[Tensor]:
    """Backward of y = act(x W^T + b) for row-concatenated weights ``names``.

    dy [M,N] contiguous (dtype of W); x rows at stride ldx.  Returns dx [M,K] (= dz W + residual) and
    stores dW / db (fp32) in env.G.  ``dx_out``/``dx_ld`` let the result land in a strided destination
    (e.g. row 0 of every sample), in which case residual is read with the same stride."""
    N, Kd = W.shape
    M = dy.numel() // N
    if act == L.ACT_NONE:
        # the bias gradient (column sums of dy) feeds nothing on this chain: helper stream, beside the GEMMs
        dz = dy
        with env.aux(dy):
            _, db = K.bias_act_bwd(dy, None, act, False, True)
    else:
        dz, db = K.bias_act_bwd(dy, y, act, True, True)
    tc = W.dtype == torch.bfloat16
    dW = torch.empty(N, Kd, device=W.device, dtype=torch.float32)
    with (env.aux(dz, x) if LN.AUX_WGRAD else contextlib.nullcontext()):
        K.gemm(dz, x, dW, m=N, n=Kd, k=M, lda=N, ldb=ldx, ldc=Kd, a_mn=True, b_mn=True,
               split_k=_split_k(N, Kd, M, tc))
    r = 0
    for nme in names:
        rows = env.P[nme + ".weight"].shape[0]
        env.grad(nme + ".weight", dW[r:r + rows])
        env.grad(nme + ".bias", db[r:r + rows])
        r += rows
    if not need_dx:
        return None
    if dx_out is None:
        dx_out = torch.empty(M, Kd, device=W.device, dtype=W.dtype)
        dx_ld = Kd
    K.gemm(dz, W, dx_out, m=M, n=Kd, k=N, lda=N, ldb=Kd, ldc=dx_ld, b_mn=True, residual=residual,
           ldr=dx_ld if residual is not None else 0)
    return dx_out


def _row0_fwd(env: Env, xfull: Tensor, name: str) -> Tensor:
    """BertPooler (Cells.py:96-102): tanh(dense(x[:, 0])) -> fp32 [B,D]."""
    B, Ln, D = xfull.shape
    out = torch.empty(B, D, device=xfull.device, dtype=torch.float32)
    K.gemm(xfull, env.W(name), out, m=B, n=D, k=D, lda=Ln * D, ldb=D, ldc=D, bias=env.b(name), act=L.ACT_TANH)
    return out


def _row0_bwd_params(env: Env, d_vec: Tensor, y_vec: Tensor, xfull: Tensor, name: str):
    """Backward of a CLS pooler up to (not including) the input gradient: tanh', bias and weight gradients.
    d_vec: fp32 gradient w.r.t. the tanh output.  Returns (dz in the compute dtype, staged W) for _row0_bwd_apply."""
    B, Ln, D = xfull.shape
    dz, db = K.bias_act_bwd(d_vec, y_vec, L.ACT_TANH, True, True)
    dzc = dz if env.cd == torch.float32 else K.cast(dz, env.cd)
    W = env.W(name)
    dW = torch.empty(D, D, device=W.device, dtype=torch.float32)
    K.gemm(dzc, xfull, dW, m=D, n=D, k=B, lda=D, ldb=Ln * D, ldc=D, a_mn=True, b_mn=True)
    env.grad(name + ".weight", dW)
    env.grad(name + ".bias", db)
    return dzc, W


def _row0_bwd_apply(pending, dx_full: Tensor) -> None:
    """dx_full[:, 0, :] += dz W (read-modify-write of row 0 of every sample)."""
    dzc, W = pending
    B, Ln, D = dx_full.shape
    K.gemm(dzc, W, dx_full, m=B, n=D, k=D, lda=D, ldb=D, ldc=Ln * D, b_mn=True, residual=dx_full, ldr=Ln * D)


def _row0_bwd(env: Env, d_vec: Tensor, y_vec: Tensor, xfull: Tensor, name: str, dx_full: Tensor) -> None:
    """d_vec: fp32 gradient w.r.t. the tanh output; accumulates into row 0 of every sample of dx_full."""
    _row0_bwd_apply(_row0_bwd_params(env, d_vec, y_vec, xfull, name), dx_full)


def _small_fwd(env: Env, x32: Tensor, name: str, act: int = L.ACT_NONE):
    """y = act(x W^T + b) for the tiny fp32 [B,D] vectors of the global branches.  bf16 mode: operands are cast
    and the product runs on tensor cores (fp32 accumulate / output); fp32 mode: CUDA-core GEMM on the masters.
    Returns (y fp32, x in the GEMM operand dtype -- kept for the weight gradient)."""
    if env.cd == torch.float32:
        return K.linear(x32, env.Wf32(name), env.b(name), act=act), x32
    xc = K.cast(x32, env.cd)
    return K.linear(xc, env.W(name), env.b(name), act=act, out_dtype=torch.float32), xc


def _small_bwd(env: Env, dy32: Tensor, xc: Tensor, name: str, *, y: Optional[Tensor] = None,
               act: int = L.ACT_NONE) -> Tensor:
    """-> dx fp32 [B,K]."""
    dz, db = K.bias_act_bwd(dy32, y, act, True, True)
    env.grad(name + ".bias", db)
    tc = env.cd == torch.bfloat16
    dzc = K.cast(dz, env.cd) if tc else dz
    W = env.W(name) if tc else env.Wf32(name)
    N, Kd = W.shape
    M = dy32.numel() // N
    dW = torch.empty(N, Kd, device=W.device, dtype=torch.float32)
    K.gemm(dzc, xc, dW, m=N, n=Kd, k=M, lda=N, ldb=Kd, ldc=Kd, a_mn=True, b_mn=True,
           split_k=1 if tc else _split_k(N, Kd, M, False))
    env.grad(name + ".weight", dW)
    dx = torch.empty(M, Kd, device=W.device, dtype=torch.float32)
    K.gemm(dzc, W, dx, m=M, n=Kd, k=N, lda=N, ldb=Kd, ldc=Kd, b_mn=True)
    return dx


# ----------------------------------------------------------------------------- attention core
def _fused_softmax(cd: torch.dtype, Lc: int) -> bool:
    """The tensor-core GEMM can finish the softmax in its epilogue when a score row fits one N tile."""
    return cd == torch.bfloat16 and Lc <= 256


def attn_fwd(q: Tensor, q_ld: int, k: Tensor, k_ld: int, v: Tensor, v_ld: int, B: int, Lq: int, Lc: int, D: int,
             heads: int, alpha: float, cd: torch.dtype, *, residual: Optional[Tensor] = None,
             epilogue: int = L.EPI_STD, c2: Optional[Tensor] = None):
    """out[b,:,h] = softmax(alpha q_h k_h^T) v_h (+ residual | SQDIFF) -> (out [B,Lq,D] cd, P [B,H,Lq,Lcp] cd)."""
    dh = D // heads
    Lcp = pad8(Lc)
    dev = q.device
    if FUSED_ATTN and K.attn_fused_supported(cd, Lc, dh):
        # one kernel: scores in TMEM, probabilities handed to the second MMA through shared memory
        out, P, _ = K.attn_fused_fwd(q, q_ld, k, k_ld, v, v_ld, B=B, Lq=Lq, Lc=Lc, D=D, heads=heads, alpha=alpha,
                                     p_ld=Lcp, residual=residual, mode=1 if epilogue == L.EPI_SQDIFF else 0, out2=c2)
        return out, P
    if _fused_softmax(cd, Lc):
        # scores stay in TMEM: the GEMM's epilogue normalises the row and writes P directly
        P = torch.empty(B, heads, Lq, Lcp, device=dev, dtype=cd)
        K.gemm(q, k, P, m=Lq, n=Lc, k=dh, lda=q_ld, ldb=k_ld, ldc=Lcp, batch=B * heads, batch_inner=heads,
               a_str=(Lq * q_ld, dh), b_str=(Lc * k_ld, dh), c_str=(heads * Lq * Lcp, Lq * Lcp), alpha=alpha,
               epilogue=L.EPI_SOFTMAX)
    else:
        S = torch.empty(B, heads, Lq, Lcp, device=dev, dtype=torch.float32)
        K.gemm(q, k, S, m=Lq, n=Lc, k=dh, lda=q_ld, ldb=k_ld, ldc=Lcp, batch=B * heads, batch_inner=heads,
               a_str=(Lq * q_ld, dh), b_str=(Lc * k_ld, dh), c_str=(heads * Lq * Lcp, Lq * Lcp), alpha=alpha)
        P = K.softmax_fwd(S, Lc, 1.0, cd)
        del S
    out = torch.empty(B, Lq, D, device=dev, dtype=cd)
    K.gemm(P, v, out, m=Lq, n=dh, k=Lc, lda=Lcp, ldb=v_ld, ldc=D, b_mn=True, batch=B * heads, batch_inner=heads,
           a_str=(heads * Lq * Lcp, Lq * Lcp), b_str=(Lc * v_ld, dh), c_str=(Lq * D, dh),
           residual=residual, ldr=D if residual is not None else 0, r_str=(Lq * D, dh), epilogue=epilogue, c2=c2)
    return out, P


def attn_bwd(dO: Tensor, do_ld: int, sign: float, P: Tensor, q: Tensor, q_ld: int, k: Tensor, k_ld: int, v: Tensor,
             v_ld: int, dq: Tensor, dq_ld: int, dk: Tensor, dk_ld: int, dv: Tensor, dv_ld: int, B: int, Lq: int,
             Lc: int, D: int, heads: int, alpha: float, cd: torch.dtype) -> None:
    """Backward of attn_fwd's core (the residual path is the caller's).  ``sign`` multiplies dO (folds the
    minus of the squared-difference epilogue).  dq/dk/dv are written in place with their own lds."""
    dh = D // heads
    Lcp = pad8(Lc)
    HS = heads * Lq * Lcp
    if FUSED_ATTN and K.attn_fused_bwd_supported(cd, Lq, Lc, dh):
        K.attn_fused_bwd(dO, do_ld, sign, P, q, q_ld, k, k_ld, v, v_ld, dq, dq_ld, dk, dk_ld, dv, dv_ld,
                         B=B, Lq=Lq, Lc=Lc, D=D, heads=heads, alpha=alpha)
        return
    K.gemm(P, dO, dv, m=Lc, n=dh, k=Lq, lda=Lcp, ldb=do_ld, ldc=dv_ld, a_mn=True, b_mn=True, batch=B * heads,
           batch_inner=heads, a_str=(HS, Lq * Lcp), b_str=(Lq * do_ld, dh), c_str=(Lc * dv_ld, dh), alpha=sign)
    if _fused_softmax(cd, Lc):
        # dS = P * (dP - sum(dP * P)) straight out of the dP accumulator; dP never reaches HBM
        dS = torch.empty(B, heads, Lq, Lcp, device=dO.device, dtype=cd)
        K.gemm(dO, v, dS, m=Lq, n=Lc, k=dh, lda=do_ld, ldb=v_ld, ldc=Lcp, batch=B * heads, batch_inner=heads,
               a_str=(Lq * do_ld, dh), b_str=(Lc * v_ld, dh), c_str=(HS, Lq * Lcp), alpha=sign,
               epilogue=L.EPI_SOFTMAX_BWD, residual=P, ldr=Lcp, r_str=(HS, Lq * Lcp))
    else:
        dP = torch.empty(B, heads, Lq, Lcp, device=dO.device, dtype=torch.float32)
        K.gemm(dO, v, dP, m=Lq, n=Lc, k=dh, lda=do_ld, ldb=v_ld, ldc=Lcp, batch=B * heads, batch_inner=heads,
               a_str=(Lq * do_ld, dh), b_str=(Lc * v_ld, dh), c_str=(HS, Lq * Lcp), alpha=sign)
        dS = K.softmax_bwd(P, dP, Lc, 1.0, cd)
        del dP
    K.gemm(dS, k, dq, m=Lq, n=dh, k=Lc, lda=Lcp, ldb=k_ld, ldc=dq_ld, b_mn=True, batch=B * heads, batch_inner=heads,
           a_str=(HS, Lq * Lcp), b_str=(Lc * k_ld, dh), c_str=(Lq * dq_ld, dh), alpha=alpha)
    K.gemm(dS, q, dk, m=Lc, n=dh, k=Lq, lda=Lcp, ldb=q_ld, ldc=dk_ld, a_mn=True, b_mn=True, batch=B * heads,
           batch_inner=heads, a_str=(HS, Lq * Lcp), b_str=(Lq * q_ld, dh), c_str=(Lc * dk_ld, dh), alpha=alpha)


# ----------------------------------------------------------------------------- cross-modal attention
class _KV:
    """Key/value projections of the raw context for every cross-modal cell of a layer: one GEMM with
    N = 2*D*n_cells (the cells share the A operand z).  XModules.py:301-302 for each cell."""

    def __init__(self, env: Env, z: Tensor, cma_prefixes: Sequence[str]):
        self.cells = list(cma_prefixes)
        self.names = [f"{c}.{kv}" for c in self.cells for kv in ("key", "value")]
        B, Lc, D = z.shape
        self.z, self.D, self.ld = z, D, 2 * D * len(self.cells)
        self.W = env.W(*self.names)
        self.kv = K.linear(z, self.W, env.b(*self.names))              # [B, Lc, 2*D*n]
        self.dkv: Optional[Tensor] = None

    def k(self, i: int) -> Tensor:
        return self.kv[:, :, 2 * i * self.D:]

    def v(self, i: int) -> Tensor:
        return self.kv[:, :, (2 * i + 1) * self.D:]

    def grads(self) -> Tensor:
        if self.dkv is None:
            self.dkv = torch.empty_like(self.kv)
        return self.dkv

    def dk(self, i: int) -> Tensor:
        return self.grads()[:, :, 2 * i * self.D:]

    def dv(self, i: int) -> Tensor:
        return self.grads()[:, :, (2 * i + 1) * self.D:]

    def backward(self, env: Env, dz_acc: Optional[Tensor]) -> Tensor:
        """dz = dKV W (+ dz_acc); weight/bias gradients of all key/value projections."""
        B, Lc, D = self.z.shape
        dz = lin_bwd(env, self.grads(), self.z, D, self.W, self.names, residual=dz_acc)
        return dz.view(B, Lc, D)


def _cma_fwd(env: Env, name: str, x: Tensor, kv: _KV, i: int, **epi):
    """q = Wq x; softmax(100 q k^T / sqrt(D)) v.  Returns (out, saved)."""
    B, Lq, D = x.shape
    Lc = kv.z.shape[1]
    q = K.linear(x, env.W(name + ".query"), env.b(name + ".query"))
    alpha = CMA_TEMPERATURE / math.sqrt(D)
    out, P = attn_fwd(q, D, kv.k(i), kv.ld, kv.v(i), kv.ld, B, Lq, Lc, D, 1, alpha, env.cd, **epi)
    return out, (q, P)


def _cma_bwd(env: Env, name: str, x: Tensor, kv: _KV, i: int, saved, dC: Tensor, sign: float,
             residual: Optional[Tensor]) -> Tensor:
    """dC: gradient w.r.t. the attended context (times ``sign``).  Returns dx = dq Wq + residual."""
    q, P = saved
    B, Lq, D = x.shape
    Lc = kv.z.shape[1]
    alpha = CMA_TEMPERATURE / math.sqrt(D)
    dq = torch.empty_like(q)
    attn_bwd(dC, D, sign, P, q, D, kv.k(i), kv.ld, kv.v(i), kv.ld, dq, D, kv.dk(i), kv.ld, kv.dv(i), kv.ld,
             B, Lq, Lc, D, 1, alpha, env.cd)
    dx = lin_bwd(env, dq.view(B * Lq, D), x, D, env.W(name + ".query"), [name + ".query"], residual=residual)
    return dx.view(B, Lq, D)


# ----------------------------------------------------------------------------- cells
def _glac_local_fwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int):
    """Token-level half of GLAC (Cells.py:149-158): sim_local = fc_1(l2norm(fc_sim_tranloc((x - ctx)^2)))."""
    d1 = torch.empty_like(x)
    sl, cma = _cma_fwd(env, c + ".CrossModalAlignment", x, kv, ki, residual=x, epilogue=L.EPI_SQDIFF, c2=d1)
    t1 = K.linear(sl, env.W(c + ".fc_sim_tranloc"), env.b(c + ".fc_sim_tranloc"))
    n1, rn1 = K.l2norm_fwd(t1)
    del t1
    t2 = K.linear(n1, env.W(c + ".fc_1"), env.b(c + ".fc_1"))                  # sim_local [B,Lq,D]
    return t2, dict(cma=cma, d1=d1, sl=sl, n1=n1, rn1=rn1, t2=t2)


def _glac_global_fwd(env: Env, c: str, x: Tensor, z: Tensor):
    """[B,D] half of GLAC (Cells.py:159-163): sim_global = fc_2(l2norm(fc_sim_tranglo((t0 - i0)^2))), fp32."""
    t0 = _row0_fwd(env, x, c + ".text_cls_pool.dense")
    i0 = _row0_fwd(env, z, c + ".image_cls_pool.dense")
    dg = K.axpby(t0, i0, 1.0, -1.0)
    sq = K.mul(dg, dg)
    g1, sq_c = _small_fwd(env, sq, c + ".fc_sim_tranglo")
    ng, rng = K.l2norm_fwd(g1)
    sg, ng_c = _small_fwd(env, ng, c + ".fc_2")                                # sim_global [B,D]
    sgc = sg if env.cd == torch.float32 else K.cast(sg, env.cd)
    return sgc, dict(t0=t0, i0=i0, dg=dg, sq_c=sq_c, ng=ng, ng_c=ng_c, rng=rng, sgc=sgc)


def _glac_saf_fwd(env: Env, c: str, sgc: Tensor, t2: Tensor):
    """Attention filtration over [sim_global ; sim_local] (XModules.py:380-384) -> out fp32 [B,D]."""
    s = c + ".SAF_module"
    P = env.P
    nbt = P.get(s + ".bn.num_batches_tracked")
    return K.saf_fwd(sgc, t2, P[s + ".attn_sim_w.weight"].detach().view(-1), P[s + ".attn_sim_w.bias"].detach(),
                     P[s + ".bn.weight"].detach(), P[s + ".bn.bias"].detach(), P[s + ".bn.running_mean"],
                     P[s + ".bn.running_var"], nbt, env.bn_training(s + ".bn"))


def _glac_fwd(env: Env, c: str, x: Tensor, z: Tensor, kv: _KV, ki: int):
    t2, saved = _glac_local_fwd(env, c, x, kv, ki)
    sgc, gsv = _glac_global_fwd(env, c, x, z)
    out, saf = _glac_saf_fwd(env, c, sgc, t2)
    saved.update(gsv)
    saved["saf"] = saf
    return out, saved


def _glac_saf_bwd(env: Env, c: str, sv, d_out: Tensor):
    """d_out: fp32 [B,D] -> (d sim_global [B,D] compute dtype, d sim_local [B,Lq,D])."""
    s = c + ".SAF_module"
    P = env.P
    d_sgc, d_t2, d_w, d_b, d_bnw, d_bnb = K.saf_bwd(
        d_out, sv["sgc"], sv["t2"], P[s + ".attn_sim_w.weight"].detach().view(-1), P[s + ".attn_sim_w.bias"].detach(),
        P[s + ".bn.weight"].detach(), P[s + ".bn.bias"].detach(), P[s + ".bn.running_mean"],
        P[s + ".bn.running_var"], env.bn_training(s + ".bn"), sv["saf"])
    env.grad(s + ".attn_sim_w.weight", d_w.view(1, -1))
    env.grad(s + ".attn_sim_w.bias", d_b)
    env.grad(s + ".bn.weight", d_bnw)
    env.grad(s + ".bn.bias", d_bnb)
    return d_sgc, d_t2


def _glac_local_bwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int, sv, d_t2: Tensor, add: Optional[Tensor]) -> Tensor:
    B, Lq, D = x.shape
    dn1 = lin_bwd(env, d_t2.view(B * Lq, D), sv["n1"], D, env.W(c + ".fc_1"), [c + ".fc_1"])
    dt1 = K.l2norm_bwd(sv["n1"].view(B * Lq, D), dn1, sv["rn1"])
    dsl = lin_bwd(env, dt1, sv["sl"], D, env.W(c + ".fc_sim_tranloc"), [c + ".fc_sim_tranloc"])
    g, gx = K.sqdiff_bwd(dsl, sv["d1"].view(B * Lq, D), add.view(B * Lq, D) if add is not None else None,
                         want_gx=True)
    # d1 = x - ctx: d ctx = -g (sign folded into the attention backward), dx gets +g (+ add)
    return _cma_bwd(env, c + ".CrossModalAlignment", x, kv, ki, sv["cma"], g.view(B, Lq, D), -1.0, gx)


def _glac_global_bwd(env: Env, c: str, x: Tensor, z: Tensor, sv, d_sgc: Tensor):
    """-> pending row-0 updates (for _row0_bwd_apply): [into dx of this cell, into dz]."""
    dsg = d_sgc if env.cd == torch.float32 else K.cast(d_sgc, torch.float32)
    dng = _small_bwd(env, dsg, sv["ng_c"], c + ".fc_2")
    dg1 = K.l2norm_bwd(sv["ng"], dng, sv["rng"])
    dsq = _small_bwd(env, dg1, sv["sq_c"], c + ".fc_sim_tranglo")
    ddg = K.mul(dsq, sv["dg"], 2.0)
    px = _row0_bwd_params(env, ddg, sv["t0"], x, c + ".text_cls_pool.dense")
    nddg = K.axpby(ddg, None, -1.0, 0.0)
    pz = _row0_bwd_params(env, nddg, sv["i0"], z, c + ".image_cls_pool.dense")
    return [px, pz]


def _glac_bwd(env: Env, c: str, x: Tensor, z: Tensor, kv: _KV, ki: int, sv, d_out: Tensor, add: Optional[Tensor],
              dz_row0: Tensor) -> Tensor:
    """d_out: fp32 [B,D].  Returns dx [B,Lq,D]; the image_cls_pool gradient goes to row 0 of dz_row0."""
    d_sgc, d_t2 = _glac_saf_bwd(env, c, sv, d_out)
    dx = _glac_local_bwd(env, c, x, kv, ki, sv, d_t2, add)
    px, pz = _glac_global_bwd(env, c, x, z, sv, d_sgc)
    _row0_bwd_apply(px, dx)
    _row0_bwd_apply(pz, dz_row0)
    return dx


def _imrc_fwd(env: Env, c: str, x: Tensor):
    """c: prefix of the SelfAttention module (e.g. '<layer>.imrc.sa')."""
    B, Lq, D = x.shape
    qkv_names = [f"{c}.att_layer.linears.{i}" for i in range(3)]
    qkv = K.linear(x, env.W(*qkv_names), env.b(*qkv_names))                    # [B,Lq,3D]
    H = env.heads
    y, P = attn_fwd(qkv, 3 * D, qkv[:, :, D:], 3 * D, qkv[:, :, 2 * D:], 3 * D, B, Lq, Lq, D, H,
                    1.0 / math.sqrt(D // H), env.cd, residual=x)               # y = x + attn
    hf = K.linear(y, env.W(c + ".feed_forward_layer.fc1"), env.b(c + ".feed_forward_layer.fc1"), act=L.ACT_RELU)
    out = K.linear(hf, env.W(c + ".feed_forward_layer.fc2"), env.b(c + ".feed_forward_layer.fc2"), residual=y)
    return out, dict(qkv=qkv, P=P, y=y, hf=hf)


def _imrc_bwd(env: Env, c: str, x: Tensor, sv, d_out: Tensor, add: Optional[Tensor]) -> Tensor:
    B, Lq, D = x.shape
    H = env.heads
    M = B * Lq
    qkv, y, hf = sv["qkv"], sv["y"], sv["hf"]
    d_out2 = d_out.view(M, D)
    dhf = lin_bwd(env, d_out2, hf, D, env.W(c + ".feed_forward_layer.fc2"), [c + ".feed_forward_layer.fc2"])
    dy = lin_bwd(env, dhf, y, D, env.W(c + ".feed_forward_layer.fc1"), [c + ".feed_forward_layer.fc1"], y=hf,
                 act=L.ACT_RELU, residual=d_out2)
    dqkv = torch.empty_like(qkv)
    attn_bwd(dy, D, 1.0, sv["P"], qkv, 3 * D, qkv[:, :, D:], 3 * D, qkv[:, :, 2 * D:], 3 * D,
             dqkv, 3 * D, dqkv[:, :, D:], 3 * D, dqkv[:, :, 2 * D:], 3 * D, B, Lq, Lq, D, H,
             1.0 / math.sqrt(D // H), env.cd)
    res = dy if add is None else K.axpby(dy, add.view(M, D), 1.0, 1.0)
    qkv_names = [f"{c}.att_layer.linears.{i}" for i in range(3)]
    dx = lin_bwd(env, dqkv.view(M, 3 * D), x, D, env.W(*qkv_names), qkv_names, residual=res)
    return dx.view(B, Lq, D)


def _cmrc_fwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int):
    """c: prefix of the Refinement module (e.g. '<layer>.cmrc.refine')."""
    B, Lq, D = x.shape
    ctx, cma = _cma_fwd(env, c + ".CrossModalAlignment", x, kv, ki)
    st_names = [c + ".fc_scale", c + ".fc_shift"]
    st = torch.empty(B, Lq, 2 * D, device=x.device, dtype=env.cd)
    # [tanh(fc_scale(ctx)) | fc_shift(ctx)]: one GEMM, tanh on the first D columns only
    K.gemm(ctx, env.W(*st_names), st, m=B * Lq, n=2 * D, k=D, lda=D, ldb=D, ldc=2 * D, bias=env.b(*st_names),
           act=L.ACT_TANH, act_cols=D)
    m = K.film_fwd(x, st)
    h = K.linear(m, env.W(c + ".fc_1"), env.b(c + ".fc_1"), act=L.ACT_RELU)
    out = K.linear(h, env.W(c + ".fc_2"), env.b(c + ".fc_2"), residual=x)
    return out, dict(cma=cma, ctx=ctx, st=st, m=m, h=h)


def _cmrc_bwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int, sv, d_out: Tensor, add: Optional[Tensor]) -> Tensor:
    B, Lq, D = x.shape
    M = B * Lq
    d_out2 = d_out.view(M, D)
    dh = lin_bwd(env, d_out2, sv["h"], D, env.W(c + ".fc_2"), [c + ".fc_2"])
    dm = lin_bwd(env, dh, sv["m"], D, env.W(c + ".fc_1"), [c + ".fc_1"], y=sv["h"], act=L.ACT_RELU)
    acc = d_out2 if add is None else K.axpby(d_out2, add.view(M, D), 1.0, 1.0)
    dx_a, dst = K.film_bwd(dm, x, sv["st"], acc)                               # tanh' already applied to d scale
    st_names = [c + ".fc_scale", c + ".fc_shift"]
    dctx = lin_bwd(env, dst.view(M, 2 * D), sv["ctx"], D, env.W(*st_names), st_names)
    return _cma_bwd(env, c + ".CrossModalAlignment", x, kv, ki, sv["cma"], dctx.view(B, Lq, D), 1.0, dx_a)


def _crcmc_fwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int):
    B, Lq, D = x.shape
    ctx, cma = _cma_fwd(env, c + ".CrossModalAlignment", x, kv, ki)
    qs = K.linear(ctx, env.W(c + ".fc_mlp_1.0"), env.b(c + ".fc_mlp_1.0"), act=L.ACT_TANH)
    ks = K.linear(x, env.W(c + ".fc_mlp_2.0"), env.b(c + ".fc_mlp_2.0"), act=L.ACT_TANH)
    qp = K.linear(qs, env.W(c + ".fc_1"), env.b(c + ".fc_1"))
    kp = K.linear(ks, env.W(c + ".fc_2"), env.b(c + ".fc_2"))
    out, P2 = attn_fwd(qp, D, kp, D, ks, D, B, Lq, Lq, D, 1, 1.0, env.cd, residual=qs)   # un-scaled softmax
    return out, dict(cma=cma, ctx=ctx, qs=qs, ks=ks, qp=qp, kp=kp, P2=P2)


def _crcmc_bwd(env: Env, c: str, x: Tensor, kv: _KV, ki: int, sv, d_out: Tensor, add: Optional[Tensor]) -> Tensor:
    B, Lq, D = x.shape
    M = B * Lq
    qs, ks, qp, kp = sv["qs"], sv["ks"], sv["qp"], sv["kp"]
    dqp, dkp, dks_a = torch.empty_like(qp), torch.empty_like(kp), torch.empty_like(ks)
    attn_bwd(d_out, D, 1.0, sv["P2"], qp, D, kp, D, ks, D, dqp, D, dkp, D, dks_a, D, B, Lq, Lq, D, 1, 1.0, env.cd)
    dqs = lin_bwd(env, dqp.view(M, D), qs, D, env.W(c + ".fc_1"), [c + ".fc_1"], residual=d_out.view(M, D))
    dks = lin_bwd(env, dkp.view(M, D), ks, D, env.W(c + ".fc_2"), [c + ".fc_2"], residual=dks_a.view(M, D))
    dctx = lin_bwd(env, dqs, sv["ctx"], D, env.W(c + ".fc_mlp_1.0"), [c + ".fc_mlp_1.0"], y=qs, act=L.ACT_TANH)
    dx_a = lin_bwd(env, dks, x, D, env.W(c + ".fc_mlp_2.0"), [c + ".fc_mlp_2.0"], y=ks, act=L.ACT_TANH,
                   residual=add.view(M, D) if add is not None else None)
    return _cma_bwd(env, c + ".CrossModalAlignment", x, kv, ki, sv["cma"], dctx.view(B, Lq, D), 1.0, dx_a)


def _gesc_fwd(env: Env, c: str, x: Tensor, z: Tensor):
    t = _row0_fwd(env, x, c + ".text_cls_pool.dense")
    i = _row0_fwd(env, z, c + ".image_cls_pool.dense")
    u = K.axpby(t, i, 1.0, 1.0)
    h1, u_c = _small_fwd(env, u, c + ".fc_mlp.0", L.ACT_TANH)
    gl, h1_c = _small_fwd(env, h1, c + ".fc_mlp.2")
    g, out = K.gate_fuse_fwd(gl, t, i)
    return out, dict(t=t, i=i, u_c=u_c, h1=h1, h1_c=h1_c, g=g)


def _gesc_bwd_params(env: Env, c: str, x: Tensor, z: Tensor, sv, d_out: Tensor):
    """d_out fp32 [B,D] -> pending row-0 updates (for _row0_bwd_apply): [into dx of this cell, into dz]."""
    d_gl, d_t, d_i = K.gate_fuse_bwd(d_out, sv["g"], sv["t"], sv["i"])
    dh1 = _small_bwd(env, d_gl, sv["h1_c"], c + ".fc_mlp.2")
    du = _small_bwd(env, dh1, sv["u_c"], c + ".fc_mlp.0", y=sv["h1"], act=L.ACT_TANH)
    d_t = K.axpby(d_t, du, 1.0, 1.0)
    d_i = K.axpby(d_i, du, 1.0, 1.0)
    px = _row0_bwd_params(env, d_t, sv["t"], x, c + ".text_cls_pool.dense")
    pz = _row0_bwd_params(env, d_i, sv["i"], z, c + ".image_cls_pool.dense")
    return [px, pz]


def _gesc_bwd(env: Env, c: str, x: Tensor, z: Tensor, sv, d_out: Tensor, dx_row0: Tensor, dz_row0: Tensor) -> None:
    """d_out fp32 [B,D]; both pooler gradients are accumulated into row 0 of dx_row0 / dz_row0."""
    px, pz = _gesc_bwd_params(env, c, x, z, sv, d_out)
    _row0_bwd_apply(px, dx_row0)
    _row0_bwd_apply(pz, dz_row0)


# ----------------------------------------------------------------------------- routers
def _routers_fwd(env: Env, routers: Sequence[str], pooled: Tensor, n_out: int, final: bool):
    """routers: prefixes of the K Router modules; pooled fp32 [K,B,D] ([1,B,D] for a shared input).
    Hidden layers of all K routers run as ONE batched GEMM (stacked staged weights) in bf16 mode."""
    Kc = len(routers)
    shared = pooled.shape[0] == 1
    B, D = pooled.shape[1:]
    Hd = env.P[f"{routers[0]}.mlp.0.weight"].shape[0]
    hid = torch.empty(Kc, B, Hd, device=pooled.device, dtype=torch.float32)
    names = [f"{rn}.mlp.0" for rn in routers]
    pooled_c = pooled if env.cd == torch.float32 else K.cast(pooled, env.cd)
    K.gemm(pooled_c, env.W(*names), hid, m=B, n=Hd, k=D, lda=D, ldb=D, ldc=Hd, batch=Kc,
           a_str=(0 if shared else B * D, 0), b_str=(Hd * D, 0), c_str=(B * Hd, 0), bias=env.b(*names),
           bias_sz=Hd, act=L.ACT_RELU)
    w2 = [env.Wf32(f"{rn}.mlp.2") for rn in routers]
    b2 = [env.b(f"{rn}.mlp.2") for rn in routers]
    raw, norm, gate = K.router_head_fwd(hid, w2, b2, n_out, final)
    return norm, gate, dict(pooled=pooled, pooled_c=pooled_c, hid=hid, raw=raw, w2=w2)


def _routers_bwd(env: Env, routers: Sequence[str], sv, d_norm: Tensor, final: bool) -> Tensor:
    """-> d_pooled fp32 [K,B,D] (or the sum over cells, [1,B,D], for a shared input)."""
    pooled, pooled_c, hid = sv["pooled"], sv["pooled_c"], sv["hid"]
    Kc, B, Hd = hid.shape
    D = pooled.shape[2]
    shared = pooled.shape[0] == 1
    d_hid, _, d_w2, d_b2 = K.router_head_bwd(d_norm.contiguous(), sv["raw"], hid, sv["w2"], final)
    names = [f"{rn}.mlp.0" for rn in routers]
    for j, r in enumerate(routers):
        env.grad(r + ".mlp.2.weight", d_w2[j])
        env.grad(r + ".mlp.2.bias", d_b2[j])
        _, db = K.bias_act_bwd(d_hid[j], None, L.ACT_NONE, False, True)    # d_hid carries the ReLU mask
        env.grad(names[j] + ".bias", db)
    dzc = d_hid if env.cd == torch.float32 else K.cast(d_hid, env.cd)       # [K,B,Hd]
    dW = torch.empty(Kc, Hd, D, device=hid.device, dtype=torch.float32)
    K.gemm(dzc, pooled_c, dW, m=Hd, n=D, k=B, lda=Hd, ldb=D, ldc=D, a_mn=True, b_mn=True, batch=Kc,
           a_str=(B * Hd, 0), b_str=(0 if shared else B * D, 0), c_str=(Hd * D, 0))
    for j, nme in enumerate(names):
        env.grad(nme + ".weight", dW[j])
    d_pooled = (torch.zeros if shared else torch.empty)(1 if shared else Kc, B, D, device=hid.device,
                                                        dtype=torch.float32)
    K.gemm(dzc, env.W(*names), d_pooled, m=B, n=D, k=Hd, lda=Hd, ldb=D, ldc=D, b_mn=True, batch=Kc,
           a_str=(B * Hd, 0), b_str=(Hd * D, 0), c_str=(0 if shared else B * D, 0), accumulate=shared)
    return d_pooled


# ----------------------------------------------------------------------------- one routing layer
def layer_forward(env: Env, pre: str, xs: Sequence[Tensor], z: Tensor, pooled: Tensor, Kc: int, final: bool):
    """xs[j] feeds cell j ([B,Lq,D], compute dtype); z is the raw other modality; pooled = mean_L(xs[j])
    (fp32 [K,B,D], or [1,B,D] when all xs are the same tensor).
    Returns (outs, pooled_of_outs | None, norm probs [B,n_out,K], state)."""
    cells = CELLS6[:Kc]
    n_out = 1 if final else Kc
    routers = [f"{pre}.{cn}.router" for cn in cells]
    # The cells read the same layer input and are independent until the aggregation.  Lanes (CUDA streams):
    #   0 (caller's)  K/V projections of the context (shared by the cross-modal cells), GLAC's token-level half,
    #                 the attention filtration that joins GLAC's two halves
    #   1, 2, 3       IMRC, CMRC, CRCMC
    #   4             everything that lives on [B,768] vectors: GLAC's global half, the routers, GESC -- some
    #                 two dozen latency-bound launches that would otherwise lengthen lane 0, the longest chain
    lanes = LN.fork(z.device, LN.CELL_LANES if LN.FWD_LANES else 1, "cells")
    with lanes.lane(1):
        imrc_out, imrc_sv = _imrc_fwd(env, pre + ".imrc.sa", xs[2])
    with lanes.lane(4):
        sgc, glac_sv = _glac_global_fwd(env, pre + ".glac", xs[1], z)
    sgc_ready = lanes.mark(4)
    with lanes.lane(4):
        norm, gate, rsv = _routers_fwd(env, routers, pooled, n_out, final)
        if Kc > 4:
            gesc_out, gesc_sv = _gesc_fwd(env, pre + ".gesc", xs[5], z)
    kv = _KV(env, z, [pre + ".glac.CrossModalAlignment", pre + ".cmrc.refine.CrossModalAlignment"] +
             ([pre + ".crcmc.CrossModalAlignment"] if Kc > 4 else []))
    lanes.catch_up(2)
    with lanes.lane(2):
        cmrc_out, cmrc_sv = _cmrc_fwd(env, pre + ".cmrc.refine", xs[3], kv, 1)
    if Kc > 4:
        lanes.catch_up(3)
        with lanes.lane(3):
            crcmc_out, crcmc_sv = _crcmc_fwd(env, pre + ".crcmc", xs[4], kv, 2)
    t2, glac_local = _glac_local_fwd(env, pre + ".glac", xs[1], kv, 0)
    lanes.wait_mark(sgc_ready)
    glac_out, glac_sv["saf"] = _glac_saf_fwd(env, pre + ".glac", sgc, t2)
    glac_sv.update(glac_local)
    st = dict(xs=list(xs), z=z, kv=kv, router=rsv, norm=norm, gate=gate, final=final, Kc=Kc, imrc=imrc_sv,
              cmrc=cmrc_sv, glac=glac_sv)
    full: List[Optional[Tensor]] = [xs[0], None, imrc_out, cmrc_out]
    bvec: List[Optional[Tensor]] = [None, glac_out, None, None]
    if Kc > 4:
        st["crcmc"], st["gesc"] = crcmc_sv, gesc_sv
        full += [crcmc_out, None]
        bvec += [None, gesc_out]
    lanes.join()
    st["full"], st["bvec"] = full, bvec
    outs, pooled_next = K.aggregate_fwd(full, bvec, norm, gate, final, inputs=list(xs) if final else None)
    return outs, pooled_next, norm, st


def layer_backward(env: Env, pre: str, st, d_outs: Sequence[Tensor], d_pooled_next: Optional[Tensor],
                   d_norm_extra: Optional[Tensor], dz_acc: Optional[Tensor], shared_input: bool):
    """-> (d_xs: list of K gradients, or [one summed gradient] for a shared input; d_pooled fp32; dz)."""
    xs, z, kv, Kc, final = st["xs"], st["z"], st["kv"], st["Kc"], st["final"]
    cells = CELLS6[:Kc]
    B, Lq, D = xs[0].shape
    d_full, d_bvec, dP = K.aggregate_bwd(st["full"], st["bvec"], st["norm"], st["gate"], final, d_outs,
                                         d_pooled_next, inputs=list(xs) if final else None)
    d_norm = dP if d_norm_extra is None else K.axpby(dP, d_norm_extra.contiguous(), 1.0, 1.0)

    # the row-0 (CLS pooler) gradients of the context are collected in a zero-initialised dz_rows and
    # folded into dz by the key/value projection backward (residual)
    dz_rows = torch.zeros_like(z) if dz_acc is None else dz_acc
    kv.grads()                                       # allocated on the caller's stream, before the fork
    d_xs: List[Optional[Tensor]] = [None] * Kc
    # same lanes as the forward (a cell's saved activations were allocated on its lane).  The CLS-pooler gradients
    # of GLAC and GESC are read-modify-writes of row 0 of tensors other lanes produce (this cell's dx, dz_rows):
    # lane 4 computes everything up to them, they are applied after the join.
    lanes = LN.fork(z.device, LN.CELL_LANES if LN.BWD_LANES else 1, "cells")
    with lanes.lane(1):
        d_xs[2] = _imrc_bwd(env, pre + ".imrc.sa", xs[2], st["imrc"], d_full[2], None)
    with lanes.lane(2):
        d_xs[3] = _cmrc_bwd(env, pre + ".cmrc.refine", xs[3], kv, 1, st["cmrc"], d_full[3], None)
    if Kc > 4:
        with lanes.lane(3):
            d_xs[4] = _crcmc_bwd(env, pre + ".crcmc", xs[4], kv, 2, st["crcmc"], d_full[4], None)
    d_sgc, d_t2 = _glac_saf_bwd(env, pre + ".glac", st["glac"], d_bvec[1])
    lanes.catch_up(4)                                # lane 4 needs d_sgc
    gesc_rows = None
    with lanes.lane(4):
        glac_rows = _glac_global_bwd(env, pre + ".glac", xs[1], z, st["glac"], d_sgc)
        d_pooled = _routers_bwd(env, [f"{pre}.{cn}.router" for cn in cells], st["router"], d_norm, final)
        if Kc > 4:
            gesc_rows = _gesc_bwd_params(env, pre + ".gesc", xs[5], z, st["gesc"], d_bvec[5])
    # GLAC's token-level half; with one shared input tensor (layer 0) it folds the RIC gradient in as residual
    dx_glac = _glac_local_bwd(env, pre + ".glac", xs[1], kv, 0, st["glac"], d_t2, d_full[0] if shared_input else None)
    if not shared_input and Kc > 4:
        d_xs[5] = torch.empty_like(xs[5]) if final else torch.zeros_like(xs[5])
    env.aux_sync()
    lanes.join()
    _row0_bwd_apply(glac_rows[0], dx_glac)
    _row0_bwd_apply(glac_rows[1], dz_rows)
    if shared_input:
        # GESC adds its row-0 gradient on top of GLAC's dx, the other cells' gradients are summed in
        acc = dx_glac
        if Kc > 4:
            _row0_bwd_apply(gesc_rows[0], acc)
            _row0_bwd_apply(gesc_rows[1], dz_rows)
        for j in range(2, min(Kc, 5)):
            acc = K.axpby(acc, d_xs[j], 1.0, 1.0)
        d_xs = [acc]
    else:
        d_xs[0], d_xs[1] = d_full[0], dx_glac
        if final:
            # gated skip of the final layer (DynamicInteraction.py:108-111): touches only gated samples; cell 5
            # (GESC) has no full-size gradient yet and is overwritten, the others are accumulated into
            K.gate_skip_bwd(d_outs[0], st["norm"], st["gate"], d_xs, 0b011110 if Kc > 4 else 0b1110)
        if Kc > 4:
            _row0_bwd_apply(gesc_rows[0], d_xs[5])
            _row0_bwd_apply(gesc_rows[1], dz_rows)
    del d_full, d_bvec                               # (kept alive until the join: read by the side lanes)
    dz = kv.backward(env, dz_rows)
    env.aux_sync()
    return d_xs, d_pooled, dz


# ----------------------------------------------------------------------------- Block fusion
def block_forward(env: Env, x0: Tensor, x1: Tensor, chunks: int, rank: int):
    """XModules.py:521-555: linear0/1 -> chunk-wise merge linears (ONE batched GEMM per side instead of the
    reference's Python loop over 20 chunks) -> fused rank-sum / signed sqrt / per-chunk L2 norm -> linear_out.
    x0, x1: [B, D] in env.cd.  Returns (out [B, output_dim] env.cd, state)."""
    B = x0.shape[0]
    a = K.linear(x0, env.W("linear0"), env.b("linear0"))
    b = K.linear(x1, env.W("linear1"), env.b("linear1"))
    mm = a.shape[1]
    S = mm // chunks
    RS = rank * S
    ms = []
    for side, src in (("merge_linears0", a), ("merge_linears1", b)):
        names = [f"{side}.{c}" for c in range(chunks)]
        m = torch.empty(B, chunks * RS, device=a.device, dtype=env.cd)
        K.gemm(src, env.W(*names), m, m=B, n=RS, k=S, lda=mm, ldb=S, ldc=chunks * RS, batch=chunks,
               a_str=(S, 0), b_str=(RS * S, 0), c_str=(RS, 0), bias=env.b(*names), bias_sz=RS)
        ms.append(m)
    z, r, inv = K.block_merge_fwd(ms[0], ms[1], chunks, rank, S)
    out = K.linear(z, env.W("linear_out"), env.b("linear_out"))
    return out, dict(x0=x0, x1=x1, a=a, b=b, m0=ms[0], m1=ms[1], z=z, r=r, inv=inv, chunks=chunks, rank=rank)


def block_backward(env: Env, st, d_out: Tensor):
    """-> (dx0, dx1) in env.cd; parameter gradients in env.G."""
    chunks, rank = st["chunks"], st["rank"]
    a, b, z = st["a"], st["b"], st["z"]
    B, mm = a.shape
    S = mm // chunks
    RS = rank * S
    d_out = d_out.contiguous() if d_out.dtype == env.cd else K.cast(d_out.contiguous(), env.cd)
    dz = lin_bwd(env, d_out, z, mm, env.W("linear_out"), ["linear_out"])
    dms = K.block_merge_bwd(dz, st["m0"], st["m1"], st["r"], st["inv"], chunks, rank, S)
    dxs = []
    for side, src, dm, lin, x in (("merge_linears0", a, dms[0], "linear0", st["x0"]),
                                  ("merge_linears1", b, dms[1], "linear1", st["x1"])):
        names = [f"{side}.{c}" for c in range(chunks)]
        _, db = K.bias_act_bwd(dm, None, L.ACT_NONE, False, True)
        dW = torch.empty(chunks * RS, S, device=dm.device, dtype=torch.float32)
        K.gemm(dm, src, dW, m=RS, n=S, k=B, lda=chunks * RS, ldb=mm, ldc=S, a_mn=True, b_mn=True, batch=chunks,
               a_str=(RS, 0), b_str=(S, 0), c_str=(RS * S, 0))
        for c, nme in enumerate(names):
            env.grad(nme + ".weight", dW[c * RS:(c + 1) * RS])
            env.grad(nme + ".bias", db[c * RS:(c + 1) * RS])
        dsrc = torch.empty(B, mm, device=dm.device, dtype=env.cd)
        K.gemm(dm, env.W(*names), dsrc, m=B, n=S, k=RS, lda=chunks * RS, ldb=S, ldc=mm, b_mn=True, batch=chunks,
               a_str=(RS, 0), b_str=(RS * S, 0), c_str=(S, 0))
        dxs.append(lin_bwd(env, dsrc, x, x.shape[1], env.W(lin), [lin]))
    return dxs[0], dxs[1]


# ----------------------------------------------------------------------------- whole stack
def layer_prefixes(R: int) -> List[str]:
    return ["dynamic_itr_l0"] + [f"dynamic_itr_l1.{i}" for i in range(R - 2)] + ["dynamic_itr_l2"]


def stack_forward(env: Env, x: Tensor, z: Tensor, R: int, Kc: int):
    """x: the branch's own stream, z: the other modality (both [B,L,D] in env.cd, contiguous).
    Returns (out fp32 [B,Lq,D], sim_paths fp32 [B,B], per-layer probs, state)."""
    B = x.shape[0]
    pres = layer_prefixes(R)
    states = []
    probs = []
    pooled = K.pool_mean([x])
    xs: Sequence[Tensor] = [x] * Kc
    for li, pre in enumerate(pres):
        final = li == len(pres) - 1
        xs, pooled, norm, st = layer_forward(env, pre, xs, z, pooled, Kc, final)
        states.append(st)
        probs.append(norm)
    out = xs[0] if env.cd == torch.float32 else K.cast(xs[0], torch.float32)
    paths = torch.cat([p.reshape(B, -1) for p in probs], dim=1).contiguous()       # [B, K^2 (R-1) + K]
    Pn = paths.shape[1]
    sim = torch.empty(B, B, device=x.device, dtype=torch.float32)
    K.gemm(paths, paths, sim, m=B, n=B, k=Pn, lda=Pn, ldb=Pn, ldc=B)                # InteractionModule.py:53
    return out, sim, probs, dict(states=states, paths=paths, R=R, Kc=Kc, x=x, z=z)


def stack_backward(env: Env, state, d_out: Optional[Tensor], d_sim: Optional[Tensor],
                   d_probs: Optional[Sequence[Optional[Tensor]]] = None):
    """-> (dx, dz) in env.cd; parameter gradients are left in env.G."""
    states, paths, R, Kc = state["states"], state["paths"], state["R"], state["Kc"]
    x, z = state["x"], state["z"]
    K.zero_arena_reset()
    B, Lq, D = x.shape
    pres = layer_prefixes(R)
    Pn = paths.shape[1]
    d_paths = None
    if d_sim is not None:
        d_sim = d_sim.contiguous()
        d_paths = torch.empty_like(paths)
        K.gemm(d_sim, paths, d_paths, m=B, n=Pn, k=B, lda=B, ldb=Pn, ldc=Pn, b_mn=True)
        K.gemm(d_sim, paths, d_paths, m=B, n=Pn, k=B, lda=B, ldb=Pn, ldc=Pn, a_mn=True, b_mn=True,
               residual=d_paths, ldr=Pn)
    if d_out is None:
        d_out = torch.zeros(B, Lq, D, device=x.device, dtype=torch.float32)
    d_cur: Sequence[Tensor] = [d_out.contiguous() if env.cd == torch.float32 else K.cast(d_out, env.cd)]
    d_pooled_next = None
    dz = None
    off = Pn
    for li in reversed(range(len(pres))):
        st = states[li]
        n_out = 1 if st["final"] else Kc
        off -= n_out * Kc
        extra = None
        if d_paths is not None:
            extra = d_paths[:, off:off + n_out * Kc].reshape(B, n_out, Kc)
        if d_probs is not None and d_probs[li] is not None:
            extra = d_probs[li] if extra is None else extra + d_probs[li]
        d_xs, d_pooled, dz = layer_backward(env, pres[li], st, d_cur, d_pooled_next, extra, dz, shared_input=li == 0)
        d_cur, d_pooled_next = d_xs, d_pooled
        states[li] = None
        if env.layer_hook is not None:
            # every parameter gradient of this routing layer is final: hand them over (data-parallel runs start
            # the layer's all-reduce here, under the backward of the earlier layers)
            env.layer_hook(pres[li], env.G)
    # layer 0: d_cur = [sum of the cell gradients]; add the routers' mean-pool gradient (broadcast over L)
    dx = d_cur[0]
    K.pool_mean_bwd_into(d_pooled_next[0], dx)
    return dx, dz
