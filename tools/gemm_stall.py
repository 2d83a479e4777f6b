"""Where do the role warps of the tcgen05 GEMM wait?  (GPU box)

Switches on the library's debug probe (d2r_gemm_set_profile): every CTA of a launch writes the SM-clock cycles
its TMA producer thread spent waiting for a free shared-memory stage, its MMA-issuing thread waiting for operand
bytes (full barrier) and for a free TMEM accumulator, and one epilogue warp waiting for a finished accumulator.
Prints one block per case: CUDA-event time without the probe, then the per-CTA averages with it.

usage: python tools/gemm_stall.py [--case a,b,...] [--out gpurun_out/gemm_stall.txt]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200 import _lib as L  # noqa: E402

SLOTS = ["start", "end", "prod_wait_empty", "prod_kblocks", "prod_loop", "mma_wait_full", "mma_wait_tmem", "mma_loop",
         "tiles", "epi_wait_full", "epi_loop", "prologue", "first_full", "smid", "epi_store_wait", "_"]
bf = torch.bfloat16


def t(*shape, dtype=bf):
    return torch.randn(*shape, device="cuda").to(dtype)


def case_linear(m, n, k, a_mn=False, b_mn=False, out=bf, bias=True, split_k=1, tile_n=0, residual=False):
    a = t(k, m) if a_mn else t(m, k)
    b = t(k, n) if b_mn else t(n, k)
    c = torch.empty(m, n, device="cuda", dtype=out)
    kw = dict(m=m, n=n, k=k, lda=a.shape[1], ldb=b.shape[1], ldc=n, a_mn=a_mn, b_mn=b_mn,
              bias=torch.randn(n, device="cuda") if bias else None, split_k=split_k, tile_n=tile_n)
    if residual:
        kw.update(residual=t(m, n, dtype=out), ldr=n)
    return (a, b, c), kw, 2.0 * m * n * k


def case_self_scores(B=256, H=16, Lq=128, D=768):
    dh = D // H
    qkv = t(B, Lq, 3 * D)
    P = torch.empty(B, H, Lq, Lq, device="cuda", dtype=bf)
    kw = dict(m=Lq, n=Lq, k=dh, lda=3 * D, ldb=3 * D, ldc=Lq, batch=B * H, batch_inner=H,
              a_str=(Lq * 3 * D, dh), b_str=(Lq * 3 * D, dh), c_str=(H * Lq * Lq, Lq * Lq), alpha=0.144,
              epilogue=L.EPI_SOFTMAX)
    return (qkv, qkv[:, :, D:], P), kw, 2.0 * Lq * Lq * dh * B * H


def case_self_pv(B=256, H=16, Lq=128, D=768):
    dh = D // H
    qkv = t(B, Lq, 3 * D)
    P = t(B, H, Lq, Lq)
    out = torch.empty(B, Lq, D, device="cuda", dtype=bf)
    kw = dict(m=Lq, n=dh, k=Lq, lda=Lq, ldb=3 * D, ldc=D, b_mn=True, batch=B * H, batch_inner=H,
              a_str=(H * Lq * Lq, Lq * Lq), b_str=(Lq * 3 * D, dh), c_str=(Lq * D, dh))
    return (P, qkv[:, :, 2 * D:], out), kw, 2.0 * Lq * dh * Lq * B * H


def case_cma_scores(B=256, Lq=128, Lc=50, D=768):
    Lcp = (Lc + 7) // 8 * 8
    q, kv = t(B, Lq, D), t(B, Lc, 6 * D)
    P = torch.empty(B, 1, Lq, Lcp, device="cuda", dtype=bf)
    kw = dict(m=Lq, n=Lc, k=D, lda=D, ldb=6 * D, ldc=Lcp, batch=B, batch_inner=1, a_str=(Lq * D, D),
              b_str=(Lc * 6 * D, D), c_str=(Lq * Lcp, Lq * Lcp), alpha=3.6, epilogue=L.EPI_SOFTMAX)
    return (q, kv, P), kw, 2.0 * Lq * Lc * D * B


def case_cma_pv(B=256, Lq=128, Lc=50, D=768):
    Lcp = (Lc + 7) // 8 * 8
    P, kv = t(B, 1, Lq, Lcp), t(B, Lc, 6 * D)
    out = torch.empty(B, Lq, D, device="cuda", dtype=bf)
    kw = dict(m=Lq, n=D, k=Lc, lda=Lcp, ldb=6 * D, ldc=D, b_mn=True, batch=B, batch_inner=1,
              a_str=(Lq * Lcp, Lq * Lcp), b_str=(Lc * 6 * D, D), c_str=(Lq * D, D))
    return (P, kv[:, :, D:], out), kw, 2.0 * Lq * D * Lc * B


def case_cma_dv(B=256, Lq=128, Lc=50, D=768):
    Lcp = (Lc + 7) // 8 * 8
    P, dO = t(B, 1, Lq, Lcp), t(B, Lq, D)
    dkv = torch.empty(B, Lc, 6 * D, device="cuda", dtype=bf)
    kw = dict(m=Lc, n=D, k=Lq, lda=Lcp, ldb=D, ldc=6 * D, a_mn=True, b_mn=True, batch=B, batch_inner=1,
              a_str=(Lq * Lcp, Lq * Lcp), b_str=(Lq * D, D), c_str=(Lc * 6 * D, D))
    return (P, dO, dkv), kw, 2.0 * Lc * D * Lq * B


CASES = {
    "fwd_nt_32768": lambda: case_linear(32768, 768, 768),
    "dgrad_nn_32768": lambda: case_linear(32768, 768, 768, b_mn=True, bias=False),
    "wgrad_tt_32768_sk16": lambda: case_linear(768, 768, 32768, a_mn=True, b_mn=True, out=torch.float32, bias=False,
                                               split_k=16),
    "fwd_nt_12800": lambda: case_linear(12800, 768, 768),
    "dgrad_nn_12800": lambda: case_linear(12800, 768, 768, b_mn=True, bias=False),
    "wgrad_tt_12800_sk16": lambda: case_linear(768, 768, 12800, a_mn=True, b_mn=True, out=torch.float32, bias=False,
                                               split_k=16),
    "fwd_longk_4608": lambda: case_linear(32768, 768, 4608, bias=False),
    "fwd_wide_4608": lambda: case_linear(32768, 4608, 768),
    "fwd_nt_32768_res": lambda: case_linear(32768, 768, 768, residual=True),
    "small_256": lambda: case_linear(256, 768, 768, out=torch.float32),
    "self_scores_128": case_self_scores,
    "self_pv_128": case_self_pv,
    "self_scores_50": lambda: case_self_scores(Lq=50 + 6),   # 56: keeps the 8-element stride rule of this harness
    "cma_scores_128x50": case_cma_scores,
    "cma_pv_128x50": case_cma_pv,
    "cma_dv_128x50": case_cma_dv,
    "cma_scores_50x128": lambda: case_cma_scores(Lq=50, Lc=128),
    "cma_pv_50x128": lambda: case_cma_pv(Lq=50, Lc=128),
}


def run(name, lines, iters=10):
    (a, b, c), kw, flops = CASES[name]()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        K.gemm(a, b, c, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.gemm(a, b, c, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2] * 1e3
    rec = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    slots = L.lib.d2r_gemm_set_profile(C.c_void_p(rec.data_ptr()))
    assert slots == 16
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.gemm(a, b, c, **kw)
    e1.record()
    torch.cuda.synchronize()
    L.lib.d2r_gemm_set_profile(None)
    r = rec.view(148, 16).cpu().double()
    mma = r[r[:, 7] > 0]          # CTAs that issued MMAs (leaders of a pair / every CTA of the single-CTA kernel)
    prod = r[r[:, 4] > 0]
    epi = r[r[:, 10] > 0]
    mean = lambda x: float(x.mean()) if x.numel() else 0.0
    loop = mean(mma[:, 7])
    lines.append(f"== {name}: {med:8.1f} us  {flops / med / 1e6:7.1f} TF/s   (with the probe on: {e0.elapsed_time(e1) * 1e3:.1f} us)  "
                 f"m={kw['m']} n={kw['n']} k={kw['k']} batch={kw.get('batch', 1)}")
    lines.append(f"   MMA issuer  ({len(mma):3d} CTAs): loop {loop:9.0f} cyc | tiles/CTA {mean(mma[:, 8]):6.1f} | wait operands "
                 f"{100 * mean(mma[:, 5]) / max(loop, 1):5.1f}% | wait TMEM accumulator {100 * mean(mma[:, 6]) / max(loop, 1):5.1f}% | "
                 f"issue+other {100 * (1 - (mean(mma[:, 5]) + mean(mma[:, 6])) / max(loop, 1)):5.1f}% | prologue {mean(mma[:, 11]):6.0f} cyc | "
                 f"first operands after {mean(mma[:, 12]):6.0f} cyc | loop max/min {float(mma[:, 7].max()) if len(mma) else 0:.0f}/{float(mma[:, 7].min()) if len(mma) else 0:.0f}")
    pl = mean(prod[:, 4])
    lines.append(f"   TMA producer({len(prod):3d} CTAs): loop {pl:9.0f} cyc | k-blocks/CTA {mean(prod[:, 3]):6.1f} | wait free stage "
                 f"{100 * mean(prod[:, 2]) / max(pl, 1):5.1f}% | cyc per k-block {pl / max(mean(prod[:, 3]), 1):6.0f}")
    el = mean(epi[:, 10])
    lines.append(f"   epilogue w0 ({len(epi):3d} CTAs): loop {el:9.0f} cyc | wait accumulator {100 * mean(epi[:, 9]) / max(el, 1):5.1f}%")
    print("\n".join(lines[-4:]), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gemm_stall.txt"))
    a = ap.parse_args()
    lines = []
    for name in (CASES if a.case == "all" else a.case.split(",")):
        run(name, lines)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
