"""Micro-benchmark of the batched small-K GEMMs of the attention blocks (GPU box), CUDA events + L2 flush."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200 import _lib as L  # noqa: E402

bf = torch.bfloat16


def timeit(fn, flush, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, H, D = 256, 16, 768
    dh = D // H
    for Lq, Lc in ((128, 128), (50, 50)):
        Lcp = (Lc + 7) // 8 * 8
        qkv = torch.randn(B, Lq, 3 * D, device="cuda").to(bf)
        P = torch.empty(B, H, Lq, Lcp, device="cuda", dtype=bf)
        S = torch.empty(B, H, Lq, Lcp, device="cuda", dtype=torch.float32)
        out = torch.empty(B, Lq, D, device="cuda", dtype=bf)
        kw = dict(m=Lq, n=Lc, k=dh, lda=3 * D, ldb=3 * D, ldc=Lcp, batch=B * H, batch_inner=H,
                  a_str=(Lq * 3 * D, dh), b_str=(Lc * 3 * D, dh), c_str=(H * Lq * Lcp, Lq * Lcp), alpha=0.14)
        t = timeit(lambda: K.gemm(qkv, qkv[:, :, D:], P, epilogue=L.EPI_SOFTMAX, **kw), flush)
        print(f"self scores+softmax L={Lq}: {t:7.1f} us", flush=True)
        dS = torch.empty_like(P)
        kwb = dict(kw, alpha=1.0)
        t = timeit(lambda: K.gemm(qkv, qkv[:, :, 2 * D:], dS, epilogue=L.EPI_SOFTMAX_BWD, residual=P, ldr=Lcp,
                                  r_str=(H * Lq * Lcp, Lq * Lcp), **kwb), flush)
        print(f"self dP+softmax_bwd L={Lq}: {t:7.1f} us", flush=True)
        t = timeit(lambda: K.gemm(qkv, qkv[:, :, D:], S, **kw), flush)
        t2 = timeit(lambda: K.softmax_fwd(S, Lc, 1.0, bf), flush)
        print(f"self scores (fp32 S) L={Lq}: {t:7.1f} us  + softmax kernel {t2:7.1f} us", flush=True)
        t = timeit(lambda: K.gemm(P, qkv[:, :, 2 * D:], out, m=Lq, n=dh, k=Lc, lda=Lcp, ldb=3 * D, ldc=D, b_mn=True,
                                  batch=B * H, batch_inner=H, a_str=(H * Lq * Lcp, Lq * Lcp), b_str=(Lc * 3 * D, dh),
                                  c_str=(Lq * D, dh)), flush)
        print(f"self P V            L={Lq}: {t:7.1f} us", flush=True)
    for Lq, Lc in ((128, 50), (50, 128), (128, 128)):
        Lcp = (Lc + 7) // 8 * 8
        q = torch.randn(B, Lq, D, device="cuda").to(bf)
        kv = torch.randn(B, Lc, 6 * D, device="cuda").to(bf)
        P = torch.empty(B, 1, Lq, Lcp, device="cuda", dtype=bf)
        out = torch.empty(B, Lq, D, device="cuda", dtype=bf)
        t = timeit(lambda: K.gemm(q, kv, P, m=Lq, n=Lc, k=D, lda=D, ldb=6 * D, ldc=Lcp, batch=B, batch_inner=1,
                                  a_str=(Lq * D, D), b_str=(Lc * 6 * D, D), c_str=(Lq * Lcp, Lq * Lcp), alpha=3.6,
                                  epilogue=L.EPI_SOFTMAX), flush)
        print(f"cross scores+softmax Lq={Lq} Lc={Lc}: {t:7.1f} us", flush=True)
        for tn in (0, 128):
            t = timeit(lambda: K.gemm(P, kv[:, :, D:], out, m=Lq, n=D, k=Lc, lda=Lcp, ldb=6 * D, ldc=D, b_mn=True,
                                      batch=B, batch_inner=1, a_str=(Lq * Lcp, Lq * Lcp), b_str=(Lc * 6 * D, D),
                                      c_str=(Lq * D, D), tile_n=tn), flush)
            print(f"cross P V tile_n={tn:3d}     Lq={Lq} Lc={Lc}: {t:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
