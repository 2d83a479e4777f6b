"""Micro-benchmark of the fused attention kernels at the shapes of the benchmark step (GPU box): CUDA events, L2
flushed, against the HBM roofline (every distinct operand read / written once) -- these products have an arithmetic
intensity far below the ridge point (34 .. 130 FLOP/B against 209), i.e. they are HBM-bound by the roofline model."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from tools.agg_bench import timeit  # noqa: E402

bf = torch.bfloat16


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1, help="index of one case (default: all)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="fwd | bwd")
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, D = 256, 768
    cases = [("cross text 128x50", 1, 128, 50, 3.6, True), ("cross image 50x128", 1, 50, 128, 3.6, True),
             ("crcmc text 128x128", 1, 128, 128, 1.0, True), ("crcmc image 50x50", 1, 50, 50, 1.0, True),
             ("self text 16h 128", 16, 128, 128, 1 / math.sqrt(48), True), ("self image 16h 50", 16, 50, 50, 1 / math.sqrt(48), True)]
    if a.case >= 0:
        cases = [cases[a.case]]
    for name, H, Lq, Lc, alpha, res in cases:
        q = (torch.randn(B, Lq, D, device="cuda") * 0.3).to(bf)
        kv = (torch.randn(B, Lc, 2 * D, device="cuda") * 0.3).to(bf)
        k, v = kv[:, :, :D], kv[:, :, D:]
        x = torch.randn(B, Lq, D, device="cuda").to(bf)
        Lcp = (Lc + 7) // 8 * 8
        f = lambda: K.attn_fused_fwd(q, D, k, 2 * D, v, 2 * D, B=B, Lq=Lq, Lc=Lc, D=D, heads=H, alpha=alpha, p_ld=Lcp,
                                     residual=x if res else None)
        out, P, _ = f()
        ms = timeit(f, a.iters, flush) if a.only != "bwd" else float("nan")
        byts = 2 * B * (Lq * D * (3 if res else 2) + 2 * Lc * D + H * Lq * Lc)
        fl = 4.0 * B * Lq * Lc * D
        print(f"fwd {name:20s} {ms * 1e3:7.1f} us  {byts / 1e6:6.1f} MB  {byts / ms / 1e6:6.0f} GB/s  {byts / ms / 1e6 / peak:5.2f} of HBM peak  "
              f"{fl / ms / 1e9:6.1f} TF/s  intensity {fl / byts:5.1f} FLOP/B", flush=True)
        dO = torch.randn(B, Lq, D, device="cuda").to(bf)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        g = lambda: K.attn_fused_bwd(dO, D, 1.0, P, q, D, k, 2 * D, v, 2 * D, dq, D, dkv, 2 * D, dkv[:, :, D:], 2 * D,
                                     B=B, Lq=Lq, Lc=Lc, D=D, heads=H, alpha=alpha)
        ms = timeit(g, a.iters, flush) if a.only != "fwd" else float("nan")
        byts = 2 * B * (3 * Lq * D + 4 * Lc * D + H * Lq * Lc)
        fl = 8.0 * B * Lq * Lc * D
        print(f"bwd {name:20s} {ms * 1e3:7.1f} us  {byts / 1e6:6.1f} MB  {byts / ms / 1e6:6.0f} GB/s  {byts / ms / 1e6 / peak:5.2f} of HBM peak  "
              f"{fl / ms / 1e9:6.1f} TF/s  intensity {fl / byts:5.1f} FLOP/B", flush=True)


if __name__ == "__main__":
    main()
