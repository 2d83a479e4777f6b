"""Micro-benchmark of the HBM-bound aggregation / router-pool kernels (GPU box): CUDA events, L2 flushed,
algorithmic bytes (each distinct full tensor read or written once) / time vs the measured copy bandwidth."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, D, Kc = 256, 768, 6
    dt = torch.bfloat16
    for name, Ln in (("text", 128), ("image", 50)):
        x0 = torch.randn(B, Ln, D, device="cuda").to(dt)
        full = [x0, None] + [torch.randn(B, Ln, D, device="cuda").to(dt) for _ in range(3)] + [None]
        bvec = [None, torch.randn(B, D, device="cuda"), None, None, None, torch.randn(B, D, device="cuda")]
        tb = x0.numel() * 2
        for final in (False, True):
            n_out = 1 if final else Kc
            P = torch.rand(B, n_out, Kc, device="cuda") + 0.1
            gate = torch.zeros(B, Kc, device="cuda")
            d_outs = [torch.randn(B, Ln, D, device="cuda").to(dt) for _ in range(n_out)]
            d_pooled = None if final else torch.randn(n_out, B, D, device="cuda")
            inputs = [x0] * Kc if final else None
            tag = f"{name}{'_final' if final else ''}"
            if not a.only or a.only in "fwd":
                ms = timeit(lambda: K.aggregate_fwd(full, bvec, P, gate, final, inputs), a.iters, flush)
                byts = (4 + n_out) * tb
                print(f"agg_fwd  {tag:12s} {ms * 1e3:8.1f} us  {byts / 1e6:7.1f} MB  {byts / ms / 1e6:7.0f} GB/s  "
                      f"{byts / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
            if not a.only or a.only in "bwd":
                ms = timeit(lambda: K.aggregate_bwd(full, bvec, P, gate, final, d_outs, d_pooled, inputs), a.iters, flush)
                byts = (n_out + 8) * tb
                print(f"agg_bwd  {tag:12s} {ms * 1e3:8.1f} us  {byts / 1e6:7.1f} MB  {byts / ms / 1e6:7.0f} GB/s  "
                      f"{byts / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
        if not a.only or a.only in "pool":
            ms = timeit(lambda: K.pool_mean([x0]), a.iters, flush)
            print(f"pool_mean {name:11s} {ms * 1e3:8.1f} us  {tb / 1e6:7.1f} MB  {tb / ms / 1e6:7.0f} GB/s  "
                  f"{tb / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
    # reference point: a plain device copy of the same size with torch (what the peak was measured with)
    src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    ms = timeit(lambda: dst.copy_(src), a.iters, flush)
    print(f"torch copy 1 GiB        {ms * 1e3:8.1f} us  {2 * src.numel() / ms / 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
