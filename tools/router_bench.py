"""Micro-benchmark of the router kernels (GPU box): pool_mean, router_head_fwd/bwd, saf_fwd/bwd at the benchmark
shape, CUDA events with an L2 flush; algorithmic bytes = every distinct operand read / written once."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from tools.agg_bench import timeit  # noqa: E402


def main():
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, D, Kc, H = 256, 768, 6, 768
    for name, Ln in (("text", 128), ("image", 50)):
        x = torch.randn(B, Ln, D, device="cuda").to(torch.bfloat16)
        ms = timeit(lambda: K.pool_mean([x]), 10, flush)
        byts = x.numel() * 2
        print(f"pool_mean       {name:6s} {ms * 1e3:8.1f} us  {byts / 1e6:7.1f} MB  {byts / ms / 1e6:7.0f} GB/s  "
              f"{byts / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
        sg = torch.randn(B, D, device="cuda").to(torch.bfloat16)
        sl = torch.randn(B, Ln, D, device="cuda").to(torch.bfloat16)
        w, bias = torch.randn(D, device="cuda") / 28, torch.zeros(1, device="cuda")
        bn_w, bn_b = torch.ones(1, device="cuda"), torch.zeros(1, device="cuda")
        rm, rv = torch.zeros(1, device="cuda"), torch.ones(1, device="cuda")
        nbt = torch.zeros((), dtype=torch.long, device="cuda")
        out, saved = K.saf_fwd(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, True)
        ms = timeit(lambda: K.saf_fwd(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, True), 10, flush)
        byts = sl.numel() * 2
        print(f"saf_fwd         {name:6s} {ms * 1e3:8.1f} us  {byts / 1e6:7.1f} MB (read once)  {byts / ms / 1e6:7.0f} GB/s  "
              f"{byts / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
        d_out = torch.randn(B, D, device="cuda")
        ms = timeit(lambda: K.saf_bwd(d_out, sg, sl, w, bias, bn_w, bn_b, rm, rv, True, saved), 10, flush)
        byts = sl.numel() * 2 * 2
        print(f"saf_bwd         {name:6s} {ms * 1e3:8.1f} us  {byts / 1e6:7.1f} MB (read + write)  {byts / ms / 1e6:7.0f} GB/s  "
              f"{byts / ms / 1e6 / peak:5.2f} of measured peak", flush=True)
    for final in (False, True):
        n_out = 1 if final else Kc
        hid = torch.relu(torch.randn(Kc, B, H, device="cuda"))
        w2 = [torch.randn(n_out, H, device="cuda") / 28 for _ in range(Kc)]
        b2 = [torch.full((n_out,), 1.5, device="cuda") for _ in range(Kc)]
        raw, norm, gate = K.router_head_fwd(hid, w2, b2, n_out, final)
        d_norm = torch.randn(B, n_out, Kc, device="cuda")
        ms = timeit(lambda: K.router_head_fwd(hid, w2, b2, n_out, final), 10, flush)
        byts = hid.numel() * 4
        print(f"router_head_fwd final={int(final)} {ms * 1e3:8.1f} us  {byts / 1e6:6.2f} MB  {byts / ms / 1e6:7.0f} GB/s", flush=True)
        ms = timeit(lambda: K.router_head_bwd(d_norm, raw, hid, w2, final), 10, flush)
        byts = hid.numel() * 4 * 2
        print(f"router_head_bwd final={int(final)} {ms * 1e3:8.1f} us  {byts / 1e6:6.2f} MB  {byts / ms / 1e6:7.0f} GB/s "
              f"(incl. the host-side zero-arena views)", flush=True)


if __name__ == "__main__":
    main()
