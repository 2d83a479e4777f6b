"""Micro-benchmark of d2r_gemm shapes (GPU box): CUDA-event timing with an L2 flush between launches.
usage: python tools/gemm_bench.py [--case NAME] [--iters N]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200 import _lib as L  # noqa: E402

CASES = {
    # name: (m, n, k, a_mn, b_mn, extra kwargs)
    "fwd_nt": (32768, 768, 768, False, False, dict(bias=True, out="bf16")),
    "fwd_nt_bn128": (32768, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=128)),
    "fwd_nt_bn192": (32768, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=192)),
    "fwd_nt_bn256": (32768, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=256)),
    "dgrad_bn192": (32768, 768, 768, False, True, dict(out="bf16", tile_n=192)),
    "img_nt": (12800, 768, 768, False, False, dict(bias=True, out="bf16")),
    "img_nt_bn192": (12800, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=192)),
    "img_nt_bn256": (12800, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=256)),
    "img_nt_bn128": (12800, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=128)),
    "fwd_nt_res": (32768, 768, 768, False, False, dict(bias=True, out="bf16", residual=True, act=L.ACT_RELU)),
    "dgrad_nn": (32768, 768, 768, False, True, dict(out="bf16")),
    "wgrad_tt": (768, 768, 32768, True, True, dict(out="f32", split_k=16)),
    "wgrad_sk8": (768, 768, 32768, True, True, dict(out="f32", split_k=8)),
    "wgrad_img_sk16": (768, 768, 12800, True, True, dict(out="f32", split_k=16)),
    "wgrad_img_sk8": (768, 768, 12800, True, True, dict(out="f32", split_k=8)),
    "wgrad_wide_sk2": (4608, 768, 32768, True, True, dict(out="f32", split_k=2)),
    "wgrad_wide_sk4": (4608, 768, 32768, True, True, dict(out="f32", split_k=4)),
    "fwd_wide": (32768, 4608, 768, False, False, dict(bias=True, out="bf16")),
    "fwd_longk": (32768, 768, 4608, False, False, dict(out="bf16")),
    "fwd_nt_quad": (32768, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=1024)),
    "fwd_nt_pair": (32768, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=512)),
    "dgrad_quad": (32768, 768, 768, False, True, dict(out="bf16", tile_n=1024)),
    "img_nt_quad": (12800, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=1024)),
    "img_nt_pair": (12800, 768, 768, False, False, dict(bias=True, out="bf16", tile_n=512)),
    "wgrad_quad": (768, 768, 32768, True, True, dict(out="f32", split_k=16, tile_n=1024)),
    "fwd_wide_quad": (32768, 4608, 768, False, False, dict(bias=True, out="bf16", tile_n=1024)),
    "fwd_longk_quad": (32768, 768, 4608, False, False, dict(out="bf16", tile_n=1024)),
    "small_tc": (256, 768, 768, False, False, dict(bias=True, out="f32")),
    "small_tc_bn128": (256, 768, 768, False, False, dict(bias=True, out="f32", tile_n=128)),
    "small_tc_bn64": (256, 768, 768, False, False, dict(bias=True, out="f32", tile_n=64)),
    "small_dgrad_bn64": (256, 768, 768, False, True, dict(out="f32", tile_n=64)),
    "small_wgrad": (768, 768, 256, True, True, dict(out="f32")),
    "small_wgrad_bn64": (768, 768, 256, True, True, dict(out="f32", tile_n=64)),
    "img_dgrad": (12800, 768, 768, False, True, dict(out="bf16")),
    "img_dgrad_bn192": (12800, 768, 768, False, True, dict(out="bf16", tile_n=192)),
    "fwd_res_bn192": (32768, 768, 768, False, False, dict(bias=True, out="bf16", residual=True, act=L.ACT_RELU, tile_n=192)),
    "small_f32": (256, 768, 768, False, False, dict(bias=True, out="f32", f32=True)),
}


def run_case(name, iters, flush):
    m, n, k, a_mn, b_mn, kw = CASES[name]
    dt = torch.float32 if kw.get("f32") else torch.bfloat16
    a = torch.randn((k, m) if a_mn else (m, k), device="cuda").to(dt)
    b = torch.randn((k, n) if b_mn else (n, k), device="cuda").to(dt)
    c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16 if kw.get("out") == "bf16" else torch.float32)
    bias = torch.randn(n, device="cuda") if kw.get("bias") else None
    res = torch.randn(m, n, device="cuda").to(c.dtype) if kw.get("residual") else None
    args = dict(m=m, n=n, k=k, lda=a.shape[1], ldb=b.shape[1], ldc=n, a_mn=a_mn, b_mn=b_mn, bias=bias,
                residual=res, ldr=n if res is not None else 0, act=kw.get("act", 0), split_k=kw.get("split_k", 1),
                tile_n=kw.get("tile_n", 0))
    for _ in range(3):
        K.gemm(a, b, c, **args)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.gemm(a, b, c, **args)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:14s} m={m} n={n} k={k} {'T' if a_mn else 'N'}{'T' if b_mn else 'N'}  median {med * 1e3:8.1f} us  "
          f"min {ts[0] * 1e3:8.1f} us  {2.0 * m * n * k / med / 1e9:8.1f} TF/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    flush = None if a.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name in (CASES if a.case == "all" else a.case.split(",")):
        run_case(name, a.iters, flush)


if __name__ == "__main__":
    main()
