"""Calibration points from the vendor libraries on the same GPU (NOT part of the product path): what cuBLAS reaches on
the projection shapes of the stack, and what torch's fused SDPA reaches on the self-attention shape.  CUDA events,
L2 flushed between iterations.  usage: python tools/calib_bench.py"""
import torch
import torch.nn.functional as F

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
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
    return ts[len(ts) // 2] * 1e3


for m, n, k in ((32768, 768, 768), (12800, 768, 768), (32768, 768, 4608), (32768, 4608, 768), (32768, 2304, 768),
                (8192, 8192, 8192)):
    a = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
    w = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: F.linear(a, w, bias))
    print(f"cuBLAS(Lt) linear  m={m} n={n} k={k}: {us:8.1f} us  {2.0 * m * n * k / us / 1e6:7.1f} TF/s", flush=True)
    at = torch.randn(k, m, device=dev, dtype=torch.bfloat16)
    if m == 32768 and k == 768 and n == 768:
        dy = torch.randn(m, n, device=dev, dtype=torch.bfloat16)
        us = timeit(lambda: dy.t() @ a)
        print(f"cuBLAS wgrad dy^T x  m={n} n={k} k={m}: {us:8.1f} us  {2.0 * m * n * k / us / 1e6:7.1f} TF/s", flush=True)

B, H, L, dk = 256, 16, 128, 48
for L in (128, 50):
    q, k_, v = (torch.randn(B, H, L, dk, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    us = timeit(lambda: F.scaled_dot_product_attention(q, k_, v))
    o = F.scaled_dot_product_attention(q, k_, v)
    do = torch.randn_like(o)
    usb = timeit(lambda: torch.autograd.grad(F.scaled_dot_product_attention(q, k_, v), (q, k_, v), do))
    print(f"torch SDPA self-attn B={B} H={H} L={L} dk={dk}: fwd {us:7.1f} us, fwd+bwd {usb:7.1f} us", flush=True)
# single-head d=768 cross attention through SDPA (Lq=128, Lc=50)
for Lq, Lc in ((128, 50), (50, 128), (128, 128)):
    q = torch.randn(B, 1, Lq, 768, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k_ = torch.randn(B, 1, Lc, 768, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(B, 1, Lc, 768, device=dev, dtype=torch.bfloat16, requires_grad=True)
    try:
        us = timeit(lambda: F.scaled_dot_product_attention(q, k_, v))
        o = F.scaled_dot_product_attention(q, k_, v)
        do = torch.randn_like(o)
        usb = timeit(lambda: torch.autograd.grad(F.scaled_dot_product_attention(q, k_, v), (q, k_, v), do))
        print(f"torch SDPA single-head d=768 Lq={Lq} Lc={Lc}: fwd {us:7.1f} us, fwd+bwd {usb:7.1f} us", flush=True)
    except Exception as e:
        print(f"torch SDPA single-head d=768 Lq={Lq} Lc={Lc}: {type(e).__name__}: {e}", flush=True)
