"""Timing of the SURVEY §8f rows built so far (GPU box): XModules.Block and XModules.js_div, forward + backward at the
benchmark batch (256), CUDA events over repeated calls, next to the oracle restatement on the host cores."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import d2r_oracle as O  # noqa: E402  (baseline timing only)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200.interaction.XModules import Block, js_div  # noqa: E402


def gpu_time(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n0 = K.L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, (K.L.launch_count() - n0) // iters


def cpu_time(fn, iters=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e6


def main():
    B = 256
    P = O.make_block_params(77)
    m = Block([768, 768], 768)
    m.load_state_dict(P)
    m = m.cuda()
    g = torch.Generator().manual_seed(1)
    h0, h1 = torch.tanh(torch.randn(B, 768, generator=g)), torch.tanh(torch.randn(B, 768, generator=g))
    for bf16 in (True, False):
        x0, x1 = h0.cuda().requires_grad_(True), h1.cuda().requires_grad_(True)

        def step():
            for p in m.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
                out = m([x0, x1])
            out.float().sum().backward()
        us, launches = gpu_time(step)
        print(f"Block fwd+bwd B={B} {'bf16' if bf16 else 'fp32'}: {us:8.1f} us/call eager ({launches} d2r launches)", flush=True)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    c0, c1 = h0.clone().requires_grad_(True), h1.clone().requires_grad_(True)
    us = cpu_time(lambda: O.block_fusion(Pg, c0, c1).sum().backward())
    print(f"Block fwd+bwd B={B} oracle (torch CPU fp32, {torch.get_num_threads()} threads): {us:8.1f} us/call", flush=True)
    p, q = (torch.randn(B, B, device="cuda") * 4).requires_grad_(True), (torch.randn(B, B, device="cuda") * 4).requires_grad_(True)
    us, launches = gpu_time(lambda: js_div(p, q).backward())
    print(f"js_div fwd+bwd [{B},{B}]: {us:8.1f} us/call eager ({launches} d2r launches)", flush=True)
    pc, qc = p.detach().cpu().requires_grad_(True), q.detach().cpu().requires_grad_(True)
    us = cpu_time(lambda: O.js_div(pc, qc).backward(), iters=20)
    print(f"js_div fwd+bwd [{B},{B}] oracle (torch CPU): {us:8.1f} us/call", flush=True)


if __name__ == "__main__":
    main()
