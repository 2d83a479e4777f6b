"""CPU-only: how far the REFERENCE ALGORITHM's own arithmetic modes are from each other, as the yardstick for the
parity tolerances.  Uses the oracle (pinned on reference-generated goldens) in float64, float32
and under torch.autocast("cpu", bfloat16).  `python tools/yardsticks.py > profiles/r02_cpu_yardsticks.txt`."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import d2r_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
mx = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
l2 = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()


def run(P, text, image, R, rev, dtype=torch.float32, autocast=False, grads=True):
    Pd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in P.items()}
    t, i = text.clone().to(dtype).requires_grad_(grads), image.clone().to(dtype).requires_grad_(grads)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out, sim, probs = O.stack_forward(Pd, t, i, R, 6, rev, True, {})
    if grads:
        (out[0].float().sum() + sim.float().sum()).backward() if dtype != torch.float64 else (out[0].sum() + sim.sum()).backward()
    return out[0].detach(), sim.detach(), [p.detach() for p in probs], t.grad, i.grad


def main():
    print("Reference algorithm (oracle) against itself, CPU, train mode, K = 6.  max = max-norm relative, L2 = relative L2.\n")
    print("1. bf16 autocast vs fp32 at the benchmark shape (B = 64, 128 + 50 tokens, R = 3): what 'bf16 parity' can mean")
    for seed, rev in ((2023, False), (2024, True)):
        P = O.make_params(seed, 3, 6)
        text, image = O.make_inputs(2023, 64, 128, 50)
        t0 = time.time()
        a, b = run(P, text, image, 3, rev), run(P, text, image, 3, rev, autocast=True)
        print(f"   {'image' if rev else 'text '} branch: outputs max {mx(b[0], a[0]):.3e} L2 {l2(b[0], a[0]):.3e} | routing probabilities max "
              f"{max(mx(x, y) for x, y in zip(b[2], a[2])):.3e} | sim_paths max {mx(b[1], a[1]):.3e} | d_own L2 "
              f"{l2(b[4] if rev else b[3], a[4] if rev else a[3]):.3e} d_context L2 {l2(b[3] if rev else b[4], a[3] if rev else a[4]):.3e}"
              f"   ({time.time() - t0:.0f} s)")
    print("   (this library on the GPU, same shape: outputs max 3.5-4.7e-2, L2 7e-3, probabilities 7e-5, sim 3e-5, input gradients "
          "L2 2.7e-2 / 6.0e-2:\n    profiles/r02_parity_report_v1.json)\n")
    print("2. fp32 vs float64: the backward behind the near-argmax softmaxes amplifies fp32 rounding on some inputs")
    for (B, Lt, Li, R, seed) in ((2, 133, 21, 3, 29), (2, 133, 21, 3, 31), (2, 100, 21, 3, 29), (3, 16, 5, 4, 7), (8, 128, 50, 3, 2023)):
        P = O.make_params(23, R, 6)
        text, image = O.make_inputs(seed, B, Lt, Li)
        a, b = run(P, text, image, R, False, torch.float64), run(P, text, image, R, False)
        print(f"   B={B} Lt={Lt} Li={Li} R={R} input seed {seed}: outputs max {mx(b[0], a[0]):.2e} | d_text max {mx(b[3], a[3]):.2e} "
              f"L2 {l2(b[3], a[3]):.2e} | d_image max {mx(b[4], a[4]):.2e} L2 {l2(b[4], a[4]):.2e}")
    print("   (this library's fp32 path on the GPU against the fp32 oracle at B = 64 / 256, 128 + 50 tokens: input gradients L2 "
          "1.8-4.0e-4 --\n    the size of the oracle's own fp32 rounding error in the last line)")
    print("   -> gradients are judged in L2 / cosine; max-norm gradient bounds are 5e-2 (tests/test_parity_gpu.py).")


if __name__ == "__main__":
    main()
