"""Micro-benchmark of the HBM-bound element-wise kernels (GPU box): CUDA events, L2 flushed between launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200 import _lib as L  # noqa: E402


def timeit(fn, flush, iters=20):
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
    for rows, cols in ((32768, 768), (12800, 768), (32768, 4608), (256, 768)):
        dy = torch.randn(rows, cols, device="cuda").bfloat16()
        y = torch.randn(rows, cols, device="cuda").bfloat16()
        for act, name, passes in ((L.ACT_NONE, "none", 1), (L.ACT_RELU, "relu", 3)):
            us = timeit(lambda: K.bias_act_bwd(dy, y, act, True, True), flush)
            gb = passes * rows * cols * 2 / 1e9
            print(f"bias_act_bwd[{name}] {rows}x{cols}: {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s", flush=True)
        x = torch.randn(rows, cols, device="cuda").bfloat16()
        us = timeit(lambda: x.clone(), flush)
        print(f"torch clone           {rows}x{cols}: {us:7.1f} us  {2 * rows * cols * 2 / 1e9 / us * 1e6:7.0f} GB/s", flush=True)
        us = timeit(lambda: x.sum(0), flush)
        print(f"torch sum(0)          {rows}x{cols}: {us:7.1f} us  {rows * cols * 2 / 1e9 / us * 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
