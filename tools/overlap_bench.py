"""Do two persistent GEMM launches on two CUDA streams fill each other's tail waves?  (GPU box)
Times n_pairs x (GEMM a ; GEMM b) issued serially on one stream vs. on two streams, CUDA events around the lot."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402


def mk(m, n, k):
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    c = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    return a, b, c, dict(m=m, n=n, k=k, lda=k, ldb=k, ldc=n)


def run(shapes, two_streams, reps=20):
    ops = [mk(*s) for s in shapes]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for it in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2e-3 * 1.9e9))
        cur = torch.cuda.current_stream()
        e0.record()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(reps):
            for i, (a, b, c, kw) in enumerate(ops):
                with torch.cuda.stream(s2 if (two_streams and i % 2) else s1):
                    K.gemm(a, b, c, **kw)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


if __name__ == "__main__":
    for shapes in ([(12800, 768, 768), (12800, 768, 768)], [(32768, 768, 768), (12800, 768, 768)],
                   [(256, 768, 768), (32768, 768, 768)], [(12800, 768, 768), (256, 768, 768)]):
        a = run(shapes, False)
        b = run(shapes, True)
        print(f"{shapes}: one stream {a:7.1f} us per pair, two streams {b:7.1f} us per pair", flush=True)
