"""Per-wrapper CUDA-event breakdown of one fwd+bwd step of both branch stacks (run on the GPU box).
Every function of d2r_b200.kernels is wrapped with a pair of events on the launching stream; GEMMs are keyed
by shape.  Writes a table sorted by total time to stdout (and gpurun_out/step_breakdown.txt)."""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import d2r_b200.lanes as LN  # noqa: E402
from d2r_b200 import kernels as K  # noqa: E402

LN.ENABLED = False        # one stream: kernels of concurrent lanes would overlap inside each other's event pairs
from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--config", default="stack", choices=["stack", "deep"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_breakdown.txt"))
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(2023)
    cfg = bench.CONFIGS[a.config]
    mt = InteractionModule(bench.make_args(), cfg["R"], bench.KC, 128).to(dev).train()
    mi = Reversed_InteractionModule(bench.make_args(), cfg["R"], bench.KC, 128).to(dev).train()
    text = torch.randn(a.batch, cfg["Lt"], bench.D, device=dev, requires_grad=True)
    image = torch.randn(a.batch, cfg["Li"], bench.D, device=dev, requires_grad=True)

    def step():
        for m in (mt, mi):
            for p in m.parameters():
                p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not a.fp32):
            o1, s1 = mt(text, image)
            o2, s2 = mi(text, image)
        (o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()).backward()

    recs = []
    names = [n for n in dir(K) if callable(getattr(K, n)) and not n.startswith("_") and
             getattr(getattr(K, n), "__module__", "") == K.__name__ and n != "linear" and
             not n.endswith("_supported") and n not in ("zeros_f32", "zero_arena_reset")]
    orig = {n: getattr(K, n) for n in names}

    def wrap(n, fn):
        def w(*args, **kw):
            key = n
            if n == "gemm":
                t = args[0]
                key = (f"gemm[{'bf16' if t.dtype == torch.bfloat16 else 'f32'}] m={kw['m']} n={kw['n']} k={kw['k']} "
                       f"b={kw.get('batch', 1)} {'T' if kw.get('a_mn') else 'N'}{'T' if kw.get('b_mn') else 'N'}"
                       f" sk={kw.get('split_k', 1)}")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            fl = 2.0 * kw["m"] * kw["n"] * kw["k"] * kw.get("batch", 1) if n == "gemm" else 0.0
            recs.append((key, e0, e1, fl))
            return r
        return w

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    for n in names:
        setattr(K, n, wrap(n, orig[n]))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # park the GPU so that the host enqueues the whole step ahead of it: the event pairs then bracket
    # back-to-back kernel execution instead of host launch latency
    torch.cuda._sleep(int(0.15 * 1.9e9))
    s0.record()
    step()
    s1.record()
    torch.cuda.synchronize()
    for n in names:
        setattr(K, n, orig[n])
    agg = collections.OrderedDict()
    for key, e0, e1, fl in recs:
        d = agg.setdefault(key, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        d[2] += fl
    total = s0.elapsed_time(s1)
    lines = [f"step {total:.2f} ms (eager, one stream, GPU parked while the host enqueues; events on); sum of wrapped calls {sum(v[1] for v in agg.values()):.2f} ms"]
    for key, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tf = f"{fl / ms / 1e9:8.1f} TF/s" if fl else " " * 13
        lines.append(f"{ms:8.3f} ms {100 * ms / total:5.1f}%  x{cnt:<4d} {ms / cnt * 1e3:8.1f} us/call {tf}  {key}")
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
