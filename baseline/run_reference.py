#!/usr/bin/env python
"""Time the UNMODIFIED reference ``InteractionModule`` + ``Reversed_InteractionModule`` (models/InteractionModule.py:9-55,
:61-108, called back to back as at models/modeling_unimo.py:842-843) on this box.  Prints ONE JSON line.

    python baseline/run_reference.py --device cpu  --batch 8   --steps 5 --warmup 2          # CPU arm (fp32)
    python baseline/run_reference.py --device cuda --batch 256 --steps 5 --warmup 2 --bf16   # stock PyTorch on the B200

The CPU arm must run with the GPU hidden (``CUDA_VISIBLE_DEVICES=""``; bench.py does that): the reference's dead-code
ContrastiveLoss moves a mask to CUDA whenever CUDA is available (XModules.py:223-227).  ``--bf16`` = the only bf16 mode
the reference supports, ``torch.autocast(bfloat16)`` around the forward.  One step = forward + backward of
both stacks, loss = out.sum() + sim_paths.sum() per branch, or forward only under no_grad with ``--eval``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--eval", action="store_true")
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--text-len", type=int, default=128)
    ap.add_argument("--image-tokens", type=int, default=50)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--full", action="store_true",
                    help="the whole UnimoModelF (encoders + stacks + head), loss = CE + js (unimo_model.py:149-162)")
    a = ap.parse_args()
    if a.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        # A parent that is one rank of a multi-GPU job may run with a CPU affinity narrowed to a few cores (inherited
        # by this process): the CPU arm is entitled to every host core.  Measured in round 2: 0.3-0.4 samples/s when
        # launched from a rank of a torchrun job against 25 from a plain process on the same box.
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except (AttributeError, OSError):
            pass
    import torch
    from baseline import ref_loader as RL
    RL.import_reference()
    from models.InteractionModule import InteractionModule, Reversed_InteractionModule   # the reference's own
    cores = os.cpu_count() or 1
    if a.device == "cpu":
        try:
            cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            pass
        torch.set_num_threads(a.threads or cores)
    dev = torch.device(a.device)
    if a.full:
        return run_full(a, dev, cores)
    torch.manual_seed(2023)
    args = RL.ref_args(DR_step=a.layers)
    mt = InteractionModule(args, num_layer_routing=a.layers, num_cells=6, path_hid=128).to(dev)
    mi = Reversed_InteractionModule(args, num_layer_routing=a.layers, num_cells=6, path_hid=128).to(dev)
    mt.train(not a.eval)
    mi.train(not a.eval)
    g = torch.Generator().manual_seed(2023)
    text = torch.randn(a.batch, a.text_len, 768, generator=g).to(dev).requires_grad_(not a.eval)
    image = torch.randn(a.batch, a.image_tokens, 768, generator=g).to(dev).requires_grad_(not a.eval)

    def step():
        if a.eval:
            with torch.no_grad(), torch.autocast(a.device, dtype=torch.bfloat16, enabled=a.bf16):
                o1, s1 = mt(text, image)
                o2, s2 = mi(text, image)
            return float(o1[0].float().sum() + o2[0].float().sum())
        for m in (mt, mi):
            for p in m.parameters():
                p.grad = None
        text.grad = None
        image.grad = None
        with torch.autocast(a.device, dtype=torch.bfloat16, enabled=a.bf16):
            o1, s1 = mt(text, image)
            o2, s2 = mi(text, image)
        loss = o1[0].float().sum() + s1.float().sum() + o2[0].float().sum() + s2.float().sum()
        loss.backward()
        return float(loss.detach())  # device -> host read of the loss, as the trainer does (train.py:123)

    for _ in range(a.warmup):
        step()
    ts = []
    if a.device == "cuda":
        torch.cuda.synchronize()
        for _ in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 1e3)
    else:
        for _ in range(a.steps):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    line = {"impl": "reference (unmodified models/InteractionModule.py from baseline/_ref)", "device": a.device,
            "batch": a.batch, "steps": a.steps, "warmup": a.warmup, "mode": "eval/no_grad" if a.eval else "train fwd+bwd",
            "dtype": "autocast-bf16" if a.bf16 else "fp32", "layers": a.layers, "text_len": a.text_len,
            "image_tokens": a.image_tokens, "median_s_per_step": med, "mean_s_per_step": sum(ts) / len(ts),
            "samples_per_s": a.batch / med, "samples_per_s_mean": a.batch * len(ts) / sum(ts),
            "cores": cores, "threads": torch.get_num_threads() if a.device == "cpu" else None,
            "torch": torch.__version__,
            "gpu": torch.cuda.get_device_name(0) if a.device == "cuda" else None,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30 if a.device == "cuda" else None}
    print(json.dumps(line), flush=True)


def run_full(a, dev, cores):
    import torch
    from baseline.full_model import build_reference_model, synthetic_batch
    model, _ = build_reference_model(a.layers, seed=2023)
    model = model.to(dev).train(not a.eval)
    batch = synthetic_batch(a.batch, a.text_len, seed=2023, device=dev)

    def step():
        for p in model.parameters():
            p.grad = None
        with torch.autocast(a.device, dtype=torch.bfloat16, enabled=a.bf16):
            loss, logits = model(*batch)
        loss.backward()
        return float(loss.detach())

    for _ in range(a.warmup):
        step()
    ts = []
    for _ in range(a.steps):
        if a.device == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        if a.device == "cuda":
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    print(json.dumps({"impl": "reference (unmodified UnimoModelF from baseline/_ref)", "device": a.device, "batch": a.batch,
                      "steps": a.steps, "warmup": a.warmup, "mode": "train fwd+bwd", "dtype": "autocast-bf16" if a.bf16 else "fp32",
                      "layers": a.layers, "text_len": a.text_len, "image_tokens": 50, "median_s_per_step": med,
                      "mean_s_per_step": sum(ts) / len(ts), "samples_per_s": a.batch / med,
                      "samples_per_s_mean": a.batch * len(ts) / sum(ts), "cores": cores,
                      "threads": torch.get_num_threads() if a.device == "cpu" else None, "torch": torch.__version__,
                      "gpu": torch.cuda.get_device_name(0) if a.device == "cuda" else None,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30 if a.device == "cuda" else None}), flush=True)


if __name__ == "__main__":
    main()
