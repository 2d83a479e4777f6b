"""Bring-up: does a CUDA-graph replay of the multi-stream schedule reproduce the single-stream results?"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import d2r_b200.lanes as LN  # noqa: E402
from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule, run_pair  # noqa: E402


def make_args():
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")


def main():
    torch.manual_seed(0)
    mt = InteractionModule(make_args(), 3, 6, 128).cuda().train()
    mi = Reversed_InteractionModule(make_args(), 3, 6, 128).cuda().train()
    B, Lt, Li = int(os.environ.get("B", 6)), int(os.environ.get("LT", 24)), int(os.environ.get("LI", 13))
    bf16 = bool(int(os.environ.get("BF16", 0)))
    tol = 3e-2 if bf16 else 1e-3
    t = torch.randn(B, Lt, 768, device="cuda", requires_grad=True)
    i = torch.randn(B, Li, 768, device="cuda", requires_grad=True)

    def step(pair):
        t.grad = None
        i.grad = None
        for m in (mt, mi):
            for p in m.parameters():
                p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            if pair:
                (o1, s1), (o2, s2) = run_pair(mt, mi, t, i)
            else:
                o1, s1 = mt(t, i)
                o2, s2 = mi(t, i)
        (o1[0].sum() + 2 * s1.sum() + 3 * o2[0].sum() + s2.sum()).backward()
        return [o1[0], s1, o2[0], s2]

    def snap(outs):
        d = {"out%d" % k: o.detach().clone() for k, o in enumerate(outs)}
        d["d_text"], d["d_image"] = t.grad.clone(), i.grad.clone()
        for b, m in (("t", mt), ("i", mi)):
            for k, p in m.named_parameters():
                if p.grad is not None:
                    d[b + "/" + k] = p.grad.clone()
        return d

    LN.ENABLED = False
    ref = snap(step(False))
    LN.ENABLED = True
    for name, pair, fwd, bwd in (("pair only", True, False, False), ("cells fwd", False, True, False),
                                 ("cells bwd", False, False, True), ("cells both", False, True, True),
                                 ("pair+cells", True, True, True)):
        LN.FWD_LANES, LN.BWD_LANES = fwd, bwd
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(pair)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = step(pair)
        worst = {}
        for rep in range(5):
            g.replay()
            torch.cuda.synchronize()
            got = snap(outs)
            for k in ref:
                sc = ref[k].abs().max().item()
                e = (got[k] - ref[k]).abs().max().item() / (sc + 1e-20)
                if sc > 1e-4 and e > worst.get(k, 0):
                    worst[k] = e
        bad = sorted(((e, k) for k, e in worst.items() if e > tol), reverse=True)
        print(f"{name:12s}: {len(bad)} tensors off; worst: {bad[:6]}", flush=True)
        del g


if __name__ == "__main__":
    main()
