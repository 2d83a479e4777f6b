"""The FLOP / byte accounting behind bench.py's roofline block, checked without a GPU: bench.kernel_roofline itself is
run on the CPU (a stand-in for the three torch.cuda calls it makes, the kernels replaced by the emulation of
tests/emu_kernels.py) over one fwd+bwd of both branch stacks at the benchmark's token counts, and the algorithmic FLOPs
it attributes to the tensor-core launches are compared with SURVEY.md section 8(a)/(d)'s analytic figure for the
reference: useful forward FLOPs per sample and routing layer 3.256 G (text branch) + 2.062 G (image branch), dead
reverse-attention branch excluded, backward = 2 x forward.  `roofline.achieved` may not be inflated by padding, by
counting the dead branch, or by launches that are not contractions."""
import argparse
import time
import types

import pytest
import torch

from oracle import d2r_oracle as O
from tests import emu_kernels as E


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)


def _fake_torch():
    cuda = types.SimpleNamespace(Event=_Event, _sleep=lambda n: None, synchronize=lambda: None)
    return types.SimpleNamespace(cuda=cuda, bfloat16=torch.bfloat16)


def test_roofline_counts_the_reference_algorithmic_flops(monkeypatch):
    from d2r_b200 import build
    build.build()
    import bench
    import d2r_b200.autograd as A
    import d2r_b200.kernels as K
    import d2r_b200.lanes as LN
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule, run_pair
    for name in dir(E):
        if not name.startswith("_") and callable(getattr(E, name)) and hasattr(K, name):
            monkeypatch.setattr(K, name, getattr(E, name))
    monkeypatch.setattr(A, "_require_cuda", lambda inputs: None)
    monkeypatch.setattr(LN, "ENABLED", False)

    args = argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")
    B, Lt, Li, R = 2, 128, 50, 3
    mt, mi = InteractionModule(args, R, 6, 128), Reversed_InteractionModule(args, R, 6, 128)
    mt.load_state_dict(O.make_params(2023, R, 6))
    mi.load_state_dict(O.make_params(2024, R, 6))
    text, image = O.make_inputs(2023, B, Lt, Li)
    t, i = text.bfloat16().requires_grad_(True), image.bfloat16().requires_grad_(True)

    def fwd_bwd():
        for m in (mt, mi):
            for p in m.parameters():
                p.grad = None
        (o1, s1), (o2, s2) = run_pair(mt, mi, t, i)
        (o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()).backward()

    tc_peak, hbm_peak = 1364.9, 6543.4
    roof, hbm = bench.kernel_roofline(fwd_bwd, K, _fake_torch(), tc_peak, hbm_peak)
    flops = roof["tflops"] * 1e12 * roof["ms"] / 1e3                       # per step, all tensor-core launches
    per_sample = flops / B
    analytic = 3 * R * (3.256e9 + 2.062e9)                                 # fwd + 2 x fwd, both branches, R layers
    print(f"counted {per_sample / 1e9:.2f} GFLOP/sample, analytic {analytic / 1e9:.2f}")
    # the first layer's input gradient of the raw context and a few [B,768] products are the only differences
    assert abs(per_sample - analytic) <= 0.04 * analytic, (per_sample / 1e9, analytic / 1e9)
    # aggregation: 12 launches per step (R layers x 2 branches x fwd/bwd)
    assert hbm["launches"] == 4 * R
    # every contraction of the step is a counted launch: 510 at the benchmark configuration
    assert roof["launches"] == 510, roof["launches"]
