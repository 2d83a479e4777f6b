"""Benchmark-shape fixtures from the UNMODIFIED reference (BASELINE configs[0] / configs[1] token counts: 128 text +
50 image tokens, K = 6, R = 3, batch 8; and configs[3]: R = 4, 256 + 197 tokens, batch 2; train mode), digests only so
that the files stay small.

Run (only where /root/reference exists):  python tests/golden/make_benchshape_golden.py

Same recipe as make_golden.py (the reference's own InteractionModule / Reversed_InteractionModule imported from
/root/reference, synthetic parameters and inputs regenerated from seeds), with bench.py's seeds: parameters
make_params(2023) for the text branch and make_params(2024) for the image branch, inputs make_inputs(2023, 8, 128, 50),
loss = out.sum() + sim.sum() (bench.py's loss).  Stored: sim_paths and the routing probabilities in full (small),
[sum, abs-sum, 256 strided samples] of the output and of both input gradients, [sum, abs-sum, 16 samples] of every
parameter gradient, the names of the parameters that get no gradient."""
import os
import sys
import tempfile

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
sys.dont_write_bytecode = True

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("D2R_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import d2r_oracle as O  # noqa: E402
from tests.golden.make_golden import ref_args  # noqa: E402

# name, parameter seed, reversed, B, Lt, Li, R   (the deep_* cases: BASELINE configs[3], R = 4, 256 + 197 tokens)
CASES = [("bench_text_b8", 2023, False, 8, 128, 50, 3), ("bench_image_b8", 2024, True, 8, 128, 50, 3),
         ("deep_text_b2", 2023, False, 2, 256, 197, 4), ("deep_image_b2", 2024, True, 2, 256, 197, 4)]


def digest(t, n):
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return torch.cat([f.sum().view(1), f.abs().sum().view(1), f[idx]]).numpy()


def main():
    from models.InteractionModule import InteractionModule, Reversed_InteractionModule
    args = ref_args(tempfile.mkdtemp())
    torch.set_num_threads(os.cpu_count() or 8)
    for name, seed, rev, B, LT, LI, R in CASES:
        torch.manual_seed(0)
        m = (Reversed_InteractionModule if rev else InteractionModule)(args, num_layer_routing=R, num_cells=6, path_hid=128)
        m.load_state_dict(O.make_params(seed, R, 6))
        m.train(True)
        text, image = O.make_inputs(2023, B, LT, LI)
        text.requires_grad_(True)
        image.requires_grad_(True)
        probs = []
        hooks = [l.register_forward_hook(lambda mod, i, o: probs.append(o[1]))
                 for l in [m.dynamic_itr_l0, *m.dynamic_itr_l1, m.dynamic_itr_l2]]
        out, sim = m(text, image)
        loss = out[0].sum() + sim.sum()
        loss.backward()
        rec = {"sim": sim.detach().numpy(), "loss": np.float64(loss.item()), "out": digest(out[0], 256),
               "d_text": digest(text.grad, 256), "d_image": digest(image.grad, 256)}
        for i, p in enumerate(probs):
            rec[f"probs{i}"] = p.detach().numpy()
        dead = []
        for k, p in m.named_parameters():
            if p.grad is None:
                dead.append(k)
            else:
                rec["gd/" + k] = digest(p.grad, 16)
        rec["dead"] = np.array(dead)
        for h in hooks:
            h.remove()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "loss", loss.item(), "size KB", os.path.getsize(path) // 1024)


if __name__ == "__main__":
    main()
