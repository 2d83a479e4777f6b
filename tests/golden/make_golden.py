"""Generate golden vectors by executing the UNMODIFIED reference in the authoring container.

Run (only where /root/reference exists):  python tests/golden/make_golden.py

The reference has no golden vectors of its own, so parity is pinned on
the reference itself: this script imports ``models.InteractionModule`` from
``/root/reference`` (read-only, nothing is copied), loads the deterministic synthetic
parameters of ``oracle.d2r_oracle.make_params`` into it, runs forward + backward on the
seeded inputs of ``make_inputs`` and stores outputs / gradient digests in
``tests/golden/*.npz``.  Weights are *not* stored: they are regenerated from the seed.

Harness-side shims only: sys.path, local BertConfig/CLIPConfig dirs,
CUDA hidden, bytecode writing off.
"""
import argparse
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
from tests.golden.cases import CASES, PARAM_SEED_BASE, INPUT_SEED_BASE, LOSS_SEED  # noqa: E402



def ref_args(tmp):
    from transformers import BertConfig, CLIPConfig
    bd, vd = os.path.join(tmp, "bert"), os.path.join(tmp, "clip")
    BertConfig().save_pretrained(bd)
    CLIPConfig().save_pretrained(vd)
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0,
                              alpha=0, margin=0.1, bert_name=bd, vit_name=vd)


def digest(t: torch.Tensor) -> np.ndarray:
    """Small, order-sensitive summary of a big gradient: sum, abs-sum, and 16 strided samples."""
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 16).long()
    return torch.cat([f.sum().view(1), f.abs().sum().view(1), f[idx]]).numpy()


def main():
    from models.InteractionModule import InteractionModule, Reversed_InteractionModule
    tmp = tempfile.mkdtemp()
    args = ref_args(tmp)
    for name, B, Lt, Li, R, rev, training, realistic, scale in CASES:
        torch.manual_seed(0)
        cls = Reversed_InteractionModule if rev else InteractionModule
        m = cls(args, num_layer_routing=R, num_cells=6, path_hid=128)
        P = O.make_params(PARAM_SEED_BASE + R, R, 6, scale)
        sd = m.state_dict()
        spec_keys = [n for n, _, _ in O.stack_param_spec(R, 6)]
        assert list(sd.keys()) == spec_keys, "state_dict key order differs from oracle spec"
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(P[k].shape), (k, v.shape, P[k].shape)
        m.load_state_dict(P)
        m.train(training)
        text, image = O.make_inputs(INPUT_SEED_BASE + B, B, Lt, Li, realistic=realistic)
        text.requires_grad_(True)
        image.requires_grad_(True)
        # capture per-layer probabilities through forward hooks on the routing layers
        probs = []
        hooks = [m.dynamic_itr_l0.register_forward_hook(lambda mod, i, o: probs.append(o[1]))]
        for l in m.dynamic_itr_l1:
            hooks.append(l.register_forward_hook(lambda mod, i, o: probs.append(o[1])))
        hooks.append(m.dynamic_itr_l2.register_forward_hook(lambda mod, i, o: probs.append(o[1])))
        out, sim = m(text, image)
        out = out[0]
        g = torch.Generator().manual_seed(LOSS_SEED)
        w_out = torch.randn(out.shape, generator=g)
        w_sim = torch.randn(sim.shape, generator=g)
        loss = (out * w_out).sum() + (sim * w_sim).sum()
        rec = {"out": out.detach().numpy(), "sim": sim.detach().numpy(), "loss": np.float64(loss.item())}
        for i, p in enumerate(probs):
            rec[f"probs{i}"] = p.detach().numpy()
        if True:
            loss.backward()
            rec["d_text"] = text.grad.numpy()
            rec["d_image"] = image.grad.numpy()
            dead = []
            for k, p in m.named_parameters():
                if p.grad is None:
                    dead.append(k)
                else:
                    rec["gd/" + k] = digest(p.grad)
            rec["dead"] = np.array(dead)
        for k, v in m.state_dict().items():
            if "SAF_module.bn.running" in k or "SAF_module.bn.num_batches" in k:
                rec["buf/" + k] = v.numpy()
        for h in hooks:
            h.remove()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        meta = dict(B=B, Lt=Lt, Li=Li, R=R, rev=rev, training=training, realistic=realistic)
        print(name, meta, "loss", loss.item(), "out|max|", out.abs().max().item(),
              "size KB", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024)


if __name__ == "__main__":
    main()
