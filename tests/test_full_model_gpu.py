"""BASELINE configs[2]: the reference's WHOLE model (UnimoModelF: BERT + CLIP-ViT towers, extra self-attention
layers, routed stacks, CLS poolers, Block fusion, js loss, classifier) with the B200 stack swapped in by
``d2r_b200.integration.accelerate`` against the same unmodified model on the same GPU, same weights, same batch.

Needs the staged reference (baseline/_ref, staged by __graft_entry__.build()); skipped when it is absent."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _models():
    from baseline import ref_loader as RL
    if not RL.available():
        pytest.skip("reference not staged (baseline/_ref)")
    from baseline.full_model import build_reference_model, synthetic_batch
    from d2r_b200.integration import accelerate
    ref, _ = build_reference_model(3, seed=5)
    acc = copy.deepcopy(ref)
    ref, acc = ref.cuda(), acc.cuda()
    keys = list(acc.state_dict().keys())
    ids = {n: id(p) for n, p in acc.named_parameters()}
    accelerate(acc)
    assert list(acc.state_dict().keys()) == keys                       # same checkpoint layout
    assert {n: id(p) for n, p in acc.named_parameters()} == ids        # the very same Parameter objects
    return ref, acc, synthetic_batch


def test_accelerated_full_model_logits_match_reference_fp32():
    ref, acc, synthetic_batch = _models()
    ref.eval()
    acc.eval()
    batch = synthetic_batch(4, 32, seed=3, device="cuda")
    with torch.no_grad():
        loss_r, logits_r = ref(*batch)
        loss_a, logits_a = acc(*batch)
    err = ((logits_a - logits_r).abs().max() / logits_r.abs().max()).item()
    assert err <= 1e-4, err
    assert abs(loss_a.item() - loss_r.item()) <= 1e-4 * max(1.0, abs(loss_r.item()))
    assert torch.equal(logits_a.argmax(-1), logits_r.argmax(-1))      # the predictions of modules/train.py:181


def test_accelerated_full_model_trains():
    """Train mode, bf16 autocast: finite loss close to the reference's, every parameter the reference trains gets a
    gradient and the never-used ones (SURVEY §8e caveat 3) stay without one."""
    ref, acc, synthetic_batch = _models()
    batch = synthetic_batch(4, 32, seed=4, device="cuda")
    for m in (ref, acc):
        m.eval()          # no dropout: the two runs must see the same network
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.train()
    out = {}
    for name, m in (("ref", ref), ("acc", acc)):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, logits = m(*batch)
        loss.backward()
        out[name] = (loss.item(), logits.float(), {n for n, p in m.named_parameters() if p.grad is None})
    assert torch.isfinite(out["acc"][1]).all()
    assert abs(out["acc"][0] - out["ref"][0]) <= 5e-2 * max(1.0, abs(out["ref"][0]))
    assert out["acc"][2] == out["ref"][2] and len(out["ref"][2]) == 110
    g_r = torch.cat([p.grad.flatten().float() for n, p in ref.named_parameters() if p.grad is not None])
    g_a = torch.cat([p.grad.flatten().float() for n, p in acc.named_parameters() if p.grad is not None])
    cos = torch.dot(g_r, g_a) / (g_r.norm() * g_a.norm())
    assert cos.item() >= 0.98, cos.item()
