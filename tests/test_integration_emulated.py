"""CPU check of ``d2r_b200.integration.accelerate`` (BASELINE configs[2]): the reference's WHOLE model
(UnimoModelF, unmodified, from the staged baseline/_ref) with the routed stacks, the CLS poolers, the Block fusion
and js_div swapped in place -- executed with the CUDA kernels replaced by the torch-CPU emulation of
tests/emu_kernels.py, so what is pinned here is the host logic: the module surgery, the parameter re-binding, the
paired execution of the two stacks inside the reference's forward (modeling_unimo.py:842-843) and the hand-written
backward feeding the encoders.  The kernels themselves run in tests/test_full_model_gpu.py (-m gpu).

Skipped when the reference is not staged (``__graft_entry__.build()`` stages it)."""
import copy

import pytest
import torch

from tests import emu_kernels as E


@pytest.fixture(scope="module")
def models():
    from baseline import ref_loader as RL
    if not RL.available():
        pytest.skip("reference not staged (baseline/_ref)")
    from baseline.full_model import build_reference_model, synthetic_batch
    ref, _ = build_reference_model(3, seed=5)
    acc = copy.deepcopy(ref)
    return ref, acc, synthetic_batch


@pytest.fixture()
def emulated(monkeypatch):
    from d2r_b200 import build
    build.build()
    import d2r_b200.autograd as A
    import d2r_b200.kernels as K
    import d2r_b200.lanes as LN
    for name in dir(E):
        if not name.startswith("_") and callable(getattr(E, name)) and hasattr(K, name):
            monkeypatch.setattr(K, name, getattr(E, name))
    monkeypatch.setattr(A, "_require_cuda", lambda inputs: None)
    monkeypatch.setattr(LN, "ENABLED", False)
    yield


def _bn_train(m):
    m.eval()              # no dropout: both runs must see the same network
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.train()


def test_accelerate_keeps_checkpoint_layout_and_matches_reference(models, emulated):
    from d2r_b200.integration import accelerate
    ref, acc, synthetic_batch = models
    keys = list(acc.state_dict().keys())
    ids = {n: id(p) for n, p in acc.named_parameters()}
    accelerate(acc)
    assert list(acc.state_dict().keys()) == keys                       # same checkpoint layout and order
    assert {n: id(p) for n, p in acc.named_parameters()} == ids        # the very same Parameter objects
    accelerate(acc)                                                    # idempotent
    assert {n: id(p) for n, p in acc.named_parameters()} == ids

    # eval forward: loss, logits, predictions (modules/train.py:181)
    ref.eval()
    acc.eval()
    batch = synthetic_batch(3, 32, seed=3, device="cpu")
    with torch.no_grad():
        loss_r, logits_r = ref(*batch)
        loss_a, logits_a = acc(*batch)
    assert ((logits_a - logits_r).abs().max() / logits_r.abs().max()).item() <= 1e-5
    assert abs(loss_a.item() - loss_r.item()) <= 1e-5 * max(1.0, abs(loss_r.item()))
    assert torch.equal(logits_a.argmax(-1), logits_r.argmax(-1))

    # a deep copy of the accelerated model (EMA / best-checkpoint copies) is an independent accelerated model
    twin = copy.deepcopy(acc)
    twin.model.itr_module.__dict__["_d2r_graph_cache"] = {"stale": object()}     # never travels with a copy
    assert "_d2r_graph_cache" not in copy.deepcopy(twin.model.itr_module).__dict__
    assert twin.model.itr_module.__dict__["_d2r_partner"] is twin.model.Reversed_itr_module
    with torch.no_grad():
        assert torch.equal(twin(*batch)[1], logits_a)
    del twin

    # train-mode arithmetic (BatchNorm batch statistics), fp32: same loss, the same parameters receive a gradient
    # (the 110 never-used tensors of SURVEY §8e caveat 3 stay without one) and the gradients agree
    batch = synthetic_batch(4, 32, seed=4, device="cpu")
    res = []
    for m in (ref, acc):
        _bn_train(m)
        for p in m.parameters():
            p.grad = None
        loss, logits = m(*batch)
        loss.backward()
        res.append((loss.item(), {n for n, p in m.named_parameters() if p.grad is None},
                    {n: p.grad.detach().double() for n, p in m.named_parameters() if p.grad is not None}))
    (l_r, none_r, g_r), (l_a, none_a, g_a) = res
    assert abs(l_a - l_r) <= 1e-5 * max(1.0, abs(l_r))
    assert none_a == none_r and len(none_r) == 110
    num = sum(((g_a[n] - g_r[n]) ** 2).sum() for n in g_r) ** 0.5
    den = sum((g_r[n] ** 2).sum() for n in g_r) ** 0.5
    assert (num / den).item() <= 1e-3, (num / den).item()
    # the reverse mix -- model.train() with the batch-norm layers frozen by .eval() (every layer reads its OWN flag in
    # the reference, XModules.py:380-384): same loss, running statistics untouched
    for m in (ref, acc):
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.eval()
    before = {n: b.clone() for n, b in acc.named_buffers()}
    with torch.no_grad():
        l_r, l_a = ref(*batch)[0].item(), acc(*batch)[0].item()
    assert abs(l_a - l_r) <= 1e-5 * max(1.0, abs(l_r)), (l_a, l_r)
    for n, b in acc.named_buffers():
        assert torch.equal(b, before[n]), n
    # BatchNorm running statistics of the attention filtration moved the same way
    for (n, b_r), (_, b_a) in zip(ref.named_buffers(), acc.named_buffers()):
        if b_r.dtype.is_floating_point:
            torch.testing.assert_close(b_a, b_r, rtol=1e-4, atol=1e-6, msg=n)


@pytest.mark.parametrize("pair,head", [(False, True), (True, False)])
def test_accelerate_partial_swaps(models, emulated, pair, head):
    """``pair=False`` keeps the reference's two back-to-back stack calls; ``head=False`` leaves the CLS poolers, the
    Block fusion and js_div on the reference's own PyTorch code.  Same logits either way."""
    import sys
    from d2r_b200.integration import accelerate
    ref, _, synthetic_batch = models
    m = copy.deepcopy(ref)
    mod = sys.modules[type(m.model).__module__]
    js_before = mod.js_div
    try:
        accelerate(m, pair=pair, head=head)
        if head:
            # js_div: the reference module's global symbol is a dispatcher now -- fused kernel inside an accelerated
            # backbone's forward only; the unmodified twin keeps the reference's own function
            import d2r_b200.integration as I
            seen = []
            fused, original = I.js_div, mod.js_div._d2r_original
            monkey = pytest.MonkeyPatch()
            monkey.setattr(I, "js_div", lambda *a, **k: (seen.append("fused"), fused(*a, **k))[1])
            monkey.setattr(mod, "js_div",
                           I._js_dispatcher(lambda *a, **k: (seen.append("reference"), original(*a, **k))[1]))
            try:
                b = synthetic_batch(2, 32, seed=1, device="cpu")
                with torch.no_grad():
                    ref.eval()(*b)
                    assert seen and set(seen) == {"reference"}, seen
                    del seen[:]
                    m.eval()(*b)
                    assert seen and set(seen) == {"fused"}, seen
            finally:
                monkey.undo()
        assert type(m.model.block_fusion).__module__.startswith("d2r_b200") == head
        assert type(m.model.itr_module).__module__.startswith("d2r_b200")
        ref.eval()
        m.eval()
        batch = synthetic_batch(3, 32, seed=11, device="cpu")
        with torch.no_grad():
            loss_r, logits_r = ref(*batch)
            loss_a, logits_a = m(*batch)
        assert ((logits_a - logits_r).abs().max() / logits_r.abs().max()).item() <= 1e-5
        assert abs(loss_a.item() - loss_r.item()) <= 1e-5 * max(1.0, abs(loss_r.item()))
    finally:
        mod.js_div = js_before
