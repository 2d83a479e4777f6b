"""CPU check of the product's host-side logic: stack.py / autograd.py / the nn.Module mirror are executed
with the CUDA kernels replaced by the torch-CPU emulation in tests/emu_kernels.py (test infrastructure, see
its header) and compared with the reference-generated golden vectors and the oracle.  This pins the launch
sequence, the stride bookkeeping and the hand-written backward pass without a GPU; the kernels themselves
are checked on the GPU (-m gpu)."""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import d2r_oracle as O
from tests import emu_kernels as E
from tests.golden.cases import CASES, PARAM_SEED_BASE, INPUT_SEED_BASE, LOSS_SEED

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture()
def emulated(monkeypatch):
    from d2r_b200 import build
    build.build()
    import d2r_b200.kernels as K
    import d2r_b200.autograd as A
    for name in dir(E):
        if not name.startswith("_") and callable(getattr(E, name)) and hasattr(K, name):
            monkeypatch.setattr(K, name, getattr(E, name))
    monkeypatch.setattr(A, "_require_cuda", lambda inputs: None)

    import d2r_b200.lanes as LN
    monkeypatch.setattr(LN, "ENABLED", False)   # no CUDA streams on the CPU
    yield


def make_args():
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")


def relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def digest(t):
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 16).long()
    return torch.cat([f.sum().view(1), f.abs().sum().view(1), f[idx]]).numpy()


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[2]], ids=[c[0] for c in CASES[:3]])
def test_emulated_stack_matches_reference_golden(emulated, case):
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule
    name, B, Lt, Li, R, rev, training, realistic, scale = case
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    P = O.make_params(PARAM_SEED_BASE + R, R, 6, scale)
    m = (Reversed_InteractionModule if rev else InteractionModule)(make_args(), R, 6, 128)
    m.load_state_dict(P)
    m.train(training)
    text, image = O.make_inputs(INPUT_SEED_BASE + B, B, Lt, Li, realistic=realistic)
    text.requires_grad_(True)
    image.requires_grad_(True)
    g = torch.Generator().manual_seed(LOSS_SEED)
    w_out = torch.randn(gold["out"].shape, generator=g)
    w_sim = torch.randn(gold["sim"].shape, generator=g)
    out, sim, probs = m(text, image, return_path_probs=True)
    ((out[0] * w_out).sum() + (sim * w_sim).sum()).backward()
    for li, p in enumerate(probs):
        assert relerr(p, gold[f"probs{li}"]) < 1e-5
    assert relerr(out[0], gold["out"]) < 1e-5 and relerr(sim, gold["sim"]) < 1e-5
    assert relerr(text.grad, gold["d_text"]) < 1e-4, relerr(text.grad, gold["d_text"])
    assert relerr(image.grad, gold["d_image"]) < 1e-4, relerr(image.grad, gold["d_image"])
    dead = set(gold["dead"].tolist())
    for k, p in m.named_parameters():
        if k in dead:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        ref, got = gold["gd/" + k], digest(p.grad)
        if O.is_zero_grad_param(k, training):
            assert abs(got[1]) < 1e-2, k
            continue
        assert abs(got[1] - ref[1]) <= 1e-3 * abs(ref[1]), (k, got[1], ref[1])
        assert np.abs(got[2:] - ref[2:]).max() <= 1e-3 * np.abs(ref[2:]).max() + 1e-7, k
    for k in gold.files:
        if k.startswith("buf/"):
            np.testing.assert_allclose(m.state_dict()[k[4:]].numpy(), gold[k], rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("K,R", [(4, 2), (6, 2)])
def test_emulated_reference_derived(emulated, K, R):
    from d2r_b200.interaction import InteractionModule
    B, Lt, Li = 3, 10, 6
    P = O.make_params(5, R, K)
    text, image = O.make_inputs(17, B, Lt, Li)
    t, i = text.clone().requires_grad_(True), image.clone().requires_grad_(True)
    ref_out, ref_sim, _ = O.stack_forward(P, t, i, R, K, False, training=True, bn_updates={})
    (ref_out[0].sum() + ref_sim.sum()).backward()
    m = InteractionModule(make_args(), R, K, 128)
    m.load_state_dict(P)
    t2, i2 = text.clone().requires_grad_(True), image.clone().requires_grad_(True)
    out, sim = m(t2, i2)
    (out[0].sum() + sim.sum()).backward()
    assert relerr(out[0], ref_out[0].detach()) < 1e-5 and relerr(sim, ref_sim.detach()) < 1e-5
    assert relerr(t2.grad, t.grad) < 1e-4 and relerr(i2.grad, i.grad) < 1e-4


def test_emulated_submodules(emulated):
    """Stand-alone use of the mirrored classes (cells / layer), forward and backward, against the oracle."""
    from d2r_b200.interaction import InteractionModule
    P = O.make_params(9, 3, 6)
    m = InteractionModule(make_args(), 3, 6, 128)
    m.load_state_dict(P)
    m.eval()
    L0, pre = m.dynamic_itr_l0, "dynamic_itr_l0"
    text, image = O.make_inputs(4, 3, 10, 7)
    cells = [
        (lambda t, i: L0.ric(t), lambda t, i: O.cell_ric(t, P, pre + ".ric")),
        (lambda t, i: L0.imrc(t), lambda t, i: O.cell_imrc(t, P, pre + ".imrc")),
        (lambda t, i: L0.cmrc(t, i), lambda t, i: O.cell_cmrc(t, i, P, pre + ".cmrc")),
        (lambda t, i: L0.glac(t, i), lambda t, i: O.cell_glac(t, i, P, pre + ".glac", training=False)),
        (lambda t, i: L0.crcmc(t, i), lambda t, i: O.cell_crcmc(t, i, P, pre + ".crcmc")),
        (lambda t, i: L0.gesc(t, i), lambda t, i: O.cell_gesc(t, i, P, pre + ".gesc")),
    ]
    g = torch.Generator().manual_seed(1)
    w = torch.randn(3, 10, 768, generator=g)
    wp = torch.randn(3, 6, generator=g)
    for ci, (mine, ref) in enumerate(cells):
        ta, ia = text.clone().requires_grad_(True), image.clone().requires_grad_(True)
        tb, ib = text.clone().requires_grad_(True), image.clone().requires_grad_(True)
        e1, p1 = mine(ta, ia)
        e2, p2 = ref(tb, ib)
        assert relerr(e1, e2.detach()) < 1e-5 and relerr(p1, p2.detach()) < 1e-5, ci
        ((e1 * w).sum() + (p1 * wp).sum()).backward()
        ((e2 * w).sum() + (p2 * wp).sum()).backward()
        assert relerr(ta.grad, tb.grad) < 1e-4, ci
        if ib.grad is not None:
            assert relerr(ia.grad, ib.grad) < 1e-4, ci
    # a routing layer on its own
    outs, allp = L0(text, image)
    embs, probs = O.run_cells([text] * 6, image, P, pre, 6, False, None, False)
    r_outs, r_allp = O.aggregate_multi(embs, probs, 6)
    assert relerr(allp, r_allp) < 1e-5
    for a, b in zip(outs, r_outs):
        assert relerr(a, b) < 1e-5


def test_emulated_bf16_mode(emulated):
    """bf16 mode takes different code paths (batched router GEMM, staged/concatenated bf16 weights, casts of the
    small vectors); the emulation also enforces the TMA 16-byte stride rules of the tcgen05 path."""
    from d2r_b200.interaction import InteractionModule
    B, Lt, Li, R, K = 3, 12, 7, 3, 6
    P = O.make_params(5, R, K)
    text, image = O.make_inputs(17, B, Lt, Li)
    t, i = text.clone().requires_grad_(True), image.clone().requires_grad_(True)
    for k, v in P.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    ref_out, ref_sim, ref_probs = O.stack_forward(P, t, i, R, K, False, training=True, bn_updates={})
    (ref_out[0].sum() + ref_sim.sum()).backward()
    m = InteractionModule(make_args(), R, K, 128)
    m.load_state_dict({k: v.detach() for k, v in P.items()})
    t2 = text.clone().to(torch.bfloat16).requires_grad_(True)
    i2 = image.clone().to(torch.bfloat16).requires_grad_(True)
    out, sim, probs = m(t2, i2, return_path_probs=True)
    assert out[0].dtype == torch.float32
    (out[0].sum() + sim.sum()).backward()
    for a, b in zip(probs, ref_probs):
        assert relerr(a, b.detach()) < 2e-2
    assert relerr(out[0], ref_out[0].detach()) < 6e-2
    assert relerr(t2.grad.float(), t.grad) < 0.15 and relerr(i2.grad.float(), i.grad) < 0.15
    errs = {}
    for k, p in m.named_parameters():
        if O.is_dead_param(k) or O.is_zero_grad_param(k, True):
            continue
        errs[k] = relerr(p.grad, P[k].grad)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    # loose: bf16 activations + the cancellation in the normalised routing probabilities (sum_j P_ij = 1)
    # make individual gradients noisy; exact gradient parity is asserted in the fp32 mode above
    assert worst[0][1] < 0.6, worst


def test_emulated_run_pair_equals_separate_calls(emulated):
    """run_pair (both branch stacks as one autograd node) == the reference's two back-to-back module calls."""
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule, run_pair
    torch.manual_seed(5)
    mt = InteractionModule(make_args(), 3, 6, 128)
    mi = Reversed_InteractionModule(make_args(), 3, 6, 128)
    text0, image0 = O.make_inputs(INPUT_SEED_BASE + 3, 3, 9, 7, realistic=True)

    def run(pair):
        for m in (mt, mi):
            m.zero_grad(set_to_none=True)
        text, image = text0.clone().requires_grad_(True), image0.clone().requires_grad_(True)
        if pair:
            (o1, s1), (o2, s2) = run_pair(mt, mi, text, image)
        else:
            o1, s1 = mt(text, image)
            o2, s2 = mi(text, image)
        (o1[0].sum() + 2 * s1.sum() + 3 * o2[0].sum() + s2.sum()).backward()
        grads = {("t", k): p.grad.clone() for k, p in mt.named_parameters() if p.grad is not None}
        grads.update({("i", k): p.grad.clone() for k, p in mi.named_parameters() if p.grad is not None})
        return [o1[0], s1, o2[0], s2, text.grad, image.grad], grads

    a, ga = run(False)
    b, gb = run(True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert ga.keys() == gb.keys()
    for k in ga:
        assert torch.equal(ga[k], gb[k]), k


def test_emulated_layer_hooks_deliver_every_live_gradient(emulated):
    """GradAllReducer.install(): the stack's backward hands over each routing layer's finished gradients (last
    layer first); after wait() the flat bucket equals the .grad of every live parameter."""
    from d2r_b200.dp import GradAllReducer
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule, run_pair
    torch.manual_seed(3)
    mt = InteractionModule(make_args(), 3, 6, 128)
    mi = Reversed_InteractionModule(make_args(), 3, 6, 128)
    red = GradAllReducer([mt, mi])
    order = []
    red.install()
    orig = red.on_layer
    red.on_layer = lambda m, layer, grads: (order.append((m, layer)), orig(m, layer, grads))[1]
    red.install()                      # re-register so that the wrapped callback is the one installed
    text, image = O.make_inputs(INPUT_SEED_BASE + 2, 2, 6, 5)
    (o1, s1), (o2, s2) = run_pair(mt, mi, text.requires_grad_(True), image.requires_grad_(True))
    (o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()).backward()
    red.wait()
    assert order == [(0, "dynamic_itr_l2"), (0, "dynamic_itr_l1.0"), (0, "dynamic_itr_l0"),
                     (1, "dynamic_itr_l2"), (1, "dynamic_itr_l1.0"), (1, "dynamic_itr_l0")]
    for name, p, v in zip(red.names, red.params, red._views):
        assert p.grad is not None, name
        assert torch.equal(v, p.grad), name
    red.uninstall()


@pytest.mark.parametrize("get_softmax", [True, False])
def test_emulated_js_div_matches_reference_formula(emulated, get_softmax):
    """XModules.js_div mirror (autograd node over the fused kernel's arithmetic) == the reference formula with
    torch autograd (oracle restatement of XModules.py:32-41), value and both gradients."""
    from d2r_b200.interaction.XModules import js_div
    g = torch.Generator().manual_seed(4)
    a = torch.randn(7, 7, generator=g) * 3
    b = torch.randn(7, 7, generator=g) * 3
    if not get_softmax:
        a, b = torch.softmax(a, -1), torch.softmax(b, -1)
    p1, q1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    p2, q2 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    l1 = js_div(p1, q1, get_softmax)
    l2 = O.js_div(p2, q2, get_softmax)
    (3.0 * l1).backward()
    (3.0 * l2).backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * max(1.0, abs(l2.item()))
    assert relerr(p1.grad, p2.grad) < 1e-5 and relerr(q1.grad, q2.grad) < 1e-5


def test_emulated_block_fusion_matches_reference_golden(emulated):
    """XModules.Block mirror vs the fixture generated by the unmodified reference class
    (tests/golden/make_block_golden.py): output, input gradients, every parameter gradient digest, and the
    seeded default initialisation."""
    from d2r_b200.interaction.XModules import Block
    from tests.golden.make_block_golden import BLOCK_PARAM_SEED, block_inputs
    gold = np.load(os.path.join(GOLD, "block_fusion.npz"))
    torch.manual_seed(2023)
    m = Block([768, 768], 768)
    assert [(n, tuple(p.shape)) for n, p in m.named_parameters()] == O.block_param_spec()
    np.testing.assert_allclose([float(p.detach().double().sum()) for p in m.parameters()], gold["init_sum"],
                               rtol=0, atol=1e-9)
    m.load_state_dict(O.make_block_params(BLOCK_PARAM_SEED))
    x0, x1 = block_inputs()
    x0.requires_grad_(True)
    x1.requires_grad_(True)
    out = m([x0, x1])
    (out * torch.from_numpy(gold["w"])).sum().backward()
    assert relerr(out, gold["out"]) < 1e-5
    assert relerr(x0.grad, gold["d_x0"]) < 1e-4 and relerr(x1.grad, gold["d_x1"]) < 1e-4
    for n, p in m.named_parameters():
        got, ref = digest(p.grad), gold["gd/" + n]
        assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6), n
    with pytest.raises(ValueError):
        Block([768, 768], 768, shared=True)


@pytest.mark.parametrize("rev", [False, True], ids=["text", "image"])
@pytest.mark.parametrize("bf16", [False, True], ids=["fp32", "bf16"])
def test_emulated_more_than_128_keys(emulated, rev, bf16):
    """Sequences past the fused attention kernel's 128-key limit (BASELINE config 4 has 256 + 197 tokens): the stack
    must take the composed GEMM path for the attentions whose keys exceed the limit and the fused entry points for the
    others, with the padded leading dimensions of P -- host logic only here, the kernels run in
    tests/test_parity_gpu.py::test_config4_long_sequences_vs_oracle."""
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule
    import d2r_b200.kernels as K
    B, Lt, Li, R = 2, 133, 21, 3
    P = O.make_params(23, R, 6)
    text, image = O.make_inputs(29, B, Lt, Li)
    # The arbiter is the oracle in FLOAT64.  On this very input the fp32 oracle itself is 1.2e-2 (max-norm; 1.7e-3 in L2)
    # away from its float64 run in d_text while its forward agrees to 5e-6 -- the chain of near-argmax softmaxes
    # amplifies fp32 rounding in the backward -- whereas the product's fp32 path stays within 4e-5.
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
    t, i = text.double().requires_grad_(True), image.double().requires_grad_(True)
    ref_out, ref_sim, ref_probs = O.stack_forward(P64, t, i, R, 6, rev, training=True, bn_updates={})
    (ref_out[0].sum() + ref_sim.sum()).backward()
    m = (Reversed_InteractionModule if rev else InteractionModule)(make_args(), R, 6, 128)
    m.load_state_dict(P)
    calls = {"fused": 0}
    fused = K.attn_fused_fwd

    def counting(*a, **kw):
        calls["fused"] += 1
        return fused(*a, **kw)
    K.attn_fused_fwd = counting
    try:
        dt = torch.bfloat16 if bf16 else torch.float32
        t2, i2 = text.clone().to(dt).requires_grad_(True), image.clone().to(dt).requires_grad_(True)
        out, sim, probs = m(t2, i2, return_path_probs=True)
        (out[0].sum() + sim.sum()).backward()
    finally:
        K.attn_fused_fwd = fused
    # fused kernel: bf16 only, and only where the keys fit (here the 21 image tokens as keys, never the 133 text tokens)
    assert (calls["fused"] > 0) == bf16
    tol_p, tol_o, tol_g = (2e-2, 6e-2, 0.25) if bf16 else (1e-5, 1e-5, 2e-4)
    for a, b in zip(probs, ref_probs):
        assert relerr(a, b.detach()) < tol_p
    assert relerr(out[0], ref_out[0].detach()) < tol_o and relerr(sim, ref_sim.detach()) < 2 * tol_p
    # gradients: max-norm in fp32; in bf16 this input is the chaotic kind (see above), judged in L2 like the GPU test
    err = (lambda a, b: ((a.double() - b).norm() / b.norm()).item()) if bf16 else relerr
    assert err(t2.grad, t.grad) < tol_g and err(i2.grad, i.grad) < tol_g, (err(t2.grad, t.grad), err(i2.grad, i.grad))


def test_emulated_module_copies_and_pickles(emulated):
    """copy.deepcopy / torch.save of a mirrored module that has already run (train.py keeps state_dicts, but users of
    the reference also deep-copy models for EMA / best-checkpoint copies): the staged bf16 weights are derived data,
    not part of either; the copy is independent and gives the same result."""
    import copy
    import io
    from d2r_b200.interaction import InteractionModule
    P = O.make_params(23, 3, 6)
    m = InteractionModule(make_args(), 3, 6, 128)
    m.load_state_dict(P)
    m.eval()
    text, image = O.make_inputs(3, 2, 9, 5)
    text, image = text.bfloat16(), image.bfloat16()
    with torch.no_grad():
        o1, _ = m(text, image)
        m2 = copy.deepcopy(m)
        assert not m2.__dict__["_d2r_stager"]._cache
        o2, _ = m2(text, image)
        assert torch.equal(o1[0], o2[0])
        for p in m2.parameters():
            p.mul_(1.5)                                   # version bump -> the copy re-stages, the original does not
        o3, _ = m2(text, image)
        o4, _ = m(text, image)
        assert not torch.equal(o3[0], o4[0]) and torch.equal(o4[0], o1[0])
        buf = io.BytesIO()
        torch.save(m, buf)
        assert len(buf.getvalue()) < 1.02 * sum(p.numel() * 4 for p in m.parameters()) + (1 << 20)
        buf.seek(0)
        m3 = torch.load(buf, weights_only=False)
        o5, _ = m3(text, image)
        assert torch.equal(o5[0], o1[0])
        assert not [k for k in m.state_dict() if "_d2r" in k]


def test_emulated_inplace_modification_between_forward_and_backward_is_detected(emulated):
    """The saved state aliases inputs and parameters (no ctx.save_for_backward): the version counters are checked by
    hand when the backward starts, with autograd's own wording."""
    from d2r_b200.interaction import InteractionModule
    m = InteractionModule(make_args(), 3, 6, 128)
    m.load_state_dict(O.make_params(23, 3, 6))
    text, image = O.make_inputs(3, 2, 9, 5)

    def forward():
        t, i = (text * 1.0).requires_grad_(True), (image * 1.0).requires_grad_(True)
        t2, i2 = t * 1.0, i * 1.0                     # non-leaf inputs, as the encoders' outputs are
        out, sim = m(t2, i2)
        return t2, i2, out[0].sum() + sim.sum()

    t2, i2, loss = forward()
    with torch.no_grad():
        i2.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        loss.backward()
    t2, i2, loss = forward()
    with torch.no_grad():
        m.dynamic_itr_l0.imrc.sa.feed_forward_layer.fc1.weight.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        loss.backward()
    t2, i2, loss = forward()
    loss.backward()                                  # untouched: fine
