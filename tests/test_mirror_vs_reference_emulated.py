"""Every mirrored class of d2r_b200/interaction against the UNMODIFIED reference class of the same name (imported
from the staged baseline/_ref), same state_dict, same inputs, train and eval mode: outputs, input gradients, every
parameter gradient, which parameters get no gradient at all, and the buffers (BatchNorm running statistics).  The
kernels are the torch-CPU emulation of tests/emu_kernels.py: this pins the stand-alone use of the classes SURVEY
section 8b says must stay importable (cells, router, layers, SelfAttention, Refinement) -- the GPU kernels behind the
same call sites are checked by the -m gpu tests.  Skipped when the reference is not staged."""
import pytest
import torch

from oracle import d2r_oracle as O
from tests import emu_kernels as E


@pytest.fixture(scope="module")
def ref():
    from baseline import ref_loader as RL
    if not RL.available():
        pytest.skip("reference not staged (baseline/_ref)")
    return RL.import_reference(), RL.ref_args()


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


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _flatten(v, out):
    if isinstance(v, torch.Tensor):
        out.append(v)
    elif isinstance(v, (list, tuple)):
        for u in v:
            _flatten(u, out)
    return out


def _run(mod, inputs, train):
    mod.train(train)
    xs = [[x.clone().requires_grad_(True) for x in inp] if isinstance(inp, list) else inp.clone().requires_grad_(True)
          for inp in inputs]
    for p in mod.parameters():
        p.grad = None
    outs = _flatten(mod(*xs), [])
    g = torch.Generator().manual_seed(1)
    sum((o * torch.randn(o.shape, generator=g)).sum() for o in outs if o.requires_grad).backward()
    return (outs, [x.grad for x in _flatten(xs, [])], {n: p.grad for n, p in mod.named_parameters()},
            {n: b.clone() for n, b in mod.named_buffers()})


def _compare(rm, mm, inputs, train):
    assert list(mm.state_dict().keys()) == list(rm.state_dict().keys())
    mm.load_state_dict(rm.state_dict())
    (fr, gr, pr, br), (fm, gm, pm, bm) = _run(rm, inputs, train), _run(mm, inputs, train)
    assert len(fr) == len(fm)
    for a, b in zip(fm, fr):
        assert a.shape == b.shape and relerr(a.detach(), b.detach()) <= 1e-5
    for a, b in zip(gm, gr):
        assert (a is None) == (b is None)
        if b is not None:
            assert relerr(a, b) <= 1e-4
    scale = max(g.abs().max().item() for g in pr.values() if g is not None)
    for n, g in pr.items():
        assert (g is None) == (pm[n] is None), n                      # same never-used parameters
        if g is not None and g.abs().max().item() > 1e-6 * scale:     # (mathematically-zero gradients: noise on both sides)
            assert relerr(pm[n], g) <= 1e-3, n
    for n, b in br.items():
        assert relerr(bm[n].float(), b.float()) <= 1e-5, n


CELLS = ["RectifiedIdentityCell", "IntraModelReasoningCell", "CrossModalRefinementCell", "GlobalLocalAlignmentCell",
         "GlobalEnhancedSemanticCell", "ContextRichCrossModalCell"]


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
def test_mirrored_classes_match_the_reference_classes(ref, emulated, train):
    import d2r_b200.interaction as M
    from tests.test_stack_emulated import make_args
    R, args = ref
    margs = make_args()
    text, image = O.make_inputs(5, 3, 9, 6)
    torch.manual_seed(0)
    _compare(R.Router.Router(6, 768, 768), M.Router.Router(6, 768, 768), [text], train)
    _compare(R.SelfAttention.SelfAttention(768, 768, 16, 0.0), M.SelfAttention.SelfAttention(768, 768, 16, 0.0),
             [text], train)
    _compare(R.Refinement.Refinement(args, 768, "clipped_l2norm", 4.0),
             M.Refinement.Refinement(margs, 768, "clipped_l2norm", 4.0), [text, image], train)
    for name in CELLS:
        ins = [text] if name in CELLS[:2] else [text, image]
        _compare(getattr(R.Cells, name)(args, 6), getattr(M.Cells, name)(margs, 6), ins, train)
    for name in ("DynamicInteraction_Layer0", "Reversed_DynamicInteraction_Layer0"):
        _compare(getattr(R.DynamicInteraction, name)(args, 6, 6), getattr(M.DynamicInteraction, name)(margs, 6, 6),
                 [text, image], train)
    # later layers: forward(ref_wrd, text, image) with the K streams of the previous layer (DynamicInteraction.py:90,
    # :210); the final one has a single output path
    g = torch.Generator().manual_seed(3)
    streams = [text + 0.1 * torch.randn(text.shape, generator=g) for _ in range(6)]
    istreams = [image + 0.1 * torch.randn(image.shape, generator=g) for _ in range(6)]
    for n_out in (6, 1):
        _compare(R.DynamicInteraction.DynamicInteraction_Layer(args, 6, n_out),
                 M.DynamicInteraction.DynamicInteraction_Layer(margs, 6, n_out), [streams, text, image], train)
        _compare(R.DynamicInteraction.Reversed_DynamicInteraction_Layer(args, 6, n_out),
                 M.DynamicInteraction.Reversed_DynamicInteraction_Layer(margs, 6, n_out), [istreams, text, image], train)
