"""CPU-side checks of the drop-in boundary: class names, constructor signatures, state_dict
keys / shapes / order and seeded initial values identical to the unmodified reference (fingerprint recorded by
tests/golden/make_init_fingerprint.py), error behaviour without CUDA, and the data-parallel plan."""
import argparse
import inspect
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_args():
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")


@pytest.fixture(scope="module", autouse=True)
def built():
    from d2r_b200 import build
    build.build()


@pytest.mark.parametrize("branch", ["text", "image"])
def test_state_dict_and_seeded_init_match_reference(branch):
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule
    fp = json.load(open(os.path.join(GOLD, "init_fingerprint.json")))[branch]
    torch.manual_seed(2023)
    m = (InteractionModule if branch == "text" else Reversed_InteractionModule)(make_args(), 3, 6, 128)
    sd = m.state_dict()
    assert list(sd.keys()) == fp["keys"]
    assert [list(v.shape) for v in sd.values()] == fp["shapes"]
    assert [n for n, _ in m.named_parameters()] == fp["params"]
    for k, v, s, a in zip(sd.keys(), sd.values(), fp["sum"], fp["abssum"]):
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, a), k
        assert abs(float(v.double().abs().sum()) - a) <= 1e-9 * max(1.0, a), k


def test_signatures_match_reference():
    import importlib
    from d2r_b200.interaction import Cells, DynamicInteraction, Refinement, Router, SelfAttention, XModules
    # the package re-exports the InteractionModule CLASS under the submodule's name, as `from models.InteractionModule
    # import InteractionModule` does in the reference; fetch the submodule itself explicitly
    InteractionModule = importlib.import_module("d2r_b200.interaction.InteractionModule")
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(InteractionModule.InteractionModule.__init__) == ["self", "args", "num_layer_routing", "num_cells", "path_hid"]
    assert sig(InteractionModule.InteractionModule.forward)[:3] == ["self", "text", "image"]
    assert sig(InteractionModule.Reversed_InteractionModule.forward)[:3] == ["self", "text", "image"]
    for cls in ("DynamicInteraction_Layer0", "DynamicInteraction_Layer", "Reversed_DynamicInteraction_Layer0",
                "Reversed_DynamicInteraction_Layer"):
        assert sig(getattr(DynamicInteraction, cls).__init__) == ["self", "args", "num_cell", "num_out_path"]
    assert sig(DynamicInteraction.DynamicInteraction_Layer0.forward) == ["self", "text", "image"]
    assert sig(DynamicInteraction.DynamicInteraction_Layer.forward) == ["self", "ref_wrd", "text", "image"]
    for cls in ("RectifiedIdentityCell", "IntraModelReasoningCell", "CrossModalRefinementCell",
                "GlobalLocalAlignmentCell", "GlobalEnhancedSemanticCell", "ContextRichCrossModalCell"):
        assert sig(getattr(Cells, cls).__init__) == ["self", "args", "num_out_path"]
    assert sig(Router.Router.__init__) == ["self", "num_out_path", "embed_size", "hid"]
    assert sig(SelfAttention.SelfAttention.__init__) == ["self", "embed_size", "hid_size", "h", "drop"]
    assert sig(Refinement.Refinement.__init__) == ["self", "args", "embed_size", "raw_feature_norm", "lambda_softmax"]
    assert sig(XModules.CrossModalAlignment.__init__) == ["self", "config", "args"]
    assert sig(XModules.AttentionFiltration.__init__) == ["self", "sim_dim"]


def test_envelope_errors():
    from d2r_b200.interaction import InteractionModule
    with pytest.raises(ValueError):
        InteractionModule(make_args(), 3, 5, 128)        # only 4 or 6 cells exist
    with pytest.raises(ValueError):
        InteractionModule(make_args(), 1, 6, 128)
    m = InteractionModule(make_args(), 2, 4, 128)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(2, 4, 768), torch.randn(2, 3, 768))
    # malformed inputs: RuntimeError like the reference (which fails inside its first matmul), raised at the boundary
    from d2r_b200.interaction import Reversed_InteractionModule, run_pair
    r = Reversed_InteractionModule(make_args(), 2, 4, 128)
    for text, image, what in ((torch.randn(2, 4, 512), torch.randn(2, 3, 512), "text input must be"),
                              (torch.randn(2, 4, 768), torch.randn(3, 3, 768), "batch sizes"),
                              (torch.randn(4, 768), torch.randn(3, 768), "text input must be"),
                              (torch.randn(2, 4, 768), torch.randn(2, 0, 768), "image input has no tokens")):
        for call in (lambda: m(text, image), lambda: r(text, image), lambda: run_pair(m, r, text, image)):
            with pytest.raises(RuntimeError, match=what):
                call()


def test_dead_parameter_plan_matches_reference_fixture():
    import numpy as np
    from d2r_b200 import dp
    gold = np.load(os.path.join(GOLD, "text_r3_train.npz"))
    dead = set(gold["dead"].tolist())
    fp = json.load(open(os.path.join(GOLD, "init_fingerprint.json")))["text"]
    for n in fp["params"]:
        assert dp.is_live(n) == (n not in dead), n


def test_shard_batch():
    from d2r_b200.dp import shard_batch
    for n, w in [(256, 8), (10, 4), (7, 2), (3, 8)]:
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _run_world2(worker, attempts=3):
    """Spawn two gloo ranks of ``worker(rank, world, port, queue)`` on 127.0.0.1 -> their results sorted by rank.  The
    rendezvous port is OS-assigned; a failed rendezvous (port taken in the meantime, a loaded host) is retried."""
    import queue as _queue
    import torch.multiprocessing as mp
    last = None
    for _ in range(attempts):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
        for p in procs:
            p.start()
        try:
            res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"worker exit codes {[p.exitcode for p in procs]}")
        except _queue.Empty as e:
            last = e
        for p in procs:
            if p.is_alive():
                p.terminate()
            p.join(timeout=10)
    raise AssertionError(f"world-2 gloo run failed {attempts} times: {last!r}")


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from d2r_b200.dp import GradAllReducer, shard_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)

    class Tiny(torch.nn.Module):           # names mimic live and dead stack parameters
        def __init__(self):
            super().__init__()
            self.live = torch.nn.Linear(6, 3)
            self.path_mapping = torch.nn.Linear(3, 2)     # dead: never used in forward

        def forward(self, x):
            return self.live(x)

    m = Tiny()
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lo, hi = shard_batch(8, rank, world)
    m(x[lo:hi]).pow(2).mean().backward()
    red = GradAllReducer([m])
    assert red.names == ["0.live.weight", "0.live.bias"]
    red.step()
    q.put((rank, m.live.weight.grad.clone(), m.live.bias.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_allreduce_gloo_world2():
    """N-rank DP == mean of the per-shard gradients (world_size 2, gloo, CPU)."""
    res = _run_world2(_dp_worker)
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    gs = []
    for lo, hi in ((0, 4), (4, 8)):
        lin.zero_grad()
        lin(x[lo:hi]).pow(2).mean().backward()
        gs.append(lin.weight.grad.clone())
    torch.testing.assert_close(res[0][1], (gs[0] + gs[1]) / 2)


def _dp_layerwise_worker(rank, world, port, q):
    """Layer-wise reducer driven the way stack_backward drives it: last routing layer first, one callback per
    layer with that layer's finished gradients; compared with the one-collective path on the same gradients."""
    import torch.distributed as dist
    from d2r_b200.dp import GradAllReducer, LAYER_HOOK_ATTR
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(4, 4)

    class TinyStack(torch.nn.Module):      # parameter names of a routed stack: l0, l1.<i>, l2 + dead ones
        def __init__(self):
            super().__init__()
            self.dynamic_itr_l0 = Layer()
            self.dynamic_itr_l1 = torch.nn.ModuleList([Layer(), Layer()])
            self.dynamic_itr_l2 = Layer()
            self.path_mapping = torch.nn.Linear(3, 2)

    torch.manual_seed(1)
    mods = [TinyStack(), TinyStack()]
    g = torch.Generator().manual_seed(100 + rank)
    grads = [{n: torch.randn(p.shape, generator=g) for n, p in m.named_parameters()} for m in mods]
    red = GradAllReducer(mods)
    assert [k for k in red.slices] == [(0, "dynamic_itr_l2"), (0, "dynamic_itr_l1.1"), (0, "dynamic_itr_l1.0"),
                                       (0, "dynamic_itr_l0"), (1, "dynamic_itr_l2"), (1, "dynamic_itr_l1.1"),
                                       (1, "dynamic_itr_l1.0"), (1, "dynamic_itr_l0")]
    red.install()
    for mi, m in enumerate(mods):
        hook = m.__dict__[LAYER_HOOK_ATTR]
        for layer in ("dynamic_itr_l2", "dynamic_itr_l1.1", "dynamic_itr_l1.0", "dynamic_itr_l0"):
            hook(layer, grads[mi])
    # autograd leaves the LOCAL gradients in p.grad during backward; wait() must re-point them at the reduced bucket
    for mi, m in enumerate(mods):
        for n, p in m.named_parameters():
            p.grad = grads[mi][n].clone()
    red.wait()
    layerwise = red.flat.clone()
    lo, hi = red.flat.data_ptr(), red.flat.data_ptr() + red.flat.numel() * 4
    for p, v in zip(red.params, red._views):
        assert lo <= p.grad.data_ptr() < hi and torch.equal(p.grad, v), "p.grad is not the reduced view after wait()"
    red.finish()                          # a stray finish() after wait() must not divide a second time
    assert torch.equal(red.flat, layerwise)
    # reference: the one-collective path on the same per-rank gradients
    for mi, m in enumerate(mods):
        for n, p in m.named_parameters():
            p.grad = grads[mi][n].clone()
    red2 = GradAllReducer(mods)
    red2.step()
    assert torch.allclose(layerwise, red2.flat, rtol=0, atol=1e-7)
    # an incomplete backward must be reported, not silently reduced
    red.on_layer(0, "dynamic_itr_l2", grads[0])
    try:
        red.wait()
        ok = False
    except RuntimeError:
        ok = True
    for w in red._works:               # (drain the collective of the deliberately incomplete round)
        w.wait()
    q.put((rank, layerwise.tolist(), ok))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_layerwise_allreduce_gloo_world2():
    res = _run_world2(_dp_layerwise_worker)
    assert res[0][1] == res[1][1]                     # both ranks hold the same averaged bucket
    assert res[0][2] and res[1][2]


def test_lanes_degrade_to_one_stream_off_cuda():
    """lanes.fork on a CPU device (the emulated tests) or with n <= 1 or ENABLED = False yields the no-op object
    with the full interface."""
    import d2r_b200.lanes as LN
    for lanes in (LN.fork(torch.device("cpu"), 5, "cells"), LN.fork(torch.device("cpu"), 1)):
        assert isinstance(lanes, LN.NoLanes)
        with lanes.lane(3):
            pass
        lanes.catch_up(2)
        lanes.wait_mark(lanes.mark(4))
        lanes.join()


def _dp_bucketed_worker(rank, world, port, q):
    """BucketedGradReducer on a small model with a never-used parameter: p.grad become views of one flat buffer, the
    buckets are reduced from inside the backward, every rank ends with the mean of the per-rank gradients."""
    import torch.distributed as dist
    from d2r_b200.dp import BucketedGradReducer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                torch.nn.Linear(16, 3))
    unused = torch.nn.Parameter(torch.zeros(5))
    params = list(model.parameters()) + [unused]
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lo, hi = rank * 4, rank * 4 + 4
    model(x[lo:hi]).pow(2).mean().backward()
    red = BucketedGradReducer(params, bucket_mb=16 * 16 * 4 / (1 << 20))      # several buckets
    red.plan()
    assert red.dead == 1 and len(red.buckets) >= 2
    for _ in range(2):                                                          # two steps: buffers are reused
        red.begin_step()
        model(x[lo:hi]).pow(2).mean().backward()
        red.finish()
    lo_ptr, hi_ptr = red.flat.data_ptr(), red.flat.data_ptr() + red.flat.numel() * 4
    assert all(lo_ptr <= p.grad.data_ptr() < hi_ptr for p in model.parameters()) and unused.grad is None
    q.put((rank, [p.grad.clone() for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_bucketed_overlapped_allreduce_gloo_world2():
    res = _run_world2(_dp_bucketed_worker)
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                torch.nn.Linear(16, 3))
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    gs = []
    for lo, hi in ((0, 4), (4, 8)):
        model.zero_grad()
        model(x[lo:hi]).pow(2).mean().backward()
        gs.append([p.grad.clone() for p in model.parameters()])
    for a, b, g0, g1 in zip(res[0][1], res[1][1], gs[0], gs[1]):
        assert torch.equal(a, b)
        torch.testing.assert_close(a, (g0 + g1) / 2)
