import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Unit tests of the kernels first, then the stack against the goldens / the oracle, then the whole reference model
# (400 M parameters, CUDA-graphed stacks): a run stopped at its first failure (-x) has then covered everything below
# the failing level.  Files not listed keep their alphabetical place in front.
ORDER = ["test_kernels_gpu.py", "test_gemm_gpu.py", "test_attn_fused_gpu.py", "test_parity_gpu.py",
         "test_parity_train_gpu.py", "test_full_model_gpu.py"]


def pytest_collection_modifyitems(config, items):
    rank = {name: i for i, name in enumerate(ORDER)}
    items.sort(key=lambda it: rank.get(os.path.basename(str(it.fspath)), -1))      # stable: order inside a file is kept
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
