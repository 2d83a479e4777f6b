"""One shape of the fused-softmax score GEMM, a few launches (for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d2r_b200 import kernels as K  # noqa: E402
from d2r_b200 import _lib as L  # noqa: E402

B, H, D, Lq = 256, 16, 768, 128
dh = D // H
qkv = torch.randn(B, Lq, 3 * D, device="cuda").bfloat16()
P = torch.empty(B, H, Lq, Lq, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    K.gemm(qkv, qkv[:, :, D:], P, m=Lq, n=Lq, k=dh, lda=3 * D, ldb=3 * D, ldc=Lq, batch=B * H, batch_inner=H,
           a_str=(Lq * 3 * D, dh), b_str=(Lq * 3 * D, dh), c_str=(H * Lq * Lq, Lq * Lq), alpha=0.14,
           epilogue=L.EPI_SOFTMAX)
torch.cuda.synchronize()
print("ok")
