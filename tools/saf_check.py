"""Bring-up check of the attention-filtration kernels against a torch fp64 reference over a sweep of L."""
import sys, os
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d2r_b200 import kernels as K

def run(B, Ln, D, dtype, training):
    torch.manual_seed(Ln * 10 + B)
    sg = torch.randn(B, D, device="cuda").to(dtype)
    sl = torch.randn(B, Ln, D, device="cuda").to(dtype)
    w = torch.randn(D, device="cuda") / 10
    bias = torch.tensor([0.1], device="cuda")
    bn_w, bn_b = torch.tensor([1.3], device="cuda"), torch.tensor([-0.2], device="cuda")
    rm, rv = torch.tensor([0.05], device="cuda"), torch.tensor([0.8], device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    rm0, rv0 = rm.clone(), rv.clone()
    out, saved = K.saf_fwd(sg, sl, w, bias, bn_w, bn_b, rm, rv, nbt, training)
    S = torch.cat([sg.unsqueeze(1), sl], 1).double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    gw, gb = bn_w.double().requires_grad_(True), bn_b.double().requires_grad_(True)
    logit = (S @ wd + bd).unsqueeze(1)
    y = F.batch_norm(logit, rm0.double().clone(), rv0.double().clone(), gw, gb, training, 0.1, 1e-5)
    a = torch.sigmoid(y)
    a = a / (a.abs().sum(-1, keepdim=True) + 1e-8)
    saf = (a @ S).squeeze(1)
    ref = saf / (saf.pow(2).sum(-1, keepdim=True).sqrt() + 1e-8)
    d_out = torch.randn(B, D, device="cuda")
    ref.backward(d_out.double())
    d_sg, d_sl, d_w, d_bias, d_bn_w, d_bn_b = K.saf_bwd(d_out, sg, sl, w, bias, bn_w, bn_b, rm, rv, training, saved)
    rel = lambda x, r: ((x.double() - r).abs().max() / (r.abs().max() + 1e-30)).item()
    return rel(out, ref), rel(d_sg, S.grad[:, 0]), rel(d_sl, S.grad[:, 1:]), rel(d_w, wd.grad)

for training in (False, True):
    for Ln in list(range(1, 41)) + [50, 128]:
        for B in (3, 6):
            e = run(B, Ln, 768, torch.float32, training)
            flag = "  <-- BAD" if max(e) > 1e-4 else ""
            if flag or Ln in (16, 21):
                print(f"train={training} B={B} L={Ln}: out {e[0]:.2e} d_sg {e[1]:.2e} d_sl {e[2]:.2e} d_w {e[3]:.2e}{flag}")
print("done")
