"""Mirror of reference models/Router.py (Router :10-26, activateFunc :6-8)."""
import torch
import torch.nn as nn

from .. import stack as S
from ..autograd import run_block
from .. import kernels as K


def activateFunc(x):
    """relu(tanh(x)) -- models/Router.py:6-8 (kept for API parity; tiny torch op, not on the hot path)."""
    return torch.relu(torch.tanh(x))


def _router_fwd(env, xs):
    pooled = K.pool_mean([xs[0]])
    norm, gate, sv = S._routers_fwd(env, ["R"], pooled, env.P["R.mlp.2.weight"].shape[0], False)
    sv["x_shape"] = xs[0].shape
    return (sv["raw"].view(sv["raw"].shape[0], -1),), sv


def _router_bwd(env, sv, grads):
    B, Ln, D = sv["x_shape"]
    d_raw = grads[0].reshape(B, -1, 1)
    d_pooled = S._routers_bwd(env, ["R"], sv, d_raw, True)      # 'final' = un-normalised probabilities
    return (K.pool_mean_bwd(d_pooled[0], Ln, env.cd),)


class Router(nn.Module):
    def __init__(self, num_out_path, embed_size, hid):
        super(Router, self).__init__()
        self.num_out_path = num_out_path
        self.mlp = nn.Sequential(nn.Linear(embed_size, hid), nn.ReLU(True), nn.Linear(hid, num_out_path))
        self.init_weights()

    def init_weights(self):
        self.mlp[2].bias.data.fill_(1.5)

    def forward(self, x):    # (bsz, L, D) -> (bsz, num_out_path)
        return run_block(self, [x], _router_fwd, _router_bwd, prefix="R.")[0]
