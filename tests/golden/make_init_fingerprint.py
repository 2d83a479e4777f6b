"""Record the UNMODIFIED reference's state_dict layout and seeded default initialisation
(run where /root/reference exists): tests/golden/init_fingerprint.json.  The drop-in mirror must create
the same keys, shapes and -- under the same torch.manual_seed -- the same initial values."""
import json
import os
import sys
import tempfile

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
sys.dont_write_bytecode = True
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("D2R_REFERENCE", "/root/reference"))
from tests.golden.make_golden import ref_args  # noqa: E402


def main():
    from models.InteractionModule import InteractionModule, Reversed_InteractionModule
    args = ref_args(tempfile.mkdtemp())
    rec = {}
    for name, cls in (("text", InteractionModule), ("image", Reversed_InteractionModule)):
        torch.manual_seed(2023)
        m = cls(args, num_layer_routing=3, num_cells=6, path_hid=128)
        sd = m.state_dict()
        rec[name] = {"keys": list(sd.keys()), "shapes": [list(v.shape) for v in sd.values()],
                     "sum": [float(v.double().sum()) for v in sd.values()],
                     "abssum": [float(v.double().abs().sum()) for v in sd.values()],
                     "params": [n for n, _ in m.named_parameters()]}
    json.dump(rec, open(os.path.join(HERE, "init_fingerprint.json"), "w"))
    print("wrote", len(rec["text"]["keys"]), "entries per branch")


if __name__ == "__main__":
    main()
