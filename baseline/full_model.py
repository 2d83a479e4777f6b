"""Build the reference's FULL model (``UnimoModelF``: BERT-base + CLIP-ViT-B/32 towers feeding the routed stacks,
``models/unimo_model.py:138-162``) from the staged, unmodified reference with random-init weights (no download), plus
MVSA-shaped synthetic batches.  Harness-side only; the product never imports this."""
from __future__ import annotations

import torch

from . import ref_loader as RL


def build_reference_model(layers: int = 3, seed: int = 2023):
    """-> (UnimoModelF instance on CPU, args).  Default BertConfig / CLIPVisionConfig (hidden 768, 12 layers each,
    ViT-B/32 at 224 px -> 50 image tokens), random init under ``seed``."""
    RL.import_reference(full_model=True)
    from transformers import BertConfig, CLIPConfig
    from models.unimo_model import UnimoModelF          # the reference's own
    args = RL.ref_args(DR_step=layers, weight_js_1=1.0, weight_js_2=1.0)
    torch.manual_seed(seed)
    vision_config = CLIPConfig().vision_config
    text_config = BertConfig()
    model = UnimoModelF(args=args, vision_config=vision_config, text_config=text_config)
    return model, args


def synthetic_batch(batch: int, max_seq: int = 128, seed: int = 2023, device="cpu"):
    """input_ids in [1000, 30000) with a random-length zero-padded tail + matching mask, token_type_ids 0,
    labels in {0,1,2}, pixel_values N(0,1) [B,3,224,224]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30000, (batch, max_seq), generator=g)
    lens = torch.randint(max_seq // 4, max_seq + 1, (batch,), generator=g)
    mask = (torch.arange(max_seq).unsqueeze(0) < lens.unsqueeze(1)).long()
    ids = ids * mask
    tt = torch.zeros_like(ids)
    labels = torch.randint(0, 3, (batch,), generator=g)
    images = torch.randn(batch, 3, 224, 224, generator=g)
    return tuple(t.to(device) for t in (ids, mask, tt, labels, images))
