"""Import the UNMODIFIED reference (SorF520/D2R) from the staged copy ``baseline/_ref/`` -- or from
``/root/reference`` in the authoring container -- with harness-side shims only (no reference file
is edited):

  1. ``sys.path`` gets the reference root (its modules import each other as ``models.X``);
  2. ``args.bert_name`` / ``args.vit_name`` point at local directories holding a default ``config.json``
     (``Cells.py:136-139`` calls ``BertConfig.from_pretrained`` only to read ``hidden_size``); no download;
  3. full model only: ``transformers.modeling_utils.apply_chunking_to_forward`` alias (moved to
     ``transformers.pytorch_utils`` in the installed transformers; ``modeling_unimo.py:8-10`` imports the old path);
  4. CPU runs hide the GPU (``CUDA_VISIBLE_DEVICES=""`` in a subprocess): the reference's dead-code
     ``ContrastiveLoss`` moves a mask to CUDA whenever CUDA is available (``XModules.py:223-227,237-241``).

``stage()`` copies the reference sources to ``baseline/_ref/`` (called by ``__graft_entry__.build()`` where
``/root/reference`` exists) so that they travel to the GPU box with the repo snapshot; the directory is
git-ignored -- reference sources never enter the repo's history.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
UPSTREAM = os.environ.get("D2R_REFERENCE", "/root/reference")


def stage(force: bool = False) -> str | None:
    """Copy the reference's Python sources (1 MB; the unused 39k-line SenticNet word list and stale .pyc files
    are left out) into baseline/_ref/.  Returns the staged path, or None when no upstream copy is reachable."""
    if not os.path.isdir(os.path.join(UPSTREAM, "models")):
        return STAGED if os.path.isdir(os.path.join(STAGED, "models")) else None
    if os.path.isdir(os.path.join(STAGED, "models")) and not force:
        return STAGED
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    shutil.copytree(UPSTREAM, STAGED, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "senticnet_word.txt",
                                                                      ".git"))
    for root, dirs, files in os.walk(STAGED):        # the upstream tree is read-only; the copy need not be
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    return STAGED


def root() -> str | None:
    if os.path.isdir(os.path.join(STAGED, "models")):
        return STAGED
    if os.path.isdir(os.path.join(UPSTREAM, "models")):
        return UPSTREAM
    return None


def available() -> bool:
    return root() is not None


_CFG_DIR = None


def config_dirs():
    """Local BertConfig / CLIPConfig directories (defaults: hidden 768, ViT-B/32 -> 50 image tokens)."""
    global _CFG_DIR
    if _CFG_DIR is None:
        from transformers import BertConfig, CLIPConfig
        _CFG_DIR = tempfile.mkdtemp(prefix="d2r_refcfg_")
        BertConfig().save_pretrained(os.path.join(_CFG_DIR, "bert"))
        CLIPConfig().save_pretrained(os.path.join(_CFG_DIR, "clip"))
    return os.path.join(_CFG_DIR, "bert"), os.path.join(_CFG_DIR, "clip")


def ref_args(**extra):
    bd, vd = config_dirs()
    ns = argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                            raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                            bert_name=bd, vit_name=vd, DR_step=3, weight_js_1=1.0, weight_js_2=1.0)
    for k, v in extra.items():
        setattr(ns, k, v)
    return ns


def import_reference(full_model: bool = False):
    """-> the reference's ``models`` package (``models.InteractionModule`` etc. imported)."""
    r = root()
    if r is None:
        raise ImportError("the reference is not staged (baseline/_ref/ missing and /root/reference absent)")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    sys.dont_write_bytecode = True
    if r not in sys.path:
        sys.path.insert(0, r)
    if full_model:
        import transformers.modeling_utils as mu
        import transformers.pytorch_utils as pu
        if not hasattr(mu, "apply_chunking_to_forward"):
            mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    import importlib
    mods = importlib.import_module("models.InteractionModule")
    if not os.path.abspath(mods.__file__).startswith(os.path.abspath(r)):
        raise ImportError(f"'models' resolved to {mods.__file__}, not to the reference under {r}")
    return importlib.import_module("models")
