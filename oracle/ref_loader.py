"""Loader for the UNMODIFIED reference modules (test / benchmark infrastructure, never the product path).

The reference is Python.  `import open_clip` fails (the package __init__ pulls tokenizer.py, which needs `ftfy`), but the
three hot-path modules import cleanly when the package __init__ is bypassed.  Two locations:

  /root/reference/src/convert_upload/open_clip   the reference itself; exists only in the build container
  baseline/_ref/open_clip                         a byte-for-byte copy of the nine files the three modules import, made by
                                                  __graft_entry__.build() (install()) while the reference is present; git-ignored
                                                  (never committed) but not gpurun-ignored, so it travels to the GPU box

Used by oracle/make_golden.py (pins the oracle, generates tests/golden/) and by bench.py (`--impl reference`, and the
reference-on-GPU comparator).
"""
from __future__ import annotations

import filecmp
import importlib
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG_DIR = "/root/reference/src/convert_upload/open_clip"
LOCAL_PKG_DIR = os.path.join(ROOT, "baseline", "_ref", "open_clip")
# what open_clip.{transformer,model,loss} import
FILES = ("transformer.py", "model.py", "loss.py", "utils.py", "pos_embed.py", "hf_model.py", "hf_configs.py",
         "modified_resnet.py", "timm_model.py")


def install() -> str:
    """Copy the reference's files (unmodified) to baseline/_ref/open_clip when the reference is present.  Returns a one-line
    record of what happened.  No __init__.py is written: the loader bypasses the package __init__ on purpose."""
    if not os.path.isdir(REF_PKG_DIR):
        return "reference absent: baseline/_ref left as it is" if os.path.isdir(LOCAL_PKG_DIR) else "reference absent, no baseline/_ref"
    os.makedirs(LOCAL_PKG_DIR, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(REF_PKG_DIR, f), os.path.join(LOCAL_PKG_DIR, f)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return f"copied {len(FILES)} unmodified files to baseline/_ref/open_clip"


def package_dir():
    for d in (REF_PKG_DIR, LOCAL_PKG_DIR):
        if os.path.isfile(os.path.join(d, "transformer.py")):
            return d
    return None


def available() -> bool:
    return package_dir() is not None


def load():
    """Returns (transformer, model, loss) modules of the reference's vendored open_clip."""
    d = package_dir()
    if d is None:
        raise RuntimeError(f"neither {REF_PKG_DIR} nor {LOCAL_PKG_DIR} holds the reference modules "
                           "(run __graft_entry__.build() in the build container)")
    if "open_clip" not in sys.modules or not hasattr(sys.modules["open_clip"], "__path__"):
        pkg = types.ModuleType("open_clip")
        pkg.__path__ = [d]
        sys.modules["open_clip"] = pkg
    tr = importlib.import_module("open_clip.transformer")
    mdl = importlib.import_module("open_clip.model")
    loss = importlib.import_module("open_clip.loss")
    return tr, mdl, loss
