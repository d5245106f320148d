"""Loader for the UNMODIFIED reference modules (test infrastructure; build container only).

`/root/reference` exists only in the build container, never on the GPU box.  `import open_clip` fails there (the
package __init__ pulls tokenizer.py, which needs `ftfy`), but the three hot-path modules import cleanly when the package
__init__ is bypassed.  Used by oracle/make_golden.py to pin the oracle and to generate the
committed golden vectors under tests/golden/.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_PKG_DIR = "/root/reference/src/convert_upload/open_clip"


def available() -> bool:
    return os.path.isdir(REF_PKG_DIR)


def load():
    """Returns (transformer, model, loss) modules of the reference's vendored open_clip."""
    if not available():
        raise RuntimeError(f"{REF_PKG_DIR} not present (the reference exists only in the build container)")
    if "open_clip" not in sys.modules or not hasattr(sys.modules["open_clip"], "__path__"):
        pkg = types.ModuleType("open_clip")
        pkg.__path__ = [REF_PKG_DIR]
        sys.modules["open_clip"] = pkg
    tr = importlib.import_module("open_clip.transformer")
    mdl = importlib.import_module("open_clip.model")
    loss = importlib.import_module("open_clip.loss")
    return tr, mdl, loss
