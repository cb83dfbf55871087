"""Imports the reference's own modules (unmodified, from /root/reference) on top of oracle/shim.

Only usable in the build container (where /root/reference is mounted); never on the GPU box and never from
`orv_b200/`.  Pitfalls handled (SURVEY.md probe P9): `transformers` must be imported before a fake `accelerate`
appears on sys.path (it version-parses the real package), so `accelerate.logging` is injected via sys.modules.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ORV_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "orv", "models", "cogvideox_control.py"))


def load():
    """Returns the imported reference module `orv.models.cogvideox_control`."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    import transformers  # noqa: F401  (before the accelerate stub exists)
    from transformers.models.t5 import T5EncoderModel, T5Tokenizer  # noqa: F401
    if "accelerate" not in sys.modules or not hasattr(sys.modules.get("accelerate.logging", None), "get_logger"):
        acc = sys.modules.get("accelerate") or types.ModuleType("accelerate")
        acc_log = types.ModuleType("accelerate.logging")
        acc_log.get_logger = lambda name, *a, **k: logging.getLogger(name)
        acc.logging = acc_log
        sys.modules.setdefault("accelerate", acc)
        sys.modules["accelerate.logging"] = acc_log
    for p in (SHIM_ROOT, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("orv.models.cogvideox_control")
