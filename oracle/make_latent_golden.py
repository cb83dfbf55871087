"""Generates tests/golden/collate_control.pt by running the REFERENCE's own `CollateFunctionControl`
(orv/dataset/dataset.py:2053-2126) on seeded items.

    python -m oracle.make_latent_golden            (build container only: needs /root/reference)

`orv.dataset.dataset` cannot be imported here (decord, omegaconf, cv2 are absent), so the class statement is located
in the file with `ast` and executed as it stands — nothing of it is copied into this repository."""
from __future__ import annotations

import ast
import os
import sys
from typing import Any

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import latent_oracle as LO  # noqa: E402

REF_FILE = os.path.join(os.environ.get("ORV_REFERENCE_ROOT", "/root/reference"), "orv", "dataset", "dataset.py")


def reference_collate_class():
    src = open(REF_FILE, encoding="utf-8").read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CollateFunctionControl")
    ns = {"torch": torch, "Any": Any}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF_FILE, "exec"), ns)  # noqa: S102 — the reference's own code
    return ns["CollateFunctionControl"]


def main():
    cls = reference_collate_class()
    out = {}
    for dt_name, dt in (("float32", torch.float32), ("bfloat16", torch.bfloat16)):
        ret = cls(weight_dtype=dt, load_tensors=True)(LO.golden_items())
        for k in ("prompts", "metainfos", "num_views", "num_frames"):  # non-tensor bookkeeping outside this path
            ret.pop(k, None)
        out[dt_name] = ret
    path = os.path.join(ROOT, "tests", "golden", "collate_control.pt")
    torch.save(out, path)
    print(path, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out["float32"].items() if k != "controls"},
          {k: tuple(v.shape) for k, v in out["float32"]["controls"].items()})


if __name__ == "__main__":
    main()
