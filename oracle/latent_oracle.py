"""TEST INFRASTRUCTURE — CPU restatement of the reference's latent read path, used only by tests/.

* `reference_sample`: what `RobotDataset.__getitem__` assembles for one clip from the pre-encoded files
  (orv/dataset/dataset.py:655-694 video / image latents — 3-D VAE moments are stored [C, F, h, w] and handed on as
  [F, C, h, w] —, :785-850 depth / label latents per view, stacked and flattened to [(v f), C, h, w], :1054-1059 the
  cached empty-prompt embedding).
* `reference_collate`: `CollateFunctionControl.__call__` (dataset.py:2053-2126) for the tensor keys of this path.

Pinning: `reference_collate` is checked against tests/golden/collate_control.pt, which oracle/make_latent_golden.py
produced by executing the reference's OWN `CollateFunctionControl` class (its source is read from
/root/reference/orv/dataset/dataset.py at generation time; the module itself cannot be imported here — decord /
omegaconf / cv2 are absent).  `reference_sample` restates a method of `RobotDataset`, whose constructor needs the real
dataset tree: that half is **unpinned**.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch


def synthetic_files(names: Sequence[str], C=32, F=5, h=6, w=8, S=7, E=16, seed: int = 0) -> Dict[Any, torch.Tensor]:
    """Seeded stand-ins for the files encode_dataset.py writes (:353-363: one [C, F, h, w] moments tensor per sample
    and folder; :1073-1094: empty_prompt.pt with a batch dimension)."""
    g = torch.Generator().manual_seed(seed)
    files: Dict[Any, torch.Tensor] = {}
    for folder, frames in (("video_latents", F), ("image1_latents", 1), ("depth_latents", F), ("label_latents", F),
                           ("depthGT_latents", F)):
        for n in names:
            files[(folder, n)] = torch.randn((C, frames, h, w), generator=g).to(torch.bfloat16)
    for n in names:
        files[("prompt_embeds", n)] = torch.randn((S, E), generator=g).to(torch.bfloat16)
    files["empty"] = torch.randn((1, S, E), generator=g).to(torch.bfloat16)
    return files


def reference_sample(files, n: str, views: Optional[Sequence[str]] = None, gt: bool = False) -> Dict[str, torch.Tensor]:
    # dataset.py:673-694 (3-D VAE latents are [C, F, H, W] on disk), :814-827, :836-848, :1056-1059
    out = {"prompt_embeds": files["empty"][0]}
    out["latents"] = files[("video_latents", n)].permute(1, 0, 2, 3)
    out["image"] = files[("image1_latents", n)].permute(1, 0, 2, 3)
    views = views or [n]
    d = "depthGT_latents" if gt else "depth_latents"
    out["latents_depth"] = torch.stack([files[(d, v)].permute(1, 0, 2, 3) for v in views]).flatten(0, 1)
    out["latents_label"] = torch.stack([files[("label_latents", v)].permute(1, 0, 2, 3) for v in views]).flatten(0, 1)
    return out


def reference_collate(items: List[Dict[str, torch.Tensor]], dtype: torch.dtype) -> Dict[str, Any]:
    # dataset.py:2076-2126: stack, cast, [B, F, C, h, w] -> [B, C, F, h, w]; image size = latent size * 8 (:2102-2106)
    ret: Dict[str, Any] = {"controls": {}}
    ret["prompt_embeds"] = torch.stack([x["prompt_embeds"] for x in items]).to(dtype=dtype)
    if "actions" in items[0]:
        ret["controls"]["actions"] = torch.stack([x["actions"] for x in items]).to(dtype=dtype)
    ret["latents"] = torch.stack([x["latents"] for x in items]).to(dtype=dtype).permute(0, 2, 1, 3, 4)
    images = torch.stack([x["image"] for x in items]).to(dtype=dtype)
    ret["images"] = images.permute(0, 2, 1, 3, 4)
    ret["image_width"], ret["image_height"] = int(images.shape[-1] * 8), int(images.shape[-2] * 8)
    for k in ("latents_depth", "latents_label"):
        ret["controls"][k] = torch.stack([x[k] for x in items]).to(dtype=dtype).permute(0, 2, 1, 3, 4)
    return ret


GOLDEN_NAMES = ["00000_04_17", "00001_04_17", "00002_04_17"]


def golden_items(with_actions: bool = True) -> List[Dict[str, torch.Tensor]]:
    """The seeded per-clip items the collate golden was made from (3 clips, 2 views for the controls, 16 actions)."""
    files = synthetic_files(GOLDEN_NAMES, C=8, F=5, h=3, w=4, S=3, E=8, seed=3)  # small: the collate is shape-agnostic
    g = torch.Generator().manual_seed(4)
    items = []
    for n in GOLDEN_NAMES:
        it = reference_sample(files, n, views=[n, GOLDEN_NAMES[0]])
        if with_actions:
            it["actions"] = torch.randn(16, 7, generator=g)
        it["prompt"] = ""
        it["metainfo"] = {"num_view": 2, "num_frame": 17}  # required by the reference collate (dataset.py:2146-2148)
        items.append(it)
    return items
