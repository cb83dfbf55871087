"""Reader for the reference's on-disk latent format + a pinned-memory condition cache.

The reference encodes every clip once (`orv/dataset/encode_dataset.py:820-900`) into per-sample `.pt` files of VAE
**moments** `[2*16, F, h, w]` (3-D VAE layout `[C, F, H, W]`) under
`<data_root>/<embeddings_folder>/<split>/{video_latents, image<ref_num>_latents, depth_latents, label_latents,
depthGT_latents, labelGT_latents, prompt_embeds}/<sample_name>.pt` plus one shared
`<data_root>/<embeddings_folder>/empty_prompt.pt`, and reads them back sample by sample with `torch.load` inside the
DataLoader workers (`orv/dataset/dataset.py:655-694`, `:785-850`, `:1054-1059`), collating to `[B, C, F, h, w]`
(`:2092-2126`) before the eval loop copies them to the GPU with pageable `.to(device, dtype)` calls
(`orv/pipeline/evaluation_control_to_video.py:320-336`).

Here the same files, keys and tensor layouts come out of `LatentStore`, and `ConditionCache` keeps the decoded
samples in **pinned** host memory, loads ahead on a background thread (the loop either side of the denoise path is
serial host I/O), and issues the host→device copies of the next clip on a side stream while the current clip is
being denoised.  No arithmetic happens here: the format is the reference's, bit for bit.
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Dict, Iterable, List, Optional, Sequence

import torch

__all__ = ["LatentStore", "ConditionCache", "sample_name", "collate_control"]


def sample_name(episode_id: int, start_frame_idx: int, num_frame: int, camera: Optional[int] = None) -> str:
    """`f'{episode_id:05d}_{start_frame_idx:02d}_{num_frame:02d}'` (+ `_<camera>` for multi-camera datasets);
    dataset.py:1048-1053."""
    name = f"{int(episode_id):05d}_{int(start_frame_idx):02d}_{int(num_frame):02d}"
    return name if camera is None else f"{name}_{int(camera)}"


class LatentStore:
    """Path layout and per-sample reader of the reference's encoded dataset."""

    def __init__(self, data_root: str, embeddings_folder: str, split: str, *, ref_num: int = 1, use_3dvae: bool = True,
                 control_keys: Sequence[str] = ("depth", "label"), load_cond_gt: bool = False,
                 empty_prompt: bool = True):
        self.data_root = str(data_root)
        self.embeddings_folder = str(embeddings_folder)
        self.split = str(split)
        self.ref_num = int(ref_num)
        self.use_3dvae = bool(use_3dvae)
        self.control_keys = tuple(control_keys)
        self.load_cond_gt = bool(load_cond_gt)
        self.empty_prompt = bool(empty_prompt)
        self._empty_prompt: Optional[torch.Tensor] = None
        self._lock = threading.Lock()

    # ---- paths (dataset.py:1056-1059, 1075-1087, 1124-1137) ----------------------------------------------------
    def paths(self, name: str) -> Dict[str, str]:
        base = os.path.join(self.data_root, self.embeddings_folder, self.split)
        gt = "GT" if self.load_cond_gt else ""
        return {
            "video_latents": os.path.join(base, "video_latents", f"{name}.pt"),
            "image_latents": os.path.join(base, f"image{self.ref_num}_latents", f"{name}.pt"),
            "depth_latents": os.path.join(base, f"depth{gt}_latents", f"{name}.pt"),
            "label_latents": os.path.join(base, f"label{gt}_latents", f"{name}.pt"),
            "prompt_embeds": os.path.join(base, "prompt_embeds", f"{name}.pt"),
            "empty_prompt": os.path.join(self.data_root, self.embeddings_folder, "empty_prompt.pt"),
        }

    @staticmethod
    def _read(path: str) -> torch.Tensor:
        with open(path, "rb") as f:
            return torch.load(f, weights_only=True)

    def _frames_first(self, t: torch.Tensor) -> torch.Tensor:
        # 3-D VAE files are [C, F, H, W]; samples are handed on as [F, C, H, W] (dataset.py:678-682, 817-820)
        return t.permute(1, 0, 2, 3) if self.use_3dvae else t

    def prompt_embeds(self, name: str) -> torch.Tensor:
        if self.empty_prompt:
            with self._lock:
                if self._empty_prompt is None:
                    self._empty_prompt = self._read(self.paths(name)["empty_prompt"])[0]  # drop the batch dim (:1058)
            return self._empty_prompt
        return self._read(self.paths(name)["prompt_embeds"])

    def load(self, name: str, *, view_names: Optional[Sequence[str]] = None, with_video: bool = True,
             frame_ids: Optional[Sequence[int]] = None, is_sliced: bool = True) -> Dict[str, torch.Tensor]:
        """One sample with the reference's keys: `latents [F, C, h, w]` and `image [F_ref, C, h, w]`
        (dataset.py:693-694), `latents_depth` / `latents_label` `[V*F, C, h, w]` (views stacked then flattened,
        :825-827, :845-848), `prompt_embeds [S, E]`.  `view_names` lists the per-view sample names of a multi-view
        clip (default: this sample alone)."""
        p = self.paths(name)
        out: Dict[str, torch.Tensor] = {"prompt_embeds": self.prompt_embeds(name)}
        if with_video and os.path.exists(p["video_latents"]):
            video = self._frames_first(self._read(p["video_latents"]))
            ids = list(frame_ids) if frame_ids is not None else list(range(video.size(0)))
            if self.use_3dvae and frame_ids is not None:
                ids = sorted({i // 4 for i in ids})  # pixel frame ids -> latent frame ids (:683)
            if is_sliced:
                ids = list(range(video.size(0)))     # (:686-687)
            if video.shape[0] <= max(ids):
                raise RuntimeError(f"Got mismatched latent video and frame ids: {tuple(video.shape)} v.s. {ids}, "
                                   f"path: {p['video_latents']}.")
            out["latents"] = video[ids]
        if os.path.exists(p["image_latents"]):
            out["image"] = self._frames_first(self._read(p["image_latents"]))
        names = list(view_names) if view_names is not None else [name]
        for key, folder in (("depth", "depth_latents"), ("label", "label_latents")):
            if key not in self.control_keys:
                continue
            views = [self._frames_first(self._read(self.paths(n)[folder])) for n in names]
            out[f"latents_{key}"] = torch.stack(views).flatten(0, 1)
        return out


def collate_control(items: List[Dict[str, Any]], weight_dtype: torch.dtype) -> Dict[str, Any]:
    """`CollateFunctionControl.__call__` for the tensor keys of this path (dataset.py:2072-2126): stack, cast,
    `[B, F, C, h, w] -> [B, C, F, h, w]`."""
    ret: Dict[str, Any] = {"controls": {}}
    keys = items[0].keys()
    if "prompt_embeds" in keys:
        ret["prompt_embeds"] = torch.stack([x["prompt_embeds"] for x in items]).to(dtype=weight_dtype)
    if "actions" in keys:
        ret["controls"]["actions"] = torch.stack([x["actions"] for x in items]).to(dtype=weight_dtype)
    if "latents" in keys:
        ret["latents"] = torch.stack([x["latents"] for x in items]).to(dtype=weight_dtype).permute(0, 2, 1, 3, 4)
    if "image" in keys:
        images = torch.stack([x["image"] for x in items]).to(dtype=weight_dtype)
        ret["images"] = images.permute(0, 2, 1, 3, 4)
        fh, fw = images.shape[-2:]
        ret["image_width"], ret["image_height"] = int(fw * 8), int(fh * 8)
    for key in ("latents_depth", "latents_label"):
        if key in keys:
            ret["controls"][key] = torch.stack([x[key] for x in items]).to(dtype=weight_dtype).permute(0, 2, 1, 3, 4)
    return ret


def _map_tensors(obj, fn):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    return obj


class ConditionCache:
    """LRU cache of collated clips in pinned host memory with read-ahead and side-stream upload.

        cache = ConditionCache(store, weight_dtype=torch.bfloat16, capacity=64)
        cache.prefetch(names[i + 1 : i + 4])            # background torch.load + collate + pin
        batch = cache.get_device([names[i]], device)    # H2D on the cache's copy stream, event-ordered

    `get()` returns exactly what `collate_control([store.load(n) ...])` returns (same values, contiguous), so a
    caller can swap it in for the DataLoader + `.to(device, dtype)` pair of the reference's eval loop.
    """

    def __init__(self, store: LatentStore, weight_dtype: torch.dtype = torch.bfloat16, capacity: int = 32,
                 workers: int = 2, pin: Optional[bool] = None):
        self.store = store
        self.weight_dtype = weight_dtype
        self.capacity = int(capacity)
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._lru: "OrderedDict[tuple, Any]" = OrderedDict()
        self._pending: Dict[tuple, Future] = {}
        self._lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="orvb-latents")
        self._copy_stream = None
        self.hits = 0
        self.misses = 0

    # ---- host side ----------------------------------------------------------------------------------------------
    def _build(self, names: tuple) -> Dict[str, Any]:
        batch = collate_control([self.store.load(n) for n in names], self.weight_dtype)

        def fix(t: torch.Tensor) -> torch.Tensor:
            t = t.contiguous()
            return t.pin_memory() if self.pin else t
        return _map_tensors(batch, fix)

    def prefetch(self, groups: Iterable[Sequence[str]]) -> None:
        """Schedules background loads; each element is the list of sample names of one batch (or a single name)."""
        for g in groups:
            key = (g,) if isinstance(g, str) else tuple(g)
            with self._lock:
                if key in self._lru or key in self._pending:
                    continue
                self._pending[key] = self._pool.submit(self._build, key)

    def get(self, names: Sequence[str]) -> Dict[str, Any]:
        key = (names,) if isinstance(names, str) else tuple(names)
        with self._lock:
            if key in self._lru:
                self._lru.move_to_end(key)
                self.hits += 1
                return self._lru[key]
            fut = self._pending.pop(key, None)
        if fut is None:
            self.misses += 1
            batch = self._build(key)
        else:
            self.hits += 1
            batch = fut.result()
        with self._lock:
            self._lru[key] = batch
            while len(self._lru) > self.capacity:
                self._lru.popitem(last=False)
        return batch

    # ---- device side --------------------------------------------------------------------------------------------
    def get_device(self, names: Sequence[str], device) -> Dict[str, Any]:
        """The batch on `device`.  Copies run on a dedicated stream (non-blocking from pinned memory) and the current
        stream waits on their event, so a clip uploaded while the previous one is denoising costs no time on the
        compute stream."""
        batch = self.get(names)
        device = torch.device(device)
        if device.type != "cuda":
            return _map_tensors(batch, lambda t: t.to(device))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(self._copy_stream):
            out = _map_tensors(batch, lambda t: t.to(device, non_blocking=True))
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        cur = torch.cuda.current_stream(device)
        cur.wait_event(done)
        _map_tensors(out, lambda t: (t.record_stream(cur), t)[1])
        return out

    def close(self) -> None:
        self._pool.shutdown(wait=True)
