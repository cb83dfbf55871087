"""CPU: the latent-format reader and the condition cache against the reference's read path as restated
in oracle/latent_oracle.py (dataset.py:655-694, 785-850, 1054-1059 for the sample; :2053-2126 for the collate), and the
collate against tests/golden/collate_control.pt — the output of the reference's OWN CollateFunctionControl
(oracle/make_latent_golden.py)."""
import os

import pytest
import torch

from orv_b200.latent_store import ConditionCache, LatentStore, collate_control, sample_name


from oracle import latent_oracle as LO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _write_dataset(root, names, **kw):
    """Files as encode_dataset.py writes them (:353-363), from the oracle's seeded stand-ins."""
    files = LO.synthetic_files(names, **kw)
    base = os.path.join(root, "emb", "val")
    for key, t in files.items():
        if key == "empty":
            torch.save(t, os.path.join(root, "emb", "empty_prompt.pt"))
            continue
        folder, n = key
        os.makedirs(os.path.join(base, folder), exist_ok=True)
        torch.save(t, os.path.join(base, folder, f"{n}.pt"))
    return files


_reference_sample = LO.reference_sample


def _reference_collate(items, dtype):
    ret = LO.reference_collate(items, dtype)
    ret.pop("image_width"), ret.pop("image_height")
    return ret


def _same(a, b):
    if isinstance(a, dict):
        assert set(a) == set(b), (set(a), set(b))
        for k in a:
            _same(a[k], b[k])
    elif isinstance(a, torch.Tensor):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)
    else:
        assert a == b


def test_sample_name_and_paths(tmp_path):
    assert sample_name(12, 3, 17) == "00012_03_17"
    assert sample_name(12, 3, 17, camera=0) == "00012_03_17_0"
    s = LatentStore(str(tmp_path), "emb", "val", ref_num=1, load_cond_gt=True)
    p = s.paths("00012_03_17")
    assert p["video_latents"].endswith(os.path.join("emb", "val", "video_latents", "00012_03_17.pt"))
    assert p["image_latents"].endswith(os.path.join("image1_latents", "00012_03_17.pt"))
    assert p["depth_latents"].endswith(os.path.join("depthGT_latents", "00012_03_17.pt"))
    assert p["empty_prompt"] == os.path.join(str(tmp_path), "emb", "empty_prompt.pt")


def test_load_matches_reference_read_path(tmp_path):
    names = [sample_name(e, 0, 17) for e in range(3)]
    files = _write_dataset(str(tmp_path), names)
    store = LatentStore(str(tmp_path), "emb", "val")
    for n in names:
        _same(store.load(n), _reference_sample(files, n))
    # multi-view clip: per-view files stacked then flattened to [(v f), C, h, w]
    got = store.load(names[0], view_names=names)
    _same(got, _reference_sample(files, names[0], views=names))
    assert got["latents_depth"].shape == (15, 32, 6, 8)
    # GT condition folders, per-sample prompt embeddings, one control key only
    store_gt = LatentStore(str(tmp_path), "emb", "val", load_cond_gt=True, control_keys=("depth",), empty_prompt=False)
    got = store_gt.load(names[1])
    assert "latents_label" not in got
    assert torch.equal(got["latents_depth"], _reference_sample(files, names[1], gt=True)["latents_depth"])
    assert torch.equal(got["prompt_embeds"], files[("prompt_embeds", names[1])])
    # frame-id handling of unsliced clips (:683-693): pixel frame ids map to latent frames, mismatches raise
    got = store.load(names[2], frame_ids=[0, 1, 2, 3, 4, 8], is_sliced=False)
    assert torch.equal(got["latents"], files[("video_latents", names[2])].permute(1, 0, 2, 3)[[0, 1, 2]])
    with pytest.raises(RuntimeError, match="mismatched latent video"):
        store.load(names[2], frame_ids=[0, 40], is_sliced=False)


def test_collate_matches_reference(tmp_path):
    names = [sample_name(e, 4, 17) for e in range(4)]
    files = _write_dataset(str(tmp_path), names)
    store = LatentStore(str(tmp_path), "emb", "val")
    items = [store.load(n) for n in names]
    got = collate_control(items, torch.float32)
    want = _reference_collate([_reference_sample(files, n) for n in names], torch.float32)
    assert got.pop("image_width") == 64 and got.pop("image_height") == 48
    _same(got, want)
    assert got["controls"]["latents_depth"].shape == (4, 32, 5, 6, 8)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_collate_matches_the_reference_classes_own_output(dt):
    """Product collate and oracle restatement, bit for bit against what the reference's CollateFunctionControl returned
    for the same seeded items (tensor keys + image size)."""
    want = torch.load(os.path.join(GOLDEN, "collate_control.pt"), weights_only=False)[str(dt).split(".")[-1]]
    items = LO.golden_items()
    _same(collate_control(items, dt), want)
    _same(LO.reference_collate(items, dt), want)


def test_condition_cache_prefetch_lru_and_values(tmp_path):
    names = [sample_name(e, 0, 17) for e in range(6)]
    files = _write_dataset(str(tmp_path), names)
    store = LatentStore(str(tmp_path), "emb", "val")
    cache = ConditionCache(store, weight_dtype=torch.bfloat16, capacity=3, workers=2, pin=False)
    cache.prefetch([[names[0]], [names[1], names[2]]])
    cache.prefetch([[names[0]]])  # already pending: no duplicate work
    a = cache.get([names[0]])
    b = cache.get([names[1], names[2]])
    assert cache.hits == 2 and cache.misses == 0
    want = _reference_collate([_reference_sample(files, names[1]), _reference_sample(files, names[2])], torch.bfloat16)
    b2 = {k: v for k, v in b.items() if k not in ("image_width", "image_height")}
    _same(b2, want)
    assert all(t.is_contiguous() for t in (b["latents"], b["controls"]["latents_label"], b["images"]))
    assert cache.get([names[0]]) is a  # resident
    c = cache.get(names[3])            # a bare name is a batch of one; cold miss
    assert cache.misses == 1 and c["latents"].shape[0] == 1
    cache.get([names[4]])              # evicts the least recently used entry ([names[1], names[2]])
    assert cache.get([names[0]]) is a
    assert cache.get([names[1], names[2]]) is not b
    # device fetch on the CPU is a plain copy with the same values
    d = cache.get_device([names[0]], "cpu")
    _same({k: v for k, v in d.items()}, {k: v for k, v in a.items()})
    cache.close()


@pytest.mark.gpu
def test_condition_cache_device_upload(tmp_path):
    names = [sample_name(e, 0, 17) for e in range(2)]
    _write_dataset(str(tmp_path), names)
    store = LatentStore(str(tmp_path), "emb", "val")
    cache = ConditionCache(store, weight_dtype=torch.bfloat16, capacity=4)
    cache.prefetch([[names[0]], [names[1]]])
    host = cache.get([names[0]])
    assert host["latents"].is_pinned()
    dev = cache.get_device([names[0]], "cuda:0")
    y = dev["controls"]["latents_depth"].float().sum()  # consumed on the current stream, ordered after the upload
    torch.cuda.synchronize()
    assert dev["latents"].is_cuda and torch.equal(dev["latents"].cpu(), host["latents"])
    assert torch.allclose(y.cpu(), host["controls"]["latents_depth"].float().sum(), rtol=1e-3)
    cache.close()
