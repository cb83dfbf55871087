"""CPU: the VAE oracle's structure and the host logic of orv_b200.AutoencoderKLCogVideoX.

The oracle (oracle/vae_oracle.py) restates diffusers' decoder — PARITY UNPINNED against the package itself (absent here).
What can be checked without it: the arithmetic facts the released checkpoint fixes (parameter names / shapes / count of
THUDM/CogVideoX-2b's `vae/`, 17 frames from 5 latent frames, the tile geometry the reference's settings produce), the
algebraic properties of the restated layers (causality, cache = continuation, nearest maps), and that the host mirror
(index maps instead of F.interpolate, one vector expression instead of the per-line blend loops, channels-last weight
packing) agrees with the restatement bit for bit.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as V
from orv_b200.models import autoencoder_kl_cogvideox as A


def test_released_geometry():
    """THUDM/CogVideoX-2b vae/config.json: 480 x 720 samples -> 240 x 360 sample tiles = 30 x 45 latent tiles stepping
    25 x 36, blended over 40 x 72 pixels, cropped to 200 x 288 (= 8 x the step: tiles abut exactly)."""
    g = V.tile_geometry(V.default_config())
    assert g == dict(latent_h=30, latent_w=45, step_h=25, step_w=36, blend_h=40, blend_w=72, limit_h=200, limit_w=288)
    m = A.AutoencoderKLCogVideoX()
    assert (m.tile_latent_min_height, m.tile_latent_min_width, m.tile_sample_min_height, m.tile_sample_min_width) == (30, 45, 240, 360)
    assert m.config.scaling_factor == 1.15258426 and m.config.temporal_compression_ratio == 4
    assert tuple(m.config.block_out_channels) == (128, 256, 256, 512)


def test_state_dict_names_match_oracle_and_checkpoint_layout():
    m = A.AutoencoderKLCogVideoX()
    shapes = V.decoder_param_shapes(V.default_config())
    sd = m.state_dict()
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == shapes[k] for k in sd)
    # the decoder half of the released checkpoint: 18 resnet blocks (2 mid + 4 x 4 up), 37 spatial norms, 3 upsamplers
    assert len([k for k in sd if k.endswith("conv1.conv.weight")]) == 18
    assert len([k for k in sd if k.endswith("norm_layer.weight")]) == 37
    assert len([k for k in sd if ".upsamplers.0.conv.weight" in k]) == 3
    assert sorted(k for k in sd if "conv_shortcut" in k) == [
        "decoder.up_blocks.1.resnets.0.conv_shortcut.bias", "decoder.up_blocks.1.resnets.0.conv_shortcut.weight",
        "decoder.up_blocks.3.resnets.0.conv_shortcut.bias", "decoder.up_blocks.3.resnets.0.conv_shortcut.weight"]
    assert sd["decoder.conv_in.conv.weight"].shape == (512, 16, 3, 3, 3)
    assert sd["decoder.conv_out.conv.weight"].shape == (3, 128, 3, 3, 3)
    # encoder / quant keys of a full checkpoint are ignored, decoder keys are strict
    full = dict(V.synthetic_state_dict(V.default_config(block_out_channels=(64, 64, 64, 64), layers_per_block=1)))
    small = A.AutoencoderKLCogVideoX(block_out_channels=(64, 64, 64, 64), layers_per_block=1)
    full["encoder.conv_in.conv.weight"] = torch.zeros(1)
    small.load_state_dict(full, strict=True)
    del full["decoder.conv_out.conv.bias"]
    with pytest.raises(RuntimeError):
        small.load_state_dict(full, strict=True)


def test_no_cpu_path():
    m = A.AutoencoderKLCogVideoX(block_out_channels=(64, 64, 64, 64), layers_per_block=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.decode(torch.zeros(1, 16, 1, 2, 2))
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1, 3, 1, 16, 16))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 13])
def test_frame_batches(n):
    got = A.frame_batches(n)
    assert got == V.frame_batches(n)
    covered = [i for s, e in got for i in range(s, e)]
    assert covered == list(range(n)) and all(e - s in (1, 2, 3) for s, e in got)
    assert n < 2 or got[0][1] - got[0][0] == 2 + n % 2  # the first batch takes the remainder


@pytest.mark.parametrize("t_in", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("compress", [True, False])
def test_upsample_frame_map_is_interpolate(t_in, compress):
    """The index map the CUDA upsampler uses = what the oracle's F.interpolate calls do (values are frame ids)."""
    x = torch.arange(t_in, dtype=torch.float32).view(1, 1, t_in, 1, 1).expand(1, 8, t_in, 2, 2).contiguous()
    sd = {"u.conv.weight": torch.zeros(8, 8, 3, 3), "u.conv.bias": torch.zeros(8)}
    sd["u.conv.weight"][torch.arange(8), torch.arange(8), 1, 1] = 1.0
    y = V.upsample3d(sd, "u", x, compress)
    assert y.shape[-2:] == (4, 4)
    assert y[0, 0, :, 0, 0].tolist() == [float(i) for i in A.upsample_frame_map(t_in, compress)]


@pytest.mark.parametrize("t_out,t_lat", [(3, 3), (5, 3), (9, 3), (2, 2), (4, 2), (8, 2), (1, 1), (7, 3), (6, 3)])
def test_spatial_norm_frame_map_is_interpolate(t_out, t_lat):
    zq = torch.arange(t_lat, dtype=torch.float32).view(1, 1, t_lat, 1, 1)
    if t_out > 1 and t_out % 2 == 1:
        z = torch.cat([F.interpolate(zq[:, :, :1], size=(1, 1, 1)), F.interpolate(zq[:, :, 1:], size=(t_out - 1, 1, 1))], 2)
    else:
        z = F.interpolate(zq, size=(t_out, 1, 1))
    assert z.flatten().tolist() == [float(i) for i in A.spatial_norm_frame_map(t_out, t_lat)]


def test_blend_matches_reference_loops_bit_for_bit():
    torch.manual_seed(0)
    for dt in (torch.bfloat16, torch.float32):
        a, b, c = torch.randn(3, 5, 40, 64).to(dt), torch.randn(3, 5, 24, 64).to(dt), torch.randn(3, 5, 40, 48).to(dt)
        want = V.blend_v(a[None].clone(), b[None].clone(), 40)[0]  # extent clipped to the shorter tile (24 rows)
        assert torch.equal(A.AutoencoderKLCogVideoX._blend(a.clone(), b.clone(), 40, 2), want)
        want = V.blend_h(a[None].clone(), c[None].clone(), 16)[0]
        assert torch.equal(A.AutoencoderKLCogVideoX._blend(a.clone(), c.clone(), 16, 3), want)
        assert not torch.equal(want, c)


def test_pack_conv_layout():
    """[c_out, c_in, kt, kh, kw] -> K-major [c_out8, (kt, kh, kw, c_in64)]: the K index the implicit GEMM walks."""
    w = torch.randn(5, 16, 3, 3, 3)
    m, bias, ker = A.AutoencoderKLCogVideoX._pack_conv(w, torch.arange(5.0), "cpu")
    assert m.shape == (8, 27 * 64) and bias.shape == (8,) and ker == (3, 3, 3)
    mm = m.view(8, 3, 3, 3, 64).float()
    assert torch.equal(mm[:5, ..., :16], w.permute(0, 2, 3, 4, 1).bfloat16().float())
    assert mm[5:].abs().sum() == 0 and mm[..., 16:].abs().sum() == 0 and bias[5:].abs().sum() == 0
    m2, _, ker2 = A.AutoencoderKLCogVideoX._pack_conv(torch.randn(64, 64, 3, 3), None, "cpu")
    assert m2.shape == (64, 9 * 64) and ker2 == (1, 3, 3)


# ---- properties of the restated layers -------------------------------------------------------------------------------
def _conv_sd(cin, cout, k=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"c.conv.weight": torch.randn(cout, cin, k, k, k, generator=g) * 0.1, "c.conv.bias": torch.randn(cout, generator=g)}


def test_causal_conv_is_causal_and_cache_continues():
    sd = _conv_sd(4, 6)
    x = torch.randn(1, 4, 5, 6, 7)
    y, cache = V.causal_conv3d(sd, "c", x, None)
    assert y.shape == (1, 6, 5, 6, 7) and torch.equal(cache, x[:, :, -2:])
    x2 = x.clone()
    x2[:, :, 3:] += 1.0  # later frames must not change earlier outputs
    y2, _ = V.causal_conv3d(sd, "c", x2, None)
    assert torch.equal(y[:, :, :3], y2[:, :, :3]) and not torch.equal(y[:, :, 3:], y2[:, :, 3:])
    # two batches with the cache = one pass over the concatenated frames
    ya, ca = V.causal_conv3d(sd, "c", x[:, :, :3], None)
    yb, _ = V.causal_conv3d(sd, "c", x[:, :, 3:], ca)
    assert torch.allclose(torch.cat([ya, yb], 2), y, atol=1e-6)
    # first batch: the first frame is repeated in front (pad_mode "constant" of diffusers, not zeros)
    xp = torch.cat([x[:, :, :1], x[:, :, :1], x], 2)
    want = F.conv3d(xp, sd["c.conv.weight"], sd["c.conv.bias"], padding=(0, 1, 1))
    assert torch.allclose(y, want, atol=1e-6)


def test_spatial_norm_commutes_with_latent_table():
    """conv_y / conv_b are 1x1x1, so evaluating them on the latent grid and gathering through the nearest map (what the
    CUDA path does) equals the restatement's resize-then-convolve."""
    torch.manual_seed(1)
    C, T, Tz, h, w, k = 64, 5, 3, 3, 4, 2
    cfg = V.default_config()
    sd = {"n.norm_layer.weight": torch.randn(C), "n.norm_layer.bias": torch.randn(C),
          "n.conv_y.conv.weight": torch.randn(C, 16, 1, 1, 1), "n.conv_y.conv.bias": torch.randn(C),
          "n.conv_b.conv.weight": torch.randn(C, 16, 1, 1, 1), "n.conv_b.conv.bias": torch.randn(C)}
    f, zq = torch.randn(1, C, T, h << k, w << k), torch.randn(1, 16, Tz, h, w)
    want = V.spatial_norm(sd, "n", f, zq, cfg)
    ty = F.conv3d(zq, sd["n.conv_y.conv.weight"], sd["n.conv_y.conv.bias"])
    tb = F.conv3d(zq, sd["n.conv_b.conv.weight"], sd["n.conv_b.conv.bias"])
    ts = torch.tensor(A.spatial_norm_frame_map(T, Tz))
    hs, ws = torch.arange(h << k) >> k, torch.arange(w << k) >> k
    gy = ty[:, :, ts][:, :, :, hs][:, :, :, :, ws]
    gb = tb[:, :, ts][:, :, :, hs][:, :, :, :, ws]
    got = F.group_norm(f, 32, sd["n.norm_layer.weight"], sd["n.norm_layer.bias"], 1e-6) * gy + gb
    assert torch.allclose(got, want, atol=1e-5)


def test_decode_shapes_and_tiling_consistency():
    """5 latent frames -> 17 frames; an untiled latent decodes the same with tiling on; a tiled decode has the full
    size and equals the untiled one far from the tile seams' influence only approximately (GroupNorm statistics are per
    tile) — so only shape and finiteness are asserted for it."""
    cfg = V.default_config(block_out_channels=(32, 32, 32, 32), layers_per_block=1, sample_height=96, sample_width=160)
    sd = V.synthetic_state_dict(cfg, seed=3)
    z = torch.randn(1, 16, 5, 4, 6)
    y = V.decode(sd, cfg, z, tiling=True)
    assert y.shape == (1, 3, 17, 32, 48) and torch.equal(y, V.decode(sd, cfg, z, tiling=False))
    zt = torch.randn(1, 16, 3, 8, 13)
    yt = V.decode(sd, cfg, zt, tiling=True)
    assert yt.shape == (1, 3, 9, 64, 104) and torch.isfinite(yt).all()
    # frame batches: the first three latent frames decode identically whether or not later frames follow (causality)
    assert torch.allclose(V.decode(sd, cfg, z[:, :, :3], tiling=False), y[:, :, :9], atol=1e-5)
