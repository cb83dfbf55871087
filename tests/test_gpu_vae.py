"""3-D VAE decode on a B200, through the C ABI, against oracle/vae_oracle.py.

The oracle restates diffusers' `AutoencoderKLCogVideoX.decode` (PARITY UNPINNED: the package is absent here and the
reference ships no decoded frames — see the oracle's header); these tests hold the CUDA path to that restatement:
operator by operator at the north-star tolerance (rtol 1e-3 / atol 1e-4, fp32 test outputs, bf16-exact operands) and
end to end (bf16 product outputs vs the fp32 oracle with the three forward gates of tests/_gates.py), including the
reference's deployment settings: frame batches with convolution caches, tiling with blending, slicing.
Reference call sites: orv/models/cogvideox_control.py:1095-1100, :1476-1479; inference_control_to_video.py:98-99.
"""
import pytest
import torch
import torch.nn.functional as F

from _gates import assert_forward_close

pytestmark = pytest.mark.gpu

DEV = "cuda"
RTOL, ATOL = 1e-3, 1e-4


@pytest.fixture(scope="module")
def ops():
    from orv_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def V():
    from oracle import vae_oracle
    return vae_oracle


def _bf(*shape, k=1.0):
    return (torch.randn(*shape, device=DEV) * k).bfloat16()


def _close(got, ref, rtol=RTOL, atol=ATOL):
    got, ref = got.double(), ref.double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all()
    bad = (got - ref).abs() > atol + rtol * ref.abs()
    assert not bad.any(), (f"{int(bad.sum())} of {bad.numel()} elements outside rtol={rtol} atol={atol}; "
                           f"max abs err {(got - ref).abs().max().item():.3e}")


def _cl(x):  # [C, T, H, W] -> [T, H, W, C]
    return x.permute(1, 2, 3, 0).contiguous()


def _pack_w(w):  # [cout, cin, kt, kh, kw] -> [cout, taps * cin]
    return w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1).contiguous()


def _ref_conv(x, w, b, cache=None, resid=None):
    """fp64 causal convolution of a channels-last tensor (oracle.causal_conv3d semantics)."""
    kt, kh, kw = w.shape[2:]
    xx = x.permute(3, 0, 1, 2)[None].double()
    if kt > 1:
        ctx = cache.permute(3, 0, 1, 2)[None].double() if cache is not None else xx[:, :, :1].expand(-1, -1, kt - 1, -1, -1)
        xx = torch.cat([ctx, xx], 2)
    y = F.conv3d(xx, w.double(), None if b is None else b.double(), padding=(0, kh // 2, kw // 2))[0].permute(1, 2, 3, 0)
    return y + resid.double() if resid is not None else y


@pytest.mark.parametrize("T,H,W,cin,cout,ker,cache,resid", [
    (3, 13, 21, 64, 72, (3, 3, 3), False, False),    # ragged patches in both directions, no cache: frame 0 repeated
    (2, 16, 32, 128, 128, (3, 3, 3), True, True),    # cache from the previous frame batch + residual epilogue
    (1, 9, 17, 64, 64, (3, 3, 3), True, False),      # single frame: both temporal taps come from the cache
    (4, 30, 45, 64, 256, (1, 3, 3), False, False),   # per-frame Conv2d (upsampler), the latent tile geometry
    (3, 8, 16, 192, 64, (1, 1, 1), False, False),    # 1x1x1 shortcut
    (2, 24, 40, 64, 8, (3, 3, 3), True, False),      # conv_out: 8 padded output channels
    (5, 60, 90, 256, 256, (3, 3, 3), True, True),    # several waves of tiles
])
def test_conv_cl_f32(ops, T, H, W, cin, cout, ker, cache, resid):
    torch.manual_seed(T * 100 + H)
    kt, kh, kw = ker
    x = _bf(T, H, W, cin, k=0.5)
    w = _bf(cout, cin, kt, kh, kw, k=(cin * kt * kh * kw) ** -0.5)
    b = _bf(cout, k=0.1)
    c = _bf(kt - 1, H, W, cin, k=0.5) if (cache and kt > 1) else None
    r = _bf(T, H, W, cout) if resid else None
    out = ops.conv_cl(x, _pack_w(w), b, ker, cache=c, resid=r, out_f32=True)
    _close(out, _ref_conv(x, w, b, c, r))
    # product path: bf16 output through the TMA store, one rounding away from the fp32 result
    out16 = ops.conv_cl(x, _pack_w(w), b, ker, cache=c, resid=r)
    assert torch.equal(out16, out.bfloat16())


@pytest.mark.parametrize("T,H,W,cin,cout,resid", [(3, 13, 21, 64, 128, False), (2, 30, 45, 128, 256, True),
                                                  (5, 60, 90, 64, 512, True), (1, 9, 17, 64, 64, False)])
def test_conv_cl_fused_groupnorm_stats(ops, T, H, W, cin, cout, resid):
    """Statistics accumulated in the convolution's epilogue = statistics of the tensor it stored (what orvb_gn_stats_cl
    computes from that tensor), deterministic, and the output itself is unchanged by asking for them."""
    torch.manual_seed(cout + T)
    x = _bf(T, H, W, cin, k=0.5)
    w = _bf(cout, cin, 3, 3, 3, k=(27 * cin) ** -0.5)
    b = _bf(cout, k=0.3) + 0.2
    r = _bf(T, H, W, cout) if resid else None
    plain = ops.conv_cl(x, _pack_w(w), b, (3, 3, 3), resid=r)
    out, st = ops.conv_cl(x, _pack_w(w), b, (3, 3, 3), resid=r, gn=(32, 1e-6))
    assert torch.equal(out, plain)
    want = ops.gn_stats_cl(out, 32, 1e-6)
    _close(st[:, 0], want[:, 0], 1e-5, 1e-6)
    _close(st[:, 1], want[:, 1], 1e-5, 1e-6)
    xd = out.double().reshape(-1, 32, cout // 32)
    _close(st[:, 0], xd.mean(dim=(0, 2)), 1e-5, 1e-6)
    _close(st[:, 1], (xd.var(dim=(0, 2), unbiased=False) + 1e-6).rsqrt(), 1e-5, 1e-6)
    out2, st2 = ops.conv_cl(x, _pack_w(w), b, (3, 3, 3), resid=r, gn=(32, 1e-6))
    assert torch.equal(st, st2) and torch.equal(out, out2)


def test_conv_cl_resid_in_place(ops):
    """conv2 of a resnet block writes over its residual input."""
    torch.manual_seed(3)
    x, r = _bf(2, 16, 16, 64, k=0.5), _bf(2, 16, 16, 64)
    w, b = _bf(64, 64, 3, 3, 3, k=0.03), _bf(64, k=0.1)
    want = ops.conv_cl(x, _pack_w(w), b, (3, 3, 3), resid=r)
    got = ops.conv_cl(x, _pack_w(w), b, (3, 3, 3), resid=r, out=r)
    assert got.data_ptr() == r.data_ptr() and torch.equal(got, want)


def test_conv_cl_rejects_bad_shapes(ops):
    x = _bf(1, 8, 8, 48)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.conv_cl(x, _bf(64, 27 * 48), None, (3, 3, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.conv_cl(x.cpu(), _bf(64, 27 * 48), None, (3, 3, 3))


@pytest.mark.parametrize("pixels,C,groups", [(1, 64, 32), (777, 128, 32), (4050, 512, 32), (100000, 256, 32), (3000, 192, 8)])
def test_gn_stats(ops, pixels, C, groups):
    torch.manual_seed(pixels)
    x = (torch.randn(pixels, C, device=DEV) * torch.linspace(0.5, 3.0, C, device=DEV) + torch.linspace(-2, 2, C, device=DEV)).bfloat16()
    st = ops.gn_stats_cl(x, groups, 1e-6)
    xd = x.double().reshape(pixels, groups, C // groups)
    mean = xd.mean(dim=(0, 2))
    var = xd.var(dim=(0, 2), unbiased=False)
    _close(st[:, 0], mean, 1e-5, 1e-6)
    _close(st[:, 1], (var + 1e-6).rsqrt(), 1e-5, 1e-6)
    assert torch.equal(st, ops.gn_stats_cl(x, groups, 1e-6))  # deterministic reduction order


@pytest.mark.parametrize("T,Tz,h,w,shift,C", [(3, 3, 6, 9, 0, 128), (5, 3, 6, 9, 1, 256), (9, 3, 5, 7, 2, 64), (8, 2, 6, 9, 3, 128),
                                              (1, 1, 4, 4, 0, 64), (2, 2, 7, 5, 1, 512)])
def test_spatial_norm_silu_f32(ops, V, T, Tz, h, w, shift, C):
    """CogVideoXSpatialNorm3D + SiLU against the oracle (which resizes zq with F.interpolate and runs the 1x1x1 convs at
    full resolution)."""
    from orv_b200.models.autoencoder_kl_cogvideox import spatial_norm_frame_map
    torch.manual_seed(T * 10 + shift)
    H, W = h << shift, w << shift
    f = _bf(C, T, H, W, k=1.5) + 0.25
    f = f.bfloat16()
    zq = _bf(16, Tz, h, w)
    sd = {"n.norm_layer.weight": _bf(C) * 0.1 + 1, "n.norm_layer.bias": _bf(C, k=0.1),
          "n.conv_y.conv.weight": _bf(C, 16, 1, 1, 1, k=0.1), "n.conv_y.conv.bias": _bf(C, k=0.1) + 1,
          "n.conv_b.conv.weight": _bf(C, 16, 1, 1, 1, k=0.25), "n.conv_b.conv.bias": _bf(C, k=0.1)}
    sd = {k: v.bfloat16() for k, v in sd.items()}
    cfg = V.default_config()
    want = F.silu(V.spatial_norm({k: v.double() for k, v in sd.items()}, "n", f[None].double(), zq[None].double(), cfg))[0]
    # table: conv_y | conv_b per latent pixel, in fp64 then rounded like the GEMM's bf16 output would be -> use exact
    # values to test THIS kernel: keep the table bf16-exact by construction (weights chosen so products are exact is not
    # possible in general), so compare against a reference built from the same bf16 table instead
    z64 = torch.zeros(Tz * h * w, 64, device=DEV, dtype=torch.bfloat16)
    z64[:, :16] = _cl(zq).reshape(-1, 16)
    wt = torch.zeros(2 * C, 64, device=DEV, dtype=torch.bfloat16)
    wt[:C, :16] = sd["n.conv_y.conv.weight"].reshape(C, 16)
    wt[C:, :16] = sd["n.conv_b.conv.weight"].reshape(C, 16)
    bt = torch.cat([sd["n.conv_y.conv.bias"], sd["n.conv_b.conv.bias"]])
    table = ops.gemm(z64, wt, bt)
    x = _cl(f)
    st = ops.gn_stats_cl(x, 32, 1e-6)
    t_src = torch.tensor(spatial_norm_frame_map(T, Tz), dtype=torch.int32, device=DEV)
    got = ops.spatial_norm_cl(x, st, sd["n.norm_layer.weight"], sd["n.norm_layer.bias"], table, 0, C, t_src, (h, w), shift,
                              y_f32=True)
    # (a) the whole operator against the oracle: the bf16 rounding of the conv_y / conv_b table (2^-9 relative, the same
    # rounding the reference's bf16 convolution output carries) bounds the difference
    assert_forward_close(got, want.permute(1, 2, 3, 0), 4e-3, 8e-3, 1.5e-2, "spatial norm vs oracle")
    # (b) this kernel's own arithmetic at the north-star tolerance: same bf16 table, float64 evaluation
    tb = table.double()
    src = (t_src.long()[:, None, None] * h + (torch.arange(H, device=DEV) >> shift)[None, :, None]) * w \
        + (torch.arange(W, device=DEV) >> shift)[None, None, :]
    yv, bv = tb[src][..., :C], tb[src][..., C:]
    nf = F.group_norm(f[None].double(), 32, sd["n.norm_layer.weight"].double(), sd["n.norm_layer.bias"].double(), 1e-6)[0]
    _close(got, F.silu(nf.permute(1, 2, 3, 0) * yv + bv))
    # product output = the fp32 values rounded once
    got16 = ops.spatial_norm_cl(x, st, sd["n.norm_layer.weight"], sd["n.norm_layer.bias"], table, 0, C, t_src, (h, w), shift)
    assert torch.equal(got16, got.bfloat16())


@pytest.mark.parametrize("T,compress", [(3, True), (2, True), (1, True), (5, False), (4, True)])
def test_upsample_matches_interpolate(ops, V, T, compress):
    """The nearest-neighbour part of CogVideoXUpsample3D, bit-exact against F.interpolate as the oracle calls it."""
    from orv_b200.models.autoencoder_kl_cogvideox import upsample_frame_map
    torch.manual_seed(T)
    C, H, W = 64, 5, 7
    x = _bf(C, T, H, W)
    sd = {"u.conv.weight": torch.zeros(C, C, 3, 3, device=DEV), "u.conv.bias": torch.zeros(C, device=DEV)}
    sd["u.conv.weight"][torch.arange(C), torch.arange(C), 1, 1] = 1.0  # identity convolution: the oracle returns the resize
    want = V.upsample3d(sd, "u", x[None].float(), compress)[0]
    t_src = torch.tensor(upsample_frame_map(T, compress), dtype=torch.int32, device=DEV)
    got = ops.upsample2x_cl(_cl(x), t_src)
    assert torch.equal(got.permute(3, 0, 1, 2).float(), want)


def test_cl_to_planar(ops):
    x = _bf(3, 10, 12, 8)
    assert torch.equal(ops.cl_to_planar(x, 3), x[..., :3].permute(3, 0, 1, 2).contiguous())


# ---------------------------------------------------------------------------------------------------------------------
# end to end
# ---------------------------------------------------------------------------------------------------------------------
def _models(V, **over):
    from orv_b200 import AutoencoderKLCogVideoX
    cfg = V.default_config(**over)
    sd32 = V.synthetic_state_dict(cfg, seed=1)
    sd = {k: v.bfloat16().float() for k, v in sd32.items()}  # bf16-exact weights on both sides
    keys = ("in_channels", "out_channels", "block_out_channels", "latent_channels", "layers_per_block", "norm_eps",
            "norm_num_groups", "temporal_compression_ratio", "sample_height", "sample_width", "scaling_factor",
            "invert_scale_latents")
    m = AutoencoderKLCogVideoX(**{k: cfg[k] for k in keys})
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV, torch.bfloat16).eval()
    return cfg, {k: v.to(DEV) for k, v in sd.items()}, m


# sample 96 x 160 -> latent tiles of 6 x 10 stepping 5 x 8, blended over 8 x 16 output pixels, cropped to 40 x 64: the
# released geometry (480 x 720 -> 30 x 45 / 25 x 36 / 40 x 72 / 200 x 288) scaled down so that, like there, the crop is
# exactly 8 x the step
SMALL = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=1, sample_height=96, sample_width=160)


@pytest.mark.parametrize("T,h,w", [(5, 4, 6), (4, 3, 5), (1, 4, 6), (3, 2, 3)])
def test_decode_untiled_small(V, T, h, w):
    """Frame batches with convolution caches (5 latent frames -> [0:3], [3:5] -> 17 frames), odd / even / single-frame
    temporal upsampling; no tiling (latent not larger than the 6 x 10 tile)."""
    cfg, sd, m = _models(V, **SMALL)
    torch.manual_seed(T)
    z = torch.randn(2, 16, T, h, w, device=DEV).bfloat16()
    want = V.decode(sd, cfg, z.float(), tiling=True)
    got = m.decode(z).sample
    frames = 4 * (T - 1) + 1 if T % 2 else 4 * T  # an even count has no single first frame: every batch doubles twice
    assert got.shape == want.shape == (2, 3, frames, 8 * h, 8 * w) and got.dtype == torch.bfloat16
    assert_forward_close(got, want, 2e-2, 4e-2, 8e-2, f"decode untiled T={T} {h}x{w}")


@pytest.mark.parametrize("h,w", [(8, 13), (12, 20)])
def test_decode_tiled_small(V, h, w):
    """enable_tiling(): 2 x 2 and 3 x 3 tile grids with ragged last tiles."""
    cfg, sd, m = _models(V, **SMALL)
    m.enable_tiling()
    m.enable_slicing()
    torch.manual_seed(h)
    z = torch.randn(1, 16, 3, h, w, device=DEV).bfloat16()
    want = V.decode(sd, cfg, z.float(), tiling=True)
    got = m.decode(z).sample
    assert got.shape == want.shape == (1, 3, 9, 8 * h, 8 * w)
    assert_forward_close(got, want, 2e-2, 4e-2, 8e-2, f"decode tiled {h}x{w}")
    # tiling off: the same latent decoded whole differs from the tiled result (so the test above does exercise tiles)
    m.disable_tiling()
    whole = m.decode(z).sample
    assert_forward_close(whole, V.decode(sd, cfg, z.float(), tiling=False), 2e-2, 4e-2, 8e-2, "decode whole")
    assert not torch.equal(whole, got)


def test_decode_deterministic_and_launch_count(V):
    cfg, sd, m = _models(V, **SMALL)
    z = torch.randn(1, 16, 5, 4, 6, device=DEV).bfloat16()
    a = m.decode(z).sample
    n = m.last_launches
    b = m.decode(z).sample
    assert torch.equal(a, b) and n == m.last_launches and n > 0


def test_decode_full_width_one_tile(V):
    """The released geometry (128/256/256/512 channels, 3 resnets per block) on one frame batch of a small latent, against
    the oracle evaluated in fp32 on the same GPU (cuDNN, TF32 off)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, sd, m = _models(V)
    torch.manual_seed(11)
    z = torch.randn(1, 16, 3, 8, 12, device=DEV).bfloat16()
    want = V.decode(sd, cfg, z.float(), tiling=True)
    got = m.decode(z).sample
    assert got.shape == (1, 3, 9, 64, 96)
    assert_forward_close(got, want, 2.5e-2, 5e-2, 1e-1, "decode full width")


def test_pipeline_decodes_without_diffusers(V):
    """pipe(..., output_type='pt' | 'pil') with orv_b200.AutoencoderKLCogVideoX attached (tiling + slicing on, as the
    reference script sets them): latents -> decode_latents -> postprocess_video, no diffusers object anywhere
    (reference :1476-1479)."""
    from oracle import flat_oracle as O
    from oracle import make_golden as G
    from orv_b200 import CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj, CogVideoXTransformer3DModelTraj
    vcfg, vsd, vae = _models(V, **SMALL)
    vae.enable_slicing()
    vae.enable_tiling()
    cfg = O.default_config(**G.BASE)
    model = CogVideoXTransformer3DModelTraj(**cfg)
    model.load_state_dict(O.synthetic_state_dict(cfg, seed=0, std=0.05), strict=False)
    model.action_embed.mask = False
    model = model.to(DEV, torch.bfloat16).eval()
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, vae, model, CogVideoXDPMScheduler(timestep_spacing="trailing"))
    inp = O.synthetic_inputs(cfg, 1, 3, 6, 8, seed=1, n_actions=8)
    moments = torch.randn(1, 32, 1, 6, 8, generator=torch.Generator().manual_seed(7)).bfloat16()
    kw = dict(image=moments, prompt="", prompt_embeds=inp["text"].cuda().bfloat16(), height=48, width=64, num_frames=9,
              num_inference_steps=2, guidance_scale=1.0, controls_or_guidances={"actions": inp["actions"]})
    lat = pipe(**kw, generator=torch.Generator().manual_seed(5), output_type="latent").frames
    vid = pipe(**kw, generator=torch.Generator().manual_seed(5), output_type="pt").frames
    assert vid.shape == (1, 9, 3, 48, 64) and float(vid.min()) >= 0.0 and float(vid.max()) <= 1.0
    want = V.decode(vsd, vcfg, (lat.float() / vcfg["scaling_factor"]).permute(0, 2, 1, 3, 4), tiling=True)
    want = (want.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
    assert (vid.float() - want).abs().max().item() < 3e-2
    pil = pipe(**kw, generator=torch.Generator().manual_seed(5), output_type="pil").frames
    assert len(pil) == 1 and len(pil[0]) == 9 and pil[0][0].size == (64, 48)
