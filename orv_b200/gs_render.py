"""3-D Gaussian rasteriser, forward pass — drop-in for the reference's `diff_gaussian_rasterization` package
(orv/ops/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:150-237) and for `render()` of
orv/dataset/gs_render.py:103-171, which turns occupancy voxels into RGB / semantic / depth / alpha maps.

Same names, argument lists and return tuple `(color, language_feature, radii, depth, alpha)`; the arithmetic runs in
liborv_b200.so (`orvb_gs_rasterize`, csrc/gs_render.cu).  Forward only: the reference's backward pass is a training
facility — tensors that require grad are accepted, the outputs carry no graph.  Spherical harmonics are not
implemented (the ORV caller passes `colors_precomp`, shs=None).  No fallback: CUDA tensors on a B200 or an error.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _lib as L

NUM_CHANNELS = 3                    # cuda_rasterizer/config.h:14
NUM_CHANNELS_LANGUAGE_FEATURE = 12  # cuda_rasterizer/config.h:15


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    include_feature: bool


_WORKSPACE = {}
_LAST_COUNT = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index)
    t = _WORKSPACE.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)  # cudaMalloc'ed blocks are 512-byte aligned
        _WORKSPACE[key] = t
    return t


def _f32(t: Optional[torch.Tensor], name: str, device) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (liborv_b200 has no CPU path)")
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, language_feature_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings: GaussianRasterizationSettings, max_instances: int = 0):
    """`_RasterizeGaussians.forward` (reference __init__.py:42-88).  Returns (color [3,H,W], language_feature
    [12,H,W] or an empty tensor, radii [P] int32, depth [1,H,W], alpha [1,H,W]).  `max_instances` (0 = automatic):
    capacity of the (Gaussian, tile) instance buffers; on overflow the call is repeated once with the exact count."""
    if sh is not None and sh.numel() != 0:
        raise NotImplementedError("spherical-harmonics colours are not implemented: pass colors_precomp "
                                  "(orv/dataset/gs_render.py:152 does)")
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    rs = raster_settings
    P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
    means = _f32(means3D, "means3D", dev)
    colors = _f32(colors_precomp, "colors_precomp", dev)
    feats = _f32(language_feature_precomp, "language_feature_precomp", dev) if rs.include_feature else None
    opac = _f32(opacities, "opacities", dev)
    sc, rot, cov = _f32(scales, "scales", dev), _f32(rotations, "rotations", dev), _f32(cov3Ds_precomp, "cov3D", dev)
    if P > 0 and (colors is None or colors.shape != (P, NUM_CHANNELS)):
        raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")  # rasterizer_impl.cu:246-249
    if rs.include_feature and P > 0 and (feats is None or feats.shape != (P, NUM_CHANNELS_LANGUAGE_FEATURE)):
        raise RuntimeError(f"language_feature_precomp must be [{P}, {NUM_CHANNELS_LANGUAGE_FEATURE}]")
    view, proj, bg = _f32(rs.viewmatrix, "viewmatrix", dev), _f32(rs.projmatrix, "projmatrix", dev), _f32(rs.bg, "bg", dev)
    color = torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    alpha = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    feat_out = (torch.zeros((NUM_CHANNELS_LANGUAGE_FEATURE, H, W), dtype=torch.float32, device=dev)
                if rs.include_feature else torch.empty(0, dtype=torch.float32, device=dev))
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    num = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = L.load()
    # capacity of the instance buffers: 1.5 x what the previous call on this device produced (scenes rendered in a row
    # are alike), a generous guess the first time; an overflow repeats the call with the exact count
    last = _LAST_COUNT.get((dev.type, dev.index, H, W))
    cap = int(max_instances) if max_instances > 0 else (max(1 << 16, int(1.5 * last)) if last else max(1 << 20, 24 * P))
    for attempt in range(2):
        nbytes = lib.orvb_gs_workspace_bytes(P, cap, H, W)
        ws = _workspace(dev, nbytes)
        a = L.GsArgs(p=P, means3d=L.ptr(means), colors=L.ptr(colors), features=L.ptr(feats), opacities=L.ptr(opac),
                     scales=L.ptr(sc), rotations=L.ptr(rot), cov3d=L.ptr(cov), scale_modifier=float(rs.scale_modifier),
                     viewmatrix=view.data_ptr(), projmatrix=proj.data_ptr(), background=bg.data_ptr(),
                     tan_fovx=float(rs.tanfovx), tan_fovy=float(rs.tanfovy), height=H, width=W,
                     out_color=color.data_ptr(), out_feature=feat_out.data_ptr() if rs.include_feature else None,
                     out_depth=depth.data_ptr(), out_alpha=alpha.data_ptr(), radii=radii.data_ptr(),
                     num_rendered=num.data_ptr(), max_instances=cap, workspace=ws.data_ptr(), workspace_bytes=ws.numel())
        L.check(lib.orvb_gs_rasterize(C.byref(a), L.current_stream()), "orvb_gs_rasterize")
        n = int(num.item())  # one scalar read-back (the reference synchronises mid-call for the same number)
        if n >= 0:
            break
        if attempt == 1:
            raise RuntimeError(f"orvb_gs_rasterize: {-n} instances exceed the capacity {cap}")
        cap = -n
    rasterize_gaussians.last_num_rendered = n
    _LAST_COUNT[(dev.type, dev.index, H, W)] = n
    return color, feat_out, radii, depth, alpha


class GaussianRasterizer(nn.Module):
    """Reference __init__.py:195-237."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """checkFrustum (forward.cu:... / auxiliary.h:139-161): view-space depth > 0.01."""
        with torch.no_grad():
            V = self.raster_settings.viewmatrix.to(positions.device, torch.float32)
            z = positions.float() @ V[:3, 2] + V[3, 2]
            return z > 0.01

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, language_feature_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if (((scales is None or rotations is None) and cov3D_precomp is None)
                or ((scales is not None or rotations is not None) and cov3D_precomp is not None)):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, language_feature_precomp, opacities, scales,
                                   rotations, cov3D_precomp, self.raster_settings)


def focal2fov(focal: float, pixels: int) -> float:
    return 2 * math.atan(pixels / (2 * focal))


def get_projection_matrix_c(fx, fy, cx, cy, W, H, znear, zfar) -> torch.Tensor:
    """Projection from pinhole intrinsics with an off-centre principal point (orv/dataset/gs_render.py:203-221)."""
    top = cy * znear / fy
    bottom = -(H - cy) * znear / fy
    right = cx * znear / fx
    left = -(W - cx) * znear / fx
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def render(extrinsics, intrinsics, image_shape, pts_xyz, pts_rgb, feat, rotations, scales, opacity, bg_color):
    """`render` of orv/dataset/gs_render.py:103-171: camera set-up on the host, rasterisation on the B200."""
    dev = pts_xyz.device
    bg = torch.tensor(bg_color, dtype=torch.float32, device=dev)
    height, width = image_shape
    fx, fy = float(intrinsics[0][0]), float(intrinsics[1][1])
    cx, cy = float(intrinsics[0][2]), float(intrinsics[1][2])
    tan_fov_x = math.tan(focal2fov(fx, width) * 0.5)
    tan_fov_y = math.tan(focal2fov(fy, height) * 0.5)
    w2c = torch.inverse(extrinsics)
    projection = get_projection_matrix_c(fx, fy, cx, cy, width, height, 0.1, 200.0).transpose(0, 1).to(dev)
    world_view = w2c.transpose(0, 1).to(dev)
    full_projection = world_view.float() @ projection
    settings = GaussianRasterizationSettings(
        image_height=height, image_width=width, tanfovx=tan_fov_x, tanfovy=tan_fov_y, bg=bg, scale_modifier=1.0,
        viewmatrix=world_view, projmatrix=full_projection, sh_degree=3, campos=world_view.inverse()[3, :3],
        prefiltered=False, debug=False, include_feature=True)
    img, fmap, radii, depth, alpha = GaussianRasterizer(settings)(
        means3D=pts_xyz, means2D=torch.zeros_like(pts_xyz), shs=None, colors_precomp=pts_rgb,
        language_feature_precomp=feat, opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
    return {"render_color": img, "radii": radii, "render_depth": depth, "render_alpha": alpha, "render_feat": fmap}
