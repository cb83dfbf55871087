"""Shared helpers of the Gaussian-rasteriser tests: scene -> torch tensors, and the call into the reference build."""
import numpy as np
import torch

from oracle import gs_oracle as G


def scene_tensors(scene, device):
    view, proj, tan_x, tan_y = G.camera(scene)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
    return dict(means=t(scene["means"]), colors=t(scene["colors"]), feats=t(scene["feats"]), opac=t(scene["opac"]),
                scales=t(scene["scales"]), rots=t(scene["rots"]), view=t(view), proj=t(proj), bg=t(scene["bg"]),
                tan_x=tan_x, tan_y=tan_y, H=scene["H"], W=scene["W"])


def run_reference(mod, s):
    """`_C.rasterize_gaussians` of the reference extension (rasterize_points.cu:35-150), argument order of
    diff_gaussian_rasterization/__init__.py:52-74."""
    dev = s["means"].device
    empty = torch.empty(0, device=dev)
    campos = torch.inverse(s["view"])[3, :3].contiguous()
    n, color, feat, depth, alpha, radii, *_ = mod.rasterize_gaussians(
        s["bg"], s["means"], s["colors"], s["feats"], s["opac"], s["scales"], s["rots"], 1.0, empty, s["view"], s["proj"],
        s["tan_x"], s["tan_y"], s["H"], s["W"], empty, 3, campos, False, False, True)
    return dict(num_rendered=int(n), color=color, feat=feat, depth=depth, alpha=alpha, radii=radii)


def run_ours(s, max_instances=0):
    from orv_b200.gs_render import GaussianRasterizationSettings, rasterize_gaussians
    rs = GaussianRasterizationSettings(image_height=s["H"], image_width=s["W"], tanfovx=s["tan_x"], tanfovy=s["tan_y"],
                                       bg=s["bg"], scale_modifier=1.0, viewmatrix=s["view"], projmatrix=s["proj"],
                                       sh_degree=3, campos=torch.inverse(s["view"])[3, :3], prefiltered=False, debug=False,
                                       include_feature=True)
    color, feat, radii, depth, alpha = rasterize_gaussians(s["means"], torch.zeros_like(s["means"]), None, s["colors"],
                                                          s["feats"], s["opac"], s["scales"], s["rots"], None, rs,
                                                          max_instances=max_instances)
    return dict(num_rendered=rasterize_gaussians.last_num_rendered, color=color, feat=feat, depth=depth, alpha=alpha,
                radii=radii)


def occupancy_scene(P=200000, H=320, W=480, seed=0):
    """An occupancy-like workload: voxels of a 0.1 m grid rendered as small isotropic Gaussians, 12 one-hot semantic
    channels, opacity 1 — what orv/dataset/gs_render.py feeds the rasteriser (sizes of the synthetic stand-in stated in
    the test / bench that uses it)."""
    g = np.random.default_rng(seed)
    # points on a few surfaces (floor, table top, back wall) + scattered objects, in camera-forward space
    n1, n2, n3 = P // 3, P // 3, P - 2 * (P // 3)
    floor = np.stack([g.uniform(-2, 2, n1), np.full(n1, 0.8), g.uniform(0.5, 5, n1)], -1)
    wall = np.stack([g.uniform(-2, 2, n2), g.uniform(-1.5, 0.8, n2), np.full(n2, 5.0)], -1)
    objs = np.stack([g.uniform(-1, 1, n3), g.uniform(-0.5, 0.8, n3), g.uniform(1.0, 3.0, n3)], -1)
    means = (np.round(np.concatenate([floor, wall, objs]) / 0.02) * 0.02).astype(np.float32)
    scene = G.synthetic_scene(P=8, H=H, W=W, seed=seed)
    Pn = means.shape[0]
    feats = np.zeros((Pn, 12), np.float32)
    feats[np.arange(Pn), g.integers(0, 12, Pn)] = 1.0
    scene.update(means=means, scales=np.full((Pn, 3), 0.012, np.float32),
                 rots=np.tile(np.array([1, 0, 0, 0], np.float32), (Pn, 1)), opac=np.ones((Pn, 1), np.float32),
                 colors=g.uniform(0, 1, (Pn, 3)).astype(np.float32), feats=feats)
    return scene
