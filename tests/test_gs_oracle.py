"""CPU: the numpy restatement of the Gaussian rasteriser (oracle/gs_oracle.py) against tests/golden/gs_render_small.pt —
the REFERENCE extension's own output for the same seeded scene (oracle/_ref build of
orv/ops/diff-gaussian-rasterization, run on a B200 by tools/make_gs_golden.py) — and against closed forms."""
import os

import numpy as np
import torch

from oracle import gs_oracle as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gs_render_small.pt")


def _run(scene):
    view, proj, tx, ty = G.camera(scene)
    return G.rasterize(scene["means"], scene["colors"], scene["feats"], scene["opac"], scene["scales"], scene["rots"], view,
                       proj, scene["bg"], tx, ty, scene["H"], scene["W"])


def test_oracle_matches_the_reference_extensions_output():
    ref = torch.load(GOLDEN, weights_only=False)
    out = _run(G.synthetic_scene(P=400, H=48, W=80, seed=0))
    assert out["num_rendered"] == ref["num_rendered"] == 960
    assert np.array_equal(out["radii"], ref["radii"].numpy())          # integer outputs: exact
    for k in ("color", "feat", "depth", "alpha"):
        np.testing.assert_allclose(out[k], ref[k].numpy(), rtol=1e-3, atol=2e-5, err_msg=k)  # fp32: north-star rtol


def test_single_gaussian_closed_form():
    """One isotropic Gaussian on the optical axis: alpha at its centre pixel = min(0.99, opacity) (forward.cu:352),
    depth = alpha * z, colour = alpha * c + (1 - alpha) * bg (:362-386), radius = ceil(3 sigma_px)."""
    scene = G.synthetic_scene(P=1, H=33, W=33, seed=0)
    scene["c2w"] = np.eye(4, dtype=np.float32)
    scene["intrinsics"] = np.array([[40.0, 0, 16.5], [0, 40.0, 16.5], [0, 0, 1]], np.float32)  # pixel (16, 16) on the axis
    z, s, op = 4.0, 0.2, 0.6
    scene.update(means=np.array([[0, 0, z]], np.float32), scales=np.full((1, 3), s, np.float32),
                 rots=np.array([[1, 0, 0, 0]], np.float32), opac=np.array([[op]], np.float32),
                 colors=np.array([[1.0, 0.5, 0.25]], np.float32))
    out = _run(scene)
    sigma2 = (40.0 * s / z) ** 2 + 0.3
    assert out["radii"][0] == int(np.ceil(3 * np.sqrt(sigma2)))
    a = out["alpha"][0, 16, 16]
    assert abs(a - op) < 1e-6
    assert abs(out["depth"][0, 16, 16] - op * z) < 1e-5
    np.testing.assert_allclose(out["color"][:, 16, 16], op * scene["colors"][0] + (1 - op) * scene["bg"], atol=1e-6)
    # Gaussian falloff two pixels off the centre
    assert abs(out["alpha"][0, 16, 18] - op * np.exp(-0.5 * 4 / sigma2)) < 1e-5


def test_opaque_front_gaussian_hides_the_ones_behind_it():
    scene = G.synthetic_scene(P=2, H=33, W=33, seed=0)
    scene["c2w"] = np.eye(4, dtype=np.float32)
    scene["intrinsics"] = np.array([[40.0, 0, 16.5], [0, 40.0, 16.5], [0, 0, 1]], np.float32)
    scene.update(means=np.array([[0, 0, 6.0], [0, 0, 3.0]], np.float32), scales=np.full((2, 3), 0.3, np.float32),
                 rots=np.tile(np.array([1, 0, 0, 0], np.float32), (2, 1)), opac=np.ones((2, 1), np.float32),
                 colors=np.array([[1, 0, 0], [0, 1, 0]], np.float32))
    out = _run(scene)
    # front (green): alpha clamps to 0.99, T = 0.01; the back one would leave T = 1e-4 - eps < 1e-4, so the pixel is
    # declared done BEFORE blending it (forward.cu:355-360): red contributes nothing, the background weighs T = 0.01
    c = out["color"][:, 16, 16]
    np.testing.assert_allclose(c, np.array([0.0, 0.99, 0.0]) + 0.01 * scene["bg"], atol=1e-5)
    assert abs(out["alpha"][0, 16, 16] - 0.99) < 1e-6 and abs(out["depth"][0, 16, 16] - 0.99 * 3.0) < 1e-5
