"""TEST INFRASTRUCTURE — numpy restatement (fp32) of the forward pass of the reference's Gaussian rasteriser
(orv/ops/diff-gaussian-rasterization), used only by tests/ and tools/.

Functions cite the reference lines they follow:
  preprocess      cuda_rasterizer/forward.cu:156-262 (in_frustum auxiliary.h:139-161, computeCov3D :118-154,
                  computeCov2D :74-116, getRect auxiliary.h:47-57, ndc2Pix :42-45)
  bin_instances   cuda_rasterizer/rasterizer_impl.cu:74-141, :283-318 (one instance per overlapped tile, sorted by the
                  64-bit key tile << 32 | depth bits; cub's radix sort is stable, so ties keep Gaussian-index order)
  render          cuda_rasterizer/forward.cu:267-398

Pinning: the reference extension builds here (oracle/build_ref.py --rasterizer -> oracle/_ref, nvcc, sm_100a) but needs
a GPU to run, so the pin happens on the GPU box: tests/test_zz_gpu_gs_render.py holds this restatement AND the CUDA
kernels to the reference's own output, and tests/golden/gs_render_small.pt is that output (written there by
tools/make_gs_golden.py, committed) for the CPU-only suite.
"""
from __future__ import annotations

import numpy as np

BLOCK_X = BLOCK_Y = 16  # config.h:16-17
f32 = np.float32


def _rect(px, py, radius, gx, gy):
    x0 = np.minimum(gx, np.maximum(0, ((px - radius.astype(f32)) / f32(BLOCK_X)).astype(np.int64)))
    y0 = np.minimum(gy, np.maximum(0, ((py - radius.astype(f32)) / f32(BLOCK_Y)).astype(np.int64)))
    x1 = np.minimum(gx, np.maximum(0, ((px + radius.astype(f32) + f32(BLOCK_X - 1)) / f32(BLOCK_X)).astype(np.int64)))
    y1 = np.minimum(gy, np.maximum(0, ((py + radius.astype(f32) + f32(BLOCK_Y - 1)) / f32(BLOCK_Y)).astype(np.int64)))
    return x0, y0, x1, y1


def preprocess(means, scales, rots, opac, view, proj, tan_fovx, tan_fovy, H, W, scale_mod=1.0):
    """Returns dict(radii int32 [P], xy [P,2], depth [P], conic_op [P,4], tiles [P]) — zeros for culled Gaussians."""
    means, scales, rots = means.astype(f32), scales.astype(f32), rots.astype(f32)
    V, M = view.astype(f32).reshape(16), proj.astype(f32).reshape(16)  # memory order of the tensors the caller passes
    P = means.shape[0]
    gx, gy = (W + BLOCK_X - 1) // BLOCK_X, (H + BLOCK_Y - 1) // BLOCK_Y
    focal_y, focal_x = f32(H) / (f32(2.0) * f32(tan_fovy)), f32(W) / (f32(2.0) * f32(tan_fovx))
    px, py, pz = means[:, 0], means[:, 1], means[:, 2]
    hx = M[0] * px + M[4] * py + M[8] * pz + M[12]
    hy = M[1] * px + M[5] * py + M[9] * pz + M[13]
    hw = M[3] * px + M[7] * py + M[11] * pz + M[15]
    p_w = f32(1.0) / (hw + f32(0.0000001))
    projx, projy = hx * p_w, hy * p_w
    tx = V[0] * px + V[4] * py + V[8] * pz + V[12]
    ty = V[1] * px + V[5] * py + V[9] * pz + V[13]
    tz = V[2] * px + V[6] * py + V[10] * pz + V[14]
    vis = tz > f32(0.01)
    tzs = np.where(vis, tz, f32(1.0))
    # computeCov3D: Sigma = R_std S^2 R_std^T with the quaternion (r, x, y, z) used as given
    s = f32(scale_mod) * scales
    r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    Rc = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),   # glm column 0
                   np.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),   # glm column 1
                   np.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1).astype(f32)
    Mc = Rc * s[:, None, :]                              # column c of M = S R: (sx R[c][0], sy R[c][1], sz R[c][2])
    Sigma = np.einsum("pik,pjk->pij", Mc, Mc).astype(f32)  # Sigma[i][j] = dot(column i, column j)
    # computeCov2D
    limx, limy = f32(1.3) * f32(tan_fovx), f32(1.3) * f32(tan_fovy)
    txc = np.minimum(limx, np.maximum(-limx, tx / tzs)) * tzs
    tyc = np.minimum(limy, np.maximum(-limy, ty / tzs)) * tzs
    j00, j02 = focal_x / tzs, -(focal_x * txc) / (tzs * tzs)
    j11, j12 = focal_y / tzs, -(focal_y * tyc) / (tzs * tzs)
    A = np.zeros((P, 2, 3), dtype=f32)
    for c in range(3):
        A[:, 0, c] = j00 * V[4 * c] + j02 * V[2 + 4 * c]
        A[:, 1, c] = j11 * V[1 + 4 * c] + j12 * V[2 + 4 * c]
    cov = np.einsum("pik,pkl,pjl->pij", A, Sigma, A).astype(f32)
    cxx, cxy, cyy = cov[:, 0, 0] + f32(0.3), cov[:, 0, 1], cov[:, 1, 1] + f32(0.3)
    det = cxx * cyy - cxy * cxy
    ok = vis & (det != 0)
    det_inv = f32(1.0) / np.where(det != 0, det, f32(1.0))
    mid = f32(0.5) * (cxx + cyy)
    root = np.sqrt(np.maximum(f32(0.1), mid * mid - det))
    radius = np.ceil(f32(3.0) * np.sqrt(np.maximum(mid + root, mid - root))).astype(f32)
    ix = (((projx.astype(np.float64) + 1.0) * W - 1.0) * 0.5).astype(f32)
    iy = (((projy.astype(np.float64) + 1.0) * H - 1.0) * 0.5).astype(f32)
    rad_i = np.where(ok, radius, 0).astype(np.int64)
    x0, y0, x1, y1 = _rect(ix, iy, rad_i, gx, gy)
    tiles = np.where(ok, (x1 - x0) * (y1 - y0), 0)
    ok = ok & (tiles > 0)
    out = dict(radii=np.where(ok, rad_i, 0).astype(np.int32), xy=np.stack([ix, iy], -1) * ok[:, None],
               depth=np.where(ok, tz, 0).astype(f32),
               conic_op=np.stack([cyy * det_inv, -cxy * det_inv, cxx * det_inv, opac.astype(f32).reshape(-1)], -1) * ok[:, None],
               tiles=np.where(ok, tiles, 0).astype(np.int64), rect=(x0, y0, x1, y1), grid=(gx, gy))
    return out


def bin_instances(pre):
    """Instance list sorted by (tile, depth bits, Gaussian index) and the per-tile [start, end) ranges."""
    gx, gy = pre["grid"]
    x0, y0, x1, y1 = pre["rect"]
    tile_ids, gids = [], []
    for g in np.nonzero(pre["tiles"])[0]:
        ys, xs = np.meshgrid(np.arange(y0[g], y1[g]), np.arange(x0[g], x1[g]), indexing="ij")
        t = (ys * gx + xs).reshape(-1)
        tile_ids.append(t)
        gids.append(np.full(t.shape, g, dtype=np.int64))
    if not tile_ids:
        return np.zeros(0, np.int64), np.zeros((gx * gy, 2), np.int64)
    tile_ids, gids = np.concatenate(tile_ids), np.concatenate(gids)
    dbits = pre["depth"].astype(f32).view(np.uint32)[gids].astype(np.int64)
    order = np.lexsort((gids, dbits, tile_ids))
    tile_ids, gids = tile_ids[order], gids[order]
    ranges = np.zeros((gx * gy, 2), np.int64)
    for t in np.unique(tile_ids):
        idx = np.nonzero(tile_ids == t)[0]
        ranges[t] = (idx[0], idx[-1] + 1)
    return gids, ranges


def render(pre, gids, ranges, colors, feats, bg, H, W):
    """Alpha blending per tile, front to back; returns (color [3,H,W], feat [12,H,W], depth [1,H,W], alpha [1,H,W])."""
    gx, gy = pre["grid"]
    colors, feats, bg = colors.astype(f32), feats.astype(f32), np.asarray(bg, dtype=f32)
    out_c, out_f = np.zeros((3, H, W), f32), np.zeros((feats.shape[1], H, W), f32)
    out_d, out_a = np.zeros((1, H, W), f32), np.zeros((1, H, W), f32)
    for ty in range(gy):
        for tx in range(gx):
            ys, xs = np.meshgrid(np.arange(ty * BLOCK_Y, min((ty + 1) * BLOCK_Y, H)),
                                 np.arange(tx * BLOCK_X, min((tx + 1) * BLOCK_X, W)), indexing="ij")
            pxf, pyf = xs.astype(f32), ys.astype(f32)
            T = np.ones(xs.shape, f32)
            done = np.zeros(xs.shape, bool)
            C = np.zeros((3,) + xs.shape, f32)
            F = np.zeros((feats.shape[1],) + xs.shape, f32)
            D = np.zeros(xs.shape, f32)
            s, e = ranges[ty * gx + tx]
            for g in gids[s:e]:
                if done.all():
                    break
                co = pre["conic_op"][g].astype(f32)
                dx, dy = pre["xy"][g, 0].astype(f32) - pxf, pre["xy"][g, 1].astype(f32) - pyf
                power = f32(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy
                alpha = np.minimum(f32(0.99), co[3] * np.exp(np.minimum(power, f32(0))).astype(f32))
                live = (~done) & (power <= 0) & (alpha >= f32(1.0 / 255.0))
                test_T = T * (f32(1) - alpha)
                stop = live & (test_T < f32(0.0001))
                done |= stop
                live &= ~stop
                w = np.where(live, alpha, f32(0))
                C += (colors[g][:, None, None] * w) * T
                D += (pre["depth"][g] * w) * T
                F += (feats[g][:, None, None] * w) * T
                T = np.where(live, test_T, T)
            out_c[:, ys, xs] = C + T * bg[:, None, None]
            out_f[:, ys, xs] = F
            out_d[0, ys, xs] = D
            out_a[0, ys, xs] = f32(1) - T
    return out_c, out_f, out_d, out_a


def rasterize(means, colors, feats, opac, scales, rots, view, proj, bg, tan_fovx, tan_fovy, H, W, scale_mod=1.0):
    pre = preprocess(means, scales, rots, opac, view, proj, tan_fovx, tan_fovy, H, W, scale_mod)
    gids, ranges = bin_instances(pre)
    c, f, d, a = render(pre, gids, ranges, colors, feats, bg, H, W)
    return dict(color=c, feat=f, depth=d, alpha=a, radii=pre["radii"], num_rendered=len(gids))


def synthetic_scene(P=400, H=48, W=80, seed=0):
    """Seeded scene in front of a pinhole camera looking down +z: Gaussians of mixed size, some behind the camera and
    some off-screen (culling paths), features one-hot-ish like the occupancy caller's semantic labels."""
    g = np.random.default_rng(seed)
    means = np.stack([g.uniform(-3, 3, P), g.uniform(-2, 2, P), g.uniform(-1.0, 9, P)], -1).astype(f32)
    scales = np.exp(g.uniform(np.log(0.02), np.log(0.4), (P, 3))).astype(f32)
    q = g.normal(size=(P, 4)).astype(f32)
    rots = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(f32)
    opac = g.uniform(0.05, 1.0, (P, 1)).astype(f32)
    colors = g.uniform(0, 1, (P, 3)).astype(f32)
    feats = np.zeros((P, 12), f32)
    feats[np.arange(P), g.integers(0, 12, P)] = 1.0
    feats += g.uniform(0, 0.05, (P, 12)).astype(f32)
    fx = fy = 0.9 * W
    cx, cy = W / 2 - 1.5, H / 2 + 0.75
    intr = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], f32)
    ang = 0.1
    c2w = np.eye(4, dtype=f32)
    c2w[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], f32)
    c2w[:3, 3] = [0.2, -0.1, -1.0]
    return dict(means=means, scales=scales, rots=rots, opac=opac, colors=colors, feats=feats, intrinsics=intr, c2w=c2w,
                H=H, W=W, bg=np.array([0.1, 0.2, 0.3], f32))


def camera(scene):
    """(viewmatrix, projmatrix, tan_fovx, tan_fovy) exactly as orv/dataset/gs_render.py:118-138 builds them (torch)."""
    import math

    import torch
    H, W, K = scene["H"], scene["W"], scene["intrinsics"]
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    tan_x = math.tan(2 * math.atan(W / (2 * fx)) * 0.5)
    tan_y = math.tan(2 * math.atan(H / (2 * fy)) * 0.5)
    znear, zfar = 0.1, 200.0
    top, bottom = cy * znear / fy, -(H - cy) * znear / fy
    right, left = cx * znear / fx, -(W - cx) * znear / fx
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    w2c = torch.inverse(torch.from_numpy(scene["c2w"]))
    view = w2c.transpose(0, 1).contiguous()
    full = (view.float() @ Pm.transpose(0, 1)).contiguous()
    return view.numpy(), full.numpy(), tan_x, tan_y
