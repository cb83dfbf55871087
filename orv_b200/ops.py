"""Torch-tensor front ends of the granular C-ABI operators (used by the parity tests and by bench.py).

Each function validates dtypes/devices, passes raw device pointers through ctypes and returns torch tensors.
None of them falls back to PyTorch arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (liborv_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def rowmap(seq_len=0, text_len=0, tokens_per_group=1, groups_per_batch=1) -> L.RowMap:
    return L.RowMap(seq_len, text_len, tokens_per_group, groups_per_batch)


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = L.EPI_BIAS,
         out: Optional[torch.Tensor] = None, ldo: Optional[int] = None, out_col_offset: int = 0,
         row_remap=(0, 0, 0), resid: Optional[torch.Tensor] = None, resid_mod: int = 0, resid_views: int = 1,
         resid_view_stride: int = 0, gate: Optional[torch.Tensor] = None, gate_text_off: int = 0,
         gate_video_off: int = 0, rm: Optional[L.RowMap] = None, qk_dim: int = 0, q_norm=None, k_norm=None,
         qk_eps: float = 1e-6, rope=None, bn: int = 0, launch: bool = True, out_f32: bool = False, k_wrap: int = 0):
    """out = epilogue(a @ w.T)  — a [M,K] bf16, w [N,K] bf16 (nn.Linear layout).  `launch=False` only fills and
    returns `(GemmArgs, out)`."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == (k_wrap or K)  # k_wrap: W [N, k_wrap] is walked cyclically along K (orvb_gemm_args.k_wrap)
    odt = torch.float32 if out_f32 else torch.bfloat16  # out_f32: tight-tolerance test mode (fp32 before the rounding)
    if out is None:
        rows = M if row_remap[0] == 0 else (M // row_remap[0]) * row_remap[1]
        out = torch.empty((rows, N), dtype=odt, device=a.device)
    _req(out, odt, "out")
    args = L.GemmArgs()
    args.out_f32 = int(out_f32)
    args.k_wrap = int(k_wrap)
    args.a, args.w = a.data_ptr(), w.data_ptr()
    args.out = out.data_ptr() + out_col_offset * out.element_size()
    args.bias = L.ptr(bias)
    args.m, args.n, args.k = M, N, K
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.ldo = ldo if ldo is not None else out.stride(0)
    args.epilogue = epilogue
    args.src_rows, args.dst_rows, args.dst_offset = row_remap
    if resid is not None:
        _req(resid, torch.bfloat16, "resid")
        args.resid, args.ldr = resid.data_ptr(), resid.stride(-2)
    args.resid_mod, args.resid_views, args.resid_view_stride = resid_mod, resid_views, resid_view_stride
    if gate is not None:
        _req(gate, torch.float32, "gate")
        args.gate, args.gate_ld = gate.data_ptr(), gate.stride(-2)
    args.gate_text_off, args.gate_video_off = gate_text_off, gate_video_off
    args.rowmap = rm if rm is not None else rowmap()
    args.qk_dim = qk_dim
    if q_norm is not None:
        args.q_norm_w, args.q_norm_b = q_norm[0].data_ptr(), q_norm[1].data_ptr()
        args.k_norm_w, args.k_norm_b = k_norm[0].data_ptr(), k_norm[1].data_ptr()
    args.qk_eps = qk_eps
    if rope is not None:
        _req(rope[0], torch.float32, "rope_cos")
        _req(rope[1], torch.float32, "rope_sin")
        args.rope_cos, args.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
    if not launch:
        return args, out
    lib = L.load()
    if bn:
        L.check(lib.orvb_gemm_bf16_bn(C.byref(args), bn, L.current_stream()), "orvb_gemm_bf16")
    else:
        L.check(lib.orvb_gemm_bf16(C.byref(args), L.current_stream()), "orvb_gemm_bf16")
    return out


def attention(qkv: torch.Tensor, batch: int, seq_len: int, heads: int, scale: float,
              out: Optional[torch.Tensor] = None, q_row0: int = 0, q_rows: int = 0, out_f32: bool = False) -> torch.Tensor:
    """Non-causal attention over the packed Q | K | V buffer; `q_row0/q_rows` select a query window (compact output);
    `out_f32` is the tight-tolerance test mode (fp32 output before the bf16 rounding)."""
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.shape == (batch * seq_len, 3 * heads * 64)
    nq = q_rows if q_rows > 0 else seq_len
    odt = torch.float32 if out_f32 else torch.bfloat16
    if out is None:
        out = torch.empty((batch * nq, heads * 64), dtype=odt, device=qkv.device)
    _req(out, odt, "out")
    a = L.AttentionArgs(qkv=qkv.data_ptr(), out=out.data_ptr(), batch=batch, seq_len=seq_len, heads=heads, scale=scale,
                        q_row0=q_row0, q_rows=q_rows, out_f32=int(out_f32))
    L.check(L.load().orvb_attention(C.byref(a), L.current_stream()), "orvb_attention")
    return out


def ln_modulate(x: torch.Tensor, ln_w, ln_b, eps: float, mod: Optional[torch.Tensor] = None, text_off: int = 0,
                video_off: int = 0, scale_first: bool = False, rm: Optional[L.RowMap] = None,
                in_video_only: bool = False, out: Optional[torch.Tensor] = None,
                ab: Optional[torch.Tensor] = None, out_f32: bool = False) -> torch.Tensor:
    """LayerNorm + AdaLN modulate.  `ab` (bf16 [groups, 4*dim] = text A | text B | video A | video B per group) is the
    folded form the forward uses: y = xhat * A + B with A = w (1 + scale), B = b (1 + scale) + shift."""
    _req(x, torch.bfloat16, "x")
    rows_in, dim = x.shape
    rmap = rm if rm is not None else rowmap()
    rows = rows_in
    if in_video_only:
        nb = rows_in // rmap.seq_len
        rows = nb * (rmap.seq_len - rmap.text_len)
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    a = L.LnArgs()
    a.y_f32 = int(out_f32)
    a.x, a.y = x.data_ptr(), out.data_ptr()
    a.ln_w, a.ln_b = L.ptr(ln_w), L.ptr(ln_b)
    a.rows, a.dim, a.eps = rows, dim, eps
    if mod is not None:
        _req(mod, torch.float32, "mod")
        a.mod, a.mod_ld = mod.data_ptr(), mod.stride(-2)
    a.text_off, a.video_off, a.scale_first = text_off, video_off, int(scale_first)
    a.rowmap = rmap
    a.in_video_only = int(in_video_only)
    if ab is not None:
        _req(ab, torch.bfloat16, "ab")
        a.ab, a.ab_ld = ab.data_ptr(), ab.stride(-2)
    L.check(L.load().orvb_ln_modulate(C.byref(a), L.current_stream()), "orvb_ln_modulate")
    return out


def skinny_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: int = 0) -> torch.Tensor:
    _req(x, torch.float32, "x")
    _req(w, torch.bfloat16, "w")
    rows, k = x.shape
    n = w.shape[0]
    y = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    L.check(L.load().orvb_skinny_linear(x.data_ptr(), w.data_ptr(), L.ptr(b), y.data_ptr(), rows, n, k, act,
                                        L.current_stream()), "orvb_skinny_linear")
    return y


def patchify(x: torch.Tensor, p: int, patch_t: int = 0) -> torch.Tensor:
    _req(x, torch.bfloat16, "x")
    B, F, Cc, H, W = x.shape
    pt = max(patch_t, 1)
    out = torch.empty((B * (F // pt) * (H // p) * (W // p), Cc * pt * p * p), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().orvb_patchify(x.data_ptr(), out.data_ptr(), B, F, Cc, H, W, p, patch_t, L.current_stream()),
            "orvb_patchify")
    return out


def unpatchify(y: torch.Tensor, B: int, F: int, Cc: int, H: int, W: int, p: int, patch_t: int = 0) -> torch.Tensor:
    _req(y, torch.bfloat16, "y")
    out = torch.empty((B, F, Cc, H, W), dtype=torch.bfloat16, device=y.device)
    L.check(L.load().orvb_unpatchify(y.data_ptr(), out.data_ptr(), B, F, Cc, H, W, p, patch_t, L.current_stream()),
            "orvb_unpatchify")
    return out


# ---------------------------------------------------------------------------------------------------------------
# 3-D VAE decode operators (channels-last bf16 [T, H, W, C], one sample) — include/orv_b200.h, "3-D VAE decode"
# ---------------------------------------------------------------------------------------------------------------
_conv_gn_scratch: dict = {}


def conv_cl(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], kernel, *, cache: Optional[torch.Tensor] = None,
            resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_f32: bool = False,
            gn: Optional[tuple] = None):
    """Causal convolution (implicit GEMM).  x [T, H, W, c_in] bf16, w [c_out, kt*kh*kw*c_in] bf16 (tap-major),
    kernel = (kt, kh, kw); cache [kt-1, H, W, c_in] = the frames in front of x (None: frame 0 repeated).
    gn = (groups, eps): also returns the GroupNorm statistics [groups, 2] of the output, accumulated in the epilogue
    (result = (out, stats))."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    T, H, W, cin = x.shape
    kt, kh, kw = kernel
    cout = w.shape[0]
    if w.shape[1] != kt * kh * kw * cin:
        raise RuntimeError(f"conv_cl: weight [{tuple(w.shape)}] does not match {kt}x{kh}x{kw}x{cin}")
    odt = torch.float32 if out_f32 else torch.bfloat16
    if out is None:
        out = torch.empty((T, H, W, cout), dtype=odt, device=x.device)
    _req(out, odt, "out")
    a = L.ConvArgs()
    a.x, a.w, a.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
    if cache is not None:
        _req(cache, torch.bfloat16, "cache")
        if tuple(cache.shape) != (kt - 1, H, W, cin):
            raise RuntimeError(f"conv_cl: cache {tuple(cache.shape)} must be {(kt - 1, H, W, cin)}")
        a.cache = cache.data_ptr()
    if bias is not None:
        _req(bias, torch.bfloat16, "bias")
        a.bias = bias.data_ptr()
    if resid is not None:
        _req(resid, torch.bfloat16, "resid")
        a.resid = resid.data_ptr()
    a.frames, a.height, a.width, a.c_in, a.c_out = T, H, W, cin, cout
    a.kt, a.kh, a.kw = kt, kh, kw
    a.out_f32 = int(out_f32)
    lib = L.load()
    stats = None
    if gn is not None:
        key = (x.device, torch.cuda.current_stream().cuda_stream)
        scratch = _conv_gn_scratch.get(key)
        if scratch is None:
            scratch = torch.empty(lib.orvb_conv_gn_scratch_bytes() + 16, dtype=torch.uint8, device=x.device)
            _conv_gn_scratch[key] = scratch
        stats = torch.empty((gn[0], 2), dtype=torch.float32, device=x.device)
        a.gn_stats, a.gn_scratch = stats.data_ptr(), (scratch.data_ptr() + 15) // 16 * 16
        a.gn_groups, a.gn_eps = int(gn[0]), float(gn[1])
    L.check(lib.orvb_conv_cl(C.byref(a), L.current_stream()), "orvb_conv_cl")
    return out if gn is None else (out, stats)


_gn_scratch: dict = {}


def _gn_scratch_for(device, nbytes: int) -> torch.Tensor:
    key = (device, torch.cuda.current_stream().cuda_stream)
    buf = _gn_scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _gn_scratch[key] = buf
    return buf


def gn_stats_cl(x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """(mean, rstd) fp32 [groups, 2] of a channels-last bf16 tensor [..., C] over all its pixels."""
    _req(x, torch.bfloat16, "x")
    Cc = x.shape[-1]
    pixels = x.numel() // Cc
    lib = L.load()
    scratch = _gn_scratch_for(x.device, lib.orvb_gn_scratch_bytes(pixels, groups))
    stats = torch.empty((groups, 2), dtype=torch.float32, device=x.device)
    L.check(lib.orvb_gn_stats_cl(x.data_ptr(), pixels, Cc, groups, eps, stats.data_ptr(), scratch.data_ptr(),
                                 L.current_stream()), "orvb_gn_stats_cl")
    return stats


def spatial_norm_cl(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, table: torch.Tensor,
                    y_off: int, b_off: int, t_src: torch.Tensor, lat_hw, shift: int, *, groups: int = 32, act: int = 1,
                    out: Optional[torch.Tensor] = None, y_f32: bool = False) -> torch.Tensor:
    """act(GroupNorm(x) * table[src, y_off:] + table[src, b_off:]) — see orvb_spatial_norm_args."""
    _req(x, torch.bfloat16, "x")
    _req(stats, torch.float32, "stats")
    _req(gamma, torch.bfloat16, "gamma")
    _req(beta, torch.bfloat16, "beta")
    _req(table, torch.bfloat16, "table")
    _req(t_src, torch.int32, "t_src")
    T, H, W, Cc = x.shape
    if t_src.numel() != T:
        raise RuntimeError("spatial_norm_cl: t_src must have one entry per frame")
    odt = torch.float32 if y_f32 else torch.bfloat16
    if out is None:
        out = torch.empty((T, H, W, Cc), dtype=odt, device=x.device)
    _req(out, odt, "out")
    a = L.SpatialNormArgs()
    a.x, a.y = x.data_ptr(), out.data_ptr()
    a.frames, a.height, a.width, a.channels, a.groups = T, H, W, Cc, groups
    a.stats, a.gamma, a.beta, a.table = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), table.data_ptr()
    a.table_ld, a.y_off, a.b_off = table.stride(0), y_off, b_off
    a.t_src = t_src.data_ptr()
    a.lat_h, a.lat_w = lat_hw
    a.shift, a.act, a.y_f32 = shift, act, int(y_f32)
    L.check(L.load().orvb_spatial_norm_cl(C.byref(a), L.current_stream()), "orvb_spatial_norm_cl")
    return out


def upsample2x_cl(x: torch.Tensor, t_src: torch.Tensor) -> torch.Tensor:
    """out[t, h, w] = x[t_src[t], h // 2, w // 2]  (channels-last bf16)."""
    _req(x, torch.bfloat16, "x")
    _req(t_src, torch.int32, "t_src")
    _, H, W, Cc = x.shape
    To = t_src.numel()
    out = torch.empty((To, 2 * H, 2 * W, Cc), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().orvb_upsample2x_cl(x.data_ptr(), out.data_ptr(), To, H, W, Cc, t_src.data_ptr(), L.current_stream()),
            "orvb_upsample2x_cl")
    return out


def cl_to_planar(x: torch.Tensor, c_keep: int) -> torch.Tensor:
    """[T, H, W, c_ld] -> [c_keep, T, H, W]."""
    _req(x, torch.bfloat16, "x")
    T, H, W, cl = x.shape
    out = torch.empty((c_keep, T, H, W), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().orvb_cl_to_planar(x.data_ptr(), out.data_ptr(), T * H * W, cl, c_keep, L.current_stream()),
            "orvb_cl_to_planar")
    return out
