"""CogVideoX DDIM / DPM-Solver++ schedulers (the two the reference pipeline accepts,
orv/models/cogvideox_control.py:1100, :1446-1457), with the per-step tensor arithmetic fused into one CUDA launch.

API mirrors diffusers' `CogVideoXDDIMScheduler` / `CogVideoXDPMScheduler` (`config`, `from_config`,
`set_timesteps`, `timesteps`, `order`, `init_noise_sigma`, `scale_model_input`, `step`).  Coefficients are computed
on the host in float64 exactly as diffusers does; `fused_step` then applies
CFG combine + v->x0 + update + bf16 cast to the resident latents through `orvb_sampler_step`.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np
import os

import torch

from . import _lib as L

_DEFAULTS = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, beta_schedule="scaled_linear",
    trained_betas=None, clip_sample=False, set_alpha_to_one=True, steps_offset=0, prediction_type="v_prediction",
    clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="trailing", rescale_betas_zero_snr=True,
    snr_shift_scale=3.0,
)


def rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    s = alphas_cumprod.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = s - sT
    s = s * (s0 / (s0 - sT))
    return s ** 2


class _CogVideoXSchedulerBase:
    order = 1

    def __init__(self, **kwargs):
        cfg = dict(_DEFAULTS)
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.trained_betas is not None:
            betas = torch.tensor(c.trained_betas, dtype=torch.float32)
        elif c.beta_schedule == "linear":
            betas = torch.linspace(c.beta_start, c.beta_end, c.num_train_timesteps, dtype=torch.float32)
        elif c.beta_schedule == "scaled_linear":
            betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps,
                                   dtype=torch.float64) ** 2
        else:
            raise NotImplementedError(f"{c.beta_schedule} is not implemented for {self.__class__}")
        self.betas = betas
        self.alphas = 1.0 - betas
        ac = torch.cumprod(self.alphas, dim=0)
        ac = ac / (c.snr_shift_scale + (1 - c.snr_shift_scale) * ac)
        if c.rescale_betas_zero_snr:
            ac = rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else ac[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c.num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        d = dict(vars(config)) if isinstance(config, SimpleNamespace) else dict(config)
        d.update(kwargs)
        return cls(**d)

    def scale_model_input(self, sample: torch.Tensor, timestep: Optional[int] = None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.config.train_timesteps`: {c.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ratio = c.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / num_inference_steps
            ts = np.round(np.arange(c.num_train_timesteps, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported.")
        self.timesteps_host = [int(t) for t in ts]  # the sampler loop reads these: no device -> host copy (= stream sync)
        t_cpu = torch.from_numpy(ts)
        if device is not None and torch.device(device).type == "cuda" and os.environ.get("ORVB_PINNED_UPLOADS", "1") != "0":
            self.timesteps = t_cpu.pin_memory().to(device, non_blocking=True)
        else:
            self.timesteps = t_cpu.to(device)

    # ---- float64 scalars -------------------------------------------------------------------------------
    def _alphas(self, timestep: int):
        prev = int(timestep) - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[int(timestep)].double()
        a_prev = (self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod).double()
        return a_t, a_prev, prev

    def _x0_coeffs(self, a_t) -> Tuple[float, float]:
        p = self.config.prediction_type
        if p == "v_prediction":
            return float(a_t ** 0.5), float(-((1 - a_t) ** 0.5))
        if p == "epsilon":
            return float(1.0 / a_t ** 0.5), float(-((1 - a_t) ** 0.5) / a_t ** 0.5)
        raise ValueError(f"prediction_type given as {p} must be one of `epsilon` or `v_prediction`")

    def _launch(self, model_out, latents, old_x0, noise, cfg_copies, guidance_scale, coeffs, next_input=None):
        if not (latents.is_cuda and latents.dtype == torch.bfloat16 and latents.is_contiguous()):
            raise RuntimeError("fused_step needs contiguous bf16 CUDA latents (the resident latents of the sampler)")
        if model_out.dtype != torch.bfloat16 or not model_out.is_contiguous():
            raise RuntimeError("fused_step needs the transformer's contiguous bf16 output")
        n = latents.numel()
        if model_out.numel() != cfg_copies * n:
            raise RuntimeError(f"model output has {model_out.numel()} elements, expected {cfg_copies}x{n}")
        a = L.SamplerStepArgs()
        a.model_out, a.latents = model_out.data_ptr(), latents.data_ptr()
        a.old_x0 = L.ptr(old_x0)
        a.noise = L.ptr(noise)
        a.n, a.cfg_copies, a.guidance_scale = n, cfg_copies, float(guidance_scale)
        a.c_x, a.c_v, a.d_cur, a.d_old, a.k_x, a.k_d, a.k_noise = [float(x) for x in coeffs]
        if next_input is not None:
            # [cfg*B, F, C_lat + C_img, h, w]: the new latents are also written into channels [0, C_lat)
            if not (next_input.is_cuda and next_input.dtype == torch.bfloat16 and next_input.is_contiguous()):
                raise RuntimeError("next_input must be a contiguous bf16 CUDA tensor")
            c_lat = latents.shape[2]
            if next_input.shape[0] != cfg_copies * latents.shape[0] or next_input.shape[2] <= c_lat:
                raise RuntimeError(f"next_input shape {tuple(next_input.shape)} does not match latents {tuple(latents.shape)}")
            a.next_input = next_input.data_ptr()
            a.lat_channels, a.img_channels = c_lat, next_input.shape[2] - c_lat
            a.hw = latents.shape[3] * latents.shape[4]
        L.check(L.load().orvb_sampler_step(C.byref(a), L.current_stream()), "orvb_sampler_step")


class CogVideoXDDIMScheduler(_CogVideoXSchedulerBase):
    def coefficients(self, timestep: int):
        a_t, a_prev, _ = self._alphas(timestep)
        c_x, c_v = self._x0_coeffs(a_t)
        a = ((1 - a_prev) / (1 - a_t)) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        return (c_x, c_v, 1.0, 0.0, float(a), float(b), 0.0)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """diffusers-compatible tensor step (fp32 torch ops on the tensors' device)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        a_t, a_prev, _ = self._alphas(int(timestep))
        x0 = (a_t ** 0.5) * sample - ((1 - a_t) ** 0.5) * model_output if self.config.prediction_type == "v_prediction" \
            else (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        a = ((1 - a_prev) / (1 - a_t)) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        prev = a * sample + b * x0
        if not return_dict:
            return (prev, x0)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def fused_step(self, model_out, timestep: int, latents, cfg_copies: int = 1, guidance_scale: float = 1.0,
                   next_input=None):
        self._launch(model_out, latents, None, None, cfg_copies, guidance_scale, self.coefficients(int(timestep)),
                     next_input)


class CogVideoXDPMScheduler(_CogVideoXSchedulerBase):
    def coefficients(self, timestep: int, timestep_back: Optional[int], have_old: bool):
        a_t, a_prev, prev = self._alphas(timestep)
        c_x, c_v = self._x0_coeffs(a_t)
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        m1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        m2 = (-2 * h).expm1() * a_prev ** 0.5
        m_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        first_order = (not have_old) or prev < 0
        if first_order:
            return (c_x, c_v, 1.0, 0.0, float(m1), float(-m2), float(m_noise)), True
        a_back = self.alphas_cumprod[int(timestep_back)].double()
        lamb_prev = ((a_back / (1 - a_back)) ** 0.5).log()
        r = (lamb - lamb_prev) / h
        m3, m4 = 1 + 1 / (2 * r), 1 / (2 * r)
        return (c_x, c_v, float(m3), float(-m4), float(m1), float(-m2), float(m_noise)), False

    def noise_draws(self, num_steps: int) -> List[int]:
        """How many `randn_tensor` calls diffusers' step makes at each loop index: one always, a second one on
        the second-order branch (the first draw is then discarded) — needed to keep the generator stream aligned."""
        ts = getattr(self, "timesteps_host", None) or self.timesteps.tolist()
        out = []
        for i, t in enumerate(ts):
            prev = t - self.config.num_train_timesteps // self.num_inference_steps
            out.append(1 if (i == 0 or prev < 0) else 2)
        return out

    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = False):
        """diffusers-compatible tensor step (torch ops)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        (c_x, c_v, d_cur, d_old, k_x, k_d, k_n), first = self.coefficients(
            int(timestep), None if timestep_back is None else int(timestep_back), old_pred_original_sample is not None)
        a_t, _, _ = self._alphas(int(timestep))
        x0 = (a_t ** 0.5) * sample - ((1 - a_t) ** 0.5) * model_output
        noise = _randn(sample.shape, generator, sample.device, sample.dtype)
        if first:
            prev = k_x * sample + k_d * x0 + k_n * noise
        else:
            d = d_cur * x0 + d_old * old_pred_original_sample
            noise = _randn(sample.shape, generator, sample.device, sample.dtype)
            prev = k_x * sample + k_d * d + k_n * noise
        if not return_dict:
            return (prev, x0)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def fused_step(self, model_out, old_x0, have_old: bool, timestep: int, timestep_back: Optional[int], latents,
                   noise, cfg_copies: int = 1, guidance_scale: float = 1.0, next_input=None):
        coeffs, _ = self.coefficients(int(timestep), timestep_back, have_old)
        self._launch(model_out, latents, old_x0, noise, cfg_copies, guidance_scale, coeffs, next_input)


def _randn(shape, generator, device, dtype):
    """diffusers randn_tensor: a CPU generator samples on the CPU, then the tensor moves to `device`."""
    if isinstance(generator, list):
        generator = generator[0]
    gen_dev = generator.device.type if generator is not None else torch.device(device).type
    if gen_dev == "cpu" and torch.device(device).type != "cpu":
        # Same CPU draw; the upload goes through PINNED memory and does not block.  A pageable cudaMemcpyAsync
        # synchronises the stream before it starts, i.e. the host would wait here for everything still queued — the tail of
        # the previous clip — and the GPU would then idle through the rest of this clip's host prologue.
        if torch.device(device).type == "cuda" and os.environ.get("ORVB_PINNED_UPLOADS", "1") != "0":
            return torch.randn(shape, generator=generator, device="cpu", dtype=dtype, pin_memory=True).to(device, non_blocking=True)
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


randn_tensor = _randn
