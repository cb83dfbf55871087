"""`CogVideoXImageToVideoPipelineTraj` — the sampler of reference orv/models/cogvideox_control.py:1090-1489 with the
same constructor, attributes and `__call__` keyword arguments, driving the B200 transformer.

What stays on the host: argument checks, integer shape bookkeeping, the CPU-generator RNG contract
(`randn_tensor`, `DiagonalGaussianDistribution.sample`) and the float64 scheduler coefficients.  What runs on the
GPU: the transformer forward (orvb_forward) and one fused launch per step for CFG combine + scheduler update +
bf16 cast (orvb_sampler_step).  The VAE and T5 encoder are out of scope: any object
exposing the diffusers attributes the reference reads (`vae.config.*`, `vae.decode`, `text_encoder`, `tokenizer`)
can be plugged in; with `prompt_embeds` and `output_type="latent"` none of them is executed.
"""
from __future__ import annotations

import inspect
import math
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from ..schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler, randn_tensor
from .cogvideox_control import CogVideoXTransformer3DModelTraj


@dataclass
class CogVideoXPipelineOutput:
    frames: torch.Tensor


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        sample = randn_tensor(self.mean.shape, generator, self.parameters.device, self.parameters.dtype)
        return self.mean + self.std * sample


def default_vae_config(scaling_factor: float = 1.15258426, invert_scale_latents: bool = False) -> SimpleNamespace:
    """The attributes of AutoencoderKLCogVideoX.config the sampler reads (THUDM/CogVideoX-2b values)."""
    return SimpleNamespace(config=SimpleNamespace(
        block_out_channels=(128, 256, 256, 512), temporal_compression_ratio=4, scaling_factor=scaling_factor,
        latent_channels=16, invert_scale_latents=invert_scale_latents))


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, **kwargs):
    if timesteps is not None:
        raise ValueError("custom `timesteps` are not supported by the CogVideoX schedulers")
    scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, num_inference_steps


class CogVideoXImageToVideoPipelineTraj:
    transformer: CogVideoXTransformer3DModelTraj

    def __init__(self, tokenizer, text_encoder, vae, transformer: CogVideoXTransformer3DModelTraj,
                 scheduler: Union[CogVideoXDDIMScheduler, CogVideoXDPMScheduler]):
        self.tokenizer, self.text_encoder, self.vae = tokenizer, text_encoder, vae
        self.transformer, self.scheduler = transformer, scheduler
        if not isinstance(self.transformer, CogVideoXTransformer3DModelTraj):
            raise ValueError("The transformer in this pipeline must be of type CogVideoXTransformer3DModelTraj")
        vc = getattr(vae, "config", None)
        self.vae_scale_factor_spatial = 2 ** (len(vc.block_out_channels) - 1) if vc is not None else 8
        self.vae_scale_factor_temporal = vc.temporal_compression_ratio if vc is not None else 4
        self.vae_scaling_factor_image = vc.scaling_factor if vc is not None else 0.7
        self._guidance_scale = 1.0
        self._interrupt = False
        self._num_timesteps = 0
        self.last_step_launches = 0
        self._staging: Dict[Tuple, torch.Tensor] = {}

    # ---- diffusers DiffusionPipeline surface the reference programs touch ----
    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def _execution_device(self):
        return self.transformer.device

    def to(self, *args, **kwargs):
        self.transformer.to(*args, **kwargs)
        for m in (self.vae, self.text_encoder):
            if hasattr(m, "to"):
                m.to(*args, **kwargs)
        return self

    def maybe_free_model_hooks(self):
        pass

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm
        cfg = getattr(self, "_progress_bar_config", {"disable": True})
        return tqdm(iterable, **cfg) if iterable is not None else tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def check_inputs(self, image, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs,
                     latents=None, prompt_embeds=None, negative_prompt_embeds=None):
        """diffusers CogVideoXImageToVideoPipeline.check_inputs.  NB the reference calls it positionally
        (:1261-1270), which shifts `prompt_embeds` into the `latents` slot — reproduced by keeping this signature."""
        if not isinstance(image, (torch.Tensor, list)) and not hasattr(image, "size"):
            raise ValueError(f"`image` has to be of type `torch.Tensor` or `PIL.Image.Image` or `List[PIL.Image.Image]` "
                             f"but is {type(image)}")
        if (height is not None and height % 8 != 0) or (width is not None and width % 8 != 0):
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make "
                             "sure to only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and "
                             "`prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True,
                      num_videos_per_prompt=1, prompt_embeds=None, negative_prompt_embeds=None,
                      max_sequence_length=226, device=None, dtype=None):
        if prompt_embeds is None or (do_classifier_free_guidance and negative_prompt_embeds is None):
            if self.text_encoder is None or self.tokenizer is None:
                raise RuntimeError("encode_prompt needs a T5 text_encoder/tokenizer; pass `prompt_embeds` "
                                   "(and `negative_prompt_embeds` for CFG) instead — the evaluation path uses the "
                                   "cached empty-prompt embedding (reference dataset.py:1056-1059)")

        def t5(texts):
            texts = [texts] if isinstance(texts, str) else texts
            tok = self.tokenizer(texts, padding="max_length", max_length=max_sequence_length, truncation=True,
                                 add_special_tokens=True, return_tensors="pt")
            emb = self.text_encoder(tok.input_ids.to(device))[0].to(dtype=dtype or self.text_encoder.dtype, device=device)
            _, s, _ = emb.shape
            return emb.repeat(1, num_videos_per_prompt, 1).view(len(texts) * num_videos_per_prompt, s, -1)

        if prompt_embeds is None:
            prompt_embeds = t5(prompt)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            n = prompt_embeds.shape[0] // num_videos_per_prompt
            neg = negative_prompt or ""
            neg = n * [neg] if isinstance(neg, str) else neg
            negative_prompt_embeds = t5(neg)
        return prompt_embeds, negative_prompt_embeds

    def prepare_extra_step_kwargs(self, generator, eta):
        kw = {}
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def _prepare_rotary_positional_embeddings(self, height: int, width: int, num_frames: int, device):
        """diffusers I2V pipeline helper."""
        from .embeddings import get_3d_rotary_pos_embed, get_resize_crop_region_for_grid
        c = self.transformer.config
        gh = height // (self.vae_scale_factor_spatial * c.patch_size)
        gw = width // (self.vae_scale_factor_spatial * c.patch_size)
        p_t = c.patch_size_t
        base_w, base_h = c.sample_width // c.patch_size, c.sample_height // c.patch_size
        if p_t is None:
            crops = get_resize_crop_region_for_grid((gh, gw), base_w, base_h)
            cos, sin = get_3d_rotary_pos_embed(c.attention_head_dim, crops, (gh, gw), num_frames, device=device)
        else:
            base_frames = (num_frames + p_t - 1) // p_t
            cos, sin = get_3d_rotary_pos_embed(c.attention_head_dim, None, (gh, gw), base_frames, grid_type="slice",
                                               max_size=(base_h, base_w), device=device)
        return cos, sin

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """diffusers CogVideoXImageToVideoPipeline.decode_latents (called at reference :1478): [B, F, C, h, w] latents ->
        [B, 3, 4(F-1)+1, 8h, 8w] frames through `vae.decode` — `orv_b200.AutoencoderKLCogVideoX` runs it on the B200
        kernels; a config-only stand-in has no decoder and raises."""
        if self.vae is None or not hasattr(self.vae, "decode"):
            raise RuntimeError("no VAE decoder attached: pass vae=orv_b200.AutoencoderKLCogVideoX(...) to the pipeline "
                               "or call it with output_type='latent'")
        latents = latents.permute(0, 2, 1, 3, 4)
        latents = 1 / self.vae_scaling_factor_image * latents
        return self.vae.decode(latents).sample

    @staticmethod
    def postprocess_video(video: torch.Tensor, output_type: str = "pil"):
        """diffusers VideoProcessor.postprocess_video (called at reference :1479): [B, C, T, H, W] in [-1, 1] ->
        'pt' [B, T, C, H, W] in [0, 1]; 'np' float32 [B, T, H, W, C]; 'pil' list (batch) of lists (frames) of images."""
        if output_type not in ("pt", "np", "pil"):
            raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil']")
        frames = (video.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return frames
        arr = frames.cpu().permute(0, 1, 3, 4, 2).float().numpy()
        if output_type == "np":
            return arr
        from PIL import Image
        u8 = (arr * 255).round().astype("uint8")
        return [[Image.fromarray(f) for f in clip] for clip in u8]

    # ---- reference :1115-1225 ---------------------------------------------------------------------------
    def prepare_latents(self, image: torch.Tensor, batch_size: int = 1, num_channels_latents: int = 16,
                        num_frames: int = 13, num_views: int = 1, height: int = 60, width: int = 90,
                        dtype: Optional[torch.dtype] = None, device: Optional[torch.device] = None,
                        generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None):
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}. Make sure the batch size matches the length of "
                             "the generators.")
        num_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
        shape = (batch_size, num_views * num_frames, num_channels_latents, height // self.vae_scale_factor_spatial,
                 width // self.vae_scale_factor_spatial)
        p_t = self.transformer.config.patch_size_t
        if p_t is not None:
            shape = shape[:1] + (shape[1] + shape[1] % p_t,) + shape[2:]
        if image.ndim == 4:
            raise RuntimeError("RGB reference images need the VAE encoder, which is outside the B200 hot path; pass "
                               "pre-encoded latents [B, C, F, h, w] (reference dataset.py:655-783)")
        elif image.ndim == 5:
            input_channel = image.size(1)
            if input_channel == num_channels_latents * 2:
                image_latents = DiagonalGaussianDistribution(image).sample(generator)
                image_latents = image_latents.permute(0, 2, 1, 3, 4)
            elif input_channel == num_channels_latents:
                image_latents = image.permute(0, 2, 1, 3, 4)
            else:
                raise RuntimeError(f"Invalid input channels {image.shape=} while {num_channels_latents=}!")
        else:
            raise RuntimeError(f"Invalid dimensions of image input: {image.shape=}")
        invert = getattr(getattr(self.vae, "config", None), "invert_scale_latents", False)
        if not invert:
            image_latents = self.vae_scaling_factor_image * image_latents
        else:
            image_latents = 1 / self.vae_scaling_factor_image * image_latents
        B = image_latents.shape[0]
        image_latents = image_latents.reshape(B, num_views, image_latents.shape[1] // num_views, *image_latents.shape[2:])
        image_frames = image_latents.size(2)
        if image_frames > num_frames:
            raise RuntimeError(f"Invalid input {image_frames=} while {num_frames=}!")
        padding_shape = (batch_size, num_views, num_frames - image_frames, num_channels_latents,
                         height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial)
        latent_padding = torch.zeros(padding_shape, device=device, dtype=dtype)
        image_latents = torch.cat([image_latents, latent_padding], dim=2)
        if p_t is not None:
            # reference quirk (:1212-1214): uses size(1) = n_views of the 6-D tensor, so with ONE view
            # and p_t = 2 a frame is prepended although num_frames was already padded by __call__, and the channel
            # concatenation of the denoise loop then fails in the reference (:1413).  Reproduced by default;
            # `pipe.fix_patch_t_padding = True` uses the frame count per view (what diffusers' own pipeline intends).
            n = image_latents.size(2) if getattr(self, "fix_patch_t_padding", False) else image_latents.size(1)
            first_frame = image_latents[:, :, : n % p_t, ...]
            image_latents = torch.cat([first_frame, image_latents], dim=2)
        image_latents = image_latents.flatten(1, 2)
        if latents is None:
            latents = randn_tensor(shape, generator, device, dtype)
        else:
            latents = latents.to(device)
        latents = latents * self.scheduler.init_noise_sigma
        return latents, image_latents

    # ---- reference :1227-1489 ---------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(
        self,
        image: torch.Tensor,
        prompt: Optional[Union[str, List[str]]] = None,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_views: int = 1,
        num_frames: int = 49,
        num_inference_steps: int = 50,
        timesteps: Optional[List[int]] = None,
        guidance_scale: float = 6,
        use_dynamic_cfg: bool = False,
        num_videos_per_prompt: int = 1,
        eta: float = 0.0,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.FloatTensor] = None,
        prompt_embeds: Optional[torch.FloatTensor] = None,
        negative_prompt_embeds: Optional[torch.FloatTensor] = None,
        output_type: str = "pil",
        return_dict: bool = True,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        max_sequence_length: int = 226,
        controls_or_guidances: Dict[str, torch.Tensor] = {},
    ) -> Union[CogVideoXPipelineOutput, Tuple]:
        self.check_inputs(image, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs,
                          prompt_embeds, negative_prompt_embeds)
        self._guidance_scale = guidance_scale
        self._interrupt = False
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self._execution_device
        do_cfg = guidance_scale > 1.0
        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt=prompt, negative_prompt=negative_prompt, do_classifier_free_guidance=do_cfg,
            num_videos_per_prompt=num_videos_per_prompt, prompt_embeds=prompt_embeds,
            negative_prompt_embeds=negative_prompt_embeds, max_sequence_length=max_sequence_length, device=device)
        prompt_embeds = prompt_embeds.to(device)
        if do_cfg:
            prompt_embeds = torch.cat([negative_prompt_embeds.to(device), prompt_embeds], dim=0)

        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps)
        self._num_timesteps = len(timesteps)

        tcfg = self.transformer.config
        latent_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
        latent_channels = tcfg.in_channels // 2 if tcfg.in_channels != 16 else tcfg.in_channels
        patch_size_t = tcfg.patch_size_t
        controls_or_guidances = dict(controls_or_guidances)
        if patch_size_t is not None and latent_frames % patch_size_t != 0:
            additional_frames = patch_size_t - latent_frames % patch_size_t
            num_frames += additional_frames * self.vae_scale_factor_temporal
            if (actions := controls_or_guidances.get("actions", None)) is not None:
                actions = torch.cat([actions, torch.zeros(
                    (actions.size(0), additional_frames * self.vae_scale_factor_temporal, actions.size(2)),
                    dtype=actions.dtype, device=actions.device)], dim=1)
                controls_or_guidances["actions"] = actions

        invert = getattr(getattr(self.vae, "config", None), "invert_scale_latents", False)
        for key in ("depths", "labels"):  # :1331-1364 — note: sampled WITHOUT the generator, as the reference
            ctl = controls_or_guidances.get(key, None)
            if ctl is not None and ctl.ndim == 5 and ctl.size(1) == latent_channels * 2:
                lat = DiagonalGaussianDistribution(ctl).sample()
                lat = self.vae_scaling_factor_image * lat if not invert else 1 / self.vae_scaling_factor_image * lat
                lat = lat.permute(0, 2, 1, 3, 4)
                controls_or_guidances[key] = torch.cat([lat, lat], dim=2)
        controls_or_guidances = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in controls_or_guidances.items()}

        image = image.to(device, dtype=prompt_embeds.dtype)  # VideoProcessor.preprocess passes latents through
        latents, image_latents = self.prepare_latents(
            image, batch_size * num_videos_per_prompt, latent_channels, num_frames, num_views, height, width,
            prompt_embeds.dtype, device, generator, latents)
        del image

        image_rotary_emb = (self._prepare_rotary_positional_embeddings(height, width, latents.size(1), device)
                            if tcfg.use_rotary_positional_embeddings else None)
        ofs_emb = None if tcfg.ofs_embed_dim is None else latents.new_full((1,), fill_value=2.0)
        if ofs_emb is not None:
            ofs_emb._orvb_host_value = 2.0  # read by the transformer instead of a device -> host copy

        # ---- denoising loop (:1402-1473), with the per-step tensor math fused on the device ----
        is_dpm = isinstance(self.scheduler, CogVideoXDPMScheduler)
        fused = latents.is_cuda and latents.dtype == torch.bfloat16 and callback_on_step_end is None
        ts_list = list(getattr(self.scheduler, "timesteps_host", None) or timesteps.tolist())
        n_cfg = 2 if do_cfg else 1
        draws = noise_dev = noise_pin = noise_evt = None
        if is_dpm and fused:
            # Same generator stream as diffusers' step(): every draw happens, the discarded ones included
            # (`noise_draws`).  A CPU generator (the reference's, inference_control_to_video.py:144) makes this host
            # work of a few ms per step; it is issued AFTER the step's forward has been enqueued, so it overlaps the
            # GPU, and reaches the device through a double-buffered pinned staging area.
            draws = self.scheduler.noise_draws(len(ts_list))
            # Ring of NS staging slots: the host may run up to 2 * NS iterations ahead of the GPU, so a descheduled host
            # thread (the draws are tens of ms of CPU work per iteration on the large configs) cannot stall the device
            # until the whole ring has drained.  (The occasional slow config-5 clip — one in four is 3 - 15 % longer — is
            # NOT this: it happens with 2 and with 8 slots and with a second of host slack, i.e. on the device side,
            # under the power cap; profiles/r02_notes.md, section 13.)
            NS = max(2, int(os.environ.get("ORVB_NOISE_SLOTS", "8")))
            nk = ("noise", tuple(latents.shape), latents.dtype, NS)
            if nk not in self._staging:
                self._staging[nk] = ([torch.empty(latents.shape, dtype=latents.dtype, device=device) for _ in range(NS)],
                                     [torch.empty(latents.shape, dtype=latents.dtype).pin_memory() for _ in range(NS)],
                                     [torch.cuda.Event() for _ in range(NS)],
                                     [torch.cuda.Event() for _ in range(NS)], torch.cuda.Stream(device=device))
            noise_dev, noise_pin, noise_evt, used_evt, copy_stream = self._staging[nk]
            for ev in used_evt:  # "the sampler step that last read this slot has run": trivially true at the start
                ev.record()
        # Persistent device staging (keyed by shape): the transformer reads its large inputs in place, so stable
        # addresses let one captured CUDA graph serve every iteration of every clip.
        def stage(name, t):
            key = (name, tuple(t.shape), t.dtype)
            buf = self._staging.get(key)
            if buf is None or buf.device != device:
                buf = torch.empty(t.shape, dtype=t.dtype, device=device)
                self._staging[key] = buf
            if buf.data_ptr() != t.data_ptr():
                buf.copy_(t, non_blocking=True)
            return buf

        latents = stage("latents", latents.contiguous())
        prompt_embeds = stage("prompt_embeds", prompt_embeds.contiguous())
        controls_or_guidances = {k: (stage(k, v.to(latents.dtype).contiguous()) if torch.is_tensor(v) else v)
                                 for k, v in controls_or_guidances.items()}
        if image_rotary_emb is not None:
            image_rotary_emb = (stage("rope_cos", image_rotary_emb[0].float().contiguous()),
                                stage("rope_sin", image_rotary_emb[1].float().contiguous()))
        Cl = latents.shape[2]
        mi_shape = (n_cfg * latents.shape[0], latents.shape[1], Cl + image_latents.shape[2]) + tuple(latents.shape[3:])
        model_input = self._staging.get(("model_input", mi_shape))
        if model_input is None or model_input.device != device:
            model_input = torch.empty(mi_shape, dtype=latents.dtype, device=device)
            self._staging[("model_input", mi_shape)] = model_input
        model_input[:, :, Cl:] = torch.cat([image_latents] * n_cfg) if do_cfg else image_latents
        model_input[:, :, :Cl] = torch.cat([latents] * 2) if do_cfg else latents
        old_x0 = None
        if is_dpm:
            old_x0 = self._staging.get(("old_x0", tuple(latents.shape)))
            if old_x0 is None or old_x0.device != device:
                old_x0 = torch.empty(latents.shape, dtype=torch.float32, device=device)
                self._staging[("old_x0", tuple(latents.shape))] = old_x0
        have_old = False
        launches = 0
        # The AdaLN tables of all steps depend on (timestep, ofs, actions) only: build them once for the whole run
        # instead of inside every forward (ORVB_MOD_SCHEDULE=0 keeps the per-step path; results are bit-identical).
        use_sched = (os.environ.get("ORVB_MOD_SCHEDULE", "1") != "0"
                     and hasattr(self.transformer, "prepare_modulation_schedule"))
        if use_sched:
            self.transformer.prepare_modulation_schedule([float(t) for t in ts_list], tuple(model_input.shape),
                                                         prompt_embeds.shape[1], controls_or_guidances, ofs=ofs_emb,
                                                         num_views=num_views)
            launches += getattr(self.transformer, "last_schedule_launches", 0)
        # Text projection and control-latent embeddings are step-invariant: computed (and kept in the forward
        # workspace) by the first iteration, restored by the others (ORVB_STATIC_CACHE=0 recomputes them every step as
        # the reference does; results are bit-identical).
        use_static = os.environ.get("ORVB_STATIC_CACHE", "1") != "0" and hasattr(self.transformer, "prepare_modulation_schedule")
        first_step = True
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i, t in enumerate(ts_list):
                if self.interrupt:
                    continue
                if not fused:
                    model_input[:, :, :Cl] = torch.cat([latents] * 2) if do_cfg else latents
                timestep = timesteps[i].expand(model_input.shape[0])
                noise_pred = self.transformer(
                    hidden_states=model_input, encoder_hidden_states=prompt_embeds, timestep=timestep, ofs=ofs_emb,
                    image_rotary_emb=image_rotary_emb, attention_kwargs=attention_kwargs,
                    controls_or_guidances=controls_or_guidances, return_dict=False, num_views=num_views,
                    _static_out=fused, _mod_step=i if use_sched else None,
                    **({"_static_mode": 1 if first_step else 2} if use_static else {}))[0]
                first_step = False
                launches += self.transformer.last_launch_count + 2
                if use_dynamic_cfg:
                    self._guidance_scale = 1 + guidance_scale * (
                        (1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)
                if fused:
                    if is_dpm:
                        slot = i % len(noise_dev)
                        if _is_cpu_gen(generator):
                            for _ in range(draws[i]):
                                nz = randn_tensor(latents.shape, generator, "cpu", latents.dtype)
                            noise_evt[slot].synchronize()  # the H2D that last used this pinned slot has finished
                            noise_pin[slot].copy_(nz)
                            # upload on a side stream WHILE this step's forward runs (on the compute stream the copy
                            # would queue behind the forward and sit in front of the sampler step: ~60 us per step)
                            copy_stream.wait_event(used_evt[slot])
                            with torch.cuda.stream(copy_stream):
                                noise_dev[slot].copy_(noise_pin[slot], non_blocking=True)
                                noise_evt[slot].record()
                            torch.cuda.current_stream().wait_event(noise_evt[slot])
                        else:
                            for _ in range(draws[i]):
                                nz = randn_tensor(latents.shape, generator, device, latents.dtype)
                            noise_dev[slot].copy_(nz)
                        self.scheduler.fused_step(noise_pred, old_x0, have_old, t, ts_list[i - 1] if i > 0 else None,
                                                  latents, noise_dev[slot], n_cfg, self.guidance_scale, model_input)
                        used_evt[slot].record()
                        have_old = True
                    else:
                        self.scheduler.fused_step(noise_pred, t, latents, n_cfg, self.guidance_scale, model_input)
                else:
                    noise_pred = noise_pred.float()
                    if do_cfg:
                        u, c = noise_pred.chunk(2)
                        noise_pred = u + self.guidance_scale * (c - u)
                    if not is_dpm:
                        latents = self.scheduler.step(noise_pred, t, latents, eta=eta, generator=generator,
                                                      return_dict=False)[0]
                    else:
                        latents, x0 = self.scheduler.step(noise_pred, old_x0 if have_old else None, t,
                                                          ts_list[i - 1] if i > 0 else None, latents, eta=eta,
                                                          generator=generator, return_dict=False)
                        old_x0, have_old = x0, True
                    latents = latents.to(prompt_embeds.dtype)
                    if callback_on_step_end is not None:
                        cb = callback_on_step_end(self, i, timesteps[i], {"latents": latents})
                        latents = cb.pop("latents", latents)
                progress_bar.update()
        if use_sched:
            self.transformer.clear_modulation_schedule()
        self.last_step_launches = launches

        B = latents.shape[0]
        latents = latents.clone()
        latents = latents.reshape(B * num_views, latent_frames if patch_size_t is None else latents.shape[1] // num_views,
                                  *latents.shape[2:])
        if not output_type == "latent":
            video = self.decode_latents(latents)
            video = self.postprocess_video(video, output_type)
        else:
            video = latents
        self.maybe_free_model_hooks()
        if not return_dict:
            return (video,)
        return CogVideoXPipelineOutput(frames=video)


def _is_cpu_gen(generator) -> bool:
    if generator is None:
        return False
    g = generator[0] if isinstance(generator, list) else generator
    return g.device.type == "cpu"
