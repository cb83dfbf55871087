"""The helpers of diffusers' CogVideoXImageToVideoPipeline that the reference subclass calls."""
import inspect
from typing import Optional

import torch

from ...models.embeddings import get_3d_rotary_pos_embed
from ..pipeline_utils import DiffusionPipeline


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
    tw, th = tgt_width, tgt_height
    h, w = src
    r = h / w
    if r > (th / tw):
        resize_height = th
        resize_width = int(round(th / h * w))
    else:
        resize_width = tw
        resize_height = int(round(tw / w * h))
    crop_top = int(round((th - resize_height) / 2.0))
    crop_left = int(round((tw - resize_width) / 2.0))
    return (crop_top, crop_left), (crop_top + resize_height, crop_left + resize_width)


def retrieve_latents(encoder_output, generator: Optional[torch.Generator] = None, sample_mode: str = "sample"):
    if hasattr(encoder_output, "latent_dist") and sample_mode == "sample":
        return encoder_output.latent_dist.sample(generator)
    elif hasattr(encoder_output, "latent_dist") and sample_mode == "argmax":
        return encoder_output.latent_dist.mode()
    elif hasattr(encoder_output, "latents"):
        return encoder_output.latents
    raise AttributeError("Could not access latents of provided encoder_output")


class CogVideoXImageToVideoPipeline(DiffusionPipeline):
    def __init__(self, tokenizer, text_encoder, vae, transformer, scheduler):
        super().__init__()
        self.register_modules(tokenizer=tokenizer, vae=vae, text_encoder=text_encoder, transformer=transformer,
                              scheduler=scheduler)
        self.vae_scale_factor_spatial = (
            2 ** (len(self.vae.config.block_out_channels) - 1) if getattr(self, "vae", None) is not None else 8)
        self.vae_scale_factor_temporal = (
            self.vae.config.temporal_compression_ratio if getattr(self, "vae", None) is not None else 4)
        self.vae_scaling_factor_image = self.vae.config.scaling_factor if getattr(self, "vae", None) is not None else 0.7

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def attention_kwargs(self):
        return self._attention_kwargs

    @property
    def interrupt(self):
        return self._interrupt

    def check_inputs(self, image, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs,
                     latents=None, prompt_embeds=None, negative_prompt_embeds=None):
        if (not isinstance(image, torch.Tensor) and not hasattr(image, "size") and not isinstance(image, list)):
            raise ValueError(f"`image` has to be of type `torch.Tensor` or `PIL.Image.Image` or `List[PIL.Image.Image]` "
                             f"but is {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and "
                             "`prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True, num_videos_per_prompt=1,
                      prompt_embeds=None, negative_prompt_embeds=None, max_sequence_length=226, device=None, dtype=None):
        if prompt_embeds is None or (do_classifier_free_guidance and negative_prompt_embeds is None):
            raise NotImplementedError("shim: pass prompt_embeds (and negative_prompt_embeds for CFG); T5 is not run")
        return prompt_embeds, negative_prompt_embeds

    def prepare_extra_step_kwargs(self, generator, eta):
        accepts_eta = "eta" in set(inspect.signature(self.scheduler.step).parameters.keys())
        extra_step_kwargs = {}
        if accepts_eta:
            extra_step_kwargs["eta"] = eta
        accepts_generator = "generator" in set(inspect.signature(self.scheduler.step).parameters.keys())
        if accepts_generator:
            extra_step_kwargs["generator"] = generator
        return extra_step_kwargs

    def _prepare_rotary_positional_embeddings(self, height: int, width: int, num_frames: int, device: torch.device):
        grid_height = height // (self.vae_scale_factor_spatial * self.transformer.config.patch_size)
        grid_width = width // (self.vae_scale_factor_spatial * self.transformer.config.patch_size)
        p = self.transformer.config.patch_size
        p_t = self.transformer.config.patch_size_t
        base_size_width = self.transformer.config.sample_width // p
        base_size_height = self.transformer.config.sample_height // p
        if p_t is None:
            grid_crops_coords = get_resize_crop_region_for_grid((grid_height, grid_width), base_size_width, base_size_height)
            freqs_cos, freqs_sin = get_3d_rotary_pos_embed(
                embed_dim=self.transformer.config.attention_head_dim, crops_coords=grid_crops_coords,
                grid_size=(grid_height, grid_width), temporal_size=num_frames, device=device)
        else:
            base_num_frames = (num_frames + p_t - 1) // p_t
            freqs_cos, freqs_sin = get_3d_rotary_pos_embed(
                embed_dim=self.transformer.config.attention_head_dim, crops_coords=None,
                grid_size=(grid_height, grid_width), temporal_size=base_num_frames, grid_type="slice",
                max_size=(base_size_height, base_size_width), device=device)
        return freqs_cos, freqs_sin

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        latents = latents.permute(0, 2, 1, 3, 4)
        latents = 1 / self.vae_scaling_factor_image * latents
        frames = self.vae.decode(latents).sample
        return frames
