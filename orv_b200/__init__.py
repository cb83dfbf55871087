"""orv_b200 — B200-native (sm_100a) implementation of ORV's denoising hot path.

Public surface mirrors the reference (orv/models/cogvideox_control.py):
    CogVideoXTransformer3DModelTraj, CogVideoXImageToVideoPipelineTraj, CogVideoXDDIMScheduler, CogVideoXDPMScheduler,
    AutoencoderKLCogVideoX (decode only: diffusers' class as the reference pipeline uses it)
The arithmetic lives in liborv_b200.so (include/orv_b200.h); importing this package never needs a GPU, running it does.
"""
from .models.cogvideox_control import CogVideoXTransformer3DModelTraj  # noqa: F401
from .models.pipeline_control import CogVideoXImageToVideoPipelineTraj, CogVideoXPipelineOutput  # noqa: F401
from .models.autoencoder_kl_cogvideox import AutoencoderKLCogVideoX  # noqa: F401
from .schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler  # noqa: F401

__all__ = ["CogVideoXTransformer3DModelTraj", "CogVideoXImageToVideoPipelineTraj", "CogVideoXPipelineOutput",
           "CogVideoXDDIMScheduler", "CogVideoXDPMScheduler", "AutoencoderKLCogVideoX"]
