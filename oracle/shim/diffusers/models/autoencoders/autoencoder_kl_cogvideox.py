from types import SimpleNamespace


class AutoencoderKLCogVideoX:
    """Config-only stand-in: the sampler reads vae.config.* and never encodes/decodes in latent-in/latent-out tests."""

    def __init__(self, scaling_factor=1.15258426, invert_scale_latents=False):
        self.config = SimpleNamespace(block_out_channels=(128, 256, 256, 512), temporal_compression_ratio=4,
                                      scaling_factor=scaling_factor, latent_channels=16,
                                      invert_scale_latents=invert_scale_latents)

    def to(self, *a, **k):
        return self
