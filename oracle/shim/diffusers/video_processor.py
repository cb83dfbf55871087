from types import SimpleNamespace


class VideoProcessor:
    """Config holder; the reference subclass overrides preprocess (latent inputs pass straight through)."""

    def __init__(self, do_resize: bool = True, vae_scale_factor: int = 8, vae_latent_channels: int = 4,
                 resample: str = "lanczos", do_normalize: bool = True, do_binarize: bool = False,
                 do_convert_rgb: bool = False, do_convert_grayscale: bool = False):
        self.config = SimpleNamespace(do_resize=do_resize, vae_scale_factor=vae_scale_factor,
                                      vae_latent_channels=vae_latent_channels, resample=resample,
                                      do_normalize=do_normalize, do_binarize=do_binarize,
                                      do_convert_rgb=do_convert_rgb, do_convert_grayscale=do_convert_grayscale)

    def postprocess_video(self, video, output_type: str = "np"):
        return video
