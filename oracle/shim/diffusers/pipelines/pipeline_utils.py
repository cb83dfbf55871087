import torch


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        for name, module in kwargs.items():
            setattr(self, name, module)

    @property
    def _execution_device(self):
        t = getattr(self, "transformer", None)
        if t is not None:
            return next(t.parameters()).device
        return torch.device("cpu")

    def to(self, *args, **kwargs):
        for name in ("transformer", "vae", "text_encoder"):
            m = getattr(self, name, None)
            if m is not None and hasattr(m, "to"):
                m.to(*args, **kwargs)
        return self

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm
        cfg = getattr(self, "_progress_bar_config", {"disable": True})
        if iterable is not None:
            return tqdm(iterable, **cfg)
        return tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def maybe_free_model_hooks(self):
        pass
