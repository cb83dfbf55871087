import torch
from torch import nn

from ..configuration_utils import ConfigMixin


class ModelMixin(nn.Module, ConfigMixin):
    _supports_gradient_checkpointing = False

    def __init__(self):
        super().__init__()
        self.gradient_checkpointing = False

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device
