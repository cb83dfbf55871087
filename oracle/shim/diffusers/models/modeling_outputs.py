from dataclasses import dataclass

import torch


@dataclass
class Transformer2DModelOutput:
    sample: "torch.Tensor"  # no default: the reference subclass adds non-default fields (components.py:13-17)
