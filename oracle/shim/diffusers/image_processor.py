from typing import List, Union

import numpy as np
import PIL.Image
import torch

PipelineImageInput = Union[PIL.Image.Image, np.ndarray, torch.Tensor, List[PIL.Image.Image], List[np.ndarray],
                           List[torch.Tensor]]


def is_valid_image(image) -> bool:
    return isinstance(image, PIL.Image.Image) or (isinstance(image, (np.ndarray, torch.Tensor)) and image.ndim in (2, 3))
