"""Image loading / preprocessing in front of the trunk (diffsim/diffsim.py:27-41)."""
from __future__ import annotations

import numpy as np
import torch


def load_image(path_or_image):
    """Path or PIL image -> RGB PIL image (the reference uses diffusers.utils.load_image, diffsim/diffsim.py:103-104)."""
    from PIL import Image

    img = Image.open(path_or_image) if isinstance(path_or_image, (str, bytes)) or hasattr(path_or_image, "__fspath__") else path_or_image
    return img.convert("RGB")


def process_image(image, img_size: int = 512) -> torch.Tensor:
    """RGB -> Lanczos resize to img_size^2 -> [-1,1] float32 NCHW, batch 1 (diffsim/diffsim.py:27-41)."""
    from PIL import Image

    image = image.convert("RGB").resize((img_size, img_size), resample=Image.LANCZOS)
    arr = np.asarray(image, dtype=np.float32)[None] / 255.0
    arr = (arr - 0.5) / 0.5
    return torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 3, 1, 2)))
