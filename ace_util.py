"""Drop-in for the reference's `ace_util` (reference ace_util.py:7-22)."""
import numpy as np
import torch


def get_pixel_grid(subsampling_factor):
    """Target pixel positions for a subsampling factor, prediction at the cell centre: 2 x 625 x 625 for factor 8
    (reference ace_util.py:7-13). The CUDA buffer-fill kernel recomputes these values on the fly."""
    pix_range = torch.arange(np.ceil(5000 / subsampling_factor), dtype=torch.float32)
    yy, xx = torch.meshgrid(pix_range, pix_range, indexing='ij')
    return subsampling_factor * (torch.stack([xx, yy]) + 0.5)


def to_homogeneous(input_tensor, dim=1):
    """Append ones along `dim` (reference ace_util.py:16-22)."""
    ones = torch.ones_like(input_tensor.select(dim, 0).unsqueeze(dim))
    return torch.cat([input_tensor, ones], dim=dim)
