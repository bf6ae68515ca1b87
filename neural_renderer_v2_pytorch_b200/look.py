"""'Look' camera transform: fixed viewing direction (reference: ``neural_renderer_torch/look.py:5-42``).
Pure torch, differentiable, O(nv): it runs before the hot path and is not accelerated."""
import torch
import torch.nn.functional as F

from .look_at import _camera_rotation, _as_batch


def look(vertices, viewpoints, direction=None, up=None):
    assert vertices.ndim == 3
    dev, B = vertices.device, vertices.shape[0]
    eye = _as_batch(viewpoints, None, B, dev)
    direction = _as_batch(direction, [0., 0., 1.], B, dev)
    up = _as_batch(up, [0., 1., 0.], B, dev)
    r = _camera_rotation(F.normalize(direction, dim=-1), up)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))
