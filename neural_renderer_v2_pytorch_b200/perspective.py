"""Perspective divide (reference: ``neural_renderer_torch/perspective.py:4-17``).
Pure torch, differentiable, O(nv): it runs before the hot path and is not accelerated."""
import torch

# the reference converts degrees with pi truncated to 3.1416 (perspective.py:9); kept for parity
_PI_REF = 3.1416


def perspective(vertices, angle=30.):
    assert vertices.ndim == 3
    if not torch.is_tensor(angle):
        angle = torch.as_tensor(float(angle), dtype=torch.float32, device=vertices.device)
    width = torch.tan(angle / 180. * _PI_REF)
    if width.ndim == 0:
        width = width[None].expand(vertices.shape[0])
    width = width[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)
