"""Perspective divide (reference: ``neural_renderer_torch/perspective.py:4-17``).
Pure torch, differentiable, O(nv): it runs before the hot path and is not accelerated."""
import torch

# the reference converts degrees with pi truncated to 3.1416 (perspective.py:9); kept for parity
_PI_REF = 3.1416


def perspective(vertices, angle=30.):
    assert vertices.ndim == 3
    if not torch.is_tensor(angle):
        # float32 arithmetic of the reference (angle / 180 * 3.1416, then tan), done on the host so
        # that no tensor has to be uploaded (CUDA-graph friendly)
        width = float(torch.tan(torch.tensor(float(angle), dtype=torch.float32) / 180. * _PI_REF))
    else:
        width = torch.tan(angle.to(vertices.device) / 180. * _PI_REF)
        if width.ndim == 0:
            width = width[None].expand(vertices.shape[0])
        width = width[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)
