"""'Look at' camera transform (reference: ``neural_renderer_torch/look_at.py:5-44``).
Pure torch, differentiable, O(nv): it runs before the hot path and is not accelerated."""
import torch
import torch.nn.functional as F


def _camera_rotation(z_axis, up):
    x_axis = F.normalize(torch.linalg.cross(up, z_axis, dim=-1), dim=-1)
    y_axis = F.normalize(torch.linalg.cross(z_axis, x_axis, dim=-1), dim=-1)
    return torch.stack((x_axis, y_axis, z_axis), dim=1)          # [B, 3, 3], rows are the new axes


def look_at(vertices, viewpoints, at=None, up=None):
    assert vertices.ndim == 3
    dev, B = vertices.device, vertices.shape[0]

    def as_batch(v, default):
        if v is None:
            v = default
        v = torch.as_tensor(v, dtype=torch.float32, device=dev)
        return v[None].expand(B, 3) if v.ndim == 1 else v

    eye = as_batch(viewpoints, None)
    at = as_batch(at, [0., 0., 0.])
    up = as_batch(up, [0., 1., 0.])
    r = _camera_rotation(F.normalize(at - eye, dim=-1), up)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))
