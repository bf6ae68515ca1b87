"""'Look at' camera transform (reference: ``neural_renderer_torch/look_at.py:5-44``).
Pure torch, differentiable, O(nv): it runs before the hot path and is not accelerated."""
import torch
import torch.nn.functional as F


_const_cache = {}


def _as_batch(v, default, B, dev):
    """list / tuple / tensor -> [B, 3] float32 on `dev`.  Python constants are uploaded once per
    device and cached, so the transform issues no host->device copy when replayed or captured in a
    CUDA graph."""
    if v is None:
        v = default
    if not torch.is_tensor(v):
        key = (str(dev), tuple(float(x) for x in v))
        t = _const_cache.get(key)
        if t is None:
            t = _const_cache[key] = torch.tensor(key[1], dtype=torch.float32, device=dev)
        v = t
    else:
        v = v.to(device=dev, dtype=torch.float32)
    return v[None].expand(B, 3) if v.ndim == 1 else v


def _camera_rotation(z_axis, up):
    x_axis = F.normalize(torch.linalg.cross(up, z_axis, dim=-1), dim=-1)
    y_axis = F.normalize(torch.linalg.cross(z_axis, x_axis, dim=-1), dim=-1)
    return torch.stack((x_axis, y_axis, z_axis), dim=1)          # [B, 3, 3], rows are the new axes


def look_at(vertices, viewpoints, at=None, up=None):
    assert vertices.ndim == 3
    dev, B = vertices.device, vertices.shape[0]
    eye = _as_batch(viewpoints, None, B, dev)
    at = _as_batch(at, [0., 0., 0.], B, dev)
    up = _as_batch(up, [0., 1., 0.], B, dev)
    r = _camera_rotation(F.normalize(at - eye, dim=-1), up)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))
