"""Light sources, same classes and constructor arguments as the reference
(``neural_renderer_torch/lights.py:4-39``).  ``color`` / ``direction`` are [B,3] tensors (one row per
view), ``alpha`` is [B].  Shading itself runs inside the fused raster / backward kernels
(``csrc/nr_common.cuh::light_weights``, reference ``rasterize.py:252-283``)."""
import torch


class Light:
    def __init__(self, color):
        self.color = color

    def to(self, device):
        self.color = self.color.to(device)


class DirectionalLight(Light):
    def __init__(self, color, direction, backside=False):
        super().__init__(color)
        self.direction = direction
        self.backside = backside

    def to(self, device):
        super().to(device)
        self.direction = self.direction.to(device)


class AmbientLight(Light):
    def __init__(self, color):
        super().__init__(color)


class SpecularLight(Light):
    def __init__(self, color, alpha=None, backside=False):
        super().__init__(color)
        self.backside = backside
        self.alpha = alpha if alpha is not None else torch.ones(color.shape[0], dtype=torch.float32)

    def to(self, device):
        super().to(device)
        self.alpha = self.alpha.to(device)


def pack_lights(lights, batch, device):
    """-> (types [L] int32, data [L,B,8] float32) on ``device``.

    The lights are CONSTANTS of the fused kernels: the backward returns gradients for the vertex normals (hence
    the vertices) and the textures, not for ``color`` / ``direction`` / ``alpha`` (in the reference those flow
    through torch ops, rasterize.py:256-283).  A light tensor that requires grad therefore raises instead of
    silently receiving none."""
    types, rows = [], []
    if len(lights) == 0:
        # `lights=[]` is "lit by nothing": a black image in the reference (the accumulated colour weight stays 0)
        lights = [AmbientLight(torch.zeros((batch, 3), dtype=torch.float32))]
    for light in lights:
        for name in ("color", "direction", "alpha"):
            t = getattr(light, name, None)
            if torch.is_tensor(t) and t.requires_grad:
                raise NotImplementedError(
                    "gradients with respect to light.%s are not implemented by the fused kernels (lights are "
                    "constants here); detach() it, or optimise it with the torch-op reference" % name)
        color = torch.as_tensor(light.color, dtype=torch.float32).to(device).detach()
        assert color.shape == (batch, 3), "light colour must be [batch, 3]"
        row = torch.zeros((batch, 8), dtype=torch.float32, device=device)
        row[:, 0:3] = color
        if isinstance(light, AmbientLight):
            t = 0
        elif isinstance(light, DirectionalLight):
            t = 1 | (4 if light.backside else 0)
            row[:, 3:6] = torch.as_tensor(light.direction, dtype=torch.float32).to(device).detach()
        elif isinstance(light, SpecularLight):
            t = 2 | (4 if light.backside else 0)
            row[:, 6] = torch.as_tensor(light.alpha, dtype=torch.float32).to(device).detach()
        else:
            raise TypeError("unknown light type %r" % type(light))
        types.append(t)
        rows.append(row)
    return (torch.tensor(types, dtype=torch.int32, device=device), torch.stack(rows, 0).contiguous())


def vertex_normals(vertices, faces):
    """Normalised per-vertex normals [B,nv,3] (``rasterize.py:167-182``): the cross products
    (v1-v0) x (v2-v1) of the faces touching a vertex, each face counted once per vertex, summed with
    index_add instead of the reference's dense [nf,nv] incidence matrix.  Differentiable torch ops."""
    import torch.nn.functional as F
    idx = faces.long()
    fv = vertices[:, idx]                                            # [B,nf,3,3]
    n = torch.linalg.cross(fv[:, :, 1] - fv[:, :, 0], fv[:, :, 2] - fv[:, :, 1], dim=-1)
    vn = torch.zeros_like(vertices)
    vn = vn.index_add(1, idx[:, 0], n)
    once1 = (idx[:, 1] != idx[:, 0]).to(n.dtype)[None, :, None]
    vn = vn.index_add(1, idx[:, 1], n * once1)
    once2 = ((idx[:, 2] != idx[:, 0]) & (idx[:, 2] != idx[:, 1])).to(n.dtype)[None, :, None]
    vn = vn.index_add(1, idx[:, 2], n * once2)
    return F.normalize(vn, dim=2)
