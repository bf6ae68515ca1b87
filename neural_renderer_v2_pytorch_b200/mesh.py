"""``Mesh`` (reference ``neural_renderer_torch/mesh.py:8-37``): an ``nn.Module`` holding the vertices
of an .obj file and a learnable legacy (v1, 6-D) texture tensor.

Outside the accelerated path (half-ported and unused by the reference's
own examples); provided so that ``import neural_renderer_v2_pytorch_b200 as neural_renderer_torch``
finds the name.  Differences: ``faces`` is kept as a tensor (the reference calls ``.expand`` on a numpy
array, mesh.py:31) and ``to()`` returns the module."""
import torch
import torch.nn as nn

from .load_obj import load_obj


class Mesh(nn.Module):
    def __init__(self, filename_obj, texture_size=4, normalization=True):
        super().__init__()
        vertices, faces = load_obj(filename_obj, normalization)
        self.vertices = torch.as_tensor(vertices)                       # mesh.py:14 (a plain tensor, not a Parameter)
        self.faces = torch.as_tensor(faces)
        self.num_vertices = self.vertices.shape[0]
        self.num_faces = self.faces.shape[0]
        shape = (self.num_faces, texture_size, texture_size, texture_size, 3)
        self.textures = nn.Parameter(torch.randn(shape))                # mesh.py:19-20
        self.texture_size = texture_size

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        device = torch._C._nn._parse_to(*args, **kwargs)[0]
        if device is not None:
            self.faces = self.faces.to(device)
            self.vertices = self.vertices.to(device)
        return self

    def get_batch(self, batch_size):
        """Broadcast to a minibatch (mesh.py:28-33); textures go through a sigmoid."""
        vertices = self.vertices.expand([batch_size] + list(self.vertices.shape))
        faces = self.faces.expand([batch_size] + list(self.faces.shape))
        textures = torch.sigmoid(self.textures.expand([batch_size] + list(self.textures.shape)))
        return vertices, faces, textures

    def set_lr(self, lr_vertices, lr_textures):
        self.vertices.lr = lr_vertices
        self.textures.lr = lr_textures
