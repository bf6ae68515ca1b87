"""``Renderer`` facade with the attributes, defaults and ``render*`` methods of the reference
(``neural_renderer_torch/renderer.py:7-75``).  Camera transforms stay differentiable torch ops;
rasterization is the fused CUDA path of ``rasterize.py``.

New here (the reference has no device logic at all): ``render*`` work on whatever CUDA device the
inputs live on, and :mod:`.parallel` shards a batch of views over the GPUs of one box."""
import math

import torch

from . import camera
from .look import look
from .look_at import look_at
from .perspective import perspective
from .rasterize import rasterize_silhouettes, rasterize_rgba, rasterize_rgb, rasterize_depth
from .rasterize_param import RasterizeParam, RasterizeHyperparam


class Renderer(object):
    def __init__(self):
        # rendering (renderer.py:9-13)
        self.image_size = 256
        self.anti_aliasing = True
        self.draw_backside = True
        self.background_color = None

        # camera (renderer.py:15-22)
        self.perspective = True
        self.viewing_angle = 30
        self.viewpoints = [0, 0, -(1. / math.tan(math.radians(self.viewing_angle)) + 1)]
        self.camera_mode = 'look_at'
        self.camera_direction = [0, 0, 1]
        self.near = 0.1
        self.far = 100

        # not in the reference: bit-reproducible gradients (see RasterizeHyperparam.deterministic)
        self.deterministic = False
        # not in the reference: one fused kernel for look_at / look + perspective (camera.py)
        self.fused_camera = True

    def transform_vertices(self, vertices, lights=None):
        """World -> screen space (renderer.py:24-35).  CUDA tensors go through the fused camera
        kernel (camera.py); set ``fused_camera = False`` for the chain of torch ops."""
        if (self.fused_camera and vertices.is_cuda and not torch.is_tensor(self.viewing_angle)
                and self.camera_mode in ('look_at', 'look')):
            return camera.transform_vertices(vertices, self.viewpoints, self.camera_mode, self.camera_direction,
                                             self.perspective, self.viewing_angle)
        if self.camera_mode == 'look_at':
            vertices = look_at(vertices, self.viewpoints)
        elif self.camera_mode == 'look':
            vertices = look(vertices, self.viewpoints, self.camera_direction)
        if self.perspective:
            vertices = perspective(vertices, angle=self.viewing_angle)
        return vertices

    def _hyperparams(self):
        hp = RasterizeHyperparam(image_size=self.image_size, near=self.near, far=self.far,
                                 anti_aliasing=self.anti_aliasing, draw_backside=self.draw_backside)
        hp.deterministic = self.deterministic
        return hp

    def render_silhouettes(self, vertices, faces, backgrounds=None):
        """[B,nv,3], [nf,3] -> [B,S,S] (renderer.py:37-46)."""
        vertices = self.transform_vertices(vertices)
        params = RasterizeParam(background_color=self.background_color, backgrounds=backgrounds)
        return rasterize_silhouettes(vertices, faces, params, self._hyperparams())

    def render(self, vertices, faces, vertices_t, faces_t, textures, backgrounds=None, lights=None):
        """RGBA [B,4,S,S] (renderer.py:48-56)."""
        vertices = self.transform_vertices(vertices)
        params = RasterizeParam(vertices_textures=vertices_t, faces_textures=faces_t, textures=textures,
                                background_color=self.background_color, backgrounds=backgrounds,
                                lights=lights)
        return rasterize_rgba(vertices, faces, params, self._hyperparams())

    def render_rgb(self, vertices, faces, vertices_t, faces_t, textures, backgrounds=None, lights=None):
        """RGB [B,3,S,S] (renderer.py:58-66)."""
        vertices = self.transform_vertices(vertices, lights)
        params = RasterizeParam(vertices_textures=vertices_t, faces_textures=faces_t, textures=textures,
                                background_color=self.background_color, backgrounds=backgrounds,
                                lights=lights)
        return rasterize_rgb(vertices, faces, params, self._hyperparams())

    def render_depth(self, vertices, faces, backgrounds=None):
        """Depth [B,S,S] (renderer.py:68-75)."""
        vertices = self.transform_vertices(vertices)
        params = RasterizeParam(background_color=self.background_color, backgrounds=backgrounds)
        return rasterize_depth(vertices, faces, params, self._hyperparams())
