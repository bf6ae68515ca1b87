"""Small helpers the reference exports next to the renderer
(``neural_renderer_torch/utils.py:18-72``): ``to_gpu``, ``create_textures``,
``get_points_from_angles``.  Host-side input preparation only."""
import math

import numpy as np
import torch


def to_gpu(data, device=None):
    """``utils.py:18-22``."""
    if isinstance(data, (tuple, list)):
        return [torch.as_tensor(d).cuda(device) for d in data]
    return torch.as_tensor(data).cuda(device)


def imread(filename):
    """``utils.py:25-27``: image file -> float32 array in [0, 1] (PIL instead of imageio)."""
    from PIL import Image
    return np.asarray(Image.open(filename), dtype='float32') / 255.


def make_gif(working_directory, filename):
    """``utils.py:10-15``: assemble ``_tmp_*.png`` frames into a gif (PIL instead of ImageMagick)."""
    import glob
    import os
    from PIL import Image
    frames = sorted(glob.glob('%s/_tmp_*.png' % working_directory))
    images = [Image.open(f).convert('RGB') for f in frames]
    if images:
        images[0].save(filename, save_all=True, append_images=images[1:], duration=80, loop=0)
    for f in frames:
        os.remove(f)


def create_textures(num_faces, texture_size=16, flatten=False):
    """Per-face texture atlas (``utils.py:30-52``): face i owns the texture_size^2 block at
    (row, column) = divmod(i, tile_width) and maps its corners to three corners of that block.
    Returns (vertices_textures [3nf,2] f32, faces_textures [nf,3] i32, textures [3,H,W] ones)."""
    if flatten:
        tile_w, tile_h = 1, num_faces
    else:
        tile_w = int((num_faces - 1.) ** 0.5) + 1
        tile_h = int((num_faces - 1.) / tile_w) + 1
    ts = texture_size
    textures = np.ones((3, tile_h * ts, tile_w * ts), 'float32')
    idx = np.arange(num_faces)
    col, row = idx % tile_w, idx // tile_w
    vt = np.zeros((num_faces, 3, 2), 'float32')
    vt[:, 0] = np.stack((col * ts, row * ts), 1)
    vt[:, 1] = np.stack((col * ts, (row + 1) * ts - 1), 1)
    vt[:, 2] = np.stack(((col + 1) * ts - 1, (row + 1) * ts - 1), 1)
    ft = np.arange(num_faces * 3, dtype='int32').reshape(num_faces, 3)
    return vt.reshape(num_faces * 3, 2), ft, textures


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """Camera position on a sphere (``utils.py:55-72``); scalars give a tuple, tensors give [B,3]."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth),
                distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    if degrees:
        # the tensor branch of the reference uses this literal for pi (utils.py:66-67)
        elevation = elevation / 180. * 3.14159265359
        azimuth = azimuth / 180. * 3.14159265359
    return torch.stack((distance * torch.cos(elevation) * torch.sin(azimuth),
                        distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)), dim=1)
