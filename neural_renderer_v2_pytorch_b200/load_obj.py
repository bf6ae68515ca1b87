"""Wavefront .obj geometry loader (reference: ``neural_renderer_torch/load_obj.py:113-166``,
geometry part).  Host-side file parsing; textures / materials are outside the hot-path scope."""
import numpy as np


def load_obj(filename_obj, normalization=True):
    """Returns (vertices [nv,3] f32, faces [nf,3] i32). Polygons are fan-triangulated; with
    ``normalization`` the mesh is scaled into the unit cube like ``load_obj.py:157-161``."""
    verts, faces = [], []
    with open(filename_obj) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                verts.append([float(t) for t in tok[1:4]])
            elif tok[0] == 'f':
                ids = [int(t.split('/')[0]) for t in tok[1:]]
                for i in range(1, len(ids) - 1):
                    faces.append((ids[0], ids[i], ids[i + 1]))
    vertices = np.asarray(verts, dtype='float32').reshape(-1, 3)
    faces = np.asarray(faces, dtype='int32').reshape(-1, 3) - 1
    if normalization:
        vertices -= vertices.min(0)[None, :]
        vertices /= np.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[None, :] / 2
    return vertices, faces
