"""Wavefront .obj loader (reference: ``neural_renderer_torch/load_obj.py:7-166``), PIL instead of
imageio.  Host-side file parsing; nothing here touches the GPU.

``load_obj(path)`` -> (vertices [nv,3] f32, faces [nf,3] i32)
``load_obj(path, load_textures=True)`` -> (vertices, faces, vertices_t [nvt,2] f32 in TEXEL units,
faces_t [nf,3] i32, textures [3,H,W] f32): every material's image (or a 2x2 block of its ``Kd`` colour)
is stacked vertically into one atlas, exactly like ``load_textures_func`` (``load_obj.py:25-110``).
"""
import os

import numpy as np


def _tokens(path):
    with open(path) as f:
        for line in f:
            tok = line.split()
            if tok:
                yield tok


def load_mtl(filename_mtl):
    """``load_obj.py:7-22``: material name -> {'texture_filename': ..} / {'color': rgb}."""
    materials, name = {}, ''
    for tok in _tokens(filename_mtl):
        if tok[0] == 'newmtl':
            name = tok[1]
            materials[name] = {}
        elif tok[0] == 'map_Kd':
            materials[name]['texture_filename'] = tok[1]
        elif tok[0] == 'Kd':
            materials[name]['color'] = np.array([float(t) for t in tok[1:4]])
    return materials


def _read_image(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert('RGB'))


def _load_textures(filename_obj, filename_mtl):
    vt = [[float(t) for t in tok[1:3]] for tok in _tokens(filename_obj) if tok[0] == 'vt']
    vertices_t = np.asarray(vt, dtype='float32').reshape(-1, 2)
    faces_t, face_material, material = [], [], ''
    for tok in _tokens(filename_obj):
        if tok[0] == 'usemtl':
            material = tok[1]
        elif tok[0] == 'f':
            ids = [int(t.split('/')[1]) if '/' in t else 0 for t in tok[1:]]
            for i in range(1, len(ids) - 1):
                faces_t.append((ids[0], ids[i], ids[i + 1]))
                face_material.append(material)
    faces_t = np.asarray(faces_t, dtype='int32').reshape(-1, 3) - 1
    face_material = np.asarray(face_material)

    atlas = np.zeros((3, 0, 0), 'float32')
    row0 = 0
    for name, mat in load_mtl(filename_mtl).items():
        mine = face_material == name
        if 'texture_filename' in mat:
            img = _read_image(os.path.join(os.path.dirname(filename_mtl), mat['texture_filename']))
            tex = (img.astype('float32') / 255.).transpose(2, 0, 1)[:, ::-1, :]
            used = np.unique(faces_t[mine].flatten())
            vertices_t[used, 0] *= tex.shape[2] - 1          # normalised uv -> texel units
            vertices_t[used, 1] *= tex.shape[1] - 1
            vertices_t[used, 1] += row0
        else:
            tex = np.ones((3, 2, 2), 'float32') * np.asarray(mat['color'])[:, None, None]
            n = vertices_t.shape[0]
            vertices_t = np.concatenate((vertices_t, np.array([[0, row0], [0, row0 + 1], [1, row0 + 1]], 'float32')), 0)
            faces_t[mine] = np.array([n, n + 1, n + 2])
        width = max(atlas.shape[2], tex.shape[2])
        pad = lambda a: np.concatenate((a, np.zeros((3, a.shape[1], width - a.shape[2]))), 2)
        atlas = np.concatenate((pad(atlas), pad(tex)), 1).astype('float32')
        row0 += tex.shape[1]
    return vertices_t, faces_t, atlas


def load_obj(filename_obj, normalization=True, load_textures=False):
    """Polygons are fan-triangulated; with ``normalization`` the mesh is scaled into the unit cube
    like ``load_obj.py:157-161``."""
    verts, faces = [], []
    mtllib = None
    for tok in _tokens(filename_obj):
        if tok[0] == 'v':
            verts.append([float(t) for t in tok[1:4]])
        elif tok[0] == 'f':
            ids = [int(t.split('/')[0]) for t in tok[1:]]
            for i in range(1, len(ids) - 1):
                faces.append((ids[0], ids[i], ids[i + 1]))
        elif tok[0] == 'mtllib':
            mtllib = tok[1]
    vertices = np.asarray(verts, dtype='float32').reshape(-1, 3)
    faces = np.asarray(faces, dtype='int32').reshape(-1, 3) - 1
    if normalization:
        vertices -= vertices.min(0)[None, :]
        vertices /= np.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[None, :] / 2
    if not load_textures:
        return vertices, faces
    if mtllib is None:
        raise Exception('Failed to load textures.')           # load_obj.py:154, same type and message
    vertices_t, faces_t, textures = _load_textures(filename_obj, os.path.join(os.path.dirname(filename_obj), mtllib))
    return vertices, faces, vertices_t, faces_t, textures
