"""Wavefront .obj writer (reference: ``neural_renderer_torch/save_obj.py:5-47``), PIL instead of
imageio.  Writes ``name.obj`` and, with textures, ``name.mtl`` + ``name.png``; texture coordinates go
back from texel units to [0,1].  Unlike the reference it does not modify ``vertices_t`` in place
(``save_obj.py:29-30`` divides the caller's array)."""
import os

import numpy as np


def save_obj(filename, vertices, faces, vertices_t=None, faces_t=None, textures=None):
    assert vertices.ndim == 2
    assert faces.ndim == 2
    stem = filename[:-4]
    with_tex = textures is not None
    with open(filename, 'w') as f:
        f.write('# %s\n#\n\n' % os.path.basename(filename))
        if with_tex:
            f.write('mtllib %s\n\n' % os.path.basename(stem + '.mtl'))
        for v in vertices:
            f.write('v %.8f %.8f %.8f\n' % (v[0], v[1], v[2]))
        f.write('\n')
        if with_tex:
            uv = np.array(vertices_t, dtype='float64').reshape(-1, 2)
            uv[:, 0] /= textures.shape[2] - 1
            uv[:, 1] /= textures.shape[1] - 1
            for t in uv:
                f.write('vt %.8f %.8f\n' % (t[0], t[1]))
            f.write('\nusemtl material_1\n')
            for a, b in zip(faces, faces_t):
                f.write('f %d/%d %d/%d %d/%d\n' % (a[0] + 1, b[0] + 1, a[1] + 1, b[1] + 1, a[2] + 1, b[2] + 1))
            f.write('\n')
        else:
            for a in faces:
                f.write('f %d %d %d\n' % (a[0] + 1, a[1] + 1, a[2] + 1))
    if with_tex:
        from PIL import Image
        img = np.clip(np.asarray(textures)[:, ::-1, :].transpose(1, 2, 0) * 255. + 0.5, 0, 255).astype('uint8')
        Image.fromarray(img).save(stem + '.png')
        with open(stem + '.mtl', 'w') as f:
            f.write('newmtl material_1\nmap_Kd %s\n' % os.path.basename(stem + '.png'))
