"""torch-CPU float32 restatement of the reference's pure-torch stages (ORACLE, tests only).

Each function cites the reference lines it follows (paths relative to
``/root/reference/neural_renderer_torch``).  The reference loops over the batch in Python
with boolean masks; here the same arithmetic is written batched so that the oracle
finishes in seconds.  The z-buffer and the weight map come from the C restatement in
``oracle/nr_oracle.c`` and are constants for autograd, exactly like the reference where
they are produced by a CUDA kernel with no autograd edge (``rasterize.py:34,75``).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import face_index_map as _c_face_index_map, weight_map as _c_weight_map


def to_map(data_in, indices):
    """``utils.py:104-114``: out[b,y,x] = data_in[b, indices[b,y,x]] where indices >= 0, else 0.
    Gradient reaches ``data_in`` only through foreground pixels (index_put accumulate)."""
    B = data_in.shape[0]
    mask = indices >= 0
    idx = indices.clamp(min=0).long()
    gathered = data_in[torch.arange(B)[:, None, None], idx]
    m = mask.reshape(mask.shape + (1,) * (gathered.ndim - mask.ndim))
    return torch.where(m, gathered, torch.zeros((), dtype=gathered.dtype))


def mask_foreground(data, fim):
    """``utils.py:117-160``: copy where fim >= 0, zeros elsewhere, same mask on the gradient."""
    m = (fim >= 0).reshape(fim.shape + (1,) * (data.ndim - fim.ndim))
    return torch.where(m, data, torch.zeros((), dtype=data.dtype))


def _shift_sum(t, axis):
    """pad_zeros(t,1,axis,'right') + pad_zeros(t,1,axis,'left') (``utils.py:75-88``)."""
    pad_r = [0, 0] * t.ndim
    pad_l = [0, 0] * t.ndim
    # F.pad counts dims from the last one backwards
    k = (t.ndim - 1 - axis) * 2
    pad_r[k + 1] = 1
    pad_l[k] = 1
    return F.pad(t, pad_r), F.pad(t, pad_l)


def maximum(data_right, data_left, eps=1e-4):
    """``utils.py:91-101``."""
    max_map = torch.max(data_right, data_left) <= 0
    abs_map = torch.abs(data_right - data_left) < eps
    rl_map = data_right > data_left
    out = torch.where(rl_map, -data_right, data_left)
    out = torch.where(abs_map, torch.zeros_like(out), out)
    out = torch.where(max_map, torch.zeros_like(out), out)
    return out


def differentiation_backward(images, grad_output):
    """``differentiation.py:13-36``: grad of the loss w.r.t. per-pixel (x, y) coordinates.

    images, grad_output: [B,R,R,C] -> [B,R,R,2]"""
    R = images.shape[1]
    step = 2. / R

    gyr = -((images[:, :-1, :] - images[:, 1:, :]) * grad_output[:, 1:, :]).sum(-1) / step
    a, b = _shift_sum(gyr[..., None], 1)
    gyr = a + b
    gyl = -((images[:, 1:, :] - images[:, :-1, :]) * grad_output[:, :-1, :]).sum(-1) / step
    a, b = _shift_sum(gyl[..., None], 1)
    gyl = b + a
    grad_y = maximum(gyr, gyl)

    gxr = -((images[:, :, :-1] - images[:, :, 1:]) * grad_output[:, :, 1:]).sum(-1) / step
    a, b = _shift_sum(gxr[..., None], 2)
    gxr = a + b
    gxl = -((images[:, :, 1:] - images[:, :, :-1]) * grad_output[:, :, :-1]).sum(-1) / step
    a, b = _shift_sum(gxl[..., None], 2)
    gxl = b + a
    grad_x = maximum(gxr, gxl)
    return torch.cat((grad_x, grad_y), -1)


class _Differentiation(torch.autograd.Function):
    """``differentiation.py:6-36``: identity forward, stencil backward."""

    @staticmethod
    def forward(ctx, images, coordinates):
        ctx.save_for_backward(images)
        return images.clone()

    @staticmethod
    def backward(ctx, gradients):
        images, = ctx.saved_tensors
        return gradients, differentiation_backward(images, gradients)


def differentiation(images, coordinates):
    return _Differentiation.apply(images, coordinates)


def sample_textures(faces, faces_textures, textures, fim, wmap, eps):
    """``rasterize.py:100-153``. faces [B,nf,3,3], faces_textures [B,nf,3,2],
    textures [B,3,H,W] -> rgb [B,R,R,3]."""
    B = faces.shape[0]
    H, W = textures.shape[2:]
    tex = textures.permute(0, 2, 3, 1).reshape(B, H * W, 3)
    z_map = to_map(faces[:, :, :, 2], fim)                         # [B,R,R,3]
    vt_map = to_map(faces_textures, fim)                           # [B,R,R,3,2]
    depth = 1. / (wmap / (z_map + 1e-10) + 1e-10).sum(-1)          # [B,R,R]
    vt_orig = vt_map.clone()
    uv = (wmap[..., None] * vt_map / (z_map[..., None] + 1e-10)).sum(-2)
    uv = uv * depth[..., None]
    uv = torch.max(uv, vt_orig.min(-2).values)
    uv = torch.min(uv, vt_orig.max(-2).values - eps)
    uv = mask_foreground(uv, fim)

    x_f, y_f = uv[..., 0], uv[..., 1]
    x_f_f, y_f_f = torch.floor(x_f), torch.floor(y_f)
    x_c_f, y_c_f = x_f_f + 1, y_f_f + 1
    x_f_i, y_f_i = x_f_f.to(torch.int32), y_f_f.to(torch.int32)
    x_c_i, y_c_i = x_c_f.to(torch.int32), y_c_f.to(torch.int32)
    vtm1 = y_f_i * W + x_f_i
    vtm2 = y_f_i * W + x_c_i
    vtm3 = y_c_i * W + x_f_i
    vtm4 = y_c_i * W + x_c_i
    w1 = (y_c_f - y_f) * (x_c_f - x_f)
    w2 = (y_c_f - y_f) * (x_f - x_f_f)
    w3 = (y_f - y_f_f) * (x_c_f - x_f)
    w4 = (y_f - y_f_f) * (x_f - x_f_f)
    rgb = (w1[..., None] * to_map(tex, vtm1) + w2[..., None] * to_map(tex, vtm2) +
           w3[..., None] * to_map(tex, vtm3) + w4[..., None] * to_map(tex, vtm4))
    return mask_foreground(rgb, fim)


def compute_depth_map(faces, fim, wmap):
    """``rasterize.py:80-88``."""
    z_map = to_map(faces[:, :, :, -1:], fim)[..., 0]
    depth = 1. / torch.sum(wmap / z_map, -1)
    return mask_foreground(depth, fim)


def compute_coordinate_map(faces, fim, wmap):
    """``rasterize.py:91-97``."""
    faces_map = to_map(faces[:, :, :, :2], fim)
    return torch.sum(faces_map * wmap[..., None], -2)


def compute_normal_map(vertices, faces_idx, faces, fim, wmap):
    """``rasterize.py:162-190`` (smooth=True): face normals summed onto their vertices (each vertex
    once per face), normalised, interpolated with the weight map."""
    v01 = faces[:, :, 1, :] - faces[:, :, 0, :]
    v12 = faces[:, :, 2, :] - faces[:, :, 1, :]
    n = torch.linalg.cross(v01, v12, dim=-1)                         # [B,nf,3]
    nf, nv = faces_idx.shape[0], vertices.shape[1]
    m = torch.zeros((nf, nv), dtype=torch.float32)
    for k in range(3):
        m[torch.arange(nf), faces_idx[:, k]] = 1
    vn = torch.matmul(n.permute(0, 2, 1), m).permute(0, 2, 1)        # [B,nv,3]
    vn = F.normalize(vn, dim=2)
    normal_map = to_map(vn[:, faces_idx], fim)                       # [B,R,R,3,3]
    return torch.sum(wmap[..., None] * normal_map, dim=-2)


def light_color_weights(normal_map, lights):
    """``rasterize.py:252-283``. lights: list of dicts {type: ambient|directional|specular, color [B,3],
    direction [B,3], alpha [B], backside}."""
    cw = torch.zeros_like(normal_map)
    for L in lights:
        color = L["color"][:, None, None, :]
        if L["type"] == "ambient":
            cw = cw + color.expand(cw.shape)
            continue
        if L["type"] == "directional":
            intensity = torch.sum(-L["direction"][:, None, None, :] * normal_map, -1)
        else:
            eye = torch.tensor([0., 0., 1.])
            intensity = torch.sum(-eye[None, None, None, :] * normal_map, -1)
        intensity = torch.abs(intensity) if L.get("backside") else torch.relu(intensity)
        if L["type"] == "specular":
            intensity = intensity ** L["alpha"][:, None, None]
        cw = cw + intensity[..., None] * color
    return cw


def rasterize(vertices, faces, image_size, anti_aliasing, near=0.1, far=100.0, eps=1e-5,
              draw_backside=True, draw_rgb=False, draw_silhouettes=True, draw_depth=False,
              vertices_textures=None, faces_textures=None, textures=None, lights=None, backgrounds=None,
              return_maps=False):
    """``rasterize.py:194-329`` (``rasterize_core``).  ``backgrounds`` [B,3,R,R] follows what
    ``blend_backgrounds`` (``rasterize.py:156-159``) is meant to do: that function fails on torch tensors
    (``.astype``, ``[::-1]``), so this restates the Chainer original it was ported from
    (``neural_renderer_chainer/rasterize.py:574-577, 722-725``) -- no reference output exists to pin it.

    vertices [B,nv,3] screen space (may require grad), faces [nf,3] int.
    Returns images [B,C,S,S] (and the internal maps when ``return_maps``)."""
    R = image_size * 2 if anti_aliasing else image_size
    fidx = torch.as_tensor(np.asarray(faces)).long()
    fv = vertices[:, fidx]                                           # rasterize.py:232
    fv_np = fv.detach().cpu().numpy()
    fim_np = _c_face_index_map(fv_np, R, near, far, draw_backside)   # rasterize.py:235
    wm_np = _c_weight_map(fv_np, fim_np)                             # rasterize.py:236
    fim = torch.from_numpy(fim_np)
    wmap = torch.from_numpy(wm_np)
    coord = compute_coordinate_map(fv, fim, wmap)                    # rasterize.py:237

    chans = []
    if draw_rgb:
        ft = vertices_textures[:, torch.as_tensor(np.asarray(faces_textures)).long()]
        rgb = sample_textures(fv, ft, textures, fim, wmap, eps)
        if lights is not None:
            rgb = rgb * light_color_weights(compute_normal_map(vertices, fidx, fv, fim, wmap), lights)
        if backgrounds is not None:
            fg = (0 <= fim).to(torch.float32)[..., None]
            rgb = fg * rgb + (1 - fg) * torch.flip(backgrounds.permute(0, 2, 3, 1), dims=(1, 2))
        chans.append(rgb)
    if draw_silhouettes:
        chans.append((0 <= fim).to(torch.float32)[..., None])
    if draw_depth:
        chans.append(compute_depth_map(fv, fim, wmap)[..., None])
    images = torch.cat(chans, -1) if len(chans) > 1 else chans[0]   # rasterize.py:295-310
    internal = images
    images = differentiation(images, coord)                          # rasterize.py:313
    images = images.permute(0, 3, 1, 2)
    images = torch.flip(images, dims=(2, 3))                         # rasterize.py:315-316
    if anti_aliasing:                                                # rasterize.py:321-328
        images = (images[:, :, 0::2, 0::2] + images[:, :, 1::2, 0::2] +
                  images[:, :, 0::2, 1::2] + images[:, :, 1::2, 1::2])
        images = images / 4.
    if return_maps:
        return images, dict(face_index_map=fim, weight_map=wmap, coordinate_map=coord,
                            internal_images=internal)
    return images
