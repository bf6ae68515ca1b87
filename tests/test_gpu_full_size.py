"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot render
64 x 512^2 x 2464 faces, let alone 1 M faces, in seconds) and through the oracle on a window."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the workload generators of the measured configurations)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import neural_renderer_v2_pytorch_b200 as nr_
    return nr_


def _cfg2(nr):
    w = bench.WORKLOADS["cfg2"]
    inp = bench.make_inputs(w, 1000, torch.device("cuda:0"), nr)
    dev = "cuda:0"
    return w, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}


def _render(nr, w, inp, v, tex, views=slice(None)):
    hp = nr.RasterizeHyperparam(image_size=w["S"], anti_aliasing=w["aa"])
    p = nr.RasterizeParam(vertices_textures=inp["vt"][views], faces_textures=inp["ft"], textures=tex[views])
    return nr.rasterize_rgba(v[views], inp["faces"], p, hp)


def test_config2_full_size_properties(nr):
    """teapot, 64 views, 512^2, RGBA (the benchmarked configuration): the forward is a pure function
    (bit-identical when repeated), views are independent (rendering a sub-batch gives the same bits),
    the silhouette channel is the foreground mask, and the backward is linear in the upstream gradient."""
    w, inp = _cfg2(nr)
    v = inp["vertices"].clone().requires_grad_(True)
    tex = inp["textures"].clone().requires_grad_(True)
    G = inp["G"]
    img = _render(nr, w, inp, v, tex)
    img.backward(G)
    gv, gt = v.grad.clone(), tex.grad.clone()
    with torch.no_grad():
        again = _render(nr, w, inp, v, tex)
        assert torch.equal(img, again)
        for views in (slice(0, 1), slice(17, 40), slice(63, 64)):
            assert torch.equal(img[views], _render(nr, w, inp, v, tex, views))
        hp = nr.RasterizeHyperparam(image_size=w["S"], anti_aliasing=False, draw_rgb=False, draw_depth=False)
        maps = nr.rasterize_maps(v.detach(), inp["faces"], nr.RasterizeParam(), hp)
        fg = (maps["face_index_map"] >= 0).flip(1, 2).float()
        assert torch.equal(img[:, 3], fg)
        assert torch.all(img[:, :3][fg[:, None].expand(-1, 3, -1, -1) == 0] == 0)
        assert 0.05 < float(fg.mean()) < 0.5
    # The texture gradient is linear in the upstream gradient; the vertex gradient goes through the
    # stencil's sign selection (utils.py:91-101), which is only positively homogeneous:
    # backward(4 G) = 4 backward(G) up to its |r - l| < 1e-4 dead zone.
    v.grad = tex.grad = None
    G2 = torch.randn(G.shape, generator=torch.Generator().manual_seed(2)).to(G.device)
    _render(nr, w, inp, v, tex).backward(2 * G + G2)
    gt3 = tex.grad.clone()
    v.grad = tex.grad = None
    _render(nr, w, inp, v, tex).backward(G2)
    want = 2 * gt + tex.grad
    assert float((gt3 - want).abs().max()) <= 1e-4 * float(want.abs().max())
    v.grad = tex.grad = None
    _render(nr, w, inp, v, tex).backward(4 * G)
    bad = ((v.grad - 4 * gv).abs() > 1e-4 * float(gv.abs().max()) * 4).float().mean()
    assert float(bad) < 1e-3, float(bad)
    assert torch.isfinite(gv).all() and torch.isfinite(gt).all()


def test_config4_window_against_the_oracle(nr):
    """1 M random triangles at 1024^2 (general multi-kernel binning, long tile lists): two views
    rendered at full size; inside a 96 x 96 window the face index map must equal the C oracle run
    on the faces whose bounding box touches the window (no other face can influence those pixels)."""
    w = dict(bench.WORKLOADS["cfg4"], views=2)
    inp = bench.make_inputs(w, 1000, torch.device("cuda:0"), nr)
    v = inp["vertices"].cuda()
    faces = inp["faces"].cuda()
    S = w["S"]
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=False, draw_rgb=False, draw_depth=False)
    maps = nr.rasterize_maps(v, faces, nr.RasterizeParam(), hp)
    fim = maps["face_index_map"]
    again = nr.rasterize_maps(v, faces, nr.RasterizeParam(), hp)["face_index_map"]
    assert torch.equal(fim, again)
    # the dense-mesh mode (8x8 tiles; tens of thousands of long lists, sorted one warp each in shared memory)
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    rz.FORCE_FINE_TILES = True
    try:
        fine = nr.rasterize_maps(v, faces, nr.RasterizeParam(), hp)["face_index_map"]
    finally:
        rz.FORCE_FINE_TILES = False
    assert torch.equal(fim, fine)
    # the face-parallel raster kernel (unsorted lists, shared-memory z-buffer, exact replay of contested pixels)
    for force in (True, False):
        rz.FORCE_DENSE_RASTER = force
        try:
            other = nr.rasterize_maps(v, faces, nr.RasterizeParam(), hp)["face_index_map"]
        finally:
            rz.FORCE_DENSE_RASTER = None
        assert torch.equal(fim, other), "dense=%s: %d px differ" % (force, int((fim != other).sum()))
    one = nr.rasterize_maps(v[1:2], faces, nr.RasterizeParam(), hp)["face_index_map"]
    assert torch.equal(fim[1:2], one)
    fv = v[:, faces.long()]                                           # [2, nf, 3, 3]
    x0, y0, n = 500, 440, 96
    lo_x, hi_x = (2 * x0 + 1 - S) / S, (2 * (x0 + n - 1) + 1 - S) / S
    lo_y, hi_y = (2 * y0 + 1 - S) / S, (2 * (y0 + n - 1) + 1 - S) / S
    for b in range(2):
        f = fv[b]
        keep = ((f[:, :, 0].max(1).values >= lo_x) & (f[:, :, 0].min(1).values <= hi_x) &
                (f[:, :, 1].max(1).values >= lo_y) & (f[:, :, 1].min(1).values <= hi_y))
        ids = torch.nonzero(keep)[:, 0]
        assert 1000 < ids.numel() < 100000
        sub = f[ids].cpu().numpy()[None]
        want = oracle.face_index_map(sub, S)[0, y0:y0 + n, x0:x0 + n]
        want = np.where(want >= 0, ids.cpu().numpy()[np.maximum(want, 0)], -1)
        got = fim[b, y0:y0 + n, x0:x0 + n].cpu().numpy()
        assert np.array_equal(got, want), "%d / %d window pixels differ" % ((got != want).sum(), got.size)
        assert (got >= 0).mean() > 0.3


def test_many_views_more_than_one_wave_of_binning_clusters(nr):
    """150 views (more binning clusters than SMs, so they run in waves): every view must equal the
    same view rendered alone."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "teapot.npz"))
    B, S = 150, 64
    g = torch.Generator().manual_seed(150)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20, torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye)).cuda()
    faces = torch.from_numpy(d["faces"]).cuda()
    hp = lambda: nr.RasterizeHyperparam(image_size=S, anti_aliasing=False)
    with torch.no_grad():
        all_views = nr.rasterize_silhouettes(vs, faces, nr.RasterizeParam(), hp())
        for b in (0, 1, 63, 64, 127, 128, 149):
            one = nr.rasterize_silhouettes(vs[b:b + 1], faces, nr.RasterizeParam(), hp())
            assert torch.equal(all_views[b:b + 1], one), b
    assert 0.02 < float(all_views.mean()) < 0.6
