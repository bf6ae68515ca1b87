"""Parity against the oracle at the EXACT configurations BASELINE.json names and bench.py measures
(inputs from bench.make_inputs, same seeds): config 2 at its full size (all 64 views, 512^2, RGBA), config 1
(one view, 256^2 with anti-aliasing, texture_size 16), config 5 (32 views, orthographic camera, tanh(textures),
2x anti-aliasing), config 3 (100 k-face shared mesh, 8 views, gradient to the ONE world-space parameter through
Renderer), and the gradients inside a window of config 4 (1 M triangles at 1024^2).

Bars: face_index_map bit-exact; images rtol 1e-5 / atol 1e-6; gradients |d| <= 1e-5 |ref| + 1e-5 max|ref|
per component AND relative L2 error <= 3e-6 over the whole tensor (float32 sums of thousands of terms whose
order differs between the atomics here, index_put in the reference and the CPU oracle).
Reference: neural_renderer_torch/rasterize.py:194-329, differentiation.py:13-36.
"""
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from oracle import pipeline as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the workload generators of the measured configurations)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import neural_renderer_v2_pytorch_b200 as nr_
    return nr_


def grad_close(got, want, what, l2=3e-6):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.abs(want).max()
    assert scale > 0, what + ": reference gradient is all zero"
    tol = 1e-5 * np.abs(want) + 1e-5 * scale
    bad = np.abs(got - want) > tol
    assert not bad.any(), "%s: %d / %d beyond tolerance, max |d| = %.3g (scale %.3g)" % (
        what, bad.sum(), bad.size, np.abs(got - want).max(), scale)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel <= l2, "%s: relative L2 error %.3g > %.3g" % (what, rel, l2)


def oracle_run(inp, w, views=None):
    """forward + backward of the oracle on the first `views` views; returns images, maps, gradients."""
    views = w["views"] if views is None else views
    rgb = w["mode"] in ("rgb", "rgba")
    v = inp["vertices"][:views].cpu().clone().requires_grad_(True)
    kw, tex = {}, None
    if rgb:
        tex = inp["textures"][:views].clone().requires_grad_(True)
        kw = dict(vertices_textures=inp["vt"][:views], faces_textures=inp["ft"].numpy(),
                  textures=torch.tanh(tex) if w.get("tanh") else tex)
    img, maps = ref.rasterize(v, inp["faces"], w["S"], w["aa"], draw_rgb=rgb,
                              draw_silhouettes=w["mode"] in ("silhouettes", "rgba"), return_maps=True, **kw)
    G = inp["G"][:views]
    img.backward(G if G.ndim == 4 else G[:, None])
    return img.detach(), maps, v.grad, (tex.grad if tex is not None else None)


def cuda_run(nr, inp, w, views=None):
    views = w["views"] if views is None else views
    rgb = w["mode"] in ("rgb", "rgba")
    S = w["S"]
    v = inp["vertices"][:views].to(DEV).clone().requires_grad_(True)
    faces = inp["faces"].to(DEV)
    tex = None
    p = nr.RasterizeParam()
    if rgb:
        tex = inp["textures"][:views].to(DEV).clone().requires_grad_(True)
        p = nr.RasterizeParam(vertices_textures=inp["vt"][:views].to(DEV), faces_textures=inp["ft"].to(DEV),
                              textures=torch.tanh(tex) if w.get("tanh") else tex)
    fn = {"rgba": nr.rasterize_rgba, "rgb": nr.rasterize_rgb, "silhouettes": nr.rasterize_silhouettes}[w["mode"]]
    img = fn(v, faces, p, nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"]))
    img.backward(inp["G"][:views].to(DEV))
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"], draw_rgb=False, draw_silhouettes=True, draw_depth=False)
    fim = nr.rasterize_maps(v.detach(), faces, nr.RasterizeParam(), hp)["face_index_map"]
    return img.detach(), fim, v.grad, (tex.grad if tex is not None else None)


def compare(nr, name, views=None):
    w = bench.WORKLOADS[name]
    inp = bench.make_inputs(w, 1000, "cpu", nr)
    img, fim, gv, gt = cuda_run(nr, inp, w, views)
    oimg, omaps, ogv, ogt = oracle_run(inp, w, views)
    nbad = int((fim.cpu() != omaps["face_index_map"]).sum())
    assert nbad == 0, "%s: face_index_map differs in %d pixels" % (name, nbad)
    oimg = oimg if img.ndim == 4 else oimg[:, 0]
    np.testing.assert_allclose(img.cpu().numpy(), oimg.numpy(), rtol=1e-5, atol=1e-6)
    grad_close(gv.cpu().numpy(), ogv.numpy(), name + " grad_vertices")
    if gt is not None:
        grad_close(gt.cpu().numpy(), ogt.numpy(), name + " grad_textures")
    assert 0.02 < float((fim >= 0).float().mean()) < 0.9


def test_config2_all_64_views(nr):
    """The benchmarked configuration at its full size: teapot, 64 views, 512^2, RGBA, texture_size 4."""
    compare(nr, "cfg2")


def test_config1(nr):
    """examples_pytorch defaults: one view from (2.732, 30, 40), 256^2 with 2x anti-aliasing, 800x800 atlas."""
    compare(nr, "cfg1")


def test_config5_all_32_views(nr):
    """example3-style texture optimisation: orthographic camera, render_rgb with 2x anti-aliasing, the gradient
    reaches the texture PARAMETER through tanh (examples_pytorch/example3.py:40,54)."""
    compare(nr, "cfg5")


def test_config3_shared_mesh_through_renderer(nr):
    """100 352-face sphere, 8 views, ONE shared [1,nv,3] world-space parameter rendered through Renderer
    (fused camera transform), squared-error loss against the unperturbed sphere.  The oracle is fed the
    screen-space vertices the camera kernel produced (a vertex moved by one ulp can flip an edge pixel, which
    is a property of the camera arithmetic, not of the rasterizer), so the comparison is exact at that
    interface: face_index_map bit for bit, images, the gradient with respect to the screen-space vertices;
    the gradient of the shared parameter is then compared with torch-CPU autograd through look_at +
    perspective fed the oracle's screen-space gradient."""
    w = bench.WORKLOADS["cfg3"]
    inp = bench.make_inputs(w, 1000, "cpu", nr)
    B, S = w["views"], w["S"]
    faces = inp["faces"].to(DEV)
    eye = inp["eye"].to(DEV)
    tv, _ = bench.sphere_mesh(225, 0.0)
    hp = lambda: nr.RasterizeHyperparam(image_size=S, anti_aliasing=False)
    with torch.no_grad():
        target = nr.rasterize_silhouettes(nr.perspective(nr.look_at(tv.to(DEV)[None].expand(B, -1, -1), eye)), faces,
                                          nr.RasterizeParam(), hp())
    rend = nr.Renderer()
    rend.image_size, rend.anti_aliasing, rend.viewpoints = S, False, eye
    param = inp["v_world"][None].to(DEV).requires_grad_(True)
    vs = rend.transform_vertices(nr.parallel.share_across_views(param, B))
    vs.retain_grad()
    images = nr.rasterize_silhouettes(vs, faces, nr.RasterizeParam(), hp())
    ((images - target) ** 2).sum().backward()
    # the same step through the facade on the [1,nv,3] mesh itself (the camera kernels broadcast it over the
    # [B,3] viewpoints and sum its gradient over the views in registers) gives the same image bits
    param2 = inp["v_world"][None].to(DEV).requires_grad_(True)
    images2 = rend.render_silhouettes(nr.parallel.share_across_ranks(param2), faces)
    ((images2 - target) ** 2).sum().backward()
    assert param2.grad.shape == param2.shape
    assert torch.equal(images, images2)
    grad_close(param2.grad.cpu().numpy(), param.grad.cpu().numpy(), "shared-mesh camera vs expanded views", l2=1e-6)

    vs_cpu = vs.detach().cpu().clone().requires_grad_(True)
    oimg, omaps = ref.rasterize(vs_cpu, inp["faces"], S, False, draw_silhouettes=True, return_maps=True)
    ((oimg[:, 0] - target.cpu()) ** 2).sum().backward()
    fim = nr.rasterize_maps(vs.detach(), faces, nr.RasterizeParam(),
                            nr.RasterizeHyperparam(image_size=S, anti_aliasing=False, draw_rgb=False))["face_index_map"]
    assert int((fim.cpu() != omaps["face_index_map"]).sum()) == 0
    assert torch.equal(images.detach().cpu(), oimg[:, 0].detach())
    grad_close(vs.grad.cpu().numpy(), vs_cpu.grad.numpy(), "cfg3 grad screen vertices")
    # camera backward + sum over the views
    pc = inp["v_world"][None].clone().requires_grad_(True)
    nr.perspective(nr.look_at(pc.expand(B, -1, -1), inp["eye"])).backward(vs_cpu.grad)
    got, want = param.grad.cpu().numpy(), pc.grad.numpy()
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5


def test_config4_window_gradients(nr):
    """1 M random triangles at 1024^2, two views at full size.  Upstream gradient non-zero only inside a 64 x 64
    window: the vertex gradients of the faces around it must equal the oracle run on the faces whose bounding
    box touches the (1-pixel larger: the stencil reads the neighbours) surrounding 96 x 96 window."""
    w = dict(bench.WORKLOADS["cfg4"], views=2)
    inp = bench.make_inputs(w, 1000, torch.device(DEV), nr)
    S = w["S"]
    v = inp["vertices"].to(DEV).clone().requires_grad_(True)
    faces = inp["faces"].to(DEV)
    # at the rim of the cloud, where the silhouette has holes (the middle is 30 layers deep: no edges, no gradient)
    x0, y0, n, m = 400, 16, 96, 16                       # window and its margin to the inner window
    mask = torch.zeros((S, S))
    mask[y0 + m:y0 + n - m, x0 + m:x0 + n - m] = 1.
    G = (inp["G"][:2].cpu() * mask.flip(0, 1)[None]).contiguous()          # output orientation is flipped
    img = nr.rasterize_silhouettes(v, faces, nr.RasterizeParam(), nr.RasterizeHyperparam(image_size=S, anti_aliasing=False))
    img.backward(G.to(DEV))
    fv = v.detach()[:, faces.long()]
    lo_x, hi_x = (2 * x0 + 1 - S) / S, (2 * (x0 + n - 1) + 1 - S) / S
    lo_y, hi_y = (2 * y0 + 1 - S) / S, (2 * (y0 + n - 1) + 1 - S) / S
    for b in range(2):
        f = fv[b]
        keep = ((f[:, :, 0].max(1).values >= lo_x) & (f[:, :, 0].min(1).values <= hi_x) &
                (f[:, :, 1].max(1).values >= lo_y) & (f[:, :, 1].min(1).values <= hi_y))
        ids = torch.nonzero(keep)[:, 0].cpu()
        sub = f[ids.to(DEV)].cpu().reshape(1, -1, 3).clone().requires_grad_(True)      # faces = arange: 3 vertices each
        sub_faces = torch.arange(ids.numel() * 3, dtype=torch.int32).reshape(-1, 3)
        oimg = ref.rasterize(sub, sub_faces, S, False, draw_silhouettes=True)
        oimg.backward(G[b:b + 1, None])
        win = (slice(S - y0 - n + m, S - y0 - m), slice(S - x0 - n + m, S - x0 - m))         # inner window, flipped
        assert torch.equal(img[b][win].detach().cpu(), oimg[0, 0][win].detach())
        vid = (3 * ids[:, None] + torch.arange(3)[None]).reshape(-1)
        got = v.grad[b, vid.to(DEV)].cpu().numpy()
        want = sub.grad[0].numpy()
        assert np.abs(want).max() > 0
        grad_close(got, want, "cfg4 window view %d" % b)
        # nothing outside the window's faces moves
        other = torch.ones(v.shape[1], dtype=torch.bool)
        other[vid] = False
        assert float(v.grad[b, other.to(DEV)].abs().max()) == 0.
