"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star):
  * face_index_map: bit-exact;
  * weight_map / depth_map / images: the forward replays the reference arithmetic operation by
    operation, so they are compared with atol 1e-6 (images) and exactly (weight map);
  * gradients: |d| <= 1e-5 * |ref| + 1e-5 * max|ref|  (sums of thousands of float32 terms whose
    order differs between atomics, index_put and the CPU oracle; the reference's own CUDA
    index_put has the same run-to-run noise).
"""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import pipeline as ref

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


@pytest.fixture(scope="module")
def nr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import neural_renderer_v2_pytorch_b200 as nr_
    return nr_


def grad_close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.abs(want).max()
    tol = 1e-5 * np.abs(want) + 1e-5 * scale
    bad = np.abs(got - want) > tol
    assert not bad.any(), "%s: %d / %d beyond tolerance, max |d| = %.3g (scale %.3g)" % (
        what, bad.sum(), bad.size, np.abs(got - want).max(), scale)


def run_cuda(nr, d, dev="cuda:0"):
    mode = str(d["mode"])
    hp = nr.RasterizeHyperparam(image_size=int(d["image_size"]), near=float(d["near"]), far=float(d["far"]),
                                anti_aliasing=bool(d["anti_aliasing"]), draw_backside=bool(d["draw_backside"]))
    v = torch.from_numpy(d["vertices"]).to(dev).requires_grad_(True)
    faces = torch.from_numpy(d["faces"]).to(dev)
    tex = vt = None
    if mode in ("rgb", "rgba"):
        tex = torch.from_numpy(d["textures"]).to(dev).requires_grad_(True)
        vt = torch.from_numpy(d["vertices_textures"]).to(dev).requires_grad_(True)
        p = nr.RasterizeParam(vertices_textures=vt, faces_textures=torch.from_numpy(d["faces_textures"]).to(dev),
                              textures=tex)
    else:
        p = nr.RasterizeParam()
    fn = {"silhouettes": nr.rasterize_silhouettes, "rgb": nr.rasterize_rgb, "rgba": nr.rasterize_rgba,
          "depth": nr.rasterize_depth}[mode]
    images = fn(v, faces, p, hp)
    (images * torch.from_numpy(d["grad_images"]).to(dev)).sum().backward()
    hp2 = nr.RasterizeHyperparam(image_size=int(d["image_size"]), near=float(d["near"]), far=float(d["far"]),
                                 anti_aliasing=bool(d["anti_aliasing"]), draw_backside=bool(d["draw_backside"]),
                                 draw_rgb=mode in ("rgb", "rgba"), draw_silhouettes=mode in ("silhouettes", "rgba"),
                                 draw_depth=mode == "depth")
    maps = nr.rasterize_maps(v.detach(), faces, p, hp2)
    return images, v, tex, vt, maps


@pytest.mark.parametrize("name", CASES)
def test_golden_case(nr, name):
    """CUDA path vs fixtures produced by the reference's own Python code (tests/golden/make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    images, v, tex, vt, maps = run_cuda(nr, d)
    if "face_index_map" in d:
        assert np.array_equal(maps["face_index_map"].cpu().numpy(), d["face_index_map"]), "face_index_map not bit-exact"
        assert np.array_equal(maps["weight_map"].cpu().numpy(), d["weight_map"]), "weight_map differs"
    np.testing.assert_allclose(images.detach().cpu().numpy(), d["images"], rtol=1e-5, atol=1e-6)
    grad_close(v.grad.cpu().numpy(), d["grad_vertices"], "grad_vertices")
    if tex is not None:
        grad_close(tex.grad.cpu().numpy(), d["grad_textures"], "grad_textures")
        grad_close(vt.grad.cpu().numpy(), d["grad_vertices_textures"], "grad_vertices_textures")


@pytest.mark.parametrize("name", ["case_rgba_64", "case_rgba_aa_32", "case_depth_aa_32", "case_sil_cull_64"])
def test_golden_case_with_fine_tiles(nr, name):
    """The dense-mesh mode (general binning over 8x8 tiles, backward over all 16x16 tiles) against the
    same reference fixtures: images and every gradient."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    rz.FORCE_FINE_TILES = True
    try:
        images, v, tex, vt, maps = run_cuda(nr, d)
    finally:
        rz.FORCE_FINE_TILES = False
    assert np.array_equal(maps["face_index_map"].cpu().numpy(), d["face_index_map"])
    np.testing.assert_allclose(images.detach().cpu().numpy(), d["images"], rtol=1e-5, atol=1e-6)
    grad_close(v.grad.cpu().numpy(), d["grad_vertices"], "grad_vertices")
    if tex is not None:
        grad_close(tex.grad.cpu().numpy(), d["grad_textures"], "grad_textures")
        grad_close(vt.grad.cpu().numpy(), d["grad_vertices_textures"], "grad_vertices_textures")


@pytest.mark.parametrize("name", CASES)
def test_golden_case_with_dense_raster(nr, name):
    """The face-parallel raster kernel for meshes of small triangles (nr_raster_dense.cu: unsorted tile lists,
    shared-memory z-buffer, exact replay of contested pixels) against the same reference fixtures: face_index_map
    bit for bit, images and every gradient."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    rz.FORCE_DENSE_RASTER = True
    try:
        images, v, tex, vt, maps = run_cuda(nr, d)
    finally:
        rz.FORCE_DENSE_RASTER = None
    if "face_index_map" in d:
        assert np.array_equal(maps["face_index_map"].cpu().numpy(), d["face_index_map"]), "face_index_map not bit-exact"
        assert np.array_equal(maps["weight_map"].cpu().numpy(), d["weight_map"]), "weight_map differs"
    np.testing.assert_allclose(images.detach().cpu().numpy(), d["images"], rtol=1e-5, atol=1e-6)
    grad_close(v.grad.cpu().numpy(), d["grad_vertices"], "grad_vertices")
    if tex is not None:
        grad_close(tex.grad.cpu().numpy(), d["grad_textures"], "grad_textures")
        grad_close(vt.grad.cpu().numpy(), d["grad_vertices_textures"], "grad_vertices_textures")


@pytest.mark.parametrize("name", ["lit_rgb_48", "lit_rgb_aa_24"])
def test_lights_golden(nr, name):
    """Directional + Ambient + Specular lights (rasterize.py:252-283) fused into the kernels vs the
    reference's own lighting code: images and gradients to vertices (through the normals too) and textures."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    dev = "cuda:0"
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    lb = bool(d["light_backside"])
    lights = [nr.DirectionalLight(t("dir_color"), t("dir_direction"), backside=lb), nr.AmbientLight(t("amb_color")),
              nr.SpecularLight(t("spec_color"), t("spec_alpha"), backside=lb)]
    v = t("vertices").requires_grad_(True)
    tex = t("textures").requires_grad_(True)
    hp = nr.RasterizeHyperparam(image_size=int(d["image_size"]), anti_aliasing=bool(d["anti_aliasing"]),
                                draw_backside=bool(d["draw_backside"]))
    p = nr.RasterizeParam(vertices_textures=t("vertices_textures"), faces_textures=t("faces_textures"), textures=tex,
                          lights=lights)
    images = nr.rasterize_rgb(v, t("faces"), p, hp)
    (images * t("grad_images")).sum().backward()
    np.testing.assert_allclose(images.detach().cpu().numpy(), d["images"], rtol=1e-5, atol=2e-6)
    grad_close(tex.grad.cpu().numpy(), d["grad_textures"], "grad_textures")
    got, want = v.grad.cpu().numpy(), d["grad_vertices"]
    # the normal path goes through torch index_add (float atomics) and powf: 1e-4 of the scale
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()


def test_renderer_end_to_end(nr):
    """World-space vertices -> look_at -> perspective -> rasterize_rgba with AA, gradients back to
    world vertices and textures, vs the reference's Renderer.render run on CPU."""
    d = np.load(os.path.join(GOLDEN, "renderer_rgba_aa_32.npz"))
    dev = "cuda:0"
    r = nr.Renderer()
    r.image_size = int(d["image_size"])
    r.viewpoints = torch.from_numpy(d["viewpoints"]).to(dev)
    vw = torch.from_numpy(d["vertices_world"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(d["textures"]).to(dev).requires_grad_(True)
    images = r.render(vw, torch.from_numpy(d["faces"]).to(dev), torch.from_numpy(d["vertices_textures"]).to(dev),
                      torch.from_numpy(d["faces_textures"]).to(dev), tex)
    (images * torch.from_numpy(d["grad_images"]).to(dev)).sum().backward()
    # look_at / perspective run as torch CUDA ops here and torch CPU ops in the fixture: a vertex
    # that moves by 1 ulp can flip a pixel, so allow a handful of differing pixels
    diff = np.abs(images.detach().cpu().numpy() - d["images"])
    assert (diff > 1e-4).mean() < 2e-3, "too many differing pixels: %g" % (diff > 1e-4).mean()
    gv, gw = vw.grad.cpu().numpy(), d["grad_vertices_world"]
    assert np.abs(gv - gw).max() <= 2e-2 * np.abs(gw).max()


def test_renderer_end_to_end_tight(nr):
    """The same Renderer.render step with the cross-device rounding of the camera transform taken out: the oracle
    rasterizes the screen-space vertices the GPU's fused camera kernel produced, so images and the gradient that
    reaches the camera transform must agree at the usual bars (test_renderer_end_to_end has to allow flipped
    pixels); the camera kernels themselves are compared with torch in test_fused_camera_*."""
    from oracle import pipeline as ref
    d = np.load(os.path.join(GOLDEN, "renderer_rgba_aa_32.npz"))
    dev = "cuda:0"
    r = nr.Renderer()
    r.image_size = int(d["image_size"])
    r.viewpoints = torch.from_numpy(d["viewpoints"]).to(dev)
    faces = torch.from_numpy(d["faces"]).to(dev)
    vt, ft = torch.from_numpy(d["vertices_textures"]), torch.from_numpy(d["faces_textures"])
    G = torch.from_numpy(d["grad_images"])
    with torch.no_grad():
        screen = r.transform_vertices(torch.from_numpy(d["vertices_world"]).to(dev)).contiguous()
    # ours: rasterize_rgba on those vertices, as Renderer.render does behind the transform
    v1 = screen.clone().requires_grad_(True)
    t1 = torch.from_numpy(d["textures"]).to(dev).requires_grad_(True)
    img1 = nr.rasterize_rgba(v1, faces, nr.RasterizeParam(vertices_textures=vt.to(dev), faces_textures=ft.to(dev), textures=t1),
                             r._hyperparams())
    (img1 * G.to(dev)).sum().backward()
    # oracle: the reference's algorithm on the SAME screen-space vertices
    v0 = screen.cpu().clone().requires_grad_(True)
    t0 = torch.from_numpy(d["textures"]).clone().requires_grad_(True)
    img0 = ref.rasterize(v0, d["faces"], r.image_size, r.anti_aliasing, near=r.near, far=r.far, draw_backside=r.draw_backside,
                         draw_rgb=True, draw_silhouettes=True, vertices_textures=vt, faces_textures=ft.numpy(), textures=t0)
    (img0 * G).sum().backward()
    np.testing.assert_allclose(img1.detach().cpu().numpy(), img0.detach().numpy(), rtol=1e-5, atol=1e-6)
    grad_close(v1.grad.cpu().numpy(), v0.grad.numpy(), "grad_vertices (screen space)")
    grad_close(t1.grad.cpu().numpy(), t0.grad.numpy(), "grad_textures")
    # ... and the whole Renderer.render agrees with the two halves put together
    vw = torch.from_numpy(d["vertices_world"]).to(dev).requires_grad_(True)
    t2 = torch.from_numpy(d["textures"]).to(dev).requires_grad_(True)
    img2 = r.render(vw, faces, vt.to(dev), ft.to(dev), t2)
    (img2 * G.to(dev)).sum().backward()
    assert torch.equal(img2, img1)
    grad_close(t2.grad.cpu().numpy(), t0.grad.numpy(), "grad_textures through Renderer.render")


def test_reference_golden_png(nr):
    """The one golden IMAGE the reference ships for this path: tests_torch/data/4e4987...png, which
    tests_torch/test_save_obj.py:13-43 and tests_chainer/test_rasterize.py:43-72 compare with a render of
    the textured ShapeNet model next to it (Renderer defaults: 256^2, anti-aliasing; draw_backside False;
    viewpoint (2.5, 10, -90)) at atol = 1e-2.  The fixture holds that model as loaded by the reference's
    own load_obj (texture atlas quantised to 8 bit: 1.6e-3) and the golden pixels.

    The coverage (alpha) channel - camera transform, z-buffer, anti-aliasing - meets the reference's bar.
    The colours do not, for the reference's torch algorithm itself: the oracle, which reproduces the
    reference's torch code bit for bit (test_oracle.py), differs from this Chainer-rendered PNG in the same
    6 % of the colour values, so the colour channels are pinned against the oracle and only loosely
    (mean |d| < 0.05) against the PNG."""
    d = np.load(os.path.join(GOLDEN, "reference_golden_png_4e4987.npz"))
    dev = "cuda:0"
    r = nr.Renderer()
    r.draw_backside = False
    r.viewpoints = [float(x) for x in d["viewpoint"]]
    v = torch.from_numpy(d["vertices"])[None]
    vt = torch.from_numpy(d["vertices_t"])[None]
    tex = torch.from_numpy(d["textures"].astype(np.float32) / 255.)[None]
    images = r.render(v.to(dev), torch.from_numpy(d["faces"]).to(dev), vt.to(dev), torch.from_numpy(d["faces_t"]).to(dev),
                      tex.to(dev))
    image = images[0].permute(1, 2, 0).cpu().numpy()
    want = d["golden_png"].astype(np.float32) / 255.
    assert image.shape == want.shape == (256, 256, 4)
    np.testing.assert_allclose(want[..., 3], image[..., 3], atol=1e-2)
    assert 0.05 < want[..., 3].mean() < 0.9
    assert np.abs(image[..., :3] - want[..., :3]).mean() < 0.05
    # the reference's torch algorithm on the same inputs
    vs = nr.perspective(nr.look_at(v, torch.tensor(r.viewpoints)[None]))
    ora = ref.rasterize(vs, d["faces"], 256, True, draw_backside=False, draw_rgb=True, draw_silhouettes=True,
                        vertices_textures=vt, faces_textures=d["faces_t"], textures=tex)[0].permute(1, 2, 0).numpy()
    # (camera transform on the GPU vs on the CPU: a vertex moved by an ulp can flip an edge pixel)
    assert (np.abs(image - ora) > 1e-4).mean() < 2e-3
    bad_ours = np.abs(image[..., :3] - want[..., :3]) > 1e-2
    bad_oracle = np.abs(ora[..., :3] - want[..., :3]) > 1e-2
    assert abs(bad_ours.mean() - bad_oracle.mean()) < 2e-3 and (bad_ours != bad_oracle).mean() < 2e-3


@pytest.mark.parametrize("mode,persp", [("look_at", True), ("look_at", False), ("look", True)])
def test_fused_camera_transform_matches_torch_ops(nr, mode, persp):
    """camera.transform_vertices (one kernel each way) vs the look_at / look + perspective torch ops:
    screen vertices to 1e-6, gradients to vertices and to the viewpoints to 1e-4."""
    g = torch.Generator().manual_seed(3)
    B, nv = 5, 777
    v0 = (torch.rand((B, nv, 3), generator=g) - 0.5)
    e0 = nr.get_points_from_angles(torch.full((B,), 2.7), torch.rand(B, generator=g) * 60 - 20, torch.rand(B, generator=g) * 360)
    G = torch.randn((B, nv, 3), generator=g).cuda()
    outs = []
    for fused in (True, False):
        v = v0.clone().cuda().requires_grad_(True)
        e = e0.clone().cuda().requires_grad_(True)
        r = nr.Renderer()
        r.camera_mode, r.perspective, r.viewpoints, r.fused_camera = mode, persp, e, fused
        r.camera_direction = [0.1, -0.2, 1.0]
        out = r.transform_vertices(v)
        (out * G).sum().backward()
        outs.append((out.detach(), v.grad, e.grad))
    (o1, gv1, ge1), (o0, gv0, ge0) = outs
    assert torch.allclose(o1, o0, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gv1, gv0, rtol=1e-4, atol=1e-4 * gv0.abs().max().item())
    assert torch.allclose(ge1, ge0, rtol=1e-3, atol=1e-4 * ge0.abs().max().item())


def _fim_cuda(nr, faces_np, R, near=0.1, far=100.0, backside=True):
    f = torch.from_numpy(np.ascontiguousarray(faces_np, dtype=np.float32)).cuda()
    B, nf = f.shape[:2]
    fim = torch.full((B * R * R,), -7, dtype=torch.int32, device="cuda")
    out = nr.face_index_map_forward_safe(f, fim, nf, R, near, far, int(backside), 1e-8, 1e-4)
    assert out.data_ptr() == fim.data_ptr()
    wm = torch.zeros((B * R * R, 3), dtype=torch.float32, device="cuda")
    nr.compute_weight_map_c(f, fim, wm, nf, R)
    return fim.reshape(B, R, R).cpu().numpy(), wm.reshape(B, R, R, 3).cpu().numpy()


def _check_vs_oracle(nr, faces_np, R, **kw):
    fim, wm = _fim_cuda(nr, faces_np, R, **kw)
    want = oracle.face_index_map(faces_np, R, kw.get("near", 0.1), kw.get("far", 100.0), kw.get("backside", True))
    nbad = int((fim != want).sum())
    assert nbad == 0, "face_index_map: %d / %d pixels differ" % (nbad, fim.size)
    want_wm = oracle.weight_map(faces_np, want)
    assert np.array_equal(wm, want_wm), "weight_map differs"
    # the fused forward itself, with both binning paths (one kernel per view / general multi-kernel)
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    B, nf = faces_np.shape[:2]
    v = torch.from_numpy(np.ascontiguousarray(faces_np, dtype=np.float32)).reshape(B, nf * 3, 3).cuda()
    idx = torch.arange(nf * 3, dtype=torch.int32).reshape(nf, 3).cuda()
    for general, fine, dense in ((False, False, False), (True, False, False), (True, True, False), (True, False, True)):
        rz.FORCE_GENERAL_BINNING, rz.FORCE_FINE_TILES, rz.FORCE_DENSE_RASTER = general, fine, dense
        try:
            hp = nr.RasterizeHyperparam(image_size=R, near=kw.get("near", 0.1), far=kw.get("far", 100.0), anti_aliasing=False,
                                        draw_backside=kw.get("backside", True), draw_rgb=False, draw_depth=False)
            maps = nr.rasterize_maps(v, idx, nr.RasterizeParam(), hp)
        finally:
            rz.FORCE_GENERAL_BINNING = rz.FORCE_FINE_TILES = False
            rz.FORCE_DENSE_RASTER = None
        got = maps["face_index_map"].cpu().numpy()
        assert np.array_equal(got, want), "fused forward (general=%s, fine=%s, dense=%s): face_index_map differs in %d px" % (
            general, fine, dense, (got != want).sum())
        assert np.array_equal(maps["weight_map"].cpu().numpy(), want_wm), "fused forward (general=%s, fine=%s, dense=%s): weight_map" % (general, fine, dense)
        assert np.array_equal(maps["images"][:, 0].flip(1, 2).cpu().numpy(), (want >= 0).astype(np.float32))
    return fim


def random_triangles(B, nf, seed, size=0.05, zlo=1.0, zhi=3.0):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-1.1, 1.1, size=(B, nf, 1, 2))
    xy = c + rng.normal(0, size, size=(B, nf, 3, 2))
    z = rng.uniform(zlo, zhi, size=(B, nf, 1, 1)) + rng.normal(0, 0.02, size=(B, nf, 3, 1))
    return np.concatenate([xy, z], -1).astype(np.float32)


@pytest.mark.parametrize("R", [16, 64, 100, 250, 256])
def test_random_small_triangles(nr, R):
    """Reference-signature operators vs the C oracle; R = 100 / 250 exercise partial tiles."""
    _check_vs_oracle(nr, random_triangles(2, 3000, R), R)


def test_random_small_triangles_cull(nr):
    _check_vs_oracle(nr, random_triangles(2, 3000, 5), 128, backside=False)


def test_large_triangles_cover_every_tile(nr):
    """Faces spanning the whole screen land in every tile list (many tiles per face)."""
    f = random_triangles(2, 200, 11, size=1.5)
    _check_vs_oracle(nr, f, 128)


def test_hysteresis_is_order_dependent(nr):
    """Three coplanar-ish full-screen faces 5e-5 apart in depth: the reference's sequential
    z-test (rasterize_cuda_kernel.cu:145) keeps the FIRST face in index order unless a later one
    is closer by more than 1e-4; a (min depth, min index) rule would give a different answer."""
    def quad(z):
        return [[-2., -2., z], [2., -2., z], [0., 2., z]]
    a = np.array([[quad(1.0), quad(0.99995), quad(0.9999)]], np.float32)
    b = np.array([[quad(0.9999), quad(0.99995), quad(1.0)]], np.float32)
    c = np.array([[quad(1.0), quad(0.9998), quad(0.99975)]], np.float32)
    for faces_np in (a, b, c):
        _check_vs_oracle(nr, faces_np, 32)
    fim_a = _check_vs_oracle(nr, a, 32)
    fim_c = _check_vs_oracle(nr, c, 32)
    assert set(np.unique(fim_a)) <= {-1, 0, 2} and set(np.unique(fim_c)) <= {-1, 1}


def test_duplicate_and_degenerate_faces(nr):
    f = random_triangles(1, 500, 3, size=0.2)
    f = np.concatenate([f, f[:, :100]], 1)          # exact duplicates later in the list never win
    f[0, 7] = f[0, 7, 0]                             # zero-area face
    f[0, 8, :, :2] = np.array([[0, 0], [0.5, 0.5], [1, 1]], np.float32)   # collinear
    _check_vs_oracle(nr, f, 96)


def test_non_finite_vertices(nr):
    f = random_triangles(1, 400, 4, size=0.3)
    f[0, 3, 0, 0] = np.nan
    f[0, 9, 1, 1] = np.inf
    f[0, 15, 2, 0] = -np.inf
    f[0, 21, 0, 2] = np.nan      # NaN depth: stays in the lists, never wins
    f[0, 27, 1, 2] = np.inf      # infinite depth at one corner: finite zp, can win
    f[0, 33, 2, 2] = 0.0         # z = 0
    f[0, 39, :, 2] = -1.0        # behind the camera
    _check_vs_oracle(nr, f, 96)


def test_vertices_on_pixel_centres(nr):
    """Edges through pixel centres make edge functions exactly zero; the reference then accepts
    every pixel of the bounding box lying on the line through v1-v2 (c2 == 0 passes both product
    tests, rasterize_cuda_kernel.cu:109,114).  The tile / block culling must keep those pixels."""
    R = 64
    rng = np.random.RandomState(12)
    for size in (6, 20, 40):
        centre = rng.randint(0, R, size=(2, 600, 1, 2))
        px = np.clip(centre + rng.randint(-size, size + 1, size=(2, 600, 3, 2)), -8, R + 8)
        xy = (2. * px + 1 - R) / R
        z = rng.uniform(1., 3., size=(2, 600, 3, 1))
        f = np.concatenate([xy, z], -1).astype(np.float32)
        f[:, :50, 1, 0] = f[:, :50, 2, 0]          # vertical edge v1-v2 on a pixel column
        f[:, 50:100, 1, 1] = f[:, 50:100, 2, 1]    # horizontal edge v1-v2 on a pixel row
        _check_vs_oracle(nr, f, R)
        _check_vs_oracle(nr, f, R, backside=False)


def test_near_far_clipping(nr):
    f = random_triangles(2, 2000, 6, size=0.1, zlo=0.05, zhi=4.0)
    _check_vs_oracle(nr, f, 128, near=1.0, far=2.5)


def test_long_tile_lists(nr):
    """> 1024 faces in one tile: global-memory sort path and multi-chunk staging."""
    rng = np.random.RandomState(8)
    f = random_triangles(1, 5000, 8, size=0.02)
    f[..., :2] = f[..., :2] * 0.05 + rng.uniform(-0.02, 0.02)     # everything inside one tile
    _check_vs_oracle(nr, f, 64)


@pytest.mark.parametrize("nf", [130, 300, 700, 1500])
def test_mid_size_tile_lists(nr, nf):
    """130 .. 1500 faces in one tile: every size class of the per-tile sort (registers up to 256 ids,
    one warp in shared memory up to 1024, the whole CTA beyond) in all three binning modes."""
    rng = np.random.RandomState(nf)
    f = random_triangles(2, nf, nf, size=0.02)
    f[..., :2] = f[..., :2] * 0.05 + rng.uniform(-0.02, 0.02)     # everything inside one 16x16 tile
    _check_vs_oracle(nr, f, 64)


def test_pair_list_overflow_falls_back_on_device(nr):
    """With no room for the (tile, face) pairs the forward must still be exact (every block scans all
    faces of its view) and the host must grow the capacity from the lazily read statistics."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    d = np.load(os.path.join(GOLDEN, "case_rgba_64.npz"))
    rz.FORCE_PAIR_CAPACITY = 0
    try:
        images, v, tex, vt, maps = run_cuda(nr, d)
    finally:
        rz.FORCE_PAIR_CAPACITY = None
    assert np.array_equal(maps["face_index_map"].cpu().numpy(), d["face_index_map"])
    np.testing.assert_allclose(images.detach().cpu().numpy(), d["images"], rtol=1e-5, atol=1e-6)
    grad_close(v.grad.cpu().numpy(), d["grad_vertices"], "grad_vertices")
    torch.cuda.synchronize()
    sc = rz._Scratch.get(torch.device("cuda", 0), torch.cuda.current_stream().cuda_stream)
    sc.poll(block=True)
    assert sc.overflows >= 1 and sc.pair_capacity > 0
    images, v, tex, vt, maps = run_cuda(nr, d)      # regular path again
    assert np.array_equal(maps["face_index_map"].cpu().numpy(), d["face_index_map"])


def test_small_mesh_binning_outgrown_switches_to_the_general_path(nr):
    """<= 8192 faces take the one-kernel binning (shared-memory pair lists).  200 screen-filling faces at
    512^2 make ~200k pairs per view, more than it holds: that call must still be exact (device-side
    fallback) and the host must bin this shape with the general path afterwards."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    f = random_triangles(2, 200, 21, size=1.5)
    R = 512
    want = oracle.face_index_map(f, R, 0.1, 100.0, True)
    v = torch.from_numpy(f).reshape(2, 600, 3).cuda()
    idx = torch.arange(600, dtype=torch.int32).reshape(200, 3).cuda()
    sc = rz._Scratch.get(torch.device("cuda", 0), torch.cuda.current_stream().cuda_stream)
    sc.general_binning.discard((200, R))
    for call in range(2):
        hp = nr.RasterizeHyperparam(image_size=R, anti_aliasing=False, draw_rgb=False, draw_depth=False)
        maps = nr.rasterize_maps(v, idx, nr.RasterizeParam(), hp)
        assert np.array_equal(maps["face_index_map"].cpu().numpy(), want), "call %d" % call
        torch.cuda.synchronize()
        sc.poll(block=True)
        assert (200, R) in sc.general_binning


def test_deterministic_mode_is_bit_reproducible(nr):
    """hyperparams.deterministic: gradients bit-identical from run to run, and still within the
    tolerance of the reference's gradients."""
    d = np.load(os.path.join(GOLDEN, "case_rgba_64.npz"))
    dev = "cuda:0"

    def run():
        hp = nr.RasterizeHyperparam(image_size=int(d["image_size"]), anti_aliasing=False)
        hp.deterministic = True
        v = torch.from_numpy(d["vertices"]).to(dev).requires_grad_(True)
        tex = torch.from_numpy(d["textures"]).to(dev).requires_grad_(True)
        vt = torch.from_numpy(d["vertices_textures"]).to(dev).requires_grad_(True)
        p = nr.RasterizeParam(vertices_textures=vt, faces_textures=torch.from_numpy(d["faces_textures"]).to(dev), textures=tex)
        img = nr.rasterize_rgba(v, torch.from_numpy(d["faces"]).to(dev), p, hp)
        (img * torch.from_numpy(d["grad_images"]).to(dev)).sum().backward()
        return v.grad.clone(), tex.grad.clone(), vt.grad.clone()

    first = run()
    for _ in range(5):
        again = run()
        for a, b in zip(first, again):
            assert torch.equal(a, b), "deterministic mode produced different bits"
    grad_close(first[0].cpu().numpy(), d["grad_vertices"], "grad_vertices")
    grad_close(first[1].cpu().numpy(), d["grad_textures"], "grad_textures")
    grad_close(first[2].cpu().numpy(), d["grad_vertices_textures"], "grad_vertices_textures")


def test_teapot_views(nr):
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    B = 4
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    g = torch.Generator().manual_seed(5)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20,
                                    torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye))
    faces_np = vs[:, torch.from_numpy(d["faces"]).long()].numpy()
    _check_vs_oracle(nr, faces_np, 256)
    vs[1] = 0                    # an all-zero view, as in tests_torch/test_rasterize.py:21-23
    _check_vs_oracle(nr, vs[:, torch.from_numpy(d["faces"]).long()].numpy(), 128)


def test_empty_inputs(nr):
    """No faces / everything off-screen -> all background."""
    fim, wm = _fim_cuda(nr, np.zeros((2, 1, 3, 3), np.float32), 32)
    assert (fim == -1).all() and (wm == 0).all()
    f = random_triangles(1, 50, 2) + np.array([5., 5., 0.], np.float32)
    fim, _ = _fim_cuda(nr, f, 32)
    assert (fim == -1).all()


def test_no_faces_and_double_precision_inputs(nr):
    """nf = 0 renders pure background with zero gradients; float64 inputs are computed in float32
    and get float64 gradients back."""
    hp = nr.RasterizeHyperparam(image_size=40, anti_aliasing=True)
    v = torch.rand(2, 5, 3, device="cuda", requires_grad=True)
    img = nr.rasterize_silhouettes(v, torch.zeros((0, 3), dtype=torch.int32, device="cuda"), nr.RasterizeParam(), hp)
    assert img.shape == (2, 40, 40) and float(img.abs().max()) == 0.0
    img.sum().backward()
    assert float(v.grad.abs().max()) == 0.0
    d = np.load(os.path.join(GOLDEN, "case_sil_64.npz"))
    v64 = torch.from_numpy(d["vertices"]).double().cuda().requires_grad_(True)
    hp = nr.RasterizeHyperparam(image_size=64, anti_aliasing=False)
    img = nr.rasterize_silhouettes(v64, torch.from_numpy(d["faces"]).long().cuda(), nr.RasterizeParam(), hp)
    (img * torch.from_numpy(d["grad_images"]).cuda()).sum().backward()
    assert v64.grad.dtype == torch.float64
    grad_close(v64.grad.cpu().numpy(), d["grad_vertices"], "grad_vertices (float64 input)")


def test_interleaved_forwards_and_repeated_backward(nr):
    """Two forwards share the cached workspace before either backward runs; backward twice through the
    same graph (retain_graph) gives the same gradient (everything the backward needs is saved per call)."""
    d1 = np.load(os.path.join(GOLDEN, "case_rgba_64.npz"))
    d2 = np.load(os.path.join(GOLDEN, "case_sil_cull_64.npz"))
    dev = "cuda:0"
    v1 = torch.from_numpy(d1["vertices"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(d1["textures"]).to(dev).requires_grad_(True)
    p1 = nr.RasterizeParam(vertices_textures=torch.from_numpy(d1["vertices_textures"]).to(dev),
                           faces_textures=torch.from_numpy(d1["faces_textures"]).to(dev), textures=tex)
    i1 = nr.rasterize_rgba(v1, torch.from_numpy(d1["faces"]).to(dev), p1, nr.RasterizeHyperparam(image_size=64, anti_aliasing=False))
    v2 = torch.from_numpy(d2["vertices"]).to(dev).requires_grad_(True)
    i2 = nr.rasterize_silhouettes(v2, torch.from_numpy(d2["faces"]).to(dev), nr.RasterizeParam(),
                                  nr.RasterizeHyperparam(image_size=64, anti_aliasing=False, draw_backside=False))
    loss = (i1 * torch.from_numpy(d1["grad_images"]).to(dev)).sum() + (i2 * torch.from_numpy(d2["grad_images"]).to(dev)).sum()
    loss.backward(retain_graph=True)
    grad_close(v1.grad.cpu().numpy(), d1["grad_vertices"], "grad_vertices of the first render")
    grad_close(v2.grad.cpu().numpy(), d2["grad_vertices"], "grad_vertices of the second render")
    grad_close(tex.grad.cpu().numpy(), d1["grad_textures"], "grad_textures")
    loss.backward()                               # accumulates: exactly twice the gradient
    grad_close(v1.grad.cpu().numpy() / 2, d1["grad_vertices"], "second backward through the same graph")


def test_capture_step_replays_the_whole_step(nr):
    """nr.capture_step: a CUDA-graph replay recomputes forward + backward on the current contents of
    the static inputs."""
    d = np.load(os.path.join(GOLDEN, "case_sil_64.npz"))
    dev = "cuda:0"
    v = torch.from_numpy(d["vertices"]).to(dev).requires_grad_(True)
    faces = torch.from_numpy(d["faces"]).to(dev)
    G = torch.from_numpy(d["grad_images"]).to(dev)

    def step():
        img = nr.rasterize_silhouettes(v, faces, nr.RasterizeParam(), nr.RasterizeHyperparam(image_size=64, anti_aliasing=False))
        img.backward(G)
        return img

    replay = nr.capture_step(step, params=[v], warmup=2)
    img = replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(img.detach().cpu().numpy(), d["images"], atol=1e-6)
    grad_close(v.grad.cpu().numpy(), d["grad_vertices"], "replayed gradient")
    with torch.no_grad():
        v.mul_(0.5)                               # new input in the same static buffer
    img2 = replay().detach().clone()
    want = nr.rasterize_silhouettes(v.detach(), faces, nr.RasterizeParam(), nr.RasterizeHyperparam(image_size=64, anti_aliasing=False))
    assert torch.equal(img2, want)


def test_fused_matches_oracle_pipeline_midsize(nr):
    """Fused forward + backward vs the torch-CPU oracle at 128^2, RGBA + depth in one call."""
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    B, S, ts = 3, 128, 3
    g = torch.Generator().manual_seed(42)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20,
                                    torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye))
    faces = torch.from_numpy(d["faces"])
    vt_np, ft_np, tex_np = nr.create_textures(faces.shape[0], ts)
    tex0 = torch.rand((B,) + tex_np.shape, generator=g)
    vt0 = torch.from_numpy(vt_np)[None].repeat(B, 1, 1)
    G = torch.randn((B, 5, S, S), generator=g)

    v0 = vs.clone().requires_grad_(True)
    t0 = tex0.clone().requires_grad_(True)
    img0 = ref.rasterize(v0, faces, S, False, draw_rgb=True, draw_silhouettes=True, draw_depth=True,
                         vertices_textures=vt0, faces_textures=ft_np, textures=t0)
    (img0 * G).sum().backward()

    v1 = vs.clone().cuda().requires_grad_(True)
    t1 = tex0.clone().cuda().requires_grad_(True)
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=False, draw_rgb=True, draw_silhouettes=True,
                                draw_depth=True)
    p = nr.RasterizeParam(vertices_textures=vt0.cuda(), faces_textures=torch.from_numpy(ft_np).cuda(), textures=t1)
    img1 = nr.rasterize_core(v1, faces.cuda(), p, hp)
    (img1 * G.cuda()).sum().backward()
    np.testing.assert_allclose(img1.detach().cpu().numpy(), img0.detach().numpy(), rtol=1e-5, atol=1e-6)
    grad_close(v1.grad.cpu().numpy(), v0.grad.numpy(), "grad_vertices")
    grad_close(t1.grad.cpu().numpy(), t0.grad.numpy(), "grad_textures")


@pytest.mark.parametrize("S", [40, 36])
def test_forward_zero_fills_the_gradient_accumulators(nr, S):
    """nrZeroFill: buffers handed to the forward come back zero (whatever their size modulo 16 bytes),
    every output element is written, and gradients of odd-sized textures still match the oracle.
    S = 36 (internal resolution 72, no multiple of the tile size) takes the scalar fill path."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    B = 2
    g = torch.Generator().manual_seed(9)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20, torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye)).cuda()
    faces = torch.from_numpy(d["faces"]).cuda()
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=True, draw_rgb=False, draw_depth=False)
    cfg, faces_d, *_ = rz._prepare(vs, faces, nr.RasterizeParam(), hp)
    bufs = [torch.full((n,), float("nan"), device="cuda") for n in (1, 3, 4, 1021, 4096 * 3 + 2)]
    for group in (bufs[:4], bufs[4:]):
        images, internal, fim, _, _, _, _ = rz._forward_call(cfg, vs.contiguous(), faces_d, None, None, None, False, zero=group)
        for t in group:
            assert float(t.abs().sum()) == 0.0 and not torch.isnan(t).any()
        assert not torch.isnan(images).any() and not torch.isnan(internal).any()
        assert int(fim.min()) >= -1 and int(fim.max()) < faces.shape[0]
    # textures whose byte size is not a multiple of 16
    ts_faces = faces.shape[0]
    vt_np, ft_np, tex_np = nr.create_textures(ts_faces, 1)
    H, W = tex_np.shape[1] + 1, tex_np.shape[2] + 3          # 2 * 3 * 51 * 53 * 4 bytes = 8 (mod 16)
    tex0 = torch.rand((B, 3, H, W), generator=g)
    assert (tex0.numel() * 4) % 16 != 0
    vt0 = torch.from_numpy(vt_np)[None].repeat(B, 1, 1)
    G = torch.randn((B, 4, S, S), generator=g)
    v0 = vs.cpu().clone().requires_grad_(True)
    t0 = tex0.clone().requires_grad_(True)
    (ref.rasterize(v0, faces.cpu(), S, True, draw_rgb=True, draw_silhouettes=True, vertices_textures=vt0,
                   faces_textures=ft_np, textures=t0) * G).sum().backward()
    v1 = vs.clone().requires_grad_(True)
    t1 = tex0.clone().cuda().requires_grad_(True)
    p = nr.RasterizeParam(vertices_textures=vt0.cuda(), faces_textures=torch.from_numpy(ft_np).cuda(), textures=t1)
    (nr.rasterize_rgba(v1, faces, p, nr.RasterizeHyperparam(image_size=S, anti_aliasing=True)) * G.cuda()).sum().backward()
    grad_close(v1.grad.cpu().numpy(), v0.grad.numpy(), "grad_vertices")
    grad_close(t1.grad.cpu().numpy(), t0.grad.numpy(), "grad_textures")


@pytest.mark.parametrize("aa,mode", [(False, "picture"), (True, "picture"), (True, "color")])
def test_backgrounds(nr, aa, mode):
    """Backgrounds.  The reference's blend_backgrounds fails on torch tensors
    (rasterize.py:156-159), so the oracle restates the Chainer original: no reference output pins this
    ("parity unpinned").  Checks images, and gradients to vertices (the stencil sees the background
    colours at silhouette edges), textures and the background picture itself."""
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    B, S, ts = 2, 48, 2
    R = 2 * S if aa else S
    g = torch.Generator().manual_seed(7)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20,
                                    torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye))
    faces = torch.from_numpy(d["faces"])
    vt_np, ft_np, tex_np = nr.create_textures(faces.shape[0], ts)
    tex0 = torch.rand((B,) + tex_np.shape, generator=g)
    vt0 = torch.from_numpy(vt_np)[None].repeat(B, 1, 1)
    G = torch.randn((B, 4, S, S), generator=g)
    color = [0.25, 0.5, 0.75]
    if mode == "picture":
        bg0 = torch.rand((B, 3, R, R), generator=g)
    else:
        bg0 = torch.tensor(color)[None, :, None, None].expand(B, 3, R, R).contiguous()

    v0 = vs.clone().requires_grad_(True)
    t0 = tex0.clone().requires_grad_(True)
    b0 = bg0.clone().requires_grad_(True)
    img0 = ref.rasterize(v0, faces, S, aa, draw_rgb=True, draw_silhouettes=True, vertices_textures=vt0,
                         faces_textures=ft_np, textures=t0, backgrounds=b0)
    (img0 * G).sum().backward()

    v1 = vs.clone().cuda().requires_grad_(True)
    t1 = tex0.clone().cuda().requires_grad_(True)
    b1 = bg0.clone().cuda().requires_grad_(True)
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=aa)
    kw = dict(backgrounds=b1) if mode == "picture" else dict(background_color=color)
    p = nr.RasterizeParam(vertices_textures=vt0.cuda(), faces_textures=torch.from_numpy(ft_np).cuda(), textures=t1, **kw)
    img1 = nr.rasterize_rgba(v1, faces.cuda(), p, hp)
    (img1 * G.cuda()).sum().backward()
    assert hp.image_size == S
    np.testing.assert_allclose(img1.detach().cpu().numpy(), img0.detach().numpy(), rtol=1e-5, atol=1e-6)
    grad_close(v1.grad.cpu().numpy(), v0.grad.numpy(), "grad_vertices")
    grad_close(t1.grad.cpu().numpy(), t0.grad.numpy(), "grad_textures")
    if mode == "picture":
        grad_close(b1.grad.cpu().numpy(), b0.grad.numpy(), "grad_backgrounds")
    # silhouettes / depth never see a background (rasterize.py:286-288 is inside the rgb branch)
    sil = nr.rasterize_silhouettes(v1.detach(), faces.cuda(), p, nr.RasterizeHyperparam(image_size=S, anti_aliasing=aa))
    assert torch.equal(sil, img1[:, 3].detach())


@pytest.mark.parametrize("C", [1, 3, 4])
def test_differentiation_known_answer(nr, C):
    """Standalone differentiation() op vs the reference's Differentiation.backward output."""
    d = np.load(os.path.join(GOLDEN, "diff_known_answer_c%d.npz" % C))
    images = torch.from_numpy(d["images"]).cuda()
    coords = torch.zeros(images.shape[:3] + (2,), device="cuda", requires_grad=True)
    out = nr.differentiation(images, coords)
    assert torch.equal(out, images)
    (out * torch.from_numpy(d["grad_output"]).cuda()).sum().backward()
    np.testing.assert_allclose(coords.grad.cpu().numpy(), d["grad_coordinates"], rtol=1e-5, atol=1e-5)


def test_differentiation_binary_image(nr):
    d = np.load(os.path.join(GOLDEN, "diff_known_answer_binary.npz"))
    images = torch.from_numpy(d["images"]).cuda()
    coords = torch.zeros(images.shape[:3] + (2,), device="cuda", requires_grad=True)
    (nr.differentiation(images, coords) * torch.from_numpy(d["grad_output"]).cuda()).sum().backward()
    np.testing.assert_allclose(coords.grad.cpu().numpy(), d["grad_coordinates"], rtol=1e-5, atol=1e-5)


def test_differentiation_finite_difference_property(nr):
    """The reference's own known-answer check (tests_torch/test_differentiation.py:31-65): |grad| equals
    the one-pixel-shift finite difference, rtol 1e-4."""
    rng = np.random.RandomState(0)
    B, S = 4, 32
    images = torch.from_numpy(rng.normal(size=(B, S, S, 3)).astype("float32")).cuda()
    noise = torch.from_numpy(rng.normal(size=(B, S, S, 3)).astype("float32")).cuda()
    coords = torch.zeros((B, S, S, 2), device="cuda", requires_grad=True)
    (nr.differentiation(images, coords) * noise).sum().backward()
    g = coords.grad
    step = 2. / S
    for _ in range(50):
        yi, xi = rng.randint(1, S - 1), rng.randint(1, S - 1)
        for axis, col in ((1, 1), (2, 0)):
            def shifted(sign):
                im = images.clone()
                if axis == 1:
                    im[:, yi - sign, xi] = images[:, yi, xi]
                    im[:, yi, xi] = images[:, yi + sign, xi]
                else:
                    im[:, yi, xi - sign] = images[:, yi, xi]
                    im[:, yi, xi] = images[:, yi, xi + sign]
                gg = ((im - images) * noise).sum((1, 2, 3)) / step
                return torch.min(gg, torch.zeros_like(gg))
            want = torch.max(shifted(1).abs(), shifted(-1).abs())
            assert torch.allclose(want, g[:, yi, xi, col].abs(), rtol=1e-4, atol=1e-5)


def test_errors(nr):
    hp = nr.RasterizeHyperparam(image_size=32, anti_aliasing=False)
    v = torch.zeros(1, 3, 3)
    f = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        nr.rasterize_silhouettes(v, f, nr.RasterizeParam(), hp)
    with pytest.raises(AssertionError):
        nr.rasterize_silhouettes(v.cuda()[0], f, nr.RasterizeParam(), hp)
    with pytest.raises(IndexError):
        nr.rasterize_silhouettes(v.cuda(), torch.tensor([[0, 1, 9]], dtype=torch.int32), nr.RasterizeParam(), hp)
    with pytest.raises(RuntimeError, match="contiguous"):
        nr.face_index_map_forward_safe(torch.zeros(1, 9, 1, device="cuda").permute(0, 2, 1).reshape(1, 1, 3, 3).transpose(2, 3),
                                       torch.zeros(32 * 32, dtype=torch.int32, device="cuda"), 1, 32, 0.1, 100., 1, 1e-8, 1e-4)


def test_square_optimisation_converges(nr):
    """tests_torch/test_rasterize.py:205-249 in spirit: a 2-triangle square must be pulled onto a target
    silhouette by Adam using the approximate gradients (IoU loss < 0.05 within 300 steps)."""
    dev = "cuda:0"
    S = 64
    target = torch.zeros(S, S, device=dev)
    target[16:40, 24:56] = 1.
    vertices = torch.tensor([[0.1, 0.1, 1.], [-0.1, 0.1, 1.], [-0.1, -0.1, 1.], [0.1, -0.1, 1.]], device=dev)
    vertices = torch.nn.Parameter(vertices)
    faces = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32, device=dev)
    opt = torch.optim.Adam([vertices], lr=0.01)
    loss = None
    for _ in range(300):
        hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=False)
        image = nr.rasterize_silhouettes(vertices[None], faces, nr.RasterizeParam(), hp)[0]
        loss = 1 - torch.sum(image * target) / torch.sum(image + target - image * target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if loss.item() < 0.05:
            break
    assert loss.item() < 0.05, "did not converge: %g" % loss.item()


def test_reference_backward_case1_with_its_own_target(nr):
    """tests_torch/test_rasterize.py:205-249 verbatim: the 2-triangle square, target = 1 - gradient.png
    (the reference's own fixture), 256^2 without anti-aliasing, Adam lr 0.005: the IoU loss must fall
    below 0.01 within 350 iterations."""
    dev = "cuda:0"
    ref_img = np.load(os.path.join(GOLDEN, "reference_gradient_png.npz"))["gradient_png"].astype(np.float32) / 255.
    target = torch.from_numpy(1 - ref_img[:, :, 0]).to(dev)
    vertices = torch.nn.Parameter(torch.tensor([[0.1, 0.1, 1.], [-0.1, 0.1, 1.], [-0.1, -0.1, 1.], [0.1, -0.1, 1.]], device=dev))
    faces = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32, device=dev)
    opt = torch.optim.Adam([vertices], lr=0.005)
    iou = None
    for i in range(350):
        hp = nr.RasterizeHyperparam(image_size=256, anti_aliasing=False)
        image = nr.rasterize_silhouettes(vertices[None], faces, nr.RasterizeParam(), hp)[0]
        iou = 1 - torch.sum(image * target) / torch.sum(image + target - image * target)
        opt.zero_grad()
        iou.backward()
        opt.step()
        if float(iou.detach()) < 0.01:
            return
    raise AssertionError("IoU loss %.4f after 350 iterations" % float(iou.detach()))


@pytest.mark.parametrize("S,aa,fine,general", [(64, False, False, False), (36, True, False, False), (40, True, True, True),
                                                (50, False, True, True), (64, True, False, True), (50, False, "dense", True),
                                                (36, True, "dense", True)])
def test_kernels_stay_inside_their_buffers(nr, S, aa, fine, general):
    """No sanitizer on this pool: every output of the forward / backward is carved out of a larger
    poisoned allocation and the guard words on both sides must survive (vector and scalar fill paths,
    16x16 and 8x8 tiles, one-kernel and general binning, zero-fill chunks, gradient atomics)."""
    import ctypes
    from neural_renderer_v2_pytorch_b200 import _lib, rasterize as rz
    L = _lib.lib()
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    B, G_ = 3, 4096
    g = torch.Generator().manual_seed(S)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), 2.732), torch.rand(B, generator=g) * 80 - 20, torch.rand(B, generator=g) * 360)
    v = nr.perspective(nr.look_at(vw, eye)).cuda().contiguous()
    faces = torch.from_numpy(d["faces"]).cuda()
    vt_np, ft_np, tex_np = nr.create_textures(faces.shape[0], 2)
    tex = torch.rand((B,) + tex_np.shape, generator=g).cuda()
    vt = torch.from_numpy(vt_np)[None].repeat(B, 1, 1).cuda()
    ft = torch.from_numpy(ft_np).cuda()
    R = 2 * S if aa else S
    dense, fine = fine == "dense", fine is True
    flags = (_lib.NR_DRAW_RGB | _lib.NR_DRAW_SILHOUETTES | _lib.NR_DRAW_BACKSIDE | (_lib.NR_ANTI_ALIASING if aa else 0) |
             (_lib.NR_FINE_TILES if fine else 0) | (_lib.NR_GENERAL_BINNING if general else 0) |
             (_lib.NR_DENSE_RASTER if dense else 0))
    cfg = _lib.RasterConfig(batch=B, num_vertices=v.shape[1], num_faces=faces.shape[0], image_size=S, flags=flags,
                            near_plane=0.1, far_plane=100., eps=1e-5, depth_min_delta=1e-4, num_tex_vertices=vt.shape[1],
                            tex_height=tex.shape[2], tex_width=tex.shape[3])
    tile = 8 if fine else 16
    ntx = (R + tile - 1) // tile
    cap = 8 * B * faces.shape[0]
    POISON_I, POISON_F = 0x5a5a5a5a, 12345.5

    def guarded(n, dtype):
        big = torch.full((n + 2 * G_,), POISON_I if dtype == torch.int32 else POISON_F, dtype=dtype, device="cuda")
        return big, big[G_:G_ + n]

    def intact(big, n):
        want = POISON_I if big.dtype == torch.int32 else POISON_F
        return bool((big[:G_] == want).all()) and bool((big[G_ + n:] == want).all())

    sizes = dict(fim=(B * R * R, torch.int32), images=(B * 4 * S * S, torch.float32), internal=(B * 4 * R * R, torch.float32),
                 aux=(B * R * R * 6, torch.float32),
                 tile_list=(8 + 16 * B * ntx * ntx, torch.int32), gv=(v.numel(), torch.float32), gtex=(tex.numel(), torch.float32),
                 gvt=(vt.numel(), torch.float32), ws=(int(L.nr_workspace_bytes(ctypes.byref(cfg), cap)) // 4 + 64, torch.int32))
    buf = {k: guarded(n, dt) for k, (n, dt) in sizes.items()}
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    zf = _lib.ZeroFill(count=3)
    for i, k in enumerate(("gv", "gtex", "gvt")):
        zf.ptr[i] = buf[k][1].data_ptr()
        zf.bytes[i] = buf[k][1].numel() * 4
    ws = buf["ws"][1]
    base = (ws.data_ptr() + 255) & ~255
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.nr_rasterize_forward(ctypes.byref(cfg), ptr(v), ptr(faces), ptr(vt), ptr(ft), ptr(tex), ptr(buf["fim"][1]), None, None,
                                ptr(buf["images"][1]), ptr(buf["internal"][1]) if aa else None, ptr(buf["aux"][1]),
                                ptr(buf["tile_list"][1]),
                                ctypes.c_void_p(base), ws.numel() * 4 - (base - ws.data_ptr()), cap, None, None,
                                ctypes.byref(zf), None, stream)
    _lib.check(rc, "forward")
    G = torch.randn((B, 4, S, S), generator=g).cuda()
    saved = buf["internal"][1] if aa else buf["images"][1]
    rc = L.nr_rasterize_backward(ctypes.byref(cfg), ptr(v), ptr(faces), ptr(vt), ptr(ft), ptr(tex), ptr(buf["fim"][1]), ptr(saved),
                                 ptr(buf["aux"][1]), None if fine else ptr(buf["tile_list"][1]), ptr(G), ptr(buf["gv"][1]),
                                 ptr(buf["gtex"][1]), ptr(buf["gvt"][1]), None, None, stream)
    _lib.check(rc, "backward")
    torch.cuda.synchronize()
    # the same backward WITHOUT the forward's aux map (weights and texel coordinates recomputed): same gradients
    again = {k: torch.zeros_like(buf[k][1]) for k in ("gv", "gtex", "gvt")}
    rc = L.nr_rasterize_backward(ctypes.byref(cfg), ptr(v), ptr(faces), ptr(vt), ptr(ft), ptr(tex), ptr(buf["fim"][1]), ptr(saved),
                                 None, None if fine else ptr(buf["tile_list"][1]), ptr(G), ptr(again["gv"]),
                                 ptr(again["gtex"]), ptr(again["gvt"]), None, None, stream)
    _lib.check(rc, "backward without aux")
    torch.cuda.synchronize()
    for k, t in again.items():
        scale = float(t.abs().max())
        assert float((t - buf[k][1]).abs().max()) <= 2e-6 * scale, k
    for k, (n, _) in sizes.items():
        assert intact(buf[k][0], n), "%s: guard words overwritten" % k
    # and everything inside was written: no poison left in the outputs, gradients finite and non-trivial
    assert not (buf["fim"][1] == POISON_I).any() and not (buf["images"][1] == POISON_F).any()
    if aa:
        assert not (buf["internal"][1] == POISON_F).any()
    for k in ("gv", "gtex", "gvt"):
        t = buf[k][1]
        assert torch.isfinite(t).all() and not (t == POISON_F).any()
    assert float(buf["gv"][1].abs().sum()) > 0 and float(buf["gtex"][1].abs().sum()) > 0
    # same numbers as the public API
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=aa)
    img = nr.rasterize_rgba(v, faces, nr.RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=tex), hp)
    assert torch.equal(img.reshape(-1), buf["images"][1])
