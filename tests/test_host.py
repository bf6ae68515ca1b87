"""Host-side logic that needs no GPU: API surface, parameter defaults, camera math, input checks."""
import math
import os

import numpy as np
import pytest
import torch

import neural_renderer_v2_pytorch_b200 as nr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_public_names_of_the_reference_path():
    # neural_renderer_torch/__init__.py:1-12, the names on the rasterize path
    for name in ("load_obj", "look", "look_at", "perspective", "rasterize_silhouettes", "rasterize_rgba",
                 "rasterize_rgb", "rasterize_depth", "Renderer", "to_gpu", "create_textures",
                 "get_points_from_angles", "differentiation", "RasterizeParam", "RasterizeHyperparam"):
        assert hasattr(nr, name), name


def test_defaults_match_the_reference():
    hp = nr.RasterizeHyperparam()          # rasterize_param.py:13-33
    assert (hp.image_size, hp.near, hp.far, hp.eps) == (256, 0.1, 100.0, 1e-5)
    assert hp.anti_aliasing and hp.draw_backside and hp.draw_rgb and hp.draw_silhouettes and hp.draw_depth
    p = nr.RasterizeParam()
    assert all(getattr(p, k) is None for k in ("vertices_textures", "faces_textures", "textures",
                                               "background_color", "backgrounds", "lights"))
    r = nr.Renderer()                      # renderer.py:8-22
    assert (r.image_size, r.anti_aliasing, r.draw_backside, r.perspective, r.viewing_angle) == (256, True, True, True, 30)
    assert r.camera_mode == "look_at" and r.near == 0.1 and r.far == 100
    assert r.viewpoints == [0, 0, -(1. / math.tan(math.radians(30)) + 1)]


def test_perspective_known_answer():
    # tests_torch/test_perspective.py style: x / z / tan(angle)
    v = torch.tensor([[[1., 2., 4.], [-3., 0.5, 2.]]])
    out = nr.perspective(v, angle=30.)
    w = math.tan(30 / 180. * 3.1416)
    want = torch.tensor([[[1 / 4. / w, 2 / 4. / w, 4.], [-3 / 2. / w, 0.5 / 2. / w, 2.]]])
    assert torch.allclose(out, want, rtol=1e-6)
    assert nr.perspective(v, angle=torch.tensor(30.)).shape == v.shape


def test_look_at_known_answers():
    # tests_torch/test_look_at.py: eye on -z looking at the origin leaves x, y and shifts z
    v = torch.tensor([[[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]]])
    out = nr.look_at(v, [0, 0, -2.])
    assert torch.allclose(out, torch.tensor([[[1., 0., 2.], [0., 1., 2.], [0., 0., 3.]]]), atol=1e-6)
    out = nr.look_at(v, [2., 0, 0])       # eye on +x: world -x becomes +z
    assert torch.allclose(out[0, 0], torch.tensor([0., 0., 1.]), atol=1e-6)
    eyes = torch.tensor([[0., 0., -2.], [2., 0., 0.]])
    out = nr.look_at(v.repeat(2, 1, 1), eyes)
    assert out.shape == (2, 3, 3)


def test_get_points_from_angles():
    x, y, z = nr.get_points_from_angles(2.0, 0., 0.)
    assert abs(x) < 1e-12 and abs(y) < 1e-12 and abs(z + 2.0) < 1e-12
    t = nr.get_points_from_angles(torch.tensor([2.0, 1.0]), torch.tensor([0., 90.]), torch.tensor([90., 0.]))
    assert t.shape == (2, 3)
    assert torch.allclose(t, torch.tensor([[2., 0., 0.], [0., 1., 0.]]), atol=1e-6)


def test_create_textures_layout():
    vt, ft, tex = nr.create_textures(2464, 4)
    assert vt.shape == (7392, 2) and ft.shape == (2464, 3) and tex.shape == (3, 200, 200)
    assert vt.dtype == np.float32 and ft.dtype == np.int32
    assert vt.max() <= 199 and (ft == np.arange(7392).reshape(-1, 3)).all()
    # face 51 -> tile (row 1, col 1) with tile_width 50
    assert (vt[51 * 3] == [4, 4]).all() and (vt[51 * 3 + 1] == [4, 7]).all() and (vt[51 * 3 + 2] == [7, 7]).all()
    vt, ft, tex = nr.create_textures(5, 2, flatten=True)
    assert tex.shape == (3, 10, 2)


def test_load_obj_matches_reference_loader(tmp_path):
    d = np.load(os.path.join(GOLDEN, "teapot.npz"))
    p = tmp_path / "m.obj"
    with open(p, "w") as f:
        f.write("# quad + triangle\nv 0 0 0\nv 2 0 0\nv 2 1 0\nv 0 1 0\nv 1 3 1\n\nf 1 2 3 4\nf 4/1 3/2 5/3\n")
    v, faces = nr.load_obj(str(p), normalization=False)
    assert v.shape == (5, 3) and faces.tolist() == [[0, 1, 2], [0, 2, 3], [3, 2, 4]]
    vn, _ = nr.load_obj(str(p))
    assert np.abs(vn).max() <= 1.0 + 1e-6
    assert d["vertices"].shape == (1292, 3) and d["faces"].shape == (2464, 3)


def test_textured_obj_loader_matches_reference_loader():
    """load_obj(load_textures=True) vs arrays produced by the REFERENCE's loader on the same files
    (tests/golden/make_golden.py::make_textured_obj_fixture): two image materials + one colour material."""
    d = np.load(os.path.join(GOLDEN, "textured_obj_reference_loader.npz"))
    v, f, vt, ft, tex = nr.load_obj(os.path.join(GOLDEN, "textured", "m.obj"), load_textures=True)
    for name, a in (("vertices", v), ("faces", f), ("vertices_t", vt), ("faces_t", ft), ("textures", tex)):
        assert a.dtype == d[name].dtype and np.array_equal(a, d[name]), name
    with pytest.raises(Exception, match="Failed to load textures"):
        nr.load_obj(os.path.join(GOLDEN, "..", "..", "tests", "golden", "textured", "no_mtl.obj")
                    if False else _obj_without_mtl(), load_textures=True)


def _obj_without_mtl():
    import tempfile
    f = tempfile.NamedTemporaryFile("w", suffix=".obj", delete=False)
    f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    f.close()
    return f.name


def test_save_obj_round_trip(tmp_path):
    """tests_torch/test_save_obj.py in spirit: save -> load returns the same mesh and (8-bit) texture."""
    v, f, vt, ft, tex = nr.load_obj(os.path.join(GOLDEN, "textured", "m.obj"), load_textures=True, normalization=False)
    vt_before = vt.copy()
    p = str(tmp_path / "out.obj")
    nr.save_obj(p, v, f, vt, ft, tex)
    assert np.array_equal(vt, vt_before), "save_obj must not modify its arguments"
    v2, f2, vt2, ft2, tex2 = nr.load_obj(p, load_textures=True, normalization=False)
    assert np.allclose(v2, v, atol=1e-6) and np.array_equal(f2, f) and np.array_equal(ft2, ft)
    assert np.allclose(vt2, vt, atol=1e-4) and tex2.shape == tex.shape
    assert np.abs(tex2 - tex).max() <= 0.5 / 255 + 1e-6
    nr.save_obj(str(tmp_path / "plain.obj"), v, f)
    v3, f3 = nr.load_obj(str(tmp_path / "plain.obj"), normalization=False)
    assert np.allclose(v3, v, atol=1e-6) and np.array_equal(f3, f)
    img = nr.imread(str(tmp_path / "out.png"))
    assert img.dtype == np.float32 and img.max() <= 1.0


def test_cpu_tensors_are_rejected_not_computed():
    """No CPU fallback: a CPU tensor raises like CHECK_CUDA (rasterize_cuda.cpp:5)."""
    hp = nr.RasterizeHyperparam(image_size=16, anti_aliasing=False)
    v = torch.zeros(1, 3, 3)
    f = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        nr.rasterize_silhouettes(v, f, nr.RasterizeParam(), hp)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        nr.face_index_map_forward_safe(torch.zeros(1, 1, 3, 3), torch.zeros(256, dtype=torch.int32), 1, 16, 0.1, 100.,
                                       1, 1e-8, 1e-4)
    with pytest.raises(AssertionError):
        nr.rasterize_silhouettes(torch.zeros(3, 3), f, nr.RasterizeParam(), hp)
    with pytest.raises(AssertionError):
        nr.rasterize_silhouettes(v, torch.zeros(1, 4, dtype=torch.int32), nr.RasterizeParam(), hp)


def test_draw_flags_are_set_like_the_reference():
    hp = nr.RasterizeHyperparam(image_size=16, anti_aliasing=False)
    with pytest.raises(RuntimeError):
        nr.rasterize_depth(torch.zeros(1, 3, 3), torch.tensor([[0, 1, 2]]), nr.RasterizeParam(), hp)
    assert (hp.draw_rgb, hp.draw_silhouettes, hp.draw_depth) == (False, False, True)   # rasterize.py:360-362
    assert hp.image_size == 16


def test_backgrounds_are_shape_checked_like_the_reference():
    """rasterize.py:216-225: [B, 3, R, R] with R doubled under anti-aliasing (AssertionError)."""
    from neural_renderer_v2_pytorch_b200.rasterize import _prepare
    v = torch.zeros(1, 3, 3)
    f = torch.tensor([[0, 1, 2]])
    hp = nr.RasterizeHyperparam(image_size=16, anti_aliasing=True, draw_depth=False)
    for bad in (torch.zeros(1, 3, 16, 16), torch.zeros(2, 3, 32, 32), torch.zeros(1, 4, 32, 32), torch.zeros(3, 32, 32)):
        with pytest.raises(AssertionError):
            _prepare(v, f, nr.RasterizeParam(backgrounds=bad), hp)
    # silhouettes never look at the background (rasterize.py:286-288 sits inside the rgb branch)
    hp.draw_rgb = False
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _prepare(v, f, nr.RasterizeParam(backgrounds=torch.zeros(1, 3, 16, 16)), hp)


def test_vertex_normals_match_the_dense_incidence_formulation():
    """lights.vertex_normals (index_add) vs the reference's [nf,nv] incidence matmul (rasterize.py:167-182),
    including a face that lists the same vertex twice (counted once)."""
    from neural_renderer_v2_pytorch_b200.lights import vertex_normals
    g = torch.Generator().manual_seed(0)
    v = torch.randn(2, 7, 3, generator=g)
    f = torch.tensor([[0, 1, 2], [2, 3, 4], [4, 5, 6], [0, 2, 4], [1, 1, 3]], dtype=torch.int32)
    fv = v[:, f.long()]
    n = torch.linalg.cross(fv[:, :, 1] - fv[:, :, 0], fv[:, :, 2] - fv[:, :, 1], dim=-1)
    m = torch.zeros(5, 7)
    for k in range(3):
        m[torch.arange(5), f[:, k].long()] = 1
    want = torch.nn.functional.normalize(torch.matmul(n.permute(0, 2, 1), m).permute(0, 2, 1), dim=2)
    assert torch.allclose(vertex_normals(v, f), want, atol=1e-6)


def test_light_classes_and_packing():
    B = 3
    lights = [nr.DirectionalLight(torch.rand(B, 3), torch.rand(B, 3), backside=True), nr.AmbientLight(torch.rand(B, 3)),
              nr.SpecularLight(torch.rand(B, 3))]
    assert torch.equal(lights[2].alpha, torch.ones(B))
    from neural_renderer_v2_pytorch_b200.lights import pack_lights
    types, data = pack_lights(lights, B, "cpu")
    assert types.tolist() == [1 | 4, 0, 2] and data.shape == (3, B, 8)
    assert torch.equal(data[0, :, 3:6], lights[0].direction) and torch.equal(data[2, :, 6], torch.ones(B))


def test_bench_bytes_formula():
    import bench
    fwd, bwd = bench.algorithmic_bytes(nv=1292, nf=2464, T=40000, P=512 * 512, C=4, S=512)
    assert abs((fwd + bwd) / 1e6 - 18.28) < 0.01
    fwd, bwd = bench.algorithmic_bytes(nv=1292, nf=2464, T=0, P=512 * 512, C=1, S=512)
    assert abs((fwd + bwd) / 1e6 - 10.55) < 0.03


def test_mesh_and_adam_exist_and_work(tmp_path):
    """Names of the reference package outside the accelerated path (mesh.py, optimizers.py)."""
    obj = tmp_path / "tri.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 3 4\n")
    m = nr.Mesh(str(obj), texture_size=2)
    v, f, t = m.get_batch(3)
    assert v.shape == (3, 4, 3) and f.shape == (3, 2, 3) and t.shape == (3, 2, 2, 2, 2, 3)
    assert float(t.min()) > 0 and float(t.max()) < 1
    # Adam: chainer's bias-corrected update with the per-parameter factor param.lr
    p = torch.nn.Parameter(torch.tensor([1.0, -2.0]))
    q = torch.nn.Parameter(torch.tensor([3.0]))
    q.lr = 0.0
    opt = nr.Adam([p, q], alpha=0.1)
    (p.pow(2).sum() + q.sum()).backward()
    opt.step()
    # first step of Adam moves every coordinate by alpha against the sign of its gradient
    assert torch.allclose(p.detach(), torch.tensor([0.9, -1.9]), atol=1e-6)
    assert float(q) == 3.0


def test_raster_batch_cursor_protocol():
    """Model of the lock-free work cursor of k_raster (nr_raster.cu, "Scheduling"): a CTA's 8 warps draw
    items from one shared word  cur = (batch / 8) << 5 | taken;  the warp that draws taken == 8 fetches
    the next batch of 8 from the global counter and publishes it with one exchange, later ones wait for
    the batch field to change.  Under arbitrary interleavings (each shared / global access is one atomic
    step; a warp may be delayed anywhere in between) every item must be processed exactly once.
    (A 4-bit `taken` field fails this test: sixteen draws in one epoch carry into the batch field.)"""
    import random

    def run(seed, total_items, nctas, nwarps=8, taken_bits=5):
        rnd = random.Random(seed)
        mask = (1 << taken_bits) - 1
        counter, done, warps = 0, [], []
        for _ in range(nctas):
            cta = {"cur": (counter >> 3) << taken_bits}
            counter += 8
            warps += [{"pc": "draw", "cta": cta, "alive": True} for _ in range(nwarps)]
        for _ in range(10 ** 6):
            alive = [w for w in warps if w["alive"]]
            if not alive:
                break
            w = rnd.choice(alive)
            cta = w["cta"]
            if w["pc"] == "draw":
                w["v"] = cta["cur"]
                cta["cur"] += 1
                w["pc"] = "decode"
            elif w["pc"] == "decode":
                taken, batch = w["v"] & mask, w["v"] >> taken_bits
                if taken < 8:
                    item = batch * 8 + taken
                    if item >= total_items:
                        w["alive"] = False
                    else:
                        done.append(item)
                        w["pc"], w["left"] = "work", rnd.randint(0, 5)
                else:
                    w["pc"] = "fetch" if taken == 8 else "wait"
            elif w["pc"] == "work":
                w["left"] -= 1
                if w["left"] <= 0:
                    w["pc"] = "draw"
            elif w["pc"] == "fetch":
                w["base"] = counter
                counter += 8
                w["pc"] = "publish"
            elif w["pc"] == "publish":
                cta["cur"] = (w["base"] >> 3) << taken_bits
                w["pc"] = "draw"
            elif w["pc"] == "wait":
                if (cta["cur"] >> taken_bits) != (w["v"] >> taken_bits):
                    w["pc"] = "draw"
        else:
            return False
        return sorted(done) == list(range(total_items))

    assert all(run(seed, total_items=150 + seed, nctas=1 + seed % 4) for seed in range(120))
    assert not all(run(seed, total_items=150, nctas=2, taken_bits=4) for seed in range(40))
