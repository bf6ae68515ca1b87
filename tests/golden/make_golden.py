"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

How the reference is run here (no GPU):
  * ``/root/reference`` is put on ``sys.path`` and ``neural_renderer_torch`` is imported
    unmodified; ``chainer`` and ``imageio`` (absent from this image, only needed by
    ``optimizers.py`` / file I/O) are stubbed in ``sys.modules``.
  * the reference's compiled module ``neural_renderer_torch.cuda.rasterize_cuda`` needs a
    GPU; it is replaced by a shim with the same three functions that runs the C
    restatement ``oracle/nr_oracle.c`` on CPU tensors.  So in these fixtures
    ``face_index_map`` / ``weight_map`` come from the C oracle, while EVERYTHING ELSE
    (``to_map``, ``compute_coordinate_map``, ``sample_textures``, ``compute_depth_map``,
    ``mask_foreground``, ``Differentiation``, flip / anti-aliasing, ``look_at``,
    ``perspective``, ``Renderer``, torch autograd through all of it) is the reference's
    own code executing.  The two kernels are pinned separately against the reference's real
    CUDA kernels on a B200 (``ref_kernel_*.npz``, written by
    ``tests/golden/make_ref_kernel_golden.py`` under gpurun).

Fixtures written (all small, np.savez_compressed):
  teapot.npz                      vertices/faces from the reference's load_obj(teapot.obj)
  diff_known_answer.npz           Differentiation backward on random data
  case_*.npz                      inputs, images, maps and input gradients of rasterize_* / Renderer
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def import_reference():
    """Import the unmodified reference package with stubbed optional deps."""
    chainer = types.ModuleType("chainer")
    chainer.optimizers = types.ModuleType("chainer.optimizers")
    chainer.optimizers.adam = types.ModuleType("chainer.optimizers.adam")
    chainer.optimizers.adam.AdamRule = type("AdamRule", (), {})
    chainer.optimizers.adam.Adam = type("Adam", (), {})
    chainer.optimizers.Adam = chainer.optimizers.adam.Adam
    chainer.optimizer = types.ModuleType("chainer.optimizer")
    chainer.cuda = types.ModuleType("chainer.cuda")
    sys.modules.setdefault("chainer", chainer)
    sys.modules.setdefault("chainer.optimizers", chainer.optimizers)
    sys.modules.setdefault("chainer.optimizers.adam", chainer.optimizers.adam)
    imageio = types.ModuleType("imageio")
    sys.modules.setdefault("imageio", imageio)

    # shim for the GPU-only extension, backed by the C oracle
    shim = types.ModuleType("neural_renderer_torch.cuda.rasterize_cuda")

    def face_index_map_forward_safe(faces, face_index, num_faces, image_size, near, far,
                                    draw_backside, eps, depth_min_delta):
        B = faces.shape[0]
        fim = oracle.face_index_map(faces.detach().numpy().reshape(B, num_faces, 3, 3), image_size,
                                    near, far, draw_backside, depth_min_delta)
        face_index.copy_(torch.from_numpy(fim.reshape(-1)))
        return face_index

    def compute_weight_map_c(faces, face_index_map, weight_map, num_faces, image_size):
        B = faces.shape[0]
        wm = oracle.weight_map(faces.detach().numpy().reshape(B, num_faces, 3, 3),
                               face_index_map.numpy().reshape(B, image_size, image_size))
        weight_map.copy_(torch.from_numpy(wm.reshape(-1, 3)))
        return face_index_map

    shim.face_index_map_forward_safe = face_index_map_forward_safe
    shim.face_index_map_forward_unsafe = None
    shim.compute_weight_map_c = compute_weight_map_c
    cuda_pkg = types.ModuleType("neural_renderer_torch.cuda")
    cuda_pkg.__path__ = []
    cuda_pkg.rasterize_cuda = shim
    sys.modules["neural_renderer_torch.cuda"] = cuda_pkg
    sys.modules["neural_renderer_torch.cuda.rasterize_cuda"] = shim

    sys.path.insert(0, REF)
    import neural_renderer_torch as nr
    return nr


def cameras(nr, B, seed, distance=2.732):
    g = torch.Generator().manual_seed(seed)
    elev = torch.rand(B, generator=g) * 80. - 20.
    azim = torch.rand(B, generator=g) * 360.
    dist = torch.full((B,), distance)
    return nr.get_points_from_angles(dist, elev, azim)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %-34s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.))


def screen_space(nr, vertices_world, viewpoints, perspective=True):
    r = nr.Renderer()
    r.viewpoints = viewpoints
    r.perspective = perspective
    return r.transform_vertices(vertices_world)


def main():
    nr = import_reference()
    from neural_renderer_torch.rasterize_param import RasterizeParam, RasterizeHyperparam
    torch.manual_seed(0)
    np.random.seed(0)

    # ---------------------------------------------------------------- teapot mesh
    v_np, f_np = nr.load_obj(os.path.join(REF, "examples_pytorch/data/teapot.obj"))
    save("teapot", vertices=v_np.astype(np.float32), faces=f_np.astype(np.int32))
    nf = f_np.shape[0]
    faces = torch.as_tensor(f_np)

    # ---------------------------------------------------------------- Differentiation known answer
    # (shapes of tests_torch/test_differentiation.py:12-26, smaller batch)
    for C in (1, 3, 4):
        g = torch.Generator().manual_seed(10 + C)
        images = torch.randn(3, 32, 32, C, generator=g)
        coords = torch.zeros(3, 32, 32, 2, requires_grad=True)
        noise = torch.randn(3, 32, 32, C, generator=g)
        (nr.differentiation(images, coords) * noise).sum().backward()
        save("diff_known_answer_c%d" % C, images=images, grad_output=noise,
             grad_coordinates=coords.grad)
    # a silhouette-like binary image exercises the exact-zero / tie branches of maximum()
    g = torch.Generator().manual_seed(20)
    images = (torch.rand(2, 24, 24, 1, generator=g) > 0.5).float()
    coords = torch.zeros(2, 24, 24, 2, requires_grad=True)
    noise = torch.randn(2, 24, 24, 1, generator=g)
    (nr.differentiation(images, coords) * noise).sum().backward()
    save("diff_known_answer_binary", images=images, grad_output=noise, grad_coordinates=coords.grad)

    # ---------------------------------------------------------------- rasterize_* cases
    def run_case(name, B, S, aa, mode, draw_backside=True, ts=4, seed=0, perspective=True,
                 distance=2.732, near=0.1, far=100.0):
        vw = torch.as_tensor(v_np)[None].repeat(B, 1, 1)
        vp = cameras(nr, B, seed, distance)
        vs = screen_space(nr, vw, vp, perspective).detach().clone().requires_grad_(True)
        hp = RasterizeHyperparam(image_size=S, near=near, far=far, anti_aliasing=aa,
                                 draw_backside=draw_backside)
        kw = {}
        if mode in ("rgb", "rgba"):
            vt_np, ft_np, tex_np = nr.create_textures(nf, texture_size=ts)
            gt = torch.Generator().manual_seed(seed + 100)
            tex = torch.rand((B,) + tex_np.shape, generator=gt).requires_grad_(True)
            vt = torch.as_tensor(vt_np)[None].repeat(B, 1, 1).requires_grad_(True)
            ft = torch.as_tensor(ft_np)
            params = RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=tex)
            kw = dict(vertices_textures=vt, faces_textures=ft, textures=tex)
        else:
            params = RasterizeParam()
        fn = {"silhouettes": nr.rasterize_silhouettes, "rgb": nr.rasterize_rgb,
              "rgba": nr.rasterize_rgba, "depth": nr.rasterize_depth}[mode]
        images = fn(vs, faces, params, hp)
        gg = torch.Generator().manual_seed(1)
        G = torch.randn(images.shape, generator=gg)
        (images * G).sum().backward()
        R = S * 2 if aa else S
        fv = vs.detach()[:, faces.long()].numpy()
        fim = oracle.face_index_map(fv, R, near, far, draw_backside)
        wm = oracle.weight_map(fv, fim)
        out = dict(vertices=vs.detach(), faces=f_np.astype(np.int32), image_size=S,
                   anti_aliasing=int(aa), draw_backside=int(draw_backside), near=near, far=far,
                   mode=mode, images=images, grad_images=G, grad_vertices=vs.grad,
                   face_index_map=fim, weight_map=wm.astype(np.float32))
        if kw:
            out.update(vertices_textures=kw["vertices_textures"].detach(),
                       faces_textures=kw["faces_textures"].numpy().astype(np.int32),
                       textures=kw["textures"].detach(),
                       grad_textures=kw["textures"].grad,
                       grad_vertices_textures=kw["vertices_textures"].grad)
        save("case_" + name, **out)

    run_case("sil_64", B=2, S=64, aa=False, mode="silhouettes")
    run_case("sil_aa_32", B=2, S=32, aa=True, mode="silhouettes", seed=3)
    run_case("sil_cull_64", B=2, S=64, aa=False, mode="silhouettes", draw_backside=False, seed=4)
    run_case("rgba_64", B=2, S=64, aa=False, mode="rgba", seed=5)
    run_case("rgba_aa_32", B=2, S=32, aa=True, mode="rgba", seed=6)
    run_case("rgb_ortho_aa_32", B=2, S=32, aa=True, mode="rgb", seed=7, perspective=False)
    run_case("rgb_cull_48", B=1, S=48, aa=False, mode="rgb", draw_backside=False, seed=8, ts=2)
    run_case("depth_64", B=2, S=64, aa=False, mode="depth", seed=9)
    run_case("depth_aa_32", B=1, S=32, aa=True, mode="depth", seed=11)
    # near/far clipping active: far plane cuts the back of the teapot
    run_case("sil_clip_64", B=2, S=64, aa=False, mode="silhouettes", seed=12, near=2.0, far=2.9)

    # ---------------------------------------------------------------- Renderer end to end
    # (world-space vertices -> look_at -> perspective -> rasterize), gradient to world vertices
    B, S = 2, 32
    vw = torch.as_tensor(v_np)[None].repeat(B, 1, 1).requires_grad_(True)
    r = nr.Renderer()
    r.image_size = S
    r.viewpoints = cameras(nr, B, 21)
    vt_np, ft_np, tex_np = nr.create_textures(nf, texture_size=2)
    tex = torch.rand((B,) + tex_np.shape, generator=torch.Generator().manual_seed(22)).requires_grad_(True)
    vt = torch.as_tensor(vt_np)[None].repeat(B, 1, 1)
    images = r.render(vw, faces, vt, torch.as_tensor(ft_np), tex)
    G = torch.randn(images.shape, generator=torch.Generator().manual_seed(1))
    (images * G).sum().backward()
    save("renderer_rgba_aa_32", vertices_world=vw.detach(), faces=f_np.astype(np.int32),
         viewpoints=r.viewpoints, image_size=S, vertices_textures=vt,
         faces_textures=ft_np.astype(np.int32), textures=tex.detach(), images=images,
         grad_images=G, grad_vertices_world=vw.grad, grad_textures=tex.grad)

    # two-triangle square of tests_torch/test_rasterize.py:205-212 (one step, gradient sign pin)
    sq = torch.tensor([[[0.1, 0.1, 1.], [-0.1, 0.1, 1.], [-0.1, -0.1, 1.], [0.1, -0.1, 1.]]],
                      requires_grad=True)
    sqf = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32)
    hp = RasterizeHyperparam(image_size=64, anti_aliasing=False)
    img = nr.rasterize_silhouettes(sq, sqf, RasterizeParam(), hp)
    G = torch.randn(img.shape, generator=torch.Generator().manual_seed(1))
    (img * G).sum().backward()
    save("case_square_64", vertices=sq.detach(), faces=sqf.numpy(), image_size=64, anti_aliasing=0,
         draw_backside=1, near=0.1, far=100.0, mode="silhouettes", images=img, grad_images=G,
         grad_vertices=sq.grad)


def make_lights_fixture(nr):
    """rasterize_rgb with Directional + Ambient + Specular lights (the setup of
    tests_torch/test_rasterize.py:114-203, per-view colours / directions from a seed), forward and
    gradients to vertices and textures: the reference's compute_normal_map (rasterize.py:162-190) and
    light accumulation (rasterize.py:252-283) executing on CPU."""
    from neural_renderer_torch.rasterize_param import RasterizeParam, RasterizeHyperparam
    d = np.load(os.path.join(HERE, "teapot.npz"))
    v_np, f_np = d["vertices"], d["faces"]
    nf = f_np.shape[0]
    for name, B, S, aa, backside, lb in (("lit_rgb_48", 2, 48, False, True, False), ("lit_rgb_aa_24", 2, 24, True, False, True)):
        g = torch.Generator().manual_seed(40 + S)
        vw = torch.as_tensor(v_np)[None].repeat(B, 1, 1)
        vs = screen_space(nr, vw, cameras(nr, B, 50 + S)).detach().clone().requires_grad_(True)
        vt_np, ft_np, tex_np = nr.create_textures(nf, texture_size=2)
        tex = torch.rand((B,) + tex_np.shape, generator=g).requires_grad_(True)
        vt = torch.as_tensor(vt_np)[None].repeat(B, 1, 1)
        dir_color, dir_dir = torch.rand((B, 3), generator=g), F_normalize(torch.randn((B, 3), generator=g))
        amb_color, spec_color = torch.rand((B, 3), generator=g) * 0.4, torch.rand((B, 3), generator=g) * 0.5
        spec_alpha = torch.rand(B, generator=g) * 3 + 1
        lights = [nr.DirectionalLight(dir_color.clone(), dir_dir.clone(), backside=lb), nr.AmbientLight(amb_color.clone()),
                  nr.SpecularLight(spec_color.clone(), spec_alpha.clone(), backside=lb)]
        hp = RasterizeHyperparam(image_size=S, anti_aliasing=aa, draw_backside=backside)
        params = RasterizeParam(vertices_textures=vt, faces_textures=torch.as_tensor(ft_np), textures=tex, lights=lights)
        images = nr.rasterize_rgb(vs, torch.as_tensor(f_np), params, hp)
        G = torch.randn(images.shape, generator=torch.Generator().manual_seed(1))
        (images * G).sum().backward()
        save(name, vertices=vs.detach(), faces=f_np.astype(np.int32), image_size=S, anti_aliasing=int(aa),
             draw_backside=int(backside), near=0.1, far=100.0, mode="rgb", images=images, grad_images=G,
             grad_vertices=vs.grad, vertices_textures=vt, faces_textures=ft_np.astype(np.int32), textures=tex.detach(),
             grad_textures=tex.grad, grad_vertices_textures=np.zeros_like(vt.numpy()),
             dir_color=dir_color, dir_direction=dir_dir, amb_color=amb_color, spec_color=spec_color,
             spec_alpha=spec_alpha, light_backside=int(lb))


def F_normalize(t):
    return t / t.norm(dim=-1, keepdim=True)


def make_textured_obj_fixture(nr):
    """A small synthetic textured mesh (two image materials of different width + one colour
    material), loaded with the REFERENCE's load_obj(load_textures=True); tests/test_host.py checks this
    repo's loader against the stored arrays.  The .obj / .mtl / .png files are committed too."""
    from PIL import Image
    d = os.path.join(HERE, "textured")
    os.makedirs(d, exist_ok=True)
    rng = np.random.RandomState(5)
    Image.fromarray(rng.randint(0, 256, size=(6, 8, 3)).astype("uint8")).save(os.path.join(d, "a.png"))
    Image.fromarray(rng.randint(0, 256, size=(4, 5, 3)).astype("uint8")).save(os.path.join(d, "b.png"))
    with open(os.path.join(d, "m.mtl"), "w") as f:
        f.write("newmtl first\nKd 1 1 1\nmap_Kd a.png\n\nnewmtl second\nmap_Kd b.png\n\nnewmtl flat\nKd 0.25 0.5 0.75\n")
    with open(os.path.join(d, "m.obj"), "w") as f:
        f.write("mtllib m.mtl\n")
        for v in [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0.5, 0.5, 1), (2, 2, 2)]:
            f.write("v %g %g %g\n" % v)
        for t in [(0, 0), (1, 0), (1, 1), (0, 1), (0.5, 0.25), (0.2, 0.9)]:
            f.write("vt %g %g\n" % t)
        f.write("usemtl first\nf 1/1 2/2 3/3 4/4\nusemtl second\nf 1/5 2/6 5/1\nf 2/2 3/3 5/4\nusemtl flat\nf 3 4 5\nf 4/1 1/2 6/3\n")
    import imageio
    imageio.imread = lambda fn: np.asarray(Image.open(fn).convert("RGB"))
    v, f, vt, ft, tex = nr.load_obj(os.path.join(d, "m.obj"), load_textures=True)
    save("textured_obj_reference_loader", vertices=v, faces=f, vertices_t=vt, faces_t=ft, textures=tex)


def make_reference_golden_png_fixture(nr):
    """The reference's own golden image: tests_torch/data/4e49873292196f02574b5684eaec43e9.png is what
    tests_torch/test_save_obj.py:43 and tests_chainer/test_rasterize.py:43-72 compare a render of the
    textured model next to it with (atol 1e-2).  The model is loaded HERE with the reference's
    load_obj(load_textures=True); the arrays and the golden pixels travel as one fixture."""
    from PIL import Image
    import imageio
    imageio.imread = lambda fn: np.asarray(Image.open(fn).convert("RGB"))
    base = "/root/reference/tests_torch/data/4e49873292196f02574b5684eaec43e9"
    v, f, vt, ft, tex = nr.load_obj(base + "/model.obj", load_textures=True)
    png = np.asarray(Image.open(base + ".png"))
    save("reference_golden_png_4e4987", vertices=v, faces=f, vertices_t=vt, faces_t=ft,
         textures=(np.asarray(tex) * 255 + 0.5).astype(np.uint8),      # the atlas is 8-bit image data
         golden_png=png, viewpoint=nr.get_points_from_angles(2.5, 10, -90))


def make_gradient_png_fixture():
    """tests_torch/data/gradient.png, the target silhouette of tests_torch/test_rasterize.py:205-249."""
    from PIL import Image
    save("reference_gradient_png", gradient_png=np.asarray(Image.open("/root/reference/tests_torch/data/gradient.png")))


if __name__ == "__main__":
    main()
    make_gradient_png_fixture()
    make_reference_golden_png_fixture(import_reference())
    make_textured_obj_fixture(import_reference())
    make_lights_fixture(import_reference())
