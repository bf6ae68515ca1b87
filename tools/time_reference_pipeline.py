#!/usr/bin/env python
"""The reference's "before" number on this GPU: its own ``rasterize_core``
(neural_renderer_torch/rasterize.py:194-329) with its own CUDA kernels, forward + backward, on the inputs
``bench.make_inputs`` produces, next to this repo's path on the same inputs.  Baseline only, never the target.

The reference package is Python + one CUDA extension:
  * the extension is ``oracle/_ref/nr_ref_rasterize_cuda.so`` (the reference's two sources compiled by
    ``oracle/build_ref.py``), registered as ``neural_renderer_torch.cuda.rasterize_cuda``;
  * the Python package is imported UNMODIFIED from ``/root/reference`` when that exists (build container: no
    GPU), else from ``oracle/_ref/pkg`` - a throw-away staging copy that ``--stage`` makes right before a gpurun
    call and ``--unstage`` removes right after it (git-ignored; reference sources never enter the history);
  * ``chainer`` and ``imageio`` (absent from this image; only ``optimizers.py`` / file I/O need them) are stubbed.

    python tools/time_reference_pipeline.py --stage        # build container, before gpurun
    python tools/time_reference_pipeline.py cfg1 cfg2:4    # on the GPU box: workload[:views]
    python tools/time_reference_pipeline.py --unstage      # build container, after gpurun
"""
import json
import os
import shutil
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
STAGE = os.path.join(ROOT, "oracle", "_ref", "pkg")
REF = "/root/reference"


def import_reference():
    import torch  # noqa: F401
    import make_ref_kernel_golden as mk
    ext = mk.load_reference_extension()
    if ext is None:
        return None, "oracle/_ref not built"
    src = REF if os.path.isdir(os.path.join(REF, "neural_renderer_torch")) else STAGE
    if not os.path.isdir(os.path.join(src, "neural_renderer_torch")):
        return None, "reference package not staged (run with --stage in the build container first)"
    chainer = types.ModuleType("chainer")
    chainer.optimizers = types.ModuleType("chainer.optimizers")
    chainer.optimizers.adam = types.ModuleType("chainer.optimizers.adam")
    chainer.optimizers.adam.AdamRule = type("AdamRule", (), {})
    chainer.optimizers.adam.Adam = type("Adam", (), {})
    chainer.optimizers.Adam = chainer.optimizers.adam.Adam
    chainer.optimizer = types.ModuleType("chainer.optimizer")
    chainer.cuda = types.ModuleType("chainer.cuda")
    for name, mod in (("chainer", chainer), ("chainer.optimizers", chainer.optimizers),
                      ("chainer.optimizers.adam", chainer.optimizers.adam), ("imageio", types.ModuleType("imageio"))):
        sys.modules.setdefault(name, mod)
    cuda_pkg = types.ModuleType("neural_renderer_torch.cuda")
    cuda_pkg.__path__ = []
    cuda_pkg.rasterize_cuda = ext
    sys.modules["neural_renderer_torch.cuda"] = cuda_pkg
    sys.modules["neural_renderer_torch.cuda.rasterize_cuda"] = ext
    sys.path.insert(0, src)
    import neural_renderer_torch as ref
    return ref, src


def main():
    args = sys.argv[1:]
    if args == ["--stage"]:
        shutil.rmtree(STAGE, ignore_errors=True)
        shutil.copytree(os.path.join(REF, "neural_renderer_torch"), os.path.join(STAGE, "neural_renderer_torch"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.so", "build", "*.cu", "*.cpp"))
        print("staged", STAGE)
        return
    if args == ["--unstage"]:
        shutil.rmtree(STAGE, ignore_errors=True)
        print("removed", STAGE)
        return
    import torch
    import bench
    import neural_renderer_v2_pytorch_b200 as nr
    ref, where = import_reference()
    if ref is None:
        print(json.dumps({"unavailable": where}))
        return
    from neural_renderer_torch.rasterize_param import RasterizeParam as RP, RasterizeHyperparam as RH
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)                # the reference launches on the current device
    for spec in (args or ["cfg1", "cfg2:4"]):
        name, _, nviews = spec.partition(":")
        w = dict(bench.WORKLOADS[name])
        if nviews:
            w["views"] = int(nviews)
        inp = bench.make_inputs(w, 1000, dev, nr)
        S, B = w["S"], w["views"]
        rgb = w["mode"] in ("rgb", "rgba")
        faces = inp["faces"].to(dev)
        G = inp["G"].to(dev)
        vt = inp["vt"].to(dev) if rgb else None
        ft = inp["ft"].to(dev) if rgb else None

        def leaves():
            v = inp["vertices"].to(dev).clone().requires_grad_(True)
            t = inp["textures"].to(dev).clone().requires_grad_(True) if rgb else None
            return v, t

        def ref_step(v, t):
            fn = {"rgba": ref.rasterize_rgba, "rgb": ref.rasterize_rgb, "silhouettes": ref.rasterize_silhouettes}[w["mode"]]
            p = RP(vertices_textures=vt, faces_textures=ft, textures=t) if rgb else RP()
            img = fn(v, faces, p, RH(image_size=S, anti_aliasing=w["aa"]))
            img.backward(G)
            return img

        def our_step(v, t):
            fn = {"rgba": nr.rasterize_rgba, "rgb": nr.rasterize_rgb, "silhouettes": nr.rasterize_silhouettes}[w["mode"]]
            p = nr.RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=t) if rgb else nr.RasterizeParam()
            img = fn(v, faces, p, nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"]))
            img.backward(G)
            return img

        def timed(step, reps):
            v, t = leaves()
            step(v, t)                              # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                v, t = leaves()
                img = step(v, t)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3, img.detach(), v.grad.detach()

        ms_ref, img_r, gv_r = timed(ref_step, 3)
        ms_our, img_o, gv_o = timed(our_step, 20)
        scale = float(gv_r.abs().max())
        print(json.dumps({
            "workload": "%s, %d view(s), %dx%d, anti_aliasing=%s, %s" % (name, B, S, S, w["aa"], w["mode"]),
            "reference_ms_per_fwd_bwd": round(ms_ref, 2), "this_repo_ms_per_fwd_bwd_eager": round(ms_our, 4),
            "speedup": round(ms_ref / ms_our, 1),
            "reference_mpix_views_per_s": round(B * S * S / 1e6 / (ms_ref / 1e3), 2),
            "images_max_abs_diff": float((img_r - img_o).abs().max()),
            "grad_vertices_max_abs_diff_over_scale": float((gv_r - gv_o).abs().max()) / max(scale, 1e-30),
            "how": "wall clock around eager calls with a synchronize on both sides (the reference synchronises with the "
                   "host several times per view); reference = its own Python (%s) + its own CUDA kernels (oracle/_ref)" % where}))


if __name__ == "__main__":
    main()
