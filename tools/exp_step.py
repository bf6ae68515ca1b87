#!/usr/bin/env python
"""Developer timing loop: ms per captured fwd+bwd step of one bench workload, for each value of the
NR_EXP environment variable given on the command line (experiment switches compiled into the library
while a change is being evaluated; none are active in a committed build), or for an alternative build
of the library: `altX` loads csrc/libnr_b200_altX.so.  Boxes differ by ~2 %, so A and B go in one call.

    python tools/exp_step.py cfg2 0 1 2 alt
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload):
    import torch
    import bench
    import neural_renderer_v2_pytorch_b200 as nr
    if os.environ.get("NR_LIB_ALT"):        # A/B against an alternative build: csrc/libnr_b200_<name>.so
        from neural_renderer_v2_pytorch_b200 import _lib
        _lib.LIB_PATH = os.path.join(_lib.CSRC, "libnr_b200_%s.so" % os.environ["NR_LIB_ALT"])
    w = bench.WORKLOADS[workload]
    dev = torch.device("cuda:0")
    inp = bench.make_inputs(w, 1000, dev, nr)
    B, S = w["views"], w["S"]
    rgb = w["mode"] in ("rgb", "rgba")
    faces, G = inp["faces"].to(dev), inp["G"].to(dev)
    vt = inp["vt"].to(dev) if rgb else None
    ft = inp["ft"].to(dev) if rgb else None
    v = inp["vertices"].to(dev).requires_grad_(True)
    tex = inp["textures"].to(dev).requires_grad_(True) if rgb else None
    fn = {"rgba": nr.rasterize_rgba, "rgb": nr.rasterize_rgb, "silhouettes": nr.rasterize_silhouettes}[w["mode"]]

    def step():
        hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"])
        p = nr.RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=tex) if rgb else nr.RasterizeParam()
        img = fn(v, faces, p, hp)
        img.backward(G)
        return img

    params = [v] + ([tex] if rgb else [])
    for _ in range(3):
        for p in params:
            p.grad = None
        step()
    run = nr.capture_step(step, params=params, warmup=2)
    for _ in range(20):
        run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200)
    print("NR_EXP=%s %s: %.4f ms/step  (grad checksum %.6g)" % (os.environ.get("NR_LIB_ALT", "") + os.environ.get("NR_EXP", "0"), workload, best,
                                                            float(v.grad.double().abs().sum())))


if __name__ == "__main__":
    if os.environ.get("NR_EXP_CHILD"):
        child(sys.argv[1])
    else:
        wl = sys.argv[1]
        for x in sys.argv[2:] or ["0"]:
            env = dict(os.environ, NR_EXP=x, NR_EXP_CHILD="1")
            if x.startswith("alt"):
                env.update(NR_EXP="0", NR_LIB_ALT=x)
            subprocess.run([sys.executable, os.path.abspath(__file__), wl], env=env)
