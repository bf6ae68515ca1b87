#!/usr/bin/env python
"""Runs a few eager fwd+bwd steps of one bench workload (for `ncu` launch lists / captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import neural_renderer_v2_pytorch_b200 as nr  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
inp = bench.make_inputs(w, 1000, dev, nr)
S = w["S"]
rgb = w["mode"] in ("rgb", "rgba")
faces, G = inp["faces"].to(dev), inp["G"].to(dev)
vt = inp["vt"].to(dev) if rgb else None
ft = inp["ft"].to(dev) if rgb else None
v = inp["vertices"].to(dev).requires_grad_(True)
tex = inp["textures"].to(dev).requires_grad_(True) if rgb else None
fn = {"rgba": nr.rasterize_rgba, "rgb": nr.rasterize_rgb, "silhouettes": nr.rasterize_silhouettes}[w["mode"]]
for _ in range(steps):
    v.grad = None
    if tex is not None:
        tex.grad = None
    hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=w["aa"])
    p = nr.RasterizeParam(vertices_textures=vt, faces_textures=ft, textures=tex) if rgb else nr.RasterizeParam()
    img = fn(v, faces, p, hp)
    img.backward(G)
torch.cuda.synchronize()
print("done", float(v.grad.abs().sum()))
