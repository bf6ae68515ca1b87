#!/usr/bin/env python
"""Times the REFERENCE's own two live CUDA kernels (oracle/_ref, compiled from /root/reference by
oracle/build_ref.py) on this GPU at the BASELINE config-2 geometry, next to this repo's fused forward
on the same inputs: the "before / after" of stages a3 + a4 of SURVEY.md section 8 (z-buffer + weight
map; the reference's remaining ~60 torch ops and per-view host synchronisations are not in this
number).  Checker infrastructure, like tests/: prints one JSON line.

    python tools/time_reference_kernels.py [views]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import bench  # noqa: E402
import make_ref_kernel_golden as mk  # noqa: E402
import neural_renderer_v2_pytorch_b200 as nr  # noqa: E402


def main():
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mod = mk.load_reference_extension()
    if mod is None:
        print(json.dumps({"unavailable": "oracle/_ref not built"}))
        return
    w = dict(bench.WORKLOADS["cfg2"], views=views)
    dev = torch.device("cuda:0")
    inp = bench.make_inputs(w, 1000, dev, nr)
    S = w["S"]
    v = inp["vertices"].to(dev)
    faces = inp["faces"].to(dev)
    fv = v[:, faces.long()].contiguous()                      # rasterize.py:232
    B, nf = fv.shape[:2]

    def ref_step():
        fim = torch.full((B * S * S,), -1, dtype=torch.int32, device=dev)      # rasterize.py:32-33 (on the device here)
        mod.face_index_map_forward_safe(fv, fim, nf, S, 0.1, 100.0, 1, 1e-8, 1e-4)
        wm = torch.zeros((B * S * S, 3), dtype=torch.float32, device=dev)      # rasterize.py:71
        mod.compute_weight_map_c(fv, fim, wm, nf, S)
        return fim, wm

    def ours_step():
        hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=False, draw_rgb=False, draw_depth=False)
        return nr.rasterize_maps(v, faces, nr.RasterizeParam(), hp)

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    t_ref, (fim_r, wm_r) = timed(ref_step, 3)
    t_our, maps = timed(ours_step, 20)
    same = bool(torch.equal(fim_r.reshape(B, S, S), maps["face_index_map"])) and bool(
        torch.equal(wm_r.reshape(B, S, S, 3), maps["weight_map"]))
    print(json.dumps({"workload": "cfg2 geometry: teapot, %d views, %dx%d" % (B, S, S),
                      "reference_kernels_ms": round(t_ref, 3), "this_repo_forward_with_maps_ms": round(t_our, 3),
                      "ratio": round(t_ref / t_our, 1), "face_index_map_and_weight_map_identical": same}))


if __name__ == "__main__":
    main()
