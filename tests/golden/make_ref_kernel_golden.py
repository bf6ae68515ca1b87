"""Run the REFERENCE's real CUDA kernels (oracle/_ref, built by oracle/build_ref.py from the
sources under /root/reference) on a B200 and store their outputs as fixtures.

    gpurun -- 'python tests/golden/make_ref_kernel_golden.py'      # writes gpurun_out/ref_kernel_*.npz
    cp gpurun_out/ref_kernel_*.npz tests/golden/                   # commit them

These pin the C restatement (oracle/nr_oracle.c) on the CPU side: tests/test_oracle.py checks the
oracle bit-for-bit against these files without needing a GPU or the reference.
Inputs are regenerated from seeds by ``ref_kernel_inputs`` (also used by the tests), so only the
outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference_extension():
    import torch  # noqa: F401  (the extension links against libtorch)
    so = os.path.join(ROOT, "oracle", "_ref", "nr_ref_rasterize_cuda.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("nr_ref_rasterize_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def teapot_faces(B, seed, distance=2.732):
    """Teapot seen from B seeded cameras, as gathered screen-space faces [B,nf,3,3] (numpy, CPU math)."""
    import torch
    import neural_renderer_v2_pytorch_b200 as nr
    d = np.load(os.path.join(HERE, "teapot.npz"))
    g = torch.Generator().manual_seed(seed)
    vw = torch.from_numpy(d["vertices"])[None].repeat(B, 1, 1)
    eye = nr.get_points_from_angles(torch.full((B,), distance), torch.rand(B, generator=g) * 80 - 20,
                                    torch.rand(B, generator=g) * 360)
    vs = nr.perspective(nr.look_at(vw, eye))
    return vs[:, torch.from_numpy(d["faces"]).long()].numpy().astype(np.float32)


def random_faces(B, nf, seed, size=0.05):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-1.1, 1.1, size=(B, nf, 1, 2))
    xy = c + rng.normal(0, size, size=(B, nf, 3, 2))
    z = rng.uniform(1.0, 3.0, size=(B, nf, 1, 1)) + rng.normal(0, 0.02, size=(B, nf, 3, 1))
    return np.concatenate([xy, z], -1).astype(np.float32)


def ref_kernel_inputs():
    """name -> (faces [B,nf,3,3], image_size, near, far, draw_backside)"""
    return {
        "teapot_128": (teapot_faces(4, 31), 128, 0.1, 100.0, 1),
        "teapot_cull_96": (teapot_faces(2, 32), 96, 0.1, 100.0, 0),
        "teapot_clip_100": (teapot_faces(2, 33), 100, 2.2, 2.9, 1),
        "random_128": (random_faces(2, 4000, 34), 128, 0.1, 100.0, 1),
        "random_big_64": (random_faces(1, 300, 35, size=0.8), 64, 0.1, 100.0, 0),
    }


def run_reference(mod, faces, S, near, far, backside):
    import torch
    f = torch.from_numpy(faces).cuda().contiguous()
    B, nf = f.shape[:2]
    fim = torch.zeros(B * S * S, dtype=torch.int32, device="cuda") - 1
    mod.face_index_map_forward_safe(f, fim, nf, S, near, far, backside, 1e-8, 1e-4)
    wm = torch.zeros((B * S * S, 3), dtype=torch.float32, device="cuda")
    mod.compute_weight_map_c(f, fim, wm, nf, S)
    torch.cuda.synchronize()
    return fim.reshape(B, S, S).cpu().numpy(), wm.reshape(B, S, S, 3).cpu().numpy()


def main():
    mod = load_reference_extension()
    assert mod is not None, "oracle/_ref/nr_ref_rasterize_cuda.so missing: run oracle/build_ref.py first"
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for name, (faces, S, near, far, bs) in ref_kernel_inputs().items():
        fim, wm = run_reference(mod, faces, S, near, far, bs)
        path = os.path.join(out_dir, "ref_kernel_%s.npz" % name)
        np.savez_compressed(path, face_index_map=fim, weight_map=wm, faces_checksum=np.float64(faces.astype(np.float64).sum()))
        print("wrote", path, os.path.getsize(path), "bytes; foreground", int((fim >= 0).sum()))


if __name__ == "__main__":
    main()
