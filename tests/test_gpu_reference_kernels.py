"""The reference's REAL CUDA kernels (oracle/_ref/nr_ref_rasterize_cuda.so, compiled from the sources
under /root/reference by oracle/build_ref.py) against (a) the C oracle and (b) this repo's CUDA path,
on the same inputs on the same GPU.  face_index_map and weight_map must be bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_ref_kernel_golden as mk  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refmod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mod = mk.load_reference_extension()
    if mod is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py in the build container)")
    return mod


@pytest.fixture(scope="module")
def nr():
    import neural_renderer_v2_pytorch_b200 as nr_
    return nr_


def ours(nr, faces, S, near, far, bs):
    f = torch.from_numpy(faces).cuda().contiguous()
    B, nf = f.shape[:2]
    fim = torch.empty(B * S * S, dtype=torch.int32, device="cuda")
    nr.face_index_map_forward_safe(f, fim, nf, S, near, far, bs, 1e-8, 1e-4)
    wm = torch.zeros((B * S * S, 3), dtype=torch.float32, device="cuda")
    nr.compute_weight_map_c(f, fim, wm, nf, S)
    return fim.reshape(B, S, S).cpu().numpy(), wm.reshape(B, S, S, 3).cpu().numpy()


def ours_dense(nr, faces, S, near, far, bs):
    """The same maps from the fused forward with the face-parallel raster kernel (nr_raster_dense.cu)."""
    from neural_renderer_v2_pytorch_b200 import rasterize as rz
    B, nf = faces.shape[:2]
    v = torch.from_numpy(np.ascontiguousarray(faces, dtype=np.float32)).reshape(B, nf * 3, 3).cuda()
    idx = torch.arange(nf * 3, dtype=torch.int32).reshape(nf, 3).cuda()
    rz.FORCE_DENSE_RASTER = True
    try:
        hp = nr.RasterizeHyperparam(image_size=S, near=near, far=far, anti_aliasing=False, draw_backside=bool(bs),
                                    draw_rgb=False, draw_depth=False)
        maps = nr.rasterize_maps(v, idx, nr.RasterizeParam(), hp)
    finally:
        rz.FORCE_DENSE_RASTER = None
    return maps["face_index_map"].cpu().numpy(), maps["weight_map"].cpu().numpy()


@pytest.mark.parametrize("name", sorted(mk.ref_kernel_inputs().keys()))
def test_three_way(refmod, nr, name):
    faces, S, near, far, bs = mk.ref_kernel_inputs()[name]
    fim_ref, wm_ref = mk.run_reference(refmod, faces, S, near, far, bs)
    fim_c = oracle.face_index_map(faces, S, near, far, bool(bs))
    wm_c = oracle.weight_map(faces, fim_c)
    fim_o, wm_o = ours(nr, faces, S, near, far, bs)
    assert np.array_equal(fim_c, fim_ref), "C oracle != reference kernel: %d px" % (fim_c != fim_ref).sum()
    assert np.array_equal(wm_c, wm_ref), "C oracle weight map != reference kernel"
    assert np.array_equal(fim_o, fim_ref), "CUDA path != reference kernel: %d px" % (fim_o != fim_ref).sum()
    assert np.array_equal(wm_o, wm_ref), "CUDA path weight map != reference kernel"
    fim_d, wm_d = ours_dense(nr, faces, S, near, far, bs)
    assert np.array_equal(fim_d, fim_ref), "dense raster kernel != reference kernel: %d px" % (fim_d != fim_ref).sum()
    assert np.array_equal(wm_d, wm_ref), "dense raster kernel weight map != reference kernel"


def test_full_size_teapot_512(refmod, nr):
    """BASELINE config 2 geometry (teapot, 512^2), 8 views: 2.1 M pixels, bit-exact vs the reference kernel."""
    faces = mk.teapot_faces(8, 77)
    fim_ref, wm_ref = mk.run_reference(refmod, faces, 512, 0.1, 100.0, 1)
    fim_o, wm_o = ours(nr, faces, 512, 0.1, 100.0, 1)
    assert (fim_ref >= 0).mean() > 0.05
    assert np.array_equal(fim_o, fim_ref), "%d px differ" % (fim_o != fim_ref).sum()
    assert np.array_equal(wm_o, wm_ref)
    fim_d, wm_d = ours_dense(nr, faces, 512, 0.1, 100.0, 1)
    assert np.array_equal(fim_d, fim_ref), "dense raster kernel: %d px differ" % (fim_d != fim_ref).sum()
    assert np.array_equal(wm_d, wm_ref)


def test_dense_random_1024(refmod, nr):
    """Config-4 style stress at reduced face count: 200k random ~5 px triangles at 1024^2."""
    faces = mk.random_faces(1, 200000, 78, size=0.005)
    fim_ref, wm_ref = mk.run_reference(refmod, faces, 1024, 0.1, 100.0, 1)
    fim_o, wm_o = ours(nr, faces, 1024, 0.1, 100.0, 1)
    assert np.array_equal(fim_o, fim_ref), "%d px differ" % (fim_o != fim_ref).sum()
    assert np.array_equal(wm_o, wm_ref)
    fim_d, wm_d = ours_dense(nr, faces, 1024, 0.1, 100.0, 1)
    assert np.array_equal(fim_d, fim_ref), "dense raster kernel: %d px differ" % (fim_d != fim_ref).sum()
    assert np.array_equal(wm_d, wm_ref)


def _stacked_faces(n, S, seed, jitter, box=0.08, views=1):
    """n triangles that all cover the same few dozen pixels at depths within `jitter` of each other: every pixel
    under them is a near-tie, i.e. the order-dependent z-test (rasterize_cuda_kernel.cu:145-148) decides."""
    rng = np.random.RandomState(seed)
    c = rng.uniform(-0.5, 0.5, size=(views, 1, 1, 2)).astype(np.float32)
    xy = (c + rng.uniform(-box, box, size=(views, n, 3, 2))).astype(np.float32)
    z = (1.5 + rng.uniform(-jitter, jitter, size=(views, n, 3, 1))).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([xy, z], -1))


@pytest.mark.parametrize("n,jitter", [(300, 2e-4), (5000, 3e-4), (5000, 5e-2)])
def test_dense_path_with_piles_of_near_ties(refmod, nr, n, jitter):
    """The z-buffer rasterizer on its worst case: thousands of faces over the same pixels with depths inside
    (and around) the 1e-4 hysteresis.  Every covered pixel is contested, a CTA's (face, contested pixel) pairs
    outgrow their shared-memory queue, pixels hold more candidates than the resolve pass sorts (it then replays
    the reference's loop over all faces), and with the default pool the candidate nodes overflow on the first
    call.  The map must still be the reference kernel's, bit for bit."""
    S = 128
    faces = _stacked_faces(n, S, 1234 + n, jitter)
    fim_ref, wm_ref = mk.run_reference(refmod, faces, S, 0.1, 100.0, 1)
    assert (fim_ref >= 0).sum() > 50
    fim_c = oracle.face_index_map(faces, S)
    assert np.array_equal(fim_c, fim_ref), "C oracle != reference kernel"
    fim_o, wm_o = ours(nr, faces, S, 0.1, 100.0, 1)
    assert np.array_equal(fim_o, fim_ref), "tile pipeline: %d px differ" % (fim_o != fim_ref).sum()
    for attempt in range(3):        # the host grows the candidate pool from the statistics of earlier calls
        fim_d, wm_d = ours_dense(nr, faces, S, 0.1, 100.0, 1)
        assert np.array_equal(fim_d, fim_ref), "z-buffer path, call %d: %d px differ" % (attempt, (fim_d != fim_ref).sum())
        assert np.array_equal(wm_d, wm_ref)
        torch.cuda.synchronize()
