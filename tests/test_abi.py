"""The C-ABI library loads and exports exactly what include/nr_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nr_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from neural_renderer_v2_pytorch_b200 import _lib
    return _lib.build()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NR_API\s+[\w\s\*]+?\b(nr_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("nr_rasterize_forward", "nr_rasterize_backward", "nr_differentiation_backward",
              "nr_face_index_map_forward_safe", "nr_compute_weight_map", "nr_workspace_bytes"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = set(re.findall(r" T (nr_\w+)", out))
    assert set(declared_symbols()) == exported


def test_python_binding_lists_every_symbol(lib_path):
    from neural_renderer_v2_pytorch_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    L = _lib.lib()
    for s in _lib.SYMBOLS:
        assert hasattr(L, s)
    assert L.nr_abi_version() == _lib.ABI_VERSION == 4


def test_no_torch_in_the_abi(lib_path):
    """plain pointers and sizes: the library does not link against torch / ATen / python."""
    out = subprocess.check_output(["ldd", lib_path], text=True)
    assert "torch" not in out and "c10" not in out and "python" not in out
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)      # declarations without comments
    assert "at::" not in code and "torch" not in code and "Tensor" not in code


def test_pure_host_entry_points(lib_path):
    from neural_renderer_v2_pytorch_b200 import _lib
    L = _lib.lib()
    f = _lib
    assert L.nr_num_channels(f.NR_DRAW_RGB | f.NR_DRAW_SILHOUETTES) == 4
    assert L.nr_num_channels(f.NR_DRAW_SILHOUETTES) == 1
    assert L.nr_num_channels(f.NR_DRAW_RGB | f.NR_DRAW_SILHOUETTES | f.NR_DRAW_DEPTH) == 5
    cfg = f.RasterConfig(batch=64, num_vertices=1292, num_faces=2464, image_size=512,
                         flags=f.NR_DRAW_RGB | f.NR_DRAW_SILHOUETTES, near_plane=0.1, far_plane=100., eps=1e-5,
                         depth_min_delta=1e-4, num_tex_vertices=7392, tex_height=200, tex_width=200)
    small = L.nr_workspace_bytes(ctypes.byref(cfg), 1000)
    big = L.nr_workspace_bytes(ctypes.byref(cfg), 1000000)
    assert big - small >= 4 * (1000000 - 1000) - 512 and small % 256 == 0
    # argument validation happens before any CUDA call
    rc = L.nr_rasterize_forward(None, *([None] * 13), 0, 0, None, None, None, None, None)
    assert rc == f.NR_ERR_INVALID_ARGUMENT and b"NULL" in L.nr_last_error()
    cfg.flags = 0
    rc = L.nr_rasterize_forward(ctypes.byref(cfg), *([None] * 13), 0, 0, None, None, None, None, None)
    assert rc == f.NR_ERR_INVALID_ARGUMENT and b"nothing to draw" in L.nr_last_error()
    assert L.nr_differentiation_backward(None, None, None, 1, 8, 3, None) == f.NR_ERR_INVALID_ARGUMENT


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from neural_renderer_v2_pytorch_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neural_renderer_v2_pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text and "nro_" not in text, fn
