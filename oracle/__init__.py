"""CPU oracle for the rasterize -> sample -> approximate-gradient path.

TEST INFRASTRUCTURE ONLY. Nothing under ``neural_renderer_v2_pytorch_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / baseline.

Two layers:

* ``libnr_oracle.so`` (``nr_oracle.c``): plain-C restatement of the reference's two live
  CUDA kernels (z-buffer ``face_index_map`` and ``weight_map``), bit-faithful to the SASS
  nvcc emits for them.  Bound here through ctypes.
* ``oracle.pipeline``: torch-CPU float32 restatement of the pure-torch stages
  (``rasterize.py:60-153,194-329``, ``differentiation.py:6-40``, ``utils.py:75-160``).

Parity status: the torch stages are pinned against the reference's own Python code
imported from ``/root/reference`` (``tests/golden/make_golden.py``); the two kernels are
pinned against the reference's real CUDA kernels compiled into ``oracle/_ref`` and run on a
B200 (``tests/golden/ref_kernel_*.npz`` + ``tests/test_gpu_reference_kernels.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnr_oracle.so")
_lib = None


def build(force=False):
    """Compile nr_oracle.c -> libnr_oracle.so with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "nr_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return True


def lib():
    global _lib
    if _lib is None:
        if not _cpu_has_fma():
            raise RuntimeError("oracle/libnr_oracle.so is built with -mfma; this CPU has no FMA")
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        for name in ("nro_face_index_map", "nro_face_index_map_rows"):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [fp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                           ctypes.c_float, ctypes.c_int, ctypes.c_float]
        L.nro_weight_map.restype = None
        L.nro_weight_map.argtypes = [fp, ip, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def face_index_map(faces, image_size, near=0.1, far=100.0, draw_backside=True,
                   depth_min_delta=1e-4, literal=False):
    """faces [B,nf,3,3] float32 (screen space) -> face_index_map [B,S,S] int32.

    Restates ``rasterize.py:27-38`` + ``rasterize_cuda_kernel.cu:52-153``.
    ``literal=True`` runs the every-pixel-scans-every-face loop; the default hoists the
    y bounding-box rejection per image row (same result, tested).
    """
    faces = _f32(faces)
    B, nf = faces.shape[:2]
    fim = np.full((B, image_size, image_size), -1, dtype=np.int32)
    fn = lib().nro_face_index_map if literal else lib().nro_face_index_map_rows
    fn(faces.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
       fim.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
       B, nf, image_size, near, far, int(bool(draw_backside)), depth_min_delta)
    return fim


def weight_map(faces, fim):
    """faces [B,nf,3,3], fim [B,S,S] -> weight_map [B,S,S,3] float32.

    Restates ``rasterize.py:67-77`` + ``rasterize_cuda_kernel.cu:246-308``.
    """
    faces = _f32(faces)
    fim = np.ascontiguousarray(fim, dtype=np.int32)
    B, nf = faces.shape[:2]
    S = fim.shape[1]
    wm = np.zeros((B, S, S, 3), dtype=np.float32)
    lib().nro_weight_map(faces.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                         fim.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                         wm.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, nf, S)
    return wm
