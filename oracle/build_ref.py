"""Build the REFERENCE's own CUDA extension into oracle/_ref/ (test infrastructure).

The reference's two live kernels only exist as CUDA (no CPU path), so the strongest oracle for
``face_index_map`` / ``weight_map`` is the reference code itself running on the B200.  This
script compiles

    /root/reference/neural_renderer_torch/cuda/rasterize_cuda.cpp
    /root/reference/neural_renderer_torch/cuda/rasterize_cuda_kernel.cu

for sm_100a into ``oracle/_ref/nr_ref_rasterize_cuda.so`` (git-ignored, travels to the GPU box).
torch 2.11 rejects ``AT_DISPATCH_FLOATING_TYPES(x.type(), ...)`` at rasterize_cuda_kernel.cu:322,
346,372,399,427, so ``.type()`` is rewritten to ``.scalar_type()`` on a temporary copy under /tmp
(no arithmetic is touched; reference sources are never copied into the repo).

Run in the build container:  python oracle/build_ref.py
"""
import os
import re
import shutil
import sys
import tempfile

REF_CUDA = "/root/reference/neural_renderer_torch/cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "nr_ref_rasterize_cuda"


def build():
    if not os.path.isdir(REF_CUDA):
        print("reference not present (%s); keeping whatever is in %s" % (REF_CUDA, OUT))
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, NAME + ".so")
    srcs = [os.path.join(REF_CUDA, f) for f in ("rasterize_cuda.cpp", "rasterize_cuda_kernel.cu")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
        return so
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    if os.path.exists("/usr/bin/gcc"):
        os.environ["CC"] = "/usr/bin/gcc"
        os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils import cpp_extension
    tmp = tempfile.mkdtemp(prefix="nr_ref_src_")
    try:
        patched = []
        for s in srcs:
            text = open(s).read()
            if s.endswith(".cu"):
                text, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)", r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()", text)
                assert n == 5, "expected 5 dispatch sites, found %d" % n
            dst = os.path.join(tmp, os.path.basename(s))
            open(dst, "w").write(text)
            patched.append(dst)
        cpp_extension.load(name=NAME, sources=patched, build_directory=OUT,
                           extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"],
                           verbose=False)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for junk in os.listdir(OUT):
        if junk.endswith((".o", ".ninja", ".ninja_deps", ".ninja_log")) or junk.startswith(".ninja"):
            os.remove(os.path.join(OUT, junk))
    return so


if __name__ == "__main__":
    print(build())
    sys.exit(0)
