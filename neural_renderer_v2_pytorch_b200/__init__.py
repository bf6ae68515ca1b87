"""B200-native drop-in for the rasterize -> sample -> approximate-gradient path of
``neural_renderer_torch`` (Rebirth-Alex/neural_renderer_v2_pytorch).

Same public names as the reference package (``neural_renderer_torch/__init__.py:1-12``) for
everything on that path; the kernels are hand-written sm_100a CUDA behind a C ABI
(``include/nr_b200.h``, ``csrc/``).  No CPU or PyTorch fallback exists: importing works anywhere,
calling an operator without the built library or without a CUDA tensor raises.
"""
from .lights import Light, DirectionalLight, AmbientLight, SpecularLight
from .load_obj import load_obj
from .look import look
from .look_at import look_at
from .mesh import Mesh
from .optimizers import Adam
from .perspective import perspective
from .rasterize_param import RasterizeParam, RasterizeHyperparam
from .rasterize import (rasterize_silhouettes, rasterize_rgba, rasterize_rgb, rasterize_depth,
                        rasterize_core, rasterize_maps, face_index_map_forward_safe,
                        compute_weight_map_c)
from .renderer import Renderer
from .save_obj import save_obj
from .utils import to_gpu, imread, make_gif, create_textures, get_points_from_angles
from .differentiation import differentiation
from .graph import capture_step
from . import parallel

__version__ = '2.0.2+b200.1'
