"""gendr_b200 -- B200-native (sm_100a) differentiable soft rasterizer behind the GenDR API.

Same top-level names as the reference package (`/root/reference/gendr/__init__.py`): the renderer module, the mesh
container, camera transforms and lighting.  The rasterization forward/backward run in hand-written CUDA
(gendr_b200/csrc, C ABI in include/gendr_b200.h); everything else here is thin PyTorch host code.
"""
from . import functional
from .lighting import AmbientLighting, DirectionalLighting, Lighting
from .losses import FlattenLoss, LaplacianLoss
from .mesh import Mesh
from .renderer import GenDR
from .transform import Look, LookAt, Projection

__all__ = ['functional', 'AmbientLighting', 'DirectionalLighting', 'Lighting', 'Mesh', 'GenDR', 'Look', 'LookAt',
           'Projection', 'LaplacianLoss', 'FlattenLoss']
__version__ = '0.1.0'
