"""rxmesh_b200 -- B200-native (sm_100a) implementation of RXMesh's static query hot path.

Host mirror of the reference interface over the C ABI in include/rxmesh_b200.h; the CUDA
library must be built (rxmesh_b200/librxmesh_b200.so) -- there is no CPU fallback.
"""
from . import meshio  # noqa: F401
from ._lib import RXMeshError, lib, LIB_PATH  # noqa: F401
from .mesh import (AoS, AoSoA, Attribute, DEVICE, HOST, INVALID64, LOCATION_ALL, Op, RXMeshStatic,  # noqa: F401
                   SoA, launch_count, load_patcher_file, rx_init, set_async)
