"""axom_b200 -- B200-native (sm_100a) spin::BVH + quest::SignedDistance behind a C ABI.

Only the hot path lives here: csrc/ (CUDA kernels + C ABI), lib/ (the built libaxb200.so),
and thin host mirrors of the reference classes (bvh.py, signed_distance.py).  There is no CPU
fallback: importing the classes works anywhere, calling them needs the CUDA library and a GPU.
"""
from .bvh import BVH, BVH_BUILD_OK, DEFAULT_SCALE_FACTOR  # noqa: F401
from .signed_distance import SignedDistance  # noqa: F401
from .distributed_closest_point import DistributedClosestPoint  # noqa: F401
from .mesh_tester import MeshTester, findTriMeshIntersectionsBVH, intersect_triangles  # noqa: F401
from .marching_cubes import MarchingCubes, MarchingCubesDataParallelism  # noqa: F401

__version__ = "0.1.0"
