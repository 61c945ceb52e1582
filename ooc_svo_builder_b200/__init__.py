"""B200-native voxelize-and-build path of Forceflow/ooc_svo_builder.

Partitioner -> Schwarz-Seidel voxelizer -> sparse voxel octree builder as
hand-written sm_100a CUDA kernels behind a C ABI (include/svo_b200.h), with the
reference's .tri/.tridata input and .octree/.octreenodes/.octreedata output
byte layout.  See DESIGN.md.
"""
from .api import (SvoBuilder, SvoError, Octree, Params, Stats, PinnedBuffer,  # noqa: F401
                  estimate_partitions, header_bytes, load_library, LIB_PATH, ABI_SYMBOLS)

__version__ = "0.1.0"
