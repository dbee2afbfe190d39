"""meshing.jl_b200 -- B200-native isosurface extraction behind the Meshing.jl `isosurface` API.

The directory name contains a dot, so the package is loaded by path under the module name
`meshing_jl_b200` (see `__graft_entry__.load_package`).

Public surface (mirrors src/Meshing.jl:9-11 of the reference): `isosurface`, `MarchingCubes`,
`MarchingTetrahedra`.  Everything is computed by the CUDA library `lib/libb200iso.so` through the C ABI
declared in include/b200iso.h; there is no CPU fallback.
"""
from . import synth  # noqa: F401
from .api import MarchingCubes, MarchingTetrahedra, isosurface, Float32, Float64  # noqa: F401
from . import capi  # noqa: F401
from . import sharding  # noqa: F401
from . import api  # noqa: F401
from . import mesh  # noqa: F401
