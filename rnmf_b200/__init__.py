"""rnmf_b200 -- B200-native (sm_100a) stencil sweep / ghost fill / residual reduction behind the
mesh, boundary-condition and solver API of wattrg/rnmf.  The compute lives in librnmf_b200.so
(CUDA, C ABI: include/rnmf_b200.h); this package is the thin host-side mirror of the reference's
interface used by the tests and bench.py.  There is no CPU fallback."""
from . import _capi
from ._capi import RnmfError
from .boundary_conditions import BCs, BcType, ComponentBCs
from .dataframe import CartesianDataFrame, CartesianDataFrame2D, CartesianDataFrame3D
from .mesh import CartesianMesh, CartesianMesh2D, CartesianMesh3D
from .runtime import Context, default_context, slab_partition
from .config import ConfigPanic, read_lua

__all__ = ["BCs", "BcType", "ComponentBCs", "CartesianDataFrame", "CartesianDataFrame2D", "CartesianDataFrame3D",
           "CartesianMesh", "CartesianMesh2D", "CartesianMesh3D", "Context", "default_context", "slab_partition",
           "RnmfError", "ConfigPanic", "read_lua"]
