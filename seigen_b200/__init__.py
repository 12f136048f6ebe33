"""seigen_b200: B200-native drop-in for the explicit ElasticLF4 path of devitocodes/seigen.

``from seigen_b200 import *`` provides what the reference scripts obtain from ``from firedrake import *`` and
``from seigen import *`` *for this path*: the utility meshes, DG function spaces, ``Function`` / ``Expression``,
``ElasticLF4`` and the scalar helpers.  All time stepping runs in hand-written sm_100a CUDA kernels behind the C
ABI of ``include/seigen_b200.h``; there is no CPU fallback.
"""
from math import pi, sqrt  # noqa: F401  (the reference scripts get these from `from firedrake import *`)

from .compat import (File, Function, FunctionSpace, TensorFunctionSpace, VectorFunctionSpace,  # noqa: F401
                     errornorm_l2, get_timers, norm, timed_region)
from .elastic import ElasticLF4, ExplicitElasticLF4, step_times  # noqa: F401
from .expression import Expression  # noqa: F401
from .forms import TestFunction, TrialFunction, dx, inner, lhs, rhs, solve  # noqa: F401
from .helpers import Vp, Vs, cfl_dt, get_dofs, log  # noqa: F401
from .mesh import (BoxMesh, IntervalMesh, Mesh, RectangleMesh, UnitCubeMesh, UnitIntervalMesh,  # noqa: F401
                   UnitSquareMesh)

__version__ = "0.1.0"
