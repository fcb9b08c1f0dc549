"""firedrake_ts_b200 -- B200-native assembly engine behind the firedrake-ts API.

Public interface = the reference's (firedrake_ts/__init__.py:1): ``DAEProblem`` and
``DAESolver``; plus the small object model needed to state a problem without
Firedrake (meshes, spaces, Functions, DirichletBC, the form catalogue).
"""
from . import fem, forms, mesh  # noqa: F401
from .function import Constant, DirichletBC, Function  # noqa: F401
from .mesh import (FunctionSpace, Partition, PeriodicIntervalMesh, SimplexMesh, UnitCubeMesh,  # noqa: F401
                   UnitIntervalMesh, UnitSquareMesh)
from .solving_utils import ConvergenceError  # noqa: F401
from .ts_solver import DAEProblem, DAESolver  # noqa: F401

__all__ = ["DAEProblem", "DAESolver"]
