"""B200-native MPM substep engine behind the reference's HybridSolver / ParticleSystem / RegularGrid /
LagrangianMesh / LevelSet interfaces (2iw31Zhv/AnisotropicElastoplasticity, HybridSolver.cpp:867-1032).

The compute path is hand-written sm_100a CUDA in csrc/, exported through the C ABI of include/aep_b200.h
(libaep_b200.so).  There is no CPU fallback: anything that steps a simulation raises if the CUDA library
is missing or no GPU is present.  `scenes` (pure numpy scene generators) imports without the library.
"""
from . import scenes  # noqa: F401

__all__ = ["scenes", "engine", "capi"]


def __getattr__(name):
    if name in ("engine", "capi", "distributed"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
