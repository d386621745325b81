"""Drop-in for the reference's `open_kitchen_pybind` module (Pybind/bindings.cpp): same `Environment` /
`RenderTargetInfo` names and methods, served by the B200 kernels, plus `BatchEnv` (zero-copy DLPack batch API)."""
from openkitchen_b200.lib.open_kitchen_pybind import Environment, RenderTargetInfo  # noqa: F401


def __getattr__(name):
    if name == "BatchEnv":
        from openkitchen_b200 import BatchEnv

        return BatchEnv
    raise AttributeError(name)
