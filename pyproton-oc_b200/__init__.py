"""protonoc-b200: the B200-native crime / recruitment hot path of PyPROTON-OC.

Host side in Python (as the reference is), hand-written sm_100a CUDA kernels behind the C ABI declared
in include/protonoc_b200.h.  There is no CPU fallback."""
from . import tables  # noqa: F401
from ._cabi import PocError, build  # noqa: F401
from .params import HOT_PATH_DEFAULTS, make_demography, make_labour, make_params  # noqa: F401
from .engine import CrimeEngine, LAYER_NAMES, REPORTER_NAMES  # noqa: F401
from .model import ProtonOC, Person, free_parameters  # noqa: F401
from . import ensemble, synth  # noqa: F401

__all__ = ["ProtonOC", "Person", "free_parameters", "ensemble", "synth", "CrimeEngine", "make_params", "make_demography", "make_labour", "HOT_PATH_DEFAULTS", "PocError", "build", "tables", "LAYER_NAMES",
           "REPORTER_NAMES"]
