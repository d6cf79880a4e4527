"""B200-native PSULVSB registration hot path: ctypes host side over libpsulvsb_b200.so."""
from . import capi  # noqa: F401
from .capi import Handle, HostProblem, PsulvsbError, default_params  # noqa: F401
from .solver import RobustRegistrationSolver, RegistrationSolution  # noqa: F401
