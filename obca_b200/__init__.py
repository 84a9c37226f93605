"""obca_b200 -- B200-native batched OBCA trajectory optimiser behind the reference's call surface.

Product = libobca.so (hand-written sm_100a CUDA, C-ABI in include/obca.h).  This package is the thin host side:
`parking` mirrors ParkingSignedDist / ParkingDist / DualMultWS / ParkingConstraints; `scenarios` holds the
host-side input producers (obstHrep twin, scenario constants, synthetic warm starts)."""
from . import parking, quadcopter, scenarios  # noqa: F401
from ._lib import ObcaError, default_opts, lib  # noqa: F401
from .parking import DualMultWS, ParkingConstraints, ParkingDist, ParkingSignedDist  # noqa: F401
from .quadcopter import QuadcopterDist, QuadcopterSignedDist, constrSatisfaction  # noqa: F401
