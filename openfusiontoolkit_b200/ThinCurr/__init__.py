"""ThinCurr model class for the B200 operator-build backend (mirror of
OpenFUSIONToolkit.ThinCurr for the dense operator-build path)."""
from ._core import ThinCurr  # noqa: F401
from .sensor import circular_flux_loop, save_sensors  # noqa: F401
from .meshing import build_ThinCurr_dummy, build_torus_vessel  # noqa: F401
