"""Import shim: the package directory required by the project layout is `gkr-msm_b200/` (not a valid
python identifier), so this importable alias just points its module search path there."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gkr-msm_b200"))

from .binding import *  # noqa: E402,F401,F403
