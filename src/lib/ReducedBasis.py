"""Drop-in for the reference's src/lib/ReducedBasis.py: re-exports the B200-backed implementation."""
from romhighcontrast_b200.lib.ReducedBasis import *  # noqa: F401,F403
from romhighcontrast_b200.lib import ReducedBasis as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
