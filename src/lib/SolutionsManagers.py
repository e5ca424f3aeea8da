"""Drop-in for the reference's src/lib/SolutionsManagers.py: re-exports the B200-backed implementation."""
from romhighcontrast_b200.lib.SolutionsManagers import *  # noqa: F401,F403
from romhighcontrast_b200.lib import SolutionsManagers as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
