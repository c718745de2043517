"""Same registry shape as /root/reference/models/__init__.py: ``models.__dict__[name](data, no_levels=N)``."""
from .FAL_netA import *  # noqa: F401,F403
from .FAL_netB import *  # noqa: F401,F403
from .FAL_netC import *  # noqa: F401,F403

__all__ = ("FAL_netA", "FAL_netB", "FAL_netC")
