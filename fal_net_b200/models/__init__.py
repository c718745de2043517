"""Same registry shape as /root/reference/models/__init__.py: ``models.__dict__[name](data, no_levels=N)``."""
from .FAL_netB import *  # noqa: F401,F403

__all__ = ("FAL_netB",)
