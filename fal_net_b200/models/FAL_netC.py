"""FAL_netC on the B200-native kernels: drop-in for /root/reference/models/FAL_netC.py (factory :29-33, default
``no_levels=33``; wider bottleneck :110-120; the encoder-decoder is registered as ``synth`` :185, so checkpoints of the
reference load unchanged).  Same kernels, same hand-scheduled backward as FAL_netB (fal_net_b200.models._falnet)."""
from __future__ import annotations

from ._falnet import SPEC_C, build

__all__ = ["FAL_netC"]


def FAL_netC(data=None, no_levels=33):
    return build(SPEC_C, data, no_levels)
