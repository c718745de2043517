"""FAL_netA on the B200-native kernels: drop-in for /root/reference/models/FAL_netA.py (factory :28-32, default
``no_levels=33``; narrower encoder-decoder :99-126 registered as ``BackBone`` :183; residual blocks with separable 3x1 / 1x3
kernels :73-76 -- run on the 3x3 tcgen05 kernels with the absent taps zero; no ``amask_conv``; ``maskR`` sampled with
grid_sample's default ``align_corners=False`` :264, which the product reproduces with a dedicated kernel)."""
from __future__ import annotations

from ._falnet import SPEC_A, build

__all__ = ["FAL_netA"]


def FAL_netA(data=None, no_levels=33):
    return build(SPEC_A, data, no_levels)
