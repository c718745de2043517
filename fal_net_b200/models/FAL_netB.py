"""FAL_netB on the B200-native kernels.

Drop-in for /root/reference/models/FAL_netB.py: same factory (``FAL_netB(data=None, no_levels=49)``,
:28-32), same ``forward(input_left, min_disp, max_disp, ret_disp, ret_subocc, ret_pan)`` signature and
return convention (:200, :228-229, :285-297), same ``weight_parameters()`` / ``bias_parameters()``
(:194-198) and the same ``state_dict`` keys / shapes / initial random stream (:130-138, :190-192), so
reference checkpoints load unchanged and ``torch.manual_seed(s); FAL_netB()`` gives the reference's
weights.  What differs is everything that executes:

  * the encoder-decoder runs in bf16 NHWC with fp32 accumulation through ``fal_net_b200.conv``
    (tcgen05 implicit-GEMM kernels; bias / ELU / residual fused in the epilogue),
  * the logit 1x1 conv (:190,215) is folded into the last 3x3 conv (both are linear, no activation
    between them: W' = W0 . W_iconv1, exact up to fp reassociation) so ``dlog`` never exists,
  * the whole MED section (:216-297) is the fused kernel pair of ``fal_net_b200.med``.

There is no CPU path: ``forward`` raises if the input is not on a CUDA device.
"""
from __future__ import annotations

from ._falnet import SPEC_B, BackBone, FAL_net, build  # noqa: F401

__all__ = ["FAL_netB"]


def FAL_netB(data=None, no_levels=49):
    return build(SPEC_B, data, no_levels)
