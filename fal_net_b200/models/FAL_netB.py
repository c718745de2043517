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

import torch
import torch.nn as nn

from .. import backbone
from .. import med

__all__ = ["FAL_netB"]

# (name, cin, cout, stride): encoder stages of BackBone (/root/reference/models/FAL_netB.py:99-112)
_ENC = (("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 256, 2),
        ("conv4", 256, 256, 2), ("conv5", 256, 256, 2), ("conv6", 256, 512, 2))
# (level, up_in, up_out, skip_ch, iconv_out): decoder stages (:116-127); level 1's iconv has no bias/act
_DEC = ((6, 512, 256, 256, 256), (5, 256, 128, 256, 256), (4, 256, 128, 256, 256), (3, 256, 128, 128, 128),
        (2, 128, 64, 64, 64), (1, 64, 64, 32, None))


def FAL_netB(data=None, no_levels=49):
    model = FAL_net(batchNorm=False, no_levels=no_levels)
    if data is not None:
        model.load_state_dict(data["state_dict"])
    return model


def _conv(cin, cout, stride=1, bias=True, k=3):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)


class _Pair(nn.Module):
    """Parameter holder named like the reference's residual_block (conv1, conv2; :69-76)."""

    def __init__(self, ch):
        super().__init__()
        self.elu = nn.ELU(inplace=True)
        self.conv1 = _conv(ch, ch, bias=False)
        self.conv2 = _conv(ch, ch, bias=False)


class _Up(nn.Module):
    """Parameter holder named like the reference's deconv (conv1; :51-55)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.elu = nn.ELU(inplace=True)
        self.conv1 = _conv(cin, cout, bias=False)


class BackBone(nn.Module):
    """Holds the parameters under the reference's names; the compute lives in FAL_net.forward."""

    def __init__(self, batchNorm=False, no_in=3, no_flow=1, no_out=64):
        super().__init__()
        if batchNorm:
            raise NotImplementedError("FAL_netB is built with batchNorm=False (reference :29)")
        self.batchNorm = batchNorm
        for name, cin, cout, stride in _ENC:
            cin = no_in if name == "conv0" else (32 + no_flow if name == "conv1" else cin)
            self.add_module(name, nn.Sequential(_conv(cin, cout, stride), nn.ELU(inplace=True)))
            self.add_module(name + "_1", _Pair(cout))
        self.elu = nn.ELU(inplace=True)
        for lvl, uin, uout, skip, iout in _DEC:
            self.add_module(f"deconv{lvl}", _Up(uin, uout))
            if iout is not None:
                self.add_module(f"iconv{lvl}", nn.Sequential(_conv(uout + skip, iout), nn.ELU(inplace=True)))
            else:
                self.iconv1 = _conv(uout + skip, no_out, bias=False)
        # constructed but never used by forward, exactly like the reference (:128; SURVEY.md 7 "unused parameters")
        self.amask_conv = nn.Sequential(_conv(96, 48), nn.ELU(inplace=True), _conv(48, 1, bias=False), nn.Sigmoid())
        for m in self.modules():                                   # :131-138
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()


class FAL_net(nn.Module):
    def __init__(self, batchNorm, no_levels):
        super().__init__()
        self.no_levels = no_levels
        self.no_fac = 1
        self.backbone = BackBone(batchNorm, no_in=3, no_flow=1, no_out=self.no_levels)
        self.softmax = nn.Softmax(dim=1)
        self.elu = nn.ELU(inplace=True)
        self.sigmoid = nn.Sigmoid()
        self.conv0 = _conv(self.no_levels, self.no_fac * self.no_levels, bias=True, k=1)     # :190
        nn.init.kaiming_normal_(self.conv0.weight.data)
        self.conv0.bias.data.zero_()

    def weight_parameters(self):
        return [p for n, p in self.named_parameters() if "weight" in n]

    def bias_parameters(self):
        return [p for n, p in self.named_parameters() if "bias" in n]

    def used_parameters(self):
        """Parameters that receive gradient (everything except the never-called amask_conv)."""
        return [(n, p) for n, p in self.named_parameters() if "amask_conv" not in n]

    # ------------------------------------------------------------------------------------------
    def logits(self, input_left, max_disp):
        """dlog0 [B,N,H,W] fp32, planar with a 16-byte-multiple row pitch (what the MED kernels stream).  The whole
        encoder-decoder is one autograd node with a hand-scheduled backward (fal_net_b200.backbone)."""
        return backbone.logits(self, input_left, max_disp)

    def forward(self, input_left, min_disp, max_disp, ret_disp=True, ret_subocc=False, ret_pan=False):
        if ret_disp and not ret_subocc and not ret_pan and not (
                torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            return backbone.disparity(self, input_left, min_disp, max_disp)      # inference: fused disparity epilogue
        dlog0 = self.logits(input_left, max_disp)
        return med.med_section(dlog0, input_left, min_disp, max_disp, ret_disp, ret_subocc, ret_pan, zero_pad=True)
