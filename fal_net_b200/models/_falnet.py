"""Spec-driven FAL-net family on the B200-native kernels: the parameter holders and ``forward`` shared by FAL_netA / B / C.

A variant is a ``Spec``: encoder / decoder channel tables, the attribute name the reference gives its encoder-decoder
(``backbone`` in B, ``synth`` in C, ``BackBone`` in A -- it is part of every ``state_dict`` key), whether the residual
blocks use separable 3x1 / 1x3 kernels (A, /root/reference/models/FAL_netA.py:73-76), whether the never-called
``amask_conv`` exists (B, C: /root/reference/models/FAL_netB.py:128) and whether ``maskR`` is sampled with
``align_corners=True`` (B, C) or the grid_sample default (A, /root/reference/models/FAL_netA.py:264).
Modules are constructed in the reference's order with the reference's initialisers, so ``torch.manual_seed(s)`` followed by
the factory reproduces the reference's initial weights (tests/golden/init_checksums.npz, variants.npz).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import backbone
from .. import med


@dataclass(frozen=True)
class Spec:
    name: str
    enc: tuple            # (name, cin, cout, stride): /root/reference/models/FAL_netB.py:99-112
    dec: tuple            # (level, up_in, up_out, skip_ch, iconv_out): :116-127; level 1's iconv has no bias / activation
    bb_attr: str = "backbone"
    separable: bool = False
    amask: bool = True
    maskr_align_corners: bool = True
    default_levels: int = 49


SPEC_B = Spec("FAL_netB",
              enc=(("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 256, 2),
                   ("conv4", 256, 256, 2), ("conv5", 256, 256, 2), ("conv6", 256, 512, 2)),
              dec=((6, 512, 256, 256, 256), (5, 256, 128, 256, 256), (4, 256, 128, 256, 256), (3, 256, 128, 128, 128),
                   (2, 128, 64, 64, 64), (1, 64, 64, 32, None)))
# /root/reference/models/FAL_netC.py:110-120: wider bottleneck; encoder-decoder registered as ``synth`` (:185)
SPEC_C = Spec("FAL_netC",
              enc=(("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 256, 2),
                   ("conv4", 256, 256, 2), ("conv5", 256, 512, 2), ("conv6", 512, 512, 2)),
              dec=((6, 512, 256, 512, 512), (5, 512, 256, 256, 256), (4, 256, 128, 256, 256), (3, 256, 128, 128, 128),
                   (2, 128, 64, 64, 64), (1, 64, 64, 32, None)),
              bb_attr="synth", default_levels=33)
# /root/reference/models/FAL_netA.py:99-126: narrower, separable residual blocks, no amask_conv, ``BackBone`` attribute (:183)
SPEC_A = Spec("FAL_netA",
              enc=(("conv0", 3, 32, 1), ("conv1", 33, 64, 2), ("conv2", 64, 128, 2), ("conv3", 128, 128, 2),
                   ("conv4", 128, 256, 2), ("conv5", 256, 256, 2), ("conv6", 256, 256, 2)),
              dec=((6, 256, 128, 256, 256), (5, 256, 128, 256, 256), (4, 256, 128, 128, 128), (3, 128, 64, 128, 128),
                   (2, 128, 64, 64, 64), (1, 64, 64, 32, None)),
              bb_attr="BackBone", separable=True, amask=False, maskr_align_corners=False, default_levels=33)


def _conv(cin, cout, stride=1, bias=True, k=3):
    if isinstance(k, tuple):
        return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=((k[0] - 1) // 2, (k[1] - 1) // 2), bias=bias)
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)


class _Pair(nn.Module):
    """Parameter holder named like the reference's residual_block (conv1, conv2; :69-76)."""

    def __init__(self, ch, separable=False):
        super().__init__()
        self.elu = nn.ELU(inplace=True)
        self.conv1 = _conv(ch, ch, bias=False, k=(3, 1) if separable else 3)
        self.conv2 = _conv(ch, ch, bias=False, k=(1, 3) if separable else 3)


class _Up(nn.Module):
    """Parameter holder named like the reference's deconv (conv1; :51-55)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.elu = nn.ELU(inplace=True)
        self.conv1 = _conv(cin, cout, bias=False)


class BackBone(nn.Module):
    """Holds the parameters under the reference's names; the compute lives in fal_net_b200.backbone."""

    def __init__(self, spec: Spec, batchNorm=False, no_in=3, no_flow=1, no_out=64):
        super().__init__()
        if batchNorm:
            raise NotImplementedError("the FAL-net factories build with batchNorm=False (reference :29)")
        self.batchNorm = batchNorm
        for name, cin, cout, stride in spec.enc:
            cin = no_in if name == "conv0" else (32 + no_flow if name == "conv1" else cin)
            self.add_module(name, nn.Sequential(_conv(cin, cout, stride), nn.ELU(inplace=True)))
            self.add_module(name + "_1", _Pair(cout, spec.separable))
        self.elu = nn.ELU(inplace=True)
        for lvl, uin, uout, skip, iout in spec.dec:
            self.add_module(f"deconv{lvl}", _Up(uin, uout))
            if iout is not None:
                self.add_module(f"iconv{lvl}", nn.Sequential(_conv(uout + skip, iout), nn.ELU(inplace=True)))
            else:
                self.iconv1 = _conv(uout + skip, no_out, bias=False)
        if spec.amask:
            # constructed but never used by forward, exactly like the reference (:128; SURVEY.md 7 "unused parameters")
            self.amask_conv = nn.Sequential(_conv(96, 48), nn.ELU(inplace=True), _conv(48, 1, bias=False), nn.Sigmoid())
        for m in self.modules():                                   # :131-138
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()


class FAL_net(nn.Module):
    def __init__(self, batchNorm, no_levels, spec: Spec = SPEC_B):
        super().__init__()
        self._spec = spec
        self.no_levels = no_levels
        self.no_fac = 1
        setattr(self, spec.bb_attr, BackBone(spec, batchNorm, no_in=3, no_flow=1, no_out=self.no_levels))
        self.softmax = nn.Softmax(dim=1)
        self.elu = nn.ELU(inplace=True)
        self.sigmoid = nn.Sigmoid()
        self.conv0 = _conv(self.no_levels, self.no_fac * self.no_levels, bias=True, k=1)     # :190
        nn.init.kaiming_normal_(self.conv0.weight.data)
        self.conv0.bias.data.zero_()

    @property
    def bb(self):
        """The encoder-decoder parameter holder, whatever the variant calls it."""
        return getattr(self, self._spec.bb_attr)

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
        return med.med_section(dlog0, input_left, min_disp, max_disp, ret_disp, ret_subocc, ret_pan, zero_pad=True,
                               maskr_align_corners=self._spec.maskr_align_corners)


def build(spec: Spec, data=None, no_levels=None):
    model = FAL_net(batchNorm=False, no_levels=spec.default_levels if no_levels is None else no_levels, spec=spec)
    if data is not None:
        model.load_state_dict(data["state_dict"])
    return model
