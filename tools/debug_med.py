import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import falnet_oracle as O
from tests.helpers import med_case_inputs
from fal_net_b200 import med
dev = torch.device("cuda:0")
g = np.load("tests/golden/med_cases.npz")
for tag in "abcdef":
    B, N, H, W = (int(v) for v in g[f"{tag}_meta"])
    logits, img, gp, gd = med_case_inputs(tag, B, N, H, W)
    mn, mx = torch.from_numpy(g[f"{tag}_min"]), torch.from_numpy(g[f"{tag}_max"])
    d, xo = O.level_tables(mn, mx, N, W)
    g0x = O.identity_grid(1, 1, 2, W)[0, 0, :, 0].contiguous().to(dev)
    for flags in (0, 1):
        r = med.med_forward_raw(logits.to(dev), img.to(dev), xo.to(dev), d.to(dev), g0x, True, True, True, flags)
        gl = med.med_backward_raw(logits.to(dev), img.to(dev), xo.to(dev), d.to(dev), g0x, r["pan"], r["disp"], r["lse0"], r["lsew"], gp.to(dev), gd.to(dev), flags)
        out = []
        for nm, t in (("pan", r["pan"]), ("disp", r["disp"]), ("maskL", r["maskL"]), ("maskR", r["maskR"]), ("glogits", gl)):
            ref = torch.from_numpy(g[f"{tag}_{nm}"])
            diff = (t.cpu() - ref).abs()
            idx = np.unravel_index(int(diff.argmax()), diff.shape)
            out.append(f"{nm} {float(diff.max() / ref.abs().max()):.2e}@{tuple(int(i) for i in idx)}")
        print(tag, (B, N, H, W), "flags", flags, " | ".join(out))
