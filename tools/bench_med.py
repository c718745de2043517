#!/usr/bin/env python
"""MED kernel microbench (BASELINE.json config #5): fwd (pan+disp+stats), fwd+masks, bwd.
Algorithmic bytes: fwd 4(N+9) B/px (4(N+7) without masks), bwd 4(2N+7) B/px.
Rotates buffer sets so L2 (126 MB) cannot serve repeats; CUDA-event timing on the launch stream."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fal_net_b200 import layout, med  # noqa: E402


def bench(B, N, H, W, iters=10, sets=3, peak=6557.8, flags=0, warm=3):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    # logits in the layout the network's last conv writes: 16-byte-multiple row pitch, zeroed pad columns
    L = []
    for _ in range(sets):
        t = layout.alloc_planar(B, N, H, W, dev)
        t.copy_(2 * torch.randn(B, N, H, W, generator=gen, device=dev))
        L.append(t)
    flags |= med.FLAG_ZERO_PAD
    I = [torch.rand(B, 3, H, W, generator=gen, device=dev) - 0.43 for _ in range(sets)]
    gp = torch.randn(B, 3, H, W, generator=gen, device=dev)
    gd = torch.randn(B, 1, H, W, generator=gen, device=dev)
    mx = torch.full((B, 1, 1), 300.0, device=dev)
    mn = mx * 2 / 300
    d, xo = med.level_tables(mn, mx, N, W)
    g0x = med.grid_row(W, dev)
    px = B * H * W
    out = {}
    res = [med.med_forward_raw(L[s], I[s], xo, d, g0x, True, True, True) for s in range(sets)]
    gl = layout.alloc_planar(B, N, H, W, dev)

    def timeit(fn):
        for s in range(warm):
            fn(s % sets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for i in range(iters):
            fn(i % sets)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    t = timeit(lambda s: med.med_forward_raw(L[s], I[s], xo, d, g0x, True, True, False, flags))
    out["fwd"] = dict(ms=t, gbs=4 * (N + 7) * px / t / 1e6)
    t = timeit(lambda s: med.med_forward_raw(L[s], I[s], xo, d, g0x, True, True, True, flags))
    out["fwd_masks"] = dict(ms=t, gbs=4 * (N + 9) * px / t / 1e6)
    t = timeit(lambda s: med.med_backward_raw(L[s], I[s], xo, d, g0x, res[s]["pan"], res[s]["disp"], res[s]["lse0"],
                                              res[s]["lsew"], gp, gd, flags, out=gl))
    out["bwd"] = dict(ms=t, gbs=4 * (2 * N + 7) * px / t / 1e6)
    t = timeit(lambda s: med.med_disp_only(L[s], d))
    out["disp_only"] = dict(ms=t, gbs=4 * (N + 1) * px / t / 1e6)
    for k in out:
        out[k]["frac_measured_peak"] = out[k]["gbs"] / peak
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--profile", default="", help="B,N,H,W: run each kernel a few times only (for ncu)")
    ap.add_argument("--flags", type=int, default=0)
    a = ap.parse_args()
    if a.profile:
        B, N, H, W = (int(v) for v in a.profile.split(","))
        print(json.dumps(bench(B, N, H, W, iters=1, sets=2, flags=a.flags, warm=1)))
        sys.exit(0)
    peak = 6557.8
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except Exception:
        pass
    cfgs = [(16, 49, 192, 640), (8, 49, 375, 1242), (2, 49, 1024, 2048)]
    if not a.quick:
        cfgs += [(8, 33, 375, 1242), (8, 65, 375, 1242), (2, 33, 1024, 2048), (2, 65, 1024, 2048)]
    for B, N, H, W in cfgs:
        r = bench(B, N, H, W, peak=peak, flags=a.flags)
        print(json.dumps(dict(B=B, N=N, H=H, W=W, **{k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in r.items()})))
