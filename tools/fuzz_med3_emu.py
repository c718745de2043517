#!/usr/bin/env python
"""Random differential test of the host-emulated third-generation MED kernels against the CPU oracle (test infrastructure;
the CPU suite runs a 12-example property test of the same kind).  usage: python tools/fuzz_med3_emu.py [seed] [cases]"""
import sys, os, ctypes, torch, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, rel_err
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'host_emu', 'libmed3_emu.so'))
fp = ctypes.c_void_p
L.emu_med3_fwd.argtypes = [fp]*11 + [ctypes.c_int]*6
L.emu_med3_bwd.argtypes = [fp]*12 + [ctypes.c_int]*5
P = lambda t: ctypes.c_void_p(t.data_ptr())
import warnings; warnings.filterwarnings("ignore")
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = {}
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 300):
    W = random.choice([random.randint(8, 64), random.randint(64, 400), random.choice([127,128,129,130,131,132,255,256,257,512,640])])
    N = random.randint(2, 40); B = random.randint(1, 2); H = 1
    maxd = random.uniform(0.5, 1.3 * W); ratio = random.uniform(1.2, 300.0)
    seed = random.randint(0, 10**6)
    g = torch.Generator().manual_seed(seed)
    logits = (random.choice([0.5, 2.0, 8.0]) * torch.randn(B, N, H, W, generator=g)).contiguous()
    img = images(B, H, W, seed + 1).contiguous()
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mn, mx = disp_range(B, maxd, maxd / ratio)
    d, xo = O.level_tables(mn, mx, N, W); d, xo = d.contiguous(), xo.contiguous()
    ref = O.med_forward_closed(logits, img, d, xo)
    ref["glogits"] = O.med_backward_closed(logits, img, d, xo, gp, gd)
    g0x = O.identity_grid(1, 1, 2, W)[0, 0, :, 0].contiguous()
    out = {k: torch.full((B, c, H, W), float("nan")) for k, c in (("pan",3),("disp",1),("maskL",1),("maskR",1),("lse0",1),("lsew",1))}
    fl = L.emu_med3_fwd(P(logits),P(img),P(g0x),P(xo),P(d),P(out["pan"]),P(out["disp"]),P(out["maskL"]),P(out["maskR"]),P(out["lse0"]),P(out["lsew"]),B,N,H,W,1,0)
    gl = torch.full((B,N,H,W), float("nan"))
    if fl == 0:
        L.emu_med3_bwd(P(logits),P(img),P(g0x),P(xo),P(d),P(out["pan"]),P(out["disp"]),P(out["lse0"]),P(out["lsew"]),P(gp.contiguous()),P(gd.contiguous()),P(gl),B,N,H,W,0)
    else:
        continue
    out["glogits"] = gl
    for nm in ("pan","disp","maskL","maskR","glogits"):
        e = rel_err(out[nm], ref[nm])
        if e > worst.get(nm, (0,))[0]: worst[nm] = (e, W, N, maxd, ratio, seed)
        if not (e < 1e-4): print("FAIL", nm, e, dict(W=W,N=N,B=B,maxd=maxd,ratio=ratio,seed=seed))
print("worst", worst)
