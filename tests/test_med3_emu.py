"""Host emulation of the third-generation MED kernels' per-thread code (fal_net_b200/csrc/med3_core.cuh, compiled with
g++ by tests/host_emu/med3_emu.cpp) against the CPU oracle.  This pins the window / clamp / alignment-class logic of
csrc/med3.cu on the CPU; the GPU parity tests (tests/test_med_gpu.py) then run the same functions inside the kernels.
Tolerance: 1e-4 relative (max|a-b| / max|b|), BASELINE.json's fp32 bound."""
import ctypes
import os
import subprocess

import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import falnet_oracle as O
from tests.helpers import disp_range, images, rel_err

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")
TOL = 1e-4


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "med3_emu.cpp")
    lib = os.path.join(HERE, "libmed3_emu.so")
    core = os.path.join(HERE, "..", "..", "fal_net_b200", "csrc", "med3_core.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", lib, src])
    L = ctypes.CDLL(lib)
    fp = ctypes.c_void_p
    L.emu_med3_fwd.argtypes = [fp] * 11 + [ctypes.c_int] * 6
    L.emu_med3_bwd.argtypes = [fp] * 12 + [ctypes.c_int] * 5
    return L


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _run(emu, logits, img, d, xo, gp, gd, force=0):
    B, N, H, W = logits.shape
    g0x = O.identity_grid(1, 1, 2, W)[0, 0, :, 0].contiguous()
    out = {k: torch.full((B, c, H, W), float("nan")) for k, c in
           (("pan", 3), ("disp", 1), ("maskL", 1), ("maskR", 1), ("lse0", 1), ("lsew", 1))}
    logits, img, d, xo = (t.contiguous() for t in (logits, img, d, xo))
    flagged = emu.emu_med3_fwd(_p(logits), _p(img), _p(g0x), _p(xo), _p(d), _p(out["pan"]), _p(out["disp"]),
                               _p(out["maskL"]), _p(out["maskR"]), _p(out["lse0"]), _p(out["lsew"]), B, N, H, W, 1, force)
    gl = torch.full((B, N, H, W), float("nan"))
    if flagged == 0:
        gp, gd = gp.contiguous(), gd.contiguous()
        emu.emu_med3_bwd(_p(logits), _p(img), _p(g0x), _p(xo), _p(d), _p(out["pan"]), _p(out["disp"]), _p(out["lse0"]),
                         _p(out["lsew"]), _p(gp), _p(gd), _p(gl), B, N, H, W, force)
    out["glogits"] = gl
    return out, flagged


@pytest.mark.parametrize("B,N,H,W,maxd,mind", [
    (2, 49, 3, 640, 300.0, 2.0),
    (1, 49, 2, 1242, 300.0, 2.0),        # W % 4 == 2: ragged last quad, pad columns
    (2, 33, 2, 321, 120.0, 1.5),         # W % 4 == 1
    (1, 17, 3, 100, 40.0, 0.5),
    (1, 65, 1, 2048, 300.0, 2.0),
    (2, 2, 3, 8, 3.0, 1.0),              # smallest supported: N = 2, W = 8
    (1, 9, 2, 40, 18.0, 0.3),            # shifts up to half the row, k0 = 0 planes
    (1, 12, 2, 64, 70.0, 3.0),           # shifts larger than the row: every tap out of range on the far planes
])
def test_emulated_kernels_match_oracle(emu, B, N, H, W, maxd, mind):
    g = torch.Generator().manual_seed(B * 1000 + N * 10 + W)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 99 + W)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mn, mx = disp_range(B, maxd, mind)
    d, xo = O.level_tables(mn, mx, N, W)
    ref = O.med_forward_closed(logits, img, d, xo)
    ref["glogits"] = O.med_backward_closed(logits, img, d, xo, gp, gd)
    out, flagged = _run(emu, logits, img, d, xo, gp, gd)
    assert flagged == 0
    for nm in ("pan", "disp", "maskL", "maskR", "lse0", "lsew", "glogits"):
        e = rel_err(out[nm], ref[nm])
        assert e < TOL, (nm, e)


def test_per_sample_level_tables(emu):
    """Every sample with its own disparity range (its own class-sorted plane table in the kernels), N = 65."""
    B, N, H, W = 3, 65, 2, 322
    g = torch.Generator().manual_seed(21)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 8)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mx = torch.tensor([300.0, 120.0, 33.3]).view(B, 1, 1)
    mn = torch.tensor([2.0, 0.7, 1.1]).view(B, 1, 1)
    d, xo = O.level_tables(mn, mx, N, W)
    ref = O.med_forward_closed(logits, img, d, xo)
    ref["glogits"] = O.med_backward_closed(logits, img, d, xo, gp, gd)
    out, flagged = _run(emu, logits, img, d, xo, gp, gd)
    assert flagged == 0
    for nm in ("pan", "disp", "maskL", "maskR", "lse0", "lsew", "glogits"):
        e = rel_err(out[nm], ref[nm])
        assert e < TOL, (nm, e)


def test_generic_code_matches_fast_code(emu):
    """Every plane forced onto the per-pixel generic functions: same results as the class-specialised windows."""
    B, N, H, W = 1, 21, 2, 322
    g = torch.Generator().manual_seed(9)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 5)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mn, mx = disp_range(B, 150.0, 1.0)
    d, xo = O.level_tables(mn, mx, N, W)
    fast, f0 = _run(emu, logits, img, d, xo, gp, gd)
    gen, f1 = _run(emu, logits, img, d, xo, gp, gd, force=1)
    assert f0 == 0 and f1 == 0
    for nm in ("pan", "disp", "maskL", "maskR", "glogits"):
        assert rel_err(gen[nm], fast[nm]) < 2e-5, nm


def test_integer_and_near_integer_shifts(emu):
    """Hand-made level tables whose pixel shift is exactly / almost an integer (tests/test_med_gpu.py has the GPU twin):
    those planes take the generic code, the others the class-specialised windows, in one row."""
    B, N, H, W = 1, 12, 2, 640
    g = torch.Generator().manual_seed(11)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 17)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    ks = torch.tensor([0.0, 1.0, 2.0, 5.0, 17.0, 64.0, 100.0, 255.0, 300.0, 638.0, 639.0, 700.0])
    eps = torch.tensor([0, 1e-7, -1e-7, 3e-5, -3e-5, 0, 1e-4, -1e-4, 0, 0, 0, 0])
    xo = ((ks + eps) * 2.0 / (W - 1)).float().view(1, N)
    d = (ks + 1.0).view(1, N).float()
    ref = O.med_forward_closed(logits, img, d, xo)
    ref["glogits"] = O.med_backward_closed(logits, img, d, xo, gp, gd)
    out, flagged = _run(emu, logits, img, d, xo, gp, gd)
    assert flagged == 0
    for nm in ("pan", "disp", "maskL", "maskR", "glogits"):
        e = rel_err(out[nm], ref[nm])
        assert e < TOL, (nm, e)


def test_overflow_rows_are_flagged(emu):
    """Huge logits: the max-free sums overflow, the row is marked (NaN in lse0[row, 0]) for the clean-up kernel."""
    B, N, H, W = 1, 6, 2, 64
    g = torch.Generator().manual_seed(3)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, 5)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mn, mx = disp_range(B, 20.0, 1.0)
    d, xo = O.level_tables(mn, mx, N, W)
    big = logits.clone()
    big[0, 2, 1, 10] = 200.0            # exp(200) overflows fp32 without a running maximum
    out, flagged = _run(emu, big, img, d, xo, gp, gd)
    assert flagged == 1 and torch.isnan(out["lse0"][0, 0, 1, 0]) and not torch.isnan(out["lse0"][0, 0, 0, 0])
    # wider row: the mark sits on the first pixel of the warp (128 pixels) that saw the overflow
    B, N, H, W = 1, 5, 1, 300
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    logits[0, 1, 0, 200] = 250.0
    img = images(B, H, W, 6)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    d, xo = O.level_tables(*disp_range(B, 20.0, 1.0), N, W)
    out, flagged = _run(emu, logits, img, d, xo, gp, gd)
    marks = torch.isnan(out["lse0"][0, 0, 0]).nonzero().flatten().tolist()
    assert flagged == 1 and 128 in marks and all(m % 128 == 0 for m in marks)


@settings(max_examples=12, deadline=None)
@given(W=st.integers(8, 200), N=st.integers(2, 24), maxd=st.floats(1.0, 260.0), ratio=st.floats(1.5, 200.0),
       seed=st.integers(0, 10_000))
def test_random_shapes_and_disparity_ranges(emu, W, N, maxd, ratio, seed):
    """Property test: arbitrary widths (every W % 4, rows shorter than a warp, shifts beyond the row), plane counts and
    disparity ranges -- the emulated kernels agree with the oracle, and with their own generic per-pixel code."""
    B, H = 2, 1
    g = torch.Generator().manual_seed(seed)
    logits = 2 * torch.randn(B, N, H, W, generator=g)
    img = images(B, H, W, seed + 1)
    gp, gd = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    mn, mx = disp_range(B, maxd, maxd / ratio)
    d, xo = O.level_tables(mn, mx, N, W)
    ref = O.med_forward_closed(logits, img, d, xo)
    ref["glogits"] = O.med_backward_closed(logits, img, d, xo, gp, gd)
    out, flagged = _run(emu, logits, img, d, xo, gp, gd)
    assert flagged == 0
    for nm in ("pan", "disp", "maskL", "maskR", "glogits"):
        e = rel_err(out[nm], ref[nm])
        assert e < TOL, (nm, e, W, N, maxd, ratio, seed)
