"""CPU oracle of the training input pipeline  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as falnet_oracle.py).

Restates in numpy (integer arithmetic, bit-exact):
  * Pillow 8-bit bicubic ``Image.resize`` as the reference calls it (/root/reference/data_transforms.py:67): the algorithm
    lives in a THIRD-PARTY dependency, Pillow (requirements: un-pinned; installed here: 12.2.0), file src/libImaging/Resample.c
    -- ``precompute_coeffs`` (double-precision bicubic a = -0.5, support 2 x max(scale, 1), window
    [int(c - s + .5), int(c + s + .5)) clipped to the image, weights normalised to sum 1), ``normalize_coeffs_8bpc``
    (22 fractional bits, round half away from zero) and the two 8-bit passes (horizontal, then vertical; accumulator starts
    at 2^21; result clamp(acc >> 22, 0, 255) after EACH pass).  Pinned against Pillow itself in tests/test_input_pipeline.py.
  * the co-transforms + input transform of the reference (/root/reference/data_transforms.py:46-157,
    /root/reference/Train_Stage1_K.py:124-128) with explicit parameters.  Pinned against the reference's own classes run
    with seeded RNGs (tests/golden/pipeline.npz, generator tests/golden/make_golden_pipeline.py).
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size, out_size):
    """(bounds [out,2], coeffs [out,ksize] int) -- Resample.c precompute_coeffs + normalize_coeffs_8bpc."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8-bit resampling pass along ``axis`` (0: vertical, 1: horizontal) of a uint8 [H,W,C] array."""
    src = img.astype(np.int64)
    if axis == 0:
        src = src.transpose(1, 0, 2)
    out = np.empty((src.shape[0], bounds.shape[0], src.shape[2]), dtype=np.int64)
    for xx in range(bounds.shape[0]):
        xmin, n = bounds[xx]
        acc = (src[:, xmin:xmin + n, :] * kk[xx, :n][None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = out.astype(np.uint8)
    return out.transpose(1, 0, 2) if axis == 0 else out


def pil_resize_bicubic(img, out_w, out_h):
    """Image.fromarray(img).resize((out_w, out_h), Image.BICUBIC) for a uint8 [H,W,3] array."""
    h, w, _ = img.shape
    out = img
    if out_w != w:
        out = _pass(out, *pil_coeffs(w, out_w), axis=1)
    if out_h != h:
        out = _pass(out, *pil_coeffs(h, out_h), axis=0)
    return out


def augment_pair(left, right, factor, x1, y1, flip, gamma, bright, cbright, size, mean=(0.411, 0.432, 0.45)):
    """The reference's co-transform chain + input transform for one pair with EXPLICIT parameters
    (data_transforms.py:61-75, 99-103, 125-129, 141-145, 157-163; Train_Stage1_K.py:124-128).  Returns two float32
    [3,th,tw] tensors."""
    th, tw = size
    inputs = [left, right]
    h, w, _ = inputs[0].shape
    inputs = [pil_resize_bicubic(a, int(w * factor), int(h * factor)) for a in inputs]
    inputs = [a[y1:y1 + th, x1:x1 + tw] for a in inputs]
    if flip:
        inputs = [np.copy(np.fliplr(inputs[1])), np.copy(np.fliplr(inputs[0]))]
    if gamma is not None:
        inputs = [255 * ((a / 255) ** gamma) for a in inputs]
    if bright is not None:
        inputs = [a * bright for a in inputs]
        for a in inputs:
            a[a > 255] = 255
    if cbright is not None:
        inputs = [np.array(a) for a in inputs]
        with np.errstate(invalid="ignore", over="ignore"):
            for i in range(2):
                for c in range(3):
                    inputs[i][:, :, c] = inputs[i][:, :, c] * cbright[i][c]
                inputs[i][inputs[i] > 255] = 255
    out = []
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
    for a in inputs:
        t = torch.from_numpy(np.transpose(a, (2, 0, 1)).copy()).float()
        t = (t - 0.0) / 255.0
        out.append((t - m) / 1.0)
    return out
