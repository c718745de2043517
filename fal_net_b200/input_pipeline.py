"""Training input pipeline on the device (csrc/input_pipe.cu): drop-in for the co-transform + input transform the
reference runs in 4 DataLoader workers (/root/reference/data_transforms.py:46-157, Train_Stage1_K.py:115-128).

    aug = GpuStereoAugment((192, 640))                       # same defaults as Train_Stage1_K.py:116-122
    left, right = aug(lefts_u8, rights_u8)                    # lists of uint8 [H,W,3] DEVICE tensors -> fp32 [B,3,192,640]

What stays on the host is what the reference also does per sample before touching pixels: drawing the random parameters
(`sample_params` consumes Python's ``random`` and ``numpy.random`` in the reference's call order, so equal seeds give equal
augmentations) and two tiny tables per image -- Pillow's fixed-point bicubic coefficients for the crop's rows / columns
(`faln_pil_bicubic_coeffs`, C) and the 3 x 256 value table of the gamma / brightness / colour / normalisation chain.  They
go to the device in ONE pinned upload per batch; all pixel work is two batched kernel launches.  Results are bit-identical
to the reference pipeline (tests/test_input_pipeline.py).  No CPU pixel path.
"""
from __future__ import annotations

import ctypes
import random as _random

import numpy as np
import torch

from . import _lib

MEAN = (0.411, 0.432, 0.45)


class AugParams:
    """One sample's draw.  factor: resize factor; x1, y1: crop origin in the resized image; flip: mirror + swap views;
    gamma / bright: None or factor; cbright: None or [[3 floats] per view]."""

    def __init__(self, factor, x1, y1, flip, gamma, bright, cbright, swap_lr=False):
        self.factor, self.x1, self.y1, self.flip = factor, x1, y1, flip
        self.gamma, self.bright, self.cbright, self.swap_lr = gamma, bright, cbright, swap_lr


def sample_params(h, w, size, down=0.75, up=1.5, gamma=(0.8, 1.2), bright=(0.5, 2.0), cbright=(0.8, 1.2), fix_order=True,
                  max_pix=300.0, rng=_random, nprng=np.random):
    """The random draws of one training sample in the reference's order: listdataset_train.py:75-86 (view order, y_pix),
    then data_transforms.py RandomResizeCrop :61-63,74-75, RandomHorizontalFlip :99, RandomGamma :125-126,
    RandomBrightness :141-142, RandomCBrightness :157-160."""
    th, tw = size
    swap = not (rng.random() < 0.5 or fix_order)
    nprng.uniform(low=-max_pix, high=max_pix)                        # y_pix (drawn, unused by the training loop)
    min_factor = max(max((th + 1) / h, (tw + 1) / w), down)
    factor = nprng.uniform(low=min_factor, high=up)
    rw, rh = int(w * factor), int(h * factor)
    x1 = rng.randint(0, rw - tw)
    y1 = rng.randint(0, rh - th)
    flip = rng.random() < 0.5
    g = rng.uniform(*gamma) if rng.random() < 0.5 else None
    b = rng.uniform(*bright) if rng.random() < 0.5 else None
    cb = None
    if rng.random() < 0.5:
        cb = [[rng.uniform(*cbright) for _ in range(3)] for _ in range(2)]
    return AugParams(factor, x1, y1, flip, g, b, cb, swap)


def value_table(gamma, bright, cbright3, mean=MEAN):
    """[3,256] float32: what the reference's value chain makes of a uint8 pixel of each channel -- RandomGamma
    (255 * (x / 255) ** g, float64), RandomBrightness (x * f, clamp 255), RandomCBrightness (x[..., c] * f_c written back
    IN PLACE, i.e. truncated and wrapped into uint8 when no earlier step promoted the array to float64 -- the reference's
    behaviour, kept), ArrayToTensor .float(), Normalize(0, 255), Normalize(mean, 1) in float32."""
    arr = np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, axis=2)       # a 256 x 1 "image" holding every value
    if gamma is not None:
        arr = 255 * ((arr / 255) ** gamma)
    if bright is not None:
        arr = arr * bright
        arr[arr > 255] = 255
    if cbright3 is not None:
        with np.errstate(invalid="ignore", over="ignore"):
            for c in range(3):
                arr[:, :, c] = arr[:, :, c] * cbright3[c]
        arr[arr > 255] = 255
    t = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 0, 1))).float()      # [3,256,1]
    t = (t - 0.0) / 255.0
    t = (t - torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)) / 1.0
    return t[:, :, 0].contiguous()


class _Desc(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int), ("row0", ctypes.c_int),
                ("rows", ctypes.c_int), ("x_tab", ctypes.c_int), ("y_tab", ctypes.c_int), ("ksx", ctypes.c_int),
                ("ksy", ctypes.c_int), ("flip", ctypes.c_int), ("dst", ctypes.c_int), ("lut", ctypes.c_int),
                ("pad_", ctypes.c_int)]


def pil_bicubic_tables(in_size, out_size, lo, n):
    """(bounds [n,2] int32, coeffs [n,ksize] int32) of Pillow's bicubic resize in_size -> out_size, outputs [lo, lo+n)."""
    L = _lib.lib()
    ks = L.faln_pil_bicubic_ksize(int(in_size), int(out_size))
    bounds = np.empty((n, 2), dtype=np.int32)
    coeffs = np.empty((n, ks), dtype=np.int32)
    _lib.check(L.faln_pil_bicubic_coeffs(int(in_size), int(out_size), int(lo), int(n),
                                         bounds.ctypes.data_as(ctypes.c_void_p), coeffs.ctypes.data_as(ctypes.c_void_p)),
               "faln_pil_bicubic_coeffs")
    return bounds, coeffs


class GpuStereoAugment:
    """RandomResizeCrop + RandomHorizontalFlip + RandomGamma + RandomBrightness + RandomCBrightness + ArrayToTensor +
    Normalize of the reference (Train_Stage1_K.py:115-128) for a batch of stereo pairs, on the device."""

    def __init__(self, size, down=0.75, up=1.5, gamma=(0.8, 1.2), bright=(0.5, 2.0), cbright=(0.8, 1.2), mean=MEAN,
                 fix_order=True, max_pix=300.0):
        self.size = (int(size[0]), int(size[1]))
        self.cfg = dict(down=down, up=up, gamma=gamma, bright=bright, cbright=cbright, fix_order=fix_order, max_pix=max_pix)
        self.mean = mean

    def sample(self, h, w, rng=_random, nprng=np.random):
        return sample_params(h, w, self.size, rng=rng, nprng=nprng, **self.cfg)

    def plan(self, shapes, params, ptrs=None):
        """Host half of a batch: per-image descriptors, the packed coefficient tables and value tables.
        shapes: [(h, w)] per pair; params: [AugParams]; ptrs: [(left_ptr, right_ptr)] device addresses (0 when only
        planning).  Returns dict(descs, tabs int32, luts float32 [n,3,256], max_rows, n_img)."""
        th, tw = self.size
        B = len(params)
        descs = (_Desc * (2 * B))()
        tabs, luts = [], []
        t_off = l_off = 0
        max_rows = 1
        for i, p in enumerate(params):
            h, w = shapes[i]
            rw, rh = int(w * p.factor), int(h * p.factor)
            bx, kx = pil_bicubic_tables(w, rw, p.x1, tw)
            by, ky = pil_bicubic_tables(h, rh, p.y1, th)
            row0 = int(by[:, 0].min())
            rows = int((by[:, 0] + by[:, 1]).max()) - row0
            max_rows = max(max_rows, rows)
            x_tab = t_off
            tabs += [bx.reshape(-1), kx.reshape(-1)]
            t_off += bx.size + kx.size
            y_tab = t_off
            tabs += [by.reshape(-1), ky.reshape(-1)]
            t_off += by.size + ky.size
            for v in range(2):
                src_view = (1 - v) if p.swap_lr else v               # listdataset_train.py:75-82: random view order
                # RandomHorizontalFlip mirrors both views AND swaps them (data_transforms.py:100-101): view v -> slot 1-v;
                # the per-channel brightness factors are drawn per OUTPUT slot, after the swap (:157-160)
                slot = (1 - v) if p.flip else v
                d = descs[2 * i + v]
                d.src = 0 if ptrs is None else int(ptrs[i][src_view])
                d.H, d.W, d.row0, d.rows = h, w, row0, rows
                d.x_tab, d.ksx, d.y_tab, d.ksy = x_tab, kx.shape[1], y_tab, ky.shape[1]
                d.flip = int(p.flip)
                d.dst = slot * B + i
                d.lut = l_off
                luts.append(value_table(p.gamma, p.bright, None if p.cbright is None else p.cbright[slot], self.mean))
                l_off += 768
        return dict(descs=descs, tabs=np.concatenate(tabs).astype(np.int32), luts=torch.stack(luts), max_rows=max_rows,
                    n_img=2 * B)

    def __call__(self, lefts, rights, params=None, rng=_random, nprng=np.random):
        """lefts / rights: sequences of uint8 [H,W,3] CUDA tensors (decoded images).  Returns (left, right) fp32
        [B,3,th,tw] and the parameters used.  ``params``: explicit list of AugParams (else drawn per pair)."""
        th, tw = self.size
        B = len(lefts)
        dev = lefts[0].device
        if not lefts[0].is_cuda:
            raise RuntimeError("GpuStereoAugment needs CUDA tensors (no CPU pixel path)")
        for a, b in zip(lefts, rights):
            assert a.dtype == b.dtype == torch.uint8 and a.dim() == 3 and a.shape == b.shape and a.shape[2] == 3
            assert a.is_contiguous() and b.is_contiguous()
        if params is None:
            params = [self.sample(lefts[i].shape[0], lefts[i].shape[1], rng, nprng) for i in range(B)]
        pl = self.plan([(t.shape[0], t.shape[1]) for t in lefts], params,
                       [(lefts[i].data_ptr(), rights[i].data_ptr()) for i in range(B)])
        desc_t = torch.frombuffer(bytearray(bytes(pl["descs"])), dtype=torch.uint8)
        tab_d = torch.from_numpy(pl["tabs"]).pin_memory().to(dev, non_blocking=True)
        lut_d = pl["luts"].reshape(-1).pin_memory().to(dev, non_blocking=True)
        desc_d = desc_t.pin_memory().to(dev, non_blocking=True)
        max_rows = pl["max_rows"]
        stride = max_rows * tw * 3
        inter = torch.empty(2 * B * stride, device=dev, dtype=torch.uint8)
        out = torch.empty(2 * B, 3, th, tw, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().faln_augment_crops_u8(_lib.ptr(desc_d), 2 * B, _lib.ptr(tab_d), _lib.ptr(lut_d), _lib.ptr(inter),
                                                    stride, max_rows, _lib.ptr(out), th, tw, _lib.cur_stream()),
                   "faln_augment_crops_u8")
        return out[:B], out[B:], params
