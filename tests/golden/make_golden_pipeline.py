#!/usr/bin/env python
"""pipeline.npz: outputs of the REFERENCE's own input pipeline (data_transforms.py co-transforms as composed at
Train_Stage1_K.py:115-128 + ArrayToTensor / Normalize) on seeded uint8 stereo pairs, with Python's ``random`` and
``numpy.random`` seeded per case; the draws around the co-transform follow Datasets/listdataset_train.py:75-86.
Run through make_golden_r2.py (build container, CPU)."""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = [(s, 120, 400, 64, 192) for s in range(12)] + [(100, 375, 1242, 192, 640)]


def source_pair(seed, h, w):
    r = np.random.RandomState(1000 + seed)
    base = r.randint(0, 256, (h // 4 + 2, w // 4 + 2, 3)).astype(np.float32)
    # smooth-ish content plus noise so that resampling, clamping and the uint8 wrap quirk are all exercised
    up = np.kron(base, np.ones((4, 4, 1), dtype=np.float32))[:h, :w]
    left = np.clip(up + r.randint(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)
    right = np.clip(np.roll(up, 7, axis=1) + r.randint(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)
    return left, right


def run():
    sys.path.insert(0, "/root/reference")
    import data_transforms as DT
    import torchvision.transforms as transforms
    from oracle import input_pipeline_oracle as IO
    sys.path.insert(0, ROOT)
    from fal_net_b200 import input_pipeline as IP
    out = {"cases": np.array(CASES)}
    for seed, h, w, th, tw in CASES:
        co = DT.Compose([DT.RandomResizeCrop((th, tw), down=0.75, up=1.5), DT.RandomHorizontalFlip(),
                         DT.RandomGamma(min=0.8, max=1.2), DT.RandomBrightness(min=0.5, max=2.0),
                         DT.RandomCBrightness(min=0.8, max=1.2)])
        tf = transforms.Compose([DT.ArrayToTensor(), transforms.Normalize(mean=[0, 0, 0], std=[255, 255, 255]),
                                 transforms.Normalize(mean=[0.411, 0.432, 0.45], std=[1, 1, 1])])
        left, right = source_pair(seed, h, w)
        random.seed(seed)
        np.random.seed(seed)
        _ = random.random() < 0.5 or True                      # listdataset_train.py:75 (fix_order=True)
        np.random.uniform(low=-300, high=300)                   # :86 y_pix
        inputs, _ = co([left.copy(), right.copy()], None)
        res = [tf(a) for a in inputs]
        # the same draws through the product's parameter sampler, then the oracle with explicit parameters
        random.seed(seed)
        np.random.seed(seed)
        p = IP.sample_params(h, w, (th, tw))
        o = IO.augment_pair(left, right, p.factor, p.x1, p.y1, p.flip, p.gamma, p.bright, p.cbright, (th, tw))
        for a, b in zip(res, o):
            assert torch.equal(a, b), (seed, float((a - b).abs().max()))
        out[f"c{seed}_left"], out[f"c{seed}_right"] = res[0].numpy(), res[1].numpy()
        out[f"c{seed}_params"] = np.array([p.factor, p.x1, p.y1, float(p.flip), -1 if p.gamma is None else p.gamma,
                                          -1 if p.bright is None else p.bright, 0 if p.cbright is None else 1])
        print("case", seed, "flip", p.flip, "gamma", p.gamma, "bright", p.bright, "cbright", p.cbright is not None)
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)
    print("pipeline.npz written")
