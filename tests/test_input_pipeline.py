"""Training input pipeline: byte / integer work, bit-exact.

CPU: (1) the oracle's numpy restatement of Pillow's 8-bit bicubic resampler against Pillow itself; (2) the C-ABI host
function `faln_pil_bicubic_coeffs` against the oracle's tables; (3) the oracle's pipeline with explicit parameters against
the golden outputs of the REFERENCE's data_transforms.py classes (tests/golden/pipeline.npz); (4) the product's parameter
sampler draws what the reference draws.  GPU: the two device kernels against the same golden outputs, bit for bit."""
import random

import numpy as np
import pytest
import torch

from oracle import input_pipeline_oracle as IO
from tests.golden.make_golden_pipeline import CASES, source_pair


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/pipeline.npz")


def _params(seed, h, w, th, tw):
    from fal_net_b200 import input_pipeline as IP
    random.seed(seed)
    np.random.seed(seed)
    return IP.sample_params(h, w, (th, tw))


def test_oracle_resampler_is_pillow_bit_for_bit():
    from PIL import Image
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (97, 211, 3), dtype=np.uint8)
    for ow, oh in [(150, 70), (211, 97), (300, 140), (180, 97), (211, 60), (106, 49), (317, 146), (1, 1), (2, 3)]:
        want = np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(IO.pil_resize_bicubic(img, ow, oh), want), (ow, oh)


def test_host_coefficient_tables_match_oracle():
    from fal_net_b200 import input_pipeline as IP
    for i, o in [(211, 150), (211, 300), (97, 97), (1242, 830), (375, 560), (1242, 1863), (5, 2), (3, 9)]:
        b, k = IP.pil_bicubic_tables(i, o, 0, o)
        b2, k2 = IO.pil_coeffs(i, o)
        assert np.array_equal(b, b2) and np.array_equal(k, k2), (i, o)
        lo, n = o // 3, max(1, o // 2)
        b3, k3 = IP.pil_bicubic_tables(i, o, lo, n)
        assert np.array_equal(b3, b2[lo:lo + n]) and np.array_equal(k3, k2[lo:lo + n])


def test_oracle_pipeline_matches_reference_golden(gold):
    assert np.array_equal(gold["cases"], np.array(CASES))
    seen = set()
    for seed, h, w, th, tw in CASES[:12]:
        left, right = source_pair(seed, h, w)
        p = _params(seed, h, w, th, tw)
        rec = gold[f"c{seed}_params"]
        assert (p.factor, p.x1, p.y1, float(p.flip)) == tuple(rec[:4])           # same draws as the reference's classes
        seen.add((p.flip, p.gamma is not None, p.bright is not None, p.cbright is not None))
        o = IO.augment_pair(left, right, p.factor, p.x1, p.y1, p.flip, p.gamma, p.bright, p.cbright, (th, tw))
        assert np.array_equal(o[0].numpy(), gold[f"c{seed}_left"]) and np.array_equal(o[1].numpy(), gold[f"c{seed}_right"])
    assert len(seen) >= 6                                                         # the cases do cover the branches


def test_value_table_is_the_reference_value_chain():
    """3 x 256 table == the chain applied to an image holding every uint8 value, incl. the in-place uint8 wrap of
    RandomCBrightness when nothing promoted the array to float first (reference behaviour, kept)."""
    from fal_net_b200 import input_pipeline as IP
    img = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    for g, b, cb in [(None, None, None), (0.9, None, None), (None, 1.7, None), (None, None, [1.19, 0.85, 1.0]),
                     (1.1, 0.6, [0.8, 1.2, 1.05])]:
        want = IO.augment_pair(img, img, 1.0, 0, 0, False, g, b, None if cb is None else [cb, cb], (1, 256))[0]
        got = IP.value_table(g, b, cb)
        assert torch.equal(got, want[:, 0, :]), (g, b, cb)


def _emulate_kernels(pl, images, th, tw):
    """numpy walk through what csrc/input_pipe.cu's two kernels do with a plan (descriptor by descriptor)."""
    out = np.zeros((pl["n_img"], 3, th, tw), dtype=np.float32)
    tabs, luts = pl["tabs"].astype(np.int64), pl["luts"].numpy().reshape(-1)
    for k in range(pl["n_img"]):
        d = pl["descs"][k]
        src = images[k].astype(np.int64)
        bx = tabs[d.x_tab:d.x_tab + 2 * tw].reshape(tw, 2)
        kx = tabs[d.x_tab + 2 * tw:d.x_tab + 2 * tw + tw * d.ksx].reshape(tw, d.ksx)
        by = tabs[d.y_tab:d.y_tab + 2 * th].reshape(th, 2)
        ky = tabs[d.y_tab + 2 * th:d.y_tab + 2 * th + th * d.ksy].reshape(th, d.ksy)
        inter = np.zeros((d.rows, tw, 3), dtype=np.int64)
        for x in range(tw):
            xmin, n = bx[x]
            acc = (src[d.row0:d.row0 + d.rows, xmin:xmin + n, :] * kx[x, :n][None, :, None]).sum(1) + (1 << 21)
            inter[:, x, :] = np.clip(acc >> 22, 0, 255)
        lut = luts[d.lut:d.lut + 768].reshape(3, 256)
        for y in range(th):
            ymin, n = by[y]
            acc = (inter[ymin - d.row0:ymin - d.row0 + n] * ky[y, :n][:, None, None]).sum(0) + (1 << 21)
            v = np.clip(acc >> 22, 0, 255)                                  # [tw,3]
            for c in range(3):
                row = lut[c][v[:, c]]
                out[d.dst, c, y] = row[::-1] if d.flip else row
    return out


def test_host_plan_through_emulated_kernels_matches_reference_golden(gold):
    """The host half of the product (parameter draw, descriptors, coefficient + value tables) driven through a numpy
    emulation of the two device kernels reproduces the reference pipeline bit for bit -- pins the glue on the CPU."""
    from fal_net_b200 import input_pipeline as IP
    group = CASES[:12]
    th, tw = group[0][3], group[0][4]
    aug = IP.GpuStereoAugment((th, tw))
    params, shapes, imgs = [], [], []
    for seed, h, w, _, _ in group:
        l, r = source_pair(seed, h, w)
        p = _params(seed, h, w, th, tw)
        params.append(p)
        shapes.append((h, w))
        imgs += [r, l] if p.swap_lr else [l, r]
    out = _emulate_kernels(aug.plan(shapes, params), imgs, th, tw)
    B = len(group)
    for i, (seed, *_r) in enumerate(group):
        assert np.array_equal(out[i], gold[f"c{seed}_left"]), seed
        assert np.array_equal(out[B + i], gold[f"c{seed}_right"]), seed


@pytest.mark.gpu
def test_device_pipeline_matches_reference_golden_bit_for_bit(gold):
    from fal_net_b200 import input_pipeline as IP
    dev = torch.device("cuda:0")
    # batch of the 12 small cases (one launch pair), then the KITTI-sized one
    for group in (CASES[:12], CASES[12:]):
        th, tw = group[0][3], group[0][4]
        aug = IP.GpuStereoAugment((th, tw))
        lefts, rights, params = [], [], []
        for seed, h, w, _, _ in group:
            l, r = source_pair(seed, h, w)
            lefts.append(torch.from_numpy(l).to(dev))
            rights.append(torch.from_numpy(r).to(dev))
            params.append(_params(seed, h, w, th, tw))
        left, right, _ = aug(lefts, rights, params=params)
        torch.cuda.synchronize()
        for i, (seed, *_rest) in enumerate(group):
            assert np.array_equal(left[i].cpu().numpy(), gold[f"c{seed}_left"]), seed
            assert np.array_equal(right[i].cpu().numpy(), gold[f"c{seed}_right"]), seed


@pytest.mark.gpu
def test_device_pipeline_draws_like_the_reference_when_seeded(gold):
    """Without explicit parameters the augmenter consumes random / numpy.random in the reference's order."""
    from fal_net_b200 import input_pipeline as IP
    dev = torch.device("cuda:0")
    seed, h, w, th, tw = CASES[4]
    l, r = source_pair(seed, h, w)
    random.seed(seed)
    np.random.seed(seed)
    left, right, p = IP.GpuStereoAugment((th, tw))([torch.from_numpy(l).to(dev)], [torch.from_numpy(r).to(dev)])
    assert np.array_equal(left[0].cpu().numpy(), gold[f"c{seed}_left"])
    assert np.array_equal(right[0].cpu().numpy(), gold[f"c{seed}_right"])
