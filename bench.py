#!/usr/bin/env python
"""bench.py -- headline benchmark of the FAL-net hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stage1|stage2|test|med] [--impl reference]

One JSON line on stdout (rank 0).  Headline workload at every N = BASELINE.json configs[1]: a Stage-1 training step
(reconstruction + smoothness loss, Adam included), batch 8 per GPU, 640x192 crops, N = 49, synthetic
KITTI-shaped stereo pairs, random-init weights (models.FAL_netB under torch.manual_seed(0)).
A "step" = forward + losses + backward + gradient all-reduce (N > 1) + Adam on one batch.

The same line carries, under "extras", driver-run sub-lines measured in the same process with the same timing rules:
  * "stage2" (every N): BASELINE configs[2], the Stage-2 step the >= 7x scaling target is stated on -- per-N values
    make the Stage-2 weak-scaling series (the headline stays ONE workload across N so the driver's own
    value_N / (N * value_1) is meaningful);
  * "test" (every N): configs[3], Test_KITTI flip post-processing, image-sharded, no collective;
  * "med" (N = 1): configs[4], the six MED microbench shapes, fwd / fwd+masks / bwd, each with its HBM roofline;
  * "reference_gpu" (N = 1): the UNMODIFIED reference (baseline/_ref) on the same B200 -- Stage-1 step B = 8 and
    inference b1 -- the practical bar;
  * "conv_layers" (N = 1): every distinct FAL_netB layer shape, our kernel vs cuDNN bf16 channels_last (fwd, dgrad, wgrad).
`--impl reference` times the reference's own code (baseline/_ref) on the host cores; `cpu_baseline` is the same thing
on a bounded sample, run in a subprocess.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MEAN = (0.411, 0.432, 0.45)


def smem_port_gbs():
    """Aggregate shared-memory port bandwidth: 128 B/clk per SM x 148 SMs x the measured max SM clock (nominal; the one
    denominator here that is not driver-measured).  tcgen05.mma with both operands in shared memory (SS mode) reads
    128x16 A + Nx16 B bf16 values per M=128 instruction, i.e. (2/N + 1/64) bytes per MAC for an N-wide tile: at N <= 64 this
    port, not the tensor pipe, bounds an implicit-GEMM convolution (ncu: sm__pipe_tc_cycles_active 74 % /
    l1tex__data_pipe_tc_wavefronts_mem_shared 63 % on the 64-channel full-resolution layers)."""
    mhz = 1965.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mhz = float(json.load(f).get("sm_max_mhz", mhz))
    except Exception:
        pass
    return 128.0 * 148 * mhz * 1e6 / 1e9


def _ss_bytes(flops, gemm_n):
    """Minimal shared-memory operand bytes of an SS-mode tcgen05 implicit GEMM with `flops` = 2 x MACs and N-tile gemm_n."""
    n = max(32, min(int(gemm_n), 256))
    return (flops / 2.0) * (2.0 / n + 1.0 / 64.0)


def peaks():
    """(HBM GB/s, sustained bf16 TFLOP/s, source): the driver's measurements on this pool's B200s, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_batch(B, H, W, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    left = torch.rand(B, 3, H, W, generator=g) - mean
    right = torch.rand(B, 3, H, W, generator=g) - mean
    if pin:
        left, right = left.pin_memory(), right.pin_memory()
    return left.to(device), right.to(device)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own code (baseline/_ref) on the host cores; oracle port as fall-back
# --------------------------------------------------------------------------------------------------
def _reference_step_fn(device, workload, B, seed_off=0):
    """Builds the reference model + optimiser exactly as the reference's entry points do (Train_Stage1_K.py:171-181,
    Train_Stage2_K.py:172-190, Test_KITTI.py:123-124,196-203) and returns (step_fn, frames_per_step, description)."""
    from baseline import ref_loader
    cpu = device == "cpu"
    RM, RL = ref_loader.load(cpu=cpu, want_losses=workload != "test")
    N = 49
    dev = torch.device(device)

    def make(seed):
        torch.manual_seed(seed)
        m = RM.FAL_netB(None, no_levels=N)
        return m.to(dev)

    mx = torch.full((B, 1, 1), 300.0, device=dev)
    mn = mx * 2 / 300
    if workload in ("stage1", "stage2"):
        H, W = 192, 640
        model = make(0)
        fix = make(1).eval() if workload == "stage2" else None
        groups = [{"params": model.bias_parameters(), "weight_decay": 0.0},
                  {"params": model.weight_parameters(), "weight_decay": 0.0}]
        opt = torch.optim.Adam(params=groups, lr=1e-4 if workload == "stage1" else 5e-5, betas=(0.5, 0.999))
        left, right = (t.to(dev) for t in synth_batch(B, H, W, 1234 + seed_off))

        def step():
            opt.zero_grad()
            if workload == "stage1":
                loss = ref_loader.ref_stage1(RL, model, left, right, mn, mx, a_p=0.0)[0]
            else:
                loss = ref_loader.ref_stage2(RL, model, fix, left, right, mn, mx, a_p=0.01)["loss"]
            loss.backward()
            opt.step()
            return loss
        return step, B * (2 if workload == "stage2" else 1), f"{workload} step, {B} of 8 {'pairs' if workload == 'stage2' else 'images'}, 192x640, N=49"
    H, W = 375, 1242
    model = make(0).eval()
    img = synth_batch(B, H, W, 1234 + seed_off)[0].to(dev)

    def step():
        with torch.no_grad():                                           # Test_KITTI.py:196-203, exact index flip
            d = model(img, mn, mx, ret_disp=True, ret_subocc=False, ret_pan=False)
            fd = model(torch.flip(img, dims=[3]), mn, mx, ret_disp=True, ret_subocc=False, ret_pan=False)
            return ((d + torch.flip(fd, dims=[3])) / 2).mean()
    return step, B, f"Test_KITTI flip-PP, {B} of 8 images, 375x1242, N=49"


def cpu_reference(workload, steps, warmup, sample_b=None, budget_s=25.0):
    """The reference's own modules (baseline/_ref, unmodified; `.cuda()` is the identity on this CPU leg) on all host
    cores, on a bounded sample of the workload.  Falls back to oracle/falnet_oracle.py (kind "port", bit-identical to
    the reference on CPU) when baseline/_ref did not travel.  Returns (value, unit, cores, kind, sample, ms_per_step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = workload if workload != "med" else "stage1"
    from baseline import ref_loader
    if not ref_loader.available():
        return _cpu_port(wl, steps, warmup, sample_b, budget_s)
    B = sample_b or 1
    step, per_step, desc = _reference_step_fn("cpu", wl, B)
    t0 = time.time()
    float(step())
    first = time.time() - t0
    if sample_b is None and wl != "test" and first * 8 * 3 < budget_s:   # the whole batch of 8 fits the budget
        B = 8
        step, per_step, desc = _reference_step_fn("cpu", wl, B)
        t0 = time.time()
        float(step())
        first = time.time() - t0
    n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.time()
    for _ in range(n):
        float(step())
    dt = (time.time() - t0) / n
    return per_step / dt, "frames/s", cores, "reference", desc + f", unmodified reference (fp32 torch CPU), {n} timed step(s)", dt * 1e3


def _cpu_port(workload, steps, warmup, sample_b=None, budget_s=25.0):
    from oracle import falnet_oracle as O
    cores = os.cpu_count() or 1
    N = 49
    B = sample_b or 1
    if workload in ("stage1", "stage2"):
        H, W = 192, 640
        p = {k: v.clone().requires_grad_("amask" not in k) for k, v in O.init_params(N, seed=0).items()}
        pfix = {k: v.clone() for k, v in O.init_params(N, seed=1).items()}
        vgg_ws = O.init_vgg(2)
        left, right = synth_batch(B, H, W, 1234)
        mx = torch.full((B, 1, 1), 300.0)
        mn = mx * 2 / 300
        m = {k: torch.zeros_like(v) for k, v in p.items()}
        v = {k: torch.zeros_like(t) for k, t in p.items()}
        tstep = [0]

        def step():
            for t in p.values():
                t.grad = None
            if workload == "stage1":
                loss = O.stage1_loss(p, left, right, mn, mx, a_p=0.0)[0]
            else:
                loss = O.stage2_loss(p, pfix, left, right, mn, mx, a_p=0.01, vgg_ws=vgg_ws, flip=lambda t: torch.flip(t, dims=[3]))["loss"]
            loss.backward()
            tstep[0] += 1
            with torch.no_grad():
                O.adam_step({k: t for k, t in p.items()}, {k: t.grad for k, t in p.items()}, m, v, tstep[0], 1e-4)
            return float(loss)
        per_step = B * (1 if workload == "stage1" else 2)
        sample = f"{workload} step on {B} of 8 {'pairs' if workload == 'stage2' else 'images'}, 192x640, N=49, oracle port (fp32 torch CPU)"
    else:
        H, W = 375, 1242
        p = O.init_params(N, seed=0)
        img = synth_batch(B, H, W, 1234)[0]
        mx = torch.full((B, 1, 1), 300.0)
        mn = mx * 2 / 300

        def step():
            with torch.no_grad():
                d = O.test_disp_fpp(p, img, mn, mx, flip=lambda t: torch.flip(t, dims=[3]))
            return float(d.mean())
        per_step = B
        sample = f"Test_KITTI flip-PP on {B} of 8 images, 375x1242, N=49, oracle port (fp32 torch CPU)"
    t0 = time.time()
    step()
    first = time.time() - t0
    n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.time()
    for _ in range(n):
        step()
    dt = (time.time() - t0) / n
    return per_step / dt, "frames/s", cores, "port", sample + f", {n} timed step(s)", dt * 1e3


def reference_gpu(steps=10, warmup=3):
    """The UNMODIFIED reference on the same B200 (its stock path: fp32 tensors, cuDNN with torch's default TF32 convolution
    setting, eager launches): Stage-1 step B = 8 at 192x640 and disparity inference b1 at 375x1242 (SURVEY.md 8d "the
    practical bar").  CUDA-event timing after warm-up."""
    out = {"impl": "reference_gpu", "torch": torch.__version__, "cudnn_conv_tf32": bool(torch.backends.cudnn.allow_tf32),
           "note": "stock reference settings; not a roofline claim"}
    for name, wl, B in (("stage1_b8_192x640", "stage1", 8), ("stage2_8pairs_192x640", "stage2", 8),
                        ("test_fpp_b1_375x1242", "test", 1)):
        try:
            step, per_step, desc = _reference_step_fn("cuda", wl, B)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "frames_per_s": per_step / (ms / 1e3), "what": desc,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            del step
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
        except Exception as e:                                   # report, do not hide
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def _dev_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def next_rows_bench(dev, models, steps, LF, FlatAdamDDP, GraphedStep):
    """SURVEY.md 8(f) rows, measured with the same rules (CUDA events, warm-up, inputs > L2 or rotating):
    f1 Test_KITTI multi-scale post-processing, f2 validate() metrics, f3 input pipeline, f4 FAL_netA / FAL_netC /
    Train_Stage1_Kslow steps.  Where the reference's own module for the row imports here (baseline/_ref myUtils.py,
    data_transforms.py: host code by construction) it is timed beside ours on one host core, as its DataLoader worker /
    validate() loop would run it."""
    import importlib.util
    import random as _random
    import time as _time
    import numpy as np
    from fal_net_b200 import input_pipeline as IP, myUtils as MU, postproc
    out = {}
    N = 49
    mx8 = torch.full((8, 1, 1), 300.0, device=dev)
    mn8 = mx8 * 2 / 300

    def ref_module(name):
        p = os.path.join(ROOT, "baseline", "_ref", name + ".py")
        if not os.path.exists(p):
            return None
        spec = importlib.util.spec_from_file_location("falnet_ref_" + name, p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    # ---- f1: multi-scale post-processing (Test_KITTI.py:287-300), 8 images at 375x1242 ----
    try:
        torch.manual_seed(0)
        model = models.FAL_netB(no_levels=N).to(dev).eval()
        imgs = [synth_batch(8, 375, 1242, 77 + i)[0].to(dev) for i in range(3)]
        it = [0]

        def f_mspp():
            it[0] += 1
            return steps.test_disp(model, imgs[it[0] % 3], mn8, mx8, ms_post_process=True)

        def f_plain():
            it[0] += 1
            return steps.test_disp(model, imgs[it[0] % 3], mn8, mx8)
        ms_m, ms_p = _dev_ms(f_mspp), _dev_ms(f_plain)
        disp = f_plain()
        small = postproc.flip_resize_bilinear(imgs[0], scale_factor=2 / 3, flip_x=True)
        d2 = model(small, mn8, mx8, ret_disp=True, ret_pan=False, ret_subocc=False)
        p95 = postproc.percentile_rows(disp, 95.0, add=1e-6)
        out["f1_ms_pp"] = {"ms_per_8_images": ms_m, "frames_per_s": 8 / (ms_m / 1e3), "plain_pass_ms": ms_p,
                           "kernels_us": {"flip_resize_bilinear": 1e3 * _dev_ms(lambda: postproc.flip_resize_bilinear(imgs[0], scale_factor=2 / 3, flip_x=True)),
                                          "percentile_rows": 1e3 * _dev_ms(lambda: postproc.percentile_rows(disp, 95.0, add=1e-6)),
                                          "mspp_blend": 1e3 * _dev_ms(lambda: postproc.mspp_blend(disp, d2, p95, 1.5))},
                           "what": "disparity pass + 2/3-scale flipped pass + per-image 95th percentile + blend, B=8 375x1242"}
        del model, imgs, disp, small, d2
    except Exception as e:
        out["f1_ms_pp"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()

    # ---- f2: validate() metrics (Train_Stage1_K.py:279-347), one batch of 8 at 375x1242 ----
    try:
        g = torch.Generator(device=dev).manual_seed(5)
        disp = 2 + 60 * torch.rand(8, 1, 375, 1242, device=dev, generator=g)
        tgt = (2 + 60 * torch.rand(8, 1, 375, 1242, device=dev, generator=g)) * (torch.rand(8, 1, 375, 1242, device=dev, generator=g) < 0.25)
        a_ = torch.rand(8, 3, 375, 1242, device=dev, generator=g) - 0.4
        b_ = torch.rand(8, 3, 375, 1242, device=dev, generator=g) - 0.4

        def f_ours():
            return MU.kitti_errors_batch(tgt, disp, "Kitti2015"), LF.realEPE(disp, tgt, sparse=True), MU.get_rmse(a_, b_)
        ms_o = _dev_ms(f_ours)
        row = {"ours_ms_per_batch8": ms_o, "what": "7 KITTI errors + EPE + RMSE for 8 images 375x1242, device-resident, no host sync"}
        RU = ref_module("myUtils")
        if RU is not None:
            def f_ref():
                td, pd_ = tgt.squeeze(1).cpu().numpy(), disp.squeeze(1).cpu().numpy()
                gt_d, pr_d = RU.disps_to_depths_kitti2015(td, pd_)
                return [RU.compute_kitti_errors(gt_d[i], pr_d[i]) for i in range(len(gt_d))], RU.get_rmse(a_, b_)
            f_ref()
            torch.cuda.synchronize()
            t0 = _time.perf_counter()
            for _ in range(3):
                f_ref()
            torch.cuda.synchronize()
            row["reference_ms_per_batch8"] = (_time.perf_counter() - t0) / 3 * 1e3
            row["reference_what"] = "reference myUtils.py as validate() calls it (.cpu().numpy() + numpy per image), wall clock"
        out["f2_metrics"] = row
        del disp, tgt, a_, b_
    except Exception as e:
        out["f2_metrics"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()

    # ---- f3: input pipeline (data_transforms.py + Train_Stage1_K.py:115-128), 8 pairs 375x1242 -> 192x640 ----
    try:
        r = np.random.RandomState(3)
        raw = [(r.randint(0, 256, (375, 1242, 3)).astype(np.uint8), r.randint(0, 256, (375, 1242, 3)).astype(np.uint8))
               for _ in range(8)]
        lefts = [torch.from_numpy(a).to(dev) for a, _ in raw]
        rights = [torch.from_numpy(b).to(dev) for _, b in raw]
        aug = IP.GpuStereoAugment((192, 640))
        _random.seed(0)
        np.random.seed(0)
        params = [aug.sample(375, 1242) for _ in range(8)]
        ms_dev = _dev_ms(lambda: aug(lefts, rights, params=params))
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(5):
            aug(lefts, rights, params=params)
        torch.cuda.synchronize()
        ms_wall = (_time.perf_counter() - t0) / 5 * 1e3
        row = {"ours_ms_per_8_pairs_device_stream": ms_dev, "ours_ms_per_8_pairs_wall_incl_host_tables": ms_wall,
               "pairs_per_s_wall": 8 / (ms_wall / 1e3),
               "what": "decoded uint8 pairs resident on the device -> normalised fp32 crops; host half = Pillow-exact coefficient / value tables"}
        DT = ref_module("data_transforms")
        if DT is not None:
            import torchvision.transforms as T
            co = DT.Compose([DT.RandomResizeCrop((192, 640), down=0.75, up=1.5), DT.RandomHorizontalFlip(),
                             DT.RandomGamma(min=0.8, max=1.2), DT.RandomBrightness(min=0.5, max=2.0),
                             DT.RandomCBrightness(min=0.8, max=1.2)])
            tf = T.Compose([DT.ArrayToTensor(), T.Normalize(mean=[0, 0, 0], std=[255, 255, 255]),
                            T.Normalize(mean=[0.411, 0.432, 0.45], std=[1, 1, 1])])
            _random.seed(0)
            np.random.seed(0)
            t0 = _time.perf_counter()
            for a, b in raw:
                inputs, _ = co([a.copy(), b.copy()], None)
                _ = [tf(x) for x in inputs]
            ms_ref = (_time.perf_counter() - t0) * 1e3
            row["reference_ms_per_8_pairs_one_worker"] = ms_ref
            row["reference_pairs_per_s_4_workers"] = 4 * 8 / (ms_ref / 1e3)
            row["reference_what"] = "reference co-transforms + input transform on one host core (its DataLoader runs 4 such workers)"
        out["f3_input_pipeline"] = row
        del lefts, rights
    except Exception as e:
        out["f3_input_pipeline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()

    # ---- f4: FAL_netA / FAL_netC Stage-1 steps and the Train_Stage1_Kslow step (B = 8, 192x640, N = 49, graph replay) ----
    left, right = [t.to(dev) for t in synth_batch(8, 192, 640, 99)]
    for name, ctor, slow in (("FAL_netA", "FAL_netA", False), ("FAL_netC", "FAL_netC", False), ("Kslow_FAL_netB", "FAL_netB", True)):
        try:
            torch.manual_seed(0)
            model = getattr(models, ctor)(no_levels=N).to(dev)
            opt = FlatAdamDDP(model, lr=1e-4)
            vgg = LF.vgg if slow else None

            def loss_fn(l, r_, model=model, slow=slow, vgg=vgg):
                if slow:
                    return steps.stage1_slow_loss(model, l, r_, mn8, mx8, a_p=0.01, vgg=vgg)["loss"]
                return steps.stage1_loss(model, l, r_, mn8, mx8, a_p=0.0)[0]
            gs = GraphedStep(opt, loss_fn, left, right, warmup=3)
            ms = _dev_ms(lambda: gs.run(left, right), iters=20)
            out["f4_" + name] = {"ms_per_step": ms, "frames_per_s": 8 / (ms / 1e3),
                                 "params_m": sum(p.numel() for p in model.parameters()) / 1e6}
            del gs, opt, model
        except Exception as e:
            out["f4_" + name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    return out


def med_microbench(hbm_peak, peak_src, iters=10):
    """BASELINE configs[4]: the six MED microbench shapes (N = 33 / 49 / 65 at 8x375x1242 and 2x1024x2048), fwd (pan + disp),
    fwd + both occlusion masks, bwd; GB/s = algorithmic bytes of SURVEY.md 8(d) / CUDA-event time; three rotating buffer sets
    of >= 0.45 GB each so the 126 MB L2 cannot serve repeats.  The kernels are the ones the training step launches."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_med
    rows = []
    for B, N, H, W in ((8, 33, 375, 1242), (8, 49, 375, 1242), (8, 65, 375, 1242),
                       (2, 33, 1024, 2048), (2, 49, 1024, 2048), (2, 65, 1024, 2048), (16, 49, 192, 640)):
        r = bench_med.bench(B, N, H, W, iters=iters, peak=hbm_peak)
        torch.cuda.empty_cache()
        row = {"shape": f"{B}x{N}x{H}x{W}"}
        for k, v in r.items():
            row[k] = {"ms": round(v["ms"], 4), "roofline": {"bound": "hbm", "achieved": round(v["gbs"], 1), "peak": hbm_peak,
                                                           "unit": "GB/s", "frac": round(v["gbs"] / hbm_peak, 4)}}
        rows.append(row)
    return {"peak_source": peak_src + " hbm_gbs", "algorithmic_bytes_per_px": "fwd 4(N+7), fwd_masks 4(N+9), bwd 4(2N+7), disp_only 4(N+1)",
            "l2": "3 rotating buffer sets per shape", "rows": rows}


# (name, cin, cout, H_in, W_in, stride) of every distinct 3x3 layer shape of FAL_netB at 192x640
_LAYERS = (("conv0_1.*", 32, 32, 192, 640, 1), ("conv1.0", 32, 64, 192, 640, 2), ("conv1_1.*", 64, 64, 96, 320, 1),
           ("conv2.0", 64, 128, 96, 320, 2), ("conv2_1.*", 128, 128, 48, 160, 1), ("conv3.0", 128, 256, 48, 160, 2),
           ("conv3_1.*", 256, 256, 24, 80, 1), ("conv4.0", 256, 256, 24, 80, 2), ("conv4_1.*", 256, 256, 12, 40, 1),
           ("conv5.0", 256, 256, 12, 40, 2), ("conv5_1.*", 256, 256, 6, 20, 1), ("conv6.0", 256, 512, 6, 20, 2),
           ("conv6_1.*", 512, 512, 3, 10, 1), ("deconv6", 512, 256, 6, 20, 1), ("iconv6", 512, 256, 6, 20, 1),
           ("deconv5", 256, 128, 12, 40, 1), ("iconv5", 384, 256, 12, 40, 1), ("deconv4", 256, 128, 24, 80, 1),
           ("iconv4", 384, 256, 24, 80, 1), ("deconv3", 256, 128, 48, 160, 1), ("iconv3", 256, 128, 48, 160, 1),
           ("deconv2", 128, 64, 96, 320, 1), ("iconv2", 128, 64, 96, 320, 1), ("deconv1", 64, 64, 192, 640, 1),
           ("iconv1+conv0 (folded)", 96, 64, 192, 640, 1))


def conv_layer_table(tf_peak, hbm_peak, B=8, iters=8):
    """Per-layer microseconds of our tcgen05 kernels against cuDNN (bf16, channels_last, cudnn.benchmark autotuned) for the
    forward, data-gradient and weight-gradient of every distinct FAL_netB layer shape at the Stage-1 batch (B = 8, 192x640).
    Each timed alone as a CUDA-graph replay of 8 back-to-back launches after 3 warm-ups (no host launch overhead on either
    side); inputs of the big layers exceed L2 only at full resolution -- the small-map layers are L2-resident for BOTH
    implementations.  `bound_us` = the layer-wise roofline
    max(flops / tensor peak, bytes / HBM peak); `bound_ss_us` adds the shared-memory operand port of SS-mode tcgen05.mma
    (smem_port_gbs) as a third bound."""
    import torch.nn.functional as F
    from fal_net_b200 import conv_native as CN
    dev = torch.device("cuda", torch.cuda.current_device())
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    CL = torch.channels_last
    g = torch.Generator(device=dev).manual_seed(11)

    def timeit(fn):
        """us per launch of a CUDA-graph replay of `iters` back-to-back launches (no host launch overhead on either side:
        an eager loop is host-bound below ~25 us per call); eager fallback if the capture fails."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(iters):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0.record()
            gr.replay()
            e1.record()
        except Exception:
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    rows, tot = [], {"ours": 0.0, "cudnn": 0.0, "bound": 0.0}
    for name, cin, cout, H, W, stride in _LAYERS:
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        x = torch.randn(B, cin, H, W, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
        w = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / (3 * cin ** 0.5))
        gy = torch.randn(B, cout, Ho, Wo, device=dev, generator=g).to(torch.bfloat16).contiguous(memory_format=CL)
        w16 = w.to(torch.bfloat16).contiguous(memory_format=CL)
        wk, wd = CN.pack_weight(w), CN.pack_weight_dgrad(w)
        dW = torch.zeros(cout, 3, 3, cin, device=dev).permute(0, 3, 1, 2)
        flops = 2 * 9 * cin * cout * B * Ho * Wo
        nbytes = 2 * B * (H * W * cin + Ho * Wo * cout) + 2 * 9 * cin * cout
        bound = max(flops / (tf_peak * 1e12), nbytes / (hbm_peak * 1e9)) * 1e6
        bound_ss = max(bound, _ss_bytes(flops, cout) / (smem_port_gbs() * 1e9) * 1e6)
        row = {"layer": name, "cin": cin, "cout": cout, "hw_in": [H, W], "stride": stride, "gflop": round(flops / 1e9, 3),
               "bound_us": round(bound, 2), "bound_ss_us": round(bound_ss, 2)}
        # a 96-channel input is two sources in the network (64-channel deconv output + 32-channel skip): the data and
        # weight gradients run per source, exactly as fal_net_b200.backbone schedules them
        parts = [(0, cin)] if (cin == 32 or cin % 64 == 0) else [(0, cin // 64 * 64), (cin // 64 * 64, cin % 64)]
        xs = [x[:, o:o + c].contiguous(memory_format=CL) for o, c in parts] if len(parts) > 1 else [x]

        def ours_dgrad():
            for o, c in parts:
                CN.conv3x3_dgrad(gy, wd, (H, W), stride=stride, rows=(o, c))

        def ours_wgrad():
            for (o, c), xp in zip(parts, xs):
                CN.conv3x3_wgrad(gy, xp, dW, cout=cout, cx=c, ci_off=o, stride=stride)
        ops = {
            "fwd": (lambda: CN.conv3x3_fwd(x, wk, None, stride, 1),
                    lambda: F.conv2d(x, w16, None, stride, 1)),
            "dgrad": (ours_dgrad,
                      lambda: torch.ops.aten.convolution_backward(gy, x, w16, None, [stride, stride], [1, 1], [1, 1], False,
                                                                  [0, 0], 1, [True, False, False])),
            "wgrad": (ours_wgrad,
                      lambda: torch.ops.aten.convolution_backward(gy, x, w16, None, [stride, stride], [1, 1], [1, 1], False,
                                                                  [0, 0], 1, [False, True, False])),
        }
        for op, (ours, lib) in ops.items():
            try:
                t_o = timeit(ours)
            except Exception as e:
                t_o = None
                row[op + "_error"] = f"{type(e).__name__}: {e}"[:160]
            t_c = timeit(lib)
            row[op] = {"ours_us": None if t_o is None else round(t_o, 1), "cudnn_bf16_us": round(t_c, 1),
                       "ours_frac_of_bound": None if t_o is None else round(bound / t_o, 3)}
            if t_o is not None:
                tot["ours"] += t_o
                tot["cudnn"] += t_c
                tot["bound"] += bound
        if name.startswith("deconv") and H % 2 == 0 and W % 2 == 0 and stride == 1:
            # the reference's deconv block = nearest 2x up-sampling + this conv: ours runs it as ONE folded kernel on the
            # low-resolution input (and one folded data-gradient kernel); the library needs interpolate + conv (+ its backward)
            xl = x[:, :, ::2, ::2].contiguous(memory_format=CL)
            wf_up, wd_up = CN.pack_up2_weights(w)
            xl32 = xl.float().contiguous(memory_format=CL)
            t_o = timeit(lambda: CN.conv3x3_up2_fwd(xl, wf_up, None, 1))
            t_c = timeit(lambda: F.elu(F.conv2d(F.interpolate(xl, scale_factor=2, mode="nearest"), w16, None, 1, 1)))
            t_od = timeit(lambda: CN.conv3x3_up2_dgrad(gy, wd_up))
            row["block_fwd"] = {"ours_folded_us": round(t_o, 1), "cudnn_interpolate_conv_elu_us": round(t_c, 1)}
            row["block_dgrad"] = {"ours_folded_us": round(t_od, 1)}
            if cin % 64 == 0 and cout % 64 == 0:
                # weight gradient of the block from the low-resolution input (conv3x3_wgrad_up2_kernel) against ours on the
                # explicitly up-sampled map (the path before that kernel) and the library's interpolate + wgrad
                t_ow = timeit(lambda: CN.conv3x3_wgrad_up2(gy, xl, dW, cout=cout, cx=cin))
                t_pw = timeit(lambda: CN.conv3x3_wgrad(gy, CN.upsample_nearest(xl, (H, W)), dW, cout=cout, cx=cin))
                t_cw = timeit(lambda: torch.ops.aten.convolution_backward(
                    gy, F.interpolate(xl, scale_factor=2, mode="nearest"), w16, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                    [False, True, False]))
                row["block_wgrad"] = {"ours_folded_us": round(t_ow, 1), "ours_upsample_plus_wgrad_us": round(t_pw, 1),
                                      "cudnn_interpolate_wgrad_us": round(t_cw, 1)}
            del xl, xl32, wf_up, wd_up
        rows.append(row)
        del x, w, gy, w16, wk, wd, dW
    torch.backends.cudnn.benchmark = old_bench
    return {"batch": B, "what": "us per launch (CUDA-graph replay), each kernel timed alone; cuDNN = torch conv2d / convolution_backward, bf16 "
            "channels_last, cudnn.benchmark=True", "sum_ours_us": round(tot["ours"], 1), "sum_cudnn_bf16_us": round(tot["cudnn"], 1),
            "sum_bound_us": round(tot["bound"], 1), "rows": rows}


def _subprocess_json(argv, timeout):
    """Run `python bench.py <argv>` and return its last JSON line (legs that must not share this process's state)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=timeout)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": f"rc={r.returncode}: {r.stderr[-300:]}"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="stage1", choices=["stage1", "stage2", "test", "med"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_gpu"])
    ap.add_argument("--budget", type=float, default=90.0, help="--impl reference: seconds of timed CPU work")
    ap.add_argument("--sample-b", type=int, default=None, help="--impl reference: images (pairs) per step of the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="stage2,test,med,reference_gpu,conv_layers,next_rows",
                    help="comma list of the sub-lines to attach (med / reference_gpu / conv_layers only at N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="launch the training step kernel by kernel instead of replaying its CUDA graph")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        val, unit, cores, kind, sample, ms = cpu_reference(a.workload, a.steps, a.warmup, sample_b=a.sample_b,
                                                           budget_s=a.budget)
        print(json.dumps({
            "impl": "reference", "metric": metric_name(a.workload), "value": val, "unit": unit, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a.workload, a.gpus),
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0
    if a.impl == "reference_gpu":
        if rank == 0:
            torch.cuda.set_device(0)
            print(json.dumps(reference_gpu(steps=min(a.steps, 10), warmup=3)))
        return 0

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from fal_net_b200 import _lib, med, models, steps
    from fal_net_b200 import loss_functions as LF
    from fal_net_b200.trainer import FlatAdamDDP, GraphedStep
    from fal_net_b200 import conv as C
    run_gpu.GraphedStep = GraphedStep
    ctx = (rank, world, dev, dist, _lib, med, models, steps, LF, FlatAdamDDP, C)

    result = run_gpu(a, a.workload, *ctx)
    wanted = [] if a.no_extras else [e for e in a.extras.split(",") if e]
    extras = {}
    for wl in ("stage2", "test"):
        if wl in wanted and wl != a.workload:
            torch.cuda.empty_cache()
            r = run_gpu(a, wl, *ctx, light=True)
            extras[wl] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "cuda_graph",
                                            "config", "roofline", "roofline_med", "comm") if k in r}
            if wl == "stage2":
                extras[wl]["pairs_per_s"] = r["value"] / 2
    if world == 1 and rank == 0:
        torch.cuda.empty_cache()
        hbm_peak, tf_peak, peak_src = peaks()
        if "med" in wanted:
            extras["med"] = med_microbench(hbm_peak, peak_src)
        if "conv_layers" in wanted:
            extras["conv_layers"] = conv_layer_table(tf_peak, hbm_peak)
        if "next_rows" in wanted:
            extras["next_rows"] = next_rows_bench(dev, models, steps, LF, FlatAdamDDP, GraphedStep)
        torch.cuda.empty_cache()
        if "reference_gpu" in wanted:
            extras["reference_gpu"] = _subprocess_json(["--impl", "reference_gpu", "--steps", "10"], timeout=240)
            try:
                extras["reference_gpu"]["speedup_stage1_device"] = \
                    result["value"] / extras["reference_gpu"]["stage1_b8_192x640"]["frames_per_s"]
            except Exception:
                pass
        if not a.no_cpu_baseline:
            cb = _subprocess_json(["--impl", "reference", "--workload", a.workload, "--steps", "3", "--budget", "20"], timeout=300)
            result["cpu_baseline"] = cb.get("cpu_baseline", cb)
            if "ms_per_step" in cb:
                result["cpu_baseline"]["ms_per_step"] = cb["ms_per_step"]
    if extras:
        result["extras"] = extras
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))
    return 0


def metric_name(workload):
    return {"stage1": "train frames/s (Stage-1 step)", "stage2": "train frames/s (Stage-2 step)",
            "test": "inference frames/s (Test_KITTI flip-PP)", "med": "MED synthesis HBM GB/s"}[workload]


def workload_config(workload, n):
    base = {"stage1": {"workload": "Stage-1 training step (L1 reconstruction + smoothness, a_p=0, Adam), batch 8/GPU, 640x192, N=49 (BASELINE configs[1])",
                       "per_gpu_batch": 8, "H": 192, "W": 640, "N": 49},
            "stage2": {"workload": "Stage-2 training step (mirrored occlusion masks + perceptual + mirror loss, Adam), 8 pairs/GPU, 640x192, N=49 (BASELINE configs[2])",
                       "per_gpu_batch": 8, "H": 192, "W": 640, "N": 49},
            "test": {"workload": "Test_KITTI inference with flip post-processing, 8 images/GPU, 375x1242, N=49 (BASELINE configs[3])",
                     "per_gpu_batch": 8, "H": 375, "W": 1242, "N": 49},
            "med": {"workload": "MED synthesis+occlusion kernel microbench fwd+bwd, B=8, 1242x375, N=49 (BASELINE configs[4])",
                    "per_gpu_batch": 8, "H": 375, "W": 1242, "N": 49}}[workload]
    base["parallelism"] = f"dp{n}"
    base["l2"] = "inputs rotate over 3 batches and each step streams >1 GB of activations/logits (>> 126 MB L2)"
    return base


def run_gpu(a, wl, rank, world, dev, dist, _lib, med, models, steps, LF, FlatAdamDDP, C, light=False):
    """Builds workload ``wl`` and measures it (device-resident value, per-kernel rooflines, e2e).  ``light``: an extras
    sub-line -- same timing rules, shorter per-kernel pass."""
    N = 49
    cfg = workload_config(wl, world)
    B, H, W = cfg["per_gpu_batch"], cfg["H"], cfg["W"]
    hbm_peak, tf_peak, peak_src = peaks()

    torch.manual_seed(0)
    model = models.FAL_netB(no_levels=N).to(dev)
    fix_model = None
    if wl == "stage2":
        torch.manual_seed(1)
        fix_model = models.FAL_netB(no_levels=N).to(dev).eval()
        for p_ in fix_model.parameters():
            p_.requires_grad_(False)
    opt = None
    if wl in ("stage1", "stage2"):
        opt = FlatAdamDDP(model, lr=1e-4 if wl == "stage1" else 5e-5)
        opt.broadcast_parameters()
    vgg = LF.vgg if wl == "stage2" else None

    # host (pinned) and device-resident batches; rank-offset seeds
    nb = 3
    host = [synth_batch(B, H, W, 1234 + 17 * rank + i, pin=True) for i in range(nb)]
    devb = [(l.to(dev), r.to(dev)) for l, r in host]
    mx = torch.full((B, 1, 1), 300.0, device=dev)
    mn = mx * 2 / 300

    def step_dev(left, right):
        if wl == "stage1":
            opt.zero_grad()
            loss = steps.stage1_loss(model, left, right, mn, mx, a_p=0.0)[0]
            loss.backward()
            opt.step()
            return loss
        if wl == "stage2":
            opt.zero_grad()
            loss = steps.stage2_loss(model, fix_model, left, right, mn, mx, a_p=0.01, vgg=vgg)["loss"]
            loss.backward()
            opt.step()
            return loss
        if wl == "test":
            return steps.test_disp(model, left, mn, mx, f_post_process=True).mean()
        # med microbench: fused fwd (with masks) + bwd on resident logits
        r = med.med_forward_raw(med_logits[0], left, med_tabs[1], med_tabs[0], med_tabs[2], True, True, True)
        g = med.med_backward_raw(med_logits[0], left, med_tabs[1], med_tabs[0], med_tabs[2], r["pan"], r["disp"], r["lse0"],
                                 r["lsew"], med_gp, med_gd, out=med_gl)
        return g[0, 0, 0, 0]

    if wl == "med":
        gen = torch.Generator(device=dev).manual_seed(7)
        med_logits = [2 * torch.randn(B, N, H, W, generator=gen, device=dev)]
        d_, xo_ = med.level_tables(mn, mx, N, W)
        med_tabs = (d_, xo_, med.grid_row(W, dev))
        med_gp = torch.randn(B, 3, H, W, generator=gen, device=dev)
        med_gd = torch.randn(B, 1, H, W, generator=gen, device=dev)
        med_gl = torch.empty_like(med_logits[0])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames_per_step = B * world * (2 if wl == "stage2" else 1)

    # training steps replay ONE captured CUDA graph (zero_grad + forward + losses + backward + all-reduce + Adam)
    graphed = None
    launches_per_step = lib_convs_per_step = None
    if wl in ("stage1", "stage2") and not a.no_graph:
        def loss_fn(left, right):
            if wl == "stage1":
                return steps.stage1_loss(model, left, right, mn, mx, a_p=0.0)[0]
            return steps.stage2_loss(model, fix_model, left, right, mn, mx, a_p=0.01, vgg=vgg)["loss"]
        graphed = run_gpu.GraphedStep(opt, loss_fn, devb[0][0], devb[0][1], warmup=3)   # 3 eager steps, then capture
        l0, c0 = _lib.launch_count(), C.LIBRARY_CALLS["conv_backward"]
        step_dev(*devb[0])                                                              # one eager step: count launches
        launches_per_step, lib_convs_per_step = _lib.launch_count() - l0, C.LIBRARY_CALLS["conv_backward"] - c0

    def run_step(left, right):
        return graphed.run(left, right) if graphed is not None else step_dev(left, right)

    # ---------------- device-resident timing (value) ----------------
    for i in range(a.warmup):
        run_step(*devb[i % nb])
    sync_all()
    lib_conv0 = C.LIBRARY_CALLS["conv_backward"]
    launches0 = _lib.launch_count()
    clocks = ClockSampler(dev.index or 0)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(a.steps):
        run_step(*devb[i % nb])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / a.steps
    clk = clocks.stop() if rank == 0 else None
    launches = (_lib.launch_count() - launches0) if graphed is None else launches_per_step * a.steps
    lib_convs = (C.LIBRARY_CALLS["conv_backward"] - lib_conv0) if graphed is None else lib_convs_per_step * a.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = frames_per_step / (ms / 1e3)

    # per-launch CUDA-event timing of our kernels (event records cannot live inside a captured graph, so the same steps
    # are run kernel by kernel once more, on ONE stream so that every kernel is timed alone; these iterations are not
    # part of `value`)
    from fal_net_b200 import backbone as BB, conv_native as CNV
    n_prof = 3 if light else max(3, min(a.steps, 10))
    BB.USE_SIDE_STREAM = False
    step_dev(*devb[0])                                               # settle allocator / caches in this mode
    med.TIMING, CNV.TIMING = [], []
    for i in range(n_prof):
        # keep the host ahead of the device: an event pair brackets its kernel in STREAM order, so if the (eager) host launch
        # arrives late the pair would also time the idle gap before it.  ~25 ms of spin first lets the host queue a whole
        # step's launches before the first timed kernel starts.
        torch.cuda._sleep(int(50e6))
        step_dev(*devb[i % nb])
    sync_all()
    timing, med.TIMING = med.TIMING, None
    ctiming, CNV.TIMING = CNV.TIMING, None
    BB.USE_SIDE_STREAM = True

    # MED kernels: HBM roofline from the algorithmic bytes of SURVEY.md 8(d)
    kinds = {}
    for kind, s0, s1, nbytes in timing:
        kinds.setdefault(kind, []).append((s0.elapsed_time(s1), nbytes))
    # an event pair brackets its kernel in STREAM order: when the eager host falls behind for one launch, the pair also times
    # the idle gap in front of the kernel (seen as one 3x sample among ten).  Samples beyond 1.5x the median are dropped.
    for kind, v in list(kinds.items()):
        med_t = sorted(x for x, _ in v)[len(v) // 2]
        kept = [(x, nb_) for x, nb_ in v if x <= 1.5 * med_t]
        kinds[kind] = kept if kept else v
    n_med_samples = {k: len(v) for k, v in kinds.items()}
    med_stats = {k: {"launches_per_step": round(sum(1 for kk, *_ in timing if kk == k) / n_prof, 3), "samples_kept": n_med_samples[k],
                     "avg_ms": sum(x for x, _ in v) / len(v),
                     "gbs": sum(nb_ for _, nb_ in v) / sum(x for x, _ in v) / 1e6,
                     "frac": sum(nb_ for _, nb_ in v) / sum(x for x, _ in v) / 1e6 / hbm_peak,
                     "share_of_step": (sum(x for x, _ in v) / len(v)) * (sum(1 for kk, *_ in timing if kk == k) / n_prof) / ms}
                 for k, v in kinds.items()}
    # convolution kernels (tcgen05 tile / row / wgrad): tensor roofline from the algorithmic FLOPs; layer by layer the
    # bound is max(flops / tensor peak, bytes / HBM peak) because the <= 96-channel layers sit below the bf16 ridge
    fam = {}
    smem_gbs = smem_port_gbs()
    # the same outlier rule per layer shape: a launch's time = the mean of its shape's samples within 1.5x their median
    groups = {}
    for kind, s0, s1, fl, nbytes, gemm_n in ctiming:
        groups.setdefault((kind, fl, nbytes, gemm_n), []).append(s0.elapsed_time(s1))
    shape_ms = {}
    for key, ts in groups.items():
        med_t = sorted(ts)[len(ts) // 2]
        kept = [x for x in ts if x <= 1.5 * med_t] or ts
        shape_ms[key] = sum(kept) / len(kept)
    for kind, s0, s1, fl, nbytes, gemm_n in ctiming:
        f = fam.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "bound_ms": 0.0, "bound_ss_ms": 0.0, "n": 0})
        f["ms"] += shape_ms[(kind, fl, nbytes, gemm_n)]
        f["flops"] += fl
        f["bytes"] += nbytes
        b0 = max(fl / (tf_peak * 1e9), nbytes / (hbm_peak * 1e6))
        f["bound_ms"] += b0
        f["bound_ss_ms"] += max(b0, _ss_bytes(fl, gemm_n) / (smem_gbs * 1e6))
        f["n"] += 1
    conv_stats = {k: {"launches_per_step": f["n"] / n_prof, "ms_per_step": f["ms"] / n_prof,
                      "tflops": f["flops"] / f["ms"] / 1e9, "frac_of_tensor_peak": f["flops"] / f["ms"] / 1e9 / tf_peak,
                      "frac_of_layerwise_roofline": f["bound_ms"] / f["ms"],
                      "frac_of_layerwise_roofline_ss": f["bound_ss_ms"] / f["ms"], "share_of_step": f["ms"] / n_prof / ms}
                  for k, f in fam.items() if f["ms"] > 0}
    roofline = None
    if conv_stats:
        tot_ms = sum(f["ms"] for f in fam.values())
        tot_fl = sum(f["flops"] for f in fam.values())
        tot_bound = sum(f["bound_ms"] for f in fam.values())
        tot_bound_ss = sum(f["bound_ss_ms"] for f in fam.values())
        roofline = {"kernel": "conv3x3 family (tcgen05 tile/row kernels: forward + dgrad, MN-major wgrad)", "bound": "tensor",
                    "achieved": tot_fl / tot_ms / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": tot_fl / tot_ms / 1e9 / tf_peak, "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained",
                    "frac_of_layerwise_roofline": tot_bound / tot_ms,
                    # the same with the shared-memory operand port of SS-mode tcgen05.mma as a third per-layer bound
                    "frac_of_layerwise_roofline_ss": tot_bound_ss / tot_ms,
                    "smem_port": {"gbs": smem_gbs, "source": "nominal: 128 B/clk x 148 SMs x sm_max_mhz"},
                    # `traffic` stays null for the family as a whole; per-launch DRAM traffic (dram__bytes_read.sum +
                    # dram__bytes_write.sum) of its largest kernels from the committed `ncu --set full` captures
                    # (profiles/r2g_conv_full.txt, profiles/r2q_new_kernels_full.txt) beside their algorithmic bytes:
                    "traffic_ncu": [
                        {"kernel": "conv3x3_row_kernel<64,64,2>", "layer": "deconv1 forward 64->64, 8x192x640", "dram_mb": 211.5,
                         "algorithmic_mb": 251.7, "note": "part of the output is still in L2 when the kernel ends"},
                        {"kernel": "conv3x3_wgrad_kernel<128,128,1>", "layer": "deconv1-sized weight gradient 64->64, 8x192x640",
                         "dram_mb": 255.6, "algorithmic_mb": 251.8},
                        {"kernel": "conv3x3_wgrad_up2_kernel", "layer": "deconv1 folded weight gradient (input 8x96x320x64)",
                         "dram_mb": 161.7, "algorithmic_mb": 157.4},
                        {"kernel": "conv3x3_row_kernel<32,32,2>", "layer": "conv0_1 forward 32->32, 8x192x640", "dram_mb": 87.5,
                         "algorithmic_mb": 125.8, "note": "part of the output is still in L2 when the kernel ends"}],
                    "launches_per_step": sum(f["n"] for f in fam.values()) / n_prof, "ms_per_step": tot_ms / n_prof,
                    "share_of_step": tot_ms / n_prof / ms, "by_kernel": conv_stats}
    roofline_med = None
    if med_stats:
        dom = max(med_stats.items(), key=lambda kv: kv[1]["avg_ms"] * kv[1]["launches_per_step"])[0]
        st = med_stats[dom]
        # `traffic` stays null (no ncu capture of THIS batch shape); `traffic_ncu` quotes dram__bytes_read.sum +
        # dram__bytes_write.sum per launch of the committed `ncu --set full` captures (profiles/r1e_med3_full_*) beside the
        # algorithmic bytes of the captured shape: DRAM traffic ~ algorithmic bytes, i.e. no wasted re-reads
        ncu_traffic = {"med_fwd": {"shape": "16x49x192x640", "dram_mb": 447.7, "algorithmic_mb": 440.4},
                       "med_fwd_masks": {"shape": "16x49x192x640", "dram_mb": 468.3, "algorithmic_mb": 456.1},
                       "med_bwd": {"shape": "8x49x375x1242", "dram_mb": 1888.7, "algorithmic_mb": 1564.9}}
        roofline_med = {"kernel": dom, "bound": "hbm", "achieved": st["gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": st["frac"], "traffic": None, "traffic_ncu": ncu_traffic.get(dom),
                        "peak_source": peak_src + " hbm_gbs",
                        "avg_launch_ms": st["avg_ms"], "share_of_step": st["share_of_step"], "all_med_kernels": med_stats}
        if roofline is None:
            roofline, roofline_med = roofline_med, None

    # ---------------- end-to-end timing (e2e): pinned host inputs, H2D + D2H inside the timed region ----------------
    h2d = 2 * B * 3 * H * W * 4 if wl in ("stage1", "stage2") else B * 3 * H * W * 4
    sync_all()
    e0.record()
    out_host = torch.empty(1, pin_memory=True)
    if graphed is not None:
        graphed.prefetch(*host[0])                        # the input pipeline: batch i+1 is uploaded while step i runs
    pending, loss_sum = None, 0.0
    for i in range(a.steps):
        l, r = host[i % nb]
        if graphed is not None:
            # consumes the staged batch (every H2D copy is inside the timed region) and queues the D2H copy of this step's loss
            ticket = graphed.run_async()
            if i + 1 < a.steps:
                graphed.prefetch(*host[(i + 1) % nb])
            if pending is not None:
                loss_sum += graphed.loss_value(pending)    # the user reads EVERY step's loss, one step behind the device
            pending = ticket
        else:
            # no captured graph (inference): the same pipeline by hand -- batch i + 1 is uploaded on a copy stream while
            # step i runs, and the result of step i - 1 is read while step i is queued
            cur = torch.cuda.current_stream()
            if i == 0:
                copy_stream = torch.cuda.Stream()
                res_ring = torch.empty(2, pin_memory=True)
                with torch.cuda.stream(copy_stream):
                    nxt = (l.to(dev, non_blocking=True), r.to(dev, non_blocking=True) if wl in ("stage1", "stage2") else None)
                    up_ev = torch.cuda.Event()
                    up_ev.record(copy_stream)
            cur.wait_event(up_ev)
            ld, rd = nxt
            loss = run_step(ld, rd)
            if i + 1 < a.steps:
                l2, r2 = host[(i + 1) % nb]
                with torch.cuda.stream(copy_stream):
                    nxt = (l2.to(dev, non_blocking=True), r2.to(dev, non_blocking=True) if wl in ("stage1", "stage2") else None)
                    up_ev = torch.cuda.Event()
                    up_ev.record(copy_stream)
            for t_ in (ld, rd):
                if t_ is not None:
                    t_.record_stream(cur)                  # allocated on the copy stream, consumed on the compute stream
            res_ring[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            ev_i = torch.cuda.Event()
            ev_i.record(cur)
            if pending is not None:
                pending[0].synchronize()                   # the user reads EVERY step's result, one step behind the device
                loss_sum += float(res_ring[pending[1]])
            pending = (ev_i, i & 1)
    if pending is not None:
        if graphed is not None:
            loss_sum += graphed.loss_value(pending)        # ... and the last one before the clock stops
        else:
            pending[0].synchronize()
            loss_sum += float(res_ring[pending[1]])
    e1.record()
    if not (loss_sum == loss_sum):
        raise RuntimeError("e2e loop produced a NaN loss")
    sync_all()
    ms_e2e = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t[0])

    # ---------------- exposed communication (N > 1): the same step captured again WITHOUT the gradient all-reduce ----------------
    comm = None
    if world > 1 and graphed is not None:
        opt.comm_enabled = False
        g2 = run_gpu.GraphedStep(opt, loss_fn, devb[0][0], devb[0][1], warmup=1)
        for i in range(a.warmup):
            g2.run(*devb[i % nb])
        sync_all()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(a.steps):
            g2.run(*devb[i % nb])
        c1.record()
        sync_all()
        t2 = torch.tensor([c0.elapsed_time(c1) / a.steps], device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        opt.comm_enabled = True
        comm = {"collective": "ncclAllReduce fp32 SUM per bucket, launched as the bucket's last gradient lands; Adam per "
                              "bucket behind it" if opt.bucket_adam else "ncclAllReduce fp32 SUM per bucket",
                "buckets_mb": [round((e_ - s_) * 4 / 1e6, 2) for s_, e_, _ in opt.buckets],
                "ms_per_step_without_allreduce": float(t2[0]), "exposed_ms": ms - float(t2[0])}

    res = {
        "metric": metric_name(wl), "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 convolutions (fp32 accumulate) + fp32 MED/losses/Adam", "data": "synthetic",
        "config": cfg, "clocks": clk,
        "e2e": {"value": frames_per_step / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "readback": ("every step's loss is copied D2H behind the step and read by the host one step late "
                             "(GraphedStep.run_async / loss_value); the last one before the clock stops") if graphed is not None
                else "inputs of step i + 1 uploaded on a copy stream while step i runs; every step's result is copied D2H "
                     "behind the step and read one step late; the last one before the clock stops"},
        "gpu_launches": launches, "cuda_graph": graphed is not None,
        "library_conv_calls_in_timed_region": lib_convs,
        "roofline": roofline,
    }
    if roofline_med is not None:
        res["roofline_med"] = roofline_med
    if comm is not None:
        res["comm"] = comm
    return res


if __name__ == "__main__":
    sys.exit(main())
