#!/usr/bin/env python
"""bench.py -- headline benchmark of the FAL-net hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stage1|stage2|test|med] [--impl reference]

One JSON line on stdout (rank 0).  Default workload = BASELINE.json configs[1]: a Stage-1 training step
(reconstruction + smoothness loss, Adam included), batch 8 per GPU, 640x192 crops, N = 49, synthetic
KITTI-shaped stereo pairs, random-init weights (models.FAL_netB under torch.manual_seed(0)).
A "step" = forward + losses + backward + gradient all-reduce (N > 1) + Adam on one batch.
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
# reference arm / cpu_baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference(workload, steps, warmup, sample_b=None, budget_s=25.0):
    """Times oracle/falnet_oracle.py (the CPU restatement of the reference, bit-identical to it on CPU) on a
    bounded sample of the workload.  Returns (value, unit, cores, sample_description, ms_per_step)."""
    from oracle import falnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N = 49
    if workload in ("stage1", "stage2"):
        H, W = 192, 640
        B = sample_b or 1
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
        unit, per_step = "frames/s", B * (1 if workload == "stage1" else 2)
        sample = f"{workload} step on {B} of 8 {'pairs' if workload == 'stage2' else 'images'}, 192x640, N=49, fp32 torch CPU"
    else:
        H, W = 375, 1242
        B = sample_b or 1
        p = O.init_params(N, seed=0)
        img = synth_batch(B, H, W, 1234)[0]
        mx = torch.full((B, 1, 1), 300.0)
        mn = mx * 2 / 300

        def step():
            with torch.no_grad():
                d = O.test_disp_fpp(p, img, mn, mx, flip=lambda t: torch.flip(t, dims=[3]))
            return float(d.mean())
        unit, per_step = "frames/s", B
        sample = f"Test_KITTI flip-PP on {B} of 8 images, 375x1242, N=49, fp32 torch CPU"
    t0 = time.time()
    for _ in range(max(1, min(warmup, 1))):
        step()
    first = time.time() - t0
    n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.time()
    for _ in range(n):
        step()
    dt = (time.time() - t0) / n
    return per_step / dt, unit, cores, sample + f", {n} timed step(s)", dt * 1e3


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="stage1", choices=["stage1", "stage2", "test", "med"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the training step kernel by kernel instead of replaying its CUDA graph")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        val, unit, cores, sample, ms = cpu_reference(a.workload if a.workload != "med" else "stage1", a.steps, a.warmup,
                                                     budget_s=90.0)
        print(json.dumps({
            "impl": "reference", "metric": metric_name(a.workload), "value": val, "unit": unit, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a.workload, a.gpus),
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
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

    result = run_gpu(a, rank, world, dev, dist, _lib, med, models, steps, LF, FlatAdamDDP, C)
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


def run_gpu(a, rank, world, dev, dist, _lib, med, models, steps, LF, FlatAdamDDP, C):
    N = 49
    wl = a.workload
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
    n_prof = max(3, min(a.steps, 10))
    BB.USE_SIDE_STREAM = False
    step_dev(*devb[0])                                               # settle allocator / caches in this mode
    med.TIMING, CNV.TIMING = [], []
    for i in range(n_prof):
        step_dev(*devb[i % nb])
    sync_all()
    timing, med.TIMING = med.TIMING, None
    ctiming, CNV.TIMING = CNV.TIMING, None
    BB.USE_SIDE_STREAM = True

    # MED kernels: HBM roofline from the algorithmic bytes of SURVEY.md 8(d)
    kinds = {}
    for kind, s0, s1, nbytes in timing:
        kinds.setdefault(kind, []).append((s0.elapsed_time(s1), nbytes))
    med_stats = {k: {"launches_per_step": len(v) / n_prof, "avg_ms": sum(x for x, _ in v) / len(v),
                     "gbs": sum(nb_ for _, nb_ in v) / sum(x for x, _ in v) / 1e6,
                     "frac": sum(nb_ for _, nb_ in v) / sum(x for x, _ in v) / 1e6 / hbm_peak,
                     "share_of_step": sum(x for x, _ in v) / n_prof / ms} for k, v in kinds.items()}
    # convolution kernels (tcgen05 tile / row / wgrad): tensor roofline from the algorithmic FLOPs; layer by layer the
    # bound is max(flops / tensor peak, bytes / HBM peak) because the <= 96-channel layers sit below the bf16 ridge
    fam = {}
    for kind, s0, s1, fl, nbytes in ctiming:
        f = fam.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "bound_ms": 0.0, "n": 0})
        f["ms"] += s0.elapsed_time(s1)
        f["flops"] += fl
        f["bytes"] += nbytes
        f["bound_ms"] += max(fl / (tf_peak * 1e9), nbytes / (hbm_peak * 1e6))
        f["n"] += 1
    conv_stats = {k: {"launches_per_step": f["n"] / n_prof, "ms_per_step": f["ms"] / n_prof,
                      "tflops": f["flops"] / f["ms"] / 1e9, "frac_of_tensor_peak": f["flops"] / f["ms"] / 1e9 / tf_peak,
                      "frac_of_layerwise_roofline": f["bound_ms"] / f["ms"], "share_of_step": f["ms"] / n_prof / ms}
                  for k, f in fam.items() if f["ms"] > 0}
    roofline = None
    if conv_stats:
        tot_ms = sum(f["ms"] for f in fam.values())
        tot_fl = sum(f["flops"] for f in fam.values())
        tot_bound = sum(f["bound_ms"] for f in fam.values())
        roofline = {"kernel": "conv3x3 family (tcgen05 tile/row kernels: forward + dgrad, MN-major wgrad)", "bound": "tensor",
                    "achieved": tot_fl / tot_ms / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": tot_fl / tot_ms / 1e9 / tf_peak, "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained",
                    "frac_of_layerwise_roofline": tot_bound / tot_ms,
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
    for i in range(a.steps):
        l, r = host[i % nb]
        if graphed is not None:
            loss = graphed.run()                          # consumes the staged batch (every H2D copy is inside the timed region)
            if i + 1 < a.steps:
                graphed.prefetch(*host[(i + 1) % nb])
        else:
            ld = l.to(dev, non_blocking=True)
            rd = r.to(dev, non_blocking=True) if wl in ("stage1", "stage2") else None
            loss = run_step(ld, rd)
        out_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the loss every step
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t[0])

    res = {
        "metric": metric_name(wl), "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 convolutions (fp32 accumulate) + fp32 MED/losses/Adam", "data": "synthetic",
        "config": cfg, "clocks": clk,
        "e2e": {"value": frames_per_step / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "cuda_graph": graphed is not None,
        "library_conv_calls_in_timed_region": lib_convs,
        "roofline": roofline,
    }
    if roofline_med is not None:
        res["roofline_med"] = roofline_med
    if rank == 0 and not a.no_cpu_baseline and world == 1:
        val, unit, cores, sample, cms = cpu_reference(wl if wl != "med" else "stage1", 3, 1, budget_s=20.0)
        res["cpu_baseline"] = {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                               "ms_per_step": cms}
    return res


if __name__ == "__main__":
    sys.exit(main())
