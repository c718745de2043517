"""The FAL_netB encoder-decoder as ONE autograd node with a hand-scheduled backward.

Forward = /root/reference/models/FAL_netB.py:140-176 + the logit 1x1 conv (:190,215), executed layer by layer on the
tcgen05 kernels (fal_net_b200.conv_native) while a tape keeps the bf16 NHWC activations.

Backward = the exact adjoint of that graph, written out by hand instead of being discovered by autograd:

  * data gradients run on the SAME tcgen05 kernel as the forward (``faln_conv3x3_dgrad``: re-packed weights, stride-2
    layers as four output-parity classes) and every chain-rule elementwise step is fused into an epilogue --
    ELU'(saved output), the residual pass-through of the res blocks, the accumulation of the second consumer of a
    skip tensor, the split of a concatenated input into its two sources (two row ranges of the packed weights);
  * the nearest-upsample backward is one kernel fused with the producer's ELU';
  * bias gradients are one channel-sum kernel each;
  * weight gradients: ``faln_conv3x3_wgrad`` (tcgen05 with MN-major operands, split-K over pixels) reduces straight
    into the fp32 gradient arena of the optimiser -- no cuDNN, no cast / add / cat launches around it
    (``LIBRARY_CALLS['conv_wgrad']`` stays 0; bench.py reports it).

Compared with per-layer autograd.Functions this removes ~700 ATen elementwise / cat / copy / reduce launches per step
(see profiles/): nothing is concatenated, no gradient is materialised twice, and the schedule is static -- which is what
lets the whole step live in one CUDA graph.
"""
from __future__ import annotations

import os

import torch

from . import conv_native as CN
from . import layout

CL = torch.channels_last
LIBRARY_CALLS = {"conv_wgrad": 0}

# The encoder / decoder stage tables ((name, cin, cout, stride) and (level, up_in, up_out, skip_ch, iconv_out)) come from
# the model's ``_spec`` (fal_net_b200.models._falnet: FAL_netA / B / C differ only in those tables, in the attribute name
# of the encoder-decoder and in A's separable residual kernels).

from .conv import GENERATION  # noqa: E402  (shared: bumped by the trainer after each in-place optimiser step)


def _cached(weight, tag, fn):
    """bf16 re-packings of a parameter, cached on the parameter object and refreshed when it changes."""
    if weight.grad_fn is not None:
        return fn()
    key = (weight._version, GENERATION[0] if weight.requires_grad else -1, weight.data_ptr())
    cache = getattr(weight, "_faln_packs", None)
    if cache is None or cache.get("key") != key:
        cache = {"key": key}
        weight._faln_packs = cache
    if tag not in cache:
        cache[tag] = fn()
    return cache[tag]


def _embed3(w):
    """A 3x1 / 1x3 kernel (FAL_netA's residual blocks, /root/reference/models/FAL_netA.py:73-76) as the 3x3 kernel with the
    absent taps zero -- exact, and it lets every layer run on the same tcgen05 3x3 kernels."""
    if w.shape[2:] == (3, 3):
        return w
    w3 = torch.zeros(w.shape[0], w.shape[1], 3, 3, device=w.device, dtype=w.dtype)
    if w.shape[2:] == (3, 1):
        w3[:, :, :, 1] = w[:, :, :, 0]
    elif w.shape[2:] == (1, 3):
        w3[:, :, 1, :] = w[:, :, 0, :]
    else:
        raise RuntimeError(f"unsupported conv kernel shape {tuple(w.shape)}")
    return w3


# Nearest 2x up-sampling folded into the deconv convolution (CN.conv3x3_up2_fwd / _dgrad: four taps per output parity class
# instead of nine, no up-sampled tensor).  Exact-2x levels only (every level of the 192x640 training crops); other sizes --
# KITTI full resolution has odd levels -- keep the explicit up-sample.  FALN_NO_UP2=1 switches it off (A/B measurements).
USE_UP2 = os.environ.get("FALN_NO_UP2", "0") in ("", "0")
# ... and its weight gradient from the low-resolution input (FALN_NO_UP2_WGRAD=1: rebuild the up-sampled map on the side stream
# and run the plain 3x3 weight-gradient kernel on it, the round-2 path before this kernel existed)
# measurement aid only (wrong gradients!): leave the bias-gradient sums out, to see what they cost inside the overlapped step
_SKIP_BIAS_SUMS = os.environ.get("FALN_DEBUG_SKIP_BIAS_SUMS", "0") not in ("", "0")
# Bias gradients from the spare operand slot of the layer's weight-gradient launch (FALN_NO_FUSED_BIAS_GRAD=1: one channel-sum
# launch per biased layer on the bias stream instead -- 13 launches that re-read every gradient map; measured as 2.7 % of the
# Stage-1 step by leaving them out, FALN_DEBUG_SKIP_BIAS_SUMS).
# FALN_PREP_ON_SIDE=1: start the small per-step weight transforms of the forward (folded deconv packs, folded logits layer,
# constant-channel table) on the side stream beside the encoder instead of inline.  Measured on B200 (100-step runs, twice
# each): Stage-1 step 3.811 ms on the side stream vs 3.784 ms inline -- SLOWER, like every other attempt to put work beside
# the full-resolution layers that open the forward.  Off by default.
PREP_ON_SIDE = os.environ.get("FALN_PREP_ON_SIDE", "0") not in ("", "0")
# The stem's weight / bias gradient from the fp32 image (stem_wgrad_mma_kernel); FALN_NO_STEM_WGRAD=1: the generic kernel on a
# 32-channel bf16 copy of the image (a 63 MB transpose + a full 32 x 32-channel launch for 27 x 32 numbers).
STEM_WGRAD_MMA = os.environ.get("FALN_NO_STEM_WGRAD", "0") in ("", "0")
# Weight gradients of the small- and mid-map layers (at most SMALL_WGRAD_CHUNKS 64-pixel chunks: the 3x10 ... 48x160 maps of the
# 192x640 crop) are collected and launched as ONE grid per kernel configuration (CN.conv3x3_wgrad_multi) when the first larger
# layer follows.  Measured on B200 (100-step runs, twice each, one box): Stage-1 step 3.734 ms without batching, 3.708 ms with a
# 96-chunk limit (3x10 ... 12x40), 3.665 ms at 1000 (... 48x160), 3.717 at 4000, 3.711 with everything at the end of backward.
# FALN_NO_WGRAD_BATCH=1: one launch per layer and source, as before.
UP2_PACKS_BATCHED = os.environ.get("FALN_NO_UP2_PACK_BATCH", "0") in ("", "0")
BATCH_SMALL_WGRAD = os.environ.get("FALN_NO_WGRAD_BATCH", "0") in ("", "0")
SMALL_WGRAD_CHUNKS = int(os.environ.get("FALN_WGRAD_BATCH_CHUNKS", "1000"))
# the same for the folded deconv weight gradients (limit on the OUTPUT map's chunks; 0 = separate launches).  Measured on B200
# (100-step runs, twice each): 3.669 / 3.665 ms batched up to 48x160 vs 3.657 / 3.668 ms separate -- no gain (four jobs that are
# not back to back in the backward); off by default, the kernel stays for shapes where the levels are smaller.
UP2_WGRAD_BATCH_CHUNKS = int(os.environ.get("FALN_WGRAD_UP2_BATCH", "0"))
FUSE_BIAS_GRAD = os.environ.get("FALN_NO_FUSED_BIAS_GRAD", "0") in ("", "0")
USE_UP2_WGRAD = os.environ.get("FALN_NO_UP2_WGRAD", "0") in ("", "0")


def _up2_packs(weight):
    """(forward, dgrad) folded packs of a deconv weight: one launch, refreshed once per optimiser step."""
    return _cached(weight, ("up2",), lambda: CN.pack_up2_weights(weight))


def _up2_packs_all(weights):
    """The folded packs of ALL deconv levels in one launch when any of them is stale (training: once per step) -- six launches
    of ~5 us each sat in the forward's main chain, right in front of their consumers.  Fills the per-weight caches."""
    stale = []
    for w in weights:
        if w.grad_fn is not None:
            return
        key = (w._version, GENERATION[0] if w.requires_grad else -1, w.data_ptr())
        cache = getattr(w, "_faln_packs", None)
        if cache is None or cache.get("key") != key or ("up2",) not in cache:
            stale.append((w, key))
    if len(stale) < 2:
        return
    packs = CN.pack_up2_weights_multi([w for w, _ in stale])
    for (w, key), pk in zip(stale, packs):
        cache = getattr(w, "_faln_packs", None)
        if cache is None or cache.get("key") != key:
            cache = {"key": key}
            w._faln_packs = cache
        cache[("up2",)] = pk


def _ctab(weight, channel):
    """Border-class table of the constant input channel (conv1.0's 33rd input), refreshed once per optimiser step."""
    return _cached(weight, ("ctab", channel), lambda: CN.const_channel_table_of(weight, channel))


def _wk(weight, cin=None):
    """Forward (KRSC bf16) weight: a view of the optimiser's bf16 shadow arena when the parameter lives in one (no pack
    kernels at all), else packed once per parameter version."""
    cin = cin or weight.shape[1]
    ar = getattr(weight, "_faln_arena", None)
    if ar is not None and weight.grad_fn is None:
        w = ar[0].packed_fwd(ar[1], cin)
        if w is not None:
            return w
    return _cached(weight, ("fwd", cin), lambda: CN.pack_weight(_embed3(weight.detach())[:, :cin]))


def _wd(weight, cin=None):
    """Data-gradient weight [Cin,3,3,Cout] bf16: the optimiser's batched per-step re-pack, else packed per version."""
    cin = cin or weight.shape[1]
    ar = getattr(weight, "_faln_arena", None)
    if ar is not None and weight.grad_fn is None:
        w = ar[0].packed_dgrad(ar[1], cin)
        if w is not None:
            return w
    return _cached(weight, ("dgrad", cin), lambda: CN.pack_weight_dgrad(_embed3(weight.detach())[:, :cin]))


def fold_logit_conv(w_iconv1, w0):
    """iconv1 (3x3, no bias, no activation; reference :127,174) followed by conv0 (1x1 + bias; :190,215)
    == one 3x3 conv with W'[o,c,kh,kw] = sum_m W0[o,m] * W_iconv1[m,c,kh,kw].  (torch form: tests / documentation; the
    product path is CN.fold_logit_conv, one kernel that writes the bf16 packs directly.)"""
    return torch.einsum("om,mckl->ockl", w0[:, :, 0, 0], w_iconv1)


def _folded_packs(model, rows_pad):
    """(forward pack, dgrad pack) of the folded logits layer; recomputed when either parameter changed (training: once per
    step, one launch), cached otherwise (frozen model, inference)."""
    wi, w0 = model.bb.iconv1.weight, model.conv0.weight
    key = (wi._version, w0._version, wi.data_ptr(), w0.data_ptr(), rows_pad,
           GENERATION[0] if (wi.requires_grad or w0.requires_grad) else -1)
    cache = getattr(model, "_faln_fold", None)
    if cache is None or cache[0] != key:
        cache = (key, CN.fold_logit_conv(wi, w0, rows_pad))
        model._faln_fold = cache
    return cache[1]


# ------------------------------------------------------------------------------------------------------------------
def forward(model, image, max_disp, tape=None, disp_lvl=None):
    """dlog0 [B,N,H,W] fp32 planar (16-byte-multiple row pitch).  ``tape`` (a dict) receives what backward needs.
    With ``disp_lvl`` [B,N] (inference, no tape) the last layer's epilogue takes the softmax-expectation itself and the
    function returns the disparity [B,1,H,W]: the logits never reach HBM."""
    bb = model.bb
    ENC, DEC = model._spec.enc, model._spec.dec
    B = image.shape[0]
    sink = getattr(model, "_faln_grad_sink", None)
    if tape is not None and sink is not None and hasattr(sink, "start_repack"):
        sink.start_repack(_side_streams(image.device)[0])      # data-gradient weight packs: overlapped with this forward
    flow_val = (max_disp.reshape(B).float() / 100.0).contiguous()                  # :208-209, constant plane per sample
    # Optional (PREP_ON_SIDE, measured slower, off): the small per-step weight transforms (six folded deconv packs, the folded
    # logits layer, the constant-channel table) start on the side stream beside the encoder; the inline calls below then hit
    # the caches.
    prep_ev = None
    if tape is not None and USE_SIDE_STREAM and PREP_ON_SIDE and image.is_cuda:
        main, side = torch.cuda.current_stream(image.device), _side_streams(image.device)[0]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if USE_UP2 and image.shape[2] % 64 == 0 and image.shape[3] % 64 == 0:      # exact 2x at every level
                for lvl, *_ in DEC:
                    up = getattr(bb, f"deconv{lvl}")
                    if up.conv1.weight.shape[1] % 32 == 0:
                        _up2_packs(up.conv1.weight)
            _folded_packs(model, None)
            _ctab(getattr(bb, ENC[1][0])[0].weight, ENC[0][2])
            prep_ev = torch.cuda.Event()
            prep_ev.record(side)
    skips = []
    for i, (name, _, cout, stride) in enumerate(ENC):
        head = getattr(bb, name)[0]
        if i == 0:
            a = CN.stem_conv(image, head.weight, head.bias, 1)                     # reads the fp32 NCHW image directly
        else:
            C1 = skips[-1].shape[1]
            if i == 1 and prep_ev is not None:
                torch.cuda.current_stream(image.device).wait_event(prep_ev)
            ctab = _ctab(head.weight, C1) if i == 1 else None
            a = CN.conv3x3_fwd(skips[-1], _wk(head.weight, C1), head.bias, stride, 1, cout=cout, ctab=ctab,
                               cscale=flow_val if i == 1 else None)
        blk = getattr(bb, name + "_1")
        r = CN.conv3x3_fwd(a, _wk(blk.conv1.weight), None, 1, 1)
        s = CN.conv3x3_fwd(r, _wk(blk.conv2.weight), None, 1, 1, residual=a)       # elu(conv2(elu(conv1(x))) + x), :79
        skips.append(s)
        if tape is not None:
            tape[name] = (a, r, s)
    h = skips[6]
    if USE_UP2 and UP2_PACKS_BATCHED and image.shape[2] % 64 == 0 and image.shape[3] % 64 == 0:    # exact 2x at every level
        _up2_packs_all([getattr(bb, f"deconv{lvl}").conv1.weight for lvl, *_ in DEC
                        if getattr(bb, f"deconv{lvl}").conv1.weight.shape[1] % 32 == 0])
    for lvl, _, uout, _, iout in DEC:
        skip = skips[lvl - 1]
        up = getattr(bb, f"deconv{lvl}")
        if USE_UP2 and skip.shape[2] == 2 * h.shape[2] and skip.shape[3] == 2 * h.shape[3] and h.shape[1] % 32 == 0:
            xu = None                                                              # :58-59 as ONE kernel, xu never exists
            u = CN.conv3x3_up2_fwd(h, _up2_packs(up.conv1.weight)[0], None, 1)
        else:
            xu = CN.upsample_nearest(h, (skip.shape[2], skip.shape[3]))            # :58
            u = CN.conv3x3_fwd(xu, _wk(up.conv1.weight), None, 1, 1)               # :59
        if iout is not None:
            ic = getattr(bb, f"iconv{lvl}")[0]
            hn = CN.conv3x3_fwd(u, _wk(ic.weight), ic.bias, 1, 1, None, skip)      # concat = second TMA source
            if tape is not None:
                tape[f"dec{lvl}"] = (h, xu, u, hn)
            h = hn
        else:
            N = model.no_levels
            Bq, _, H, W = u.shape
            fused = disp_lvl is not None and tape is None and CN.logits_disp_supported(W, N)
            # W' = conv0 . iconv1 as the forward pack and the data-gradient pack, one launch; frozen weights: cached
            wfk, wfd = _folded_packs(model, 64 if fused else None)
            if fused:
                return CN.conv3x3_logits_disp(u, skip, wfk, model.conv0.bias, disp_lvl)
            out = layout.alloc_planar(Bq, N, H, W, u.device)
            CN.conv3x3_fwd(u, wfk, model.conv0.bias, 1, 0, None, skip, cout=N, planar_out=out)
            if tape is not None:
                tape["dec1"] = (h, xu, u, None)
                tape["wfd"] = wfd
                tape["flow_val"] = flow_val
                tape["image"] = image
            return out
    raise AssertionError("unreachable")


def disparity(model, image, min_disp, max_disp):
    """Inference path (/root/reference/models/FAL_netB.py:228-229, Test_KITTI.py:196): disparity only, no autograd."""
    from . import med
    if not image.is_cuda:
        raise RuntimeError("fal_net_b200.FAL_netB runs on CUDA (sm_100a) only; there is no CPU path")
    N = model.no_levels
    d_lvl, _ = med.level_tables(min_disp, max_disp, N, image.shape[3])
    out = forward(model, image, max_disp, None, disp_lvl=d_lvl)
    if out.shape[1] == 1 and N != 1:
        return out
    return med.med_disp_only(out, d_lvl)


# ------------------------------------------------------------------------------------------------------------------
_SIDE_STREAMS: dict = {}
USE_SIDE_STREAM = True      # bench.py switches this off for its per-kernel timing pass (kernels then run one at a time)
# Parameter-gradient tasks (weight gradients, bias sums) of different layers are independent and can be dealt round-robin
# onto several side streams (FALN_SIDE_STREAMS).  Measured on B200 (gpurun_out/s6_*): the data-gradient chain on the main
# stream is the critical path and more concurrency beside it only takes SMs away from it -- Stage-1 step 5.04 ms with one
# side stream, 5.26 ms with two to four; Stage-2 15.67 vs 15.97 ms.  Default: one.
N_SIDE_STREAMS = max(1, int(os.environ.get("FALN_SIDE_STREAMS", "1")))
# The bias-gradient sums are deliberately small-footprint, long-latency kernels (two blocks per SM): on the weight-gradient
# stream they add ~0.28 ms of serial latency to a side chain that is now as long as the data-gradient chain.  FALN_BIAS_STREAM=1
# gives them a stream of their own (same footprint, no serialisation with the weight gradients).
BIAS_STREAM = os.environ.get("FALN_BIAS_STREAM", "1") not in ("", "0")


def _side_streams(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    n = N_SIDE_STREAMS + (1 if BIAS_STREAM else 0)
    if st is None or len(st) != n:
        st = _SIDE_STREAMS[key] = [torch.cuda.Stream(device=key) for _ in range(n)]
    return st


class _DictSink:
    """Gradient destination when no flat arena is attached: fresh zeroed fp32 tensors, returned to autograd."""

    def __init__(self, model, device):
        self.shapes = {n: p.shape for n, p in model.named_parameters()}
        self.device = device
        self.grads = {}

    def grad_view(self, name):
        if name not in self.grads:
            shp = self.shapes[name]
            t = torch.zeros(shp, device=self.device, dtype=torch.float32)
            self.grads[name] = t.contiguous(memory_format=CL) if len(shp) == 4 and shp[2:] == (3, 3) else t
        return self.grads[name]

    def mark_ready(self, name):
        pass


def backward(model, tape, g_logits, sink=None):
    """Hand-scheduled backward.  Parameter gradients are ACCUMULATED into ``sink.grad_view(name)`` (fp32, the flat gradient
    arena of trainer.FlatAdamDDP when one is attached to the model -- then ``sink.mark_ready(name)`` releases all-reduce
    buckets as soon as they are complete); returns the sink."""
    bb = model.bb
    ENC, DEC = model._spec.enc, model._spec.dec
    pfx = model._spec.bb_attr + "."
    N = model.no_levels
    B, _, H, W = g_logits.shape
    dev = g_logits.device
    sink = sink or _DictSink(model, dev)

    # Parameter gradients are off the critical path (only the data-gradient chain feeds the next layer): they run on a
    # side stream, ordered after their producer by an event, and are joined before returning.  Inside a CUDA-graph capture
    # this becomes a parallel branch of the graph, so the latency-bound small layers of both chains overlap.
    main = torch.cuda.current_stream(dev)
    sides = _side_streams(dev)
    keep = []                                                  # tensors the side streams read stay alive until the join
    turn = [0]
    used = []                                                  # side streams that received work in THIS backward: only those are
    #                                                            joined (inside a graph capture a stream that never forked from the
    #                                                            capturing stream must not be waited on)

    def on_side(fn, *tensors, bias=False):
        if not USE_SIDE_STREAM:
            fn()
            return
        keep.extend(tensors)
        if bias and BIAS_STREAM:
            side = sides[-1]                                   # the bias-sum stream
        else:
            side = sides[turn[0] % N_SIDE_STREAMS]
            turn[0] += 1
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        if side not in used:
            used.append(side)
        with torch.cuda.stream(side):
            fn()

    def ready(name):
        """Release ``name`` to the gradient sink.  When this completes an all-reduce bucket, the collective is issued from
        the current side stream, so that stream first waits for the gradients its siblings are still producing."""
        if USE_SIDE_STREAM and len(sides) > 1 and getattr(sink, "completes_bucket", lambda n: False)(name):
            cur = torch.cuda.current_stream(dev)
            for other in used:
                if other is not cur:
                    ev = torch.cuda.Event()
                    ev.record(other)
                    cur.wait_event(ev)
        sink.mark_ready(name)

    pending, pending_up2, pending_names = [], [], []

    def flush_small():
        """Launch the collected small-map weight gradients as batched grids on the side stream."""
        if not pending and not pending_up2:
            return
        jobs, jobs_up2, names = list(pending), list(pending_up2), list(pending_names)
        del pending[:], pending_up2[:], pending_names[:]

        def run():
            CN.conv3x3_wgrad_up2_multi(jobs_up2)
            CN.conv3x3_wgrad_multi(jobs)
            for n_ in names:
                ready(n_)
        on_side(run, *[t for j in jobs + jobs_up2 for t in (j["g"], j["x"])])

    def bias_grad(name, g, C):
        def run():
            if not _SKIP_BIAS_SUMS:
                CN.channel_sum(g, sink.grad_view(name), C)
            ready(name)
        on_side(run, g, bias=True)

    def wgrad(name, g_pre, sources, cout, stride=1, const=None, bias=None):
        """sources: the conv's (concatenated) inputs, in channel order; const = (value[B], in_hw) of a trailing constant
        input plane; bias = name of the layer's bias: its gradient (sum of g_pre over pixels) is taken along by the first
        source's launch (the kernel's spare operand slot reads ones, csrc/conv_wgrad.cu)."""
        if bias is not None and not FUSE_BIAS_GRAD:
            bias_grad(bias, g_pre, cout)
            bias = None
        dst0 = sink.grad_view(name)
        chunks = g_pre.shape[0] * ((g_pre.shape[3] + 15) // 16) * ((g_pre.shape[2] + 3) // 4)
        if (BATCH_SMALL_WGRAD and chunks <= SMALL_WGRAD_CHUNKS and const is None and dst0.shape[2:] == (3, 3)
                and not any(callable(x) for x in sources)):
            # a small-map layer: its launch is a ~15 us latency chain that holds SMs beside the data-gradient chain.  Collected
            # and launched with its neighbours as ONE grid (flush_small: when the first large layer follows, or at the end)
            off = 0
            for x in sources:
                cx = min(x.shape[1], dst0.shape[1] - off)
                pending.append(dict(g=g_pre, x=x, dW=dst0, cout=cout, cx=cx, ci_off=off, stride=stride,
                                    dbias=sink.grad_view(bias) if (bias is not None and off == 0) else None))
                off += cx
            pending_names.extend([n_ for n_ in (bias, name) if n_ is not None])
            return
        flush_small()

        def run():
            dW = dst = sink.grad_view(name)
            if dW.shape[2:] != (3, 3):          # separable kernel: full 3x3 gradient into scratch, keep the taps that exist
                dW = torch.zeros(dW.shape[0], 3, 3, dW.shape[1], device=dev, dtype=torch.float32).permute(0, 3, 1, 2)
            off = 0
            for x in sources:
                x = x() if callable(x) else x                  # lazily built input (the up-sampled map of a folded deconv)
                cx = min(x.shape[1], dW.shape[1] - off)
                CN.conv3x3_wgrad(g_pre, x, dW, cout=cout, cx=cx, ci_off=off, stride=stride,
                                 dbias=sink.grad_view(bias) if (bias is not None and off == 0) else None)
                off += cx
            if dW is not dst:
                dst.add_(dW[:, :, :, 1:2] if dst.shape[2:] == (3, 1) else dW[:, :, 1:2, :])
            if const is not None:
                CN.const_channel_wgrad_into(g_pre, const[0], const[1], stride, cout, dW, off)
            if bias is not None:
                ready(bias)
            ready(name)
        on_side(run, g_pre, *[x for x in sources if not callable(x)])

    # ---------------------------------------------------------------- folded logits conv (iconv1 o conv0)
    Np = (N + 31) // 32 * 32
    g = layout.planar_to_nhwc_bf16(g_logits, Np).permute(0, 3, 1, 2)            # bf16 [B,Np,H,W] channels_last view
    h2, xu, u, _ = tape["dec1"]
    s0 = tape["conv0"][2]
    if not FUSE_BIAS_GRAD:
        bias_grad("conv0.bias", g, N)

    def folded():
        gwf = torch.zeros(N, 3, 3, u.shape[1] + s0.shape[1], device=dev, dtype=torch.float32).permute(0, 3, 1, 2)
        CN.conv3x3_wgrad(g, u, gwf, cout=N, ci_off=0, dbias=sink.grad_view("conv0.bias") if FUSE_BIAS_GRAD else None)
        if FUSE_BIAS_GRAD:
            ready("conv0.bias")
        CN.conv3x3_wgrad(g, s0, gwf, cout=N, ci_off=u.shape[1])
        # adjoint of the fold, one launch: dW0 = <gW', W_iconv1>, dW_iconv1 = W0^T gW'
        CN.fold_logit_conv_bwd(gwf, bb.iconv1.weight, model.conv0.weight, sink.grad_view(pfx + "iconv1.weight"),
                               sink.grad_view("conv0.weight"))
        ready("conv0.weight")
        ready(pfx + "iconv1.weight")
    on_side(folded, g, u, s0)
    wd = tape["wfd"]                                                             # [96,3,3,Np], written by the fold kernel
    C1 = u.shape[1]
    g_u = CN.conv3x3_dgrad(g, wd, (H, W), rows=(0, C1), dact=1, ysave=u)
    G_skip = {0: CN.conv3x3_dgrad(g, wd, (H, W), rows=(C1, s0.shape[1]))}        # second consumer arrives in the encoder pass
    del g

    # ---------------------------------------------------------------- decoder, level 1 .. 6
    g_h = None
    for lvl in (1, 2, 3, 4, 5, 6):
        h_in, xu, u, hn = tape[f"dec{lvl}"]
        skip = tape[ENC[lvl - 1][0]][2]
        up = getattr(bb, f"deconv{lvl}")
        if lvl > 1:
            ic = getattr(bb, f"iconv{lvl}")[0]
            cout = ic.weight.shape[0]
            wgrad(pfx + f"iconv{lvl}.0.weight", g_h, (u, skip), cout, bias=pfx + f"iconv{lvl}.0.bias")
            wd = _wd(ic.weight)
            C1 = u.shape[1]
            hw = (u.shape[2], u.shape[3])
            g_u = CN.conv3x3_dgrad(g_h, wd, hw, rows=(0, C1), dact=1, ysave=u)
            G_skip[lvl - 1] = CN.conv3x3_dgrad(g_h, wd, hw, rows=(C1, skip.shape[1]))
        if xu is None:
            # folded level: the weight gradient still wants the up-sampled input -- rebuilt on the SIDE stream, off the
            # critical path -- while the data gradient goes straight to the low-resolution tensor (up-sample backward,
            # 3x3 data gradient and ELU' of the producer in one kernel)
            hw_up = (u.shape[2], u.shape[3])
            keep.append(h_in)
            if USE_UP2_WGRAD and h_in.shape[1] % 64 == 0 and g_u.shape[1] % 64 == 0 and up.conv1.weight.shape[2:] == (3, 3):
                # ... and so does the weight gradient: sixteen quarter-resolution correlations of (h, g) folded into the nine
                # taps (csrc/conv_wgrad.cu, conv3x3_wgrad_up2_kernel) -- no up-sampled tensor on either stream
                lo_chunks = h_in.shape[0] * ((h_in.shape[3] + 15) // 16) * ((h_in.shape[2] + 3) // 4)
                if BATCH_SMALL_WGRAD and 4 * lo_chunks <= UP2_WGRAD_BATCH_CHUNKS:
                    # a small level: collected and launched with its neighbours as one grid (flush_small)
                    pending_up2.append(dict(g=g_u, x=h_in, dW=sink.grad_view(pfx + f"deconv{lvl}.conv1.weight"),
                                            cout=up.conv1.weight.shape[0]))
                    pending_names.append(pfx + f"deconv{lvl}.conv1.weight")
                else:
                    def run(name=pfx + f"deconv{lvl}.conv1.weight", g_=g_u, h_=h_in, co=up.conv1.weight.shape[0]):
                        CN.conv3x3_wgrad_up2(g_, h_, sink.grad_view(name), cout=co)
                        ready(name)
                    on_side(run, g_u, h_in)
            else:
                wgrad(pfx + f"deconv{lvl}.conv1.weight", g_u, (lambda h_=h_in, hw_=hw_up: CN.upsample_nearest(h_, hw_),),
                      up.conv1.weight.shape[0])
            g_h = CN.conv3x3_up2_dgrad(g_u, _up2_packs(up.conv1.weight)[1], dact=1, ysave=h_in)
            del g_u
        else:
            wgrad(pfx + f"deconv{lvl}.conv1.weight", g_u, (xu,), up.conv1.weight.shape[0])
            g_xu = CN.conv3x3_dgrad(g_u, _wd(up.conv1.weight), (xu.shape[2], xu.shape[3]))
            # nearest-upsample backward fused with ELU' of the producer (h_{l+1}, or the bottleneck skip s6 for level 6)
            g_h = CN.upsample_nearest_bwd(g_xu, (h_in.shape[2], h_in.shape[3]), ysave=h_in, dact=1)
            del g_u, g_xu

    # ---------------------------------------------------------------- encoder, level 6 .. 0
    g_s = g_h                                                                     # pre-activation gradient of s6
    for i in (6, 5, 4, 3, 2, 1, 0):
        name, _, cout, stride = ENC[i]
        a, r, s = tape[name]
        head = getattr(bb, name)[0]
        blk = getattr(bb, name + "_1")
        hw = (a.shape[2], a.shape[3])
        wgrad(pfx + f"{name}_1.conv2.weight", g_s, (r,), cout)
        g_r = CN.conv3x3_dgrad(g_s, _wd(blk.conv2.weight), hw, dact=1, ysave=r)
        wgrad(pfx + f"{name}_1.conv1.weight", g_r, (a,), cout)
        g_a = CN.conv3x3_dgrad(g_r, _wd(blk.conv1.weight), hw, dact=1, ysave=a, residual=g_s)   # (dgrad + skip path) * ELU'
        wname, bname = pfx + f"{name}.0.weight", pfx + f"{name}.0.bias"
        if i == 0:
            if STEM_WGRAD_MMA and cout == 32 and g_a.shape[1] == 32 and head.weight.shape[2:] == (3, 3):
                # weight + bias gradient of the stem from the fp32 image itself (patch rows rebuilt in shared memory)
                def run(g_=g_a, img=tape["image"]):
                    CN.stem_wgrad(img, g_, sink.grad_view(wname), sink.grad_view(bname))
                    ready(bname)
                    ready(wname)
                on_side(run, g_a, tape["image"])
                break
            # the 3-channel image as a 32-channel (zero-padded) bf16 NHWC tensor: same tensor-core path, cx = 3
            img16 = layout.planar_to_nhwc_bf16(tape["image"].float().contiguous(), 32).permute(0, 3, 1, 2)
            wgrad(wname, g_a, (img16,), cout, bias=bname)
            break
        prev = tape[ENC[i - 1][0]][2]
        Cp = prev.shape[1]
        # conv1.0's 33rd input is the constant max_disp/100 plane
        wgrad(wname, g_a, (prev,), cout, stride,
              const=(tape["flow_val"], (prev.shape[2], prev.shape[3])) if i == 1 else None, bias=bname)
        # stride-2 dgrad into the previous skip: add to what the decoder left there, then ELU'(s_{i-1})
        g_s = CN.conv3x3_dgrad(g_a, _wd(head.weight, Cp), (prev.shape[2], prev.shape[3]), stride=stride,
                               out=G_skip.pop(i - 1), accum=True, dact=1, ysave=prev)
        del g_r, g_a
    flush_small()
    for side in used:                                                             # join: gradients complete, `keep` may go
        main.wait_stream(side)
    del keep
    return sink


class BackboneFn(torch.autograd.Function):
    """logits = backbone(image); parameters are passed positionally so autograd routes their gradients."""

    @staticmethod
    def forward(ctx, model, names, image, max_disp, *params):
        tape = {}
        out = forward(model, image, max_disp, tape)
        ctx.model, ctx.names, ctx.tape = model, names, tape
        ctx.sink = getattr(model, "_faln_grad_sink", None)
        return out

    @staticmethod
    def backward(ctx, g_logits):
        tape, ctx.tape = ctx.tape, None
        Bq, Nq, Hq, Wq = g_logits.shape
        sb, sn, sh, sw = g_logits.stride()
        if sw != 1 or sn != Hq * sh or sb != Nq * Hq * sh or sh < Wq:            # need a uniformly pitched planar tensor
            g = layout.alloc_planar(Bq, Nq, Hq, Wq, g_logits.device)
            g.copy_(g_logits)
            g_logits = g
        sink = ctx.sink if (ctx.sink is not None and ctx.sink.accepts(ctx.model)) else None
        out = backward(ctx.model, tape, g_logits, sink)
        if sink is not None:                       # gradients already sit in the flat arena the parameters' .grad alias
            return (None, None, None, None) + (None,) * len(ctx.names)
        return (None, None, None, None) + tuple(out.grads.get(n) for n in ctx.names)


def logits(model, image, max_disp):
    """Training-aware entry: tape + hand-scheduled backward when gradients are required, plain forward otherwise."""
    if not image.is_cuda:
        raise RuntimeError("fal_net_b200.FAL_netB runs on CUDA (sm_100a) only; there is no CPU path")
    named = [(n, p) for n, p in model.named_parameters() if "amask_conv" not in n]
    if torch.is_grad_enabled() and any(p.requires_grad for _, p in named):
        names = tuple(n for n, _ in named)
        return BackboneFn.apply(model, names, image, max_disp, *[p for _, p in named])
    return forward(model, image, max_disp, None)
