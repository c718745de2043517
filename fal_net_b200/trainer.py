"""Optimiser + data-parallel plumbing for the Stage-1 / Stage-2 steps.

* All trainable parameters that actually receive gradient (the reference's ``amask_conv`` never does,
  SURVEY.md 7) are re-homed into ONE flat fp32 arena; their ``.grad`` are views into a second arena.
  The Adam update of /root/reference/Train_Stage1_K.py:177-181 (two param groups, betas (0.5, 0.999),
  wd 0) is then a single fused kernel over the arena (csrc/misc.cu) instead of ~150 foreach launches.
* Data parallelism is one process per GPU (torch.distributed / NCCL over NVLink): the gradient arena is
  cut into contiguous buckets laid out in the order gradients become ready (decoder first); each bucket
  is all-reduced asynchronously as soon as its last gradient has been accumulated, overlapping the rest
  of backward.  ``step()`` waits for the buckets and applies Adam with grad_scale = 1/world_size.
  (The reference has no distributed code: it wraps the model in single-device DataParallel.)
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import optim


class FlatAdamDDP:
    def __init__(self, model, lr, betas=(0.5, 0.999), eps=1e-8, weight_decay=0.0, bias_decay=0.0,
                 bucket_mb: float = 16.0, process_group=None, overlap=True, _update=None, tail_mb: float = 1.0,
                 bucket_adam=None):
        named = model.used_parameters() if hasattr(model, "used_parameters") else list(model.named_parameters())
        named = [(n, p) for n, p in named if p.requires_grad]
        # arena order = reverse registration order ~ the order in which backward produces gradients
        named = list(reversed(named))
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        offs, o = [], 0
        for s in sizes:
            offs.append(o)
            o += (s + 3) // 4 * 4                       # keep every tensor 16-byte aligned inside the arena
        self.n = (o + 3) // 4 * 4
        self.offsets = offs
        self.shapes = [tuple(p.shape) for p in self.params]
        self.p = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.g = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.n, device=dev, dtype=torch.float32)
        for i, p in enumerate(self.params):
            view = self._view(self.p, i)
            view.copy_(p.data)
            p.data = view
            p.grad = self._view(self.g, i)
            p._faln_arena = (self, i)
        # ---- bf16 shadow of the arena (kept current by the Adam kernel) and the per-step dgrad re-packing: 3x3 conv weights
        # live in the arena in KRSC order, so the shadow IS the packed forward weight and the weight gradient is written
        # coalesced; the [Cin,3,3,Cout] packs of the data-gradient kernels come from one batched transpose launch
        self.w16 = self.wd16 = self._jobs = None
        self._dgrad_off = {}
        self._versions = [p._version for p in self.params]
        if dev.type == "cuda":
            self.w16 = torch.zeros(self.n, device=dev, dtype=torch.bfloat16)
            jobs, o, max_tiles, total_tiles = [], 0, 1, 0
            for i, shp in enumerate(self.shapes):
                if len(shp) == 4 and shp[2:] == (3, 3) and shp[0] % 32 == 0 and shp[1] >= 32:
                    cout, cin = shp[0], shp[1]
                    used = cin // 32 * 32
                    jobs.append([offs[i], o, cout, cin, used, total_tiles])     # [5]: first tile of the job in the flat list
                    self._dgrad_off[i] = (o, used)
                    o += used * 9 * cout
                    max_tiles = max(max_tiles, 9 * (cout // 32) * (used // 32))
                    total_tiles += 9 * (cout // 32) * (used // 32)
            self.wd16 = torch.zeros(max(o, 8), device=dev, dtype=torch.bfloat16)
            self._jobs = torch.tensor(jobs, dtype=torch.int64, device=dev) if jobs else None
            self._max_tiles, self._total_tiles = max_tiles, total_tiles
            self.sync_shadow()
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        # The reference's two Adam param groups (Train_Stage1_K.py:177-181: bias_parameters with bias_decay, weight_parameters
        # with weight_decay).  Equal decays (the default, both 0) ride in the fused kernel's scalar; distinct decays use a
        # per-element decay arena folded into the gradient before the kernel (g += wd_elem * p, Adam's L2 form).
        self.wdv = None
        if weight_decay != bias_decay:
            self.wd = 0.0
            self.wdv = torch.zeros(self.n, device=dev, dtype=torch.float32)
            for i, nme in enumerate(self.names):
                n_i = 1
                for d_ in self.shapes[i]:
                    n_i *= d_
                self.wdv[offs[i]:offs[i] + n_i] = bias_decay if "bias" in nme else weight_decay
        self.t = 0
        # FALN_DEFER_REPACK=1: start the data-gradient re-pack on the side stream at the beginning of the next forward instead of
        # right after Adam.  Measured on B200 (100-step runs): Stage-1 4.45 ms deferred vs 4.42 ms immediate, Stage-2 14.47 vs
        # 14.36 ms -- the side-stream transpose competes with the full-resolution layers that open the forward.  Default: off.
        self.defer_repack = dev.type == "cuda" and os.environ.get("FALN_DEFER_REPACK", "0") not in ("", "0")
        self._repack_stale = False
        self._repack_event = None
        # device-side copy of (lr, step) for CUDA-graph replay (optim.adam_step_dev_); None on the CPU test path
        self.hp = torch.tensor([lr, 0.0, 0.0, 0.0], device=dev, dtype=torch.float32) if dev.type == "cuda" else None
        self.device_hp = False                          # GraphedStep switches this on
        self._update = _update or optim.adam_step_     # tests inject a CPU update to exercise the host logic on gloo
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.overlap = overlap and self.world > 1
        # ---- buckets: contiguous arena ranges of ~bucket_mb
        self.buckets = []                                # (start, end, [param indices])
        cap = int(bucket_mb * (1 << 20) / 4)
        start, members = 0, []
        for i, (off, s) in enumerate(zip(offs, sizes)):
            members.append(i)
            end = off + (s + 3) // 4 * 4
            if end - start >= cap or i == len(sizes) - 1:
                self.buckets.append((start, end, members))
                start, members = end, []
        # The gradients that arrive LAST (the encoder's first layers) are the smallest tensors of the model: they get a
        # bucket of their own (<= tail_mb), so that the only all-reduce nothing can hide is a latency-sized one, and the
        # rest of the former last bucket (7 MB at 16 MB buckets) is reduced under the backward of those layers.
        if os.environ.get("FALN_TAIL_MB", "") != "":
            tail_mb = float(os.environ["FALN_TAIL_MB"])            # A/B switch (0 = no tail bucket)
        if tail_mb and tail_mb > 0 and self.buckets:
            s_, e_, mem = self.buckets[-1]
            cap_t, tot, k = int(tail_mb * (1 << 20) / 4), 0, len(mem)
            while k > 1 and tot + (sizes[mem[k - 1]] + 3) // 4 * 4 <= cap_t:
                tot += (sizes[mem[k - 1]] + 3) // 4 * 4
                k -= 1
            if 0 < k < len(mem):
                cut = offs[mem[k]]
                self.buckets[-1] = (s_, cut, mem[:k])
                self.buckets.append((cut, e_, mem[k:]))
        # Adam bucket by bucket, right behind each bucket's all-reduce (or, on one GPU, as soon as the bucket's gradients
        # are complete) on an optimiser stream beside the rest of backward; step() joins it and sweeps up what is left.
        # Nothing in backward reads the fp32 / bf16 parameter arenas after a bucket's gradients exist (the data-gradient
        # kernels read the separate [Cin,3,3,Cout] packs, refreshed after the join).  FALN_BUCKET_ADAM=0 switches it off.
        # Measured on one B200 (100-step A/B/A/B): Stage-1 step 4.115 / 4.117 ms with the single Adam launch, 4.121 / 4.123 ms
        # bucket by bucket -- on one GPU nothing waits, so the default there is the single launch; with an all-reduce in the
        # step (world > 1) the bucket-wise update is the default.
        if bucket_adam is None:
            env = os.environ.get("FALN_BUCKET_ADAM", "")
            bucket_adam = (self.world > 1) if env == "" else env != "0"
        self.bucket_adam = bool(bucket_adam) and self.wdv is None
        self.track = self.overlap or self.bucket_adam  # count gradient arrivals per bucket
        self.comm_enabled = True                       # bench.py switches the all-reduce off to measure what it costs
        self._opt_stream = None
        self._adam_done = [False] * len(self.buckets)
        self._ticked = False
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._pending = [0] * len(self.buckets)
        self._works = []
        if self.track:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))
        self._reset_counts()
        # hand-scheduled backward passes (fal_net_b200.backbone) reduce their gradients straight into the arena
        self._index = {n: i for i, n in enumerate(self.names)}
        self._model_id = id(model)
        try:
            model._faln_grad_sink = self
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _view(self, arena, i):
        """Logical-shape view of parameter i inside ``arena``; 3x3 conv weights are stored KRSC (= torch.channels_last)."""
        shp, off = self.shapes[i], self.offsets[i]
        n = 1
        for d in shp:
            n *= d
        flat = arena[off:off + n]
        if len(shp) == 4 and shp[2:] == (3, 3):
            return flat.view(shp[0], 3, 3, shp[1]).permute(0, 3, 1, 2)
        return flat.view(shp)

    def sync_shadow(self):
        """Refresh the bf16 shadow and the dgrad packs from the fp32 arena (after anything but our Adam changed it)."""
        if self.w16 is None:
            return
        from . import _lib
        _lib.check(_lib.lib().faln_f32_to_bf16(_lib.ptr(self.p), _lib.ptr(self.w16), self.n, _lib.cur_stream()),
                   "faln_f32_to_bf16")
        self._repack_dgrad()
        self._versions = [p._version for p in self.params]
        from . import conv
        conv.invalidate_packed_weights()               # per-parameter packs cached by backbone._cached are stale now

    def _repack_dgrad(self):
        if self._jobs is None:
            return
        from . import _lib
        if os.environ.get("FALN_PACK_DGRAD_2D", "0") not in ("", "0"):        # the first version: a max_tiles x njobs grid
            _lib.check(_lib.lib().faln_pack_dgrad_batched(_lib.ptr(self.w16), _lib.ptr(self.wd16), _lib.ptr(self._jobs),
                                                          self._jobs.shape[0], self._max_tiles, _lib.cur_stream()),
                       "faln_pack_dgrad_batched")
            return
        _lib.check(_lib.lib().faln_pack_dgrad_flat(_lib.ptr(self.w16), _lib.ptr(self.wd16), _lib.ptr(self._jobs),
                                                   self._jobs.shape[0], self._total_tiles, _lib.cur_stream()),
                   "faln_pack_dgrad_flat")

    def _fresh(self, i):
        if self.params[i]._version != self._versions[i]:      # someone wrote the parameter through torch: re-derive
            self.sync_shadow()

    def packed_fwd(self, i, cin):
        """bf16 [Cout,3,3,Cin] forward weight of parameter i straight out of the shadow arena, or None if the layer needs
        an explicit pack (partial input-channel range, Cout not a multiple of 32)."""
        shp = self.shapes[i]
        if self.w16 is None or len(shp) != 4 or shp[2:] != (3, 3) or cin != shp[1] or shp[0] % 32 or shp[1] % 8:
            return None
        self._fresh(i)
        off = self.offsets[i]
        return self.w16[off:off + shp[0] * 9 * shp[1]].view(shp[0], 3, 3, shp[1])

    def packed_dgrad(self, i, cin):
        """bf16 [Cin_used,3,3,Cout] data-gradient weight of parameter i (refreshed once per step), or None."""
        ent = self._dgrad_off.get(i)
        if ent is None or ent[1] != cin:
            return None
        self._fresh(i)
        if self._repack_stale:                         # nobody started the deferred re-pack: do it here, in stream order
            self._repack_dgrad()
            self._repack_stale = False
        if self._repack_event is not None:             # started on the side stream at the beginning of this step's forward
            torch.cuda.current_stream().wait_event(self._repack_event)
            self._repack_event = None
        o, used = ent
        cout = self.shapes[i][0]
        return self.wd16[o:o + used * 9 * cout].view(used, 3, 3, cout)

    # gradient-sink protocol used by backbone.backward
    def accepts(self, model):
        """True if this arena is where ``model``'s gradients live right now (p.grad aliases the arena)."""
        if id(model) != self._model_id:
            return False
        p, off = self.params[0], self.offsets[0]
        return p.grad is not None and p.grad.data_ptr() == self.g.data_ptr() + 4 * off

    def grad_view(self, name):
        return self._view(self.g, self._index[name])

    def mark_ready(self, name):
        if self.track:
            self._hook(self._index[name])

    def completes_bucket(self, name):
        """True if marking ``name`` ready will complete its bucket (launching its all-reduce and / or its Adam update)."""
        return bool(self.track) and self._pending[self._bucket_of[self._index[name]]] == 1

    def _hook(self, i):
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            s, e, _ = self.buckets[b]
            work = None
            if self.overlap and self.comm_enabled:
                work = dist.all_reduce(self.g[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                self._works.append(work)
            if self.bucket_adam:
                self._adam_bucket(b, work)

    def _adam_range(self, s, e, tick):
        """Adam over arena elements [s, e): device-side hyper-parameters under a CUDA graph, host-side otherwise."""
        w16 = self.w16[s:e] if self.w16 is not None else None
        if self.device_hp:
            optim.adam_range_dev_(self.p[s:e], self.g[s:e], self.m[s:e], self.v[s:e], self.hp, w16, tick=tick,
                                  beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.wd,
                                  grad_scale=1.0 / self.world)
        else:
            self._update(self.p[s:e], self.g[s:e], self.m[s:e], self.v[s:e], w16, lr=self.lr, beta1=self.betas[0],
                         beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.t + 1,
                         grad_scale=1.0 / self.world)

    def _adam_bucket(self, b, work):
        s, e, _ = self.buckets[b]
        if self.p.is_cuda:
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)                              # the caller has joined the streams that produced the bucket's gradients
            self._opt_stream.wait_event(ev)
            with torch.cuda.stream(self._opt_stream):
                if work is not None:
                    work.wait()                         # stream-ordered: the optimiser stream waits for the reduced bucket
                self._adam_range(s, e, tick=not self._ticked)
        else:
            if work is not None:
                work.wait()
            self._adam_range(s, e, tick=not self._ticked)
        self._ticked = True
        self._adam_done[b] = True

    def _reset_counts(self):
        for b, (_, _, mem) in enumerate(self.buckets):
            self._pending[b] = len(mem)
        self._works = []
        self._adam_done = [False] * len(self.buckets)
        self._ticked = False

    def _make_hook(self, i):
        def hook(_p):
            self._hook(i)
        return hook

    def zero_grad(self):
        self.g.zero_()
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):      # autograd may have replaced .grad; re-point it
            if p.grad is None or p.grad.data_ptr() != self.g.data_ptr() + 4 * off:
                p.grad = self._view(self.g, i)
        self._reset_counts()

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            dist.broadcast(self.p, src=src, group=self.pg)
            self.sync_shadow()

    def step(self):
        if self.world > 1 and self.comm_enabled:
            if self.overlap:
                for w in self._works:
                    w.wait()
                # buckets whose hooks did not all fire (should not happen) are reduced here
                for b, (s, e, _) in enumerate(self.buckets):
                    if self._pending[b] > 0:
                        dist.all_reduce(self.g[s:e], op=dist.ReduceOp.SUM, group=self.pg)
            else:
                dist.all_reduce(self.g, op=dist.ReduceOp.SUM, group=self.pg)
        if self.bucket_adam and any(self._adam_done):
            # the buckets updated behind their all-reduce: join the optimiser stream, then sweep up what never completed
            if self._opt_stream is not None:
                torch.cuda.current_stream().wait_stream(self._opt_stream)
            for b, (s, e, _) in enumerate(self.buckets):
                if not self._adam_done[b]:
                    self._adam_range(s, e, tick=not self._ticked)
                    self._ticked = True
            self.t += 1
            self._adam_done = [False] * len(self.buckets)
            self._ticked = False
            self._finish_step()
            return
        self.t += 1
        if self.wdv is not None:                       # distinct decays: the kernel scales g by 1/world afterwards
            self.g.addcmul_(self.wdv, self.p, value=float(self.world))
        if self.device_hp:
            optim.adam_step_dev_(self.p, self.g, self.m, self.v, self.hp, self.w16, beta1=self.betas[0], beta2=self.betas[1],
                                 eps=self.eps, weight_decay=self.wd, grad_scale=1.0 / self.world)
        else:
            self._update(self.p, self.g, self.m, self.v, self.w16, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                         eps=self.eps, weight_decay=self.wd, step=self.t, grad_scale=1.0 / self.world)
        self._finish_step()

    def _finish_step(self):
        # The [Cin,3,3,Cout] data-gradient packs are only needed by the NEXT step's backward: instead of re-packing here, on
        # the critical path right after Adam (55 us per step), the re-pack is started on the side stream at the beginning of
        # the next forward (start_repack) and overlaps it; packed_dgrad() waits for it / falls back to a synchronous re-pack.
        if self.defer_repack:
            self._repack_stale = True
        else:
            self._repack_dgrad()
        from . import conv
        conv.invalidate_packed_weights()      # the arena changed behind torch's version counters

    def start_repack(self, side_stream):
        """Launch the pending data-gradient re-pack on ``side_stream`` (ordered after everything queued on the current
        stream so far, i.e. after the Adam update that made it stale).  Called at the start of the backbone's forward."""
        if not self._repack_stale or self._jobs is None:
            self._repack_stale = False
            return
        main = torch.cuda.current_stream()
        side_stream.wait_stream(main)
        with torch.cuda.stream(side_stream):
            self._repack_dgrad()
            ev = torch.cuda.Event()
            ev.record(side_stream)
        self._repack_event = ev
        self._repack_stale = False

    # torch.optim-like conveniences used by the entry points
    @property
    def param_groups(self):
        return [{"lr": self.lr}]

    def set_lr(self, lr):
        self.lr = lr
        if self.hp is not None:
            self.hp[0:1].fill_(lr)


class GraphedStep:
    """One whole training step (zero_grad, forward, losses, backward, gradient all-reduce, Adam) captured once in a
    CUDA graph and replayed: the ~1,500 kernel launches of a step cost one ``cudaGraphLaunch`` on the host.

    ``loss_fn(left, right) -> (loss, *aux)`` must be shape-static and free of host synchronisation (the step bodies of
    fal_net_b200.steps are).  ``run(left, right)`` copies the batch into the static input buffers (device-to-device or
    pinned-host-to-device, on the current stream), replays the graph and returns the static loss tensor.
    The Adam step counter / learning rate live on the device (FlatAdamDDP.hp), so replays stay exact."""

    def __init__(self, opt: "FlatAdamDDP", loss_fn, left, right, warmup: int = 3):
        self.opt, self.loss_fn = opt, loss_fn
        self.left = torch.empty_like(left, device=opt.p.device)
        self.right = torch.empty_like(right, device=opt.p.device)
        self.left.copy_(left)
        self.right.copy_(right)
        opt.device_hp = True
        opt.hp[1:2].fill_(float(opt.t))
        # Warm-up (allocator, lazy inits, NCCL) runs real steps: snapshot the optimiser state and put it back afterwards, so
        # capturing a graph does not train the model (three silent Adam updates on the first batch).
        snap = [t.clone() for t in (opt.p, opt.m, opt.v, opt.hp)] if warmup > 0 else None
        t0 = opt.t
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                    # warm-up on a side stream
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            for dst, src in zip((opt.p, opt.m, opt.v, opt.hp), snap):
                dst.copy_(src)
            opt.t = t0
            opt.sync_shadow()
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # FALN_MAIN_PRIORITY=1: capture on a high-priority stream, so the kernels of the critical (forward / data-gradient)
        # chain are scheduled ahead of the parameter-gradient work on the default-priority side stream
        prio = int(os.environ.get("FALN_MAIN_PRIORITY", "0"))
        cap_stream = torch.cuda.Stream(priority=-1) if prio else None
        t_cap = opt.t
        # every replay must refresh the data-gradient weight packs (deferred re-pack, FlatAdamDDP.step): make sure the
        # captured forward contains that launch even when the packs happen to be current right now
        opt._repack_stale, opt._repack_event = opt.defer_repack, None
        with torch.cuda.graph(self.graph, stream=cap_stream):
            self.loss = self._body()
        opt.t = t_cap                                    # capturing executed nothing: the host step counter must not move
        self.replays = 0
        # input prefetch: the NEXT batch's host->device copy runs on a copy stream while the current step executes
        self._copy_stream = torch.cuda.Stream()
        self._stage = None
        self._staged = None                              # event: staged batch has landed on the device
        self._consumed = torch.cuda.Event()              # event: staged batch has been moved into the graph's inputs
        self._consumed.record()
        # loss read-back ring (run_async / loss_value): every step's loss is copied to pinned host memory right behind the
        # step; the host reads it one step late, so the device never waits for the host between two steps
        self._loss_host = torch.empty(self.LOSS_RING, dtype=torch.float32).pin_memory()
        self._loss_ev = [None] * self.LOSS_RING

    def _body(self):
        self.opt.zero_grad()
        out = self.loss_fn(self.left, self.right)
        loss = out[0] if isinstance(out, (tuple, list)) else (out["loss"] if isinstance(out, dict) else out)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def prefetch(self, left, right):
        """Start copying the next batch (pinned host or device tensors) into a staging buffer on the copy stream; the next
        ``run()`` without arguments consumes it.  Lets the H2D transfer of step i+1 overlap the kernels of step i."""
        if self._stage is None:
            self._stage = (torch.empty_like(self.left), torch.empty_like(self.right))
        cs = self._copy_stream
        cs.wait_event(self._consumed)                    # the staging buffers are free again
        with torch.cuda.stream(cs):
            self._stage[0].copy_(left, non_blocking=True)
            self._stage[1].copy_(right, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(cs)

    def run(self, left=None, right=None):
        if left is None and right is None and self._staged is not None:
            main = torch.cuda.current_stream()
            main.wait_event(self._staged)
            self.left.copy_(self._stage[0], non_blocking=True)     # device-to-device, ~10 us
            self.right.copy_(self._stage[1], non_blocking=True)
            self._consumed = torch.cuda.Event()
            self._consumed.record(main)
            self._staged = None
        if left is not None:
            self.left.copy_(left, non_blocking=True)
        if right is not None:
            self.right.copy_(right, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.opt.t += 1
        return self.loss

    LOSS_RING = 4

    def run_async(self, left=None, right=None) -> int:
        """``run()`` plus an asynchronous device-to-host copy of this step's loss; returns a ticket for ``loss_value``.
        The reference reads ``loss.item()`` right after every step (Train_Stage1_K.py:249,259), which drains the device
        before the next step can be queued; here the host queues step i + 1 first and then reads the loss of step i."""
        self.run(left, right)
        k = self.replays % self.LOSS_RING
        self._loss_host[k:k + 1].copy_(self.loss.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._loss_ev[k] = (self.replays, ev)
        return self.replays

    def loss_value(self, ticket: int) -> float:
        """Blocks until the loss of the step ``run_async`` returned ``ticket`` for is on the host, and returns it."""
        k = ticket % self.LOSS_RING
        ent = self._loss_ev[k]
        if ent is None or ent[0] != ticket:
            raise RuntimeError(f"loss of step {ticket} is no longer in the read-back ring (depth {self.LOSS_RING})")
        ent[1].synchronize()
        return float(self._loss_host[k])
