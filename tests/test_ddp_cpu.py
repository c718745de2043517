"""CPU, world_size 2 on gloo: the data-parallel host logic (flat arenas, gradient-ready bucketing, asynchronous
bucket all-reduce from post-accumulate-grad hooks, 1/world scaling) of fal_net_b200.trainer.FlatAdamDDP.
The fused-Adam CUDA kernel is replaced by the oracle's Adam through the trainer's test hook, so that only host
logic is exercised here (the kernel itself is tested on the GPU)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(7, 5)
        self.b = torch.nn.Linear(5, 3)
        self.unused = torch.nn.Linear(3, 3)

    def used_parameters(self):
        return [(n, p) for n, p in self.named_parameters() if "unused" not in n]

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _cpu_update(p, g, m, v, w16, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale):
    from oracle import falnet_oracle as O
    O.adam_step({"p": p}, {"p": g * grad_scale}, {"p": m}, {"p": v}, step, lr, beta1, beta2, eps)


def _worker(rank, world, port, out, bucket_adam=True):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fal_net_b200.trainer import FlatAdamDDP
    torch.manual_seed(0)
    model = _Tiny()
    opt = FlatAdamDDP(model, lr=1e-2, bucket_mb=1e-4, _update=_cpu_update,      # tiny buckets -> several all-reduces
                      bucket_adam=bucket_adam)
    assert len(opt.buckets) >= 2 and opt.bucket_adam == bucket_adam
    opt.broadcast_parameters()
    # gradient-sink protocol of the hand-scheduled backward: marking the last member of a bucket launches its all-reduce
    # (the backward then makes its side stream wait for the sibling streams first, backbone.backward.ready)
    opt.zero_grad()
    b0 = opt._bucket_of[0]
    members = [n for n in opt.names if opt._bucket_of[opt._index[n]] == b0]
    for n in members[:-1]:
        assert not opt.completes_bucket(n)
        opt.mark_ready(n)
    assert opt.completes_bucket(members[-1])
    opt.mark_ready(members[-1])
    assert len(opt._works) == 1
    for w in opt._works:
        w.wait()
    assert opt._adam_done[b0] == bucket_adam                                    # ... and, bucket by bucket, its Adam update
    g = torch.Generator().manual_seed(100 + rank)                               # rank-offset data
    for _ in range(3):
        x = torch.randn(4, 7, generator=g)
        opt.zero_grad()
        model(x).pow(2).mean().backward()
        if bucket_adam:
            assert all(opt._adam_done)                                          # every bucket was updated behind its all-reduce
        opt.step()
        assert opt.t == _ + 1 and not any(opt._adam_done)
    torch.save({k: v.detach().clone() for k, v in model.state_dict().items()}, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("bucket_adam", [True, False])
def test_bucketed_allreduce_matches_single_process(tmp_path, bucket_adam):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), bucket_adam), nprocs=world, join=True)
    sd = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for k in sd[0]:
        assert torch.equal(sd[0][k], sd[1][k]), k                               # replicas stay in lock-step
    # single-process reference: average the two ranks' gradients by hand, torch.optim.Adam
    torch.manual_seed(0)
    ref = _Tiny()
    opt = torch.optim.Adam([p for _, p in ref.used_parameters()], lr=1e-2, betas=(0.5, 0.999))
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for _ in range(3):
        xs = [torch.randn(4, 7, generator=g) for g in gens]
        opt.zero_grad()
        (sum(ref(x).pow(2).mean() for x in xs) / world).backward()
        opt.step()
    for k, v in ref.state_dict().items():
        assert torch.allclose(v, sd[0][k], rtol=1e-5, atol=1e-6), k
    assert torch.equal(ref.unused.weight, sd[0]["unused.weight"])               # never touched


def test_gradient_buckets_of_the_real_model_end_in_a_small_tail():
    """FAL_netB's gradient arena at the default 16 MB buckets: contiguous, complete, in gradient-ready order, and the bucket
    that completes last (the encoder's first layers) is a latency-sized one (<= 1 MB) -- the only all-reduce that cannot
    hide behind backward."""
    sys.path.insert(0, ROOT)
    from fal_net_b200 import models
    from fal_net_b200.trainer import FlatAdamDDP
    torch.manual_seed(0)
    m = models.__dict__["FAL_netB"](None, no_levels=49)
    opt = FlatAdamDDP(m, lr=1e-4, _update=_cpu_update)
    assert opt.buckets[0][0] == 0 and opt.buckets[-1][1] == opt.n
    for (s0, e0, _), (s1, _, _) in zip(opt.buckets, opt.buckets[1:]):
        assert e0 == s1 and e0 > s0
    assert sorted(i for _, _, mem in opt.buckets for i in mem) == list(range(len(opt.params)))
    tail = opt.buckets[-1]
    assert (tail[1] - tail[0]) * 4 <= (1 << 20) and len(opt.buckets) >= 4
    assert "conv0.0.weight" in opt.names[tail[2][-1]]                          # the stem's weight gradient arrives last
    one = FlatAdamDDP(models.__dict__["FAL_netB"](None, no_levels=49), lr=1e-4, _update=_cpu_update, tail_mb=0)
    assert len(one.buckets) == len(opt.buckets) - 1
