import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fal_net_b200 import loss_functions as LF, conv_native as CN
from fal_net_b200.conv import VGG_CFG
dev = torch.device("cuda:0")
vgg = LF.Vgg19_pc().to(dev)
g = torch.Generator().manual_seed(3)
x = (torch.rand(2, 3, 48, 80, generator=g) - 0.43).to(dev).requires_grad_(True)
outs = vgg(x)
cots = [torch.randn(o.shape, generator=g).to(dev) for o in outs]
loss = sum((o.float() * c).sum() for o, c in zip(outs, cots))
(gx,) = torch.autograd.grad(loss, x)
xr = x.detach().clone().requires_grad_(True)
h, i, refs = xr, 0, []
for v in VGG_CFG:
    if v == "M":
        h = F.max_pool2d(h, 2, 2); refs.append(h)
    else:
        w16 = vgg.weights[i].to(torch.bfloat16).float() if i else vgg.weights[i].float()
        h = F.relu(F.conv2d(h, w16, vgg.biases[i].float(), 1, 1))
        h = h + (h.to(torch.bfloat16).float() - h).detach()
        i += 1
lref = sum((o * c).sum() for o, c in zip(refs, cots))
(gr,) = torch.autograd.grad(lref, xr)
for o, r in zip(outs, refs):
    print("fwd", tuple(o.shape), float((o.float() - r).abs().max() / r.abs().max()), float((o.float() - r).norm() / r.norm()))
print("grad rel-L2", float((gx - gr).norm() / gr.norm()))
# single-layer dgrad checks with ReLU' and residual at VGG shapes
CL = torch.channels_last
gen = torch.Generator().manual_seed(5)
for (B, H, W, Cin, Cout, dact, res) in [(2, 48, 80, 64, 64, 2, False), (2, 24, 40, 128, 128, 2, False), (2, 24, 40, 64, 128, 0, True),
                                        (2, 12, 20, 256, 256, 2, False), (2, 12, 20, 128, 256, 0, True), (2, 6, 10, 512, 512, 2, False),
                                        (2, 6, 10, 256, 512, 0, True)]:
    w = (torch.randn(Cout, Cin, 3, 3, generator=gen) * (2.0 / (9 * Cout)) ** 0.5).bfloat16().to(dev)
    gg = torch.randn(B, Cout, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
    ys = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL) if dact else None
    r = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL) if res else None
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), gg.float(), 1, 1)
    if r is not None: ref = ref + r.float()
    if dact: ref = ref * (ys.float() > 0)
    got = CN.conv3x3_dgrad(gg, CN.pack_weight_dgrad(w), (H, W), 1, dact=dact, ysave=ys, residual=r)
    print("dgrad", (B, H, W, Cin, Cout, dact, res), float((got.float() - ref).norm() / ref.norm()))
    xx = torch.randn(B, Cin, H, W, generator=gen).bfloat16().to(dev).contiguous(memory_format=CL)
    y = CN.conv3x3_fwd(xx, CN.pack_weight(w), None, 1, 2)
    yr = F.relu(F.conv2d(xx.float(), w.float(), None, 1, 1))
    print("fwd  ", (B, H, W, Cin, Cout), float((y.float() - yr).norm() / yr.norm()))
