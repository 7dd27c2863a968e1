"""CPU study: how far are bf16-operand gradients from fp32 gradients for the standard NerfModel?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.nn.functional as F
from oracle import ref_mlp, ref_pe


class RoundBf16(th.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(th.bfloat16).to(th.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(th.bfloat16).to(th.float32)   # dY is rounded to bf16 as well


def rb(x):
    return RoundBf16.apply(x)


def emu_forward(sd, cfg, P, D):
    n_hidden, n_segments = cfg["n_hidden"], cfg["n_segments"]
    P, D = rb(P), rb(D)
    z = th.zeros((P.shape[0], 0))
    for i in range(n_segments):
        z = th.cat((z, P), dim=1)
        for k in range(n_hidden + 1):
            if k > 0:
                z = rb(th.relu(z))
            z = F.linear(z, rb(sd[f"model_segments.{i}.{2*k}.weight"]), sd[f"model_segments.{i}.{2*k}.bias"])
        if i < n_segments - 1:
            z = rb(th.relu(z))
    dens = z[:, -1]
    zz = rb(z[:, :-1])
    h = rb(th.relu(F.linear(th.cat((zz, D), 1), rb(sd["model_color.0.weight"]), sd["model_color.0.bias"])))
    out = F.linear(h, rb(sd["model_color.2.weight"]), sd["model_color.2.bias"])
    return ref_mlp.softplus8(dens), th.sigmoid(out[:, :3])


import importlib
sys.path.insert(0, "/root/repo/oracle/_stubs")
th.manual_seed(2)
import torch.nn as nn
# same construction order as NerfModel (first, last, mids per segment)
def seg(d_in, h, d_out, n_hidden):
    first = nn.Linear(d_in, h); last = nn.Linear(h, d_out); mids = []
    for _ in range(n_hidden - 1): mids += [nn.ReLU(), nn.Linear(h, h)]
    return nn.Sequential(first, *mids, nn.ReLU(), last)
segs = nn.ModuleList([seg(63, 256, 256, 4), seg(319, 256, 257, 4)])
color = nn.Sequential(nn.Linear(283, 128), nn.ReLU(), nn.Linear(128, 3))
sd = {}
for i, s in enumerate(segs):
    for k, m in enumerate(s):
        if isinstance(m, nn.Linear):
            sd[f"model_segments.{i}.{k}.weight"] = m.weight; sd[f"model_segments.{i}.{k}.bias"] = m.bias
sd["model_color.0.weight"] = color[0].weight; sd["model_color.0.bias"] = color[0].bias
sd["model_color.2.weight"] = color[2].weight; sd["model_color.2.bias"] = color[2].bias
cfg = dict(n_hidden=4, n_segments=2, delayed_direction=True, delayed_density=False)
n = 657
g = th.Generator().manual_seed(3)
pos = (th.rand((n, 3), generator=g) * 2 - 1) * 1.5
d = F.normalize(th.randn((n, 3), generator=g), dim=1)
gs = th.randn((n,), generator=g) * 0.1
gc = th.randn((n, 3), generator=g)
P = ref_pe.barf_encoding(pos, 10, 1.0, True, th.tensor(6.5)); D = ref_pe.barf_encoding(d, 4, 1.0, True, th.tensor(4.0))
s1, c1 = ref_mlp.nerf_model_forward(sd, cfg, P, D)
((s1 * gs).sum() + (c1 * gc).sum()).backward()
ref = {k: v.grad.clone() for k, v in sd.items()}
for v in sd.values(): v.grad = None
s2, c2 = emu_forward(sd, cfg, P, D)
((s2 * gs).sum() + (c2 * gc).sum()).backward()
for k, v in sd.items():
    print(f"{k:32s} {((v.grad - ref[k]).norm() / ref[k].norm()).item():.4f}")
