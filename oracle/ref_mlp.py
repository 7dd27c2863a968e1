"""Oracle: radiance / proposal networks (a5-a9) as pure functions of a state dict.
Test infrastructure only.  fp32 on the CPU, the arithmetic of the reference modules."""
import torch as th
import torch.nn.functional as F


def softplus8(x):
    """torch.nn.Softplus(threshold=8), reference barf/model_interpolation_architecture.py:89."""
    return F.softplus(x, beta=1.0, threshold=8.0)


class _RoundBf16(th.autograd.Function):
    """bf16 rounding of a stored operand: values forward, gradients backward."""

    @staticmethod
    def forward(ctx, x):
        return x.to(th.bfloat16).to(th.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(th.bfloat16).to(th.float32)


class _RoundGradBf16(th.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(th.bfloat16).to(th.float32)


def nerf_model_forward(sd: dict, cfg: dict, pe_pos: th.Tensor, pe_dir: th.Tensor, emulate_bf16: bool = False):
    """reference barf/model_interpolation_architecture.py:104-141 (NerfModel.forward) on
    already-encoded inputs.  sd: state dict with keys model_segments.{i}.{2k}.weight/bias and
    model_color.{0,2}.weight/bias; cfg: n_hidden, n_segments, delayed_direction, delayed_density.

    emulate_bf16=True keeps the reference's arithmetic but rounds to bf16 exactly the tensors the
    CUDA path stores in bf16 (encodings, weights, every layer's stored activations, and the
    gradients w.r.t. them), with fp32 accumulation — the tight comparison target for the kernels;
    the fp32 result (emulate_bf16=False) is the reference proper."""
    rb = _RoundBf16.apply if emulate_bf16 else (lambda t: t)
    rg = _RoundGradBf16.apply if emulate_bf16 else (lambda t: t)
    n_hidden, n_segments = cfg["n_hidden"], cfg["n_segments"]
    pe_pos, pe_dir = rb(pe_pos), rb(pe_dir)
    z = th.zeros((pe_pos.shape[0], 0))
    density = None
    for i in range(n_segments):
        if not cfg["delayed_direction"]:
            z = th.cat((z, pe_dir), dim=1)
        z = th.cat((z, pe_pos), dim=1)
        if n_hidden == 0:
            z = F.linear(z, rb(sd[f"model_segments.{i}.weight"]), sd[f"model_segments.{i}.bias"])
        else:
            for k in range(n_hidden + 1):
                if k > 0:
                    z = rb(th.relu(z))
                z = F.linear(z, rb(sd[f"model_segments.{i}.{2 * k}.weight"]), sd[f"model_segments.{i}.{2 * k}.bias"])
        if i < n_segments - 1:
            z = rb(th.relu(z))
    length = z.shape[1] - (0 if cfg["delayed_density"] else 1)
    if not cfg["delayed_density"]:
        density = rg(z[:, -1])
    zz = rb(z[:, :length])
    final_in = th.cat((zz, pe_dir), dim=1) if cfg["delayed_direction"] else zz
    h = rb(th.relu(F.linear(final_in, rb(sd["model_color.0.weight"]), sd["model_color.0.bias"])))
    out = rg(F.linear(h, rb(sd["model_color.2.weight"]), sd["model_color.2.bias"]))
    if cfg["delayed_density"]:
        density = out[:, -1]
    return softplus8(density), th.sigmoid(out[:, :3])


def gauss_act(x, inv_std):
    """reference barf/gaussian.py:8-63: exp(-x^2 * (inv_std^2 + 1e-6))."""
    return th.exp(-x ** 2 * (inv_std ** 2 + 1e-6))


def sarf_act(x, f):
    """reference sarf/activation.py:40-66 (autograd path actually used)."""
    xs = (th.signbit(x) * 2 - 1) * (x.abs() + 1e-4)
    return th.cos(f / (xs ** 2 + 1 / f ** 2)) * th.exp(-xs ** 2)


def gabor_act(x, v, s):
    """reference gaborf/gabor.py:8-29: exp(-v x^2) * cos(s x)."""
    return th.exp(-v * x ** 2) * th.cos(s * x)
