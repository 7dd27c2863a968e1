"""Oracle: inverse-CDF importance resampling and the PropNet sampling chain (a12).
Test infrastructure only.

PARITY UNPINNED.  The reference delegates this arithmetic to the third-party package nerfacc
(environment.yml:26, unpinned; the PropNetEstimator API implies >= 0.5.0), which is neither part
of /root/reference nor installed here, and the reference holds no test or fixture at that
boundary.  What follows restates nerfacc 0.5.x's published algorithm as summarised in SURVEY.md
§8(a12)/(c), anchored on the reference's call sites garf/model_garf.py:210-230,257:
  * targets u_k = u_floor + (k + bias) * (u_ceil - u_floor) / S with one `bias` per ray
    (uniform jitter if stratified, else 0.5);
  * right-bisect u_k into the cdf, linear interpolation inside the bin (bin midpoint when the
    bin's cdf mass is < 1e-10);
  * new interval edges = midpoints between consecutive samples; the first / last edge is
    extrapolated by half a step and clipped to the ray's first / last edge;
  * "lindisp" maps s in [0,1] to t = 1 / (s/far + (1-s)/near).
nerfacc draws its jitter from an internal Philox stream, which cannot be reproduced from torch;
here (and in the CUDA op) the uniforms are an explicit input.
"""
import numpy as np

f32 = np.float32


def importance_sampling(edges, cdf, n_out: int, u_ray=None):
    """edges, cdf: (B, Sc+1).  Returns new edges (B, n_out+1) and bin indices (B, n_out) int32."""
    edges = np.asarray(edges, dtype=f32)
    cdf = np.asarray(cdf, dtype=f32)
    B, E = edges.shape
    Sc = E - 1
    out = np.zeros((B, n_out + 1), dtype=f32)
    idx = np.zeros((B, n_out), dtype=np.int32)
    for r in range(B):
        c, e = cdf[r], edges[r]
        u_floor, u_ceil = c[0], c[Sc]
        u_step = f32(f32(u_ceil - u_floor) / f32(n_out))
        bias = f32(0.5) if u_ray is None else f32(u_ray[r])
        s = np.zeros(n_out, dtype=f32)
        for k in range(n_out):
            u = f32(u_floor + f32(f32(f32(k) + bias) * u_step))
            p = int(np.searchsorted(c, u, side="right")) - 1
            p = min(max(p, 0), Sc - 1)
            dc = f32(c[p + 1] - c[p])
            if dc < f32(1e-10):
                s[k] = f32(f32(e[p] + e[p + 1]) * f32(0.5))
            else:
                scale = f32(f32(e[p + 1] - e[p]) / dc)
                s[k] = f32(f32(f32(u - c[p]) * scale) + e[p])
            idx[r, k] = p
        if n_out == 1:
            out[r, 0], out[r, 1] = e[0], e[Sc]
        else:
            out[r, 0] = max(f32(s[0] - f32(f32(s[1] - s[0]) * f32(0.5))), e[0])
            out[r, n_out] = min(f32(s[-1] + f32(f32(s[-1] - s[-2]) * f32(0.5))), e[Sc])
            out[r, 1:n_out] = (s[:-1] + s[1:]).astype(f32) * f32(0.5)
    return out, idx


def lindisp_s_to_t(s, near: float, far: float):
    import torch as th
    s = th.as_tensor(s)
    return 1.0 / (s / far + (1.0 - s) / near)
