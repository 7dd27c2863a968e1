"""Oracle: deterministic pdf resampling (a11).  Test infrastructure only.

Restates NerfInterpolation._sample_t_pdf_weighted (reference barf/model_interpolation.py:193-277)
per ray in numpy float32 with every rounding step explicit.  Two orders that PyTorch leaves
unspecified are pinned here (and shared with the CUDA kernel):
  * the row sum of the weights (:215) — `lane_strided_sum` below;
  * ties of the remainder ranking (:225) — lowest index first (what torch's CPU sort produced).
"""
import numpy as np

f32 = np.float32


def lane_strided_sum(row: np.ndarray) -> np.float32:
    """32 partial sums (lane l adds elements l, l+32, ... left to right) reduced by a
    xor-butterfly: offsets 16, 8, 4, 2, 1."""
    lanes = np.zeros(32, dtype=f32)
    for i, v in enumerate(row.astype(f32)):
        lanes[i % 32] = f32(lanes[i % 32] + v)
    for off in (16, 8, 4, 2, 1):
        lanes = (lanes + lanes[np.arange(32) ^ off]).astype(f32)
    return lanes[0]


def counts_for_ray(w: np.ndarray, n_samples: int):
    """per-bin sample counts (reference :215-227) and the success flag of :235."""
    n_bins = w.shape[0]
    n_new = f32(n_samples - n_bins)
    with np.errstate(all="ignore"):
        p = (w.astype(f32) / lane_strided_sum(w)).astype(f32)
        raw = (p * n_new).astype(f32)
        fl = np.floor(raw).astype(f32)
        err = (raw - fl).astype(f32)
        excess = f32(n_new - fl.sum(dtype=f32))
        # rank = argsort(argsort(err)), stable
        rank = np.empty(n_bins, dtype=np.int64)
        rank[np.argsort(err, kind="stable")] = np.arange(n_bins)
        add = (rank.astype(f32) >= f32(f32(n_bins) - excess)).astype(f32)
        n = (fl + add + f32(1)).astype(f32)
        ok = bool(np.all(n >= 0) and n.sum(dtype=f32) == f32(n_samples))
    return n, ok


def expand_ray(t_c: np.ndarray, delta_c: np.ndarray, n: np.ndarray, n_samples: int):
    """t_k = t_c[i] + ((k - cum_i) * delta_i) / n_i for cum_i <= k < cum_{i+1} (reference :262-269)."""
    cum = np.concatenate(([f32(0)], np.cumsum(n, dtype=f32)))
    t = np.zeros(n_samples, dtype=f32)
    for k in range(n_samples):
        i = int(np.searchsorted(cum, f32(k), side="right")) - 1
        i = min(max(i, 0), n.shape[0] - 1)
        num = f32(f32(f32(k) - cum[i]) * delta_c[i])
        t[k] = f32(t_c[i] + f32(num / n[i]))
    return t


def sample_pdf_weighted(t_coarse, weights, delta_coarse, n_samples: int, near: float, far: float,
                        fallback_u=None):
    """Full a11 on (B, Sc) numpy arrays.  Returns t_start, t_end (B, Sf), counts (B, Sc) int32,
    failed (bool).  On failure the WHOLE batch becomes equidistant samples with offset -1
    (reference :273-275), drawing fallback_u (B,) as the per-ray offset uniforms."""
    t_coarse = np.asarray(t_coarse, dtype=f32)
    weights = np.asarray(weights, dtype=f32)
    delta_coarse = np.asarray(delta_coarse, dtype=f32)
    B, Sc = t_coarse.shape
    counts = np.zeros((B, Sc), dtype=f32)
    t = np.zeros((B, n_samples), dtype=f32)
    failed = False
    for r in range(B):
        n, ok = counts_for_ray(weights[r], n_samples)
        counts[r] = n
        failed |= not ok
        if ok:
            t[r] = expand_ray(t_coarse[r], delta_coarse[r], n, n_samples)
    if failed:
        import torch as th
        from .ref_sampling import sample_uniform
        u = None if fallback_u is None else th.as_tensor(fallback_u, dtype=th.float32)
        ts, te = sample_uniform(near, far, B, n_samples, None, u, -1.0 if u is not None else 0.0)
        return ts.numpy(), te.numpy(), np.nan_to_num(counts, nan=-1).astype(np.int32), True
    t_end = np.concatenate((t[:, 1:], np.full((B, 1), f32(far), dtype=f32)), axis=1)
    return t, t_end, counts.astype(np.int32), False
